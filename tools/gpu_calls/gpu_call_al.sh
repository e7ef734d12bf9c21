#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
true
true
MICRB_SKIP_LEX=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_micrb_ring$' --launch-skip 12 -c 4 -o $out/${tag}_micrb python tools/micrb_bench.py 512 4 "${2:-8x8}" > $out/${tag}_prof.log 2>&1
tail -2 $out/${tag}_prof.log
ncu -i $out/${tag}_micrb.ncu-rep --page raw --csv > $out/${tag}_micrb_raw.csv 2>/dev/null
ncu -i $out/${tag}_micrb.ncu-rep --page source --csv --kernel-name regex:'^k_micrb_ring$' --launch-skip 1 --launch-count 1 > $out/${tag}_micrb_source.csv 2>/dev/null
ls -la $out/${tag}_micrb*

#!/bin/bash
# the whole GPU suite WITHOUT -x (every failure listed), then an ncu --set full capture of the advection kernels
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -rf > $out/${tag}_pytest_gpu_full.txt 2>&1
tail -8 $out/${tag}_pytest_gpu_full.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_semi_lagrange|k_mc_rest' --launch-skip 5 -c 5 -o $out/${tag}_advect python tools/prof_advect.py 512 2 > $out/${tag}_prof_advect.log 2>&1
tail -2 $out/${tag}_prof_advect.log
ncu -i $out/${tag}_advect.ncu-rep --page raw --csv > $out/${tag}_advect_raw.csv 2>/dev/null
ls -la $out/${tag}_advect*

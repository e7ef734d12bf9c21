#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
FLIP_BENCH_ONLY=mapPartsToMAC timeout 300 python tools/flip_bench.py 256 > $out/${tag}_flip_bench.txt 2>&1
cat $out/${tag}_flip_bench.txt
FLIP_BENCH_ONLY=mapPartsToMAC timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_mapparts.csv python tools/flip_bench.py 256 > $out/${tag}_prof.log 2>&1
FLIP_BENCH_ONLY=mapPartsToMAC timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_parts_cells" -s 4 -c 4 -o $out/${tag}_ncu_mapparts python tools/flip_bench.py 256 > $out/${tag}_ncu.log 2>&1
tail -2 $out/${tag}_ncu.log | cut -c1-100

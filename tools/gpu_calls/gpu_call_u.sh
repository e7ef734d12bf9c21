#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_liquid.csv python tools/liquid_bench.py 512 > $out/${tag}_liquid.log 2>&1
tail -6 $out/${tag}_liquid.log

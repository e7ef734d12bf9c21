#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_mg_l0_fused -s 8 -c 2 -o $out/${tag}_ncu_l0_fused python tools/prof_solve.py --res 512 --pc 3 --reps 1 > $out/${tag}_ncu.log 2>&1
tail -1 $out/${tag}_ncu.log

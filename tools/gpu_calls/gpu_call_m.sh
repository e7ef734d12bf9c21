#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multigrid.py -m gpu -q -rf -x > $out/${tag}_pytest_mg.txt 2>&1
tail -4 $out/${tag}_pytest_mg.txt
timeout 600 python tools/mg_bench.py 512 $out/${tag}_mg_bench.json > $out/${tag}_mg_bench.txt 2>&1
grep '"res": 512' $out/${tag}_mg_bench.txt | cut -c1-330
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_mg_l0_fused -c 2 -o $out/${tag}_ncu_l0_fused python tools/prof_solve.py --res 512 --pc 3 --reps 1 > $out/${tag}_ncu.log 2>&1
tail -1 $out/${tag}_ncu.log

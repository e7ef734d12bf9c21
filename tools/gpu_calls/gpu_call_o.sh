#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_parity.py -m gpu -q -rf -x > $out/${tag}_pytest_mg.txt 2>&1
tail -4 $out/${tag}_pytest_mg.txt
timeout 600 python tools/mg_bench.py 512 $out/${tag}_mg_bench.json > $out/${tag}_mg_bench.txt 2>&1
grep '"res": 512' $out/${tag}_mg_bench.txt | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file $out/${tag}_launches_pcmg_512.csv env MP_MG_L0FUSED=0 python tools/prof_solve.py --res 512 --pc 3 --reps 2 > $out/${tag}_prof.log 2>&1
tail -1 $out/${tag}_prof.log

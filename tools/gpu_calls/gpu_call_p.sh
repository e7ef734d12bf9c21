#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_zzz_flip.py tests/test_gpu_golden.py -m gpu -q -rf > $out/${tag}_pytest.txt 2>&1
tail -8 $out/${tag}_pytest.txt
timeout 600 python tools/mg_bench.py 512 $out/${tag}_mg_bench.json > $out/${tag}_mg_bench.txt 2>&1
grep '"res": 512' $out/${tag}_mg_bench.txt | cut -c1-250
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file $out/${tag}_launches_pcmg_512.csv python tools/prof_solve.py --res 512 --pc 3 --reps 2 > $out/${tag}_prof.log 2>&1
tail -1 $out/${tag}_prof.log | cut -c1-200

#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_multigrid.py -m gpu -q -rf -x > $out/${tag}_pytest.txt 2>&1
tail -3 $out/${tag}_pytest.txt
timeout 600 python tools/mg_bench.py 512 $out/${tag}_mg_bench.json > $out/${tag}_mg_bench.txt 2>&1
grep '"res": 512' $out/${tag}_mg_bench.txt | cut -c1-200 | sed -n '2p;7p'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file $out/${tag}_launches_pcmgdyn_512.csv python tools/prof_solve.py --res 512 --pc 2 --reps 2 > $out/${tag}_prof.log 2>&1
tail -1 $out/${tag}_prof.log | cut -c1-150

#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multigrid.py -m gpu -q -x > $out/${tag}_pytest_mg.txt 2>&1
tail -12 $out/${tag}_pytest_mg.txt
for cfg in "1 1" "0 0" "1 0" "0 1"; do set -- $cfg
MP_MG_FULL=$1 MP_MG_L0VEC=$2 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu --pc 3 > $out/${tag}_bench_mg_full$1_vec$2.json 2> $out/${tag}_bench_mg_full$1_vec$2.err
python -c "import json;d=json.load(open('$out/${tag}_bench_mg_full$1_vec$2.json'));print('full $1 vec $2',d['solve_ms'],d['iterations'],d['kernel_ms'])"
done
timeout 900 python -m pytest tests -m gpu -q -rf -s > $out/${tag}_pytest_gpu_full.txt 2>&1
grep -v "^\.\|^$" $out/${tag}_pytest_gpu_full.txt | tail -40
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches_pcmgstatic_512.csv python tools/prof_solve.py --res 512 --pc 3 > /dev/null 2>&1

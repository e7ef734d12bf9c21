#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_baseline_configs.py -m gpu -q -rf -x > $out/${tag}_pytest.txt 2>&1
tail -5 $out/${tag}_pytest.txt
timeout 600 python tools/mg_bench.py 512 $out/${tag}_mg_bench.json > $out/${tag}_mg_bench.txt 2>&1
grep '"res": 512' $out/${tag}_mg_bench.txt | cut -c1-200
MP_MG_REGULAR=0 timeout 600 python tools/mg_bench.py 512 $out/${tag}_mg_bench_noreg.json > $out/${tag}_mg_bench_noreg.txt 2>&1
grep '"res": 512' $out/${tag}_mg_bench_noreg.txt | cut -c1-200 | tail -2

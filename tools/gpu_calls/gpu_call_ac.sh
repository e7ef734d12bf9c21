#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_parity.py -m gpu -q -rf -x > $out/${tag}_pytest.txt 2>&1
tail -3 $out/${tag}_pytest.txt
timeout 600 python tools/mg_bench.py 512 $out/${tag}_mg_bench.json > $out/${tag}_mg_bench.txt 2>&1
grep '"res": 512' $out/${tag}_mg_bench.txt | cut -c1-200 | sed -n '2p;5p;7p'
MP_MG_COARSE_ROWS=0 timeout 600 python tools/mg_bench.py 512 > $out/${tag}_mg_bench_oldcg.txt 2>&1
grep '"res": 512' $out/${tag}_mg_bench_oldcg.txt | cut -c1-200 | sed -n '2p'

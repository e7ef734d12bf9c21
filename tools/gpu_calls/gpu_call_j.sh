#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -rf -x > $out/${tag}_pytest_gpu_x.txt 2>&1
tail -6 $out/${tag}_pytest_gpu_x.txt
timeout 300 python tools/liquid_bench.py 512 $out/${tag}_liquid_bench.json > $out/${tag}_liquid_bench.txt 2>&1
cat $out/${tag}_liquid_bench.txt
MP_LIQUID_FRONTIER=0 timeout 300 python tools/liquid_bench.py 512 $out/${tag}_liquid_bench_sweeps.json > $out/${tag}_liquid_bench_sweeps.txt 2>&1
cat $out/${tag}_liquid_bench_sweeps.txt
timeout 300 python tools/dam_bench.py 192 $out/${tag}_dam_bench.json > $out/${tag}_dam_bench.txt 2>&1
head -12 $out/${tag}_dam_bench.txt

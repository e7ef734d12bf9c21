#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_step_liquid.py -m gpu -q -rf > $out/${tag}_pytest.txt 2>&1
tail -3 $out/${tag}_pytest.txt
timeout 300 python tools/liquid_bench.py 512 $out/${tag}_liquid_bench.json > $out/${tag}_liquid_bench.txt 2>&1
cat $out/${tag}_liquid_bench.txt

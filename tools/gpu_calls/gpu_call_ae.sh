#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -q -rf > $out/${tag}_pytest_step.txt 2>&1
tail -8 $out/${tag}_pytest_step.txt
timeout 600 python tools/step_bench.py 512 128 > $out/${tag}_step_bench.txt 2>&1
cat $out/${tag}_step_bench.txt | cut -c1-200

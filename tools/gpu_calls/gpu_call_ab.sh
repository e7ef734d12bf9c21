#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -rf -x > $out/${tag}_pytest_gpu_x.txt 2>&1
tail -6 $out/${tag}_pytest_gpu_x.txt
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
tail -2 $out/${tag}_bench_n1.err; cut -c1-600 $out/${tag}_bench_n1.json

#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_micrb.py -m gpu -q -rf -x > $out/${tag}_pytest_micrb.txt 2>&1
tail -3 $out/${tag}_pytest_micrb.txt | cut -c1-250
MICRB_SKIP_LEX=1 timeout 600 python tools/micrb_bench.py 512 4 "${2:-8x4,8x8,16x8,16x16}" $out/${tag}_micrb_512.json 2>&1 | cut -c1-200

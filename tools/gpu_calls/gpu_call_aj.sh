#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_micrb.py -m gpu -q -rf -x > $out/${tag}_pytest_micrb.txt 2>&1
tail -4 $out/${tag}_pytest_micrb.txt | cut -c1-250
MICRB_SKIP_LEX=1 timeout 900 python tools/micrb_bench.py 512 4 "${2:-8x4,8x8,16x8}" $out/${tag}_micrb_512.json > $out/${tag}_micrb_512.txt 2>&1
cat $out/${tag}_micrb_512.txt | cut -c1-200

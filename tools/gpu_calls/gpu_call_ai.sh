#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_micrb.py -m gpu -q -rf -x -s > $out/${tag}_pytest_micrb.txt 2>&1
tail -25 $out/${tag}_pytest_micrb.txt | cut -c1-250
timeout 600 python tools/micrb_bench.py 256 4 "8x4,8x8,16x8,16x16" > $out/${tag}_micrb_256.txt 2>&1
cat $out/${tag}_micrb_256.txt | cut -c1-200
timeout 900 python tools/micrb_bench.py 512 4 "8x4,8x8,16x8,16x16,32x16" $out/${tag}_micrb_512.json > $out/${tag}_micrb_512.txt 2>&1
cat $out/${tag}_micrb_512.txt | cut -c1-200

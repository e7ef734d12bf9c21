#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_micrb.py tests/test_gpu_parity.py -m gpu -q -rf -x > $out/${tag}_pytest_mic.txt 2>&1
tail -3 $out/${tag}_pytest_mic.txt | cut -c1-250
timeout 600 python tools/micrb_bench.py 512 4 "0x0,8x4,8x8" $out/${tag}_micrb_512.json > $out/${tag}_micrb_512.txt 2>&1
timeout 600 python tools/micrb_bench.py 256 4 "0x0,8x8,16x8" $out/${tag}_micrb_256.json > $out/${tag}_micrb_256.txt 2>&1
timeout 600 python tools/micrb_bench.py 128 4 "0x0" $out/${tag}_micrb_128.json > $out/${tag}_micrb_128.txt 2>&1
timeout 600 python tools/micrb_bench.py 512 8 "0x0" $out/${tag}_micrb_512_f64.json > $out/${tag}_micrb_512_f64.txt 2>&1
for r in 512 256 128 512_f64; do echo "== $r"; cut -c1-200 $out/${tag}_micrb_$r.txt; done

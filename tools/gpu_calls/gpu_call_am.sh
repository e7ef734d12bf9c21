#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_micrb.py -m gpu -q -rf -x > $out/${tag}_pytest_micrb.txt 2>&1
tail -3 $out/${tag}_pytest_micrb.txt | cut -c1-250
for pre in 0 2 4; do
echo "MP_MIC_RB_PRE=$pre"
MP_MIC_RB_PRE=$pre MICRB_SKIP_LEX=1 timeout 600 python tools/micrb_bench.py 512 4 "${2:-8x4,8x8,16x8}" 2>&1 | cut -c1-200
done

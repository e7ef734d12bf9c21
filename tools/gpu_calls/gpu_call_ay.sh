#!/bin/bash
# the per-config table of the final code: every BASELINE configuration on the GPU next to the reference on the host cores
tag=${1:-rX}
mkdir -p gpurun_out
timeout 1500 python tools/run_configs.py --tag ${tag} > gpurun_out/${tag}_run_configs.log 2>&1
tail -3 gpurun_out/${tag}_run_configs.log | cut -c1-300
cp profiles/${tag}_configs.md profiles/${tag}_configs.json gpurun_out/ 2>/dev/null
cat profiles/${tag}_configs.md | cut -c1-260

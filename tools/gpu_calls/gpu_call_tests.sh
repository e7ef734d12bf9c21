#!/bin/bash
# One gpurun call: the whole GPU suite WITHOUT -x (every failure is listed, none hides the rest), then the default bench line and the
# plugin timings of the liquid / FLIP neighbours.   usage: gpurun --timeout 1500 -- 'bash tools/gpu_calls/gpu_call_tests.sh r2a'
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -rf --durations=15 > $out/${tag}_pytest_gpu_full.txt 2>&1
tail -40 $out/${tag}_pytest_gpu_full.txt
timeout 400 python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
cat $out/${tag}_bench_n1.json
timeout 300 python tools/liquid_bench.py 512 $out/${tag}_liquid_bench.json > $out/${tag}_liquid_bench.txt 2>&1
timeout 300 python tools/flip_bench.py 256 $out/${tag}_flip_bench.json > $out/${tag}_flip_bench.txt 2>&1
timeout 300 python tools/dam_bench.py 192 $out/${tag}_dam_bench.json > $out/${tag}_dam_bench.txt 2>&1
tail -30 $out/${tag}_liquid_bench.txt $out/${tag}_flip_bench.txt $out/${tag}_dam_bench.txt

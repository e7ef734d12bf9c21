#!/bin/bash
# What the first gpurun call of a round should bring back, in one go (everything lands in gpurun_out/<tag>_*):
#   1. the GPU test suite                      2. the default bench.py line (N=1)
#   3. the liquid-neighbour plugin timings     4. an ncu launch list of those plugins (cold-cache, serialised: compare SHARES)
#   5. one `--set full` capture of the extrapolation pass kernels (read here with `ncu -i ... --page raw --csv`)
#   6. the FLIP particle plugin timings (tools/flip_bench.py) + launch list + one full capture
# usage:  gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh r2'
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest_gpu.txt
python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python tools/liquid_bench.py 512 $out/${tag}_liquid_bench.json > $out/${tag}_liquid_bench.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_liquid_512.csv \
    python tools/liquid_bench.py 512 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_liquid_cells -s 4 -c 6 -o $out/${tag}_ncu_liquid_passes -f \
    python tools/liquid_bench.py 512 > /dev/null 2>&1
# 6. the FLIP particle plugins (never timed so far): timings at 256^3, a launch list, one full capture of the gather of mapPartsToMAC
timeout 600 python tools/flip_bench.py 256 $out/${tag}_flip_bench.json > $out/${tag}_flip_bench.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_flip_256.csv \
    python tools/flip_bench.py 128 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_parts_cells -c 4 -o $out/${tag}_ncu_flip_cells -f \
    python tools/flip_bench.py 128 > /dev/null 2>&1
timeout 600 python tools/dam_bench.py 192 $out/${tag}_dam_bench.json > $out/${tag}_dam_bench.txt 2>&1      # 7. the whole benchmark_dam step, plugin by plugin
ls -la $out | tail -12
cat $out/${tag}_pytest_gpu.txt $out/${tag}_liquid_bench.txt $out/${tag}_flip_bench.txt $out/${tag}_dam_bench.txt

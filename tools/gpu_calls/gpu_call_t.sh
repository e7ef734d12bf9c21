#!/bin/bash
# 2 GPUs: sharded parity check, sharded tests, bench at N = 2 (PcNone line + PcMGStatic config)
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NP:-2} --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > $out/${tag}_sharded_check_${NP:-2}gpu.txt 2>&1
grep "sharded_check\|Error\|error" $out/${tag}_sharded_check_${NP:-2}gpu.txt | tail -30
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -rf > $out/${tag}_pytest_sharded.txt 2>&1
tail -3 $out/${tag}_pytest_sharded.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NP:-2} --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus ${NP:-2} --steps 3 --warmup 3 > $out/${tag}_bench_n${NP:-2}.json 2> $out/${tag}_bench_n${NP:-2}.err
tail -2 $out/${tag}_bench_n${NP:-2}.err; cut -c1-1500 $out/${tag}_bench_n${NP:-2}.json

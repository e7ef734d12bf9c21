#!/bin/bash
# N GPUs: the default bench line (PcNone + the PcMGStatic config on the sharded grid), and on 4 GPUs the sharded parity check
tag=${1:-rX}; N=${2:-8}
out=gpurun_out
mkdir -p $out
if [ "$N" = "4" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > $out/${tag}_sharded_check_4gpu.txt 2>&1
grep "sharded_check\|Error\|error" $out/${tag}_sharded_check_4gpu.txt | tail -40
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > $out/${tag}_bench_n$N.json 2> $out/${tag}_bench_n$N.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench_n$N.json') if l.startswith('{')][-1])
print('N=$N pc 0', d['value'], d['iterations'], d['solve_ms'], d['kernel_ms'], d.get('exchange')); print(json.dumps(d.get('configs')))" || tail -20 $out/${tag}_bench_n$N.err
nvidia-smi --query-gpu=memory.used --format=csv | head -3

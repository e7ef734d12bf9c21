#!/bin/bash
# 2 GPUs: the sharded parity check (PcNone, block-Jacobi MIC, global GridMg on slabs), then weak-scaling bench lines for PcNone and PcMGStatic
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > $out/${tag}_sharded_check_2gpu.txt 2>&1
grep "sharded_check\|Error\|error" $out/${tag}_sharded_check_2gpu.txt | tail -40
for pc in 0 3; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --pc $pc > $out/${tag}_bench_n2_pc$pc.json 2> $out/${tag}_bench_n2_pc$pc.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench_n2_pc$pc.json') if l.startswith('{')][-1])
print('N=2 pc $pc', d['value'], d['iterations'], d['solve_ms'], d['kernel_ms'], d.get('exchange'))" || tail -5 $out/${tag}_bench_n2_pc$pc.err
done
MP_MG_SLAB_GLOBAL=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 2 --pc 3 > $out/${tag}_bench_n2_pc3_blockjacobi.json 2> $out/${tag}_bench_n2_pc3_blockjacobi.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench_n2_pc3_blockjacobi.json') if l.startswith('{')][-1])
print('N=2 pc 3 block-Jacobi', d['value'], d['iterations'], d['solve_ms'])"

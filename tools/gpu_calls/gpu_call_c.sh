#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tma or kernel_selection or iterate_state" -s > $out/${tag}_pytest_tma16.txt 2>&1
tail -8 $out/${tag}_pytest_tma16.txt
MP_TMA_TY=8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tma or kernel_selection" > $out/${tag}_pytest_tma8.txt 2>&1
tail -3 $out/${tag}_pytest_tma8.txt
for ty in 16 8; do
MP_TMA_TY=$ty timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_bench_ty$ty.json 2> $out/${tag}_bench_ty$ty.err
python -c "import json;d=json.load(open('$out/${tag}_bench_ty$ty.json'));print('TY',$ty,d['cg_iter_per_s'],d['kernel_ms'],d['roofline']['frac'])"
done
MP_TMA_TY=16 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu --prec 8 > $out/${tag}_bench_f64.json 2> $out/${tag}_bench_f64.err
python -c "import json;d=json.load(open('$out/${tag}_bench_f64.json'));print('f64',d['cg_iter_per_s'],d['kernel_ms'],d['roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_matvec_fused_tma -s 20 -c 1 -o $out/${tag}_ncu_matvec_tma16 -f python tools/prof_solve.py --res 512 --pc 0 --iters 40 > $out/${tag}_ncu.log 2>&1
tail -2 $out/${tag}_ncu.log

#!/bin/bash
# second call of round 2: the new TMA matvec -- parity tests, the suite, the bench line, the launch list and one full ncu capture
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tma or kernel_selection or iterate_state" > $out/${tag}_pytest_tma.txt 2>&1
tail -30 $out/${tag}_pytest_tma.txt
timeout 1200 python -m pytest tests -m gpu -q -rf > $out/${tag}_pytest_gpu_full.txt 2>&1
tail -15 $out/${tag}_pytest_gpu_full.txt
timeout 400 python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
cat $out/${tag}_bench_n1.json; tail -5 $out/${tag}_bench_n1.err
MP_CG_FUSED=1 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_bench_n1_fused1.json 2> $out/${tag}_bench_n1_fused1.err
cat $out/${tag}_bench_n1_fused1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_matvec_fused_tma -s 20 -c 2 -o $out/${tag}_ncu_matvec_tma -f python tools/prof_solve.py --res 512 --pc 0 --iters 40 > $out/${tag}_ncu.log 2>&1
tail -3 $out/${tag}_ncu.log

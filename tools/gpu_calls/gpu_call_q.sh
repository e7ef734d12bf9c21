#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_zzz_flip.py tests/test_gpu_step_liquid.py -m gpu -q -rf > $out/${tag}_pytest.txt 2>&1
tail -5 $out/${tag}_pytest.txt
timeout 300 python tools/flip_bench.py 256 $out/${tag}_flip_bench.json > $out/${tag}_flip_bench.txt 2>&1
cat $out/${tag}_flip_bench.txt
MP_MAPPARTS=0 MP_UNION_SORTED=0 timeout 300 python tools/flip_bench.py 256 $out/${tag}_flip_bench_old.json > $out/${tag}_flip_bench_old.txt 2>&1
grep -i "mapParts\|union" $out/${tag}_flip_bench_old.txt
timeout 300 python tools/dam_bench.py 192 $out/${tag}_dam_bench.json > $out/${tag}_dam_bench.txt 2>&1
head -8 $out/${tag}_dam_bench.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_make_rhs|k_scan_flags|k_correct_velocity|k_make_matrix" -c 5 -o $out/${tag}_ncu_setup python tools/prof_solve.py --res 512 --pc 3 --reps 1 > $out/${tag}_ncu.log 2>&1
tail -1 $out/${tag}_ncu.log | cut -c1-100

#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_baseline_configs.py -m gpu -q -rf -x > $out/${tag}_pytest.txt 2>&1
tail -3 $out/${tag}_pytest.txt | cut -c1-250
timeout 300 python tools/prof_solve.py --res 512 --pc 3 --reps 3 2>&1 | grep -o "'msRhs': [0-9.]*\|'msMatrix': [0-9.]*\|'msCorrect': [0-9.]*\|'msSolve': [0-9.]*\|'msTotal': [0-9.]*" | paste - - - - -

#!/bin/bash
# Round-end rehearsal: the driver's own sequence (pytest -m gpu -x, smoke(), bench.py default, bench.py --impl reference).
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -rf -x > $out/${tag}_pytest_gpu_x.txt 2>&1
tail -6 $out/${tag}_pytest_gpu_x.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.txt 2>&1
tail -3 $out/${tag}_smoke.txt
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
tail -2 $out/${tag}_bench_n1.err; cut -c1-600 $out/${tag}_bench_n1.json
timeout 600 python bench.py --impl reference > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
tail -2 $out/${tag}_bench_ref.err; cut -c1-900 $out/${tag}_bench_ref.json

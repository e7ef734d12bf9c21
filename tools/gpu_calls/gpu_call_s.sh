#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_zzz_flip.py tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -rf > $out/${tag}_pytest.txt 2>&1
tail -5 $out/${tag}_pytest.txt
timeout 300 python tools/flip_bench.py 256 $out/${tag}_flip_bench.json > $out/${tag}_flip_bench.txt 2>&1
cat $out/${tag}_flip_bench.txt
timeout 300 python tools/dam_bench.py 192 $out/${tag}_dam_bench.json > $out/${tag}_dam_bench.txt 2>&1
head -8 $out/${tag}_dam_bench.txt
FLIP_BENCH_ONLY=mapPartsToMAC timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_mapparts.csv python tools/flip_bench.py 256 > $out/${tag}_prof.log 2>&1
timeout 600 python tools/mg_bench.py 512 $out/${tag}_mg_bench.json > $out/${tag}_mg_bench.txt 2>&1
grep '"res": 512' $out/${tag}_mg_bench.txt | cut -c1-250

#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_zzz_flip.py -m gpu -q -rf > $out/${tag}_pytest.txt 2>&1
tail -4 $out/${tag}_pytest.txt
timeout 300 python tools/flip_bench.py 256 $out/${tag}_flip_bench.json > $out/${tag}_flip_bench.txt 2>&1
cat $out/${tag}_flip_bench.txt
timeout 300 python tools/dam_bench.py 192 $out/${tag}_dam_bench.json > $out/${tag}_dam_bench.txt 2>&1
head -4 $out/${tag}_dam_bench.txt
DAM_PC=2 timeout 300 python tools/dam_bench.py 192 $out/${tag}_dam_bench_pcmgdynamic.json > $out/${tag}_dam_bench_pcmgdynamic.txt 2>&1
head -8 $out/${tag}_dam_bench_pcmgdynamic.txt
FLIP_BENCH_ONLY=mapPartsToMAC timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_mapparts.csv python tools/flip_bench.py 256 > $out/${tag}_prof.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_liquid_cells_list" -c 2 -o $out/${tag}_ncu_liquid_list python tools/liquid_bench.py 512 > $out/${tag}_ncu.log 2>&1
timeout 600 python tools/mg_bench.py 512 $out/${tag}_mg_bench.json > $out/${tag}_mg_bench.txt 2>&1
grep '"res": 512' $out/${tag}_mg_bench.txt | cut -c1-100 | head -2; grep -o '"stage_ms": {[^}]*}' $out/${tag}_mg_bench.txt | tail -5 | head -2

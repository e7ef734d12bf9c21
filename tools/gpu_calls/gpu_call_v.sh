#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_step_liquid.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_step.py -m gpu -q -rf > $out/${tag}_pytest.txt 2>&1
tail -4 $out/${tag}_pytest.txt
timeout 300 python tools/liquid_bench.py 512 $out/${tag}_liquid_bench.json > $out/${tag}_liquid_bench.txt 2>&1
cat $out/${tag}_liquid_bench.txt
timeout 600 python tools/mg_bench.py 512 $out/${tag}_mg_bench.json > $out/${tag}_mg_bench.txt 2>&1
grep '"res": 512' $out/${tag}_mg_bench.txt | cut -c1-220 | head -3
timeout 300 python tools/dam_bench.py 192 $out/${tag}_dam_bench.json > $out/${tag}_dam_bench.txt 2>&1
head -4 $out/${tag}_dam_bench.txt
DAM_PC=2 timeout 300 python tools/dam_bench.py 192 $out/${tag}_dam_bench_pcmgdynamic.json > $out/${tag}_dam_bench_pcmgdynamic.txt 2>&1
head -4 $out/${tag}_dam_bench_pcmgdynamic.txt

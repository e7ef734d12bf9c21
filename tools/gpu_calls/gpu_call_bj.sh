#!/bin/bash
# liquid / FLIP / dam timings of the final code
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 300 python tools/liquid_bench.py 512 $out/${tag}_liquid_bench.json > $out/${tag}_liquid_bench.txt 2>&1
timeout 300 python tools/flip_bench.py 256 $out/${tag}_flip_bench.json > $out/${tag}_flip_bench.txt 2>&1
timeout 300 python tools/dam_bench.py 192 $out/${tag}_dam_bench.json > $out/${tag}_dam_bench.txt 2>&1
DAM_PC=2 timeout 300 python tools/dam_bench.py 192 $out/${tag}_dam_bench_pcmgdynamic.json > $out/${tag}_dam_bench_pcmgdynamic.txt 2>&1
MP_MIC_RB=1 timeout 300 python tools/dam_bench.py 192 $out/${tag}_dam_bench_pcmic_blockrb.json > $out/${tag}_dam_bench_pcmic_blockrb.txt 2>&1
tail -30 $out/${tag}_liquid_bench.txt $out/${tag}_flip_bench.txt $out/${tag}_dam_bench.txt $out/${tag}_dam_bench_pcmgdynamic.txt $out/${tag}_dam_bench_pcmic_blockrb.txt | cut -c1-160

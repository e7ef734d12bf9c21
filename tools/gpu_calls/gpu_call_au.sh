#!/bin/bash
# ncu launch list of the bench command itself (first 400 launches) + a fresh --set full capture of the dominant kernel + launch list of a PcMIC solve in block red-black ordering
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_pcnone_512.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-configs > $out/${tag}_bench_under_ncu.log 2>&1
tail -1 $out/${tag}_bench_under_ncu.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_matvec_fused_tma --launch-skip 30 -c 1 -o $out/${tag}_matvec_tma python tools/prof_solve.py --res 512 --pc 0 --iters 60 > $out/${tag}_prof_tma.log 2>&1
ncu -i $out/${tag}_matvec_tma.ncu-rep --page raw --csv > $out/${tag}_matvec_tma_raw.csv 2>/dev/null
MP_MIC_RB=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_pcmic_blockrb_512.csv python tools/prof_solve.py --res 512 --pc 1 --iters 40 > $out/${tag}_prof_rb.log 2>&1
ls -la $out/${tag}_*

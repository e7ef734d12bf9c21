#!/bin/bash
tag=${1:-rX}
out=gpurun_out
mkdir -p $out/bound_run
timeout 1500 python -m pytest tests -m gpu -q -rf -x > $out/${tag}_pytest_gpu_x.txt 2>&1
tail -6 $out/${tag}_pytest_gpu_x.txt
# the reference's own binary, solvePressure bound to libmantapress: scenes/simpleplume.py unmodified, GPU projection vs the binary's own CPU body
R=$PWD
( cd $out/bound_run
  SCENE=$R/oracle/_ref/bound/scenes/simpleplume.py OUT=gpu MANTA_DEBUG=1 timeout 600 $R/oracle/_ref/bound/manta $R/tools/bound_scene_driver.py > gpu.log 2>&1
  SCENE=$R/oracle/_ref/bound/scenes/simpleplume.py OUT=gpu2 MANTA_DEBUG=2 timeout 600 $R/oracle/_ref/bound/manta $R/tools/bound_scene_driver.py > gpu_debug2.log 2>&1
  SCENE=$R/oracle/_ref/bound/scenes/simpleplume.py OUT=cpu MANTA_CPU_PRESSURE=1 timeout 900 $R/oracle/_ref/bound/manta $R/tools/bound_scene_driver.py > cpu.log 2>&1
  tail -2 gpu.log cpu.log; grep -c "libmantapress" gpu_debug2.log; grep "libmantapress" gpu_debug2.log | tail -3
  python - <<'PY'
import json, numpy as np
g, c = np.load("gpu.npz"), np.load("cpu.npz")
res = {"gpu_wall_s": json.load(open("gpu.json"))["wall_s"], "cpu_wall_s": json.load(open("cpu.json"))["wall_s"], "grids": {}}
for k in c.files:
    a, b = g[k].astype(np.float64), c[k].astype(np.float64)
    res["grids"][k] = {"bit_identical": bool(np.array_equal(g[k], c[k])), "max_abs_diff": float(np.abs(a - b).max()), "rel_l2": float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))}
print(json.dumps(res))
json.dump(res, open("compare.json", "w"), indent=1)
PY
  rm -f gpu.npz gpu2.npz cpu.npz waveletNoiseTile.bin )
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python -c "
import json;d=json.load(open('$out/${tag}_bench_n1.json'))
print(d['value'],d['cg_iter_per_s'],d['roofline']['frac'])
for k,v in d['configs'].items(): print(k,v['iterations'],round(v['solve_ms'],2),round(v['solve_ms_cold'],2),v.get('dominant_kernel',{}).get('frac_of_measured_peak'))
"

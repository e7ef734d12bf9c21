"""One advectSemiLagrange call per kind at res^3 for ncu (the kernels of tools/step_bench.py's advection rows).
    ncu --set full -k regex:k_semi_lagrange|k_mc_rest --launch-skip 6 -c 6 python tools/prof_advect.py 512"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
flags_h, vel_h = scenes.smoke_plume((res, res, res), 4)
dens_h = np.random.default_rng(3).random((res, res, res)).astype(np.float32)
s = mf.Solver(gridSize=(res, res, res), dim=3, prec=4)
F, V, D = mf.FlagGrid(s, flags_h), mf.MACGrid(s, vel_h), mf.RealGrid(s, dens_h)
for _ in range(reps):
    mf.advectSemiLagrange(F, V, D, order=1)          # k_semi_lagrange
    mf.advectSemiLagrange(F, V, D, order=2)          # k_semi_lagrange, k_mc_rest
    mf.advectSemiLagrange(F, V, V, order=2)          # k_semi_lagrange_mac, k_mc_rest_mac
    s.synchronize()
print("done")

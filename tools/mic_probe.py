"""Timing probe for the MIC(0) sweeps: ms per ApplyPreconditionModifiedIncompCholesky2 (forward + backward) on an all-fluid box.
   python tools/mic_probe.py 512,512,512 [prec] [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import cg, scenes  # noqa: E402

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "512,512,512").split(","))
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
sx, sy, sz = shape
s = mf.Solver(gridSize=shape, dim=3, prec=prec)
flags = scenes.closed_box_flags(*shape)
F = mf.FlagGrid(s, flags)
A0, Ai, Aj, Ak, Pc, r, z = (mf.RealGrid(s) for _ in range(7))
cg.MakeLaplaceMatrix(F, A0, Ai, Aj, Ak)
rng = np.random.default_rng(1)
r.copyFromArray(rng.standard_normal(flags.shape).astype(s.real))
cg.InitPreconditionModifiedIncompCholesky2(F, Pc, A0, Ai, Aj, Ak)
s.synchronize()
ts = []
for i in range(reps + 2):
    s.synchronize()
    t0 = time.perf_counter()
    cg.ApplyPreconditionModifiedIncompCholesky2(z, r, F, Pc, A0, Ai, Aj, Ak)
    s.synchronize()
    ts.append(1e3 * (time.perf_counter() - t0))
ts = ts[2:]
cells = sx * sy * sz
ms = float(np.median(ts))
ch = 32 // prec
iters = (sx + ch - 1) // ch + 10
print(f"mic_probe {shape} prec {prec} MP_MIC={os.environ.get('MP_MIC', '-')} EXP={os.environ.get('MP_MIC_EXP', '0')}: {ms:.3f} ms per apply "
      f"({2 * cells * 6 * prec / ms / 1e6:.0f} GB/s algorithmic; {iters} iterations per column, {1e3 * ms / 2 / iters:.2f} us per iteration if one column)")

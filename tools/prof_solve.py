"""One capped solvePressure for profiling under ncu (not a benchmark): python tools/prof_solve.py --res 512 --pc 0 --iters 40"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--prec", type=int, default=4)
ap.add_argument("--pc", type=int, default=0)
ap.add_argument("--iters", type=int, default=40)
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
flags, vel = scenes.smoke_plume(a.res, a.prec)
s = mf.Solver(gridSize=(a.res,) * 3, dim=3, prec=a.prec)
F, V, P = mf.FlagGrid(s, flags), mf.MACGrid(s, vel), mf.RealGrid(s)
fac = (a.iters + 0.5) / a.res if a.pc < 2 else 99
for r in range(a.reps):
    V.copyFromArray(vel)
    mf.solvePressure(vel=V, pressure=P, flags=F, cgAccuracy=1e-4 if a.pc >= 2 else 1e-12, cgMaxIterFac=fac, preconditioner=a.pc, zeroPressureFixing=(a.pc >= 2))
    print(mf.lastSolveInfo(), flush=True)

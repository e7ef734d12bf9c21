"""Small solves of every path for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402

for prec in (4, 8):
    for name, (flags, vel, phi) in {"smoke": scenes.smoke_plume((20, 18, 22), prec, random_vel=True) + (None,),
                                    "smoke_ragged": scenes.smoke_plume((19, 17, 15), prec, random_vel=True) + (None,),
                                    "liquid": scenes.liquid_basin((20, 22, 18), prec),
                                    "smoke2d": scenes.smoke_plume((24, 20, 1), prec, random_vel=True) + (None,)}.items():
        sz, sy, sx = flags.shape
        s = mf.Solver(gridSize=(sx, sy, sz), dim=3 if sz > 1 else 2, prec=prec)
        F, PH = mf.FlagGrid(s, flags), (mf.RealGrid(s, phi) if phi is not None else None)
        for pc in (0, 1, 2, 3):
            V, P = mf.MACGrid(s, vel), mf.RealGrid(s)
            mf.solvePressure(vel=V, pressure=P, flags=F, phi=PH, cgAccuracy=1e-5, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=(pc >= 2))
            print(name, prec, pc, mf.lastSolveInfo()["iterations"], float(np.abs(P.numpy()).max()), flush=True)
        mf.releaseMG(s)
        s.close()
print("sanitize_run done")

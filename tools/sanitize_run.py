"""Small solves of every path for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402

NEW_ONLY = "--new-only" in sys.argv       # only the kernels added after the first sanitizer pass of the round
for prec in (() if NEW_ONLY else (4, 8)):
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
# the MIC warp-column sweeps (large grids get them by default; forced here), vector and ragged rows
os.environ["MP_MIC"] = "4"
for prec in (4, 8):
    for shape in ((24, 21, 13), (19, 17, 15)):
        flags, vel = scenes.smoke_plume(shape, prec, random_vel=True)
        sz, sy, sx = flags.shape
        s = mf.Solver(gridSize=(sx, sy, sz), dim=3, prec=prec)
        V, P = mf.MACGrid(s, vel), mf.RealGrid(s)
        mf.solvePressure(vel=V, pressure=P, flags=mf.FlagGrid(s, flags), cgAccuracy=1e-5, cgMaxIterFac=99, preconditioner=mf.PcMIC)
        print("mic warp columns", prec, shape, mf.lastSolveInfo()["iterations"], flush=True)
        s.close()
del os.environ["MP_MIC"]
# the plugins either side of the projection and the other GridCg callers
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers  # noqa: E402
from cuda_impl import CudaImpl  # noqa: E402
for prec in (4, 8):
    I = CudaImpl(prec)
    for name in ("box3d", "box2d"):
        flags, vel, dens, obvel = helpers.step_scene(name, prec)
        for case in helpers.STEP_CASES:
            helpers.run_step_case(I, case, flags, vel, dens, obvel)
        print("step plugins", prec, name, flush=True)
    helpers.run_wave_steps(I, "we3d", prec, True, steps=1)
    helpers.run_guiding(I, "guide3d", prec)
    helpers.run_plume_steps(I, (12, 18, 12), prec, steps=2)
print("sanitize_run done")

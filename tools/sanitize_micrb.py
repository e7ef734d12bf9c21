"""Small PcMIC solves in block red-black ordering (k_micrb_mask, k_micrb<0|1|2>, whole-chunk and ragged rows, tiles cut by the grid, several sub-blocks per tile)
for compute-sanitizer:   compute-sanitizer --tool memcheck python tools/sanitize_micrb.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402

for prec in (4, 8):
    for shape, liquid in (((24, 21, 13), False), ((19, 17, 15), False), ((32, 20, 18), True), ((40, 37, 29), False)):
        flags, vel, phi = (scenes.liquid_basin(shape, prec) if liquid else scenes.smoke_plume(shape, prec, random_vel=True) + (None,))
        sz, sy, sx = flags.shape
        for tile in ((8, 4), (8, 8), (16, 12), (0, 0)):
            s = mf.Solver(gridSize=(sx, sy, sz), dim=3, prec=prec)
            s.setMicOrdering(1, *tile)
            V, P = mf.MACGrid(s, vel), mf.RealGrid(s)
            PH = mf.RealGrid(s, phi) if phi is not None else None
            mf.solvePressure(vel=V, pressure=P, flags=mf.FlagGrid(s, flags), phi=PH, cgAccuracy=1e-5, cgMaxIterFac=99, preconditioner=mf.PcMIC)
            print("micrb", prec, shape, tile, s.micOrdering(), mf.lastSolveInfo()["iterations"], flush=True)
            s.close()
print("done")

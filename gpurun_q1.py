import sys, time; sys.path.insert(0,'.')
import numpy as np
import mantaflow_b200 as mf
from mantaflow_b200 import scenes
for res in (128, 256):
    flags, vel = scenes.smoke_plume(res, 4)
    s = mf.Solver(gridSize=(res,res,res), dim=3, prec=4)
    F,V,P = mf.FlagGrid(s,flags), mf.MACGrid(s,vel), mf.RealGrid(s)
    for pc in (0,1):
        V.copyFromArray(vel)
        mf.solvePressure(vel=V,pressure=P,flags=F,cgAccuracy=1e-4,cgMaxIterFac=99,preconditioner=pc)
        i=mf.lastSolveInfo(); print(res,pc,i, flush=True)
    s.close()

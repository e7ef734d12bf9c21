"""One step of the main loop of scenes/benchmark_dam.py:100-134 (the reference's FLIP benchmark), device resident, timed plugin by plugin.
    python tools/dam_bench.py [res] [out.json]       # default 192: a dam of res/3 x 2res/3 x res cells with 8 particles per cell
Every call is bracketed by a stream synchronisation, so the per-plugin times include the launch gaps a scene would overlap; the `step` line is one
un-bracketed pass.  Shares of the step are what to compare with the reference's profile (mapPartsToMAC 8.7 %, gridParticleIndex + unionParticleLevelset 13.5 %,
extrapolateMACSimple 14.8 % of a CPU step)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 192
outp = sys.argv[2] if len(sys.argv) > 2 else None
sx = sy = sz = res
flags_h = scenes.closed_box_flags(sx, sy, sz)
rng = np.random.default_rng(5)
ii, jj, kk = np.meshgrid(np.arange(1, sx // 3), np.arange(1, 2 * sy // 3), np.arange(1, sz - 1), indexing="ij")
base = np.repeat(np.stack([ii.ravel(), jj.ravel(), kk.ravel()], 1).astype(np.float32), 8, 0)
pos = base + rng.random(base.shape, dtype=np.float32)
pos = np.ascontiguousarray(pos[rng.permutation(len(pos))]); del base
N = len(pos)
k, j, i = np.ogrid[0:sz, 0:sy, 0:sx]
phiObs_h = np.minimum(np.minimum(np.minimum(i + 0.5 - 1, sx - 1.5 - i), np.minimum(j + 0.5 - 1, sy - 1.5 - j)), np.minimum(k + 0.5 - 1, sz - 1.5 - k)).astype(np.float32)

s = mf.Solver(gridSize=(sx, sy, sz), dim=3, prec=4)
s.timestep = 0.8
flags, vel, velOld, pressure = mf.FlagGrid(s, flags_h), s.create(mf.MACGrid), s.create(mf.MACGrid), s.create(mf.RealGrid)
phi, phiObs, index = s.create(mf.LevelsetGrid), mf.LevelsetGrid(s, phiObs_h), s.create(mf.IntGrid)
pp = s.create(mf.BasicParticleSystem)
pV, pVtmp, pT = pp.create(mf.PdataVec3), pp.create(mf.PdataVec3), pp.create(mf.PdataInt)
pindex = s.create(mf.ParticleIndexSystem)
pp.setParticles(pos)
pT.setConst(mf.FlagFluid)
mf.markFluidCells(parts=pp, flags=flags, ptype=pT)
grav, dx, bnd = -0.01, 1.0, 1
FlagFluid, FlagEmpty = mf.FlagFluid, mf.FlagEmpty

PC = int(os.environ.get("DAM_PC", "1"))          # the scene's own call takes the plugin's default PcMIC; DAM_PC=2: PcMGDynamic (BASELINE config 2 names PcMG)
PC_NAME = {0: "PcNone", 1: "PcMIC", 2: "PcMGDynamic", 3: "PcMGStatic"}[PC]
calls = [   # scenes/benchmark_dam.py:100-134, ghost-fluid variant
    ("mapPartsToMAC", lambda: mf.mapPartsToMAC(vel=vel, flags=flags, velOld=velOld, parts=pp, partVel=pV, ptype=pT, exclude=FlagEmpty)),
    ("getMaxAbs (adaptTimestep)", lambda: vel.getMaxAbs()),
    ("addGravityNoScale", lambda: mf.addGravityNoScale(flags=flags, vel=vel, gravity=(0, grav, 0))),
    ("gridParticleIndex", lambda: mf.gridParticleIndex(parts=pp, flags=flags, indexSys=pindex, index=index)),
    ("unionParticleLevelset", lambda: mf.unionParticleLevelset(parts=pp, indexSys=pindex, flags=flags, index=index, phi=phi, radiusFactor=1.0)),
    ("extrapolateLsSimple", lambda: mf.extrapolateLsSimple(phi=phi, distance=4, inside=True)),
    ("setWallBcs", lambda: mf.setWallBcs(flags=flags, vel=vel)),
    ("solvePressure (%s, phi)" % PC_NAME, lambda: mf.solvePressure(flags=flags, vel=vel, pressure=pressure, phi=phi, cgAccuracy=1e-3, preconditioner=PC)),
    ("setWallBcs 2", lambda: mf.setWallBcs(flags=flags, vel=vel)),
    ("extrapolateMACSimple", lambda: mf.extrapolateMACSimple(flags=flags, vel=vel)),
    ("flipVelocityUpdate", lambda: mf.flipVelocityUpdate(vel=vel, velOld=velOld, flags=flags, parts=pp, partVel=pV, flipRatio=0.97, ptype=pT, exclude=FlagEmpty)),
    ("addForcePvel", lambda: mf.addForcePvel(vel=pV, a=(0, grav, 0), dt=s.timestep, ptype=pT, exclude=FlagFluid)),
    ("getPosPdata", lambda: pp.getPosPdata(target=pVtmp)),
    ("advectInGrid RK4", lambda: pp.advectInGrid(flags=flags, vel=vel, integrationMode=mf.IntRK4, deleteInObstacle=False, ptype=pT, exclude=FlagEmpty)),
    ("eulerStep", lambda: mf.eulerStep(parts=pp, vel=pV, ptype=pT, exclude=FlagFluid)),
    ("projectOutOfBnd", lambda: pp.projectOutOfBnd(flags=flags, bnd=bnd + dx * 0.5, plane="xXyYzZ", ptype=pT)),
    ("pushOutofObs", lambda: mf.pushOutofObs(parts=pp, flags=flags, phiObs=phiObs, thresh=dx * 0.5, ptype=pT)),
    ("updateVelocityFromDeltaPos", lambda: mf.updateVelocityFromDeltaPos(parts=pp, vel=pV, x_prev=pVtmp, dt=s.timestep, ptype=pT, exclude=FlagFluid)),
    ("markFluidCells", lambda: mf.markFluidCells(parts=pp, flags=flags, ptype=pT)),
    ("setPartType 1", lambda: mf.setPartType(parts=pp, ptype=pT, mark=FlagFluid, stype=FlagEmpty, flags=flags, cflag=FlagFluid)),
    ("markIsolatedFluidCell", lambda: mf.markIsolatedFluidCell(flags=flags, mark=FlagEmpty)),
    ("setPartType 2", lambda: mf.setPartType(parts=pp, ptype=pT, mark=FlagEmpty, stype=FlagFluid, flags=flags, cflag=FlagEmpty)),
]


def step(timed=None):
    for name, fn in calls:
        if timed is None:
            fn()
            continue
        s.synchronize(); t0 = time.perf_counter(); fn(); s.synchronize()
        timed[name] = timed.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)


for _ in range(2):
    step()
s.synchronize()
reps, per = 3, {}
for _ in range(reps):
    step(per)
s.synchronize(); t0 = time.perf_counter()
for _ in range(reps):
    step()
s.synchronize()
whole = 1e3 * (time.perf_counter() - t0) / reps
total = sum(per.values()) / reps
out = {"res": res, "particles": N, "step_ms": whole, "sum_of_plugins_ms": total, "plugins_ms": {k: v / reps for k, v in per.items()},
       "last_solve": {k: v for k, v in dict(mf.lastSolveInfo() or {}).items() if isinstance(v, (int, float))}}
print(f"# benchmark_dam loop, {res}^3 float, {N} particles, one B200: {whole:.2f} ms per step ({total:.2f} ms as the sum of synchronised plugins)")
for name, ms in sorted(out["plugins_ms"].items(), key=lambda kv: -kv[1]):
    print(f"{name:32s} {ms:9.3f} ms  {100 * ms / total:5.1f} %")
if outp:
    json.dump(out, open(outp, "w"), indent=1)

"""Timing of the FLIP particle <-> grid plugins on the device (second slice) on a basin + drop filled with 8 particles per
liquid cell (the sampling density of scenes/benchmark_dam.py / flip02_surface.py: `sampleLevelsetWithParticles(discretization=2)`).
    python tools/flip_bench.py [res] [out.json]       # default 256 (about 34 M particles at 256^3 ... 8 per liquid cell)
Bytes: what a plugin must move at least per particle (P) or per cell (C), in float: pos 12, flag 4, vel 12, flags/index/phi 4, MAC 12."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
outp = sys.argv[2] if len(sys.argv) > 2 else None
PEAK = 6546.6
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
k, j, i = np.ogrid[0:res, 0:res, 0:res]
drop = np.sqrt(((i + 0.5 - 0.5 * res) ** 2 + (j + 0.5 - 0.5 * res) ** 2 + (k + 0.5 - 0.5 * res) ** 2).astype(np.float32)) - np.float32(0.125 * res)
phi_h = np.minimum(drop, ((j + 0.5) - 0.2 * res).astype(np.float32))
flags_h = scenes.closed_box_flags(res, res, res, boundaryWidth=1)
liquid = (phi_h < 0) & ((flags_h & mf.FlagObstacle) == 0)
kk, jj, ii = np.nonzero(liquid)
rng = np.random.default_rng(3)
base = np.repeat(np.stack([ii, jj, kk], 1).astype(np.float32), 8, 0)
pos = base + rng.random(base.shape, dtype=np.float32)
perm = rng.permutation(len(pos))                       # particles of a running simulation are not sorted by cell
pos = np.ascontiguousarray(pos[perm]); del base, perm
pvel = (rng.random(pos.shape, dtype=np.float32) * 2 - 1) * np.float32(0.3)
N, n = len(pos), res ** 3

s = mf.Solver(gridSize=(res, res, res), dim=3, prec=4)
F = mf.FlagGrid(s, flags_h)
vel, velOld, weight = s.create(mf.MACGrid), s.create(mf.MACGrid), s.create(mf.VecGrid)
phi, index = s.create(mf.LevelsetGrid), s.create(mf.IntGrid)
pp = s.create(mf.BasicParticleSystem)
pVel, pindex = pp.create(mf.PdataVec3), s.create(mf.ParticleIndexSystem)
phiObs = mf.LevelsetGrid(s, np.minimum(np.minimum(np.minimum(i + 0.5 - 1, res - 1.5 - i), np.minimum(j + 0.5 - 1, res - 1.5 - j)), np.minimum(k + 0.5 - 1, res - 1.5 - k)).astype(np.float32))
s.timestep = 0.5
pp.setParticles(pos)
pVel.copyFromArray(pvel)
mf.markFluidCells(pp, F); s.synchronize()


def timed(fn, reps=3):
    fn(); s.synchronize()
    ts = []
    for _ in range(reps):
        s.synchronize(); t0 = time.perf_counter(); fn(); s.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts))


rows = [   # name, call, minimum bytes
    ("markFluidCells", lambda: mf.markFluidCells(pp, F), 8 * n + 16 * N),
    ("mapPartsToMAC (+ weight)", lambda: mf.mapPartsToMAC(F, vel, velOld, pp, pVel, weight=weight), 28 * N + 36 * n),
    ("mapPartsToMAC", lambda: mf.mapPartsToMAC(F, vel, velOld, pp, pVel), 28 * N + 24 * n),
    ("mapMACToParts", lambda: mf.mapMACToParts(F, vel, pp, pVel), 28 * N + 12 * n),
    ("flipVelocityUpdate", lambda: mf.flipVelocityUpdate(F, vel, velOld, pp, pVel, 0.97), 40 * N + 24 * n),
    ("advectInGrid RK4", lambda: pp.advectInGrid(F, vel, mf.IntRK4, deleteInObstacle=False), 32 * N + 12 * n),
    ("projectOutOfBnd", lambda: pp.projectOutOfBnd(F, 1.5), 28 * N),
    ("pushOutofObs", lambda: mf.pushOutofObs(pp, F, phiObs, thresh=0.5), 28 * N + 4 * n),
    ("gridParticleIndex", lambda: mf.gridParticleIndex(pp, pindex, F, index), 20 * N + 8 * n),
    ("unionParticleLevelset", lambda: mf.unionParticleLevelset(pp, pindex, F, index, phi), 16 * N + 8 * n),
]
only = os.environ.get("FLIP_BENCH_ONLY")          # e.g. FLIP_BENCH_ONLY=mapPartsToMAC under ncu
if only:
    rows = [r for r in rows if r[0] == only]
out = {"res": res, "prec": 4, "particles": N, "peak_gbs": PEAK, "plugins": {}}
print(f"# {res}^3 float, {N} particles (8 per liquid cell, shuffled), one B200", flush=True)
for name, fn, nbytes in rows:
    ms = timed(fn)
    gbs = nbytes / ms / 1e6
    out["plugins"][name] = {"ms": ms, "min_bytes": nbytes, "gbs": gbs, "frac_of_peak": gbs / PEAK, "ns_per_particle": 1e6 * ms / N}
    print(f"{name:28s} {ms:9.3f} ms  {1e6 * ms / N:7.3f} ns/particle  {gbs:7.0f} GB/s of minimum traffic  {gbs / PEAK:5.2f} of measured HBM peak", flush=True)
    if outp:
        json.dump(out, open(outp, "w"), indent=1)

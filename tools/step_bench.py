"""Timing of the device plugins either side of the projection and of a whole simpleplume-like step, with the
reference's CPU plugins timed beside them on a bounded grid.
    python tools/step_bench.py [res] [cpu_res]        # default 512 (grid res x 1.5res x res would not fit the CPU leg: cubic here)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cpu_res = int(sys.argv[2]) if len(sys.argv) > 2 else 128
prec = 4
PEAK = 6546.6
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def plume_fields(n, real):
    shape = (n, n, n)
    flags, vel = scenes.smoke_plume(shape, prec)
    rng = np.random.default_rng(3)
    dens = rng.random(shape).astype(real)
    return flags, vel, dens


def timed(s, fn, reps=5):
    fn(); s.synchronize()
    ts = []
    for _ in range(reps):
        s.synchronize(); t0 = time.perf_counter(); fn(); s.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts))


real = np.float32
flags_h, vel_h, dens_h = plume_fields(res, real)
s = mf.Solver(gridSize=(res, res, res), dim=3, prec=prec)
F, V, D, P = mf.FlagGrid(s, flags_h), mf.MACGrid(s, vel_h), mf.RealGrid(s, dens_h), mf.RealGrid(s)
F.dev(); V.dev(); D.dev()
n = res ** 3
w = prec
# algorithmic bytes per cell: what one pass must read and write at least (flags 4 B, Real w, Vec3 3w)
rows = [
    ("setWallBcs", lambda: mf.setWallBcs(F, V), 4 + 6 * w),
    ("addBuoyancy", lambda: mf.addBuoyancy(F, D, V, (0, -6e-4, 0)), 4 + w + 6 * w),
    ("addGravity", lambda: mf.addGravity(F, V, (0, -1e-3, 0)), 4 + 6 * w),
    ("advectSemiLagrange density order 1", lambda: mf.advectSemiLagrange(F, V, D, order=1), 3 * w + 2 * w),      # vel + src read, result written
    # two passes: forward trace (vel + src read, fwd written), then backward trace + correction + clamping (vel, fwd, orig, flags read, result written)
    ("advectSemiLagrange density order 2 (MacCormack)", lambda: mf.advectSemiLagrange(F, V, D, order=2), (3 * w + 2 * w) + (4 + 3 * w + 3 * w)),
    ("advectSemiLagrange vel order 2 (MacCormack)", lambda: mf.advectSemiLagrange(F, V, V, order=2), (3 * w + 3 * w + 3 * w) + (4 + 3 * w + 3 * w + 3 * w + 3 * w)),
    # round 2: cubic lookups (64 reads per lookup, all from the neighbourhood of the traced position) and explicit-midpoint back-tracing
    ("advectSemiLagrange density order 2, orderSpace 2", lambda: mf.advectSemiLagrange(F, V, D, order=2, orderSpace=2), (3 * w + 2 * w) + (4 + 3 * w + 3 * w)),
    ("advectSemiLagrange density order 2, orderTrace 2", lambda: mf.advectSemiLagrange(F, V, D, order=2, orderTrace=2), (3 * w + 2 * w) + (4 + 3 * w + 3 * w)),
    ("advectSemiLagrange vel order 2, orderSpace 2 orderTrace 2", lambda: mf.advectSemiLagrange(F, V, V, order=2, orderSpace=2, orderTrace=2), (3 * w + 3 * w + 3 * w) + (4 + 3 * w + 3 * w + 3 * w + 3 * w)),
]
out = {"res": res, "prec": prec, "peak_gbs": PEAK, "plugins": {}}
print(f"# {res}^3 float, one B200; algorithmic bytes = compulsory reads + writes of every pass of the plugin")
for name, fn, bpc in rows:
    ms = timed(s, fn)
    gbs = bpc * n / ms / 1e6
    out["plugins"][name] = {"ms": ms, "bytes_per_cell": bpc, "gbs": gbs, "frac_of_peak": gbs / PEAK}
    print(f"{name:50s} {ms:8.3f} ms  {bpc:4d} B/cell  {gbs:7.0f} GB/s  {gbs / PEAK:5.2f} of measured HBM peak")

# whole step, device resident (scenes/simpleplume.py:48-60 without the noise inflow), PcMGStatic as a production setting and the scene's default PcMIC
V.copyFromArray(np.zeros_like(vel_h)); D.copyFromArray(np.zeros_like(dens_h))
src = np.zeros(dens_h.shape, bool); src[int(0.08 * res):int(0.14 * res), int(0.4 * res):int(0.6 * res), int(0.4 * res):int(0.6 * res)] = True
d = D.numpy(writable=True); d[src.transpose(1, 0, 2)] = 1
for pc, pcname in ((mf.PcMGStatic, "PcMGStatic"), (mf.PcMIC, "PcMIC")):
    def step():
        mf.advectSemiLagrange(flags=F, vel=V, grid=D, order=2)
        mf.advectSemiLagrange(flags=F, vel=V, grid=V, order=2, strength=1.0)
        mf.setWallBcs(flags=F, vel=V)
        mf.addBuoyancy(density=D, vel=V, gravity=(0, -6e-4, 0), flags=F)
        mf.solvePressure(flags=F, vel=V, pressure=P, preconditioner=pc, zeroPressureFixing=(pc != mf.PcMIC))
    for _ in range(3):
        step()
    ms = timed(s, step, reps=3)
    it = mf.lastSolveInfo()["iterations"]
    out["step_" + pcname] = {"ms": ms, "iterations_last": it}
    print(f"whole step ({pcname}, {it} CG iterations in the last solve)        {ms:8.2f} ms")

# the reference's CPU plugins on cpu_res^3
try:
    from oracle.oracle_api import Oracle, available
    if available("reference", prec):
        R = Oracle("reference", prec)
        fl, ve, de = plume_fields(cpu_res, real)
        cpu = {}
        for name, fn in (("setWallBcs", lambda: R.set_wall_bcs_obvel(fl, ve.copy(), None)),
                         ("addBuoyancy", lambda: R.add_buoyancy(fl, de, ve.copy(), (0, -6e-4, 0))),
                         ("advectSemiLagrange density order 2 (MacCormack)", lambda: R.advect_semi_lagrange(fl, ve, de.copy(), order=2)),
                         ("advectSemiLagrange vel order 2 (MacCormack)", lambda: R.advect_semi_lagrange(fl, ve, ve.copy(), order=2))):
            fn(); t0 = time.perf_counter(); fn(); ms = 1e3 * (time.perf_counter() - t0)
            cpu[name] = ms
            g = out["plugins"][name]["ms"] * (cpu_res / res) ** 3
            print(f"reference CPU ({os.cpu_count()} cores) {cpu_res}^3 {name:50s} {ms:8.2f} ms   (device, scaled to {cpu_res}^3: {g:.3f} ms, x{ms / g:.0f})")
        out["cpu_reference"] = {"res": cpu_res, "cores": os.cpu_count(), "ms": cpu}
except Exception as e:      # the reference library is test infrastructure; the bench still reports the device numbers without it
    print("cpu leg skipped:", e)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "step_bench.json"), "w"), indent=1)

"""PD_fluid_guiding on the device with the reference's CPU plugin beside it on a bounded grid.
    python tools/guiding_bench.py [res] [cpu_res]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import helpers  # noqa: E402
import mantaflow_b200 as mf  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cpu_res = int(sys.argv[2]) if len(sys.argv) > 2 else 96
prec = 4
out = {}


def scene(n):
    helpers.GUIDING_SCENES["bench"] = ((n, n, n), 3)
    try:
        return helpers.guiding_scene("bench", prec)
    finally:
        del helpers.GUIDING_SCENES["bench"]


kw = dict(blurRadius=2, sigma=0.99, maxIters=40, cgAccuracy=1e-4)
for n in (cpu_res, res):
    flags, vel, velT, w, _ = scene(n)
    s = mf.Solver(gridSize=(n, n, n), dim=3, prec=prec)
    F, VT, W, P = mf.FlagGrid(s, flags), mf.MACGrid(s, velT), mf.RealGrid(s, w), mf.RealGrid(s)
    ts = []
    for rep in range(3):
        V = mf.MACGrid(s, vel); V.dev(); s.synchronize()
        t0 = time.perf_counter()
        mf.PD_fluid_guiding(V, VT, P, F, W, preconditioner=mf.PcMGStatic, zeroPressureFixing=True, **kw)
        s.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    it = mf.lastGuidingIterations()
    out["gpu_%d" % n] = {"ms": min(ts[1:]), "first_ms": ts[0], "pd_iterations": it}
    print(f"device {n}^3 float PcMGStatic: {it + 1} primal-dual iterations (solvePressure calls), {min(ts[1:]):.1f} ms  (first call incl. hierarchy build {ts[0]:.1f} ms), "
          f"{min(ts[1:]) / (it + 1):.2f} ms per iteration")
    mf.releaseMG(s); s.close()
try:
    from oracle.oracle_api import Oracle, available
    if available("reference", prec):
        R = Oracle("reference", prec)
        flags, vel, velT, w, _ = scene(cpu_res)
        v = vel.copy(); t0 = time.perf_counter()
        p, it = R.pd_fluid_guiding(flags, v, velT, w, preconditioner=3, zeroPressureFixing=True, **kw)
        ms = 1e3 * (time.perf_counter() - t0)
        out["cpu_reference_%d" % cpu_res] = {"ms": ms, "pd_iterations": it, "cores": os.cpu_count()}
        g = out["gpu_%d" % cpu_res]
        print(f"reference CPU ({os.cpu_count()} cores) {cpu_res}^3: {it + 1} iterations, {ms:.0f} ms  -> device x{ms / g['ms']:.0f} at equal size")
except Exception as e:
    print("cpu leg skipped:", e)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r1_guiding_bench.json"), "w"), indent=1)

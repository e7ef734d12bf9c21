"""Timing of the liquid-neighbour plugins on the device (first slice) on a basin + drop level set.
    python tools/liquid_bench.py [res] [out.json]       # default 512
Algorithmic bytes per cell = what the passes of a plugin must move at least: the mark pass reads its input (flags 4 B or phi w) and writes
4 B of marks, every extrapolation pass reads the 4 B of marks again (values are only touched in the thin band that advances)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
outp = sys.argv[2] if len(sys.argv) > 2 else None
PEAK = 6546.6
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
n, w = res ** 3, 4
k, j, i = np.ogrid[0:res, 0:res, 0:res]
drop = np.sqrt(((i + 0.5 - 0.5 * res) ** 2 + (j + 0.5 - 0.5 * res) ** 2 + (k + 0.5 - 0.5 * res) ** 2).astype(np.float32)) - np.float32(0.125 * res)
phi_h = np.minimum(drop, ((j + 0.5) - 0.2 * res).astype(np.float32))
s = mf.Solver(gridSize=(res, res, res), dim=3, prec=4)
F = mf.FlagGrid(s, scenes.closed_box_flags(res, res, res, boundaryWidth=1))
P = mf.LevelsetGrid(s, phi_h)
V = s.create(mf.MACGrid)
v = V.numpy(writable=True); v[:, : res // 4, :, 1] = -0.3
F.updateFromLevelset(P); s.synchronize()


def timed(fn, reps=3):
    fn(); s.synchronize()
    ts = []
    for _ in range(reps):
        s.synchronize(); t0 = time.perf_counter(); fn(); s.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts))


rows = [
    ("FlagGrid.updateFromLevelset", lambda: F.updateFromLevelset(P), 4 + w + 4),
    ("extrapolateLsSimple distance 5 outside", lambda: mf.extrapolateLsSimple(P, distance=5, inside=False), (w + 4) + 4 * 4),
    ("extrapolateLsSimple distance 5 inside", lambda: mf.extrapolateLsSimple(P, distance=5, inside=True), (w + 4) + 4 * 4),
    ("extrapolateMACSimple distance 5", lambda: mf.extrapolateMACSimple(F, V, distance=5), (4 + 4) + 5 * 4),
    ("Grid.setBound (outer layers only)", lambda: P.setBound(1, 1), 0),
]
out = {"res": res, "prec": 4, "peak_gbs": PEAK, "plugins": {}}
print(f"# {res}^3 float, one B200", flush=True)
for name, fn, bpc in rows:
    ms = timed(fn)
    gbs = bpc * n / ms / 1e6
    out["plugins"][name] = {"ms": ms, "bytes_per_cell": bpc, "gbs": gbs, "frac_of_peak": gbs / PEAK}
    print(f"{name:44s} {ms:8.3f} ms  {bpc:3d} B/cell  {gbs:7.0f} GB/s  {gbs / PEAK:5.2f} of measured HBM peak", flush=True)
    if outp:
        json.dump(out, open(outp, "w"), indent=1)

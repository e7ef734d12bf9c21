"""PcMIC solves of the bench's plume at res^3 with the reference's lexicographic MIC(0) and with the block red-black ordering (mp_set_mic_ordering)
for a list of tile shapes: iterations, solve time, time per application of the preconditioner.
    python tools/micrb_bench.py [res] [prec] [tiles "8x4,8x8,16x8"] [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tiles = [tuple(int(v) for v in t.split("x")) for t in (sys.argv[3] if len(sys.argv) > 3 else "0x0,8x4,8x8,16x8").split(",")]       # 0x0: chosen from the grid
lex = os.environ.get("MICRB_SKIP_LEX", "0") != "1"
flags, vel = scenes.smoke_plume((res, res, res), prec)
s = mf.Solver(gridSize=(res, res, res), dim=3, prec=prec)
s.setProfiling(4)
F, V0, V, P = mf.FlagGrid(s, flags), mf.MACGrid(s, vel), mf.MACGrid(s), mf.RealGrid(s)
F.dev(); V0.dev()
cells = res ** 3
rows = []
for t in ([None] if lex else []) + tiles:
    if t is None:
        s.setMicOrdering(0)
    else:
        s.setMicOrdering(1, *t)
    best = None
    for rep in range(2):
        mf._lib.check(s.lib.mp_grid_copy_from(V.dev(), V0.dev()))
        mf.solvePressure(vel=V, pressure=P, flags=F, cgAccuracy=1e-4, cgMaxIterFac=99, preconditioner=mf.PcMIC)
        info = mf.lastSolveInfo()
        if best is None or info["msTotal"] < best["msTotal"]:
            best = info
    name = "lexicographic (reference ordering)" if t is None else "block red-black %dx%d" % s.micOrdering()[1:]
    # per application and cell: 4 Real read + 2 written + 2 mask bytes (+ the edge rows, not counted)
    bpc = 6 * prec + 2 if t is not None else 12 + 12 * prec
    gbs = bpc * cells / (best["msPrecondAvg"] * 1e-3) / 1e9 if best["msPrecondAvg"] > 0 else 0.0
    row = {"ordering": name, "iterations": best["iterations"], "solve_ms": best["msTotal"], "ms_per_iteration": best["msTotal"] / max(best["iterations"], 1),
           "precond_ms": best["msPrecondAvg"], "precond_bytes_per_cell": bpc, "precond_gbs": gbs, "matvec_ms": best["msMatvecAvg"], "axpy_ms": best["msAxpyAvg"],
           "update_ms": best["msUpdateAvg"], "max_divergence": None}
    rows.append(row)
    print("%-36s iterations %4d  solve %9.2f ms  %.3f ms/iteration  preconditioner %.3f ms (%5.0f GB/s on %d B/cell)" %
          (name, row["iterations"], row["solve_ms"], row["ms_per_iteration"], row["precond_ms"], gbs, bpc), flush=True)
if len(sys.argv) > 4:
    json.dump({"res": res, "prec": prec, "rows": rows}, open(sys.argv[4], "w"), indent=1)

"""PcMGStatic / PcMGDynamic solves with the level-0 V-cycle kernels fused (default) and per colour (MP_MG_L0FUSED=0), one B200:
python tools/mg_bench.py [res] [out.json].  Device-resident solvePressure calls, kernel times from the library's sampled CUDA events."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
out_path = sys.argv[2] if len(sys.argv) > 2 else None
rows = []


def run(prec, pc, fused, reps, envs=None):
    os.environ["MP_MG_L0FUSED"] = str(fused)
    for k, v in (envs or {}).items():
        os.environ[k] = v
    flags, vel = scenes.smoke_plume(res, prec)
    s = mf.Solver(gridSize=(res,) * 3, dim=3, prec=prec)
    s.setProfiling(1)
    F, V0, V, P = mf.FlagGrid(s, flags), mf.MACGrid(s, vel), mf.MACGrid(s), mf.RealGrid(s)
    F.dev(); V0.dev()
    ms = []
    for r in range(reps):
        mf._lib.check(s.lib.mp_grid_copy_from(V.dev(), V0.dev()))
        mf.solvePressure(vel=V, pressure=P, flags=F, cgAccuracy=1e-4, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=True)
        info = mf.lastSolveInfo()
        ms.append(info["msTotal"])
    p = P.numpy()
    row = {"res": res, "prec": prec, "pc": pc, "l0fused": fused, "env": envs or {}, "iterations": info["iterations"], "solve_ms_cold": ms[0], "solve_ms_warm": min(ms[1:]),
           "vcycle_ms": info["msPrecondAvg"], "matvec_ms": info["msMatvecAvg"], "axpy_ms": info["msAxpyAvg"], "update_ms": info["msUpdateAvg"],
           "stage_ms": {k: info[k] for k in ("msRhs", "msMatrix", "msSolve", "msCorrect")},
           "pressure_checksum": float(abs(p).sum(dtype="float64"))}
    rows.append(row)
    print(json.dumps(row), flush=True)
    if pc == 3:
        mf.releaseMG(s)
    s.close()
    for k in (envs or {}):
        os.environ.pop(k, None)


# lazy module loading: touch every kernel once on a small grid
_r = res; res = 32
for prec in (4, 8):
    run(prec, 3, 1, 2); run(prec, 3, 0, 2)
rows.clear(); res = _r
for prec in (4, 8):
    run(prec, 3, 1, 4)
    run(prec, 3, 0, 4)
    run(prec, 3, 0, 4, {"MP_MG_L0MASK": "0"})
run(4, 2, 0, 3)
if out_path:
    json.dump(rows, open(out_path, "w"), indent=1)

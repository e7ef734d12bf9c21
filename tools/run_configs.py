"""Runs the BASELINE.json configurations on one GPU next to the reference's CPU implementation on the host cores and
writes a markdown + json table (profiles/<tag>_configs.{md,json}).  Not the bench contract (that is bench.py); this is the
per-config evidence: iterations, device ms, plugin (host-buffer) ms, CPU reference ms, parity of the converged pressure.

    python tools/run_configs.py --tag r1 [--skip-cpu] [--max-cpu-res 256]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402

PC = {0: "PcNone", 1: "PcMIC", 2: "PcMGDynamic", 3: "PcMGStatic"}


def rel_l2(a, b):
    return float(np.linalg.norm((a.astype(np.float64) - b.astype(np.float64)).ravel()) / max(np.linalg.norm(b.astype(np.float64).ravel()), 1e-300))


def demean(p, flags):
    fl = (flags & 1) != 0
    q = p.astype(np.float64).copy()
    q[fl] -= q[fl].mean()
    return q


def gpu_solve(flags, vel, phi, prec, pc, acc, fac, fix, reps=2, mic_rb=False):
    sz, sy, sx = flags.shape
    s = mf.Solver(gridSize=(sx, sy, sz), dim=3 if sz > 1 else 2, prec=prec)
    if mic_rb:            # PcMIC reformulated: MIC(0) in block red-black ordering, tile chosen from the grid
        s.setMicOrdering(1)
    F, V, P = mf.FlagGrid(s, flags), mf.MACGrid(s, vel), mf.RealGrid(s)
    PH = mf.RealGrid(s, phi) if phi is not None else None
    kw = dict(cgAccuracy=acc, cgMaxIterFac=fac, preconditioner=pc, zeroPressureFixing=fix)
    best = None
    for r in range(reps):
        V.copyFromArray(vel); V.dev()
        mf.solvePressure(vel=V, pressure=P, flags=F, phi=PH, **kw)
        info = mf.lastSolveInfo()
        if r == 0:
            cold = info["msTotal"]
        best = info if best is None or info["msTotal"] < best["msTotal"] else best
    p, v = P.numpy().copy(), V.numpy().copy()
    # plugin through host buffers (H2D + solve + D2H)
    hv, hp = vel.copy(), np.zeros(flags.shape, vel.dtype)
    t0 = time.perf_counter()
    mf.solvePressureHost(s, hv, hp, flags, phi=phi, **kw)
    host_ms = 1e3 * (time.perf_counter() - t0)
    mf.releaseMG(s)
    order = s.micOrdering() if mic_rb else (0, 0, 0)
    s.close()
    return dict(order=order, iterations=best["iterations"], resNorm=best["resNorm"], ms=best["msTotal"], ms_cold=cold, ms_solve=best["msSolve"], host_ms=host_ms, p=p, v=v)


def cpu_solve(O, flags, vel, phi, pc, acc, fac, fix):
    v = vel.copy()
    t0 = time.perf_counter()
    if pc == 0:
        # the reference plugin asserts on PcNone: rhs + matrix + GridCg + correctVelocity driven directly
        rhs, _, _ = O.compute_rhs(flags, v, phi=phi)
        A = O.make_matrix(flags, phi=phi)
        maxdim = max(flags.shape)
        p, it, rn = O.cg_solve(flags, rhs, *A, pc=0, accuracy=acc, maxIter=int(np.float32(fac) * maxdim))
        O.correct_velocity(flags, v, p, phi=phi)
    else:
        p, it, rn = O.solve_pressure(flags, v, phi=phi, cgAccuracy=acc, cgMaxIterFac=fac, preconditioner=pc, zeroPressureFixing=fix)
    return dict(iterations=it, resNorm=rn, ms=1e3 * (time.perf_counter() - t0), p=p, v=v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r1")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--max-cpu-res", type=int, default=256)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    cores = os.cpu_count()
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    from oracle.oracle_api import Oracle, available

    cfgs = []
    # config 1: scenes/simpleplume.py geometry 64x96x64, PcMIC, cgAccuracy 1e-3 (scene defaults), float
    cfgs.append(("1 simpleplume-like 64x96x64 smoke", lambda prec: scenes.smoke_plume((64, 96, 64), prec, obstacle=False) + (None,), 4, [(1, 1e-3, 1.5, False), (1, 1e-3, 1.5, False, True), (3, 1e-3, 1.5, True)]))
    # config 2: benchmark_dam.py geometry 88x83x33 liquid with free surface (phi ghost fluid)
    cfgs.append(("2 dam-like 88x83x33 liquid+phi", lambda prec: scenes.liquid_basin((88, 83, 33), prec), 4, [(1, 1e-3, 1.5, False), (1, 1e-3, 1.5, False, True), (2, 1e-3, 1.5, False)]))
    # config 3: 256^3 smoke plume with obstacle, PcNone vs PcMIC vs PcMGStatic, cgAccuracy 1e-4
    cfgs.append(("3 synthetic 256^3 smoke+obstacle", lambda prec: scenes.smoke_plume(256, prec) + (None,), 4, [(0, 1e-4, 99, False), (1, 1e-4, 99, False), (1, 1e-4, 99, False, True), (3, 1e-4, 99, True)]))
    # config 4: 512^3 single GPU, float and double
    cfgs.append(("4 synthetic 512^3 smoke+obstacle", lambda prec: scenes.smoke_plume(512, prec) + (None,), 4, [(0, 1e-4, 99, False), (1, 1e-4, 99, False), (1, 1e-4, 99, False, True), (3, 1e-4, 99, True)]))
    cfgs.append(("4 synthetic 512^3 smoke+obstacle", lambda prec: scenes.smoke_plume(512, prec) + (None,), 8, [(0, 1e-4, 99, False), (3, 1e-4, 99, True)]))
    rows = []
    for name, make, prec, runs in cfgs:
        if a.only and a.only not in name:
            continue
        flags, vel, phi = make(prec)
        res = max(flags.shape)
        O = None
        if not a.skip_cpu and res <= a.max_cpu_res:
            O = Oracle("reference" if available("reference", prec) else "port", prec)
        cpu_cache = {}
        for run in runs:
            pc, acc, fac, fix = run[:4]
            rb = len(run) > 4 and run[4]
            g = gpu_solve(flags, vel, phi, prec, pc, acc, fac, fix, mic_rb=rb)
            pcname = PC[pc] if not rb else "PcMIC reformulated (block red-black %dx%d; CPU column: the reference's PcMIC)" % g["order"][1:]
            row = dict(config=name, prec="f32" if prec == 4 else "f64", pc=pcname, cgAccuracy=acc, gpu_iterations=g["iterations"], gpu_resNorm=g["resNorm"],
                       gpu_ms=g["ms"], gpu_ms_cold=g["ms_cold"], gpu_host_ms=g["host_ms"], max_div_after=scenes.max_divergence(flags, g["v"]) if phi is None else None)
            if O is not None:
                if (pc, acc, fac, fix) not in cpu_cache:
                    cpu_cache[(pc, acc, fac, fix)] = cpu_solve(O, flags, vel, phi, pc, acc, fac, fix)
                c = cpu_cache[(pc, acc, fac, fix)]
                row.update(cpu_kind=O.kind, cpu_cores=cores, cpu_iterations=c["iterations"], cpu_ms=c["ms"], speedup_device=c["ms"] / g["ms"], speedup_plugin=c["ms"] / g["host_ms"],
                           p_rel_l2=rel_l2(g["p"], c["p"]), p_rel_l2_demeaned=rel_l2(demean(g["p"], flags), demean(c["p"], flags)), vel_rel_l2=rel_l2(g["v"], c["v"]),
                           cpu_max_div_after=scenes.max_divergence(flags, c["v"]) if phi is None else None)
            rows.append(row)
            print(json.dumps(row), flush=True)
    out = os.path.join(ROOT, "profiles", "%s_configs" % a.tag)
    json.dump(rows, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write("| config | build | preconditioner | GPU iters | GPU device ms (warm / cold) | GPU plugin ms (host buffers) | CPU ref iters | CPU ref ms (%d cores) | speed-up device / plugin | pressure rel-L2 vs ref | vel rel-L2 | max div after (GPU / ref) |\n" % cores)
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write("| %s | %s | %s | %d | %.2f / %.2f | %.2f | %s | %s | %s | %s | %s | %s |\n" % (
                r["config"], r["prec"], r["pc"], r["gpu_iterations"], r["gpu_ms"], r["gpu_ms_cold"], r["gpu_host_ms"],
                r.get("cpu_iterations", "-"), ("%.0f" % r["cpu_ms"]) if "cpu_ms" in r else "-",
                ("%.0fx / %.0fx" % (r["speedup_device"], r["speedup_plugin"])) if "cpu_ms" in r else "-",
                ("%.1e" % r["p_rel_l2"]) if "p_rel_l2" in r else "-", ("%.1e" % r["vel_rel_l2"]) if "vel_rel_l2" in r else "-",
                ("%.1e / %s" % (r["max_div_after"], ("%.1e" % r["cpu_max_div_after"]) if r.get("cpu_max_div_after") is not None else "-")) if r["max_div_after"] is not None else "-"))
    print("wrote", out + ".md")


if __name__ == "__main__":
    main()

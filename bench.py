#!/usr/bin/env python
"""bench.py -- the pressure-projection hot path on B200 (see DESIGN.md "Measurement").

A step = ONE solvePressure (computePressureRhs + MakeLaplaceMatrix + GridCg to convergence + correctVelocity,
plugin/pressure.cpp:480-521) over one synthetic smoke-plume grid:
  N=1 : BASELINE.json configs[3] "synthetic 512^3 smoke plume single-GPU pressure projection", float build,
        preconditioner PcNone by default (the memory-bound matvec/axpy path the north_star roofline is about),
        cgAccuracy 1e-4, cgMaxIterFac 99.
  N>1 : the same 512^3 cells PER GPU, z-slab sharded (global grid 512x512x1024 / 512x1024x1024 / 1024^3 for N = 2 / 4 / 8;
        N=8 is BASELINE.json's 1024^3 config), one-plane halo exchange + per-iteration scalar all-gather over NCCL -> "scaling": "weak".
metric  = cells x CG iterations / second over the whole job ("Gcell-iter/s"); a 512^3 PcNone iteration at the HBM roofline
          (64 B/cell, MEASURED_PEAKS hbm_gbs) is the ceiling.  cg_iter_per_s and solve_ms are reported beside it.
value   = inputs resident in HBM when the timed region starts; e2e = same metric through mp_solve_pressure_host with
          pinned HOST buffers (H2D of flags+vel and D2H of vel+pressure inside the timed region).
--impl reference : the reference's own CPU code (oracle/_ref, the unmodified reference compiled here; else the C port)
          on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PC_NAMES = {0: "PcNone", 1: "PcMIC", 2: "PcMGDynamic", 3: "PcMGStatic"}
# algorithmic bytes per cell (SURVEY 8d / DESIGN.md): matvec 4+6w, axpy2+norms 6w, update 3w
def bytes_per_cell(w):
    return {"matvec": 4 + 6 * w, "axpy": 6 * w, "update": 3 * w, "iter_none": 4 + 15 * w}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """All host cores for the reference's OpenMP build (torchrun exports OMP_NUM_THREADS=1 to its workers, which must not
    silently serialise the CPU arm); MP_CPU_THREADS overrides.  Must run before the oracle library is loaded."""
    n = int(os.environ.get("MP_CPU_THREADS", 0)) or len(os.sched_getaffinity(0)) or os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


def make_scene(res, prec, n_slabs=1):
    from mantaflow_b200 import scenes
    sx = sy = res
    sz = res * n_slabs
    return scenes.smoke_plume((sx, sy, sz), prec)


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_step(O, flags, vel, pc, max_iter, accuracy):
    """One bounded pass of the hot path on the CPU: rhs + matrix + GridCg capped at max_iter + correctVelocity.
    GridCg is driven directly because solvePressure(PcNone) asserts in the reference.  Returns (seconds, iterations)."""
    t0 = time.perf_counter()
    rhs, _, _ = O.compute_rhs(flags, vel)
    A = O.make_matrix(flags)
    x, it, rn = O.cg_solve(flags, rhs, *A, pc={0: 0, 1: 1, 2: 2, 3: 2}[pc], accuracy=accuracy, maxIter=max_iter)
    O.correct_velocity(flags, vel, x)
    return time.perf_counter() - t0, it


def cpu_sample(O, flags, vel, pc, cap_lo, cap_hi):
    """The reference's cost split into set-up and per-iteration parts from two capped passes over the same input:
    a pass costs setup + iterations x per_iter, so per_iter = (t_hi - t_lo) / (it_hi - it_lo) and setup = t_lo - it_lo x per_iter
    (rhs, matrix, GridCg::doInit, preconditioner set-up, correctVelocity).  A full solve takes the same number of iterations as ours
    (same algorithm, same arithmetic: 1598 at 512^3 PcNone, tests/golden/fullsize_divergence.json), so the per-iteration rate is the
    honest CPU figure for the Gcell-iter/s metric; the capped-sample rate (which amortises the set-up over a handful of iterations) is
    kept beside it."""
    t_lo, it_lo = cpu_step(O, flags, vel.copy(), pc, cap_lo, 1e-4)
    t_hi, it_hi = cpu_step(O, flags, vel.copy(), pc, cap_hi, 1e-4)
    if it_hi > it_lo:
        per_iter = max((t_hi - t_lo) / (it_hi - it_lo), 1e-9)
    else:                      # converged below the low cap (multigrid): nothing to separate
        per_iter = t_hi / max(it_hi, 1)
    setup = max(t_lo - it_lo * per_iter, 0.0)
    return {"per_iter_s": per_iter, "setup_s": setup, "t_hi_s": t_hi, "it_hi": it_hi, "t_lo_s": t_lo, "it_lo": it_lo}


def cpu_fields(smp, cells, cores, kind, what):
    """the cpu_baseline object (and the reference line's extra keys) from one cpu_sample"""
    per_iter_rate = cells / smp["per_iter_s"] / 1e9
    capped_rate = cells * smp["it_hi"] / smp["t_hi_s"] / 1e9
    return {"value": per_iter_rate, "unit": "Gcell-iter/s", "cores": cores, "kind": kind,
            "value_is": "cells / per-iteration time (set-up excluded: the figure a full solve of ~3.1 x res iterations converges to)",
            "ref_ms_per_iter": 1e3 * smp["per_iter_s"], "ref_setup_ms": 1e3 * smp["setup_s"],
            "value_capped_sample": capped_rate,
            "sample": "%s: two passes of rhs+matrix+GridCg+correctVelocity capped at %d and %d iterations (%.1f s + %.1f s)" % (
                what, smp["it_lo"], smp["it_hi"], smp["t_lo_s"], smp["t_hi_s"])}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle_api import Oracle, available
    cores = host_threads()
    kind = "reference" if available("reference", args.prec) else "port"
    O = Oracle(kind, args.prec)
    flags, vel = make_scene(args.res, args.prec)
    cells = flags.size
    cap_hi = args.cpu_iters
    cap_lo = max(2, cap_hi // 4)
    smps = []
    for s in range(args.warmup + args.steps):
        # a warm-up pass is a short one (page faults, thread pool); every timed step is the bounded two-pass sample
        if s < args.warmup:
            cpu_step(O, flags, vel.copy(), args.pc, 2, 1e-4)
        else:
            smps.append(cpu_sample(O, flags, vel, args.pc, cap_lo, cap_hi))
    smp = {k: float(np.mean([q[k] for q in smps])) for k in smps[0]}
    smp["it_hi"], smp["it_lo"] = int(smps[0]["it_hi"]), int(smps[0]["it_lo"])
    what = "one %d^3 block of the workload (the per-GPU share%s), %s" % (
        args.res, "" if args.gpus == 1 else "; the global grid of the N-GPU arm does not fit the host, so for N > 1 this is a machine-vs-machine figure, not the same problem", PC_NAMES[args.pc])
    cb = cpu_fields(smp, cells, cores, kind, what)
    value = cb["value"]
    line = {"impl": "reference", "metric": "pressure-solve CG throughput (cells x iterations / s)", "value": value, "unit": "Gcell-iter/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (smp["t_lo_s"] + smp["t_hi_s"]),
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.prec == 4 else "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cg_iter_per_s": 1.0 / smp["per_iter_s"],
            "ref_ms_per_iter": cb["ref_ms_per_iter"], "ref_setup_ms": cb["ref_setup_ms"], "value_capped_sample": cb["value_capped_sample"],
            "same_problem_as_repo_arm": args.gpus == 1,
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "Gcell-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


from mantaflow_b200.sharded import global_grid  # noqa: E402


def workload_config(args, n):
    gx, gy, gz = global_grid(args.res, n)
    return {"workload": "synthetic smoke plume with sphere obstacle, %d^3 cells per GPU (global %dx%dx%d), solvePressure %s, cgAccuracy 1e-4, cgMaxIterFac 99, %s build"
            % (args.res, gx, gy, gz, PC_NAMES[args.pc], "float" if args.prec == 4 else "double"),
            "preconditioner": PC_NAMES[args.pc], "grid": [gx, gy, gz], "sharding": "z-slabs x%d (%d planes per GPU)" % (n, gz // n),
            "l2": "inputs larger than L2 (one Real grid = %d MiB per GPU, 9 streamed per iteration)" % (args.res ** 3 * args.prec >> 20)}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import mantaflow_b200 as mf
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run --nproc-per-node N)" % (args.gpus, world))
    if not torch.cuda.is_available() or mf.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prec, res, pc = args.prec, args.res, args.pc
    real = np.float32 if prec == 4 else np.float64

    if world > 1:
        from mantaflow_b200 import sharded
        runner = sharded.ShardedBench(args, rank, world, local, dist)
    else:
        runner = SingleBench(args, local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ----
    for _ in range(args.warmup):
        runner.step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = runner.launches()
    barrier()
    t0 = time.perf_counter()
    ms_dev, iters = [], 0
    for _ in range(args.steps):
        info = runner.step_resident()
        ms_dev.append(info["msTotal"]); iters = info["iterations"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    launches = runner.launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = float(np.mean(ms_dev))          # CUDA-event time of the step on the launching stream
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = t.tolist()
    prof = runner.last_info

    # ---- end-to-end arm (host buffers through the C-ABI plugin entry point) ----
    e2e_ms, bi, bo = runner.run_e2e(args.warmup, args.steps, barrier)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item()

    extra = None
    if world > 1 and not args.no_configs and args.pc == 0:
        extra = runner.extra_configs()          # collective
        t = torch.tensor([extra["pcmgstatic"]["solve_ms"], extra["pcmgstatic"]["solve_ms_cold"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        extra["pcmgstatic"]["solve_ms"], extra["pcmgstatic"]["solve_ms_cold"] = t.tolist()
    if rank == 0:
        cells = res ** 3 * world          # = product of global_grid(res, world)
        peak, peak_src = load_peaks()
        bpc = bytes_per_cell(prec)
        mvk = prof.get("matvecKernel", 1)
        mv_name = {0: "k_matvec_dot", 1: "k_matvec_zmarch", 2: "k_matvec_zmarch_masked", 3: "k_matvec_fused", 4: "k_matvec_fused_tma"}[mvk]
        if mvk == 2:
            bpc["matvec"] = 4 + 3 * prec          # coupling-mask fast path: cmask 4 + A0 w + s w + t w
        ax_name, up_name = "k_axpy2_norm", "k_update_search"
        if mvk == 3:                              # fused PcNone iteration: R r, s_old, x, cmask, A0; W s, t, x  |  R r, t; W r
            bpc["matvec"] = 4 + 7 * prec
            bpc["axpy"] = 3 * prec
            bpc["update"] = 0
            ax_name, up_name = "k_axpy1_norm", None
        if mvk == 4:                              # the same iteration, staged by TMA, matrix as 2 bytes per cell: R r, s_old, x, mask16; W s, t, x
            bpc["matvec"] = 2 + 6 * prec
            bpc["axpy"] = 3 * prec
            bpc["update"] = 0
            ax_name, up_name = "k_axpy1_norm", None
        mv_ms = prof.get("msMatvecAvg", 0.0)
        cells_gpu = res ** 3
        kt = {"matvec": mv_ms, "axpy": prof.get("msAxpyAvg") or 0.0, "update": prof.get("msUpdateAvg") or 0.0}
        kn = {"matvec": mv_name, "axpy": ax_name, "update": up_name}
        if up_name is None:                       # the search-vector update rides on the matvec: two kernels per iteration
            del kt["update"], kn["update"]
        gbs = {k: (bpc[k] * cells_gpu / (kt[k] * 1e-3) / 1e9) if kt[k] > 0 else None for k in kt}
        dom = max(kt, key=lambda k: kt[k])        # the kernel with the largest share of the timed solve
        achieved = gbs[dom]
        line = {"metric": "pressure-solve CG throughput (cells x iterations / s)", "value": cells * iters / (dev_ms * 1e-3) / 1e9, "unit": "Gcell-iter/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if prec == 4 else "f64", "data": "synthetic", "config": workload_config(args, world),
                "solve_ms": dev_ms, "wall_ms_per_step": wall_ms, "iterations": iters, "cg_iter_per_s": iters / (dev_ms * 1e-3),
                "stage_ms": {k: prof.get(k) for k in ("msRhs", "msMatrix", "msSolve", "msCorrect")},
                "kernel_ms": dict([(mv_name, mv_ms), (ax_name, prof.get("msAxpyAvg"))] + ([(up_name, prof.get("msUpdateAvg"))] if up_name else []) +
                                  [("precond", prof.get("msPrecondAvg")), ("samples", prof.get("profSamples"))]),
                "kernel_gbs": {kn[k]: gbs[k] for k in kt},
                "kernel_frac_of_peak": {kn[k]: (gbs[k] / peak if gbs[k] else None) for k in kt},
                "kernel_bytes_per_cell": {kn[k]: bpc[k] for k in kt},
                "roofline": {"bound": "hbm", "kernel": kn[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": (achieved / peak) if achieved else None, "traffic": runner.ncu_traffic(kn[dom]),
                             "traffic_source": "ncu --set full capture of this kernel at this size (profiles/traffic.json), not this run", "peak_source": peak_src,
                             "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None,
                             "algorithmic_bytes_per_cell": bpc[dom]},
                "exchange": runner.exchange(),
                "e2e": {"value": cells * iters / (e2e_ms * 1e-3) / 1e9, "unit": "Gcell-iter/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bo},
                "gpu_launches": int(launches), "clocks": clocks}
        if world == 1 and not args.no_configs and args.pc == 0 and args.prec == 4:
            runner.release()
            line["configs"] = extra_configs(args, peak)
        if extra is not None:
            line["configs"] = extra
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


class SingleBench:
    def __init__(self, args, device):
        import ctypes as C
        import mantaflow_b200 as mf
        self.mf, self.args = mf, args
        res, prec = args.res, args.prec
        flags, vel = make_scene(res, prec)
        self.s = mf.Solver(gridSize=(res, res, res), dim=3, prec=prec, device=device)
        self.s.setProfiling(16)
        self.F, self.V0, self.V, self.P = mf.FlagGrid(self.s, flags), mf.MACGrid(self.s, vel), mf.MACGrid(self.s), mf.RealGrid(self.s)
        self.F.dev(); self.V0.dev()
        self.kw = dict(cgAccuracy=1e-4, cgMaxIterFac=99, preconditioner=args.pc, zeroPressureFixing=(args.pc >= 2))
        self.last_info = {}
        # pinned host staging for the e2e arm
        lib = self.s.lib
        self._pinned = []

        def pinned_like(a):
            p = C.c_void_p()
            mf._lib.check(lib.mp_host_alloc(C.byref(p), C.c_ulonglong(a.nbytes)))
            self._pinned.append(p)
            buf = (C.c_char * a.nbytes).from_address(p.value)
            out = np.frombuffer(buf, dtype=a.dtype).reshape(a.shape)
            out[...] = a
            return out
        self.h_flags, self.h_vel0 = pinned_like(flags), pinned_like(vel)
        self.h_vel, self.h_p = pinned_like(vel), pinned_like(np.zeros(flags.shape, vel.dtype))

    def launches(self):
        return self.s.kernelLaunches()

    def step_resident(self):
        mf = self.mf
        # restore the un-projected velocity on the device (D2D, part of the step) and project it
        mf._lib.check(self.s.lib.mp_grid_copy_from(self.V.dev(), self.V0.dev()))
        mf.solvePressure(vel=self.V, pressure=self.P, flags=self.F, **self.kw)
        self.last_info = mf.lastSolveInfo()
        return self.last_info

    def run_e2e(self, warmup, steps, barrier):
        mf = self.mf
        ts = []
        for i in range(warmup + steps):
            self.h_vel[...] = self.h_vel0          # host-side reset, outside the timed region
            barrier()
            t0 = time.perf_counter()
            mf.solvePressureHost(self.s, self.h_vel, self.h_p, self.h_flags, **self.kw)   # synchronous: returns after the D2H
            float(self.h_p.ravel()[self.h_p.size // 2])
            barrier()
            if i >= warmup:
                ts.append(1e3 * (time.perf_counter() - t0))
        bi = self.h_flags.nbytes + self.h_vel.nbytes
        bo = self.h_vel.nbytes + self.h_p.nbytes
        return float(np.mean(ts)), bi, bo

    def exchange(self):
        return "none (1 GPU)"

    def release(self):
        """give the device memory of the timed workload back before the other configurations run"""
        for g in (self.F, self.V0, self.V, self.P):
            g.close()
        self.s.close()

    def ncu_traffic(self, kernel):
        p = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(p):
            try:
                return json.load(open(p)).get("%s_f%d_%d" % (kernel, self.args.prec * 8, self.args.res))
            except Exception:
                return None
        return None


def cpu_baseline(args):
    from oracle.oracle_api import Oracle, available
    cores = host_threads()
    kind = "reference" if available("reference", args.prec) else "port"
    O = Oracle(kind, args.prec)
    res = args.cpu_res or args.res
    flags, vel = make_scene(res, args.prec)
    cpu_step(O, flags, vel.copy(), args.pc, 2, 1e-4)       # warm-up (page faults, thread pool)
    smp = cpu_sample(O, flags, vel, args.pc, max(2, args.cpu_iters // 4), args.cpu_iters)
    return cpu_fields(smp, flags.size, cores, kind, "%d^3 %s" % (res, PC_NAMES[args.pc]))


# ------------------------------------------------------------------------------------------------ the other configurations (N = 1)
def extra_configs(args, peak):
    """BASELINE.json's other configurations and preconditioners, each as device-resident solvePressure calls on this GPU: PcMIC / PcMGStatic
    (cold = first solve incl. GridMg::setA, warm = hierarchy reused) / PcMGDynamic at the bench size, the double build, config 1
    (64x96x64 smoke, PcMIC) and config 2 (88x83x33 liquid + phi, PcMIC / PcMGDynamic).  Not part of `value`; measured after the timed region."""
    import mantaflow_b200 as mf
    from mantaflow_b200 import scenes
    res = args.res
    out = {}

    def run(name, flags, vel, phi, prec, pc, acc, fac, fix, reps, mic_rb=False):
        sz, sy, sx = flags.shape
        s = mf.Solver(gridSize=(sx, sy, sz), dim=3, prec=prec)
        s.setProfiling(4 if pc else 16)
        if mic_rb:                   # PcMIC reformulated: MIC(0) of the block red-black ordering, tile chosen from the grid (mp_set_mic_ordering)
            s.setMicOrdering(1)
        F, V0, V, P = mf.FlagGrid(s, flags), mf.MACGrid(s, vel), mf.MACGrid(s), mf.RealGrid(s)
        PH = mf.RealGrid(s, phi) if phi is not None else None
        F.dev(); V0.dev()
        ms, info = [], None
        first_ms = None
        for r in range(reps + (1 if pc == 3 else 0)):
            mf._lib.check(s.lib.mp_grid_copy_from(V.dev(), V0.dev()))
            mf.solvePressure(vel=V, pressure=P, flags=F, phi=PH, cgAccuracy=acc, cgMaxIterFac=fac, preconditioner=pc, zeroPressureFixing=fix)
            info = mf.lastSolveInfo()
            if pc == 3 and r == 0:
                # the very first solve of a solver also pays cudaMalloc of ~12 grids and of the hierarchy (hundreds of ms after another
                # configuration's memory was freed): reported on its own; "cold" below = the hierarchy released and set up again (GridMg::setA
                # included, allocations reused), which is what a scene pays when its flags change
                first_ms = info["msTotal"]
                mf.releaseMG(s)
                continue
            ms.append(info["msTotal"])
        cells = flags.size
        w = prec
        row = {"grid": [sx, sy, sz], "dtype": "f32" if prec == 4 else "f64", "preconditioner": PC_NAMES[pc], "cgAccuracy": acc,
               "iterations": info["iterations"], "solve_ms": min(ms[1:]) if len(ms) > 1 else ms[0], "solve_ms_cold": ms[0],
               "ms_per_iteration": (min(ms[1:]) if len(ms) > 1 else ms[0]) / max(info["iterations"], 1), "exchange": "none (1 GPU)",
               "kernel_ms": {"matvec": info["msMatvecAvg"], "axpy": info["msAxpyAvg"], "precond": info["msPrecondAvg"], "update": info["msUpdateAvg"]},
               "matvecKernel": info["matvecKernel"], "mgLevels": info["mgLevels"]}
        if first_ms is not None:
            row["solve_ms_first_incl_allocation"] = first_ms
        # dominant kernel of the iteration against its algorithmic bytes: MIC apply 12+12w, V-cycle ~(4+28w)+(3+55w)/7, matvec per kernel
        mvb = {0: 4 + 6 * w, 1: 4 + 6 * w, 2: 4 + 3 * w, 3: 4 + 7 * w, 4: 2 + 6 * w}[info["matvecKernel"]]
        cand = {"matvec": (info["msMatvecAvg"], mvb)}
        if pc == 1 and mic_rb:
            mode, ty, tz = s.micOrdering()
            row["preconditioner"] = "PcMIC, block red-black ordering (reformulated; %dx%d rows per tile)" % (ty, tz) if mode else "PcMIC (ordering request fell back to lexicographic)"
            cand["precond (MIC apply, block red-black)"] = (info["msPrecondAvg"], 2 + 6 * w)       # 4 Real read + 2 written + 2 mask bytes per cell
        elif pc == 1:
            cand["precond (MIC apply)"] = (info["msPrecondAvg"], 12 + 12 * w)
        if pc >= 2:
            cand["precond (GridMg V-cycle)"] = (info["msPrecondAvg"], (4 + 28 * w) + (3 + 55 * w) / 7.0)
        dom = max(cand, key=lambda k: cand[k][0] or 0.0)
        t, bpc = cand[dom]
        if t and t > 0:
            gbs = bpc * cells / (t * 1e-3) / 1e9
            row["dominant_kernel"] = {"name": dom, "ms": t, "bytes_per_cell": bpc, "achieved_gbs": gbs, "frac_of_measured_peak": gbs / peak}
        if pc == 3:
            mf.releaseMG(s)
        s.close()
        out[name] = row

    # CUDA loads kernels lazily: the first PcMIC / PcMG* solve of the process would pay ~1 s of module loading that has nothing to do with the
    # hierarchy set-up the "cold" figure is about -- run every preconditioner once on a small grid in both precisions first
    for wprec in (4, 8):
        wflags, wvel = make_scene(32, wprec)
        for wpc in (0, 1, 2, 3):
            run("_warm", wflags, wvel, None, wprec, wpc, 1e-4, 99, wpc >= 2, 1)
    out.pop("_warm", None)
    f32 = make_scene(res, 4)
    run("pcmic_f32", f32[0], f32[1], None, 4, 1, 1e-4, 99, False, 2)
    run("pcmic_blockrb_f32", f32[0], f32[1], None, 4, 1, 1e-4, 99, False, 2, mic_rb=True)
    run("pcmgstatic_f32", f32[0], f32[1], None, 4, 3, 1e-4, 99, True, 3)
    run("pcmgdynamic_f32", f32[0], f32[1], None, 4, 2, 1e-4, 99, True, 2)
    del f32
    f64 = make_scene(res, 8)
    run("pcnone_f64", f64[0], f64[1], None, 8, 0, 1e-4, 99, False, 2)
    run("pcmgstatic_f64", f64[0], f64[1], None, 8, 3, 1e-4, 99, True, 3)
    del f64
    c1 = scenes.smoke_plume((64, 96, 64), 4, obstacle=False)
    run("cfg1_64x96x64_pcmic", c1[0], c1[1], None, 4, 1, 1e-3, 1.5, False, 3)
    run("cfg1_64x96x64_pcmic_blockrb", c1[0], c1[1], None, 4, 1, 1e-3, 1.5, False, 3, mic_rb=True)
    c2 = scenes.liquid_basin((88, 83, 33), 4)
    run("cfg2_88x83x33_phi_pcmic", c2[0], c2[1], c2[2], 4, 1, 1e-3, 1.5, False, 3)
    run("cfg2_88x83x33_phi_pcmgdynamic", c2[0], c2[1], c2[2], 4, 2, 1e-3, 1.5, False, 3)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--prec", type=int, default=4, choices=[4, 8])
    ap.add_argument("--pc", type=int, default=0, choices=[0, 1, 2, 3], help="0 PcNone 1 PcMIC 2 PcMGDynamic 3 PcMGStatic")
    ap.add_argument("--cpu-res", type=int, default=0, help="grid of the bounded cpu_baseline sample in our arm (0 = --res)")
    ap.add_argument("--cpu-iters", type=int, default=24, help="GridCg iteration cap of a CPU step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (other preconditioners / precisions / BASELINE configs 1-2)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

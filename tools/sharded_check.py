"""Parity of the z-slab sharded solve against the single-process oracle.  Run under torchrun:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import mantaflow_b200 as mf  # noqa: E402
from mantaflow_b200 import scenes, sharded  # noqa: E402
from oracle.oracle_api import Oracle  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True


def gather(arr):
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(sharded.owned(arr)))
    return sharded.assemble(parts)


def rel(a, b):
    return float(np.linalg.norm((a - b).astype(np.float64).ravel()) / max(np.linalg.norm(b.astype(np.float64).ravel()), 1e-300))


cases = [("smoke", 4, dict()), ("smoke", 8, dict()), ("smoke_pin", 4, dict(zeroPressureFixing=True)), ("liquid", 4, dict()), ("smoke_compat_l2", 4, dict(enforceCompatibility=True, useL2Norm=True))]
for name, prec, extra in cases:
    shape = (44, 36, 41)      # ragged in z: uneven slabs
    phi = None
    if name == "liquid":
        flags, vel, phi = scenes.liquid_basin(shape, prec)
    else:
        flags, vel = scenes.smoke_plume(shape, prec, random_vel=True)
    acc = 1e-5 if prec == 4 else 1e-11
    uid = sharded.exchange_unique_id(dist, rank)      # one NCCL communicator per solver
    s = sharded.ShardedSolver(shape, rank, world, uid, prec=prec, device=local)
    F = mf.FlagGrid(s, sharded.local_slab(flags, rank, world)); V = mf.MACGrid(s, sharded.local_slab(vel, rank, world)); P = mf.RealGrid(s)
    PH = mf.RealGrid(s, sharded.local_slab(phi, rank, world)) if phi is not None else None
    mf.solvePressure(vel=V, pressure=P, flags=F, phi=PH, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=0, **extra)
    info = mf.lastSolveInfo()
    p_all, v_all = gather(P.numpy()), gather(V.numpy())
    if rank == 0:
        O = Oracle("port", prec)
        v_o = vel.copy()
        p_o, it_o, rn_o = O.solve_pressure(flags, v_o, phi=phi, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=0, **extra)
        e_p, e_v = rel(p_all, p_o), rel(v_all, v_o)
        good = abs(info["iterations"] - it_o) <= 1 and e_p <= (1e-4 if prec == 4 else 1e-10) and e_v <= (1e-4 if prec == 4 else 1e-10)
        print("sharded_check %-16s f%d world=%d iterations %d (oracle %d) relL2 p %.2e vel %.2e fixed %d %s" % (name, prec * 8, world, info["iterations"], it_o, e_p, e_v, info["fixedCell"], "OK" if good else "FAIL"), flush=True)
        ok = ok and good
    # preconditioned solves: PcMIC is block-Jacobi over the slabs (different operator than the global preconditioner, same solution within
    # the solver tolerance; iteration counts are reported, not asserted); PcMG is the global hierarchy here (sx = 44: rows 16-byte aligned
    # in float) -- asserted in the section below
    if name in ("smoke_pin", "liquid"):
        for pc in (1, 2):
            fixp = pc >= 2 or bool(extra.get("zeroPressureFixing"))
            V.copyFromArray(sharded.local_slab(vel, rank, world))
            mf.solvePressure(vel=V, pressure=P, flags=F, phi=PH, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=fixp)
            info = mf.lastSolveInfo()
            p_all, v_all = gather(P.numpy()), gather(V.numpy())
            if rank == 0:
                v_o = vel.copy()
                p_o, it_o, rn_o = O.solve_pressure(flags, v_o, phi=phi, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=fixp)
                e_p, e_v = rel(p_all, p_o), rel(v_all, v_o)
                good = info["resNorm"] < acc and e_v <= 2e-3 and e_p <= 2e-3
                print("sharded_check %-16s f%d world=%d %s on slabs: iterations %d (single-process oracle %d) resNorm %.2e relL2 p %.2e vel %.2e %s"
                      % (name, prec * 8, world, {1: "PcMIC", 2: "PcMGDynamic"}[pc], info["iterations"], it_o, info["resNorm"], e_p, e_v, "OK" if good else "FAIL"), flush=True)
                ok = ok and good
    s.close()

# the global GridMg on z-slabs (rows aligned to 16 bytes): the hierarchy of the GLOBAL grid on every rank, level-0 work of the V-cycle sharded
# with halo exchanges -- the same V-cycle as a single-GPU solve, so the iteration count is the oracle's and the float result its bits
# shapes are (sx, sy, sz)
for name, prec, shape in [("smoke_pin", 4, (48, 36, 45)), ("liquid", 4, (64, 40, 44)), ("smoke_pin", 8, (48, 36, 45)), ("liquid", 8, (32, 30, 38))]:
    phi = None
    if name == "liquid":
        flags, vel, phi = scenes.liquid_basin(shape, prec)
    else:
        flags, vel = scenes.smoke_plume(shape, prec, random_vel=True)
    acc = 1e-5 if prec == 4 else 1e-11
    uid = sharded.exchange_unique_id(dist, rank)
    s = sharded.ShardedSolver(shape, rank, world, uid, prec=prec, device=local)
    F = mf.FlagGrid(s, sharded.local_slab(flags, rank, world)); V = mf.MACGrid(s, sharded.local_slab(vel, rank, world)); P = mf.RealGrid(s)
    PH = mf.RealGrid(s, sharded.local_slab(phi, rank, world)) if phi is not None else None
    if rank == 0:
        O = Oracle("port", prec)
    for step, pc in enumerate((3, 3, 2, 2)):          # PcMGStatic twice (second solve reuses the hierarchy), then PcMGDynamic twice:
        # the last solve shards every level that leaves each rank >= 4 planes (production shards levels of >= 4 M vertices only)
        os.environ["MP_MG_SHARD_MIN_N"] = "1000" if step == 3 else str(4 << 20)
        V.copyFromArray(sharded.local_slab(vel, rank, world))
        mf.solvePressure(vel=V, pressure=P, flags=F, phi=PH, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=True)
        info = mf.lastSolveInfo()
        p_all, v_all = gather(P.numpy()), gather(V.numpy())
        if rank == 0:
            v_o = vel.copy()
            p_o, it_o, rn_o = O.solve_pressure(flags, v_o, phi=phi, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=True)
            e_p, e_v = rel(p_all, p_o), rel(v_all, v_o)
            good = abs(info["iterations"] - it_o) <= 1 and e_p <= (1e-4 if prec == 4 else 1e-10) and e_v <= (1e-4 if prec == 4 else 1e-10)
            print("sharded_check %-16s f%d world=%d %s GLOBAL GridMg on slabs%s: iterations %d (single-process oracle %d) levels %d relL2 p %.2e vel %.2e %s"
                  % (name, prec * 8, world, {2: "PcMGDynamic", 3: "PcMGStatic"}[pc], " (coarse levels sharded too)" if step == 3 else "", info["iterations"], it_o, info["mgLevels"], e_p, e_v, "OK" if good else "FAIL"), flush=True)
            ok = ok and good
    mf.releaseMG(s)
    s.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)

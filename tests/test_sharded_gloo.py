"""CPU (not gpu), world_size 2 over gloo: the N>1 host logic -- slab rule, scatter with ghost planes, one-plane halo
exchange and all-gathered partial reductions combined in rank order -- drives a slab-decomposed CG (numpy stand-in for the
kernels) to the same iterates as the unsharded oracle.  This is the algorithm csrc/mp_dist.cu + mp_cg.cu implement."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _apply(flags, s, A0, Ai, Aj, Ak):
    """ApplyMatrix (conjugategrad.h:118-133) on a slab with ghost planes; result valid on owned planes [1:-1]"""
    fl = (flags & 1) != 0
    t = s.copy()
    c = (slice(1, -1), slice(1, -1), slice(1, -1))
    v = (s[c] * A0[c] + s[1:-1, 1:-1, :-2] * Ai[1:-1, 1:-1, :-2] + s[1:-1, 1:-1, 2:] * Ai[c]
         + s[1:-1, :-2, 1:-1] * Aj[1:-1, :-2, 1:-1] + s[1:-1, 2:, 1:-1] * Aj[c]
         + s[:-2, 1:-1, 1:-1] * Ak[:-2, 1:-1, 1:-1] + s[2:, 1:-1, 1:-1] * Ak[c])
    t[c] = np.where(fl[c], v, s[c])
    return t


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from mantaflow_b200 import scenes, sharded
    from oracle.oracle_api import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prec = 8
    flags, vel = scenes.smoke_plume((14, 12, 17), prec, random_vel=True)      # 17 planes: uneven slabs
    O = Oracle("port", prec)
    rhs, _, _ = O.compute_rhs(flags, vel)
    A = O.make_matrix(flags)
    x_ref, it_ref, rn_ref = O.cg_solve(flags, rhs, *A, pc=0, accuracy=1e-10, maxIter=500)

    def halo(a):       # one-plane exchange with <= 2 neighbours (mp_dist_halo)
        reqs = []
        lo, hi = torch.from_numpy(a[0]), torch.from_numpy(a[-1])
        if rank > 0:
            reqs += [dist.isend(torch.from_numpy(np.ascontiguousarray(a[1])), rank - 1), dist.irecv(lo, rank - 1)]
        if rank < world - 1:
            reqs += [dist.isend(torch.from_numpy(np.ascontiguousarray(a[-2])), rank + 1), dist.irecv(hi, rank + 1)]
        for r in reqs:
            r.wait()

    def gathered(vals):    # all-gather of the partials, combined in rank order (mp_dist_allgather + k_cg_combine)
        out = [torch.zeros(len(vals), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(out, torch.tensor(vals, dtype=torch.float64))
        return np.stack([o.numpy() for o in out])

    L = lambda a: sharded.local_slab(a, rank, world)
    f, b = L(flags), L(rhs)
    A0, Ai, Aj, Ak = (L(a) for a in A)
    x, r = np.zeros_like(b), b.copy()
    s = b.copy(); halo(s)
    own = lambda a: a[1:-1]
    sigma = gathered([float(np.sum(own(r) * own(r)))])[:, 0].sum()
    its = 0
    for it in range(500):
        its += 1
        t = _apply(f, s, A0, Ai, Aj, Ak)
        dp = gathered([float(np.sum(own(t) * own(s)))])[:, 0].sum()
        alpha = sigma / dp if abs(dp) > 0 else 0.0
        own(x)[...] += alpha * own(s); own(r)[...] -= alpha * own(t)
        g = gathered([float(np.abs(own(r)).max()), float(np.sum(own(r) * own(r)))])
        res, sig_new = g[:, 0].max(), g[:, 1].sum()
        if res < 1e-10:
            break
        beta = sig_new / sigma; sigma = sig_new
        own(s)[...] = own(r) + beta * own(s); halo(s)
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(own(x)))
    if rank == 0:
        x_all = sharded.assemble(parts)
        err = float(np.linalg.norm(x_all - x_ref) / np.linalg.norm(x_ref))
        q.put((its, it_ref, err))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_slab_cg_over_gloo_matches_unsharded_oracle():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    its, it_ref, err = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert abs(its - it_ref) <= 1 and err < 1e-9, (its, it_ref, err)

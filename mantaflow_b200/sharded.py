"""z-slab sharding of one pressure solve over the GPUs of one box (one process per GPU).

The reference has no distributed mode.  torch.distributed is only the plumbing here: it carries the
128-byte NCCL unique id to the ranks (and the max-over-ranks timing in bench.py); halo exchange and the scalar
all-gathers of the CG loop run inside libmantapress over NCCL/NVLink (csrc/mp_dist.cu).

Host-side layout rule (pure numpy, testable without a GPU): rank r owns the global planes [k0,k1) and stores them
with one ghost plane on each side; ghost planes outside the domain are zero."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .grid import Solver


def global_grid(res, n):
    """res^3 cells per GPU, as cubic as possible: N=1 res^3, N=2 res x res x 2res, N=4 res x 2res x 2res, N=8 (2res)^3
    (res=512, N=8 is the 1024^3 configuration); z is the sharded axis."""
    dims = [res, res, res]
    f, axis = n, 2
    while f > 1:
        dims[axis] *= 2
        f //= 2
        axis = (axis - 1) % 3
    return tuple(dims)


def slab(sz, rank, world):
    """global planes [k0,k1) owned by `rank` -- the same rule as mp_dist_slab"""
    base, rem = divmod(sz, world)
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


def local_slab(global_arr, rank, world):
    """owned planes plus one ghost plane on each side, zero-filled outside the domain; axis 0 is z"""
    sz = global_arr.shape[0]
    k0, k1 = slab(sz, rank, world)
    out = np.zeros((k1 - k0 + 2,) + global_arr.shape[1:], global_arr.dtype)
    lo, hi = max(k0 - 1, 0), min(k1 + 1, sz)
    out[lo - (k0 - 1):hi - (k0 - 1)] = global_arr[lo:hi]
    return out


def owned(local_arr):
    """strip the two ghost planes"""
    return local_arr[1:-1]


def assemble(owned_parts):
    """global array from the ranks' owned planes (rank order)"""
    return np.concatenate(list(owned_parts), axis=0)


def unique_id():
    buf = (C.c_char * 128)()
    check(_lib.load().mp_dist_unique_id(buf))
    return bytes(buf)


def exchange_unique_id(dist, rank):
    """rank 0 creates the NCCL id, torch.distributed broadcasts it (plumbing only)"""
    box = [unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def nccl_library_path():
    """the NCCL copy torch ships (nvidia-nccl wheel); None lets the library pick the one already loaded"""
    try:
        import os
        import nvidia.nccl
        p = os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2")
        return p if os.path.exists(p) else None
    except Exception:
        return None


class ShardedSolver(Solver):
    """FluidSolver whose 3-D grids are this rank's slab (owned planes + 2 ghost planes) of a global grid."""

    def __init__(self, globalGridSize, rank, world, uid, prec=4, device=0):
        gsx, gsy, gsz = (int(v) for v in globalGridSize)
        self.rank, self.world, self.globalGridSize = rank, world, (gsx, gsy, gsz)
        self.k0, self.k1 = slab(gsz, rank, world)
        super().__init__(gridSize=(gsx, gsy, self.k1 - self.k0 + 2), dim=3, prec=prec, device=device)
        path = nccl_library_path()
        check(self.lib.mp_dist_init(self._ctx, C.c_int(rank), C.c_int(world), uid, path.encode() if path else None))
        check(self.lib.mp_dist_set_domain(self._ctx, C.c_int(gsz)))

    def exchangeHalo(self, grid):
        check(self.lib.mp_dist_exchange_halo(self._ctx, grid.dev()))
        grid.markDeviceWritten()


class ShardedBench:
    """bench.py's N>1 runner: res^3 cells per GPU, global grid bench.global_grid(res, N) (1024^3 at res=512, N=8), PcNone."""

    def __init__(self, args, rank, world, device, dist):
        import mantaflow_b200 as mf
        from . import scenes
        self.mf, self.args, self.rank, self.world = mf, args, rank, world
        res, prec = args.res, args.prec
        uid = exchange_unique_id(dist, rank)
        gsx, gsy, gsz = global_grid(res, world)
        self.s = ShardedSolver((gsx, gsy, gsz), rank, world, uid, prec=prec, device=device)
        self.s.setProfiling(16)
        flags, vel = scenes.smoke_plume((gsx, gsy, gsz), prec, zrange=(self.s.k0 - 1, self.s.k1 + 1))
        self.F, self.V0, self.V, self.P = mf.FlagGrid(self.s, flags), mf.MACGrid(self.s, vel), mf.MACGrid(self.s), mf.RealGrid(self.s)
        self.F.dev(); self.V0.dev()
        self.kw = dict(cgAccuracy=1e-4, cgMaxIterFac=99, preconditioner=args.pc, zeroPressureFixing=(args.pc >= 2))
        self.last_info = {}
        self.h_flags, self.h_vel0 = flags, vel
        self.h_vel, self.h_p = vel.copy(), np.zeros(flags.shape, vel.dtype)

    def launches(self):
        return self.s.kernelLaunches()

    def step_resident(self):
        mf = self.mf
        check(self.s.lib.mp_grid_copy_from(self.V.dev(), self.V0.dev()))
        mf.solvePressure(vel=self.V, pressure=self.P, flags=self.F, **self.kw)
        self.last_info = mf.lastSolveInfo()
        return self.last_info

    def run_e2e(self, warmup, steps, barrier):
        import time
        mf = self.mf
        ts = []
        for i in range(warmup + steps):
            self.h_vel[...] = self.h_vel0
            barrier()
            t0 = time.perf_counter()
            mf.solvePressureHost(self.s, self.h_vel, self.h_p, self.h_flags, **self.kw)
            float(self.h_p.ravel()[self.h_p.size // 2])
            barrier()
            if i >= warmup:
                ts.append(1e3 * (time.perf_counter() - t0))
        bi = (self.h_flags.nbytes + self.h_vel.nbytes) * self.world
        bo = (self.h_vel.nbytes + self.h_p.nbytes) * self.world
        return float(np.mean(ts)), bi, bo

    def extra_configs(self):
        """PcMGStatic on the same sharded grid (collective: every rank calls it): the GLOBAL GridMg hierarchy on every rank, levels of >= 4 M
        vertices swept over the rank's planes with boundary-plane exchanges -- first solve incl. setA (cold) and two with the hierarchy reused"""
        mf = self.mf
        kw = dict(self.kw)
        kw.update(preconditioner=3, zeroPressureFixing=True)
        ms, info = [], None
        for _ in range(3):
            check(self.s.lib.mp_grid_copy_from(self.V.dev(), self.V0.dev()))
            mf.solvePressure(vel=self.V, pressure=self.P, flags=self.F, **kw)
            info = mf.lastSolveInfo()
            ms.append(info["msTotal"])
        mf.releaseMG(self.s)
        gx, gy, gz = self.s.globalGridSize
        return {"pcmgstatic": {"grid": [gx, gy, gz], "preconditioner": "PcMGStatic", "cgAccuracy": kw["cgAccuracy"], "iterations": info["iterations"],
                               "solve_ms": min(ms[1:]), "solve_ms_cold": ms[0], "mgLevels": info["mgLevels"],
                               "kernel_ms": {"matvec": info["msMatvecAvg"], "axpy": info["msAxpyAvg"], "precond": info["msPrecondAvg"], "update": info["msUpdateAvg"]},
                               "exchange": "CG loop: " + self.exchange() + "; V-cycle: NCCL send/recv of boundary planes per colour, one all-reduce at the first replicated level"}}

    def exchange(self):
        """how the per-iteration halo planes and scalar partials travelled (mp_dist_exchange_mode)"""
        mode = C.c_int(0)
        check(self.s.lib.mp_dist_exchange_mode(self.s._ctx, C.byref(mode)))
        return {0: "none (1 GPU)", 1: "nccl (send/recv halo planes + all-gather of the partials)",
                2: "p2p (NVLink peer stores into CUDA-IPC mapped arenas + release/acquire flags; NCCL only carries the set-up)"}[mode.value]

    def ncu_traffic(self, kernel):
        return None

"""Host-side mirrors of the solver-level reference classes and kernels, bound to the C-ABI:
GridCg (conjugategrad.h:65-114), GridMg (multigrid.h:31-137), MakeLaplaceMatrix / ApplyMatrix
(conjugategrad.h:118-187), the MIC(0) helpers (conjugategrad.cpp:66-97,:135-159) and the small kernels of
plugin/pressure.cpp that are not PYTHON()-exposed (MakeRhs, ApplyGhostFluidDiagonal, CountEmptyCells, fixPressure)."""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import MP_CG_PC_MGP, MP_CG_PC_MICP, MP_CG_PC_NONE, check


def _d(g):
    return None if g is None else g.dev()


def MakeRhs(flags, rhs, vel, perCellCorr=None, fractions=None, obvel=None, phi=None, curv=None, surfTens=0., gfClamp=1e-4):
    """pressure.cpp:32-84; returns (sum, cnt) like the kernel's reduce members"""
    s = flags.parent
    sm, cnt = C.c_double(0), C.c_int(0)
    check(s.lib.mp_make_rhs(s._ctx, flags.dev(), rhs.dev(), vel.dev(), _d(perCellCorr), _d(fractions), _d(obvel), _d(phi), _d(curv),
                            C.c_double(surfTens), C.c_double(gfClamp), C.byref(sm), C.byref(cnt)))
    rhs.markDeviceWritten()
    return sm.value, cnt.value


def MakeLaplaceMatrix(flags, A0, Ai, Aj, Ak, fractions=None):
    s = flags.parent
    check(s.lib.mp_make_laplace_matrix(s._ctx, flags.dev(), A0.dev(), Ai.dev(), Aj.dev(), Ak.dev(), _d(fractions)))
    for g in (A0, Ai, Aj, Ak):
        g.markDeviceWritten()


def ApplyGhostFluidDiagonal(A0, flags, phi, gfClamp):
    s = flags.parent
    check(s.lib.mp_apply_ghost_fluid_diagonal(s._ctx, A0.dev(), flags.dev(), phi.dev(), C.c_double(gfClamp)))
    A0.markDeviceWritten()


def CountEmptyCells(flags):
    s = flags.parent
    n = C.c_longlong(0)
    check(s.lib.mp_count_empty_cells(s._ctx, flags.dev(), C.byref(n)))
    return n.value


def chooseFixCell(flags):
    s = flags.parent
    n = C.c_longlong(0)
    check(s.lib.mp_choose_fix_cell(s._ctx, flags.dev(), C.byref(n)))
    return n.value


def fixPressure(fixPidx, value, rhs, A0, Ai, Aj, Ak):
    s = rhs.parent
    check(s.lib.mp_fix_pressure(s._ctx, C.c_longlong(fixPidx), C.c_double(value), rhs.dev(), A0.dev(), Ai.dev(), Aj.dev(), Ak.dev()))
    for g in (rhs, A0, Ai, Aj, Ak):
        g.markDeviceWritten()


def ApplyMatrix(flags, dst, src, A0, Ai, Aj, Ak):
    """ApplyMatrix / ApplyMatrix2D by dimensionality"""
    s = flags.parent
    check(s.lib.mp_apply_matrix(s._ctx, flags.dev(), dst.dev(), src.dev(), A0.dev(), Ai.dev(), Aj.dev(), Ak.dev()))
    dst.markDeviceWritten()


def InitPreconditionModifiedIncompCholesky2(flags, Aprecond, A0, Ai, Aj, Ak):
    s = flags.parent
    check(s.lib.mp_mic_init(s._ctx, flags.dev(), Aprecond.dev(), A0.dev(), Ai.dev(), Aj.dev(), Ak.dev()))
    Aprecond.markDeviceWritten()


def ApplyPreconditionModifiedIncompCholesky2(dst, Var1, flags, Aprecond, A0, Ai, Aj, Ak):
    s = flags.parent
    check(s.lib.mp_mic_apply(s._ctx, dst.dev(), Var1.dev(), flags.dev(), Aprecond.dev(), A0.dev(), Ai.dev(), Aj.dev(), Ak.dev()))
    dst.markDeviceWritten()


def InitPreconditionIncompCholesky(flags, A0, Ai, Aj, Ak, orgA0, orgAi, orgAj, orgAk):
    """IC(0) "a la Wavelet Turbulence" conjugategrad.cpp:26-63: A0..Ak receive the factor of orgA0..orgAk"""
    s = flags.parent
    check(s.lib.mp_ic_init(s._ctx, flags.dev(), A0.dev(), Ai.dev(), Aj.dev(), Ak.dev(), orgA0.dev(), orgAi.dev(), orgAj.dev(), orgAk.dev()))
    for g in (A0, Ai, Aj, Ak):
        g.markDeviceWritten()


def ApplyPreconditionIncompCholesky(dst, Var1, flags, A0, Ai, Aj, Ak, orgA0=None, orgAi=None, orgAj=None, orgAk=None):
    """conjugategrad.cpp:109-132 (the org* arguments are unused there as well)"""
    s = flags.parent
    check(s.lib.mp_ic_apply(s._ctx, dst.dev(), Var1.dev(), flags.dev(), A0.dev(), Ai.dev(), Aj.dev(), Ak.dev()))
    dst.markDeviceWritten()


def GridDotProduct(a, b):
    s = a.parent
    out = C.c_double(0)
    check(s.lib.mp_grid_dot(s._ctx, a.dev(), b.dev(), C.byref(out)))
    return out.value


def cgSolveDiffusion(flags, grid, alpha=0.25, cgMaxIterFac=1.0, cgAccuracy=1e-4):
    """conjugategrad.cpp:350-423 (PYTHON() plugin): implicit diffusion of a Real or Vec3/MAC grid on the device GridCg"""
    from ._lib import SolveInfo
    s = flags.parent
    info = SolveInfo()
    check(s.lib.mp_cg_solve_diffusion(s._ctx, flags.dev(), grid.dev(), C.c_double(alpha), C.c_double(cgMaxIterFac), C.c_double(cgAccuracy), C.byref(info)))
    grid.markDeviceWritten()
    return info.as_dict()


def cgSolveWE(flags, ut, utm1, out, crankNic=False, cSqr=0.25, cgMaxIterFac=1.5, cgAccuracy=1e-5):
    """plugin/waves.cpp:86-147 (PYTHON() plugin): implicit wave-equation step on the device GridCg; utm1 <- ut, ut <- out"""
    from ._lib import SolveInfo
    s = flags.parent
    info = SolveInfo()
    check(s.lib.mp_cg_solve_we(s._ctx, flags.dev(), ut.dev(), utm1.dev(), out.dev(), C.c_int(int(bool(crankNic))), C.c_double(cSqr),
                               C.c_double(cgMaxIterFac), C.c_double(cgAccuracy), C.c_double(s.timestep), C.byref(info)))
    for g in (ut, utm1, out):
        g.markDeviceWritten()
    return info.as_dict()


def vicPoisson(vel, flags, vorticity, cgMaxIterFac=1.5, cgAccuracy=1e-3, scale=0.01, precondition=0):
    """The grid half of VICintegration (plugin/vortexplugins.cpp:253-299; parameter names and defaults of the plugin :195-196): from the vorticity
    grid the plugin's Peskin kernel leaves (:203-250) to `vel` (MACGrid: shifted components, VecGrid: centred) through three GridCg solves
    preconditioned with PC_ICP (precondition=1) or PC_mICP (2).  precondition=0, the plugin's default, raises setICPreconditioner's error as it
    does in the reference (conjugategrad.cpp:312).  Returns the three iteration counts."""
    from .grid import VecGrid
    s = flags.parent
    its = (C.c_int * 3)()
    check(s.lib.mp_vic_poisson(s._ctx, flags.dev(), vorticity.dev(), vel.dev(), C.c_int(0 if isinstance(vel, VecGrid) else 1), C.c_double(cgMaxIterFac),
                               C.c_double(cgAccuracy), C.c_double(scale), C.c_int(precondition), its))
    vel.markDeviceWritten()
    return list(its)


class GridMg:
    """multigrid.h:31-137"""

    def __init__(self, solver):
        self.solver = solver
        sx, sy, sz = solver.gridSize
        self._h = C.c_void_p()
        check(solver.lib.mp_mg_create(solver._ctx, C.c_int(solver.prec), C.c_int(sx), C.c_int(sy), C.c_int(sz), C.byref(self._h)))
        solver._adopt(self)

    def setA(self, A0, Ai, Aj, Ak):
        check(self.solver.lib.mp_mg_set_a(self._h, A0.dev(), Ai.dev(), Aj.dev(), Ak.dev()))

    def setRhs(self, rhs):
        check(self.solver.lib.mp_mg_set_rhs(self._h, rhs.dev()))

    def isASet(self):
        v = C.c_int(0)
        check(self.solver.lib.mp_mg_is_a_set(self._h, C.byref(v)))
        return bool(v.value)

    def doVCycle(self, dst, src=None):
        res = C.c_double(0)
        check(self.solver.lib.mp_mg_do_vcycle(self._h, dst.dev(), _d(src), C.byref(res)))
        dst.markDeviceWritten()
        return res.value

    def setCoarsestLevelAccuracy(self, accuracy):
        check(self.solver.lib.mp_mg_set_coarsest_level_accuracy(self._h, C.c_double(accuracy)))

    def setSmoothing(self, numPreSmooth, numPostSmooth):
        check(self.solver.lib.mp_mg_set_smoothing(self._h, C.c_int(numPreSmooth), C.c_int(numPostSmooth)))

    # parity probes
    def numLevels(self):
        v = C.c_int(0)
        check(self.solver.lib.mp_mg_num_levels(self._h, C.byref(v)))
        return v.value

    def levelInfo(self, l):
        a, b, c, st = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        check(self.solver.lib.mp_mg_level_info(self._h, C.c_int(l), C.byref(a), C.byref(b), C.byref(c), C.byref(st)))
        return (a.value, b.value, c.value), st.value

    def level0Fused(self):
        """True when level 0 of the V-cycle runs as the fused single-pass kernels (mp_mg_level0_fused)"""
        v = C.c_int(0)
        check(self.solver.lib.mp_mg_level0_fused(self._h, C.byref(v)))
        return bool(v.value)

    def download(self, what, l):
        (sx, sy, sz), st = self.levelInfo(l)
        n = sx * sy * sz
        if what == "type":
            out = np.zeros(n, np.int8)
        elif what == "a":
            out = np.zeros(n * st, self.solver.real)
        else:
            out = np.zeros(n, self.solver.real)
        check(self.solver.lib.mp_mg_download(self._h, C.c_int(l), what.encode(), out.ctypes.data_as(C.c_void_p)))
        return out

    def close(self):
        if self._h and self.solver._ctx:          # the handle points into the context: once that is gone (Solver.close released us first) nothing is left to free
            self.solver.lib.mp_mg_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GridCg:
    """GridCg<ApplyMatrix|ApplyMatrix2D> (conjugategrad.h:65-114); the template argument is chosen by dimensionality."""
    PC_None, PC_ICP, PC_mICP, PC_MGP = 0, 1, 2, 3

    def __init__(self, dst, rhs, residual, search, flags, tmp, A0, Ai, Aj, Ak):
        self.solver = s = flags.parent
        self._grids = (dst, rhs, residual, search, flags, tmp, A0, Ai, Aj, Ak)
        self._written = (dst, residual, search, tmp)
        self._h = C.c_void_p()
        check(s.lib.mp_cg_create(s._ctx, dst.dev(), rhs.dev(), residual.dev(), search.dev(), flags.dev(), tmp.dev(),
                                 A0.dev(), Ai.dev(), Aj.dev(), Ak.dev(), C.byref(self._h)))
        s._adopt(self)
        self._keep = []

    def setAccuracy(self, v):
        check(self.solver.lib.mp_cg_set_accuracy(self._h, C.c_double(v)))

    def setUseL2Norm(self, v):
        check(self.solver.lib.mp_cg_set_use_l2_norm(self._h, C.c_int(int(bool(v)))))

    def setICPreconditioner(self, method, A0=None, Ai=None, Aj=None, Ak=None):
        self._keep = [A0, Ai, Aj, Ak]
        check(self.solver.lib.mp_cg_set_ic_preconditioner(self._h, C.c_int(method), _d(A0), _d(Ai), _d(Aj), _d(Ak)))

    def setMGPreconditioner(self, method, mg):
        self._keep = [mg]
        check(self.solver.lib.mp_cg_set_mg_preconditioner(self._h, C.c_int(method), mg._h))

    def forceReinit(self):
        check(self.solver.lib.mp_cg_force_reinit(self._h))

    def _mark(self):
        for g in self._written:
            g.markDeviceWritten()
        for g in self._keep:          # the preconditioner grids of PC_ICP / PC_mICP are written by doInit
            if hasattr(g, "markDeviceWritten"):
                g.markDeviceWritten()

    def iterate(self):
        for g in self._grids:
            g.dev()
        keep = C.c_int(0)
        rc = self.solver.lib.mp_cg_iterate(self._h, C.byref(keep))
        self._mark()
        check(rc)
        return bool(keep.value)

    def solve(self, maxIter):
        for g in self._grids:
            g.dev()
        rc = self.solver.lib.mp_cg_solve(self._h, C.c_int(maxIter))
        self._mark()
        check(rc)

    def _get(self):
        it, rn, sg = C.c_int(0), C.c_double(0), C.c_double(0)
        check(self.solver.lib.mp_cg_get(self._h, C.byref(it), C.byref(rn), C.byref(sg)))
        return it.value, rn.value, sg.value

    def getIterations(self): return self._get()[0]
    def getResNorm(self): return self._get()[1]
    def getSigma(self): return self._get()[2]

    def close(self):
        if self._h and self.solver._ctx:
            self.solver.lib.mp_cg_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""The PYTHON() plugins of plugin/pressure.cpp with unchanged names, parameter names, order and defaults
(pressure.cpp:252, :277-292, :312-326, :455-468, :480-495), running on the device through the C-ABI.
Unknown keyword arguments are an error, as in the reference (PbArgs::check pconvert.cpp:459-473)."""
import ctypes as C

from . import _lib
from ._lib import PcMIC, PressureParams, SolveInfo, check

_last_info = None


def lastSolveInfo():
    """Iterations / residual of the last solve: what the reference only prints at debug level 2 (pressure.cpp:440)."""
    return _last_info


def _params(cgAccuracy, gfClamp, cgMaxIterFac, precondition, preconditioner, enforceCompatibility, useL2Norm, zeroPressureFixing, surfTens):
    p = PressureParams()
    p.cgAccuracy, p.gfClamp, p.cgMaxIterFac = float(cgAccuracy), float(gfClamp), float(cgMaxIterFac)
    p.precondition, p.preconditioner = int(bool(precondition)), int(preconditioner)
    p.enforceCompatibility, p.useL2Norm, p.zeroPressureFixing = int(bool(enforceCompatibility)), int(bool(useL2Norm)), int(bool(zeroPressureFixing))
    p.surfTens = float(surfTens)
    return p


def _d(g):
    return None if g is None else g.dev()


def releaseMG(solver=None):
    """pressure.cpp:252-266"""
    if solver is not None:
        check(solver.lib.mp_release_mg(solver._ctx))


def computePressureRhs(rhs, vel, pressure, flags, cgAccuracy=1e-3, phi=None, perCellCorr=None, fractions=None, obvel=None,
                       gfClamp=1e-04, cgMaxIterFac=1.5, precondition=True, preconditioner=PcMIC, enforceCompatibility=False,
                       useL2Norm=False, zeroPressureFixing=False, curv=None, surfTens=0.):
    s = flags.parent
    p = _params(cgAccuracy, gfClamp, cgMaxIterFac, precondition, preconditioner, enforceCompatibility, useL2Norm, zeroPressureFixing, surfTens)
    check(s.lib.mp_compute_pressure_rhs(s._ctx, rhs.dev(), vel.dev(), pressure.dev(), flags.dev(), _d(phi), _d(perCellCorr),
                                        _d(fractions), _d(obvel), _d(curv), C.byref(p)))
    rhs.markDeviceWritten()


def solvePressureSystem(rhs, vel, pressure, flags, cgAccuracy=1e-3, phi=None, perCellCorr=None, fractions=None,
                        gfClamp=1e-04, cgMaxIterFac=1.5, precondition=True, preconditioner=PcMIC, enforceCompatibility=False,
                        useL2Norm=False, zeroPressureFixing=False, curv=None, surfTens=0.):
    global _last_info
    s = flags.parent
    p = _params(cgAccuracy, gfClamp, cgMaxIterFac, precondition, preconditioner, enforceCompatibility, useL2Norm, zeroPressureFixing, surfTens)
    info = SolveInfo()
    rc = s.lib.mp_solve_pressure_system(s._ctx, rhs.dev(), vel.dev(), pressure.dev(), flags.dev(), _d(phi), _d(perCellCorr),
                                        _d(fractions), _d(curv), C.byref(p), C.byref(info))
    rhs.markDeviceWritten()          # fixPressure edits rhs (pressure.cpp:229-239)
    pressure.markDeviceWritten()
    _last_info = info.as_dict()
    check(rc)


def correctVelocity(vel, pressure, flags, cgAccuracy=1e-3, phi=None, perCellCorr=None, fractions=None,
                    gfClamp=1e-04, cgMaxIterFac=1.5, precondition=True, preconditioner=PcMIC, enforceCompatibility=False,
                    useL2Norm=False, zeroPressureFixing=False, curv=None, surfTens=0.):
    s = flags.parent
    p = _params(cgAccuracy, gfClamp, cgMaxIterFac, precondition, preconditioner, enforceCompatibility, useL2Norm, zeroPressureFixing, surfTens)
    check(s.lib.mp_correct_velocity(s._ctx, vel.dev(), pressure.dev(), flags.dev(), _d(phi), _d(curv), C.byref(p)))
    vel.markDeviceWritten()


def solvePressure(vel, pressure, flags, cgAccuracy=1e-3, phi=None, perCellCorr=None, fractions=None, obvel=None,
                  gfClamp=1e-04, cgMaxIterFac=1.5, precondition=True, preconditioner=PcMIC, enforceCompatibility=False,
                  useL2Norm=False, zeroPressureFixing=False, curv=None, surfTens=0., retRhs=None):
    """pressure.cpp:480-521.  Grids stay resident in HBM; host copies are refreshed lazily on .numpy()."""
    global _last_info
    s = flags.parent
    p = _params(cgAccuracy, gfClamp, cgMaxIterFac, precondition, preconditioner, enforceCompatibility, useL2Norm, zeroPressureFixing, surfTens)
    info = SolveInfo()
    rc = s.lib.mp_solve_pressure(s._ctx, vel.dev(), pressure.dev(), flags.dev(), _d(phi), _d(perCellCorr), _d(fractions),
                                 _d(obvel), _d(curv), _d(retRhs), C.byref(p), C.byref(info))
    vel.markDeviceWritten()
    pressure.markDeviceWritten()
    if retRhs is not None:
        retRhs.markDeviceWritten()
    _last_info = info.as_dict()
    check(rc)


def solvePressureHost(solver, vel, pressure, flags, phi=None, perCellCorr=None, fractions=None, obvel=None, curv=None, retRhs=None,
                      **kw):
    """mp_solve_pressure_host: the plugin on raw HOST arrays (numpy, reference layout); H2D + solve + D2H in one call.
    Returns the mp_solve_info dict.  This is the call a maintainer's pressure.cpp makes for grids without a device mirror."""
    defaults = dict(cgAccuracy=1e-3, gfClamp=1e-04, cgMaxIterFac=1.5, precondition=True, preconditioner=PcMIC,
                    enforceCompatibility=False, useL2Norm=False, zeroPressureFixing=False, surfTens=0.)
    for k in kw:
        if k not in defaults:
            raise TypeError("solvePressureHost: unknown keyword argument '%s'" % k)
    defaults.update(kw)
    p = _params(**defaults)
    info = SolveInfo()
    sx, sy, sz = solver.gridSize

    def ptr(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = solver.lib.mp_solve_pressure_host(solver._ctx, C.c_int(solver.prec), C.c_int(sx), C.c_int(sy), C.c_int(sz),
                                           ptr(vel), ptr(pressure), ptr(flags), ptr(phi), ptr(perCellCorr), ptr(fractions),
                                           ptr(obvel), ptr(curv), ptr(retRhs), C.byref(p), C.byref(info))
    check(rc)
    return info.as_dict()

"""The PYTHON() plugins either side of the pressure projection, on the device -- same names, arguments and
defaults as the reference, so the main loop of scenes/simpleplume.py runs with every field resident in HBM:

    setWallBcs          plugin/extforces.cpp:307-316
    addGravity          plugin/extforces.cpp:61-65      addGravityNoScale :67-69
    addBuoyancy         plugin/extforces.cpp:86-90
    advectSemiLagrange  plugin/advection.cpp:442-461
    PD_fluid_guiding    plugin/fluidguiding.cpp:294-353  (+ getSpiralVelocity :171-192, setGradientYWeight :195-207 for its scenes)

and the liquid neighbours (first slice) with which the level-set loop of scenes/freesurface.py:54-84 does the same:

    extrapolateMACSimple   fastmarch.cpp:337-375      extrapolateLsSimple :470-507      extrapolateVec3Simple :510-542
    extrapolateMACFromWeight :410-432
    getLaplacian, getCurvature   plugin/flip.cpp:710-716
    updateFractions, setObstacleFlags   plugin/initplugins.cpp:437-440,:473-475
    (FlagGrid.updateFromLevelset and Grid.setBound are methods of the grid classes, grid.py)
"""
import ctypes as C

from ._lib import MantaError, check


def _d(g):
    return None if g is None else g.dev()


def _vec3(v):
    try:
        x, y, z = (float(c) for c in v)
    except TypeError:
        x = y = z = float(v)
    return C.c_double(x), C.c_double(y), C.c_double(z)


def setWallBcs(flags, vel, obvel=None, fractions=None, phiObs=None, boundaryWidth=0):
    s = flags.parent
    check(s.lib.mp_set_wall_bcs(s._ctx, flags.dev(), vel.dev(), _d(obvel), _d(fractions), _d(phiObs), C.c_int(boundaryWidth)))
    vel.markDeviceWritten()


def addGravity(flags, vel, gravity, exclude=None, scale=True):
    s = flags.parent
    check(s.lib.mp_add_gravity(s._ctx, flags.dev(), vel.dev(), *_vec3(gravity), _d(exclude), C.c_int(int(bool(scale))), C.c_double(s.timestep)))
    vel.markDeviceWritten()


def addGravityNoScale(flags, vel, gravity, exclude=None):
    addGravity(flags, vel, gravity, exclude, False)


def addBuoyancy(flags, density, vel, gravity, coefficient=1., scale=True):
    s = flags.parent
    check(s.lib.mp_add_buoyancy(s._ctx, flags.dev(), density.dev(), vel.dev(), *_vec3(gravity), C.c_double(coefficient), C.c_int(int(bool(scale))),
                                C.c_double(s.timestep)))
    vel.markDeviceWritten()


def advectSemiLagrange(flags, vel, grid, order=1, strength=1.0, orderSpace=1, openBounds=False, boundaryWidth=-1, clampMode=2, orderTrace=1):
    """openBounds / boundaryWidth are deprecated in the reference and have no effect (advection.cpp:446)"""
    s = flags.parent
    if order not in (1, 2):
        raise MantaError(1, "AdvectSemiLagrange: Only order 1 (regular SL) and 2 (MacCormack) supported")
    from .grid import VecGrid
    # a Grid<Vec3> has the storage of a MACGrid but is cell-centred: fnAdvectSemiLagrange<Grid<Vec3>> (advection.cpp:455-457), not the MAC kernels
    fn = s.lib.mp_advect_semi_lagrange_vec3 if isinstance(grid, VecGrid) else s.lib.mp_advect_semi_lagrange
    check(fn(s._ctx, flags.dev(), vel.dev(), grid.dev(), C.c_int(order), C.c_double(strength), C.c_int(orderSpace),
             C.c_int(clampMode), C.c_int(orderTrace), C.c_double(s.timestep)))
    grid.markDeviceWritten()


def extrapolateMACSimple(flags, vel, distance=4, phiObs=None, intoObs=False):
    s = flags.parent
    check(s.lib.mp_extrapolate_mac_simple(s._ctx, flags.dev(), vel.dev(), C.c_int(int(distance)), _d(phiObs), C.c_int(int(bool(intoObs)))))
    vel.markDeviceWritten()


def extrapolateMACFromWeight(vel, weight, distance=2):
    s = vel.parent
    check(s.lib.mp_extrapolate_mac_from_weight(s._ctx, vel.dev(), weight.dev(), C.c_int(int(distance))))
    vel.markDeviceWritten(); weight.markDeviceWritten()


def extrapolateLsSimple(phi, distance=4, inside=False):
    s = phi.parent
    check(s.lib.mp_extrapolate_ls_simple(s._ctx, phi.dev(), C.c_int(int(distance)), C.c_int(int(bool(inside)))))
    phi.markDeviceWritten()


def extrapolateVec3Simple(vel, phi, distance=4, inside=False):
    s = vel.parent
    check(s.lib.mp_extrapolate_vec3_simple(s._ctx, vel.dev(), phi.dev(), C.c_int(int(distance)), C.c_int(int(bool(inside)))))
    vel.markDeviceWritten()


def updateFractions(flags, phiObs, fractions, boundaryWidth=0, fracThreshold=0.01):
    """plugin/initplugins.cpp:437-440"""
    s = flags.parent
    check(s.lib.mp_update_fractions(s._ctx, flags.dev(), phiObs.dev(), fractions.dev(), C.c_int(int(boundaryWidth)), C.c_double(fracThreshold)))
    fractions.markDeviceWritten()


def setObstacleFlags(flags, phiObs, fractions=None, phiOut=None, phiIn=None, boundaryWidth=1):
    """plugin/initplugins.cpp:473-475"""
    s = flags.parent
    check(s.lib.mp_set_obstacle_flags(s._ctx, flags.dev(), phiObs.dev(), _d(fractions), _d(phiOut), _d(phiIn), C.c_int(int(boundaryWidth))))
    flags.markDeviceWritten()


def getLaplacian(laplacian, grid):
    """plugin/flip.cpp:710-712"""
    s = grid.parent
    check(s.lib.mp_get_laplacian(s._ctx, laplacian.dev(), grid.dev()))
    laplacian.markDeviceWritten()


def getCurvature(curv, grid, h=1.0):
    """plugin/flip.cpp:714-716: the `curv` argument of solvePressure's surface-tension variant"""
    s = grid.parent
    check(s.lib.mp_get_curvature(s._ctx, curv.dev(), grid.dev(), C.c_double(h)))
    curv.markDeviceWritten()


_last_guiding = {"iterations": -1}


def PD_fluid_guiding(vel, velT, pressure, flags, weight, blurRadius=5, theta=1.0, tau=1.0, sigma=1.0, epsRel=1e-3, epsAbs=1e-3, maxIters=200,
                     phi=None, perCellCorr=None, fractions=None, obvel=None, gfClamp=1e-04, cgMaxIterFac=1.5, cgAccuracy=1e-3,
                     preconditioner=1, zeroPressureFixing=False, curv=None, surfTens=0.):
    """primal-dual fluid guiding around solvePressure, every temporary on the device; the loop count the reference prints is kept in
    lastGuidingIterations()"""
    s = flags.parent
    it = C.c_int(-1)
    check(s.lib.mp_pd_fluid_guiding(s._ctx, vel.dev(), velT.dev(), pressure.dev(), flags.dev(), weight.dev(), C.c_int(blurRadius), C.c_double(theta),
                                    C.c_double(tau), C.c_double(sigma), C.c_double(epsRel), C.c_double(epsAbs), C.c_int(maxIters), _d(phi), _d(perCellCorr),
                                    _d(fractions), _d(obvel), C.c_double(gfClamp), C.c_double(cgMaxIterFac), C.c_double(cgAccuracy), C.c_int(preconditioner),
                                    C.c_int(int(bool(zeroPressureFixing))), _d(curv), C.c_double(surfTens), C.byref(it)))
    vel.markDeviceWritten(); pressure.markDeviceWritten()
    _last_guiding["iterations"] = it.value


def lastGuidingIterations():
    return _last_guiding["iterations"]


def releaseBlurPrecomp():
    """fluidguiding.cpp:356-360: the reference caches one blur kernel in globals; here it is rebuilt per call, nothing to release"""

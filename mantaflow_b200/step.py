"""The PYTHON() plugins either side of the pressure projection, on the device -- same names, arguments and
defaults as the reference, so the main loop of scenes/simpleplume.py runs with every field resident in HBM:

    setWallBcs          plugin/extforces.cpp:307-316
    addGravity          plugin/extforces.cpp:61-65      addGravityNoScale :67-69
    addBuoyancy         plugin/extforces.cpp:86-90
    advectSemiLagrange  plugin/advection.cpp:442-461
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
    check(s.lib.mp_advect_semi_lagrange(s._ctx, flags.dev(), vel.dev(), grid.dev(), C.c_int(order), C.c_double(strength), C.c_int(orderSpace),
                                        C.c_int(clampMode), C.c_int(orderTrace), C.c_double(s.timestep)))
    grid.markDeviceWritten()

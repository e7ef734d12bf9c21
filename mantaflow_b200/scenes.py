"""Synthetic inputs for the pressure path (numpy, reference layout [Z,Y,X] / [Z,Y,X,3]).

These are the BASELINE.json setups: the closed-box smoke plume with a sphere obstacle and a box
velocity source, a pseudo-random divergence-ful variant, and a liquid basin + drop with a level set for the
ghost-fluid path.  Scene construction is host code outside the hot path; the same arrays feed the CUDA path,
the oracle and the reference."""
import numpy as np

FlagFluid, FlagObstacle, FlagEmpty, FlagInflow, FlagOutflow, FlagOpen, FlagStick = 1, 2, 4, 8, 16, 32, 64


def _coords(sx, sy, sz):
    k = np.arange(sz).reshape(sz, 1, 1)
    j = np.arange(sy).reshape(1, sy, 1)
    i = np.arange(sx).reshape(1, 1, sx)
    return i, j, k


def closed_box_flags(sx, sy, sz, boundaryWidth=0):
    """flags.initDomain(boundaryWidth) + fillGrid()  (grid.cpp:732-861)"""
    f = np.full((sz, sy, sx), FlagFluid, np.int32)
    w = boundaryWidth + 1
    f[:, :, :w] = FlagObstacle; f[:, :, sx - w:] = FlagObstacle
    f[:, :w, :] = FlagObstacle; f[:, sy - w:, :] = FlagObstacle
    if sz > 1:
        f[:w, :, :] = FlagObstacle; f[sz - w:, :, :] = FlagObstacle
    return f


def set_wall_bcs(flags, vel):
    """KnSetWallBcs without obvel (plugin/extforces.cpp:186-218): zero normal velocity on obstacle faces."""
    fl = (flags & FlagFluid) != 0
    ob = (flags & FlagObstacle) != 0
    act = fl | ob
    is3d = flags.shape[0] > 1
    m = np.zeros_like(ob); m[:, :, 1:] = ob[:, :, :-1] | (ob[:, :, 1:] & fl[:, :, :-1]); vel[..., 0][m & act] = 0
    m = np.zeros_like(ob); m[:, 1:, :] = ob[:, :-1, :] | (ob[:, 1:, :] & fl[:, :-1, :]); vel[..., 1][m & act] = 0
    if is3d:
        m = np.zeros_like(ob); m[1:, :, :] = ob[:-1, :, :] | (ob[1:, :, :] & fl[:-1, :, :]); vel[..., 2][m & act] = 0
    else:
        vel[..., 2][act] = 0
    return vel


def smoke_plume(res, prec=4, obstacle=True, random_vel=False, seed=1234, scale=1.0, zrange=None):
    """Closed box, sphere obstacle r=0.12 res at (0.5,0.6,0.5) res, velocity (0.15,0.3,0.21) inside the box
    (0.3,0.1,0.3)-(0.7,0.3,0.7) res, then setWallBcs.  `res` is an int or (sx,sy,sz).
    zrange=(za,zb) builds only the global planes [za,zb) (a z-slab with its ghost planes; planes outside the
    domain are zero) without ever allocating the global grid."""
    sx, sy, sz = (res, res, res) if np.isscalar(res) else res
    real = np.float32 if prec == 4 else np.float64
    if zrange is None:
        ks = np.arange(sz)
    else:
        assert sz > 1 and not random_vel
        ks = np.arange(zrange[0] - 1, zrange[1])          # one extra plane below for the z wall condition, cropped at the end
    nz = len(ks)
    k = ks.reshape(nz, 1, 1)
    j = np.arange(sy).reshape(1, sy, 1)
    i = np.arange(sx).reshape(1, 1, sx)
    flags = np.full((nz, sy, sx), FlagFluid, np.int32)
    flags[:, :, 0] = FlagObstacle; flags[:, :, sx - 1] = FlagObstacle
    flags[:, 0, :] = FlagObstacle; flags[:, sy - 1, :] = FlagObstacle
    if sz > 1:
        flags[(ks == 0) | (ks == sz - 1)] = FlagObstacle
        flags[(ks < 0) | (ks >= sz)] = 0
    if obstacle:
        r2 = (i + 0.5 - 0.5 * sx) ** 2 + (j + 0.5 - 0.6 * sy) ** 2 + ((k + 0.5 - 0.5 * sz) ** 2 if sz > 1 else 0)
        flags[np.broadcast_to(r2 <= (0.12 * max(sx, sy, sz)) ** 2, flags.shape) & (flags == FlagFluid)] = FlagObstacle
    vel = np.zeros((nz, sy, sx, 3), real)
    if random_vel:
        rng = np.random.Generator(np.random.PCG64(seed))
        vel[...] = (0.1 * (rng.random(vel.shape) - 0.5)).astype(real)
    inbox = (i > 0.3 * sx) & (i < 0.7 * sx) & (j > 0.1 * sy) & (j < 0.3 * sy)
    if sz > 1:
        inbox = inbox & (k > 0.3 * sz) & (k < 0.7 * sz)
    inbox = np.broadcast_to(inbox, flags.shape)
    vel[inbox] = np.array([0.15, 0.3, 0.21 if sz > 1 else 0.0], real) * scale
    set_wall_bcs(flags, vel)
    if zrange is not None:
        flags, vel = np.ascontiguousarray(flags[1:]), np.ascontiguousarray(vel[1:])
        vel[(flags == 0)] = 0
    return flags, vel


def liquid_basin(res, prec=4, seed=7, open_top=False):
    """Basin (y < 0.2 res) plus a drop of radius 0.15 res at (0.5,0.5,0.5) res (test_2050_freesurface.py:39-43),
    flags from the level set (updateFromLevelset), velocity: gravity kick + small noise inside the liquid."""
    sx, sy, sz = (res, res, res) if np.isscalar(res) else res
    real = np.float32 if prec == 4 else np.float64
    flags = closed_box_flags(sx, sy, sz)
    i, j, k = _coords(sx, sy, sz)
    basin = (j + 0.5) - 0.2 * sy
    drop = np.sqrt((i + 0.5 - 0.5 * sx) ** 2 + (j + 0.5 - 0.5 * sy) ** 2 + ((k + 0.5 - 0.5 * sz) ** 2 if sz > 1 else 0)) - 0.15 * max(sx, sy, sz)
    phi = np.broadcast_to(np.minimum(basin, drop), flags.shape).astype(real)
    nonobs = (flags & (FlagObstacle | FlagOutflow)) == 0
    flags[nonobs] = np.where(phi[nonobs] <= 0, FlagFluid, FlagEmpty).astype(np.int32)
    rng = np.random.Generator(np.random.PCG64(seed))
    vel = (0.02 * (rng.random((sz, sy, sx, 3)) - 0.5)).astype(real)
    vel[..., 1] -= real(0.25)
    if sz == 1:
        vel[..., 2] = 0
    vel[(flags & (FlagFluid | FlagEmpty)) == 0] = 0
    set_wall_bcs(flags, vel)
    return flags, vel, np.ascontiguousarray(phi)


def random_fractions(flags, prec=4, seed=11):
    """Face fractions in [0.25,1] and an obstacle velocity field for the 2nd-order-boundary code paths."""
    real = np.float32 if prec == 4 else np.float64
    rng = np.random.Generator(np.random.PCG64(seed))
    frac = (0.25 + 0.75 * rng.random(flags.shape + (3,))).astype(real)
    obvel = (0.05 * (rng.random(flags.shape + (3,)) - 0.5)).astype(real)
    return frac, obvel


def max_divergence(flags, vel):
    """max |div v| over fluid cells (what computePressureRhs + getMaxAbs gives in the reference, SURVEY 8c iv)"""
    fl = (flags & FlagFluid) != 0
    d = np.zeros(flags.shape, np.float64)
    v = vel.astype(np.float64)
    d[:, :, :-1] += v[:, :, 1:, 0] - v[:, :, :-1, 0]
    d[:, :-1, :] += v[:, 1:, :, 1] - v[:, :-1, :, 1]
    if flags.shape[0] > 1:
        d[:-1, :, :] += v[1:, :, :, 2] - v[:-1, :, :, 2]
    return float(np.abs(d[fl]).max()) if fl.any() else 0.0

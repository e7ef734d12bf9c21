"""The fused level-0 V-cycle kernels (mantaflow_b200/csrc/mp_mg_l0_fused.cuh: zero iterate + both colours + residual in one pass;
both colours of a sweep in one pass; the operator as 2 bytes per cell) walked on the host, thread by thread in the kernel's phase order
(tests/emul/mg_l0_emul.cpp), against a plain numpy statement of knSmoothColor / knCalcResidual (multigrid.cpp:668-711,:739-771) on
level 0 -- bit for bit.  The GPU run of the same code is compared with the unfused kernels in tests/test_gpu_multigrid.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "emul", "mg_l0_emul.cpp")
    hdr = os.path.join(HERE, "..", "mantaflow_b200", "csrc", "mp_mg_l0_fused.cuh")
    out = os.path.join(HERE, "emul", "_build", "libmg_l0_emul.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", src, "-o", out])
    return ctypes.CDLL(out)


def make_level0(shape, dtype, seed, is3D=True, ghost=True):
    """a level-0 operator the way GridMg::setA sees it after MakeLaplaceMatrix (+ ghost-fluid diagonals, + a pinned trivial row)"""
    rng = np.random.default_rng(seed)
    sz, sy, sx = shape
    # fluid blobs inside a one-cell wall; some obstacles and some empty cells inside
    kind = np.zeros(shape, np.int8)                      # 0 obstacle, 1 fluid, 2 empty
    inner = (slice(1, -1) if is3D else slice(None), slice(1, -1), slice(1, -1))
    r = rng.random(shape)
    kind[inner] = np.where(r[inner] < 0.75, 1, np.where(r[inner] < 0.9, 2, 0))
    fluid = kind == 1
    A = np.zeros((4,) + shape, dtype)
    nonobs = kind != 0
    cnt = np.zeros(shape, np.int32)
    for ax in ([0, 1, 2] if is3D else [1, 2]):
        for sh in (1, -1):
            nb = np.roll(nonobs, sh, axis=ax)
            sl = [slice(None)] * 3
            sl[ax] = 0 if sh == 1 else -1
            nb[tuple(sl)] = False
            cnt += nb
    A[0][fluid] = cnt[fluid].astype(dtype)
    def upper(ax):
        nbf = np.roll(fluid, -1, axis=ax)
        sl = [slice(None)] * 3
        sl[ax] = -1
        nbf[tuple(sl)] = False
        return np.where(fluid & nbf, dtype(-1), dtype(0))
    A[1] = upper(2); A[2] = upper(1)
    if is3D:
        A[3] = upper(0)
    if ghost:      # ghost-fluid diagonals: fluid cells with an empty neighbour get a non-integer diagonal
        sel = fluid & (rng.random(shape) < 0.1)
        A[0][sel] = (A[0][sel] + rng.random(sel.sum()).astype(dtype) * dtype(3.7)).astype(dtype)
    typ = np.where(A[0] != 0, 1, 0).astype(np.int8)
    # a pinned cell: identity row, couplings removed on both sides (fixPressure), diagonal scaled by 1e-6 in the level-0 copy
    idx = np.argwhere(fluid)
    pz, py, px = idx[len(idx) // 2]
    A[0][pz, py, px] = dtype(1) * dtype(1e-6)
    A[1][pz, py, px] = 0; A[2][pz, py, px] = 0; A[3][pz, py, px] = 0
    A[1][pz, py, px - 1] = 0; A[2][pz, py - 1, px] = 0
    if is3D:
        A[3][pz - 1, py, px] = 0
    typ[pz, py, px] = 2
    return A, typ


def row_sum(A, b, typ, bscale, x, is3D):
    """b - sum of the off-diagonal terms in the reference's order (-x, +x, -y, +y, -z, +z), for every vertex"""
    s = b.copy()
    if bscale != 0:
        s[typ == 2] = s[typ == 2] * b.dtype.type(bscale)
    s[:, :, 1:] = s[:, :, 1:] - A[1][:, :, :-1] * x[:, :, :-1]
    s[:, :, :-1] = s[:, :, :-1] - A[1][:, :, :-1] * x[:, :, 1:]
    s[:, 1:, :] = s[:, 1:, :] - A[2][:, :-1, :] * x[:, :-1, :]
    s[:, :-1, :] = s[:, :-1, :] - A[2][:, :-1, :] * x[:, 1:, :]
    if is3D:
        s[1:] = s[1:] - A[3][:-1] * x[:-1]
        s[:-1] = s[:-1] - A[3][:-1] * x[1:]
    return s


def colour_mask(shape, c):
    z, y, x = np.indices(shape)
    return ((x + y + z + c) & 1) == 0


def smooth(A, b, typ, bscale, x, c, is3D):
    s = row_sum(A, b, typ, bscale, x, is3D)
    sel = (typ != 0) & colour_mask(x.shape, c)
    out = x.copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        out[sel] = (s / A[0])[sel]
    return out


def residual(A, b, typ, bscale, x, is3D):
    s = row_sum(A, b, typ, bscale, x, is3D)
    s = s - A[0] * x
    return np.where(typ != 0, s, b.dtype.type(0))


def run_emul(lib, mode, A, typ, b, bscale, xin, kchunk, c_first, c_second, order, is3D):
    dtype = b.dtype.type
    sz, sy, sx = b.shape
    suffix = "f32" if dtype is np.float32 else "f64"
    mask = np.zeros(b.shape, np.uint16)
    P = ctypes.c_void_p
    bad = getattr(lib, "mgl0_build_mask_" + suffix)(sx, sy, sz, int(is3D), P(A.ctypes.data), P(typ.ctypes.data), P(mask.ctypes.data))
    assert bad == 0
    xout = np.full(b.shape, 7.25, b.dtype); rout = np.full(b.shape, -3.5, b.dtype)
    cfloat = ctypes.c_float if dtype is np.float32 else ctypes.c_double
    fn = getattr(lib, "mgl0_run_" + suffix)
    fn.argtypes = [ctypes.c_int] * 7 + [P, P, cfloat, P, P, P, P, ctypes.c_int]
    A0 = np.ascontiguousarray(A[0])
    xin_ = np.ascontiguousarray(xin) if xin is not None else np.zeros(1, b.dtype)
    fn(mode, sx, sy, sz, kchunk, c_first, c_second, P(A0.ctypes.data), P(b.ctypes.data), cfloat(bscale), P(mask.ctypes.data), P(xin_.ctypes.data),
       P(xout.ctypes.data), P(rout.ctypes.data), order)
    return xout, rout


def bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


CASES = [
    # shape (z, y, x), kchunk
    ((21, 37, 132), 8),
    ((9, 16, 128), 4),
    ((5, 50, 264), 64),
    ((34, 19, 64), 5),
    ((1, 45, 140), 1),      # 2-D
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,kchunk", CASES)
def test_fused_down_equals_colour_sweeps_and_residual(lib, dtype, shape, kchunk):
    is3D = shape[0] > 1
    A, typ = make_level0(shape, dtype, seed=sum(shape), is3D=is3D)
    rng = np.random.default_rng(5)
    b = (rng.standard_normal(shape) * 3).astype(dtype)
    for bscale in (0.0, 1e-6):
        x = np.zeros(shape, dtype)
        x = smooth(A, b, typ, bscale, x, 0, is3D)
        x = smooth(A, b, typ, bscale, x, 1, is3D)
        r = residual(A, b, typ, bscale, x, is3D)
        for order in (0, 1):
            xe, re = run_emul(lib, 0, A, typ, b, bscale, None, kchunk, 0, 0, order, is3D)
            assert np.array_equal(bits(xe), bits(x))
            assert np.array_equal(bits(re), bits(r))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,kchunk", CASES)
@pytest.mark.parametrize("first", [1, 0])
def test_fused_sweep_equals_two_colour_sweeps(lib, dtype, shape, kchunk, first):
    is3D = shape[0] > 1
    A, typ = make_level0(shape, dtype, seed=1 + sum(shape), is3D=is3D)
    rng = np.random.default_rng(6)
    b = (rng.standard_normal(shape) * 3).astype(dtype)
    xin = rng.standard_normal(shape).astype(dtype)
    xin[typ == 0] = 0
    x = smooth(A, b, typ, 1e-6, xin, first, is3D)
    x = smooth(A, b, typ, 1e-6, x, 1 - first, is3D)
    xe, _ = run_emul(lib, 1, A, typ, b, 1e-6, xin, kchunk, first, 1 - first, 0, is3D)
    assert np.array_equal(bits(xe), bits(x))


def test_mask_rejects_face_fractions(lib):
    shape = (6, 10, 16)
    A, typ = make_level0(shape, np.float32, 3, ghost=False)
    mask = np.zeros(shape, np.uint16)
    P = ctypes.c_void_p
    assert lib.mgl0_build_mask_f32(16, 10, 6, 1, P(A.ctypes.data), P(typ.ctypes.data), P(mask.ctypes.data)) == 0
    z, y, x = np.argwhere(A[1] == -1)[0]
    A[1][z, y, x] = np.float32(-0.5)
    assert lib.mgl0_build_mask_f32(16, 10, 6, 1, P(A.ctypes.data), P(typ.ctypes.data), P(mask.ctypes.data)) == 1

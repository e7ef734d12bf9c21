"""CPU (not gpu): the serial coarse-vertex selection of GridMg::setA that libmantapress runs on the host for levels whose parallel
closure leaves undecided vertices (mantaflow_b200/csrc/mp_mg_coarsen.h: per-count stacks with lazy deletion) picks, level by level,
exactly the coarse vertices of genCoarseGrid (multigrid.cpp:520-578) -- compared with the unmodified reference's hierarchy (oracle/_ref)
and with the C restatement on random obstacle fields, liquids with isolated drops (the order-dependent cases), 2-D and ragged sizes.
These inputs do depend on the visiting order: the same code with first-in-first-out stacks differs from the reference on 9 of their 21
levels (checked once while writing the test)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mantaflow_b200 import scenes

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "emul", "mg_coarsen_host.cpp")
    out = os.path.join(HERE, "emul", "_build", "libmg_coarsen_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    hdr = os.path.join(HERE, "..", "mantaflow_b200", "csrc", "mp_mg_coarsen.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    return C.CDLL(out)


def _field(shape, seed, liquid):
    rng = np.random.Generator(np.random.PCG64(seed))
    sz, sy, sx = shape
    flags = scenes.closed_box_flags(sx, sy, sz)
    inner = flags == 1
    flags[inner & (rng.random(shape) < 0.25)] = 2
    if liquid:      # scattered fluid cells in air: many isolated features -> vertices with several free interpolation vertices
        fl = flags == 1
        flags[fl & (rng.random(shape) < 0.6)] = 4
    return flags


@pytest.mark.parametrize("kind", ["port", "reference"])
@pytest.mark.parametrize("shape,liquid", [((20, 24, 28), False), ((21, 19, 33), True), ((1, 40, 44), True), ((36, 36, 36), True), ((9, 50, 17), True)])
def test_selection_equals_reference_hierarchy(lib, kind, shape, liquid, port32, request):
    O = port32 if kind == "port" else request.getfixturevalue("ref32")
    for seed in (1, 2, 3):
        flags = _field(shape, seed * 11 + shape[1], liquid)
        A = O.make_matrix(flags)
        sz, sy, sx = shape
        O.mg_create(sx, sy, sz)
        try:
            O.mg_set_a(*A)
            nlev = O.mg_num_levels()
            types = [O.mg_get("type", l) for l in range(nlev)]
            sizes = [O.mg_level_size(l) for l in range(nlev)]
        finally:
            O.mg_destroy()
        assert nlev >= 2
        for l in range(1, nlev):
            f, c = sizes[l - 1], sizes[l]
            tf = np.ascontiguousarray(types[l - 1], np.int8)
            tc = np.zeros(c[0] * c[1] * c[2], np.int8)
            lib.mg_coarsen_level(f[0], f[1], f[2], c[0], c[1], c[2], int(sz > 1), tf.ctypes.data_as(C.c_void_p), tc.ctypes.data_as(C.c_void_p))
            want = (types[l] != 0).astype(np.int8)
            assert np.array_equal(tc, want), (shape, seed, l, int((tc != want).sum()))

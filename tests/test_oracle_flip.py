"""CPU (not gpu): the FLIP particle <-> grid plugins (plugin/flip.cpp: markFluidCells, gridParticleIndex, unionParticleLevelset, mapPartsToMAC,
mapMACToParts, flipVelocityUpdate -- SURVEY 8f-4) restated in the oracle ahead of their device versions: bit for bit the golden vectors of
the unmodified reference, and the reference itself on another seed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers
from oracle.oracle_api import Oracle
from helpers import FLIP_SCENES, load_golden, run_flip_plugins

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(FLIP_SCENES))
def test_port_reproduces_flip_golden(name, prec, port32, port64):
    g = load_golden("step_" + name, prec)
    out = run_flip_plugins(port32 if prec == 4 else port64, name, prec)
    assert set(out) == set(g)
    for key in out:
        assert np.array_equal(out[key], g[key]), (name, prec, key)
    # the fixtures are not vacuous
    assert (g["mark"] & 1).sum() > 50 and len(g["index_sys"]) > 300 and (g["union"] < 0).sum() > 50 and np.abs(g["map_vel"]).max() > 0.5
    assert not np.array_equal(g["pic"], g["flip"])


@pytest.mark.parametrize("prec", [4, 8])
def test_port_equals_reference_on_another_scene(prec, port32, port64, ref32, ref64, monkeypatch):
    P, R = (port32, ref32) if prec == 4 else (port64, ref64)
    monkeypatch.setitem(helpers.FLIP_SCENES, "other", (9, 11, 13))
    a, b = run_flip_plugins(P, "other", prec), run_flip_plugins(R, "other", prec)
    for key in a:
        assert np.array_equal(a[key], b[key]), key


# ---------------------------------------------------------------- host emulation of the CUDA kernels (mp_particles_cells.cuh)
class FlipEmulation(Oracle):
    """the oracle_api FLIP interface over tests/emul/particles_emul.cpp: the per-cell / per-particle code and the pass sequences the CUDA
    kernels of mantaflow_b200/csrc/mp_particles.cu run, walked by host loops (the build container has no GPU)"""
    kind = "emulation"

    def __init__(self, lib, prec, order, port=None):
        self.lib, self.prec, self.order, self.port = lib, prec, order, port
        self.real = np.float32 if prec == 4 else np.float64

    def _f(self, name, restype=C.c_int):
        if self.port is not None and not hasattr(self.lib, "emu_" + name):
            return self.port._f(name, restype)          # what is not a particle kernel (the projection, the extrapolations) comes from the restatement
        f = getattr(self.lib, "emu_" + name)
        f.restype = restype
        return lambda *a: f(C.c_int(self.prec), C.c_int(self.order), *a)

    def _chk(self, rc):
        assert rc == 0


@pytest.fixture(scope="module")
def parts_emul_lib():
    src = os.path.join(HERE, "emul", "particles_emul.cpp")
    out = os.path.join(HERE, "emul", "_build", "libparticles_emul.so")
    csrc = os.path.join(ROOT, "mantaflow_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("mp_particles_cells.cuh", "mp_liquid_cells.cuh", "mp_common.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
            pytest.skip("cuda_runtime.h not found: the kernel header cannot be compiled for the host emulation")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        # -ffp-contract=off: the kernels are compiled -fmad=false, the reference build has no FMA either
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-w", "-I" + cuda_inc, "-shared", "-fPIC", src, "-o", out])
    return C.CDLL(out)


@pytest.mark.parametrize("order", [0, 1, 2, 3])
@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(FLIP_SCENES))
def test_kernel_emulation_reproduces_flip_golden(name, prec, order, parts_emul_lib):
    """the code the CUDA kernels run, on the host, in four cell / particle orders: bit-identical to the unmodified reference -- including
    mapPartsToMAC, whose faces gather their particles in ascending particle order (the order of the reference's serial scatter)"""
    g = load_golden("step_" + name, prec)
    out = run_flip_plugins(FlipEmulation(parts_emul_lib, prec, order), name, prec)
    for key in g:
        assert np.array_equal(out[key], g[key]), (name, prec, order, key)


@pytest.mark.parametrize("mode", ["0", "2"])
@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(FLIP_SCENES))
def test_kernel_emulation_map_parts_27_way_walk(name, prec, mode, parts_emul_lib, monkeypatch):
    """mapPartsToMAC has three forms on the device: a tree of 3-way merges whose walk looks up per-particle records of the 24 (w, w v) pairs
    (default in 3-D, used by every other test here), the same tree with the MAC weights evaluated in the walk (MP_MAPPARTS=2; what 2-D grids
    take) and the one-kernel 27-way walk (MP_MAPPARTS=0) -- the same particles in the same order: all give the golden bits"""
    monkeypatch.setenv("MP_MAPPARTS", mode)
    g = load_golden("step_" + name, prec)
    out = run_flip_plugins(FlipEmulation(parts_emul_lib, prec, 3), name, prec)
    for key in g:
        assert np.array_equal(out[key], g[key]), (name, prec, key)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("shape", [(6, 5, 7), (1, 9, 530)])
def test_kernel_emulation_equals_port_on_other_scenes(shape, prec, parts_emul_lib, port32, port64, monkeypatch):
    """another seed, a row wider than one block's 4 x 128 cells, and no particles at all"""
    monkeypatch.setitem(helpers.FLIP_SCENES, "other", shape)
    P, E = (port32 if prec == 4 else port64), FlipEmulation(parts_emul_lib, prec, 3)
    a, b = run_flip_plugins(P, "other", prec), run_flip_plugins(E, "other", prec)
    for key in a:
        assert np.array_equal(a[key], b[key]), key
    flags, pos, pflag, ptype, pvel, phiObs = helpers.flip_scene("other", prec)
    none = slice(0, 0)
    for I in (P, E):
        index, isys = I.grid_particle_index(flags.shape, pos[none], pflag[none])
        assert not index.any() and len(isys) == 0
        assert np.array_equal(I.union_particle_levelset(pos[none], index, isys), P.union_particle_levelset(pos[none], index, isys))
        vel, velOld = I.map_parts_to_mac(flags.shape, pos[none], pflag[none], pvel[none])
        assert not vel.any() and not velOld.any()
        assert np.array_equal(I.mark_fluid_cells(flags.copy(), pos[none], pflag[none]), P.mark_fluid_cells(flags.copy(), pos[none], pflag[none]))


# ---------------------------------------------------------------- ParticleSystem::advectInGrid (particle.h:512-536)
@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(FLIP_SCENES))
def test_port_reproduces_advect_golden(name, prec, port32, port64):
    """Euler / RK2 / RK4 through the MAC grid with the four combinations of deleteInObstacle / stopInObstacle, skipNew and an excluded type:
    positions and flags bit for bit"""
    g = load_golden("step_adv_" + name, prec)
    out = helpers.run_advect_cases(port32 if prec == 4 else port64, name, prec)
    assert set(out) == set(g)
    for key in g:
        assert np.array_equal(out[key], g[key]), (name, prec, key)
    # not vacuous: particles are deleted in obstacles, stopped at them (bisection: neither the old nor the integrated position), and moved
    _, _, pos, pflag, _ = helpers.advect_scene(name, prec)
    assert ((g["rk4_tracer_flag"] & 1024) != 0).sum() > ((pflag & 1024) != 0).sum() + 20
    assert (np.abs(g["rk4_flip_pos"] - pos).max(1) > 0.5).sum() > 100
    assert not np.array_equal(g["rk4_flip_pos"], g["rk2_nostop_pos"])


@pytest.mark.parametrize("order", [0, 1, 2])
@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(FLIP_SCENES))
def test_kernel_emulation_reproduces_advect_golden(name, prec, order, parts_emul_lib):
    """the fused per-particle kernel (all Runge-Kutta stages and the clamping in one pass) on the host"""
    g = load_golden("step_adv_" + name, prec)
    out = helpers.run_advect_cases(FlipEmulation(parts_emul_lib, prec, order), name, prec)
    for key in g:
        assert np.array_equal(out[key], g[key]), (name, prec, order, key)


# ---------------------------------------------------------------- the whole main loop of scenes/benchmark_dam.py:100-134
@pytest.mark.parametrize("prec", [4, 8])
def test_port_reproduces_the_benchmark_dam_loop(prec, port32, port64):
    """twelve passes of the reference's FLIP benchmark loop (20 plugin calls per pass, ghost-fluid PcMIC solves included): the restatement ends with the
    reference's particles, types, flags and fields -- bit for bit in float, within the reduction order in double"""
    g = load_golden("step_dam", prec)
    out = helpers.run_dam_loop(port32 if prec == 4 else port64, prec)
    assert np.array_equal(out["iterations"], g["iterations"])
    for key in ("flags", "ptype"):
        assert np.array_equal(out[key], g[key]), key
    for key in ("pos", "pvel", "vel", "phi", "pressure"):
        if prec == 4:
            assert np.array_equal(out[key], g[key]), key
        else:
            assert float(np.abs(out[key] - g[key]).max()) <= 1e-12, key
    # not vacuous: the dam has collapsed along x, particles changed type in both directions
    assert g["pos"][:, 0].max() > 10 and 0 < (g["ptype"] == 4).sum() < 60


@pytest.mark.parametrize("order", [0, 3])
def test_kernel_emulation_runs_the_benchmark_dam_loop(order, parts_emul_lib, port32):
    """the same twelve passes with every particle kernel taken from the device code (host emulation) and the grid plugins from the restatement:
    still the reference's bits at the end"""
    g = load_golden("step_dam", 4)
    out = helpers.run_dam_loop(FlipEmulation(parts_emul_lib, 4, order, port=port32), 4)
    for key in g:
        assert np.array_equal(out[key], g[key]), key


@pytest.mark.parametrize("prec", [4, 8])
def test_vec_max_abs_port_equals_reference(prec, port32, port64, ref32, ref64):
    """Grid<Vec3>::getMaxAbs (grid.cpp:330-332), what the liquid scenes hand to adaptTimestep"""
    P, R = (port32, ref32) if prec == 4 else (port64, ref64)
    vel = load_golden("step_dam", prec)["vel"]
    assert P.vec_max_abs(vel) == R.vec_max_abs(vel) > 1.0


# ---------------------------------------------------------------- element-wise Grid<T> arithmetic (grid.cpp:258-284)
@pytest.mark.parametrize("comps", [1, 3])
@pytest.mark.parametrize("dtype", [np.int32, np.float32, np.float64])
@pytest.mark.parametrize("op", helpers.GRID_OPS)
def test_kernel_emulation_grid_arithmetic(op, dtype, comps, parts_emul_lib):
    """the element functor of mp_gridops.cuh (what k_grid_arith runs) against the numpy restatement of the reference's one-line kernels"""
    if dtype == np.int32 and comps == 3:
        pytest.skip("there are no Vec3i grids")
    me, other, c, want = helpers.grid_arith_case(op, dtype, comps)
    got = me.copy()
    f = parts_emul_lib.emu_grid_arith
    rc = f(C.c_int(0 if dtype == np.int32 else np.dtype(dtype).itemsize), C.c_int(2), C.c_longlong(me.size // comps), C.c_int(comps), got.ctypes.data_as(C.c_void_p),
           C.c_int(helpers.GRID_OPS.index(op)), None if other is None else other.ctypes.data_as(C.c_void_p), C.c_double(c[0]), C.c_double(c[1]), C.c_double(c[2]))
    assert rc == 0 and np.array_equal(got, want), (op, dtype, comps)


@pytest.mark.parametrize("trial", range(12))
def test_kernel_emulation_fuzz_against_port(trial, parts_emul_lib, port32, port64):
    """random tiny grids (down to two cells per axis, 2-D and 3-D), particles outside the domain on every side, on cell faces and centres exactly, deleted and
    excluded ones, dozens per cell: the device code equals the restatement bit for bit in three cell / particle orders"""
    rng = np.random.default_rng(100 + trial)
    prec = 4 if trial % 2 == 0 else 8
    real = np.float32 if prec == 4 else np.float64
    shape = (1, int(rng.integers(3, 9)), int(rng.integers(3, 9))) if trial % 3 == 0 else tuple(int(v) for v in rng.integers(2, 7, 3))
    sz, sy, sx = shape
    n = int(rng.integers(1, 600))
    pos = (rng.random((n, 3)) * (np.array([sx, sy, sz]) + 2) - 1).astype(real)
    snap = rng.random(n) < 0.2
    pos[snap] = np.round(pos[snap] * 2) / 2
    if sz == 1:
        pos[:, 2] = 0.5
    pflag = np.where(rng.random(n) < 0.1, 1024, 0).astype(np.int32)
    ptype = np.where(rng.random(n) < 0.2, 4, 1).astype(np.int32)
    pvel = (rng.random((n, 3)) * 2 - 1).astype(real)
    if sz == 1:
        pvel[:, 2] = 0
    P = port32 if prec == 4 else port64
    a = P.map_parts_to_mac(shape, pos, pflag, pvel, want_weight=True, ptype=ptype, exclude=4)
    ia, sa = P.grid_particle_index(shape, pos, pflag)
    ua = P.union_particle_levelset(pos, ia, sa, 1.3, ptype=ptype, exclude=4)
    fa = P.flip_velocity_update(a[0], a[1], pos, pflag, pvel.copy(), 0.9, ptype=ptype, exclude=4)
    for order in (0, 2, 3):
        E = FlipEmulation(parts_emul_lib, prec, order)
        b = E.map_parts_to_mac(shape, pos, pflag, pvel, want_weight=True, ptype=ptype, exclude=4)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), (shape, n, order)
        ib, sb = E.grid_particle_index(shape, pos, pflag)
        assert np.array_equal(ia, ib) and np.array_equal(sa, sb)
        assert np.array_equal(ua, E.union_particle_levelset(pos, ib, sb, 1.3, ptype=ptype, exclude=4))
        assert np.array_equal(fa, E.flip_velocity_update(a[0], a[1], pos, pflag, pvel.copy(), 0.9, ptype=ptype, exclude=4))


@pytest.mark.parametrize("trial", range(10))
def test_kernel_emulation_fuzz_particle_movers(trial, parts_emul_lib, port32, port64):
    """random obstacle cells inside small boxes, fast velocity fields, every integration mode x obstacle policy, new / deleted / excluded particles:
    advectInGrid, pushOutofObs, setPartType, markIsolatedFluidCell and markFluidCells(phiObs) of the device code equal the restatement bit for bit"""
    from mantaflow_b200 import scenes
    rng = np.random.default_rng(500 + trial)
    prec = 4 if trial % 2 == 0 else 8
    real = np.float32 if prec == 4 else np.float64
    sx, sy, sz = (int(rng.integers(4, 12)), int(rng.integers(4, 12)), 1) if trial % 3 == 0 else tuple(int(v) for v in rng.integers(4, 10, 3))
    flags = scenes.closed_box_flags(sx, sy, sz)
    flags[((flags & 2) == 0) & (rng.random(flags.shape) < 0.15)] = 2
    free = (flags & 2) == 0
    flags[free] = np.where(rng.random(flags.shape) < 0.5, 1, 4)[free]
    n, nd = int(rng.integers(1, 400)), (3 if sz > 1 else 2)
    pos = np.full((n, 3), 0.5)
    pos[:, :nd] = 1 + rng.random((n, nd)) * (np.array([sx, sy, sz])[:nd] - 2)
    pos = pos.astype(real)
    pflag = (np.where(rng.random(n) < 0.1, 1024, 0) | np.where(rng.random(n) < 0.2, 1, 0)).astype(np.int32)
    ptype = np.where(rng.random(n) < 0.2, 4, 1).astype(np.int32)
    vel = ((rng.random(flags.shape + (3,)) - 0.5) * 8).astype(real)
    if sz == 1:
        vel[..., 2] = 0
    phiObs = (rng.random(flags.shape) * 3 - 1).astype(real)
    P, E = (port32 if prec == 4 else port64), FlipEmulation(parts_emul_lib, prec, trial % 3)
    for mode in (0, 1, 2):
        for dele in (False, True):
            for stop in (False, True):
                kw = dict(integrationMode=mode, deleteInObstacle=dele, stopInObstacle=stop, skipNew=bool(trial & 1), ptype=ptype if trial % 4 else None, exclude=4)
                a, b = P.advect_in_grid(flags, vel, pos.copy(), pflag.copy(), 0.7, **kw), E.advect_in_grid(flags, vel, pos.copy(), pflag.copy(), 0.7, **kw)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (mode, dele, stop)
    moved = P.advect_in_grid(flags, vel, pos.copy(), pflag.copy(), 0.7, integrationMode=1, deleteInObstacle=False, stopInObstacle=False)[0]
    assert np.array_equal(P.push_out_of_obs(flags.shape, moved.copy(), pflag, phiObs, shift=0.1, thresh=0.4, ptype=ptype, exclude=4),
                          E.push_out_of_obs(flags.shape, moved.copy(), pflag, phiObs, shift=0.1, thresh=0.4, ptype=ptype, exclude=4))
    assert np.array_equal(P.set_part_type(flags, moved, ptype.copy(), 1, 4, 1), E.set_part_type(flags, moved, ptype.copy(), 1, 4, 1))
    assert np.array_equal(P.mark_isolated_fluid_cell(flags.copy(), 4), E.mark_isolated_fluid_cell(flags.copy(), 4))
    assert np.array_equal(P.mark_fluid_cells(flags.copy(), moved, pflag, phiObs=phiObs, ptype=ptype, exclude=4),
                          E.mark_fluid_cells(flags.copy(), moved, pflag, phiObs=phiObs, ptype=ptype, exclude=4))

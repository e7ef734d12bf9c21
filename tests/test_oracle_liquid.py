"""CPU (not gpu): the liquid neighbours of the projection (first slice: extrapolateMACSimple, extrapolateLsSimple,
extrapolateVec3Simple, FlagGrid::updateFromLevelset, Grid::setBound).

* the C restatement reproduces the golden vectors of the unmodified reference bit for bit, and the reference itself on another seed;
* the host emulation of the CUDA kernels (tests/emul/liquid_emul.cpp: the per-cell operations and pass sequences of
  mantaflow_b200/csrc/mp_liquid_cells.cuh, walked by a host loop instead of one thread per cell) does the same for every walk order --
  the build container has no GPU, so this is how the kernels' arithmetic and pass structure are checked here; the launch itself is
  covered by tests/test_gpu_liquid.py on the B200."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import (SECORDER_SCENES, check_sec_order_bnd_against_golden, FREESURFACE_SCENES, LIQUID_CASES, LIQUID_SCENES, check_freesurface_against_golden, check_liquid_against_golden, liquid_scene,
                     run_liquid_case)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(LIQUID_SCENES))
def test_port_reproduces_liquid_golden(name, prec, port32, port64):
    check_liquid_against_golden(port32 if prec == 4 else port64, name, prec)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(FREESURFACE_SCENES))
def test_port_reproduces_freesurface_steps(name, prec, port32, port64):
    """scenes/freesurface.py:54-84, six steps: float 3-D (PcMIC) bit-identical; the multigrid solves of the 2-D case and the double build
    agree up to the reduction order"""
    tol = {("fs3d", 4): 0.0, ("fs2d", 4): 1e-3, ("fs3d", 8): 1e-11, ("fs2d", 8): 1e-10}[(name, prec)]
    check_freesurface_against_golden(port32 if prec == 4 else port64, name, prec, tol)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(SECORDER_SCENES))
def test_port_reproduces_second_order_boundary_scenario(name, prec, port32, port64):
    """tools/tests/test_1040_secOrderBnd.py: updateFractions, setObstacleFlags, setWallBcs(fractions, phiObs), extrapolateMACSimple and the
    projection with fractions; float 3-D (PcMIC) bit-identical, the rest up to the reduction order"""
    tol = {("sob3d", 4): 0.0, ("sob2d", 4): 1e-4, ("sob3d", 8): 1e-10, ("sob2d", 8): 1e-10}[(name, prec)]
    check_sec_order_bnd_against_golden(port32 if prec == 4 else port64, name, prec, tol)


@pytest.mark.parametrize("prec", [4, 8])
def test_port_equals_reference_on_another_seed(prec, port32, port64, ref32, ref64):
    P, R = (port32, ref32) if prec == 4 else (port64, ref64)
    flags, vel, phi, phiObs = liquid_scene("liqragged", prec)
    flags = np.ascontiguousarray(flags[::-1]); vel = np.ascontiguousarray(vel[:, ::-1]); phi = np.ascontiguousarray(phi[::-1]); phiObs = np.ascontiguousarray(phiObs[:, :, ::-1])
    for case in LIQUID_CASES:
        assert np.array_equal(run_liquid_case(P, case, flags, vel, phi, phiObs), run_liquid_case(R, case, flags, vel, phi, phiObs)), case


# ---------------------------------------------------------------- host emulation of the CUDA kernels
class Emulation:
    """the oracle_api interface over tests/emul/liquid_emul.cpp"""
    kind = "emulation"

    def __init__(self, lib, prec, order):
        self.lib, self.prec, self.order = lib, prec, order
        self.real = np.float32 if prec == 4 else np.float64

    def _head(self, a):
        sz, sy, sx = a.shape[:3]
        return C.c_int(self.prec), C.c_int(self.order), C.c_int(sx), C.c_int(sy), C.c_int(sz)

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def extrapolate_mac_simple(self, flags, vel, distance=4, phiObs=None, intoObs=False):
        assert self.lib.emu_extrapolate_mac_simple(*self._head(flags), self._p(flags), self._p(vel), C.c_int(distance), self._p(phiObs), C.c_int(int(intoObs))) == 0
        return vel

    def extrapolate_mac_from_weight(self, vel, weight, distance=2):
        assert self.lib.emu_extrapolate_mac_from_weight(*self._head(vel), self._p(vel), self._p(weight), C.c_int(distance)) == 0
        return vel

    def extrapolate_ls_simple(self, phi, distance=4, inside=False):
        assert self.lib.emu_extrapolate_ls_simple(*self._head(phi), self._p(phi), C.c_int(distance), C.c_int(int(inside))) == 0
        return phi

    def extrapolate_vec3_simple(self, vel, phi, distance=4, inside=False):
        assert self.lib.emu_extrapolate_vec3_simple(*self._head(phi), self._p(vel), self._p(phi), C.c_int(distance), C.c_int(int(inside))) == 0
        return vel

    def update_from_levelset(self, flags, phi):
        assert self.lib.emu_update_from_levelset(*self._head(flags), self._p(flags), self._p(phi)) == 0
        return flags

    def set_wall_bcs_frac(self, flags, vel, phiObs):
        assert self.lib.emu_set_wall_bcs_frac(*self._head(flags), self._p(flags), self._p(vel), self._p(phiObs)) == 0
        return vel

    def update_fractions(self, flags, phiObs, boundaryWidth=0, fracThreshold=0.01):
        fr = np.full(flags.shape + (3,), 7.0, self.real)       # every entry is written
        assert self.lib.emu_update_fractions(*self._head(flags), self._p(flags), self._p(phiObs), self._p(fr), C.c_int(boundaryWidth), C.c_double(fracThreshold)) == 0
        return fr

    def set_obstacle_flags(self, flags, phiObs, fractions=None, phiOut=None, phiIn=None, boundaryWidth=1):
        assert self.lib.emu_set_obstacle_flags(*self._head(flags), self._p(flags), self._p(phiObs), self._p(fractions), self._p(phiOut), self._p(phiIn), C.c_int(boundaryWidth)) == 0
        return flags

    def get_laplacian(self, grid):
        out = np.zeros_like(grid)
        assert self.lib.emu_stencil(*self._head(grid), self._p(out), self._p(grid), C.c_double(1.0), C.c_int(0)) == 0
        return out

    def get_curvature(self, grid, h=1.0):
        out = np.zeros_like(grid)
        assert self.lib.emu_stencil(*self._head(grid), self._p(out), self._p(grid), C.c_double(h), C.c_int(1)) == 0
        return out

    def set_bound(self, grid, value, boundaryWidth=1):
        assert self.lib.emu_set_bound(*self._head(grid), self._p(grid), C.c_int(1 if grid.ndim == 3 else 3), C.c_double(value), C.c_int(boundaryWidth)) == 0
        return grid


@pytest.fixture(scope="module")
def emul_lib():
    src = os.path.join(HERE, "emul", "liquid_emul.cpp")
    out = os.path.join(HERE, "emul", "_build", "libliquid_emul.so")
    deps = [src, os.path.join(ROOT, "mantaflow_b200", "csrc", "mp_liquid_cells.cuh"), os.path.join(ROOT, "mantaflow_b200", "csrc", "mp_common.cuh")]
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
@pytest.mark.parametrize("name", list(LIQUID_SCENES))
def test_kernel_emulation_reproduces_liquid_golden(name, prec, order, emul_lib):
    """the code the CUDA kernels run, cell by cell on the host, in four different cell orders (3 = block by block, thread by thread as the CUDA launch maps them): bit-identical to the reference"""
    check_liquid_against_golden(Emulation(emul_lib, prec, order), name, prec)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("shape", [(5, 7, 1100), (1, 9, 530)])
def test_kernel_emulation_launch_geometry_on_wide_rows(shape, prec, emul_lib, port32, port64):
    """rows wider than one block's 4 x 128 cells and row counts that are no multiple of 4: the CUDA launch's cell mapping covers every cell"""
    import helpers
    helpers.LIQUID_SCENES["wide"] = shape
    try:
        flags, vel, phi, phiObs = liquid_scene("wide", prec)
    finally:
        del helpers.LIQUID_SCENES["wide"]
    O, E = (port32 if prec == 4 else port64), Emulation(emul_lib, prec, 3)
    for case in LIQUID_CASES:
        assert np.array_equal(run_liquid_case(E, case, flags, vel, phi, phiObs), run_liquid_case(O, case, flags, vel, phi, phiObs)), case

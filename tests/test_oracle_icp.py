"""CPU (not gpu): the IC(0) preconditioner PC_ICP (InitPreconditionIncompCholesky / ApplyPreconditionIncompCholesky, conjugategrad.cpp:26-63,
:109-132 -- the preconditioner of the VIC Poisson solve).

* the C restatement reproduces the golden vectors of the unmodified reference (factor grids and sweeps bit for bit, GridCg with PC_ICP to the
  iteration), and the reference itself on another system;
* the host emulation of the CUDA kernels (tests/emul/ic_emul.cpp: the per-cell code and hyperplane geometry of mp_ic_cells.cuh) gives the same
  bits in both in-plane orders."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers
from helpers import ICP_SCENES, check_icp_against_golden, load_golden
from oracle.oracle_api import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", ICP_SCENES)
def test_port_reproduces_icp_golden(name, prec, port32, port64):
    check_icp_against_golden(port32 if prec == 4 else port64, name, prec, exact_reductions=False)


@pytest.mark.parametrize("prec", [4, 8])
def test_port_equals_reference_on_another_system(prec, port32, port64, ref32, ref64):
    """the diffusion matrix I + alpha L with obstacle rows (cgSolveDiffusion's system): off-diagonals -alpha, obstacles inside the domain"""
    P, R = (port32, ref32) if prec == 4 else (port64, ref64)
    k = load_golden("smoke16", prec)
    flags = k["flags"]
    real = np.float32 if prec == 4 else np.float64
    rng = np.random.default_rng(3)
    fluid = (flags & 1) != 0
    A0 = np.where(fluid, 1 + 6 * 0.3, 1.0).astype(real) + (rng.random(flags.shape) * 0.1).astype(real)
    Ai, Aj, Ak = [np.where(fluid, -0.3, 0).astype(real) * (rng.random(flags.shape) < 0.9) for _ in range(3)]
    Ai, Aj, Ak = [np.ascontiguousarray(a.astype(real)) for a in (Ai, Aj, Ak)]
    src = (rng.random(flags.shape) - 0.5).astype(real)
    Pp, Pr = P.ic_init(flags, A0, Ai, Aj, Ak), R.ic_init(flags, A0, Ai, Aj, Ak)
    for a, b in zip(Pp, Pr):
        assert np.array_equal(a, b)
    assert np.array_equal(P.ic_apply(flags, src, *Pp)[fluid], R.ic_apply(flags, src, *Pr)[fluid])
    xp, itp, _ = P.cg_solve(flags, src * fluid, A0, Ai, Aj, Ak, pc=3, accuracy=1e-5 if prec == 4 else 1e-11, maxIter=500)
    xr, itr, _ = R.cg_solve(flags, src * fluid, A0, Ai, Aj, Ak, pc=3, accuracy=1e-5 if prec == 4 else 1e-11, maxIter=500)
    assert abs(itp - itr) <= 1 and helpers.rel_l2(xp, xr) <= (1e-4 if prec == 4 else 1e-10)


class IcEmulation(Oracle):
    """ic_init / ic_apply over tests/emul/ic_emul.cpp; everything else (the CG loop around them) from the restatement"""
    kind = "emulation"

    def __init__(self, lib, port, order):
        self.emu, self.port, self.order = lib, port, order
        self.prec, self.real, self.lib, self.pfx = port.prec, port.real, port.lib, port.pfx

    def _f(self, name, restype=C.c_int):
        if name in ("ic_init", "ic_apply"):
            f = getattr(self.emu, "emu_" + name)
            f.restype = restype
            return lambda *a: f(C.c_int(self.prec), C.c_int(self.order), *a)
        return Oracle._f(self, name, restype)


@pytest.fixture(scope="module")
def ic_emul_lib():
    src = os.path.join(HERE, "emul", "ic_emul.cpp")
    out = os.path.join(HERE, "emul", "_build", "libic_emul.so")
    csrc = os.path.join(ROOT, "mantaflow_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("mp_ic_cells.cuh", "mp_common.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
            pytest.skip("cuda_runtime.h not found: the kernel header cannot be compiled for the host emulation")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-w", "-I" + cuda_inc, "-shared", "-fPIC", src, "-o", out])
    return C.CDLL(out)


@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", ICP_SCENES)
def test_kernel_emulation_reproduces_icp_golden(name, prec, order, ic_emul_lib, port32, port64):
    """the gather form of the factorisation and the hyperplane schedule give the reference's bits, whatever the order inside a plane"""
    g, k = load_golden("icp_" + name, prec), load_golden(name, prec)
    E = IcEmulation(ic_emul_lib, port32 if prec == 4 else port64, order)
    flags, A = k["flags"], [k[n] for n in "A0 Ai Aj Ak".split()]
    P = E.ic_init(flags, *A)
    for n, p in zip("0ijk", P):
        assert np.array_equal(p, g["ic_P" + n]), n
    fluid = (flags & 1) != 0
    assert np.array_equal(E.ic_apply(flags, k["src"], *P)[fluid], g["ic_apply"][fluid])

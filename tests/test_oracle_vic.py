"""CPU (not gpu): the grid half of VICintegration (plugin/vortexplugins.cpp:253-299 -- the VIC Poisson solve): MakeLaplaceMatrix,
CurlOp, per component Get(Shifted)Component, GridCg<ApplyMatrix> with PC_ICP / PC_mICP and the L2 stop test, scaling, SetComponent.

* the golden vectors come from the plugin of the UNMODIFIED reference run on a triangle sheet (tests/golden/make_golden.py --only-vic; the
  Peskin mapping of the mesh onto the vorticity grid is the reference's own and is part of the fixture);
* the C restatement reproduces them from the fixture's vorticity grid, and equals the reference bit for bit in float when that is built here."""
import numpy as np
import pytest

import helpers
from oracle.oracle_api import OracleError


@pytest.mark.parametrize("prec", [4, 8])
def test_port_reproduces_vic_golden(prec, port32, port64):
    helpers.check_vic_against_golden(port32 if prec == 4 else port64, prec, exact_reductions=(prec == 4))


@pytest.mark.parametrize("prec", [4, 8])
def test_port_equals_reference_on_the_scene(prec, port32, port64, ref32, ref64):
    """the fixture is what the reference computes today (no stale file), and the restatement follows it to the bit in float"""
    P, R = (port32, ref32) if prec == 4 else (port64, ref64)
    fx = helpers.run_vic_reference(R, prec)
    g = helpers.load_golden("vic_sheet24", prec)
    assert np.array_equal(fx["vorticity"], g["vorticity"])
    for mac, pc in helpers.VIC_CASES:
        tag = "%s_pc%d" % ("mac" if mac else "vec", pc)
        vel, its = P.vic_poisson(fx["flags"], fx["vorticity"], np.zeros_like(fx["vorticity"]), velIsMac=mac, cgMaxIterFac=5, cgAccuracy=helpers.vic_accuracy(prec),
                                 scale=0.01, precondition=pc)
        assert its == list(fx["its_" + tag])
        if prec == 4:
            assert np.array_equal(vel, fx["vel_" + tag]) and np.array_equal(fx["vel_" + tag], g["vel_" + tag])
        else:
            assert helpers.rel_l2(vel, fx["vel_" + tag]) < 1e-13


def test_default_precondition_is_an_error_as_in_the_reference(port32, ref32):
    """VICintegration's default precondition = 0 reaches setICPreconditioner(PC_None), which asserts (conjugategrad.cpp:312)"""
    flags, vel0, tri, tv = helpers.vic_scene(4)
    with pytest.raises(OracleError, match="setICPreconditioner: Invalid method"):
        ref32.vic_integration(flags, tri, tv, 2.0, vel0, precondition=0)
    with pytest.raises(OracleError, match="setICPreconditioner: Invalid method"):
        port32.vic_poisson(flags, np.zeros(flags.shape + (3,), np.float32), vel0, precondition=0)

"""CPU (not gpu): the restatement of the steps either side of the projection (setWallBcs, addGravity, addBuoyancy,
advectSemiLagrange) reproduces the golden vectors of the unmodified reference bit for bit, and the reference itself where
oracle/_ref is built."""
import numpy as np
import pytest

from helpers import (GUIDING_SCENES, STEP_CASES, STEP_SCENES, WE_SCENES, check_guiding_against_golden, check_step_against_golden, check_waves_against_golden, load_golden, run_plume_steps,
                     run_step_case, step_scene)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(STEP_SCENES))
def test_port_reproduces_step_golden(name, prec, port32, port64):
    check_step_against_golden(port32 if prec == 4 else port64, name, prec)


@pytest.mark.parametrize("tag,shape", [("3d", (24, 36, 24)), ("2d", (1, 48, 32))])
def test_port_reproduces_plume_steps(tag, shape, port32):
    g = load_golden("plume" + tag, 4)
    dens, vel, p, its = run_plume_steps(port32, shape, 4, steps=6)
    assert its == [int(v) for v in g["iterations"]]
    assert np.array_equal(dens, g["density"]) and np.array_equal(vel, g["vel"]) and np.array_equal(p, g["pressure"])


@pytest.mark.parametrize("prec", [4, 8])
def test_port_equals_reference_on_another_seed(prec, port32, port64, ref32, ref64):
    """not only the committed vectors: a scene the goldens do not contain, run through both"""
    P, R = (port32, ref32) if prec == 4 else (port64, ref64)
    flags, vel, dens, obvel = step_scene("ragged3d", prec)
    vel = np.ascontiguousarray(vel[::-1].copy()); dens = np.ascontiguousarray(dens[:, ::-1].copy())
    for case in STEP_CASES:
        assert np.array_equal(run_step_case(P, case, flags, vel, dens, obvel), run_step_case(R, case, flags, vel, dens, obvel)), case


def test_unsupported_orders_are_errors(port32):
    from oracle.oracle_api import OracleError
    flags, vel, dens, _ = step_scene("box2d", 4)
    with pytest.raises(OracleError):
        port32.advect_semi_lagrange(flags, vel, dens.copy(), order=3)
    with pytest.raises(OracleError):
        port32.advect_semi_lagrange(flags, vel, dens.copy(), order=1, orderSpace=3)
    with pytest.raises(OracleError):
        port32.advect_semi_lagrange(flags, vel, dens.copy(), order=1, orderTrace=3)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", ["box3d", "box2d"])
def test_port_reproduces_higher_order_advection_golden(name, prec, port32, port64):
    """orderSpace 2 (cubic lookups) / orderTrace 2 (explicit midpoint): the restatement against the reference's output, bit for bit"""
    from helpers import check_step_hi_against_golden
    check_step_hi_against_golden(port32 if prec == 4 else port64, name, prec)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(WE_SCENES))
def test_port_reproduces_wave_equation_golden(name, prec, port32, port64):
    """cgSolveWE: float bit-identical (double accumulators of float products), double up to the reduction order"""
    check_waves_against_golden(port32 if prec == 4 else port64, name, prec, tol=0.0 if prec == 4 else 1e-13)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(GUIDING_SCENES))
def test_port_reproduces_fluid_guiding_golden(name, prec, port32, port64):
    """PD_fluid_guiding: ~22 primal-dual iterations, each with blurs and a multigrid solve; float bit-identical"""
    check_guiding_against_golden(port32 if prec == 4 else port64, name, prec, tol=0.0 if prec == 4 else 1e-11)

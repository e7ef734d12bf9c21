"""GPU: the device plugins either side of the projection -- setWallBcs, addGravity, addBuoyancy, advectSemiLagrange --
bit for bit against the reference's golden vectors and against the oracle, and six steps of the simpleplume main loop with every
field resident on the device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from helpers import (GUIDING_SCENES, STEP_CASES, STEP_SCENES, WE_SCENES, check_guiding_against_golden, check_step_against_golden,  # noqa: E402
                     check_waves_against_golden, load_golden,
                     run_plume_steps, run_step_case, step_scene)


@pytest.fixture(scope="module")
def mf():
    import mantaflow_b200 as m
    if m.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests need a B200")
    return m


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(STEP_SCENES))
def test_cuda_reproduces_step_golden(name, prec):
    from cuda_impl import CudaImpl
    check_step_against_golden(CudaImpl(prec), name, prec)


@pytest.mark.parametrize("prec", [4, 8])
def test_cuda_equals_oracle_on_a_larger_scene(prec):
    from cuda_impl import CudaImpl
    from oracle.oracle_api import Oracle
    import helpers
    helpers.STEP_SCENES["large"] = (40, 52, 64)
    try:
        flags, vel, dens, obvel = step_scene("large", prec)
    finally:
        del helpers.STEP_SCENES["large"]
    O, I = Oracle("port", prec), CudaImpl(prec)
    for case in STEP_CASES:
        assert np.array_equal(run_step_case(I, case, flags, vel, dens, obvel), run_step_case(O, case, flags, vel, dens, obvel)), case


@pytest.mark.parametrize("tag,shape", [("3d", (24, 36, 24)), ("2d", (1, 48, 32))])
def test_plume_steps_match_the_reference(tag, shape):
    """scenes/simpleplume.py:48-60 through the adapter (host arrays in, host arrays out, every plugin on the device)"""
    from cuda_impl import CudaImpl
    g = load_golden("plume" + tag, 4)
    dens, vel, p, its = run_plume_steps(CudaImpl(4), shape, 4, steps=6)
    assert its == [int(v) for v in g["iterations"]]
    assert np.array_equal(dens, g["density"]) and np.array_equal(vel, g["vel"]) and np.array_equal(p, g["pressure"])


def test_plume_steps_device_resident(mf):
    """the same loop written like the scene: grids are created once and never leave the device until the end"""
    import helpers
    shape = (24, 36, 24)
    g = load_golden("plume3d", 4)
    flags_h, src, real = helpers.plume_scene(shape, 4)
    s = mf.Solver(gridSize=(shape[2], shape[1], shape[0]), dim=3, prec=4)
    flags = mf.FlagGrid(s, flags_h)
    vel, density, pressure = s.create(mf.MACGrid), s.create(mf.RealGrid), s.create(mf.RealGrid)
    launches0 = s.kernelLaunches()
    for _ in range(6):
        d = density.numpy(writable=True); d[src] = 1                      # the scene's densityInflow / applyToGrid (host side here)
        mf.advectSemiLagrange(flags=flags, vel=vel, grid=density, order=2)
        mf.advectSemiLagrange(flags=flags, vel=vel, grid=vel, order=2, strength=1.0)
        mf.setWallBcs(flags=flags, vel=vel)
        mf.addBuoyancy(density=density, vel=vel, gravity=(0, -6e-4, 0), flags=flags)
        mf.solvePressure(flags=flags, vel=vel, pressure=pressure)
    assert s.kernelLaunches() > launches0
    assert np.array_equal(density.numpy(), g["density"]) and np.array_equal(vel.numpy(), g["vel"]) and np.array_equal(pressure.numpy(), g["pressure"])


def test_unsupported_variants_fail_loudly(mf):
    flags_h, vel_h, dens_h, _ = step_scene("box2d", 4)
    s = mf.Solver(gridSize=(30, 24, 1), dim=2, prec=4)
    F, V, D = mf.FlagGrid(s, flags_h), mf.MACGrid(s, vel_h), mf.RealGrid(s, dens_h)
    with pytest.raises(mf.MantaError):
        mf.advectSemiLagrange(F, V, D, order=3)
    with pytest.raises(mf.MantaError):
        mf.advectSemiLagrange(F, V, D, orderSpace=3)
    with pytest.raises(mf.MantaError):
        mf.advectSemiLagrange(F, V, D, orderTrace=3)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", ["box3d", "box2d"])
def test_cuda_reproduces_higher_order_advection_golden(name, prec):
    """advectSemiLagrange with orderSpace 2 (cubic lookups, util/interpolHigh.h) / orderTrace 2 (explicit midpoint, advection.cpp:32-37, :58-73),
    Real and MAC grids, plain and MacCormack: bit-identical to the reference's output"""
    from cuda_impl import CudaImpl
    from helpers import check_step_hi_against_golden
    check_step_hi_against_golden(CudaImpl(prec), name, prec)


@pytest.mark.parametrize("prec", [4, 8])
def test_cuda_higher_order_advection_equals_oracle_on_a_larger_scene(prec):
    from cuda_impl import CudaImpl
    from oracle.oracle_api import Oracle
    import helpers
    helpers.STEP_SCENES["large"] = (40, 52, 64)
    try:
        flags, vel, dens, obvel = step_scene("large", prec)
    finally:
        del helpers.STEP_SCENES["large"]
    O, I = Oracle("port", prec), CudaImpl(prec)
    for case in helpers.STEP_HI_CASES:
        assert np.array_equal(helpers.run_step_hi_case(I, case, flags, vel, dens, obvel), helpers.run_step_hi_case(O, case, flags, vel, dens, obvel)), case


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(WE_SCENES))
def test_cuda_reproduces_wave_equation_golden(name, prec):
    """cgSolveWE on the device GridCg: float bit-identical to the reference, double within the reduction order"""
    from cuda_impl import CudaImpl
    check_waves_against_golden(CudaImpl(prec), name, prec, tol=0.0 if prec == 4 else 1e-12)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(GUIDING_SCENES))
def test_cuda_reproduces_fluid_guiding_golden(name, prec):
    """PD_fluid_guiding on the device: same loop count, float within the multigrid solves' tolerance of the reference"""
    from cuda_impl import CudaImpl
    check_guiding_against_golden(CudaImpl(prec), name, prec, tol=2e-5 if prec == 4 else 1e-9)

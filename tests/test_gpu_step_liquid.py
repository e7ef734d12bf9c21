"""GPU: the liquid neighbours of the projection on the device (first slice) -- extrapolateMACSimple, extrapolateLsSimple,
extrapolateVec3Simple, FlagGrid.updateFromLevelset, Grid.setBound -- bit for bit against the reference's golden vectors and against the
oracle on larger grids, and six steps of the level-set free-surface loop of scenes/freesurface.py:54-84 with every field resident on
the device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import helpers  # noqa: E402
from helpers import (SECORDER_SCENES, check_sec_order_bnd_against_golden, FREESURFACE_SCENES, LIQUID_CASES, LIQUID_SCENES, check_freesurface_against_golden, check_liquid_against_golden,  # noqa: E402
                     liquid_scene, load_golden, run_liquid_case)


@pytest.fixture(scope="module")
def mf():
    import mantaflow_b200 as m
    if m.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests need a B200")
    return m


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(LIQUID_SCENES))
def test_cuda_reproduces_liquid_golden(name, prec):
    from cuda_impl import CudaImpl
    check_liquid_against_golden(CudaImpl(prec), name, prec)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("shape", [(34, 44, 150), (1, 70, 260)])
def test_cuda_equals_oracle_on_larger_scenes(shape, prec):
    """rows longer than one thread block (128 cells) and many blocks in y / z"""
    from cuda_impl import CudaImpl
    from oracle.oracle_api import Oracle
    helpers.LIQUID_SCENES["large"] = shape
    try:
        flags, vel, phi, phiObs = liquid_scene("large", prec)
    finally:
        del helpers.LIQUID_SCENES["large"]
    O, I = Oracle("port", prec), CudaImpl(prec)
    for case in LIQUID_CASES:
        a, b = run_liquid_case(I, case, flags, vel, phi, phiObs), run_liquid_case(O, case, flags, vel, phi, phiObs)
        if case in helpers.LIQUID_ULP_CASES:          # pow() in double: last-bit differences between libm and the device
            assert np.allclose(a, b, rtol=3e-7 if prec == 4 else 4e-15, atol=0), case
        else:
            assert np.array_equal(a, b), case


@pytest.mark.parametrize("prec", [4, 8])
def test_extrapolation_schedules_agree(prec, monkeypatch):
    """The extrapolation plugins run their passes 2 .. distance on lists of the cells the pass before marked (the frontier schedule of
    mp_liquid.cu) instead of sweeping the grid per pass: the same per-cell code on the only cells a pass can change.  Both schedules give
    the same grids bit for bit (3-D with rows longer than a block, 2-D, distances 2 .. 9), and both equal the oracle."""
    from cuda_impl import CudaImpl
    from oracle.oracle_api import Oracle
    O, I = Oracle("port", prec), CudaImpl(prec)
    for shape in ((37, 45, 150), (1, 70, 66)):
        helpers.LIQUID_SCENES["sched"] = shape
        try:
            flags, vel, phi, phiObs = liquid_scene("sched", prec)
        finally:
            del helpers.LIQUID_SCENES["sched"]
        for distance in (2, 3, 5, 9):
            res = {}
            for mode in ("1", "0"):
                monkeypatch.setenv("MP_LIQUID_FRONTIER", mode)
                res[mode] = (I.extrapolate_mac_simple(flags, vel.copy(), distance=distance), I.extrapolate_mac_simple(flags, vel.copy(), distance=distance, phiObs=phiObs, intoObs=True),
                             I.extrapolate_ls_simple(phi.copy(), distance=distance), I.extrapolate_ls_simple(phi.copy(), distance=distance, inside=True),
                             I.extrapolate_vec3_simple(vel.copy(), phi, distance=distance))
            ref = (O.extrapolate_mac_simple(flags, vel.copy(), distance=distance), O.extrapolate_mac_simple(flags, vel.copy(), distance=distance, phiObs=phiObs, intoObs=True),
                   O.extrapolate_ls_simple(phi.copy(), distance=distance), O.extrapolate_ls_simple(phi.copy(), distance=distance, inside=True),
                   O.extrapolate_vec3_simple(vel.copy(), phi, distance=distance))
            for q in range(5):
                assert np.array_equal(res["1"][q], res["0"][q]), (shape, distance, q, "frontier vs sweeps")
                assert np.array_equal(res["1"][q], ref[q]), (shape, distance, q, "vs the oracle")


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(FREESURFACE_SCENES))
def test_freesurface_steps_match_the_reference(name, prec):
    """scenes/freesurface.py:54-84 through the adapter: identical fluid / empty cells after six steps, fields within the solver tolerance"""
    from cuda_impl import CudaImpl
    check_freesurface_against_golden(CudaImpl(prec), name, prec, tol=1e-4 if prec == 4 else 1e-8)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(SECORDER_SCENES))
def test_second_order_boundary_scenario_matches_the_reference(name, prec):
    """tools/tests/test_1040_secOrderBnd.py on the device: identical fill fractions and flags; fields within 3e-4 in float (the reference's own
    threshold for this test is 1e-4 per field; two float multigrid implementations already differ by 2e-5 after its ten steps, and a 1-ulp
    perturbation of the solves moves the result by 1e-5)"""
    from cuda_impl import CudaImpl
    check_sec_order_bnd_against_golden(CudaImpl(prec), name, prec, tol=3e-4 if prec == 4 else 1e-8)


def test_freesurface_steps_device_resident(mf):
    """the same loop written like the scene: grids are created once and stay on the device until the end"""
    shape, pc = FREESURFACE_SCENES["fs3d"]
    sz, sy, sx = shape
    g = load_golden("step_fs3d", 4)
    from mantaflow_b200 import scenes
    k, j, i = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    basin = (j + 0.5) - 0.2 * sy
    drop = np.sqrt((i + 0.5 - 0.5 * sx) ** 2 + (j + 0.5 - 0.5 * sy) ** 2 + (k + 0.5 - 0.5 * sz) ** 2) - 0.125 * sx
    s = mf.Solver(gridSize=(sx, sy, sz), dim=3, prec=4)
    flags = mf.FlagGrid(s, scenes.closed_box_flags(sx, sy, sz, boundaryWidth=1))
    phi = mf.LevelsetGrid(s, np.minimum(basin, drop).astype(np.float32))
    vel, pressure = s.create(mf.MACGrid), s.create(mf.RealGrid)
    flags.updateFromLevelset(phi)
    launches0 = s.kernelLaunches()
    for _ in range(6):
        mf.extrapolateLsSimple(phi=phi, distance=5, inside=False)
        mf.extrapolateLsSimple(phi=phi, distance=5, inside=True)
        mf.extrapolateMACSimple(flags=flags, vel=vel, distance=5)
        mf.advectSemiLagrange(flags=flags, vel=vel, grid=phi, order=2, clampMode=2)
        phi.setBound(1, 1.)
        flags.updateFromLevelset(phi)
        mf.advectSemiLagrange(flags=flags, vel=vel, grid=vel, order=2)
        mf.addGravity(flags=flags, vel=vel, gravity=(0, -0.025, 0))
        mf.setWallBcs(flags=flags, vel=vel)
        mf.solvePressure(flags=flags, vel=vel, pressure=pressure, cgMaxIterFac=5, cgAccuracy=1e-5, phi=phi)
    assert s.kernelLaunches() > launches0
    assert not (flags._hostDirty or phi._hostDirty or vel._hostDirty), "a grid went back to the host inside the loop"
    assert np.array_equal(flags.numpy(), g["flags"])
    for a, key, tol in ((phi, "phi", 1e-4), (vel, "vel", 1e-4), (pressure, "pressure", 1e-3)):
        assert float(np.abs(a.numpy().astype(np.float64) - g[key]).max()) <= tol, key


def test_invalid_arguments_fail_loudly(mf):
    flags_h, vel_h, phi_h, _ = liquid_scene("liq2d", 4)
    s = mf.Solver(gridSize=(24, 28, 1), dim=2, prec=4)
    F, V, P = mf.FlagGrid(s, flags_h), mf.MACGrid(s, vel_h), mf.LevelsetGrid(s, phi_h)
    with pytest.raises(mf.MantaError):
        mf.extrapolateMACSimple(F, V, distance=251)
    with pytest.raises(mf.MantaError):
        mf.extrapolateMACSimple(V, V)                 # flags is not a FlagGrid
    with pytest.raises(mf.MantaError):
        mf.extrapolateLsSimple(V)                     # not a real grid
    with pytest.raises(mf.MantaError):
        mf.extrapolateVec3Simple(P, P)                # not a Vec3 grid
    with pytest.raises(mf.MantaError):
        F.updateFromLevelset(V)                       # not a level set
    t = mf.Solver(gridSize=(2, 28, 1), dim=2, prec=4)
    with pytest.raises(mf.MantaError):
        mf.extrapolateLsSimple(mf.LevelsetGrid(t))    # no interior cells


@pytest.mark.parametrize("ext", [".uni", ".raw", ".npz"])
def test_grid_save_load_from_device_mirror(mf, ext, tmp_path):
    """Grid.save writes what the DEVICE holds (the host copy is stale after a device plugin), Grid.load reaches the device with the next plugin"""
    flags_h, vel_h, phi_h, _ = liquid_scene("liq3d", 4)
    sz, sy, sx = flags_h.shape
    s = mf.Solver(gridSize=(sx, sy, sz), dim=3, prec=4)
    P, V, F = mf.LevelsetGrid(s, phi_h), mf.MACGrid(s, vel_h), mf.FlagGrid(s, flags_h)
    P.setBound(0.5, 1); F.updateFromLevelset(P); mf.extrapolateMACSimple(F, V, distance=3)      # device copies are now the newer ones
    names = [str(tmp_path / (n + ext)) for n in ("phi", "vel", "flags")]
    for g, n in zip((P, V, F), names):
        assert g._devDirty and g.save(n) == 1
    P2, V2, F2 = mf.LevelsetGrid(s), mf.MACGrid(s), mf.FlagGrid(s)
    for g, n in zip((P2, V2, F2), names):
        assert g.load(n) == 1 and g._hostDirty
    assert np.array_equal(P2.numpy(), P.numpy()) and np.array_equal(V2.numpy(), V.numpy()) and np.array_equal(F2.numpy(), F.numpy())
    assert np.array_equal(P.numpy()[0], np.full((sy, sx), 0.5, np.float32))
    mf.extrapolateLsSimple(P2, distance=3); mf.extrapolateLsSimple(P, distance=3)               # the loaded grid is uploaded by its first plugin
    assert np.array_equal(P2.numpy(), P.numpy())
    R = mf.RealGrid(s)
    if ext == ".uni":
        assert R.load(names[0]) == 1                  # real <-> levelset are interchangeable (unifyGridType)
        with pytest.raises(mf.MantaError):
            R.load(names[2])                          # a flag grid is not

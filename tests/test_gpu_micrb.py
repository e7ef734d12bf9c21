"""GPU parity of the reformulated preconditioner -- MIC(0) in block red-black ordering (csrc/mp_micrb.cu, mp_set_mic_ordering) -- against its
specification (oracle/mf_oracle.c micrb_init / micrb_apply, pinned on the CPU in tests/test_micrb.py): factor and sweeps bit for bit in both
precisions, PCG iteration counts equal, float solves bit-identical; solvePressure(PcMIC) with the ordering switched on converges to the
reference's pressure within north_star's tolerance, with the iteration counts of both orderings reported."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from mantaflow_b200 import scenes  # noqa: E402

TOL = {4: 1e-4, 8: 1e-10}


@pytest.fixture(scope="module")
def mf():
    import mantaflow_b200 as m
    if m.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests need a B200")
    return m


def oracle(prec):
    from oracle.oracle_api import Oracle
    return Oracle("port", prec)


def mk(mf, shape, prec):
    sz, sy, sx = shape
    return mf.Solver(gridSize=(sx, sy, sz), dim=3, prec=prec)


def rel_l2(a, b):
    d = np.linalg.norm((a.astype(np.float64) - b.astype(np.float64)).ravel())
    n = np.linalg.norm(b.astype(np.float64).ravel())
    return d / n if n > 0 else d


SCENES = {
    "smoke24": lambda prec: scenes.smoke_plume(24, prec, random_vel=True) + (None,),
    "smoke_ragged": lambda prec: scenes.smoke_plume((37, 26, 19), prec, random_vel=True) + (None,),       # rows that are not whole chunks
    "liquid28": lambda prec: scenes.liquid_basin(28, prec),
    "smoke_vec": lambda prec: scenes.smoke_plume((40, 21, 13), prec, random_vel=True) + (None,),          # tiles cut by the grid in y and z
    "liquid_vec": lambda prec: scenes.liquid_basin((48, 36, 30), prec),
}
TILES = [(8, 4), (8, 8), (16, 8), (16, 12), (64, 64)]


@pytest.mark.parametrize("tile", TILES)
@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("scene", list(SCENES))
def test_factor_and_sweeps_bit_exact(mf, scene, prec, tile):
    from mantaflow_b200 import cg
    flags, vel, phi = SCENES[scene](prec)
    O = oracle(prec)
    A_o = O.make_matrix(flags, phi=phi)
    tiles = (0, tile[0], tile[1])
    P_o = O.micrb_init(flags, *A_o, tiles=tiles)
    src = (np.random.Generator(np.random.PCG64(10)).random(flags.shape) - 0.5).astype(vel.dtype)
    z_o = O.micrb_apply(flags, src, P_o, *A_o, tiles=tiles)
    s = mk(mf, flags.shape, prec)
    s.setMicOrdering(1, *tile)
    F = mf.FlagGrid(s, flags)
    A = [mf.RealGrid(s, a) for a in A_o]
    P = mf.RealGrid(s)
    cg.InitPreconditionModifiedIncompCholesky2(F, P, *A)
    assert np.array_equal(P.numpy(), P_o)
    S, D = mf.RealGrid(s, src), mf.RealGrid(s)
    cg.ApplyPreconditionModifiedIncompCholesky2(D, S, F, P, *A)
    assert np.array_equal(D.numpy(), z_o)
    if tile == (64, 64):        # one tile: the reference's own MIC(0)
        assert np.array_equal(P_o, O.mic_init(flags, *A_o)) and np.array_equal(z_o, O.mic_apply(flags, src, P_o, *A_o))
    # switched off again the same calls give the lexicographic factor
    s.setMicOrdering(0)
    cg.InitPreconditionModifiedIncompCholesky2(F, P, *A)
    assert np.array_equal(P.numpy(), O.mic_init(flags, *A_o)) and s.micOrdering() == (0, 0, 0)
    s.close()


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("scene", ["smoke24", "liquid28", "smoke_vec"])
def test_gridcg_with_the_block_red_black_factor(mf, scene, prec):
    from mantaflow_b200 import cg
    flags, vel, phi = SCENES[scene](prec)
    O = oracle(prec)
    rhs_o, _, _ = O.compute_rhs(flags, vel, phi=phi)
    A_o = O.make_matrix(flags, phi=phi)
    acc = 1e-5 if prec == 4 else 1e-11
    O.set_mic_tiles((0, 8, 8))
    x_o, it_o, _ = O.cg_solve(flags, rhs_o, *A_o, pc=4, accuracy=acc, maxIter=3000)
    O.set_mic_tiles((0, 0, 0))
    _, it_lex, _ = O.cg_solve(flags, rhs_o, *A_o, pc=1, accuracy=acc, maxIter=3000)
    s = mk(mf, flags.shape, prec)
    s.setMicOrdering(1, 8, 8)
    F = mf.FlagGrid(s, flags)
    A = [mf.RealGrid(s, a) for a in A_o]
    x, b, r, se, t = mf.RealGrid(s), mf.RealGrid(s, rhs_o), mf.RealGrid(s), mf.RealGrid(s), mf.RealGrid(s)
    g = cg.GridCg(x, b, r, se, F, t, *A)
    g.setAccuracy(acc)
    g.setUseL2Norm(False)
    pcs = [mf.RealGrid(s) for _ in range(4)]
    g.setICPreconditioner(2, *pcs)
    g.solve(3000)
    print("%s prec %d: iterations block red-black %d (specification %d), lexicographic %d" % (scene, prec, g.getIterations(), it_o, it_lex))
    assert abs(g.getIterations() - it_o) <= (0 if prec == 4 else 1) and g.getResNorm() < acc
    if prec == 4:
        assert np.array_equal(x.numpy(), x_o)           # same arithmetic in the same order: the same bits
    assert rel_l2(x.numpy(), x_o) <= TOL[prec]
    s.close()


@pytest.mark.parametrize("prec", [4, 8])
def test_solve_pressure_pcmic_reformulated(mf, prec):
    """the plugin with PcMIC and the ordering switched on: pressure within tolerance of the reference ordering's, divergence at its level"""
    flags, vel = scenes.smoke_plume((64, 48, 40), prec, random_vel=True)
    O = oracle(prec)
    acc = 1e-6 if prec == 4 else 1e-12
    v_o = vel.copy()
    p_o, it_o, _ = O.solve_pressure(flags, v_o, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=1)
    s = mk(mf, flags.shape, prec)
    F, V, P = mf.FlagGrid(s, flags), mf.MACGrid(s, vel), mf.RealGrid(s)
    s.setMicOrdering(1, 8, 8)
    mf.solvePressure(vel=V, pressure=P, flags=F, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=mf.PcMIC)
    it_rb = mf.lastSolveInfo()["iterations"]
    assert s.micOrdering() == (1, 8, 8)
    fl = (flags & 1) != 0
    pg, po = P.numpy().astype(np.float64), p_o.astype(np.float64)
    pg[fl] -= pg[fl].mean(); po[fl] -= po[fl].mean()
    print("prec %d: iterations lexicographic %d, block red-black 8x8 %d" % (prec, it_o, it_rb))
    assert rel_l2(pg[fl], po[fl]) <= TOL[prec] and it_o <= it_rb <= 2.5 * it_o
    assert scenes.max_divergence(flags, V.numpy()) <= max(2 * scenes.max_divergence(flags, v_o), 10 * acc)
    # the default ordering is untouched by all this
    s.setMicOrdering(0)
    V.copyFromArray(vel)
    mf.solvePressure(vel=V, pressure=P, flags=F, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=mf.PcMIC)
    assert abs(mf.lastSolveInfo()["iterations"] - it_o) <= (0 if prec == 4 else 1) and s.micOrdering() == (0, 0, 0)
    if prec == 4:
        assert np.array_equal(P.numpy(), p_o)
    # tile 0 x 0: chosen from the grid (8 x 4 on a grid this small)
    s.setMicOrdering(1)
    V.copyFromArray(vel)
    mf.solvePressure(vel=V, pressure=P, flags=F, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=mf.PcMIC)
    assert s.micOrdering() == (1, 8, 4) and mf.lastSolveInfo()["iterations"] >= it_o
    s.close()


def test_face_fractions_keep_the_lexicographic_kernels(mf):
    """a matrix with off-diagonals other than 0 / -1 cannot use the byte mask: the call falls back to the reference ordering (still on the device)"""
    from mantaflow_b200 import cg
    prec = 4
    flags, vel, _ = SCENES["smoke24"](prec)
    O = oracle(prec)
    A_o = [a.copy() for a in O.make_matrix(flags)]
    A_o[1][A_o[1] == -1] = -0.5
    s = mk(mf, flags.shape, prec)
    s.setMicOrdering(1, 8, 4)
    F = mf.FlagGrid(s, flags)
    A = [mf.RealGrid(s, a) for a in A_o]
    P = mf.RealGrid(s)
    cg.InitPreconditionModifiedIncompCholesky2(F, P, *A)
    assert np.array_equal(P.numpy(), O.mic_init(flags, *A_o)) and s.micOrdering() == (0, 0, 0)
    s.close()

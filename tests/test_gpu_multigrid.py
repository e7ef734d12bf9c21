"""GPU parity tests for GridMg (multigrid.cpp) and the PcMGDynamic / PcMGStatic preconditioners.

The vertex types of every level must equal the oracle's exactly (the coarse-vertex selection is integer work);
level-0/1 operators are bit-exact for the integer-valued Laplacian, coarser Galerkin operators and the V-cycle are
compared with a floating-point tolerance because the coarse solve reduces in a different (fixed) order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from mantaflow_b200 import scenes  # noqa: E402
from test_gpu_parity import SCENES, mf, mk, oracle, rel_l2, TOL  # noqa: E402,F401


def setup_system(prec, scene, pin=True):
    flags, vel, phi = SCENES[scene](prec)
    O = oracle(prec)
    rhs, _, _ = O.compute_rhs(flags, vel, phi=phi)
    A = O.make_matrix(flags, phi=phi)
    if pin:
        fix = O.choose_fix_cell(flags)
        if fix >= 0:
            O.fix_pressure(flags, fix, 0.0, rhs, *A)
    return flags, vel, phi, O, rhs, A


def isolated_features_flags(prec):
    """a domain with isolated single fluid cells and thin diagonal features: forces the order-dependent phase of
    genCoarseGrid (multigrid.cpp:543-575), i.e. the exact serial fallback"""
    flags, vel = scenes.smoke_plume((26, 22, 18), prec, random_vel=True)
    rng = np.random.Generator(np.random.PCG64(21))
    inner = np.zeros(flags.shape, bool); inner[1:-1, 1:-1, 1:-1] = True
    keep = rng.random(flags.shape) < 0.18
    flags[inner & ~keep] = 2          # most of the interior becomes obstacle: many isolated fluid cells
    scenes.set_wall_bcs(flags, vel)
    return flags, vel


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("scene", ["smoke24", "smoke_ragged", "liquid28", "smoke2d", "liquid2d"])
def test_hierarchy_matches_oracle(mf, scene, prec):
    from mantaflow_b200 import cg
    flags, vel, phi, O, rhs, A_o = setup_system(prec, scene)
    sz, sy, sx = flags.shape
    O.mg_create(sx, sy, sz); O.mg_set_a(*A_o)
    s = mk(mf, flags.shape, prec)
    A = [mf.RealGrid(s, a) for a in A_o]
    mg = cg.GridMg(s)
    assert not mg.isASet()
    mg.setA(*A)
    assert mg.isASet()
    assert mg.numLevels() == O.mg_num_levels()
    for l in range(mg.numLevels()):
        assert mg.levelInfo(l)[0] == O.mg_level_size(l)
        t, t_o = mg.download("type", l), O.mg_get("type", l)
        assert np.array_equal(t, t_o), "vertex types differ on level %d" % l
        a, a_o = mg.download("a", l), O.mg_get("a", l)
        st = mg.levelInfo(l)[1]
        act = np.repeat(t_o != 0, st)
        if l == 0 or phi is None:
            assert np.array_equal(a[act], a_o[act]), "operator differs on level %d" % l
        else:
            assert np.allclose(a[act], a_o[act], rtol=1e-5 if prec == 4 else 1e-13, atol=1e-6 if prec == 4 else 1e-14)
    O.mg_destroy()


@pytest.mark.parametrize("prec", [4, 8])
def test_hierarchy_isolated_features_exact_fallback(mf, prec):
    from mantaflow_b200 import cg
    flags, vel = isolated_features_flags(prec)
    O = oracle(prec)
    A_o = O.make_matrix(flags)
    sz, sy, sx = flags.shape
    O.mg_create(sx, sy, sz); O.mg_set_a(*A_o)
    s = mk(mf, flags.shape, prec)
    A = [mf.RealGrid(s, a) for a in A_o]
    mg = cg.GridMg(s)
    mg.setA(*A)
    for l in range(mg.numLevels()):
        assert np.array_equal(mg.download("type", l), O.mg_get("type", l)), "vertex types differ on level %d" % l
        st = mg.levelInfo(l)[1]
        act = np.repeat(O.mg_get("type", l) != 0, st)
        assert np.array_equal(mg.download("a", l)[act], O.mg_get("a", l)[act])
    O.mg_destroy()


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("scene", ["smoke24", "smoke_ragged", "liquid28", "smoke2d"])
def test_vcycle_matches_oracle(mf, scene, prec):
    from mantaflow_b200 import cg
    flags, vel, phi, O, rhs, A_o = setup_system(prec, scene)
    sz, sy, sx = flags.shape
    O.mg_create(sx, sy, sz); O.mg_set_a(*A_o)
    z_o = O.mg_vcycle(rhs, coarsestAccuracy=1e-9, pre=1, post=1)
    s = mk(mf, flags.shape, prec)
    A = [mf.RealGrid(s, a) for a in A_o]
    mg = cg.GridMg(s)
    B, Z = mf.RealGrid(s, rhs), mf.RealGrid(s)
    with pytest.raises(mf.MantaError, match="not been set"):
        mg.setRhs(B)
    mg.setA(*A)
    mg.setCoarsestLevelAccuracy(1e-9)
    mg.setSmoothing(1, 1)
    mg.setRhs(B)
    res = mg.doVCycle(Z)
    assert rel_l2(Z.numpy(), z_o) <= (1e-5 if prec == 4 else 1e-9)
    assert res >= 0
    # second cycle starting from the first result (src != NULL path, multigrid.cpp:457)
    Z2 = mf.RealGrid(s)
    res2 = mg.doVCycle(Z2, Z)
    assert res2 < res
    O.mg_destroy()


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("pc", [2, 3])
@pytest.mark.parametrize("scene", ["smoke24", "smoke_ragged", "liquid28", "smoke2d"])
def test_solve_pressure_mg(mf, scene, pc, prec):
    flags, vel, phi = SCENES[scene](prec)
    O = oracle(prec)
    acc = 1e-5 if prec == 4 else 1e-11
    v_o = vel.copy()
    p_o, it_o, rn_o = O.solve_pressure(flags, v_o, phi=phi, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=True)
    s = mk(mf, flags.shape, prec)
    F, V, P = mf.FlagGrid(s, flags), mf.MACGrid(s, vel), mf.RealGrid(s)
    PH = mf.RealGrid(s, phi) if phi is not None else None
    mf.solvePressure(vel=V, pressure=P, flags=F, phi=PH, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=True)
    info = mf.lastSolveInfo()
    assert abs(info["iterations"] - it_o) <= 1, (info["iterations"], it_o)
    assert info["maxIter"] == 100 and info["mgLevels"] >= 2
    assert rel_l2(P.numpy(), p_o) <= TOL[prec]
    assert rel_l2(V.numpy(), v_o) <= TOL[prec]
    mf.releaseMG(s)


def test_static_mg_reuses_hierarchy_like_test_0110(mf):
    """tools/tests/test_0110_mgsolve.py:57-68: two PcMGStatic solves with different velocities on one hierarchy,
    preceded by a PcMGDynamic solve that must drop any previous hierarchy (pressure.cpp:423-426)."""
    prec = 4
    flags, vel = scenes.smoke_plume(28, prec)
    O = oracle(prec)
    s = mk(mf, flags.shape, prec)
    F, V, P = mf.FlagGrid(s, flags), mf.MACGrid(s), mf.RealGrid(s)
    key = 77
    for step, (pc, scale) in enumerate([(3, 1.0), (3, -1.3), (2, 0.7), (3, 2.0)]):
        _, v = scenes.smoke_plume(28, prec, scale=scale)
        v_o = v.copy()
        p_o, it_o, _ = O.solve_pressure(flags, v_o, cgAccuracy=1e-5, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=True, solver_key=key)
        V.copyFromArray(v)
        mf.solvePressure(vel=V, pressure=P, flags=F, cgAccuracy=1e-5, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=True)
        info = mf.lastSolveInfo()
        assert abs(info["iterations"] - it_o) <= 1, (step, info["iterations"], it_o)
        assert rel_l2(P.numpy(), p_o) <= 1e-4 and rel_l2(V.numpy(), v_o) <= 1e-4
    O.release_solver(key)
    mf.releaseMG(s)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("shape,liquid", [((24, 24, 24), False), ((40, 36, 64), False), ((33, 47, 72), True), ((1, 40, 48), False), ((70, 66, 132), False)])
def test_vcycle_kernel_forms_agree_bit_for_bit(mf, shape, liquid, prec, monkeypatch):
    """The V-cycle's bandwidth forms -- level 0 with 16 bytes of cells per thread (k_mg_l0_vec, k_mg_restrict_l0_vec, k_mg_interp_add_l0_vec)
    and the colour-major full rows of levels > 0 (k_mg_build_full / k_mg_sweep_full) -- do the arithmetic of the one-cell-per-thread
    kernels term for term: the iterate after two V-cycles is the same bit for bit (both are compared with the oracle elsewhere in this file)."""
    from mantaflow_b200 import cg
    from test_gpu_parity import random_domain
    flags, vel, phi = random_domain(shape, prec, seed=shape[1], liquid=liquid)
    O = oracle(prec)
    rhs, _, _ = O.compute_rhs(flags, vel, phi=phi)
    A_o = O.make_matrix(flags, phi=phi)
    fix = O.choose_fix_cell(flags)
    if fix >= 0:
        O.fix_pressure(flags, fix, 0.0, rhs, *A_o)
    out = {}
    for full, vec in ((1, 1), (0, 0), (1, 0), (0, 1)):
        monkeypatch.setenv("MP_MG_FULL", str(full)); monkeypatch.setenv("MP_MG_L0VEC", str(vec))
        s = mk(mf, flags.shape, prec)
        A = [mf.RealGrid(s, a) for a in A_o]
        mg = cg.GridMg(s)
        B, Z, Z2 = mf.RealGrid(s, rhs), mf.RealGrid(s), mf.RealGrid(s)
        mg.setA(*A)
        mg.setSmoothing(2, 1)
        mg.setRhs(B)
        r1 = mg.doVCycle(Z)
        r2 = mg.doVCycle(Z2, Z)
        out[(full, vec)] = (Z.numpy().copy(), Z2.numpy().copy(), r1, r2)
        assert r2 < r1
        mg.close(); s.close()
    ref = out[(0, 0)]
    for k, v in out.items():
        assert np.array_equal(v[0], ref[0]) and np.array_equal(v[1], ref[1]), ("V-cycle forms differ", k)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("shape,liquid", [((24, 24, 24), False), ((40, 36, 64), False), ((33, 47, 72), True), ((1, 40, 48), False), ((70, 66, 132), False), ((37, 50, 260), True)])
def test_vcycle_level0_fused_agrees_bit_for_bit(mf, shape, liquid, prec, monkeypatch):
    """Level 0 of the V-cycle on the 2-byte operator mask -- per colour (k_mg_l0_vecm) and as the fused single-pass kernels (mp_mg_l0_fused.cuh: zero iterate + both colours + residual in one pass on the
    way down, both colours of the post-smoothing in one pass on the way up, the operator as 2 bytes per vertex) does the arithmetic of the
    per-colour kernels term for term: iterates, residual norms and a PcMGStatic solve are the same bit for bit.  (The host walk of the same
    phase functions is checked against numpy in tests/test_mg_l0_fused_emul.py.)"""
    from mantaflow_b200 import cg
    from test_gpu_parity import random_domain
    flags, vel, phi = random_domain(shape, prec, seed=shape[1] + 1, liquid=liquid)
    O = oracle(prec)
    rhs, _, _ = O.compute_rhs(flags, vel, phi=phi)
    A_o = O.make_matrix(flags, phi=phi)
    fix = O.choose_fix_cell(flags)
    if fix >= 0:
        O.fix_pressure(flags, fix, 0.0, rhs, *A_o)
    out = {}
    for fused, masked in ((1, 1), (0, 1), (0, 0)):
        monkeypatch.setenv("MP_MG_L0FUSED", str(fused)); monkeypatch.setenv("MP_MG_L0MASK", str(masked))
        s = mk(mf, flags.shape, prec)
        A = [mf.RealGrid(s, a) for a in A_o]
        mg = cg.GridMg(s)
        B, Z, Z2 = mf.RealGrid(s, rhs), mf.RealGrid(s), mf.RealGrid(s)
        mg.setA(*A)
        assert mg.level0Fused() == bool(fused)
        mg.setRhs(B)
        r1 = mg.doVCycle(Z)                     # fused down + fused up
        r2 = mg.doVCycle(Z2, Z)                 # initial guess given: per-colour pre-smoothing, fused up
        # the preconditioner path (rhs scaled on the fly, no residual norm)
        F, V, P = mf.FlagGrid(s, flags), mf.MACGrid(s, vel), mf.RealGrid(s)
        PH = mf.RealGrid(s, phi) if phi is not None else None
        mf.solvePressure(vel=V, pressure=P, flags=F, cgAccuracy=1e-5 if prec == 4 else 1e-9, phi=PH, preconditioner=mf.PcMGStatic, cgMaxIterFac=99,
                         zeroPressureFixing=True)
        out[(fused, masked)] = (Z.numpy().copy(), Z2.numpy().copy(), r1, r2, P.numpy().copy(), V.numpy().copy(), mf.lastSolveInfo()["iterations"])
        assert r2 < r1
        mf.releaseMG(s)
        mg.close(); s.close()
    b = out[(0, 0)]      # the per-colour kernels on the coefficient arrays
    for key in ((1, 1), (0, 1)):
        a = out[key]
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), ("level-0 V-cycle form differs from the per-colour kernels", key)
        assert a[2] == b[2] and a[3] == b[3], key
        assert a[6] == b[6] and np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5]), ("PcMGStatic solve differs", key)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("res,liquid", [(48, False), (72, False), (64, True)])
def test_galerkin_regular_vertices_copy_the_same_row(mf, res, liquid, prec, monkeypatch):
    """setA's fast path: coarse vertices whose whole neighbourhood is the unperturbed operator (k_mg_classify1 / k_mg_classifyN) take the row of the
    level's first regular vertex instead of recomputing it.  Every level's operator equals the one computed vertex by vertex (MP_MG_REGULAR=0),
    bit for bit -- on a smoke plume with a sphere obstacle (mostly regular) and on a liquid basin (ghost-fluid diagonals along the surface)."""
    from mantaflow_b200 import cg, scenes
    if liquid:
        flags, vel, phi = scenes.liquid_basin((res, res, res), prec)
    else:
        (flags, vel), phi = scenes.smoke_plume(res, prec), None
    O = oracle(prec)
    rhs, _, _ = O.compute_rhs(flags, vel, phi=phi)
    A_o = O.make_matrix(flags, phi=phi)
    fix = O.choose_fix_cell(flags)
    if fix >= 0:
        O.fix_pressure(flags, fix, 0.0, rhs, *A_o)
    ops = {}
    for regular in (1, 0):
        monkeypatch.setenv("MP_MG_REGULAR", str(regular))
        s = mk(mf, flags.shape, prec)
        mg = cg.GridMg(s)
        mg.setA(*[mf.RealGrid(s, a) for a in A_o])
        ops[regular] = [mg.download("a", l) for l in range(mg.numLevels())]
        mg.close(); s.close()
    assert len(ops[1]) >= 3
    for l, (a, b) in enumerate(zip(ops[1], ops[0])):
        assert np.array_equal(a, b), "level %d operator differs with the regular-vertex fast path" % l
    assert np.count_nonzero(ops[1][1]) > 0


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("res,liquid", [(72, False), (64, True)])
def test_sweeps_with_constant_rows_agree_bit_for_bit(mf, res, liquid, prec, monkeypatch):
    """levels > 0: vertices whose whole 27-entry row equals the row of the level's first regular vertex (compared number by number in setA) take
    their coefficients from that constant row instead of streaming them (k_mg_sweep_full with rowreg).  Same numbers, same order: V-cycle iterates,
    residual norms and a PcMGStatic solve are the same bit for bit with MP_MG_ROWREG=0."""
    from mantaflow_b200 import cg, scenes
    if liquid:
        flags, vel, phi = scenes.liquid_basin((res, res, res), prec)
    else:
        (flags, vel), phi = scenes.smoke_plume(res, prec), None
    O = oracle(prec)
    rhs, _, _ = O.compute_rhs(flags, vel, phi=phi)
    A_o = O.make_matrix(flags, phi=phi)
    fix = O.choose_fix_cell(flags)
    if fix >= 0:
        O.fix_pressure(flags, fix, 0.0, rhs, *A_o)
    out = {}
    for rowreg in (1, 0):
        monkeypatch.setenv("MP_MG_ROWREG", str(rowreg))
        s = mk(mf, flags.shape, prec)
        mg = cg.GridMg(s)
        B, Z, Z2 = mf.RealGrid(s, rhs), mf.RealGrid(s), mf.RealGrid(s)
        mg.setA(*[mf.RealGrid(s, a) for a in A_o])
        mg.setRhs(B)
        r1 = mg.doVCycle(Z)
        r2 = mg.doVCycle(Z2, Z)
        F, V, P = mf.FlagGrid(s, flags), mf.MACGrid(s, vel), mf.RealGrid(s)
        PH = mf.RealGrid(s, phi) if phi is not None else None
        mf.solvePressure(vel=V, pressure=P, flags=F, cgAccuracy=1e-5 if prec == 4 else 1e-9, phi=PH, preconditioner=mf.PcMGStatic, cgMaxIterFac=99, zeroPressureFixing=True)
        out[rowreg] = (Z.numpy().copy(), Z2.numpy().copy(), r1, r2, P.numpy().copy(), mf.lastSolveInfo()["iterations"])
        mf.releaseMG(s); mg.close(); s.close()
    a, b = out[1], out[0]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3]
    assert a[5] == b[5] and np.array_equal(a[4], b[4])

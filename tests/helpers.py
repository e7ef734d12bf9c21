"""Shared helpers of the parity tests: golden fixtures and one generic checker that runs any implementation with the
oracle_api interface (the CPU restatement, the reference build, or the CUDA path through tests/cuda_impl.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)

KERNEL_FIXTURES = ["smoke16", "liquid14", "smoke2d"]


def load_golden(name, prec):
    """kernels_<name>_f<bits>.npz, or <name>_f<bits>.npz for the step_* / plume* fixtures"""
    stem = name if name.startswith(("step_", "plume", "icp_", "vic_")) else "kernels_" + name
    return dict(np.load(os.path.join(GOLDEN, "%s_f%d.npz" % (stem, prec * 8))))


def rel_l2(a, b):
    d = np.linalg.norm((np.asarray(a, np.float64) - np.asarray(b, np.float64)).ravel())
    n = np.linalg.norm(np.asarray(b, np.float64).ravel())
    return d / n if n > 0 else d


def check_kernels_against_golden(I, g, prec, exact_reductions):
    """I: implementation with the Oracle interface.  Integer / per-cell arithmetic must be bit-exact; whatever passes
    through a global reduction (CG scalars) is bit-exact only when `exact_reductions` (serial float oracle)."""
    flags, vel = g["flags"], g["vel"]
    phi = g.get("phi")
    sz = flags.shape[0]
    tol = 1e-4 if prec == 4 else 1e-10
    rhs, s, c = I.compute_rhs(flags, vel, phi=phi)
    assert c == int(g["rhs_cnt"])
    assert np.array_equal(rhs, g["rhs"]), "rhs not bit-exact"
    assert abs(s - float(g["rhs_sum"])) <= 1e-10 * max(1.0, c)
    A = I.make_matrix(flags, phi=phi)
    for n, a in zip("A0 Ai Aj Ak".split(), A):
        assert np.array_equal(a, g[n]), n + " not bit-exact"
    assert np.array_equal(I.apply_matrix(flags, g["src"], *A), g["apply_matrix"]), "ApplyMatrix not bit-exact"
    acc = 1e-5 if prec == 4 else 1e-11
    x, it, rn = I.cg_solve(flags, rhs, *A, pc=0, accuracy=acc, maxIter=4000)
    assert abs(it - int(g["cg_none_it"])) <= 1, ("PcNone iterations", it, int(g["cg_none_it"]))
    assert rel_l2(x, g["cg_none_x"]) <= tol
    if exact_reductions:
        assert it == int(g["cg_none_it"]) and np.array_equal(x, g["cg_none_x"])
    if sz > 1:
        P = I.mic_init(flags, *A)
        assert np.array_equal(P, g["mic_P"]), "MIC factor not bit-exact"
        assert np.array_equal(I.mic_apply(flags, g["src"], P, *A), g["mic_apply"]), "MIC sweeps not bit-exact"
        x, it, rn = I.cg_solve(flags, rhs, *A, pc=1, accuracy=acc, maxIter=4000)
        assert abs(it - int(g["cg_mic_it"])) <= 1
        assert rel_l2(x, g["cg_mic_x"]) <= tol
    if "fix_idx" in g:
        fix = int(g["fix_idx"])
        if hasattr(I, "choose_fix_cell") and getattr(I, "kind", "") != "reference":
            assert I.choose_fix_cell(flags) == fix, "pinned-cell choice differs from the reference"
        rhs_f = rhs.copy(); Af = [a.copy() for a in A]
        I.fix_pressure(flags, fix, 0.0, rhs_f, *Af)
        assert np.array_equal(rhs_f, g["fix_rhs"])
        for n, a in zip("fix_A0 fix_Ai fix_Aj fix_Ak".split(), Af):
            assert np.array_equal(a, g[n]), n
    else:
        rhs_f, Af = rhs, A
    sx, sy = flags.shape[2], flags.shape[1]
    I.mg_create(sx, sy, sz)
    I.mg_set_a(*Af)
    assert I.mg_num_levels() == int(g["mg_levels"])
    for l in range(int(g["mg_levels"])):
        assert tuple(I.mg_level_size(l)) == tuple(g["mg_size_%d" % l])
        t = I.mg_get("type", l)
        assert np.array_equal(t, g["mg_type_%d" % l]), "GridMg vertex types differ on level %d" % l
        a, a_g = I.mg_get("a", l), g["mg_a_%d" % l]
        act = np.repeat(t != 0, a.size // t.size)
        if l == 0 or phi is None:
            assert np.array_equal(a[act], a_g[act]), "GridMg operator differs on level %d" % l
        else:   # std::sort leaves the order of equal-key coarsening paths unspecified: last-bit differences allowed
            assert np.allclose(a[act], a_g[act], rtol=1e-5 if prec == 4 else 1e-13, atol=1e-6 if prec == 4 else 1e-14)
    z = I.mg_vcycle(rhs_f, coarsestAccuracy=1e-9, pre=1, post=1)
    assert rel_l2(z, g["mg_vcycle"]) <= (1e-5 if prec == 4 else 1e-9)
    I.mg_destroy()
    v = I.correct_velocity(flags, vel.copy(), g["cv_pressure"], phi=phi)
    assert np.array_equal(v, g["cv_vel"]), "correctVelocity not bit-exact"
    # the plugin
    v = vel.copy()
    p, it, rn = I.solve_pressure(flags, v, phi=phi, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=2, zeroPressureFixing=True)
    assert abs(it - int(g["plugin_mg_it"])) <= 1
    assert rel_l2(p, g["plugin_mg_p"]) <= tol and rel_l2(v, g["plugin_mg_vel"]) <= tol
    if sz > 1:
        v = vel.copy()
        p, it, rn = I.solve_pressure(flags, v, phi=phi, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=1)
        assert abs(it - int(g["plugin_mic_it"])) <= 1
        assert rel_l2(p, g["plugin_mic_p"]) <= tol and rel_l2(v, g["plugin_mic_vel"]) <= tol


def check_diffusion_against_golden(I, g, prec):
    tol = 1e-5 if prec == 4 else 1e-12
    d = I.cg_solve_diffusion(g["flags"], g["src"].copy(), alpha=0.7, cgMaxIterFac=2.0, cgAccuracy=1e-7)
    assert np.abs(d.astype(np.float64) - g["diff_real"]).max() <= tol
    v = I.cg_solve_diffusion(g["flags"], g["vel"].copy())
    assert np.abs(v.astype(np.float64) - g["diff_vec"]).max() <= tol


def check_psolve52(I, thr=1e-4):
    """tools/tests/test_0100_psolve.py and test_0110_mgsolve.py: max abs per-cell difference (gridMaxDiff, grid.cpp:400-430)
    against the reference's result below the float-build threshold 1e-4 (test_0100_psolve.py:40-41,:54-55)."""
    from make_golden import box_source
    from mantaflow_b200 import scenes
    g = dict(np.load(os.path.join(GOLDEN, "psolve52_f32.npz")))
    res, prec = 52, 4
    flags = scenes.closed_box_flags(res, res, res)
    kw = dict(cgMaxIterFac=99, cgAccuracy=1e-4)
    md = lambda a, b: float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
    v = box_source(res, (0.15, 0.3, 0.21), prec)
    p, it, _ = I.solve_pressure(flags, v, zeroPressureFixing=False, **kw)
    assert md(p, g["t0100_pressure0"]) <= thr and md(v, g["t0100_vel0"]) <= thr and abs(it - int(g["t0100_it0"])) <= 1
    v = box_source(res, (1.5, 3, 2.1), prec); scenes.set_wall_bcs(flags, v)
    p, it, _ = I.solve_pressure(flags, v, zeroPressureFixing=False, **kw)
    assert md(p, g["t0100_pressure"]) <= thr * 10 and md(v, g["t0100_vel"]) <= thr * 10 and abs(it - int(g["t0100_it1"])) <= 1   # 10x larger source
    key = 4242
    v = box_source(res, (0.15, 0.3, 0.21), prec)
    p, it, _ = I.solve_pressure(flags, v, zeroPressureFixing=True, preconditioner=2, solver_key=key, **kw)
    assert md(p, g["t0110_p0"]) <= thr and abs(it - int(g["t0110_it0"])) <= 1
    v = box_source(res, (1.5, 3, 2.1), prec); scenes.set_wall_bcs(flags, v)
    p, it, _ = I.solve_pressure(flags, v, zeroPressureFixing=True, preconditioner=2, solver_key=key, **kw)
    assert md(p, g["t0110_p1"]) <= thr * 10 and abs(it - int(g["t0110_it1"])) <= 1
    v = box_source(res, (1.1, 2, -2.1), prec); scenes.set_wall_bcs(flags, v)
    p, it, _ = I.solve_pressure(flags, v, zeroPressureFixing=True, preconditioner=3, solver_key=key, **kw)
    assert abs(it - int(g["t0110_it2"])) <= 1
    v = g["t0110_vel_in3"].copy()
    p, it, _ = I.solve_pressure(flags, v, zeroPressureFixing=True, preconditioner=3, solver_key=key, **kw)
    assert md(p, g["t0110_p2"]) <= thr * 10 and md(v, g["t0110_v2"]) <= thr * 10 and abs(it - int(g["t0110_it3"])) <= 1
    I.release_solver(key)


# ---------------------------------------------------------------- the steps either side of the projection
STEP_SCENES = {"box3d": (16, 18, 20), "box2d": (1, 24, 30), "ragged3d": (13, 17, 22)}      # (sz, sy, sx)
STEP_CASES = ["wall_obvel", "wall", "gravity_excl", "gravity_noscale", "buoyancy", "adv_real_o1", "adv_real_o2_c2", "adv_real_o2_c1",
              "adv_mac_o1", "adv_mac_o2_c2", "adv_mac_o2_c1", "adv_self_o2"]


def step_scene(name, prec):
    """closed box with random obstacle / empty cells, an outflow band below the top wall, random velocity / density / obstacle velocity"""
    from mantaflow_b200 import scenes
    shape = STEP_SCENES[name]
    sz, sy, sx = shape
    real = np.float32 if prec == 4 else np.float64
    rng = np.random.default_rng(5 + sx)
    flags = scenes.closed_box_flags(sx, sy, sz)
    inner = (slice(1, -1) if sz > 1 else slice(None), slice(1, -1), slice(1, -1))
    r = rng.random(flags[inner].shape)
    f = flags[inner]
    f[r < 0.08] = 2            # obstacle
    f[(r > 0.08) & (r < 0.14)] = 4   # empty
    flags[inner] = f
    flags[(slice(1, -1) if sz > 1 else slice(None)), sy - 2, 1:-1] = 16 | 4      # outflow | empty
    vel = (rng.random(shape + (3,)) * 3 - 1.5).astype(real)
    if sz == 1:
        vel[..., 2] = 0
    dens = rng.random(shape).astype(real)
    obvel = (rng.random(shape + (3,)) - 0.5).astype(real)
    return flags, vel, dens, obvel


def run_step_case(I, case, flags, vel, dens, obvel):
    real = vel.dtype
    if case == "wall_obvel":
        return I.set_wall_bcs_obvel(flags, vel.copy(), obvel)
    if case == "wall":
        return I.set_wall_bcs_obvel(flags, vel.copy(), None)
    if case == "gravity_excl":
        return I.add_gravity(flags, vel.copy(), (0.1, -0.3, 0.2), exclude=(dens - 0.3).astype(real), scale=True, dt=0.7)
    if case == "gravity_noscale":
        return I.add_gravity(flags, vel.copy(), (0.1, -0.3, 0.2), scale=False, dt=0.7)
    if case == "buoyancy":
        return I.add_buoyancy(flags, dens, vel.copy(), (0, -6e-4, 1e-4), coefficient=1.3, scale=True, dt=0.9)
    if case == "adv_real_o1":
        return I.advect_semi_lagrange(flags, vel, dens.copy(), order=1, dt=0.8)
    if case == "adv_real_o2_c2":
        return I.advect_semi_lagrange(flags, vel, dens.copy(), order=2, clampMode=2, dt=0.8)
    if case == "adv_real_o2_c1":
        return I.advect_semi_lagrange(flags, vel, dens.copy(), order=2, clampMode=1, strength=0.8, dt=0.8)
    if case == "adv_mac_o1":
        return I.advect_semi_lagrange(flags, vel, obvel.copy(), order=1, dt=0.8)
    if case == "adv_mac_o2_c2":
        return I.advect_semi_lagrange(flags, vel, obvel.copy(), order=2, clampMode=2, dt=0.8)
    if case == "adv_mac_o2_c1":
        return I.advect_semi_lagrange(flags, vel, obvel.copy(), order=2, clampMode=1, dt=1.7)
    if case == "adv_self_o2":      # advectSemiLagrange(vel=vel, grid=vel): the grid advects itself
        v = vel.copy()
        return I.advect_semi_lagrange(flags, v, v, order=2, dt=0.8)
    raise KeyError(case)


# advectSemiLagrange with cubic lookups (orderSpace 2, util/interpolHigh.h) and explicit-midpoint back-tracing (orderTrace 2, advection.cpp:32-37, :58-73)
STEP_HI_SCENES = ["box3d", "box2d"]
STEP_HI_ORDERS = [(2, 1), (1, 2), (2, 2)]          # (orderSpace, orderTrace)
STEP_HI_KINDS = ["real_o1", "real_o2_c2", "real_o2_c1", "mac_o1", "mac_o2_c2", "mac_o2_c1", "self_o2", "vec3_o1", "vec3_o2_c2", "vec3_o2_c1"]
# a cell-centred Grid<Vec3> (fnAdvectSemiLagrange<Grid<Vec3>>, advection.cpp:455-457) also with the default orders
STEP_HI_CASES = (["advhi_%s_s%dt%d" % (k, os_, ot) for (os_, ot) in STEP_HI_ORDERS for k in STEP_HI_KINDS] +
                 ["advhi_%s_s1t1" % k for k in STEP_HI_KINDS if k.startswith("vec3")])


def run_step_hi_case(I, case, flags, vel, dens, obvel):
    kind, orders = case[len("advhi_"):].rsplit("_", 1)
    kw = dict(orderSpace=int(orders[1]), orderTrace=int(orders[3]))
    if kind == "real_o1":
        return I.advect_semi_lagrange(flags, vel, dens.copy(), order=1, dt=0.8, **kw)
    if kind == "real_o2_c2":
        return I.advect_semi_lagrange(flags, vel, dens.copy(), order=2, clampMode=2, dt=0.8, **kw)
    if kind == "real_o2_c1":
        return I.advect_semi_lagrange(flags, vel, dens.copy(), order=2, clampMode=1, strength=0.8, dt=0.8, **kw)
    if kind == "mac_o1":
        return I.advect_semi_lagrange(flags, vel, obvel.copy(), order=1, dt=0.8, **kw)
    if kind == "mac_o2_c2":
        return I.advect_semi_lagrange(flags, vel, obvel.copy(), order=2, clampMode=2, dt=0.8, **kw)
    if kind == "mac_o2_c1":
        return I.advect_semi_lagrange(flags, vel, obvel.copy(), order=2, clampMode=1, dt=1.7, **kw)
    if kind == "self_o2":
        v = vel.copy()
        return I.advect_semi_lagrange(flags, v, v, order=2, dt=0.8, **kw)
    if kind == "vec3_o1":
        return I.advect_semi_lagrange(flags, vel, obvel.copy(), order=1, dt=0.8, vec3=True, **kw)
    if kind == "vec3_o2_c2":
        return I.advect_semi_lagrange(flags, vel, obvel.copy(), order=2, clampMode=2, dt=0.8, vec3=True, **kw)
    if kind == "vec3_o2_c1":
        return I.advect_semi_lagrange(flags, vel, obvel.copy(), order=2, clampMode=1, strength=0.9, dt=1.3, vec3=True, **kw)
    raise KeyError(case)


def check_step_hi_against_golden(I, name, prec):
    """the higher-order advection variants bit for bit against the reference's output on the same seeded inputs"""
    g = load_golden("step_hi_" + name, prec)
    flags, vel, dens, obvel = step_scene(name, prec)
    assert np.array_equal(flags, g["flags"]) and np.array_equal(vel, g["vel"])
    lin = run_step_case(I, "adv_real_o1", flags, vel, dens, obvel)
    for case in STEP_HI_CASES:
        out = run_step_hi_case(I, case, flags, vel, dens, obvel)
        assert np.array_equal(out, g[case]), (name, prec, case, float(np.abs(out.astype(np.float64) - g[case]).max()))
        if case.startswith("advhi_real_o1"):
            assert not np.array_equal(out, lin), case           # the variant really takes another path than the defaults


def check_step_against_golden(I, name, prec):
    """every step plugin bit for bit against the reference's output on the same seeded inputs"""
    g = load_golden("step_" + name, prec)
    flags, vel, dens, obvel = step_scene(name, prec)
    assert np.array_equal(flags, g["flags"]) and np.array_equal(vel, g["vel"])
    for case in STEP_CASES:
        out = run_step_case(I, case, flags, vel, dens, obvel)
        assert np.array_equal(out, g[case]), (name, prec, case, float(np.abs(out.astype(np.float64) - g[case]).max()))


def plume_scene(shape, prec):
    """simpleplume-like: closed box, density source blob near the floor"""
    from mantaflow_b200 import scenes
    sz, sy, sx = shape
    real = np.float32 if prec == 4 else np.float64
    flags = scenes.closed_box_flags(sx, sy, sz)
    k, j, i = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    src = ((i + 0.5 - 0.5 * sx) ** 2 + ((k + 0.5 - 0.5 * sz) ** 2 if sz > 1 else 0) <= (0.14 * sx) ** 2) & (j >= 0.08 * sy) & (j <= 0.14 * sy) & (flags == 1)
    return flags, src, real


def run_plume_steps(I, shape, prec, steps, pc=1):
    """the main loop of scenes/simpleplume.py:48-60 with a constant density source instead of the noise inflow"""
    flags, src, real = plume_scene(shape, prec)
    vel = np.zeros(shape + (3,), real); dens = np.zeros(shape, real); p = np.zeros(shape, real)
    its = []
    for _ in range(steps):
        dens[src] = 1
        I.advect_semi_lagrange(flags, vel, dens, order=2)
        I.advect_semi_lagrange(flags, vel, vel, order=2, strength=1.0)
        I.set_wall_bcs_obvel(flags, vel, None)
        I.add_buoyancy(flags, dens, vel, (0, -6e-4, 0))
        p, it, _ = I.solve_pressure(flags, vel, preconditioner=pc, zeroPressureFixing=(pc >= 2))
        its.append(it)
    return dens, vel, p, its


# ---------------------------------------------------------------- cgSolveWE (plugin/waves.cpp:86-147)
WE_SCENES = {"we2d": (1, 40, 36), "we3d": (14, 18, 20)}


def run_wave_steps(I, name, prec, crankNic, steps=3):
    from mantaflow_b200 import scenes
    sz, sy, sx = shape = WE_SCENES[name]
    real = np.float32 if prec == 4 else np.float64
    flags = scenes.closed_box_flags(sx, sy, sz)
    rng = np.random.default_rng(3)
    ut, utm1 = rng.random(shape).astype(real), rng.random(shape).astype(real)
    out = None
    for _ in range(steps):
        out = I.cg_solve_we(flags, ut, utm1, crankNic=crankNic, cSqr=0.3, dt=0.9)
    return ut, utm1, out


def check_waves_against_golden(I, name, prec, tol):
    g = load_golden("step_" + name, prec)
    for cn in (False, True):
        for a, key in zip(run_wave_steps(I, name, prec, cn), ("ut", "utm1", "out")):
            ref = g["%s_cn%d" % (key, int(cn))]
            assert np.abs(a.astype(np.float64) - ref).max() <= tol, (name, prec, cn, key)


# ---------------------------------------------------------------- PD_fluid_guiding (plugin/fluidguiding.cpp:294-353)
GUIDING_SCENES = {"guide2d": ((1, 40, 36), 3), "guide3d": ((14, 18, 20), 2)}      # shape, preconditioner (the scenes use the multigrid ones)


def guiding_scene(name, prec):
    """closed box with an obstacle block, small random velocity, the spiral target field of getSpiralVelocity (fluidguiding.cpp:171-192,
    strength 0.5) and the two-band weight of scenes/guiding_2d.py:55-56"""
    from mantaflow_b200 import scenes
    shape, pc = GUIDING_SCENES[name]
    sz, sy, sx = shape
    real = np.float32 if prec == 4 else np.float64
    flags = scenes.closed_box_flags(sx, sy, sz)
    if sz > 1:
        flags[5:8, 6:9, 7:11] = 2
    else:
        flags[0, 10:14, 12:16] = 2
    rng = np.random.default_rng(3)
    vel = ((rng.random(shape + (3,)) - 0.5) * 0.2).astype(real)
    if sz == 1:
        vel[..., 2] = 0
    scenes.set_wall_bcs(flags, vel)
    velT = np.zeros(shape + (3,), real)
    j, i = np.meshgrid(np.arange(sy), np.arange(sx), indexing="ij")
    dx = real(0.5 * (sx - 1)) - i.astype(real); dy = real(0.5 * (sy - 1)) - j.astype(real)
    h = np.sqrt(dx * dx + dy * dy)
    with np.errstate(invalid="ignore", divide="ignore"):
        velT[:, :, :, 0] = np.where(h > 0, dy / h, 0)[None]; velT[:, :, :, 1] = np.where(h > 0, -dx / h, 0)[None]
    velT *= real(0.5)
    w = np.ones(shape, real); w[:, sy // 2:, :] = 5
    return flags, vel, velT, w, pc


def run_guiding(I, name, prec):
    flags, vel, velT, w, pc = guiding_scene(name, prec)
    v = vel.copy()
    p, it = I.pd_fluid_guiding(flags, v, velT, w, blurRadius=2, sigma=0.99, maxIters=40, cgAccuracy=1e-4, preconditioner=pc, zeroPressureFixing=True)
    return v, p, it


def check_guiding_against_golden(I, name, prec, tol):
    g = load_golden("step_" + name, prec)
    v, p, it = run_guiding(I, name, prec)
    assert it == int(g["iterations"]), (it, int(g["iterations"]))
    assert np.abs(v.astype(np.float64) - g["vel"]).max() <= tol and np.abs(p.astype(np.float64) - g["pressure"]).max() <= tol * 10


# ---------------------------------------------------------------- liquid neighbours: fastmarch.cpp:337-542, grid.cpp:585-593,:844-854
LIQUID_SCENES = {"liq3d": (12, 16, 14), "liq2d": (1, 28, 24), "liqragged": (11, 13, 9)}      # (sz, sy, sx)
LIQUID_CASES = ["mac_d4", "mac_d3_into", "mac_d5_phiobs", "mac_d0", "ls_out_d4", "ls_in_d5", "ls_out_d1", "ls_in_d2", "v3_out_d4", "v3_in_d2",
                "v3_out_d1", "from_levelset", "bound_real_w0", "bound_real_w2", "bound_vec_w1", "laplacian", "curvature_h1", "curvature_h07", "wall_frac", "macw_d2", "macw_d4", "macw_marks_d4",
                "fractions_w0", "fractions_open_w0", "fractions_open_w1", "obsflags_frac", "obsflags_phi_w2"]
LIQUID_ULP_CASES = ("curvature_h1", "curvature_h07")      # end in a double pow(), which neither libm nor the device rounds correctly: last bit may differ


def liquid_scene(name, prec):
    """basin + drop level set with noise, flags from it plus random obstacle cells, random velocity, a random obstacle level set"""
    from mantaflow_b200 import scenes
    sz, sy, sx = shape = LIQUID_SCENES[name]
    real = np.float32 if prec == 4 else np.float64
    flags, _, phi = scenes.liquid_basin((sx, sy, sz), prec)
    rng = np.random.default_rng(11 + sx)
    inner = np.zeros(shape, bool)
    inner[(slice(1, -1) if sz > 1 else slice(None)), 1:-1, 1:-1] = True
    flags[inner & (rng.random(shape) < 0.06)] = 2
    vel = (rng.random(shape + (3,)) - 0.5).astype(real)
    if sz == 1:
        vel[..., 2] = 0
    phi = (phi + (rng.random(shape) - 0.5).astype(real)).astype(real)
    phi[rng.random(shape) < 0.03] = -2000          # below invalidTimeValue: updateFromLevelset leaves those cells alone
    phiObs = (rng.random(shape) * 6 - 4).astype(real)
    return flags, vel, phi, phiObs


def run_liquid_case(I, case, flags, vel, phi, phiObs):
    if case == "mac_d4":
        return I.extrapolate_mac_simple(flags, vel.copy(), distance=4)
    if case == "mac_d3_into":
        return I.extrapolate_mac_simple(flags, vel.copy(), distance=3, intoObs=True)
    if case == "mac_d5_phiobs":
        return I.extrapolate_mac_simple(flags, vel.copy(), distance=5, phiObs=phiObs)
    if case == "mac_d0":
        return I.extrapolate_mac_simple(flags, vel.copy(), distance=0)
    if case.startswith("ls_"):
        return I.extrapolate_ls_simple(phi.copy(), distance=int(case[-1]), inside="_in_" in case)
    if case.startswith("v3_"):
        return I.extrapolate_vec3_simple(vel.copy(), phi, distance=int(case[-1]), inside="_in_" in case)
    if case == "from_levelset":
        return I.update_from_levelset(flags.copy(), phi)
    if case == "bound_real_w0":
        return I.set_bound(phi.copy(), 0.5, 0)
    if case == "bound_real_w2":
        return I.set_bound(phi.copy(), -3.0, 2)
    if case == "bound_vec_w1":
        return I.set_bound(vel.copy(), 0.25, 1)
    if case.startswith("macw"):          # weights as mapPartsToMAC leaves them: positive where a face got a contribution
        w = np.where(np.repeat((phi < 0)[..., None], 3, axis=3), np.abs(vel) + 0.1, 0).astype(vel.dtype)
        v = I.extrapolate_mac_from_weight(vel.copy(), w, distance=int(case[-1]))
        return w if "marks" in case else v
    if case.startswith("fractions") or case.startswith("obsflags"):
        f2 = flags.copy()
        if "open" in case or case.startswith("obsflags"):      # open / outflow / inflow walls instead of solid ones
            f2[:, :, 0] = 32; f2[:, -1, :] = 16 | 4; f2[:, 0, :] = 8 | 1
            if flags.shape[0] > 1:
                f2[-1] = 32; f2[0] = 16
        if case.startswith("fractions"):
            return I.update_fractions(f2, phiObs, boundaryWidth=int(case[-1]))
        if case == "obsflags_frac":
            return I.set_obstacle_flags(f2, phiObs, fractions=I.update_fractions(f2, phiObs, boundaryWidth=0), phiOut=phi, boundaryWidth=1)
        return I.set_obstacle_flags(f2, phiObs, phiIn=phi, boundaryWidth=2)
    if case == "wall_frac":
        return I.set_wall_bcs_frac(flags, vel.copy(), phiObs)
    if case == "laplacian":
        return I.get_laplacian(np.where(phi < -1000, 0, phi).astype(phi.dtype))
    if case.startswith("curvature"):
        return I.get_curvature(np.where(phi < -1000, 0, phi).astype(phi.dtype), 1.0 if case.endswith("h1") else 0.7)
    raise KeyError(case)


def check_liquid_against_golden(I, name, prec):
    g = load_golden("step_" + name, prec)
    flags, vel, phi, phiObs = liquid_scene(name, prec)
    assert np.array_equal(flags, g["flags"]) and np.array_equal(vel, g["vel"]) and np.array_equal(phi, g["phi"])
    for case in LIQUID_CASES:
        out = run_liquid_case(I, case, flags, vel, phi, phiObs)
        if case in LIQUID_ULP_CASES and getattr(I, "kind", "") == "cuda":
            assert np.allclose(out, g[case], rtol=3e-7 if prec == 4 else 4e-15, atol=0), (name, prec, case)
            continue
        assert np.array_equal(out, g[case]), (name, prec, case, float(np.abs(out.astype(np.float64) - g[case]).max()))


FREESURFACE_SCENES = {"fs3d": ((20, 24, 20), 1), "fs2d": ((1, 40, 36), 2)}       # shape, preconditioner (mICP only supports 3-D, conjugategrad.cpp:222)


def run_freesurface_steps(I, name, prec, steps):
    """the main loop of scenes/freesurface.py:54-84 with useMarching = False, ghostFluid = True, doOpen = False: two-cell walls (bWidth 1),
    basin + drop, every plugin of the loop through the implementation under test.  The solver tolerance is tighter than the scene's 5e-4
    (and the iteration cap wider) so that implementations whose iteration counts differ by one still agree to ~1e-5."""
    from mantaflow_b200 import scenes
    shape, pc = FREESURFACE_SCENES[name]
    sz, sy, sx = shape
    real = np.float32 if prec == 4 else np.float64
    flags = scenes.closed_box_flags(sx, sy, sz, boundaryWidth=1)
    k, j, i = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    basin = (j + 0.5) - 0.2 * sy
    drop = np.sqrt((i + 0.5 - 0.5 * sx) ** 2 + (j + 0.5 - 0.5 * sy) ** 2 + ((k + 0.5 - 0.5 * sz) ** 2 if sz > 1 else 0)) - 0.125 * sx
    phi = np.ascontiguousarray(np.minimum(basin, drop).astype(real))
    I.update_from_levelset(flags, phi)
    vel = np.zeros(shape + (3,), real); p = np.zeros(shape, real)
    its = []
    for _ in range(steps):
        I.extrapolate_ls_simple(phi, distance=5, inside=False)
        I.extrapolate_ls_simple(phi, distance=5, inside=True)
        I.extrapolate_mac_simple(flags, vel, distance=5)
        I.advect_semi_lagrange(flags, vel, phi, order=2, clampMode=2)
        I.set_bound(phi, 1.0, 1)
        I.update_from_levelset(flags, phi)
        I.advect_semi_lagrange(flags, vel, vel, order=2)
        I.add_gravity(flags, vel, (0, -0.025, 0))
        I.set_wall_bcs_obvel(flags, vel, None)
        p, it, _ = I.solve_pressure(flags, vel, phi=phi, cgMaxIterFac=5, cgAccuracy=1e-5 if prec == 4 else 1e-10, preconditioner=pc)
        its.append(it)
    return flags, phi, vel, p, its


def check_freesurface_against_golden(I, name, prec, tol):
    """six steps of the free-surface loop: identical flags, iteration counts within one, fields within tol (0: bit-identical)"""
    g = load_golden("step_" + name, prec)
    flags, phi, vel, p, its = run_freesurface_steps(I, name, prec, steps=6)
    assert np.array_equal(flags, g["flags"]), "fluid / empty cells differ after six steps"
    assert all(abs(a - int(b)) <= 1 for a, b in zip(its, g["iterations"])), (its, g["iterations"])
    for a, key, f in ((phi, "phi", 1), (vel, "vel", 1), (p, "pressure", 10)):
        err = float(np.abs(a.astype(np.float64) - g[key]).max())
        assert err <= tol * f, (name, prec, key, err)


# ---------------------------------------------------------------- size-independent properties of the projection (checked at BASELINE's full sizes on the GPU)
def check_projection_properties(I, res, prec, preconditioners, accuracy=1e-4, random_vel=True, seed=5, reference=None):
    """What must hold at ANY grid size, with no second implementation at hand (the reference cannot run 512^3 inside a test):
    (1) ApplyMatrix is linear and symmetric (x.Ay == y.Ax), identity rows included;
    (2) the right-hand side of a closed box is compatible: sum(rhs) ~ 0, and the device reduction agrees with the grid it reduced;
    (3) after solvePressure the divergence of every fluid cell is below the solver's max-norm tolerance (the residual of the last
        iterate IS the divergence after correctVelocity), for every preconditioner; a closed box is singular, so with a pinned cell the
        pinned row alone may carry the sum of all other residuals;
    (4) projecting an already projected field changes it by no more than the tolerance allows (idempotence).
    `reference` = {preconditioner: what the UNMODIFIED reference left behind on the same input (tests/golden/fullsize_divergence.json,
    written by tools/ref_fullsize_divergence.py)}: over ~1600 float iterations the TRUE residual b - A x drifts away from the recursively
    updated one that meets the tolerance (512^3 PcNone: max |div| 3.6e-4 at cgAccuracy 1e-4 in the reference itself), so there the bound
    of (3) is the reference's own divergence -- north_star: "post-projection max divergence at or below the reference's" -- and, in float,
    iteration count and the SHA-1 of the pressure grid must equal the reference's.
    Returns {preconditioner: iterations}."""
    from mantaflow_b200 import scenes
    real = np.float32 if prec == 4 else np.float64
    eps = 1.2e-7 if prec == 4 else 2.3e-16
    flags, vel = scenes.smoke_plume(res, prec, random_vel=random_vel)
    shape = flags.shape
    fluid = (flags & 1) != 0
    rng = np.random.default_rng(seed)
    dot = lambda a, b: float(np.dot(a.ravel().astype(np.float64), b.ravel().astype(np.float64)))
    # (1)
    A = I.make_matrix(flags)
    x = (rng.random(shape) - 0.5).astype(real); y = (rng.random(shape) - 0.5).astype(real)
    Ax, Ay = I.apply_matrix(flags, x, *A), I.apply_matrix(flags, y, *A)
    a = real(0.375)                                   # exactly representable: a*x carries no rounding of its own
    Axy = I.apply_matrix(flags, (a * x + y).astype(real), *A)
    lin = float(np.abs(Axy.astype(np.float64) - (float(a) * Ax.astype(np.float64) + Ay.astype(np.float64))).max())
    assert lin <= 64 * eps * 12, ("ApplyMatrix not linear", lin)             # |A| rows sum to <= 12, |x| <= 1
    sxy, syx = dot(x, Ay), dot(y, Ax)
    assert abs(sxy - syx) <= 1e3 * eps * (np.sqrt(dot(x, x) * dot(Ay, Ay)) + 1), ("ApplyMatrix not symmetric", sxy, syx)
    assert np.array_equal(Ax[~fluid], x[~fluid]), "non-fluid rows are not the identity"
    del Ax, Ay, Axy, x, y, A
    # (2)
    rhs, s, cnt = I.compute_rhs(flags, vel)
    assert cnt == int(fluid.sum())
    absum = float(np.abs(rhs.astype(np.float64)).sum())
    assert abs(s) <= 1e-5 * absum + 1e-6, ("closed box, wall conditions set: the fluxes must cancel", s, absum)
    assert abs(float(rhs.astype(np.float64).sum()) - s) <= 1e-6 * absum + 1e-9, "the reduction disagrees with the grid it reduced"
    its = {}
    for pc in preconditioners:
        # (3)
        v = vel.copy()
        p, it, rn = I.solve_pressure(flags, v, cgAccuracy=accuracy, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=(pc >= 2))
        its[pc] = it
        assert np.isfinite(p).all() and rn <= accuracy, (pc, it, rn)
        div, _, _ = I.compute_rhs(flags, v)
        d = np.abs(div[fluid].astype(np.float64))
        pinned = (pc >= 2) or accuracy < 1e-7         # plugin/pressure.cpp:349: zeroPressureFixing || cgAccuracy < 1e-7 pins a cell
        over = int((d > 2 * accuracy).sum())
        ref = (reference or {}).get(pc)
        if ref is not None:
            assert it == ref["iterations"], ("iteration count differs from the reference's", pc, it, ref["iterations"])
            assert float(d.max()) <= ref["max_div"] * (1 + 1e-6) and over <= ref["cells_over_2acc"], ("divergence above the reference's", pc, over, float(d.max()), ref)
            if prec == 4 and "pressure_sha1" in ref:
                import hashlib
                assert hashlib.sha1(np.ascontiguousarray(p).tobytes()).hexdigest() == ref["pressure_sha1"], "the pressure grid is not the reference's bit for bit"
        else:
            assert over <= (1 if pinned else 0), ("divergence above the solver tolerance after the projection", pc, over, float(d.max()))
        assert np.array_equal(div[~fluid], np.zeros_like(div[~fluid]))
        # (4)
        v2 = v.copy()
        p2, it2, _ = I.solve_pressure(flags, v2, cgAccuracy=accuracy, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=(pc >= 2))
        assert it2 <= max(2, it // 4), ("a projected field needed a full solve again", pc, it, it2)
        d2 = np.abs(I.compute_rhs(flags, v2)[0][fluid].astype(np.float64))
        if ref is not None:
            assert float(d2.max()) <= max(2 * accuracy, ref["max_div"] * (1 + 1e-6)), ("second projection", pc, float(d2.max()))
        else:
            assert int((d2 > 2 * accuracy).sum()) <= (1 if pinned else 0), ("second projection", pc, float(d2.max()))
    return its


# ---------------------------------------------------------------- second-order obstacle boundaries: the scenario of tools/tests/test_1040_secOrderBnd.py
SECORDER_SCENES = {"sob2d": (1, 32, 32), "sob3d": (18, 20, 22)}
# the reference test is 2-D and runs 10 steps; the 3-D variant of the set-up amplifies last-bit differences by ~1e3 per step from step 5 on
# (measured with a solver perturbed by 1 ulp), so it is compared after 4 steps
SECORDER_STEPS = {"sob2d": 10, "sob3d": 4}


def run_sec_order_bnd(I, name, prec, steps=None):
    """test_1040_secOrderBnd.py:20-66: a spherical container given by an obstacle level set, fill fractions from it, obstacle flags from the
    fractions, a vortex inside; every step advects density and velocity (MacCormack, clampMode 1), applies the second-order wall conditions
    (setWallBcs with fractions + phiObs), extrapolates one cell, projects with the fractions, and repeats the wall conditions.
    The level set and the initial vortex are built in numpy (Sphere.computeLevelset / initVortexVelocity are scene set-up, not on the path)."""
    from mantaflow_b200 import scenes
    steps = SECORDER_STEPS[name] if steps is None else steps
    sz, sy, sx = shape = SECORDER_SCENES[name]
    real = np.float32 if prec == 4 else np.float64
    k, j, i = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    c = np.array([0.5 * sx, 0.5 * sy, 0.5 * sz if sz > 1 else 0.5])
    radius = 0.4 * min(sx, sy)
    dist = np.sqrt((i + 0.5 - c[0]) ** 2 + (j + 0.5 - c[1]) ** 2 + ((k + 0.5 - c[2]) ** 2 if sz > 1 else 0))
    phiObs = np.ascontiguousarray((-(dist - radius)).astype(real))          # sphere.computeLevelset(); phiObs.multConst(-1)
    vel = np.zeros(shape + (3,), real)
    inside = phiObs >= -1.0                                                  # kninitVortexVelocity initplugins.cpp:478-498
    dx = i - c[0]; dx = np.where(dx >= 0, dx - .5, dx + .5); dy = j - c[1]
    vel[..., 0] = np.where(inside, -np.sin(np.arctan2(dy, dx)) * (np.sqrt(dx * dx + dy * dy) / radius), 0)
    dx = i - c[0]; dy = j - c[1]; dy = np.where(dy >= 0, dy - .5, dy + .5)
    vel[..., 1] = np.where(inside, np.cos(np.arctan2(dy, dx)) * (np.sqrt(dx * dx + dy * dy) / radius), 0)
    # the exactly symmetric vortex puts back-traced positions ON cell boundaries, where the last bit of a solve decides which cells the
    # MacCormack clamp looks at; a small seeded perturbation removes those ties so that implementations can be compared by tolerance
    rng = np.random.default_rng(17)
    vel[..., :2] += np.where(inside[..., None], 0.02 * (rng.random(shape + (2,)) - 0.5), 0).astype(real)
    flags = scenes.closed_box_flags(sx, sy, sz)
    flags[flags == 1] = 4                                                    # flags.initDomain(): empty inside the walls
    fractions = I.update_fractions(flags, phiObs)
    I.set_obstacle_flags(flags, phiObs, fractions=fractions)
    flags[(flags & (2 | 8 | 16 | 32)) == 0] = 1                              # flags.fillGrid()
    dens = np.zeros(shape, real); dens[inside & (j > 0.5 * sy)] = 1
    its = []
    p = None
    for _ in range(steps):
        I.advect_semi_lagrange(flags, vel, dens, order=2, clampMode=1)
        I.advect_semi_lagrange(flags, vel, vel, order=2, strength=1.0, clampMode=1)
        I.set_wall_bcs_frac(flags, vel, phiObs)
        I.extrapolate_mac_simple(flags, vel, distance=1)
        p, it, _ = I.solve_pressure(flags, vel, fractions=fractions, cgMaxIterFac=5, cgAccuracy=1e-5 if prec == 4 else 1e-10,
                                    preconditioner=1 if sz > 1 else 2, zeroPressureFixing=(sz == 1))
        its.append(it)
        I.set_wall_bcs_frac(flags, vel, phiObs)
        I.extrapolate_mac_simple(flags, vel, distance=1)
    return flags, fractions, dens, vel, p, its


def check_sec_order_bnd_against_golden(I, name, prec, tol):
    g = load_golden("step_" + name, prec)
    flags, fractions, dens, vel, p, its = run_sec_order_bnd(I, name, prec)
    assert np.array_equal(flags, g["flags"]) and np.array_equal(fractions, g["fractions"]), "flags / fill fractions differ"
    assert all(abs(a - int(b)) <= 1 for a, b in zip(its, g["iterations"])), (its, g["iterations"])
    for a, key, f in ((dens, "density", 1), (vel, "vel", 1), (p, "pressure", 10)):
        err = float(np.abs(a.astype(np.float64) - g[key]).max())
        assert err <= tol * f, (name, prec, key, err)


# ---------------------------------------------------------------- FLIP particle <-> grid plugins (plugin/flip.cpp), oracle side (the device versions come next)
FLIP_SCENES = {"flip3d": (10, 14, 12), "flip2d": (1, 20, 18)}


def flip_scene(name, prec):
    """particles sampled in a basin + drop (about 4 per liquid cell in 2-D, 6 in 3-D), some outside the domain, some deleted, a few typed
    as FlagEmpty (excluded); random particle velocities; an obstacle level set"""
    from mantaflow_b200 import scenes
    sz, sy, sx = shape = FLIP_SCENES[name]
    real = np.float32 if prec == 4 else np.float64
    flags, _, phi = scenes.liquid_basin((sx, sy, sz), prec)
    rng = np.random.default_rng(23 + sx)
    k, j, i = np.nonzero(phi < 0)
    per = 6 if sz > 1 else 4
    base = np.repeat(np.stack([i, j, k], 1), per, 0).astype(np.float64)
    pos = base + rng.random(base.shape)
    if sz == 1:
        pos[:, 2] = 0.5
    pos[rng.random(len(pos)) < 0.01] += np.array([sx, 0, 0])            # left the domain
    pos[rng.random(len(pos)) < 0.01] *= -1
    pos = np.ascontiguousarray(pos.astype(real))
    pflag = np.zeros(len(pos), np.int32)
    pflag[rng.random(len(pos)) < 0.03] |= 1 << 10                       # PDELETE
    pflag[rng.random(len(pos)) < 0.05] |= 1                             # PNEW: still active
    ptype = np.where(rng.random(len(pos)) < 0.1, 4, 1).astype(np.int32)  # FlagEmpty-typed particles are excluded by the scenes
    pvel = (rng.random(pos.shape) * 2 - 1).astype(real)
    if sz == 1:
        pvel[:, 2] = 0
    phiObs = (rng.random(shape) * 6 - 2).astype(real)
    return flags, pos, pflag, ptype, pvel, phiObs


def run_flip_plugins(I, name, prec):
    """every plugin once, on the outputs of the previous ones where the scene chains them (scenes/benchmark_dam.py:100-125)"""
    flags, pos, pflag, ptype, pvel, phiObs = flip_scene(name, prec)
    shape = flags.shape
    out = {}
    out["mark"] = I.mark_fluid_cells(flags.copy(), pos, pflag, ptype=ptype, exclude=4)
    out["mark_phiobs"] = I.mark_fluid_cells(flags.copy(), pos, pflag, phiObs=phiObs)
    index, isys = I.grid_particle_index(shape, pos, pflag)
    out["index"], out["index_sys"] = index, isys
    out["union"] = I.union_particle_levelset(pos, index, isys, radiusFactor=1.0)
    out["union_excl"] = I.union_particle_levelset(pos, index, isys, radiusFactor=1.5, ptype=ptype, exclude=4)
    vel, velOld, w = I.map_parts_to_mac(shape, pos, pflag, pvel, want_weight=True, ptype=ptype, exclude=4)
    out["map_vel"], out["map_weight"] = vel, w
    assert np.array_equal(vel, velOld)
    out["map_vel_noweight"] = I.map_parts_to_mac(shape, pos, pflag, pvel)[0]
    rng = np.random.default_rng(5)
    vnew = (vel + (rng.random(vel.shape) - 0.5).astype(vel.dtype) * 0.1).astype(vel.dtype)
    out["pic"] = I.flip_velocity_update(vnew, vel, pos, pflag, pvel.copy(), -1.0, ptype=ptype, exclude=4)
    out["flip"] = I.flip_velocity_update(vnew, vel, pos, pflag, pvel.copy(), 0.97, ptype=ptype, exclude=4)
    return out


# ---------------------------------------------------------------- IC(0) preconditioner (PC_ICP, conjugategrad.cpp:26-63,:109-132)
ICP_SCENES = ["smoke16", "liquid14"]          # the 3-D systems of the kernels_* fixtures (ICP only supports 3-D grids, conjugategrad.cpp:218)


def run_icp(I, g, prec):
    """factor, one application and a GridCg solve with PC_ICP (pc = 3 in the oracle interface) on the system of a kernels_* fixture"""
    flags, A = g["flags"], [g[n] for n in "A0 Ai Aj Ak".split()]
    P = I.ic_init(flags, *A)
    out = {"ic_P%s" % n: p for n, p in zip("0ijk", P)}
    out["ic_apply"] = I.ic_apply(flags, g["src"], *P)
    x, it, rn = I.cg_solve(flags, g["rhs"], *A, pc=3, accuracy=1e-5 if prec == 4 else 1e-11, maxIter=4000)
    out.update(cg_ic_x=x, cg_ic_it=np.array(it), cg_ic_res=np.array(rn))
    return out


def check_icp_against_golden(I, name, prec, exact_reductions):
    g, k = load_golden("icp_" + name, prec), load_golden(name, prec)
    out = run_icp(I, k, prec)
    fluid = (k["flags"] & 1) != 0
    for n in "0ijk":
        assert np.array_equal(out["ic_P" + n], g["ic_P" + n]), "IC factor P%s not bit-exact" % n
    assert np.array_equal(out["ic_apply"][fluid], g["ic_apply"][fluid]), "IC sweeps not bit-exact"
    assert abs(int(out["cg_ic_it"]) - int(g["cg_ic_it"])) <= 1
    assert rel_l2(out["cg_ic_x"], g["cg_ic_x"]) <= (1e-4 if prec == 4 else 1e-10)
    if exact_reductions:
        assert int(out["cg_ic_it"]) == int(g["cg_ic_it"]) and np.array_equal(out["cg_ic_x"], g["cg_ic_x"])
    # the fixture is not vacuous: ICP beats the unpreconditioned solve
    assert int(g["cg_ic_it"]) < int(k["cg_none_it"])


# ---------------------------------------------------------------- ParticleSystem::advectInGrid (particle.h:512-536)
ADVECT_CASES = {   # integrationMode, deleteInObstacle, stopInObstacle, skipNew, use ptype / exclude
    "rk4_flip": (2, False, True, False, True),        # the call of scenes/benchmark_dam.py:121 / flip02_surface.py
    "rk4_tracer": (2, True, True, False, False),      # the defaults (tracer particles are deleted in obstacles)
    "rk2_nostop": (1, False, False, True, False),
    "euler_delete_nostop": (0, True, False, False, True),
}


def advect_scene(name, prec):
    """the particles of flip_scene (active ones inside the domain; a few deleted), an obstacle block inside the basin, and a velocity field
    fast enough (up to 2.5 cells per step) to push particles into the walls and the block"""
    flags, pos, pflag, ptype, pvel, _ = flip_scene(name, prec)
    sz, sy, sx = flags.shape
    real = np.float32 if prec == 4 else np.float64
    flags = flags.copy()
    flags[(sz // 3 if sz > 1 else 0):(2 * sz // 3 if sz > 1 else 1), 2:sy // 4, sx // 3:sx // 2] = 2       # obstacle block
    nd = 3 if sz > 1 else 2
    inside = np.all((pos[:, :nd] >= 1) & (pos[:, :nd] < np.array([sx - 1, sy - 1, sz - 1])[:nd]), axis=1)
    pos, pflag, ptype = np.ascontiguousarray(pos[inside]), np.ascontiguousarray(pflag[inside]), np.ascontiguousarray(ptype[inside])
    rng = np.random.default_rng(77 + sx)
    vel = ((rng.random(flags.shape + (3,)) - 0.5) * 5).astype(real)
    if sz == 1:
        vel[..., 2] = 0
    return flags, vel, pos, pflag, ptype


def run_advect_cases(I, name, prec):
    flags, vel, pos, pflag, ptype = advect_scene(name, prec)
    out = {}
    for case, (mode, dele, stop, skip, typed) in ADVECT_CASES.items():
        p, f = I.advect_in_grid(flags, vel, pos.copy(), pflag.copy(), 0.8, integrationMode=mode, deleteInObstacle=dele, stopInObstacle=stop, skipNew=skip,
                                ptype=ptype if typed else None, exclude=4 if typed else 0)
        out[case + "_pos"], out[case + "_flag"] = p, f
    # pushOutofObs (flip.cpp:528-545) and projectOutOfBnd (particle.h:565-590) on the advected particles, as scenes/benchmark_dam.py:123-124 chains them
    sz, sy, sx = flags.shape
    phiObs = obstacle_levelset(flags.shape, prec)
    moved, mflag = out["rk2_nostop_pos"], out["rk2_nostop_flag"]          # some of these left the domain or sit inside the block
    out["project"] = I.project_out_of_bnd(flags.shape, moved.copy(), mflag, 1.5, plane="xXyYzZ", ptype=ptype, exclude=4)
    out["project_xY"] = I.project_out_of_bnd(flags.shape, moved.copy(), mflag, 2.25, plane="xY")
    out["push"] = I.push_out_of_obs(flags.shape, out["project"].copy(), mflag, phiObs, shift=0.0, thresh=0.5, ptype=ptype, exclude=4)
    out["push_shift"] = I.push_out_of_obs(flags.shape, moved.copy(), mflag, phiObs, shift=0.25, thresh=0.0)
    # the Lagrangian-particle helpers of scenes/benchmark_dam.py:118-134 (plugin/ptsplugins.cpp, grid.cpp:866-890); type 4 plays the free particles
    pv0 = (np.random.default_rng(31).random(pos.shape) * 2 - 1).astype(pos.dtype)
    if sz == 1:
        pv0[:, 2] = 0
    out["lag_force"] = I.add_force_pvel(pv0.copy(), (0.0, -0.3, 0.1 if sz > 1 else 0.0), 0.8, ptype=ptype, exclude=1)
    out["lag_force_all"] = I.add_force_pvel(pv0.copy(), (0.2, -0.3, 0.0), 0.7)
    out["lag_euler"] = I.euler_step(pos.copy(), pv0, 0.8, ptype=ptype, exclude=1)
    out["lag_delta"] = I.update_velocity_from_delta_pos(out["lag_euler"], pv0.copy(), pos, 0.8, ptype=ptype, exclude=1)
    sparse = I.mark_fluid_cells(flags.copy(), np.ascontiguousarray(pos[::5]), np.ascontiguousarray(pflag[::5]))       # sparse particles leave isolated fluid cells
    out["lag_type1"] = I.set_part_type(sparse, pos, ptype.copy(), 1, 4, 1)
    out["lag_isolated"] = I.mark_isolated_fluid_cell(sparse.copy(), 4)
    out["lag_type2"] = I.set_part_type(out["lag_isolated"], pos, out["lag_type1"].copy(), 4, 1, 4)
    return out


def obstacle_levelset(shape, prec):
    """signed distance to the walls of the box (one cell thick) and to a sphere standing where advect_scene puts its obstacle block"""
    sz, sy, sx = shape
    k, j, i = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    x, y, z = i + 0.5, j + 0.5, k + 0.5
    d = np.minimum(np.minimum(x - 1, sx - 1 - x), np.minimum(y - 1, sy - 1 - y))
    if sz > 1:
        d = np.minimum(d, np.minimum(z - 1, sz - 1 - z))
    cz = 0.5 * sz if sz > 1 else 0.5
    sphere = np.sqrt((x - 5 * sx / 12.0) ** 2 + (y - (1 + sy / 8.0)) ** 2 + ((z - cz) ** 2 if sz > 1 else 0)) - sx / 8.0
    return np.ascontiguousarray(np.minimum(d, sphere).astype(np.float32 if prec == 4 else np.float64))


# ---------------------------------------------------------------- the main loop of scenes/benchmark_dam.py:100-134 (the reference's FLIP benchmark)
def run_dam_loop(I, prec, steps=12, shape=(12, 20, 16)):
    """every plugin call of that loop, in its order and with its arguments (ghost-fluid variant, `gfm`), on a small dam: particles typed
    FlagFluid take part in the grid solve, particles typed FlagEmpty fly ballistically (the scene's `exclude` arguments)"""
    from mantaflow_b200 import scenes
    sz, sy, sx = shape
    real = np.float32 if prec == 4 else np.float64
    flags = scenes.closed_box_flags(sx, sy, sz)
    rng = np.random.default_rng(11)
    k, j, i = np.nonzero((flags & 2) == 0)
    dam = (i < sx // 3) & (j < 2 * sy // 3)                                    # a column of liquid in one corner
    base = np.repeat(np.stack([i[dam], j[dam], k[dam]], 1), 4, 0).astype(np.float64)
    pos = np.ascontiguousarray((base + rng.random(base.shape)).astype(real))
    spray = (np.array([sx // 2, 2 * sy // 3, sz // 3]) + rng.random((60, 3)) * np.array([sx // 3, sy // 5, sz // 3])).astype(real)      # free particles above the floor
    pos = np.ascontiguousarray(np.concatenate([pos, spray]))
    pflag = np.zeros(len(pos), np.int32)
    ptype = np.full(len(pos), 1, np.int32)
    ptype[-len(spray):] = 4
    pvel = np.zeros_like(pos)
    pvel[-len(spray):] = ((rng.random((len(spray), 3)) - 0.5) * 0.6).astype(real)
    phiObs = obstacle_levelset(shape, prec)
    dt, grav, dx = 0.8, -0.1, 1.0
    flags = I.mark_fluid_cells(flags, pos, pflag, ptype=ptype)
    its = []
    for _ in range(steps):
        vel, velOld = I.map_parts_to_mac(shape, pos, pflag, pvel, ptype=ptype, exclude=4)
        I.add_gravity(flags, vel, (0, grav, 0), scale=False, dt=dt)
        index, isys = I.grid_particle_index(shape, pos, pflag)
        phi = I.union_particle_levelset(pos, index, isys, radiusFactor=1.0)
        I.extrapolate_ls_simple(phi, distance=4, inside=True)
        I.set_wall_bcs_obvel(flags, vel)
        p, it, _ = I.solve_pressure(flags, vel, phi=phi, cgAccuracy=1e-5 if prec == 4 else 1e-10, cgMaxIterFac=3.0)
        its.append(it)
        I.set_wall_bcs_obvel(flags, vel)
        I.extrapolate_mac_simple(flags, vel)
        I.flip_velocity_update(vel, velOld, pos, pflag, pvel, 0.97, ptype=ptype, exclude=4)
        I.add_force_pvel(pvel, (0, grav, 0), dt, ptype=ptype, exclude=1)
        x_prev = pos.copy()                                                   # pp.getPosPdata(target=pVtmp)
        I.advect_in_grid(flags, vel, pos, pflag, dt, integrationMode=2, deleteInObstacle=False, ptype=ptype, exclude=4)
        I.euler_step(pos, pvel, dt, ptype=ptype, exclude=1)
        I.project_out_of_bnd(shape, pos, pflag, 1 + dx * 0.5, plane="xXyYzZ", ptype=ptype)
        I.push_out_of_obs(shape, pos, pflag, phiObs, thresh=dx * 0.5, ptype=ptype)
        I.update_velocity_from_delta_pos(pos, pvel, x_prev, dt, ptype=ptype, exclude=1)
        I.mark_fluid_cells(flags, pos, pflag, ptype=ptype)
        I.set_part_type(flags, pos, ptype, 1, 4, 1)
        I.mark_isolated_fluid_cell(flags, 4)
        I.set_part_type(flags, pos, ptype, 4, 1, 4)
    return dict(flags=flags, pos=pos, pvel=pvel, ptype=ptype, vel=vel, phi=phi, pressure=p, iterations=np.array(its))


# ---------------------------------------------------------------- element-wise Grid<T> arithmetic (grid.cpp:258-284, grid.h:472-480)
GRID_OPS = ["setConst", "addConst", "multConst", "add", "sub", "mult", "addScaled", "clamp", "stomp", "safeDivide"]


def grid_arith_case(op, dtype, comps, seed=3):
    """inputs and the numpy restatement of one operation: (me, other | None, constant (x, y, z), expected)"""
    rng = np.random.default_rng(seed + comps)
    shape = (5, 6, 7) + ((3,) if comps == 3 else ())
    if dtype == np.int32:
        me, other = rng.integers(-9, 9, shape).astype(np.int32), rng.integers(-3, 3, shape).astype(np.int32)
        c = (3.0, -2.0, 5.0)
    else:
        me, other = ((rng.random(shape) - 0.5) * 4).astype(dtype), ((rng.random(shape) - 0.5) * 4).astype(dtype)
        other[rng.random(shape) < 0.2] = 0
        c = (0.3, -1.7, 2.9)
    cv = np.array(c[:3] if comps == 3 else c[:1], dtype=dtype)           # T(value): one constant per component
    lo, hi = dtype(-1), dtype(1.25) if dtype != np.int32 else dtype(2)
    with np.errstate(divide="ignore", invalid="ignore"):
        safe_div = (me // other if dtype != np.int32 else (np.trunc(me / np.where(other == 0, 1, other))).astype(np.int32)) if dtype == np.int32 else me / other
    want = {"setConst": np.broadcast_to(cv, shape).astype(dtype), "addConst": me + cv, "multConst": me * cv, "add": me + other, "sub": me - other, "mult": me * other,
            "addScaled": me + cv * other, "clamp": np.minimum(np.maximum(me, lo), hi), "stomp": np.where(me < cv, dtype(0), me),
            "safeDivide": np.where(other != 0, safe_div, me)}[op].astype(dtype)
    binary = op in ("add", "sub", "mult", "addScaled", "safeDivide")
    const = (float(lo), float(hi), 0.0) if op == "clamp" else c
    return me, (other if binary else None), const, np.ascontiguousarray(want)


# ---------------------------------------------------------------------------------------------- VICintegration, grid half
VIC_CASES = [(True, 1), (True, 2), (False, 1), (False, 2)]          # (vel is a MACGrid, precondition: 1 PC_ICP, 2 PC_mICP)


def vic_scene(prec, res=24):
    """a closed box whose top layers are empty (Dirichlet cells: the Poisson problem is regular) and a vortex sheet of 60 triangles kept so
    far from every non-fluid cell that the curl vanishes there -- on non-fluid cells the reference's IC / MIC sweeps leave the stale content of
    their output grid, and a right-hand side that is non-zero there stalls its preconditioned CG (measured: residual stuck at 0.059)"""
    import mantaflow_b200.scenes as scenes
    real = np.float32 if prec == 4 else np.float64
    flags, vel = scenes.smoke_plume(res, prec, obstacle=False)
    top = flags[:, -4:-1, 1:-1]
    top[top == 1] = 4
    rng = np.random.default_rng(1)
    c = np.array([res * 0.5, res * 0.45, res * 0.5])
    tri, vort = [], []
    for _ in range(60):
        a = rng.random(3) * res * 0.3 + res * 0.33
        tri.append(np.stack([a, a + rng.random(3) * 2 - 1, a + rng.random(3) * 2 - 1]))
        vort.append(np.cross(a - c, [0.3, 1.0, 0.2]) * 0.2)
    return flags, np.zeros_like(vel), np.array(tri, real), np.array(vort, real)


def vic_accuracy(prec):
    return 1e-6 if prec == 4 else 1e-12


def run_vic_reference(R, prec):
    """the fixture: VICintegration of the UNMODIFIED reference on vic_scene -> the vorticity grid its Peskin kernel leaves and, per case, the velocity"""
    flags, vel0, tri, tv = vic_scene(prec)
    out = {"flags": flags}
    for mac, pc in VIC_CASES:
        vort, vel, its = R.vic_integration(flags, tri, tv, 2.0, vel0, velIsMac=mac, cgMaxIterFac=5, cgAccuracy=vic_accuracy(prec), scale=0.01, precondition=pc)
        tag = "%s_pc%d" % ("mac" if mac else "vec", pc)
        out["vorticity"] = vort
        out["vel_" + tag] = vel
        out["its_" + tag] = np.array(its)
    return out


def check_vic_against_golden(I, prec, exact_reductions):
    """I.vic_poisson (the port, or the CUDA path) on the golden vorticity grid against the reference's velocity"""
    g = load_golden("vic_sheet24", prec)
    flags, vort = g["flags"], g["vorticity"]
    for mac, pc in VIC_CASES:
        tag = "%s_pc%d" % ("mac" if mac else "vec", pc)
        vel, its = I.vic_poisson(flags, vort, np.zeros(flags.shape + (3,), vort.dtype), velIsMac=mac, cgMaxIterFac=5, cgAccuracy=vic_accuracy(prec), scale=0.01,
                                 precondition=pc)
        ref, rits = g["vel_" + tag], list(g["its_" + tag])
        if exact_reductions:
            assert its == rits and np.array_equal(vel, ref), tag
        else:
            assert all(abs(a - b) <= 1 for a, b in zip(its, rits)), (tag, its, rits)
            assert rel_l2(vel, ref) <= (1e-4 if prec == 4 else 1e-10), (tag, rel_l2(vel, ref))

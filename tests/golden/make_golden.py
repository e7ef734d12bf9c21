"""Generates the golden vectors in tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by `make -C oracle ref`).  The reference ships no golden data (tools/testdata/readme.txt:1), so
these are the reference's own outputs on seeded inputs, produced in the build container; the fixtures travel to the
GPU box, the reference tree does not.

    python tests/golden/make_golden.py            # rewrites every fixture

Fixtures
  kernels_<scene>_f{32,64}.npz : per-kernel outputs (rhs, A0..Ak, pinned system, ApplyMatrix, MIC factor + sweeps,
                                 GridMg vertex types / operators / V-cycle, GridCg runs for PC_None/mICP/MGP,
                                 correctVelocity, the solvePressure plugin with PcMIC / PcMGDynamic)
  step_<scene>_f{32,64}.npz    : setWallBcs (with obvel), addGravity, addBuoyancy, advectSemiLagrange (Real / MAC, order 1 / 2, both clamp
                                 modes, outflow cells) on seeded random boxes; plume{3d,2d}_f32.npz: six steps of the simpleplume main loop;
                                 step_liq*: the extrapolation plugins, updateFromLevelset, setBound; step_fs*: six steps of the free-surface loop; step_sob*: the scenario of test_1040_secOrderBnd.py
  psolve52_f32.npz             : the scenario of tools/tests/test_0100_psolve.py and test_0110_mgsolve.py (52^3 closed box,
                                 box velocity source, solves with PcMIC / PcMGDynamic / PcMGStatic), float build
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from mantaflow_b200 import scenes  # noqa: E402
from oracle.oracle_api import Oracle  # noqa: E402

KERNEL_SCENES = {
    "smoke16": lambda prec: scenes.smoke_plume((16, 14, 12), prec, random_vel=True) + (None,),
    "liquid14": lambda prec: scenes.liquid_basin((14, 16, 12), prec),
    "smoke2d": lambda prec: scenes.smoke_plume((24, 20, 1), prec, random_vel=True) + (None,),
}


def kernels_fixture(R, name, prec):
    flags, vel, phi = KERNEL_SCENES[name](prec)
    sz, sy, sx = flags.shape
    out = dict(flags=flags, vel=vel)
    if phi is not None:
        out["phi"] = phi
    rhs, s, c = R.compute_rhs(flags, vel, phi=phi)
    out.update(rhs=rhs, rhs_sum=s, rhs_cnt=c)
    A = R.make_matrix(flags, phi=phi)
    for n, a in zip("A0 Ai Aj Ak".split(), A):
        out[n] = a
    rng = np.random.Generator(np.random.PCG64(42))
    src = (rng.random(flags.shape) - 0.5).astype(vel.dtype)
    out["src"] = src
    out["apply_matrix"] = R.apply_matrix(flags, src, *A)
    # unpinned GridCg runs (PcNone via GridCg directly: SURVEY F4)
    acc = 1e-5 if prec == 4 else 1e-11
    x, it, rn = R.cg_solve(flags, rhs, *A, pc=0, accuracy=acc, maxIter=4000)
    out.update(cg_none_x=x, cg_none_it=it, cg_none_res=rn)
    if sz > 1:
        P = R.mic_init(flags, *A)
        out["mic_P"] = P
        out["mic_apply"] = R.mic_apply(flags, src, P, *A)
        x, it, rn = R.cg_solve(flags, rhs, *A, pc=1, accuracy=acc, maxIter=4000)
        out.update(cg_mic_x=x, cg_mic_it=it, cg_mic_res=rn)
    # pinned system for multigrid.  The reference does not expose its cell choice; it is recovered from the plugin:
    # solvePressure with zeroPressureFixing pins exactly one cell to p == 0 with an identity row.
    v = vel.copy()
    p_mg, it_mg, rn_mg = R.solve_pressure(flags, v, phi=phi, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=2, zeroPressureFixing=True)
    out.update(plugin_mg_p=p_mg, plugin_mg_vel=v, plugin_mg_it=it_mg, plugin_mg_res=rn_mg)
    if not (flags & 4).any():
        from oracle.oracle_api import Oracle as _O
        fix = _O("port", prec).choose_fix_cell(flags)      # restatement's choice, validated below against the plugin result
        assert fix >= 0 and p_mg.ravel()[fix] == 0
        rhs_f = rhs.copy(); Af = [a.copy() for a in A]
        R.fix_pressure(flags, fix, 0.0, rhs_f, *Af)
        out.update(fix_idx=fix, fix_rhs=rhs_f, fix_A0=Af[0], fix_Ai=Af[1], fix_Aj=Af[2], fix_Ak=Af[3])
        # the pinned system solved through GridCg+GridMg must reproduce the plugin result bit for bit -> the choice is the reference's
        x, it, rn = R.cg_solve(flags, rhs_f, *Af, pc=2, accuracy=acc, maxIter=100)
        same = np.array_equal(x, p_mg) if prec == 4 else np.allclose(x, p_mg, rtol=0, atol=1e-9 * np.abs(p_mg).max())   # f64: OpenMP reduction order
        assert it == it_mg and same, "pinned-cell choice differs from the reference plugin"
    else:
        rhs_f, Af = rhs, A
    R.mg_create(sx, sy, sz)
    R.mg_set_a(*Af)
    nl = R.mg_num_levels()
    out["mg_levels"] = nl
    for l in range(nl):
        out["mg_size_%d" % l] = np.array(R.mg_level_size(l))
        out["mg_type_%d" % l] = R.mg_get("type", l)
        out["mg_a_%d" % l] = R.mg_get("a", l)
    out["mg_vcycle"] = R.mg_vcycle(rhs_f, coarsestAccuracy=1e-9, pre=1, post=1)
    R.mg_destroy()
    pr = rng.random(flags.shape).astype(vel.dtype)
    out["cv_pressure"] = pr
    out["cv_vel"] = R.correct_velocity(flags, vel.copy(), pr, phi=phi)
    if sz > 1:
        v = vel.copy()
        p, it, rn = R.solve_pressure(flags, v, phi=phi, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=1)
        out.update(plugin_mic_p=p, plugin_mic_vel=v, plugin_mic_it=it, plugin_mic_res=rn)
    # cgSolveDiffusion (conjugategrad.cpp:350-423): Real grid with explicit parameters, Vec3 grid with the defaults
    out["diff_real"] = R.cg_solve_diffusion(flags, src.copy(), alpha=0.7, cgMaxIterFac=2.0, cgAccuracy=1e-7)
    out["diff_vec"] = R.cg_solve_diffusion(flags, vel.copy())
    return out


def box_source(res, value, prec, vel=None):
    """velocity `value` in the cells whose centre lies inside Box(p0=gs*(0.3,0.4,0.3), p1=gs*(0.7,0.8,0.7)) (test_0100_psolve.py:29)"""
    real = np.float32 if prec == 4 else np.float64
    if vel is None:
        vel = np.zeros((res, res, res, 3), real)
    c = np.arange(res) + 0.5
    m = lambda lo, hi: (c >= lo * res) & (c <= hi * res)
    inside = m(0.3, 0.7)[:, None, None] & m(0.4, 0.8)[None, :, None] & m(0.3, 0.7)[None, None, :]
    vel[inside] = np.array(value, real)
    return vel


def psolve52(R, prec=4):
    res = 52
    flags = scenes.closed_box_flags(res, res, res)
    out = dict(flags=flags)
    kw = dict(cgMaxIterFac=99, cgAccuracy=1e-4)
    # --- test_0100_psolve.py: two PcMIC solves, the second after setWallBcs ---
    v = box_source(res, (0.15, 0.3, 0.21), prec)
    out["t0100_vel_in0"] = v.copy()
    p, it, rn = R.solve_pressure(flags, v, zeroPressureFixing=False, **kw)
    out.update(t0100_pressure0=p, t0100_vel0=v.copy(), t0100_it0=it)
    v = box_source(res, (1.5, 3, 2.1), prec); scenes.set_wall_bcs(flags, v)
    out["t0100_vel_in1"] = v.copy()
    p, it, rn = R.solve_pressure(flags, v, zeroPressureFixing=False, **kw)
    out.update(t0100_pressure=p, t0100_vel=v.copy(), t0100_it1=it)
    # --- test_0110_mgsolve.py: PcMGDynamic x2, then PcMGStatic x2 on one hierarchy ---
    key = 110
    v = box_source(res, (0.15, 0.3, 0.21), prec)
    p, it, rn = R.solve_pressure(flags, v, zeroPressureFixing=True, preconditioner=2, solver_key=key, **kw)
    out.update(t0110_p0=p, t0110_it0=it)
    v = box_source(res, (1.5, 3, 2.1), prec); scenes.set_wall_bcs(flags, v)
    p, it, rn = R.solve_pressure(flags, v, zeroPressureFixing=True, preconditioner=2, solver_key=key, **kw)
    out.update(t0110_p1=p, t0110_it1=it)
    v = box_source(res, (1.1, 2, -2.1), prec); scenes.set_wall_bcs(flags, v)
    p, it, rn = R.solve_pressure(flags, v, zeroPressureFixing=True, preconditioner=3, solver_key=key, **kw)
    out.update(t0110_it2=it)
    v = box_source(res, (-1.1, -2, 2.1), prec, vel=v); scenes.set_wall_bcs(flags, v)      # applied on top of the projected field (:66)
    out["t0110_vel_in3"] = v.copy()
    p, it, rn = R.solve_pressure(flags, v, zeroPressureFixing=True, preconditioner=3, solver_key=key, **kw)
    out.update(t0110_p2=p, t0110_v2=v.copy(), t0110_it3=it)
    R.release_solver(key)
    return out


def step_fixtures():
    """step_<scene>_f{32,64}.npz: setWallBcs / addGravity / addBuoyancy / advectSemiLagrange outputs of the reference on the seeded
    scenes of tests/helpers.py; plume_f32.npz: six steps of the simpleplume main loop (advect density + velocity with MacCormack,
    setWallBcs, addBuoyancy, solvePressure PcMIC) run by the reference."""
    sys.path.insert(0, os.path.dirname(HERE))
    import helpers
    for prec in (4, 8):
        R = Oracle("reference", prec)
        for name in helpers.STEP_SCENES:
            flags, vel, dens, obvel = helpers.step_scene(name, prec)
            fx = dict(flags=flags, vel=vel)
            for case in helpers.STEP_CASES:
                fx[case] = helpers.run_step_case(R, case, flags, vel, dens, obvel)
            path = os.path.join(HERE, "step_%s_f%d.npz" % (name, prec * 8))
            np.savez_compressed(path, **fx)
            print("%-34s %7.1f KiB" % (os.path.basename(path), os.path.getsize(path) / 1024))
        step_hi_fixtures(R, prec)
        for name in helpers.WE_SCENES:       # cgSolveWE: three implicit wave steps, plain and Crank-Nicolson
            fx = {}
            for cn in (False, True):
                for a, key in zip(helpers.run_wave_steps(R, name, prec, cn), ("ut", "utm1", "out")):
                    fx["%s_cn%d" % (key, int(cn))] = a
            path = os.path.join(HERE, "step_%s_f%d.npz" % (name, prec * 8))
            np.savez_compressed(path, **fx)
            print("%-34s %7.1f KiB" % (os.path.basename(path), os.path.getsize(path) / 1024))
        for name in helpers.GUIDING_SCENES:  # PD_fluid_guiding: the whole primal-dual loop incl. its multigrid solves
            v, p, it = helpers.run_guiding(R, name, prec)
            path = os.path.join(HERE, "step_%s_f%d.npz" % (name, prec * 8))
            np.savez_compressed(path, vel=v, pressure=p, iterations=np.array(it))
            print("%-34s %7.1f KiB  %d PD iterations" % (os.path.basename(path), os.path.getsize(path) / 1024, it))
    for shape, tag in (((24, 36, 24), "3d"), ((1, 48, 32), "2d")):
        dens, vel, p, its = helpers.run_plume_steps(Oracle("reference", 4), shape, 4, steps=6)
        path = os.path.join(HERE, "plume%s_f32.npz" % tag)
        np.savez_compressed(path, density=dens, vel=vel, pressure=p, iterations=np.array(its))
        print("%-34s %7.1f KiB  its %s  max density %.3f" % (os.path.basename(path), os.path.getsize(path) / 1024, its, float(dens.max())))


def step_hi_fixtures(R=None, prec=None):
    """step_hi_<scene>_f{32,64}.npz: the reference's advectSemiLagrange with orderSpace 2 (cubic lookups) and / or orderTrace 2 (explicit midpoint),
    Real and MAC grids, plain and MacCormack, on the seeded scenes of tests/helpers.py"""
    sys.path.insert(0, os.path.dirname(HERE))
    import helpers
    for pr in ((4, 8) if prec is None else (prec,)):
        Rr = Oracle("reference", pr) if R is None or prec is None else R
        for name in helpers.STEP_HI_SCENES:
            flags, vel, dens, obvel = helpers.step_scene(name, pr)
            fx = dict(flags=flags, vel=vel)
            for case in helpers.STEP_HI_CASES:
                fx[case] = helpers.run_step_hi_case(Rr, case, flags, vel, dens, obvel)
            path = os.path.join(HERE, "step_hi_%s_f%d.npz" % (name, pr * 8))
            np.savez_compressed(path, **fx)
            print("%-34s %7.1f KiB" % (os.path.basename(path), os.path.getsize(path) / 1024))


def liquid_fixtures():
    """step_liq*_f{32,64}.npz: the reference's extrapolateMACSimple / extrapolateLsSimple / extrapolateVec3Simple, FlagGrid::updateFromLevelset and
    Grid::setBound on the seeded scenes of tests/helpers.py; step_fs*: six steps of the free-surface loop; step_sob*: the scenario of test_1040_secOrderBnd.py of scenes/freesurface.py:54-84."""
    sys.path.insert(0, os.path.dirname(HERE))
    import helpers
    for prec in (4, 8):
        R = Oracle("reference", prec)
        for name in helpers.LIQUID_SCENES:   # extrapolateMACSimple / LsSimple / Vec3Simple, updateFromLevelset, setBound
            flags, vel, phi, phiObs = helpers.liquid_scene(name, prec)
            fx = dict(flags=flags, vel=vel, phi=phi)
            for case in helpers.LIQUID_CASES:
                fx[case] = helpers.run_liquid_case(R, case, flags, vel, phi, phiObs)
            path = os.path.join(HERE, "step_%s_f%d.npz" % (name, prec * 8))
            np.savez_compressed(path, **fx)
            print("%-34s %7.1f KiB" % (os.path.basename(path), os.path.getsize(path) / 1024))
        for name in helpers.SECORDER_SCENES:      # the scenario of tools/tests/test_1040_secOrderBnd.py (fill fractions, second-order wall conditions)
            flags, fractions, dens, vel, p, its = helpers.run_sec_order_bnd(R, name, prec)
            path = os.path.join(HERE, "step_%s_f%d.npz" % (name, prec * 8))
            np.savez_compressed(path, flags=flags, fractions=fractions, density=dens, vel=vel, pressure=p, iterations=np.array(its))
            print("%-34s %7.1f KiB  its %s  fluid cells %d" % (os.path.basename(path), os.path.getsize(path) / 1024, its, int((flags & 1).sum())))
        for name in helpers.FREESURFACE_SCENES:   # six steps of the level-set free-surface loop (scenes/freesurface.py:54-84)
            flags, phi, vel, p, its = helpers.run_freesurface_steps(R, name, prec, steps=6)
            path = os.path.join(HERE, "step_%s_f%d.npz" % (name, prec * 8))
            np.savez_compressed(path, flags=flags, phi=phi, vel=vel, pressure=p, iterations=np.array(its))
            print("%-34s %7.1f KiB  its %s  fluid cells %d" % (os.path.basename(path), os.path.getsize(path) / 1024, its, int((flags & 1).sum())))


def flip_fixtures():
    """step_flip*_f{32,64}.npz: the reference's markFluidCells, gridParticleIndex, unionParticleLevelset, mapPartsToMAC, mapMACToParts and
    flipVelocityUpdate (plugin/flip.cpp) on the seeded particle scenes of tests/helpers.py"""
    sys.path.insert(0, os.path.dirname(HERE))
    import helpers
    for prec in (4, 8):
        R = Oracle("reference", prec)
        for name in helpers.FLIP_SCENES:
            fx = helpers.run_flip_plugins(R, name, prec)
            path = os.path.join(HERE, "step_%s_f%d.npz" % (name, prec * 8))
            np.savez_compressed(path, **fx)
            print("%-34s %7.1f KiB  %d indexed particles, %d fluid cells" % (os.path.basename(path), os.path.getsize(path) / 1024, len(fx["index_sys"]), int((fx["mark"] & 1).sum())))
            fx = helpers.run_advect_cases(R, name, prec)          # ParticleSystem::advectInGrid on the same particles
            path = os.path.join(HERE, "step_adv_%s_f%d.npz" % (name, prec * 8))
            np.savez_compressed(path, **fx)
            print("%-34s %7.1f KiB  %d particles" % (os.path.basename(path), os.path.getsize(path) / 1024, len(fx["rk4_flip_flag"])))


def dam_fixtures():
    """step_dam_f{32,64}.npz: twelve passes of the main loop of scenes/benchmark_dam.py:100-134 through the unmodified reference (tests/helpers.py::run_dam_loop)"""
    sys.path.insert(0, os.path.dirname(HERE))
    import helpers
    for prec in (4, 8):
        fx = helpers.run_dam_loop(Oracle("reference", prec), prec)
        path = os.path.join(HERE, "step_dam_f%d.npz" % (prec * 8))
        np.savez_compressed(path, **fx)
        print("%-34s %7.1f KiB  its %s  %d particles, %d typed FlagEmpty at the end" % (os.path.basename(path), os.path.getsize(path) / 1024, list(fx["iterations"]), len(fx["pos"]),
                                                                                       int((fx["ptype"] == 4).sum())))


def icp_fixtures():
    """icp_<scene>_f{32,64}.npz: the reference's IC(0) preconditioner (InitPreconditionIncompCholesky / ApplyPreconditionIncompCholesky,
    conjugategrad.cpp:26-63,:109-132) and a GridCg solve with PC_ICP on the systems of the 3-D kernels_* fixtures"""
    sys.path.insert(0, os.path.dirname(HERE))
    import helpers
    for prec in (4, 8):
        R = Oracle("reference", prec)
        for name in helpers.ICP_SCENES:
            fx = helpers.run_icp(R, helpers.load_golden(name, prec), prec)
            path = os.path.join(HERE, "icp_%s_f%d.npz" % (name, prec * 8))
            np.savez_compressed(path, **fx)
            print("%-34s %7.1f KiB  PC_ICP %d its" % (os.path.basename(path), os.path.getsize(path) / 1024, int(fx["cg_ic_it"])))


def vic_fixtures():
    """vic_sheet24_f{32,64}.npz: VICintegration (plugin/vortexplugins.cpp:195-300) of the unmodified reference on helpers.vic_scene: the vorticity grid
    of its Peskin kernel, and the velocity / iteration counts of its Poisson solves for MACGrid and Grid<Vec3> targets with PC_ICP and PC_mICP"""
    sys.path.insert(0, os.path.dirname(HERE))
    import helpers
    for prec in (4, 8):
        fx = helpers.run_vic_reference(Oracle("reference", prec), prec)
        path = os.path.join(HERE, "vic_sheet24_f%d.npz" % (prec * 8))
        np.savez_compressed(path, **fx)
        print("%-34s %7.1f KiB  its %s" % (os.path.basename(path), os.path.getsize(path) / 1024, {k: list(v) for k, v in fx.items() if k.startswith("its_")}))


def io_fixtures():
    """tests/golden/io/ref_<kind>_<2d|3d>_f{32,64}.uni: grid files written by the unmodified reference's Grid<T>::save (fileio/iogrids.cpp)
    from the seeded arrays of tests/test_fileio.py::sample"""
    sys.path.insert(0, os.path.dirname(HERE))
    import test_fileio as T
    os.makedirs(os.path.join(HERE, "io"), exist_ok=True)
    for prec in (4, 8):
        R = Oracle("reference", prec)
        for kind in T.KINDS:
            for tag in T.SHAPES:
                path = os.path.join(HERE, "io", "ref_%s_%s_f%d.uni" % (kind, tag, prec * 8))
                R.grid_file(path, T.sample(kind, tag, prec), kind, load=False)
                print("%-34s %6d B" % (os.path.basename(path), os.path.getsize(path)))


def main():
    if "--only-io" in sys.argv:
        return io_fixtures()
    if "--only-flip" in sys.argv:
        return flip_fixtures()
    if "--only-icp" in sys.argv:
        return icp_fixtures()
    if "--only-vic" in sys.argv:
        return vic_fixtures()
    if "--only-dam" in sys.argv:
        return dam_fixtures()
    if "--only-step-hi" in sys.argv:
        return step_hi_fixtures()
    if "--only-liquid" in sys.argv:
        return liquid_fixtures()
    if "--only-step" in sys.argv:
        step_fixtures()
        return liquid_fixtures()
    step_fixtures()
    liquid_fixtures()
    flip_fixtures()
    io_fixtures()
    icp_fixtures()
    vic_fixtures()
    dam_fixtures()
    for prec in (4, 8):
        R = Oracle("reference", prec)
        for name in KERNEL_SCENES:
            fx = kernels_fixture(R, name, prec)
            path = os.path.join(HERE, "kernels_%s_f%d.npz" % (name, prec * 8))
            np.savez_compressed(path, **fx)
            print("%-34s %7.1f KiB  cg_none %d its%s" % (os.path.basename(path), os.path.getsize(path) / 1024, fx["cg_none_it"],
                                                         ", mic %d, mg %d" % (fx.get("cg_mic_it", -1), fx["plugin_mg_it"])))
    fx = psolve52(Oracle("reference", 4), 4)
    path = os.path.join(HERE, "psolve52_f32.npz")
    np.savez_compressed(path, **fx)
    print("%-34s %7.1f KiB  its %s" % (os.path.basename(path), os.path.getsize(path) / 1024, [int(fx[k]) for k in sorted(fx) if "_it" in k]))


if __name__ == "__main__":
    main()

"""GPU: parity at BASELINE.json's OWN grid sizes, against the UNMODIFIED reference compiled into oracle/_ref (the C restatement only when
_ref is absent): config 1 (scenes/simpleplume.py geometry 64x96x64, PcMIC / PcMGStatic), config 2 (scenes/benchmark_dam.py geometry
88x83x33 liquid with the free surface as ghost-fluid Dirichlet cells, PcMIC / PcMGDynamic), config 3 (256^3 smoke plume with obstacle,
PcNone vs PcMIC vs PcMGStatic, cgAccuracy 1e-4) and a 128^3 cut of it in both precisions.  Launch geometry is size dependent (z-chunks,
TMA tile grid, MIC warp columns, the multigrid's level count), so the small parity grids of test_gpu_parity.py do not cover it.

Per case: right-hand side and A0/Ai/Aj/Ak bit for bit in the same Real; iteration count within 1 (PcNone, PcMIC -- both are the
reference's arithmetic in the reference's order) or within 1 (multigrid, same V-cycle); converged pressure within relative L2 1e-4
(float) / 1e-10 (double); max |div| of the projected field at or below the reference's (north_star).
The reference side costs ~70 s of host CPU at 256^3 on the GPU box's cores (profiles/r1_configs.md)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from mantaflow_b200 import scenes  # noqa: E402

TOL = {4: 1e-4, 8: 1e-10}


def _ref(prec):
    from oracle.oracle_api import Oracle, available
    return Oracle("reference" if available("reference", prec) else "port", prec)


def rel_l2(a, b):
    d = np.linalg.norm((a.astype(np.float64) - b.astype(np.float64)).ravel())
    return float(d / max(np.linalg.norm(b.astype(np.float64).ravel()), 1e-300))


def demean(p, flags):
    fl = (flags & 1) != 0
    q = p.astype(np.float64).copy()
    q[fl] -= q[fl].mean()
    return q


CASES = {
    # name: (scene builder -> flags, vel, phi), prec, [(preconditioner, cgAccuracy, cgMaxIterFac, zeroPressureFixing)]
    "cfg1_64x96x64": (lambda prec: scenes.smoke_plume((64, 96, 64), prec, obstacle=False) + (None,), 4, [(1, 1e-3, 1.5, False), (3, 1e-3, 1.5, True)]),
    "cfg2_88x83x33_phi": (lambda prec: scenes.liquid_basin((88, 83, 33), prec), 4, [(1, 1e-3, 1.5, False), (2, 1e-3, 1.5, False)]),
    "smoke128_f32": (lambda prec: scenes.smoke_plume(128, prec) + (None,), 4, [(0, 1e-4, 99, False), (1, 1e-4, 99, False), (3, 1e-4, 99, True)]),
    # double build: north_star's 1e-10 is a statement about CONVERGED pressures, so the double cases converge to cgAccuracy 1e-10 (where the
    # restatement and the reference agree to 1.5e-12 on this grid for all three preconditioners).  At cgAccuracy 1e-4 the double
    # PcMIC solve of the UNMODIFIED reference is itself not reproducible: 78 iterations with one OpenMP thread count, 81 with another,
    # pressures 1.7e-7 apart in relative L2 (measured here, 128^3) -- the order of its dot-product reduction decides; that case is kept
    # with the reference's own spread as the bar.
    "smoke128_f64": (lambda prec: scenes.smoke_plume(128, prec) + (None,), 8, [(0, 1e-10, 99, False), (1, 1e-10, 99, False), (3, 1e-10, 99, True), (0, 1e-4, 99, False), (1, 1e-4, 99, False)]),
    "cfg3_256_f32": (lambda prec: scenes.smoke_plume(256, prec) + (None,), 4, [(0, 1e-4, 99, False), (1, 1e-4, 99, False), (3, 1e-4, 99, True)]),
}


def _ref_solve(O, flags, vel, phi, pc, acc, fac, fix):
    v = vel.copy()
    if pc == 0:
        # the reference plugin asserts on PcNone in 3-D: rhs + matrix + GridCg + correctVelocity driven directly
        rhs, _, _ = O.compute_rhs(flags, v, phi=phi)
        A = O.make_matrix(flags, phi=phi)
        if fix or acc < 1e-7:                  # pressure.cpp:349-387: zeroPressureFixing || cgAccuracy < 1e-7 pins a cell of a domain without empty cells
            from oracle.oracle_api import Oracle
            idx = Oracle("port", O.prec).choose_fix_cell(flags)          # integer work; only the restatement exposes the choice
            if idx >= 0:
                O.fix_pressure(flags, idx, 0.0, rhs, *A)                  # the reference's own fixPressure (ref_harness.cpp)
        p, it, rn = O.cg_solve(flags, rhs, *A, pc=0, accuracy=acc, maxIter=int(np.float32(fac) * max(flags.shape)))
        O.correct_velocity(flags, v, p, phi=phi)
    else:
        p, it, rn = O.solve_pressure(flags, v, phi=phi, cgAccuracy=acc, cgMaxIterFac=fac, preconditioner=pc, zeroPressureFixing=fix)
    return p, v, it


@pytest.mark.parametrize("name", list(CASES))
def test_baseline_config_against_reference(name):
    import mantaflow_b200 as mf
    from cuda_impl import CudaImpl
    make, prec, runs = CASES[name]
    flags, vel, phi = make(prec)
    O, I = _ref(prec), CudaImpl(prec)
    # assembly: bit for bit in the same Real
    rhs_o, sum_o, cnt_o = O.compute_rhs(flags, vel, phi=phi)
    rhs_g, sum_g, cnt_g = I.compute_rhs(flags, vel, phi=phi)
    assert np.array_equal(rhs_g, rhs_o) and cnt_g == cnt_o, name
    A_o, A_g = O.make_matrix(flags, phi=phi), I.make_matrix(flags, phi=phi)
    for q, nm in enumerate(("A0", "Ai", "Aj", "Ak")):
        assert np.array_equal(A_g[q], A_o[q]), (name, nm)
    del A_o, A_g, rhs_o, rhs_g
    sz, sy, sx = flags.shape
    s = mf.Solver(gridSize=(sx, sy, sz), dim=3, prec=prec)
    F, V, P = mf.FlagGrid(s, flags), mf.MACGrid(s, vel), mf.RealGrid(s)
    PH = mf.RealGrid(s, phi) if phi is not None else None
    report = []
    for pc, acc, fac, fix in runs:
        p_o, v_o, it_o = _ref_solve(O, flags, vel, phi, pc, acc, fac, fix)
        V.copyFromArray(vel)
        mf.solvePressure(vel=V, pressure=P, flags=F, phi=PH, cgAccuracy=acc, cgMaxIterFac=fac, preconditioner=pc, zeroPressureFixing=fix)
        info = mf.lastSolveInfo()
        p_g, v_g = P.numpy(), V.numpy()
        div_o, div_g = scenes.max_divergence(flags, v_o), scenes.max_divergence(flags, v_g)
        e = rel_l2(demean(p_g, flags), demean(p_o, flags)) if (fix or phi is None) else rel_l2(p_g, p_o)
        report.append((pc, info["iterations"], it_o, e, div_g, div_o))
        print("%s %s pc %d: iterations %d (reference %d) pressure rel-L2 %.2e max|div| %.3e (reference %.3e)" % (name, O.kind, pc, info["iterations"], it_o, e, div_g, div_o))
        loose = prec == 8 and acc > 1e-9 and pc == 1         # see the comment at CASES["smoke128_f64"]
        # the iteration count of the reference's double PcMIC solve moves with the order of its own OpenMP reductions at either accuracy (78 / 81 at
        # 1e-4; at 1e-10: 241 on one B200 box, 239-240 on others; this path: 239 every time), so its own spread is the bar there; the converged
        # pressures are still held to 1e-10 below
        assert abs(info["iterations"] - it_o) <= (4 if (prec == 8 and pc == 1) else 1), report[-1]
        assert e <= (1e-6 if loose else TOL[prec]), report[-1]
        assert rel_l2(v_g, v_o) <= (1e-5 if loose else TOL[prec]), report[-1]      # the velocity correction is a difference of pressures: ~20x the pressure's relative spread
        if phi is None:
            # at or below the reference's; two double solves that stop a step apart both sit below cgAccuracy, either may be the smaller one
            assert div_g <= div_o * (1 + 1e-3) + (1e-7 if prec == 4 else 1e-15) or (prec == 8 and div_g <= acc), report[-1]
        if prec == 4 and pc in (0, 1):
            # float, reference arithmetic in the reference's order: the same bits
            assert info["iterations"] == it_o and np.array_equal(p_g, p_o), report[-1]
    mf.releaseMG(s)
    s.close()

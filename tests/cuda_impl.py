"""Adapter giving the CUDA path (through the C-ABI / host mirror) the same interface as oracle_api.Oracle, so the generic
golden checkers of tests/helpers.py run unchanged against it."""
import numpy as np

import mantaflow_b200 as mf
from mantaflow_b200 import cg


class CudaImpl:
    kind = "cuda"

    def __init__(self, prec):
        self.prec = prec
        self.real = np.float32 if prec == 4 else np.float64
        self._solvers = {}
        self._mg = None

    def _solver(self, flags, key=None):
        sz, sy, sx = flags.shape
        k = (sx, sy, sz, key)
        if k not in self._solvers:
            self._solvers[k] = mf.Solver(gridSize=(sx, sy, sz), dim=3 if sz > 1 else 2, prec=self.prec)
        return self._solvers[k]

    def _g(self, s, cls, a):
        return None if a is None else cls(s, a)

    def compute_rhs(self, flags, vel, phi=None, perCellCorr=None, fractions=None, obvel=None, curv=None, gfClamp=1e-4, surfTens=0.0, enforceCompatibility=False):
        s = self._solver(flags)
        F, V, rhs = mf.FlagGrid(s, flags), mf.MACGrid(s, vel), mf.RealGrid(s)
        sm, cnt = cg.MakeRhs(F, rhs, V, perCellCorr=self._g(s, mf.RealGrid, perCellCorr), fractions=self._g(s, mf.MACGrid, fractions),
                             obvel=self._g(s, mf.MACGrid, obvel), phi=self._g(s, mf.RealGrid, phi), curv=self._g(s, mf.RealGrid, curv),
                             surfTens=surfTens, gfClamp=gfClamp)
        return rhs.numpy().copy(), sm, cnt

    def make_matrix(self, flags, fractions=None, phi=None, gfClamp=1e-4):
        s = self._solver(flags)
        F = mf.FlagGrid(s, flags)
        A = [mf.RealGrid(s) for _ in range(4)]
        cg.MakeLaplaceMatrix(F, *A, fractions=self._g(s, mf.MACGrid, fractions))
        if phi is not None:
            cg.ApplyGhostFluidDiagonal(A[0], F, mf.RealGrid(s, phi), gfClamp)
        return [a.numpy().copy() for a in A]

    def choose_fix_cell(self, flags):
        s = self._solver(flags)
        return cg.chooseFixCell(mf.FlagGrid(s, flags))

    def fix_pressure(self, flags, idx, value, rhs, A0, Ai, Aj, Ak):
        s = self._solver(flags)
        gs = [mf.RealGrid(s, a) for a in (rhs, A0, Ai, Aj, Ak)]
        cg.fixPressure(idx, value, *gs)
        for a, g in zip((rhs, A0, Ai, Aj, Ak), gs):
            a[...] = g.numpy()

    def apply_matrix(self, flags, src, A0, Ai, Aj, Ak):
        s = self._solver(flags)
        D = mf.RealGrid(s)
        cg.ApplyMatrix(mf.FlagGrid(s, flags), D, mf.RealGrid(s, src), *[mf.RealGrid(s, a) for a in (A0, Ai, Aj, Ak)])
        return D.numpy().copy()

    def mic_init(self, flags, A0, Ai, Aj, Ak):
        s = self._solver(flags)
        P = mf.RealGrid(s)
        cg.InitPreconditionModifiedIncompCholesky2(mf.FlagGrid(s, flags), P, *[mf.RealGrid(s, a) for a in (A0, Ai, Aj, Ak)])
        return P.numpy().copy()

    def mic_apply(self, flags, src, P, A0, Ai, Aj, Ak):
        s = self._solver(flags)
        D = mf.RealGrid(s)
        cg.ApplyPreconditionModifiedIncompCholesky2(D, mf.RealGrid(s, src), mf.FlagGrid(s, flags), mf.RealGrid(s, P), *[mf.RealGrid(s, a) for a in (A0, Ai, Aj, Ak)])
        return D.numpy().copy()

    def ic_init(self, flags, A0, Ai, Aj, Ak):
        s = self._solver(flags)
        P = [mf.RealGrid(s, np.full(flags.shape, 3.0, self.real)) for _ in range(4)]      # every cell is written
        cg.InitPreconditionIncompCholesky(mf.FlagGrid(s, flags), *P, *[mf.RealGrid(s, a) for a in (A0, Ai, Aj, Ak)])
        return [p.numpy().copy() for p in P]

    def ic_apply(self, flags, src, P0, Pi, Pj, Pk):
        s = self._solver(flags)
        D = mf.RealGrid(s)
        cg.ApplyPreconditionIncompCholesky(D, mf.RealGrid(s, src), mf.FlagGrid(s, flags), *[mf.RealGrid(s, a) for a in (P0, Pi, Pj, Pk)])
        return D.numpy().copy()

    def cg_solve(self, flags, rhs, A0, Ai, Aj, Ak, pc=0, accuracy=1e-4, useL2=False, maxIter=1000):
        s = self._solver(flags)
        F = mf.FlagGrid(s, flags)
        A = [mf.RealGrid(s, a) for a in (A0, Ai, Aj, Ak)]
        x, b, r, se, t = mf.RealGrid(s), mf.RealGrid(s, rhs), mf.RealGrid(s), mf.RealGrid(s), mf.RealGrid(s)
        g = cg.GridCg(x, b, r, se, F, t, *A)
        g.setAccuracy(accuracy); g.setUseL2Norm(useL2)
        keep = None
        if pc == 1:
            keep = [mf.RealGrid(s) for _ in range(4)]
            g.setICPreconditioner(cg.GridCg.PC_mICP, *keep)
        elif pc == 2:
            keep = cg.GridMg(s)
            g.setMGPreconditioner(cg.GridCg.PC_MGP, keep)
        elif pc == 3:
            keep = [mf.RealGrid(s) for _ in range(4)]
            g.setICPreconditioner(cg.GridCg.PC_ICP, *keep)
        g.solve(maxIter)
        return x.numpy().copy(), g.getIterations(), g.getResNorm()

    def correct_velocity(self, flags, vel, pressure, phi=None, curv=None, gfClamp=1e-4, surfTens=0.0):
        s = self._solver(flags)
        V = mf.MACGrid(s, vel)
        mf.correctVelocity(V, mf.RealGrid(s, pressure), mf.FlagGrid(s, flags), phi=self._g(s, mf.RealGrid, phi), curv=self._g(s, mf.RealGrid, curv),
                           gfClamp=gfClamp, surfTens=surfTens)
        vel[...] = V.numpy()
        return vel

    def solve_pressure(self, flags, vel, pressure=None, phi=None, perCellCorr=None, fractions=None, obvel=None, curv=None, retRhs=False, solver_key=0, **kw):
        s = self._solver(flags, solver_key or None)
        F, V, P = mf.FlagGrid(s, flags), mf.MACGrid(s, vel), mf.RealGrid(s)
        RR = mf.RealGrid(s) if retRhs else None
        mf.solvePressure(vel=V, pressure=P, flags=F, phi=self._g(s, mf.RealGrid, phi), perCellCorr=self._g(s, mf.RealGrid, perCellCorr),
                         fractions=self._g(s, mf.MACGrid, fractions), obvel=self._g(s, mf.MACGrid, obvel), curv=self._g(s, mf.RealGrid, curv), retRhs=RR, **kw)
        info = mf.lastSolveInfo()
        vel[...] = V.numpy()
        out = (P.numpy().copy(), info["iterations"], info["resNorm"])
        return out + (RR.numpy().copy(),) if retRhs else out

    def cg_solve_diffusion(self, flags, grid, alpha=0.25, cgMaxIterFac=1.0, cgAccuracy=1e-4):
        s = self._solver(flags)
        G = (mf.RealGrid if grid.ndim == 3 else mf.MACGrid)(s, grid)
        mf.cgSolveDiffusion(mf.FlagGrid(s, flags), G, alpha=alpha, cgMaxIterFac=cgMaxIterFac, cgAccuracy=cgAccuracy)
        grid[...] = G.numpy()
        return grid

    # the steps either side of the projection: same signatures as oracle_api.Oracle, arrays updated in place
    def set_wall_bcs_obvel(self, flags, vel, obvel=None):
        s = self._solver(flags)
        V = mf.MACGrid(s, vel)
        mf.setWallBcs(mf.FlagGrid(s, flags), V, obvel=self._g(s, mf.MACGrid, obvel))
        vel[...] = V.numpy()
        return vel

    def add_gravity(self, flags, vel, gravity, exclude=None, scale=True, dt=1.0):
        s = self._solver(flags); s.timestep = dt
        V = mf.MACGrid(s, vel)
        mf.addGravity(mf.FlagGrid(s, flags), V, gravity, exclude=self._g(s, mf.RealGrid, exclude), scale=scale)
        vel[...] = V.numpy()
        return vel

    def add_buoyancy(self, flags, density, vel, gravity, coefficient=1.0, scale=True, dt=1.0):
        s = self._solver(flags); s.timestep = dt
        V = mf.MACGrid(s, vel)
        mf.addBuoyancy(mf.FlagGrid(s, flags), mf.RealGrid(s, density), V, gravity, coefficient=coefficient, scale=scale)
        vel[...] = V.numpy()
        return vel

    def advect_semi_lagrange(self, flags, vel, grid, order=1, strength=1.0, orderSpace=1, clampMode=2, orderTrace=1, dt=1.0, vec3=False):
        s = self._solver(flags); s.timestep = dt
        V = mf.MACGrid(s, vel)
        G = V if grid is vel else (mf.RealGrid if grid.ndim == 3 else (mf.VecGrid if vec3 else mf.MACGrid))(s, grid)      # self-advection: one grid, as in the scenes
        mf.advectSemiLagrange(mf.FlagGrid(s, flags), V, G, order=order, strength=strength, orderSpace=orderSpace, clampMode=clampMode, orderTrace=orderTrace)
        grid[...] = G.numpy()
        return grid

    # -- liquid neighbours --
    def extrapolate_mac_simple(self, flags, vel, distance=4, phiObs=None, intoObs=False):
        s = self._solver(flags)
        V = mf.MACGrid(s, vel)
        mf.extrapolateMACSimple(mf.FlagGrid(s, flags), V, distance=distance, phiObs=self._g(s, mf.LevelsetGrid, phiObs), intoObs=intoObs)
        vel[...] = V.numpy()
        return vel

    def extrapolate_mac_from_weight(self, vel, weight, distance=2):
        s = self._solver(vel[..., 0])
        V, W = mf.MACGrid(s, vel), mf.VecGrid(s, weight)
        mf.extrapolateMACFromWeight(V, W, distance=distance)
        vel[...] = V.numpy(); weight[...] = W.numpy()
        return vel

    def extrapolate_ls_simple(self, phi, distance=4, inside=False):
        s = self._solver(phi)
        P = mf.LevelsetGrid(s, phi)
        mf.extrapolateLsSimple(P, distance=distance, inside=inside)
        phi[...] = P.numpy()
        return phi

    def extrapolate_vec3_simple(self, vel, phi, distance=4, inside=False):
        s = self._solver(phi)
        V = mf.VecGrid(s, vel)
        mf.extrapolateVec3Simple(V, mf.LevelsetGrid(s, phi), distance=distance, inside=inside)
        vel[...] = V.numpy()
        return vel

    def update_from_levelset(self, flags, phi):
        s = self._solver(flags)
        F = mf.FlagGrid(s, flags)
        F.updateFromLevelset(mf.LevelsetGrid(s, phi))
        flags[...] = F.numpy()
        return flags

    def set_bound(self, grid, value, boundaryWidth=1):
        s = self._solver(grid[..., 0] if grid.ndim == 4 else grid)
        G = (mf.RealGrid if grid.ndim == 3 else mf.VecGrid)(s, grid)
        G.setBound(value, boundaryWidth)
        grid[...] = G.numpy()
        return grid

    def set_wall_bcs_frac(self, flags, vel, phiObs):
        s = self._solver(flags)
        V = mf.MACGrid(s, vel)
        mf.setWallBcs(mf.FlagGrid(s, flags), V, fractions=mf.MACGrid(s), phiObs=mf.RealGrid(s, phiObs))
        vel[...] = V.numpy()
        return vel

    def update_fractions(self, flags, phiObs, boundaryWidth=0, fracThreshold=0.01):
        s = self._solver(flags)
        Fr = mf.MACGrid(s, np.full(flags.shape + (3,), 7.0, self.real))         # every entry is written
        mf.updateFractions(mf.FlagGrid(s, flags), mf.RealGrid(s, phiObs), Fr, boundaryWidth=boundaryWidth, fracThreshold=fracThreshold)
        return Fr.numpy().copy()

    def set_obstacle_flags(self, flags, phiObs, fractions=None, phiOut=None, phiIn=None, boundaryWidth=1):
        s = self._solver(flags)
        F = mf.FlagGrid(s, flags)
        mf.setObstacleFlags(F, mf.RealGrid(s, phiObs), fractions=self._g(s, mf.MACGrid, fractions), phiOut=self._g(s, mf.RealGrid, phiOut),
                            phiIn=self._g(s, mf.RealGrid, phiIn), boundaryWidth=boundaryWidth)
        flags[...] = F.numpy()
        return flags

    def get_laplacian(self, grid):
        s = self._solver(grid)
        L = mf.RealGrid(s)
        mf.getLaplacian(L, mf.RealGrid(s, grid))
        return L.numpy().copy()

    def get_curvature(self, grid, h=1.0):
        s = self._solver(grid)
        Cv = mf.RealGrid(s)
        mf.getCurvature(Cv, mf.RealGrid(s, grid), h)
        return Cv.numpy().copy()

    # -- FLIP particle <-> grid plugins (second slice) --
    def _parts(self, s, pos, pflag, ptype=None, pvel=None):
        P = mf.BasicParticleSystem(s)
        T = P.create(mf.PdataInt) if ptype is not None else None
        V = P.create(mf.PdataVec3) if pvel is not None else None
        P.setParticles(pos, pflag)
        if T is not None:
            T.copyFromArray(ptype)
        if V is not None:
            V.copyFromArray(pvel)
        return P, T, V

    def mark_fluid_cells(self, flags, pos, pflag, phiObs=None, ptype=None, exclude=0):
        s = self._solver(flags)
        P, T, _ = self._parts(s, pos, pflag, ptype)
        F = mf.FlagGrid(s, flags)
        mf.markFluidCells(P, F, phiObs=self._g(s, mf.RealGrid, phiObs), ptype=T, exclude=exclude)
        flags[...] = F.numpy()
        return flags

    def grid_particle_index(self, shape, pos, pflag):
        s = self._solver(np.zeros(shape, np.int32))
        P, _, _ = self._parts(s, pos, pflag)
        I, Ix = mf.ParticleIndexSystem(s), mf.IntGrid(s, np.full(shape, 77, np.int32))       # every cell is written
        mf.gridParticleIndex(P, I, mf.FlagGrid(s), Ix)
        return Ix.numpy().copy(), I.numpy().copy()

    def union_particle_levelset(self, pos, index, indexSys, radiusFactor=1.0, ptype=None, exclude=0):
        s = self._solver(index)
        P, T, _ = self._parts(s, pos, np.zeros(len(pos), np.int32), ptype)
        I = mf.ParticleIndexSystem(s)
        I._a.set(indexSys); I._count = len(indexSys)
        phi = mf.LevelsetGrid(s, np.full(index.shape, 9.0, self.real))                          # every cell is written
        mf.unionParticleLevelset(P, I, mf.FlagGrid(s), mf.IntGrid(s, index), phi, radiusFactor=radiusFactor, ptype=T, exclude=exclude)
        return phi.numpy().copy()

    def map_parts_to_mac(self, shape, pos, pflag, pvel, want_weight=False, ptype=None, exclude=0):
        s = self._solver(np.zeros(shape, np.int32))
        P, T, V = self._parts(s, pos, pflag, ptype, pvel)
        junk = np.full(tuple(shape) + (3,), 5.0, self.real)                                        # every entry is written
        vel, velOld, w = mf.MACGrid(s, junk), mf.MACGrid(s, junk), (mf.VecGrid(s, junk) if want_weight else None)
        mf.mapPartsToMAC(mf.FlagGrid(s), vel, velOld, P, V, weight=w, ptype=T, exclude=exclude)
        return (vel.numpy().copy(), velOld.numpy().copy(), w.numpy().copy()) if want_weight else (vel.numpy().copy(), velOld.numpy().copy())

    def flip_velocity_update(self, vel, velOld, pos, pflag, pvel, flipRatio, ptype=None, exclude=0):
        s = self._solver(vel[..., 0])
        P, T, V = self._parts(s, pos, pflag, ptype, pvel)
        if flipRatio < 0:
            mf.mapMACToParts(mf.FlagGrid(s), mf.MACGrid(s, vel), P, V, ptype=T, exclude=exclude)
        else:
            mf.flipVelocityUpdate(mf.FlagGrid(s), mf.MACGrid(s, vel), mf.MACGrid(s, velOld), P, V, flipRatio, ptype=T, exclude=exclude)
        pvel[...] = V.numpy()
        return pvel

    def advect_in_grid(self, flags, vel, pos, pflag, dt, integrationMode=2, deleteInObstacle=True, stopInObstacle=True, skipNew=False, ptype=None, exclude=0):
        s = self._solver(flags); s.timestep = dt
        P, T, _ = self._parts(s, pos, pflag, ptype)
        P.advectInGrid(mf.FlagGrid(s, flags), mf.MACGrid(s, vel), integrationMode, deleteInObstacle=deleteInObstacle, stopInObstacle=stopInObstacle, skipNew=skipNew,
                       ptype=T, exclude=exclude)
        pos[...] = P.positions(); pflag[...] = P.flags()
        return pos, pflag

    def push_out_of_obs(self, shape, pos, pflag, phiObs, shift=0.0, thresh=0.0, ptype=None, exclude=0):
        s = self._solver(np.zeros(shape, np.int32))
        P, T, _ = self._parts(s, pos, pflag, ptype)
        mf.pushOutofObs(P, mf.FlagGrid(s), mf.RealGrid(s, phiObs), shift=shift, thresh=thresh, ptype=T, exclude=exclude)
        pos[...] = P.positions()
        return pos

    def project_out_of_bnd(self, shape, pos, pflag, bnd, plane="xXyYzZ", ptype=None, exclude=0):
        s = self._solver(np.zeros(shape, np.int32))
        P, T, _ = self._parts(s, pos, pflag, ptype)
        P.projectOutOfBnd(mf.FlagGrid(s), bnd, plane=plane, ptype=T, exclude=exclude)
        pos[...] = P.positions()
        return pos

    # -- the Lagrangian-particle helpers (plugin/ptsplugins.cpp, grid.cpp:866-890); arrays updated in place and returned like the oracle's --
    def _lag(self, pos, ptype=None, pvel=None):
        s = self._solver(np.zeros((4, 4, 4), np.int32))
        return (s,) + self._parts(s, pos, np.zeros(len(pos), np.int32), ptype, pvel)

    def add_force_pvel(self, pvel, a, dt, ptype=None, exclude=0):
        s, P, T, V = self._lag(np.zeros_like(pvel), ptype, pvel)
        mf.addForcePvel(V, a, dt, T, exclude)
        pvel[...] = V.numpy()
        return pvel

    def update_velocity_from_delta_pos(self, pos, pvel, x_prev, dt, ptype=None, exclude=0):
        s, P, T, V = self._lag(pos, ptype, pvel)
        X = P.create(mf.PdataVec3); X.copyFromArray(x_prev)
        mf.updateVelocityFromDeltaPos(P, V, X, dt, T, exclude)
        pvel[...] = V.numpy()
        return pvel

    def euler_step(self, pos, pvel, dt, ptype=None, exclude=0):
        s, P, T, V = self._lag(pos, ptype, pvel)
        s.timestep = dt
        before = P.create(mf.PdataVec3)
        P.getPosPdata(before)                       # ParticleSystem::getPosPdata on the way: the copy holds the old positions
        mf.eulerStep(P, V, T, exclude)
        assert np.array_equal(before.numpy(), pos)
        pos[...] = P.positions()
        return pos

    def set_part_type(self, flags, pos, ptype, mark, stype, cflag):
        s = self._solver(flags)
        P, T, _ = self._parts(s, pos, np.zeros(len(pos), np.int32), ptype)
        mf.setPartType(P, T, mark, stype, mf.FlagGrid(s, flags), cflag)
        ptype[...] = T.numpy()
        return ptype

    def mark_isolated_fluid_cell(self, flags, mark):
        s = self._solver(flags)
        F = mf.FlagGrid(s, flags)
        mf.markIsolatedFluidCell(F, mark)
        flags[...] = F.numpy()
        return flags

    def vec_max_abs(self, vel):
        s = self._solver(vel[..., 0])
        return mf.MACGrid(s, vel).getMaxAbs()

    def cg_solve_we(self, flags, ut, utm1, crankNic=False, cSqr=0.25, cgMaxIterFac=1.5, cgAccuracy=1e-5, dt=1.0):
        s = self._solver(flags); s.timestep = dt
        U, Um, O = mf.RealGrid(s, ut), mf.RealGrid(s, utm1), mf.RealGrid(s)
        mf.cgSolveWE(mf.FlagGrid(s, flags), U, Um, O, crankNic=crankNic, cSqr=cSqr, cgMaxIterFac=cgMaxIterFac, cgAccuracy=cgAccuracy)
        ut[...] = U.numpy(); utm1[...] = Um.numpy()
        return O.numpy().copy()

    def pd_fluid_guiding(self, flags, vel, velT, weight, blurRadius=5, theta=1.0, tau=1.0, sigma=1.0, epsRel=1e-3, epsAbs=1e-3, maxIters=200,
                         cgMaxIterFac=1.5, cgAccuracy=1e-3, preconditioner=1, zeroPressureFixing=False):
        s = self._solver(flags)
        V, P = mf.MACGrid(s, vel), mf.RealGrid(s)
        mf.PD_fluid_guiding(V, mf.MACGrid(s, velT), P, mf.FlagGrid(s, flags), mf.RealGrid(s, weight), blurRadius=blurRadius, theta=theta, tau=tau, sigma=sigma,
                            epsRel=epsRel, epsAbs=epsAbs, maxIters=maxIters, cgMaxIterFac=cgMaxIterFac, cgAccuracy=cgAccuracy, preconditioner=preconditioner,
                            zeroPressureFixing=zeroPressureFixing)
        mf.releaseMG(s)
        vel[...] = V.numpy()
        return P.numpy().copy(), mf.lastGuidingIterations()

    def release_solver(self, key):
        for k, s in list(self._solvers.items()):
            if k[3] == key:
                mf.releaseMG(s)

    # GridMg probes
    def mg_create(self, sx, sy, sz):
        self._mg_solver = mf.Solver(gridSize=(sx, sy, sz), dim=3 if sz > 1 else 2, prec=self.prec)
        self._mg = cg.GridMg(self._mg_solver)

    def mg_destroy(self):
        if self._mg is not None:
            self._mg.close(); self._mg = None

    def mg_set_a(self, A0, Ai, Aj, Ak):
        self._mg.setA(*[mf.RealGrid(self._mg_solver, a) for a in (A0, Ai, Aj, Ak)])

    def mg_num_levels(self):
        return self._mg.numLevels()

    def mg_level_size(self, l):
        return self._mg.levelInfo(l)[0]

    def mg_get(self, what, l):
        return self._mg.download(what, l)

    def mg_vcycle(self, rhs, coarsestAccuracy=1e-8, pre=1, post=1):
        s = self._mg_solver
        self._mg.setCoarsestLevelAccuracy(coarsestAccuracy); self._mg.setSmoothing(pre, post)
        self._mg.setRhs(mf.RealGrid(s, rhs))
        Z = mf.RealGrid(s)
        self._mg.doVCycle(Z)
        return Z.numpy().copy()

    def vic_poisson(self, flags, vort, vel, velIsMac=True, cgMaxIterFac=1.5, cgAccuracy=1e-3, scale=0.01, precondition=0):
        s = self._solver(flags)
        F, W = mf.FlagGrid(s, flags), mf.VecGrid(s, vort)
        V = (mf.MACGrid if velIsMac else mf.VecGrid)(s, vel)
        its = cg.vicPoisson(V, F, W, cgMaxIterFac=cgMaxIterFac, cgAccuracy=cgAccuracy, scale=scale, precondition=precondition)
        return V.numpy().copy(), its

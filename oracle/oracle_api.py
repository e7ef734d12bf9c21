"""TEST INFRASTRUCTURE ONLY: ctypes driver shared by the two CPU oracles.

* ``Oracle("port", prec)``      -> oracle/libmf_oracle_f{32,64}.so  (plain-C restatement, mf_oracle.c)
* ``Oracle("reference", prec)`` -> oracle/_ref/libmanta_ref_f{32,64}.so (the UNMODIFIED reference
  compiled from /root/reference by oracle/Makefile, C entry points in ref_harness.cpp)

Both export the same functions with prefix ``mfo_`` / ``ref_``.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / ``--impl reference`` legs may import this module; the product package
(mantaflow_b200/) never does.

Arrays follow the reference layout (grid.h:70): numpy shape [Z, Y, X] (C order) for Real/flag grids,
[Z, Y, X, 3] for MAC grids -- the same convention as the reference's numpy bridge
(plugin/numpyconvert.cpp:145-183).
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path(kind, prec):
    tag = "f32" if prec == 4 else "f64"
    if kind == "port":
        return os.path.join(_HERE, "libmf_oracle_%s.so" % tag)
    return os.path.join(_HERE, "_ref", "libmanta_ref_%s.so" % tag)


def available(kind, prec=4):
    return os.path.exists(lib_path(kind, prec))


class OracleError(RuntimeError):
    pass


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, kind="port", prec=4):
        assert kind in ("port", "reference") and prec in (4, 8)
        self.kind, self.prec = kind, prec
        self.real = np.float32 if prec == 4 else np.float64
        self.pfx = "mfo_" if kind == "port" else "ref_"
        path = lib_path(kind, prec)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle`)")
        self.lib = C.CDLL(path)
        assert self._f("real_size", C.c_int)() == prec
        self._f("set_debug_level")(C.c_int(0))   # the reference chats on stdout at level 1

    def _f(self, name, restype=C.c_int):
        f = getattr(self.lib, self.pfx + name)
        f.restype = restype
        return f

    def _chk(self, rc):
        if rc != 0:
            raise OracleError(self._f("last_error", C.c_char_p)().decode())

    def _r(self, a, shape=None):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=self.real)
        return a

    @staticmethod
    def dims(flags):
        sz, sy, sx = flags.shape
        return C.c_int(sx), C.c_int(sy), C.c_int(sz)

    # -- plugin/extforces.cpp:307 setWallBcs (no obvel / fractions) --
    def set_wall_bcs(self, flags, vel):
        assert vel.dtype == self.real and vel.flags.c_contiguous
        self._chk(self._f("set_wall_bcs")(*self.dims(flags), _p(flags), _p(vel)))
        return vel

    # -- the steps either side of the projection: arrays are updated in place and returned --
    def set_wall_bcs_obvel(self, flags, vel, obvel=None):
        assert vel.dtype == self.real and vel.flags.c_contiguous
        self._chk(self._f("set_wall_bcs_obvel")(*self.dims(flags), _p(flags), _p(vel), _p(self._r(obvel))))
        return vel

    def set_wall_bcs_frac(self, flags, vel, phiObs):
        """setWallBcs(flags, vel, fractions=..., phiObs=...): the second-order variant, plugin/extforces.cpp:220-303,:307-316"""
        assert vel.dtype == self.real and vel.flags.c_contiguous
        self._chk(self._f("set_wall_bcs_frac")(*self.dims(flags), _p(flags), _p(vel), _p(self._r(phiObs))))
        return vel

    def add_gravity(self, flags, vel, gravity, exclude=None, scale=True, dt=1.0):
        assert vel.dtype == self.real and vel.flags.c_contiguous
        self._chk(self._f("add_gravity")(*self.dims(flags), _p(flags), _p(vel), *[C.c_double(float(g)) for g in gravity],
                                         _p(self._r(exclude)), C.c_int(int(scale)), C.c_double(dt)))
        return vel

    def add_buoyancy(self, flags, density, vel, gravity, coefficient=1.0, scale=True, dt=1.0):
        assert vel.dtype == self.real and vel.flags.c_contiguous
        self._chk(self._f("add_buoyancy")(*self.dims(flags), _p(flags), _p(self._r(density)), _p(vel), *[C.c_double(float(g)) for g in gravity],
                                          C.c_double(coefficient), C.c_int(int(scale)), C.c_double(dt)))
        return vel

    def advect_semi_lagrange(self, flags, vel, grid, order=1, strength=1.0, orderSpace=1, clampMode=2, orderTrace=1, dt=1.0, vec3=False):
        """grid: (sz,sy,sx) Real grid or (sz,sy,sx,3) MAC grid (vec3=True: a cell-centred Grid<Vec3>), advected in place"""
        assert grid.dtype == self.real and grid.flags.c_contiguous
        kind = (2 if vec3 else 1) if grid.ndim == 4 else 0
        self._chk(self._f("advect_semi_lagrange")(*self.dims(flags), _p(flags), _p(self._r(vel)), _p(grid), C.c_int(kind), C.c_int(order),
                                                  C.c_double(strength), C.c_int(orderSpace), C.c_int(clampMode), C.c_int(orderTrace), C.c_double(dt)))
        return grid

    def cg_solve_we(self, flags, ut, utm1, crankNic=False, cSqr=0.25, cgMaxIterFac=1.5, cgAccuracy=1e-5, dt=1.0):
        """plugin/waves.cpp:86 cgSolveWE; ut / utm1 are advanced in place, returns `out`"""
        assert ut.dtype == self.real and utm1.dtype == self.real and ut.flags.c_contiguous and utm1.flags.c_contiguous
        out = np.zeros(flags.shape, self.real)
        self._chk(self._f("cg_solve_we")(*self.dims(flags), _p(flags), _p(ut), _p(utm1), _p(out), C.c_int(int(crankNic)), C.c_double(cSqr),
                                         C.c_double(cgMaxIterFac), C.c_double(cgAccuracy), C.c_double(dt)))
        return out

    def vic_integration(self, flags, tri, triVort, sigma, vel, velIsMac=True, cgMaxIterFac=1.5, cgAccuracy=1e-3, scale=0.01, precondition=0):
        """plugin/vortexplugins.cpp:195 VICintegration on a vortex sheet given as triangles (reference only: the mesh code is not restated);
        returns (vorticity grid, vel, the three GridCg iteration counts)"""
        assert self.kind == "reference"
        tri, triVort = self._r(tri), self._r(triVort)
        vel = np.array(vel, dtype=self.real, order="C")
        vort = np.zeros(flags.shape + (3,), self.real)
        its = (C.c_int * 3)()
        self._chk(self._f("vic_integration")(*self.dims(flags), _p(flags), C.c_int(len(triVort)), _p(tri), _p(triVort), C.c_double(sigma), _p(vel), C.c_int(int(velIsMac)),
                                             _p(vort), C.c_double(cgMaxIterFac), C.c_double(cgAccuracy), C.c_double(scale), C.c_int(precondition), its))
        return vort, vel, list(its)

    def vic_poisson(self, flags, vort, vel, velIsMac=True, cgMaxIterFac=1.5, cgAccuracy=1e-3, scale=0.01, precondition=0):
        """the grid half of VICintegration (vortexplugins.cpp:253-299): vorticity grid -> velocity; returns (vel, iteration counts).  Port only."""
        assert self.kind == "port"
        vort = self._r(vort)
        vel = np.array(vel, dtype=self.real, order="C")
        its = (C.c_int * 3)()
        self._chk(self._f("vic_poisson")(*self.dims(flags), _p(flags), _p(vort), _p(vel), C.c_int(int(velIsMac)), C.c_double(cgMaxIterFac), C.c_double(cgAccuracy),
                                         C.c_double(scale), C.c_int(precondition), its))
        return vel, list(its)

    def pd_fluid_guiding(self, flags, vel, velT, weight, blurRadius=5, theta=1.0, tau=1.0, sigma=1.0, epsRel=1e-3, epsAbs=1e-3, maxIters=200,
                         cgMaxIterFac=1.5, cgAccuracy=1e-3, preconditioner=1, zeroPressureFixing=False):
        """plugin/fluidguiding.cpp:294 PD_fluid_guiding; vel is replaced by the guided, divergence-free field; returns (pressure, iterations)"""
        assert vel.dtype == self.real and vel.flags.c_contiguous
        p = np.zeros(flags.shape, self.real)
        it = C.c_int(0)
        self._chk(self._f("pd_fluid_guiding")(*self.dims(flags), _p(flags), _p(vel), _p(self._r(velT)), _p(p), _p(self._r(weight)), C.c_int(blurRadius),
                                              C.c_double(theta), C.c_double(tau), C.c_double(sigma), C.c_double(epsRel), C.c_double(epsAbs), C.c_int(maxIters),
                                              C.c_double(cgMaxIterFac), C.c_double(cgAccuracy), C.c_int(preconditioner), C.c_int(int(zeroPressureFixing)),
                                              C.byref(it)))
        return p, it.value

    def compute_rhs(self, flags, vel, phi=None, perCellCorr=None, fractions=None, obvel=None, curv=None,
                    gfClamp=1e-4, surfTens=0.0, enforceCompatibility=False):
        vel, phi, perCellCorr, fractions, obvel, curv = map(self._r, (vel, phi, perCellCorr, fractions, obvel, curv))
        rhs = np.zeros(flags.shape, self.real)
        s, c = C.c_double(0), C.c_int(0)
        self._chk(self._f("compute_rhs")(*self.dims(flags), _p(flags), _p(vel), _p(rhs), _p(phi), _p(perCellCorr),
                                         _p(fractions), _p(obvel), _p(curv), C.c_double(gfClamp), C.c_double(surfTens),
                                         C.c_int(int(enforceCompatibility)), C.byref(s), C.byref(c)))
        return rhs, s.value, c.value

    def make_matrix(self, flags, fractions=None, phi=None, gfClamp=1e-4):
        fractions, phi = self._r(fractions), self._r(phi)
        A = [np.zeros(flags.shape, self.real) for _ in range(4)]
        self._chk(self._f("make_matrix")(*self.dims(flags), _p(flags), _p(fractions), _p(phi), C.c_double(gfClamp),
                                         *[_p(a) for a in A]))
        return A

    def choose_fix_cell(self, flags):
        assert self.kind == "port", "cell choice is only exposed by the restatement"
        return self._f("choose_fix_cell", C.c_longlong)(*self.dims(flags), _p(flags))

    def fix_pressure(self, flags, idx, value, rhs, A0, Ai, Aj, Ak):
        self._chk(self._f("fix_pressure")(*self.dims(flags), C.c_longlong(idx), C.c_double(value), _p(rhs), _p(A0), _p(Ai), _p(Aj), _p(Ak)))

    def apply_matrix(self, flags, src, A0, Ai, Aj, Ak):
        dst = np.zeros(flags.shape, self.real)
        self._chk(self._f("apply_matrix")(*self.dims(flags), _p(flags), _p(dst), _p(self._r(src)), _p(A0), _p(Ai), _p(Aj), _p(Ak)))
        return dst

    def mic_init(self, flags, A0, Ai, Aj, Ak):
        P = np.zeros(flags.shape, self.real)
        self._chk(self._f("mic_init")(*self.dims(flags), _p(flags), _p(P), _p(A0), _p(Ai), _p(Aj), _p(Ak)))
        return P

    def mic_apply(self, flags, src, P, A0, Ai, Aj, Ak, dst=None):
        dst = np.zeros(flags.shape, self.real) if dst is None else dst
        self._chk(self._f("mic_apply")(*self.dims(flags), _p(flags), _p(dst), _p(self._r(src)), _p(P), _p(A0), _p(Ai), _p(Aj), _p(Ak)))
        return dst

    # -- MIC(0) in block red-black ordering: the reformulated preconditioner (no reference counterpart; mf_oracle.c micrb_*), port only.
    #    cg_solve(pc=4) runs PCG with it.  tiles = (T0, T1, T2), 0 = the whole extent; (0, 0, 0) is the reference's ordering.
    def set_mic_tiles(self, tiles):
        assert self.kind == "port"
        self._chk(self._f("set_mic_tiles")(*[C.c_int(int(t)) for t in tiles]))

    def micrb_init(self, flags, A0, Ai, Aj, Ak, tiles):
        self.set_mic_tiles(tiles)
        P = np.zeros(flags.shape, self.real)
        self._chk(self._f("micrb_init")(*self.dims(flags), _p(flags), _p(P), _p(A0), _p(Ai), _p(Aj), _p(Ak)))
        return P

    def micrb_apply(self, flags, src, P, A0, Ai, Aj, Ak, tiles, dst=None):
        self.set_mic_tiles(tiles)
        dst = np.zeros(flags.shape, self.real) if dst is None else dst
        self._chk(self._f("micrb_apply")(*self.dims(flags), _p(flags), _p(dst), _p(self._r(src)), _p(P), _p(A0), _p(Ai), _p(Aj), _p(Ak)))
        return dst

    def ic_init(self, flags, A0, Ai, Aj, Ak):
        """InitPreconditionIncompCholesky conjugategrad.cpp:26-63 (PC_ICP); returns the factor grids (P0, Pi, Pj, Pk)"""
        P = [np.zeros(flags.shape, self.real) for _ in range(4)]
        self._chk(self._f("ic_init")(*self.dims(flags), _p(flags), *[_p(a) for a in P], _p(A0), _p(Ai), _p(Aj), _p(Ak)))
        return P

    def ic_apply(self, flags, src, P0, Pi, Pj, Pk, dst=None):
        """ApplyPreconditionIncompCholesky conjugategrad.cpp:109-132"""
        dst = np.zeros(flags.shape, self.real) if dst is None else dst
        self._chk(self._f("ic_apply")(*self.dims(flags), _p(flags), _p(dst), _p(self._r(src)), _p(P0), _p(Pi), _p(Pj), _p(Pk)))
        return dst

    def cg_solve(self, flags, rhs, A0, Ai, Aj, Ak, pc=0, accuracy=1e-4, useL2=False, maxIter=1000):
        x = np.zeros(flags.shape, self.real)
        it, rn = C.c_int(0), C.c_double(0)
        self._chk(self._f("cg_solve")(*self.dims(flags), _p(flags), _p(self._r(rhs)), _p(x), _p(A0), _p(Ai), _p(Aj), _p(Ak),
                                      C.c_int(pc), C.c_double(accuracy), C.c_int(int(useL2)), C.c_int(maxIter),
                                      C.byref(it), C.byref(rn)))
        return x, it.value, rn.value

    def correct_velocity(self, flags, vel, pressure, phi=None, curv=None, gfClamp=1e-4, surfTens=0.0):
        assert vel.dtype == self.real and vel.flags.c_contiguous
        self._chk(self._f("correct_velocity")(*self.dims(flags), _p(flags), _p(vel), _p(self._r(pressure)), _p(self._r(phi)),
                                              _p(self._r(curv)), C.c_double(gfClamp), C.c_double(surfTens)))
        return vel

    def solve_pressure(self, flags, vel, pressure=None, phi=None, perCellCorr=None, fractions=None, obvel=None, curv=None,
                       cgAccuracy=1e-3, gfClamp=1e-4, cgMaxIterFac=1.5, precondition=True, preconditioner=1,
                       enforceCompatibility=False, useL2Norm=False, zeroPressureFixing=False, surfTens=0.0,
                       retRhs=False, solver_key=0):
        """plugin/pressure.cpp:480-521; vel is updated in place; returns (pressure, iterations, resNorm[, rhs])"""
        assert vel.dtype == self.real and vel.flags.c_contiguous
        if pressure is None:
            pressure = np.zeros(flags.shape, self.real)
        rr = np.zeros(flags.shape, self.real) if retRhs else None
        it, rn = C.c_int(-1), C.c_double(-1)
        self._chk(self._f("solve_pressure")(C.c_longlong(solver_key), *self.dims(flags), _p(flags), _p(vel), _p(pressure),
                                            _p(self._r(phi)), _p(self._r(perCellCorr)), _p(self._r(fractions)), _p(self._r(obvel)),
                                            _p(self._r(curv)), _p(rr), C.c_double(cgAccuracy), C.c_double(gfClamp),
                                            C.c_double(cgMaxIterFac), C.c_int(int(precondition)), C.c_int(preconditioner),
                                            C.c_int(int(enforceCompatibility)), C.c_int(int(useL2Norm)),
                                            C.c_int(int(zeroPressureFixing)), C.c_double(surfTens), C.byref(it), C.byref(rn)))
        out = (pressure, it.value, rn.value)
        return out + (rr,) if retRhs else out

    def cg_solve_diffusion(self, flags, grid, alpha=0.25, cgMaxIterFac=1.0, cgAccuracy=1e-4):
        """conjugategrad.cpp:350-423; grid ([Z,Y,X] or [Z,Y,X,3]) is diffused in place"""
        assert grid.dtype == self.real and grid.flags.c_contiguous
        ncomp = 1 if grid.ndim == 3 else 3
        self._chk(self._f("cg_solve_diffusion")(*self.dims(flags), _p(flags), _p(grid), C.c_int(ncomp), C.c_double(alpha),
                                                C.c_double(cgMaxIterFac), C.c_double(cgAccuracy)))
        return grid

    # -- liquid neighbours: arrays are updated in place and returned --
    def extrapolate_mac_simple(self, flags, vel, distance=4, phiObs=None, intoObs=False):
        """fastmarch.cpp:337-375"""
        assert vel.dtype == self.real and vel.flags.c_contiguous
        self._chk(self._f("extrapolate_mac_simple")(*self.dims(flags), _p(flags), _p(vel), C.c_int(distance), _p(self._r(phiObs)), C.c_int(int(intoObs))))
        return vel

    def update_fractions(self, flags, phiObs, boundaryWidth=0, fracThreshold=0.01):
        """plugin/initplugins.cpp:437-440; returns the fractions grid [Z,Y,X,3]"""
        fr = np.zeros(flags.shape + (3,), self.real)
        self._chk(self._f("update_fractions")(*self.dims(flags), _p(flags), _p(self._r(phiObs)), _p(fr), C.c_int(boundaryWidth), C.c_double(fracThreshold)))
        return fr

    def set_obstacle_flags(self, flags, phiObs, fractions=None, phiOut=None, phiIn=None, boundaryWidth=1):
        """plugin/initplugins.cpp:473-475; flags are updated in place and returned"""
        assert flags.dtype == np.int32 and flags.flags.c_contiguous
        self._chk(self._f("set_obstacle_flags")(*self.dims(flags), _p(flags), _p(self._r(phiObs)), _p(self._r(fractions)), _p(self._r(phiOut)), _p(self._r(phiIn)),
                                                C.c_int(boundaryWidth)))
        return flags

    def extrapolate_mac_from_weight(self, vel, weight, distance=2):
        """fastmarch.cpp:410-432; vel and weight are both updated in place (the weight grid ends up holding the marks), vel is returned"""
        assert vel.dtype == self.real and weight.dtype == self.real and vel.flags.c_contiguous and weight.flags.c_contiguous
        self._chk(self._f("extrapolate_mac_from_weight")(*self.dims(vel[..., 0]), _p(vel), _p(weight), C.c_int(distance)))
        return vel

    def extrapolate_ls_simple(self, phi, distance=4, inside=False):
        """fastmarch.cpp:470-507"""
        assert phi.dtype == self.real and phi.flags.c_contiguous
        self._chk(self._f("extrapolate_ls_simple")(*self.dims(phi), _p(phi), C.c_int(distance), C.c_int(int(inside))))
        return phi

    def extrapolate_vec3_simple(self, vel, phi, distance=4, inside=False):
        """fastmarch.cpp:510-542"""
        assert vel.dtype == self.real and vel.flags.c_contiguous
        self._chk(self._f("extrapolate_vec3_simple")(*self.dims(phi), _p(vel), _p(self._r(phi)), C.c_int(distance), C.c_int(int(inside))))
        return vel

    def update_from_levelset(self, flags, phi):
        """FlagGrid::updateFromLevelset grid.cpp:844-854"""
        assert flags.dtype == np.int32 and flags.flags.c_contiguous
        self._chk(self._f("update_from_levelset")(*self.dims(flags), _p(flags), _p(self._r(phi))))
        return flags

    def set_bound(self, grid, value, boundaryWidth=1):
        """Grid<T>::setBound grid.cpp:591-593 ([Z,Y,X] or [Z,Y,X,3]; the same value in every component)"""
        assert grid.dtype == self.real and grid.flags.c_contiguous
        ncomp = 1 if grid.ndim == 3 else 3
        self._chk(self._f("set_bound")(*self.dims(grid[..., 0] if ncomp == 3 else grid), _p(grid), C.c_int(ncomp), C.c_double(value), C.c_int(boundaryWidth)))
        return grid

    def get_laplacian(self, grid):
        """plugin/flip.cpp:710-712; returns the Laplacian (outer layer 0)"""
        out = np.zeros(grid.shape, self.real)
        self._chk(self._f("get_laplacian")(*self.dims(grid), _p(out), _p(self._r(grid))))
        return out

    def get_curvature(self, grid, h=1.0):
        """plugin/flip.cpp:714-716; returns the curvature (outer layer 0)"""
        out = np.zeros(grid.shape, self.real)
        self._chk(self._f("get_curvature")(*self.dims(grid), _p(out), _p(self._r(grid)), C.c_double(h)))
        return out

    # -- FLIP particle <-> grid plugins (plugin/flip.cpp); particles are arrays: pos [N,3], pflag [N] int32, optional ptype [N] int32, pvel [N,3] --
    def _parts(self, pos, pflag, ptype=None):
        assert pos.dtype == self.real and pos.flags.c_contiguous and pflag.dtype == np.int32
        return C.c_longlong(len(pos)), _p(pos), _p(pflag), (None if ptype is None else _p(np.ascontiguousarray(ptype, np.int32)))

    def mark_fluid_cells(self, flags, pos, pflag, phiObs=None, ptype=None, exclude=0):
        """plugin/flip.cpp:158-177; flags updated in place and returned"""
        n, pp, pf, pt = self._parts(pos, pflag, ptype)
        self._chk(self._f("mark_fluid_cells")(*self.dims(flags), _p(flags), n, pp, pf, _p(self._r(phiObs)), pt, C.c_int(exclude)))
        return flags

    def grid_particle_index(self, shape, pos, pflag):
        """plugin/flip.cpp:260-306; returns (index [Z,Y,X] int32, indexSys [count] int32)"""
        index = np.zeros(shape, np.int32); isys = np.zeros(len(pos), np.int32); cnt = C.c_longlong(0)
        n, pp, pf, _ = self._parts(pos, pflag)
        self._chk(self._f("grid_particle_index")(*self.dims(index), n, pp, pf, _p(index), _p(isys), C.byref(cnt)))
        return index, isys[:cnt.value].copy()

    def union_particle_levelset(self, pos, index, indexSys, radiusFactor=1.0, ptype=None, exclude=0):
        """plugin/flip.cpp:340-350; returns phi"""
        phi = np.zeros(index.shape, self.real)
        pt = None if ptype is None else _p(np.ascontiguousarray(ptype, np.int32))
        self._chk(self._f("union_particle_levelset")(*self.dims(index), C.c_longlong(len(pos)), _p(pos), _p(index), _p(indexSys), C.c_longlong(len(indexSys)), _p(phi),
                                                     C.c_double(radiusFactor), pt, C.c_int(exclude)))
        return phi

    def map_parts_to_mac(self, shape, pos, pflag, pvel, want_weight=False, ptype=None, exclude=0):
        """plugin/flip.cpp:573-595; returns (vel, velOld[, weight])"""
        vel = np.zeros(tuple(shape) + (3,), self.real); velOld = np.zeros_like(vel); w = np.zeros_like(vel) if want_weight else None
        n, pp, pf, pt = self._parts(pos, pflag, ptype)
        self._chk(self._f("map_parts_to_mac")(*self.dims(vel[..., 0]), _p(vel), _p(velOld), n, pp, pf, _p(self._r(pvel)), _p(w), pt, C.c_int(exclude)))
        return (vel, velOld, w) if want_weight else (vel, velOld)

    def flip_velocity_update(self, vel, velOld, pos, pflag, pvel, flipRatio, ptype=None, exclude=0):
        """flipVelocityUpdate plugin/flip.cpp:669-677; flipRatio < 0: mapMACToParts :651-656 (pure PIC).  pvel updated in place and returned"""
        assert pvel.dtype == self.real and pvel.flags.c_contiguous
        n, pp, pf, pt = self._parts(pos, pflag, ptype)
        self._chk(self._f("flip_velocity_update")(*self.dims(vel[..., 0]), _p(self._r(vel)), _p(self._r(velOld)), n, pp, pf, _p(pvel), C.c_double(flipRatio), pt, C.c_int(exclude)))
        return pvel

    def advect_in_grid(self, flags, vel, pos, pflag, dt, integrationMode=2, deleteInObstacle=True, stopInObstacle=True, skipNew=False, ptype=None, exclude=0):
        """ParticleSystem::advectInGrid particle.h:512-536 (IntEuler 0, IntRK2 1, IntRK4 2); pos and pflag are updated in place and returned"""
        assert pos.dtype == self.real and pos.flags.c_contiguous and pflag.dtype == np.int32 and pflag.flags.c_contiguous
        pt = None if ptype is None else _p(np.ascontiguousarray(ptype, np.int32))
        self._chk(self._f("advect_in_grid")(*self.dims(flags), _p(flags), _p(self._r(vel)), C.c_longlong(len(pos)), _p(pos), _p(pflag), C.c_double(dt), C.c_int(integrationMode),
                                            C.c_int(int(deleteInObstacle)), C.c_int(int(stopInObstacle)), C.c_int(int(skipNew)), pt, C.c_int(exclude)))
        return pos, pflag

    def push_out_of_obs(self, shape, pos, pflag, phiObs, shift=0.0, thresh=0.0, ptype=None, exclude=0):
        """pushOutofObs plugin/flip.cpp:528-545; pos updated in place and returned"""
        n, pp, pf, pt = self._parts(pos, pflag, ptype)
        self._chk(self._f("push_out_of_obs")(*self.dims(np.empty(shape, np.int8)), n, pp, pf, _p(self._r(phiObs)), C.c_double(shift), C.c_double(thresh), pt, C.c_int(exclude)))
        return pos

    def project_out_of_bnd(self, shape, pos, pflag, bnd, plane="xXyYzZ", ptype=None, exclude=0):
        """ParticleSystem::projectOutOfBnd particle.h:565-590; pos updated in place and returned"""
        n, pp, pf, pt = self._parts(pos, pflag, ptype)
        axis = sum(1 << q for q, ch in enumerate("xXyYzZ") if ch in plane)
        self._chk(self._f("project_out_of_bnd")(*self.dims(np.empty(shape, np.int8)), n, pp, pf, C.c_double(bnd), C.c_int(axis), pt, C.c_int(exclude)))
        return pos

    # -- plugin/ptsplugins.cpp:17-70, grid.cpp:866-890: the Lagrangian-particle helpers of scenes/benchmark_dam.py; arrays updated in place and returned --
    def _pt(self, ptype):
        return None if ptype is None else _p(np.ascontiguousarray(ptype, np.int32))

    def add_force_pvel(self, pvel, a, dt, ptype=None, exclude=0):
        self._chk(self._f("add_force_pvel")(C.c_longlong(len(pvel)), _p(pvel), C.c_double(a[0]), C.c_double(a[1]), C.c_double(a[2]), C.c_double(dt), self._pt(ptype), C.c_int(exclude)))
        return pvel

    def update_velocity_from_delta_pos(self, pos, pvel, x_prev, dt, ptype=None, exclude=0):
        self._chk(self._f("update_velocity_from_delta_pos")(C.c_longlong(len(pos)), _p(pos), _p(pvel), _p(self._r(x_prev)), C.c_double(dt), self._pt(ptype), C.c_int(exclude)))
        return pvel

    def euler_step(self, pos, pvel, dt, ptype=None, exclude=0):
        self._chk(self._f("euler_step")(C.c_longlong(len(pos)), _p(pos), _p(self._r(pvel)), C.c_double(dt), self._pt(ptype), C.c_int(exclude)))
        return pos

    def set_part_type(self, flags, pos, ptype, mark, stype, cflag):
        assert ptype.dtype == np.int32 and ptype.flags.c_contiguous
        self._chk(self._f("set_part_type")(*self.dims(flags), C.c_longlong(len(pos)), _p(pos), _p(ptype), C.c_int(mark), C.c_int(stype), _p(flags), C.c_int(cflag)))
        return ptype

    def mark_isolated_fluid_cell(self, flags, mark):
        self._chk(self._f("mark_isolated_fluid_cell")(*self.dims(flags), _p(flags), C.c_int(mark)))
        return flags

    def vec_max_abs(self, vel):
        """Grid<Vec3>::getMaxAbs grid.cpp:330-332"""
        out = C.c_double(0)
        self._chk(self._f("vec_max_abs")(*self.dims(vel[..., 0]), _p(self._r(vel)), C.byref(out)))
        return out.value

    def grid_file(self, name, array, kind, load=False):
        """Grid<T>::save / load of the unmodified reference (grid.cpp:113-156; reference build only).
        kind: "real" | "mac" | "flags" | "levelset" | "vec3"; `array` is written to / filled from the file `name`."""
        assert self.kind == "reference", "file I/O is the reference's own code (fileio/iogrids.cpp); the restatement has none"
        k = {"real": 0, "mac": 1, "flags": 2, "levelset": 3, "vec3": 4}[kind]
        assert array.flags.c_contiguous and array.dtype == (np.int32 if kind == "flags" else self.real)
        self._chk(self._f("grid_file")(*self.dims(array[..., 0] if array.ndim == 4 else array), C.c_int(k), _p(array), name.encode(), C.c_int(int(load))))
        return array

    def release_solver(self, key):
        self._chk(self._f("release_solver")(C.c_longlong(key)))

    # -- GridMg probes --
    def mg_create(self, sx, sy, sz):
        self._chk(self._f("mg_create")(C.c_int(sx), C.c_int(sy), C.c_int(sz)))

    def mg_destroy(self):
        self._f("mg_destroy")()

    def mg_set_a(self, A0, Ai, Aj, Ak):
        self._chk(self._f("mg_set_a")(_p(A0), _p(Ai), _p(Aj), _p(Ak)))

    def mg_num_levels(self):
        return self._f("mg_num_levels")()

    def mg_level_size(self, l):
        o = (C.c_int * 3)()
        self._f("mg_level_size")(C.c_int(l), o)
        return tuple(o)

    def mg_get(self, what, l):
        sx, sy, sz = self.mg_level_size(l)
        n = sx * sy * sz
        if what == "type":
            out = np.zeros(n, np.int8)
        elif what == "a":
            out = np.zeros(n * self._f("mg_stencil_size")(C.c_int(l)), self.real)
        else:
            out = np.zeros(n, self.real)
        self._f("mg_get_" + what)(C.c_int(l), _p(out))
        return out

    def mg_vcycle(self, rhs, coarsestAccuracy=1e-8, pre=1, post=1):
        dst = np.zeros(rhs.shape, self.real)
        self._chk(self._f("mg_vcycle")(_p(self._r(rhs)), _p(dst), C.c_double(coarsestAccuracy), C.c_int(pre), C.c_int(post)))
        return dst

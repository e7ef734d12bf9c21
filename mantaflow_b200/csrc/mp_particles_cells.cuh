// FLIP particle <-> grid plugins on the device (second slice; scenes/benchmark_dam.py:100-125):
//   markFluidCells         plugin/flip.cpp:158-177  (knClearFluidFlags :137-141, knSetNbObstacle :142-157)
//   gridParticleIndex      plugin/flip.cpp:260-306
//   unionParticleLevelset  plugin/flip.cpp:340-350  (ComputeUnionLevelsetPindex :308-338) + phi.setBound(0.5, 0)
//   mapPartsToMAC          plugin/flip.cpp:573-595  (knMapLinearVec3ToMACGrid :562-569, setInterpolMAC util/interpol.h:159-203)
//   mapMACToParts          plugin/flip.cpp:651-656, flipVelocityUpdate :669-677 (interpolMAC util/interpol.h:127-157)
//   ParticleSystem::advectInGrid  particle.h:512-536 (Euler / RK2 / RK4 through the MAC grid, clamping / deletion in obstacles)
//   pushOutofObs           plugin/flip.cpp:528-545,  ParticleSystem::projectOutOfBnd  particle.h:565-590
//
// Particles are device arrays: pos [N][3] Real, pflag [N] int (BasicParticleData particle.h:182-191; active = !(flag & PDELETE)),
// per-particle data (ParticleDataImpl<Vec3|int>) [N][3] Real / [N] int.
//
// As in mp_liquid_cells.cuh, the per-cell / per-particle operations and the pass sequences are written against an executor `Exec`:
//     template <class F> int cells(const Dims& d, const F& f);            // f(d, i, j, k, idx) once per cell, any order
//     template <class F> int parts(IndexInt np, const F& f);              // f(idx) once per particle, any order
//     int zero(void* p, size_t bytes);
//     int exclusiveScan(int* data, IndexInt n, IndexInt* total);          // in place; *total = sum (on the host)
//     int sortPairs(int* keys, int* keysTmp, int* vals, int* valsOut, IndexInt n, int keyBits);   // STABLE sort by key; keys may be clobbered
// mp_particles.cu instantiates it with CUDA launches (+ cub for the scan and the radix sort); tests/emul/liquid_emul.cpp with host loops.
//
// Determinism.  The reference's scatter kernels are serial (knMapLinearVec3ToMACGrid is a `KERNEL(pts, single)`): a face sums its
// contributions in particle order.  Floating-point atomics would give a different, run-dependent sum.  Here the scatter is turned into
// a gather: particles are bucketed by cell (integer atomics for the counts, a scan, a stable radix sort -- all order independent), and a
// face walks the particles of the 3 x 3 (x 3) cells around it in ASCENDING PARTICLE ORDER (a k-way merge of the cells' ascending lists),
// adding the same products in the same order as the reference -> bit-identical grids, independent of the launch geometry.
#pragma once
#include "mp_liquid_cells.cuh"

namespace parts {

enum : int { PDELETE = 1 << 10 };      // particle.h:41

MP_HD int atomicIncr(int* p) {
#ifdef __CUDA_ARCH__
	return atomicAdd(p, 1);
#else
	return (*p)++;
#endif
}
MP_HD bool inBounds0(const Dims& d, int x, int y, int z) {      // GridBase::isInBounds(p, 0)
	return x >= 0 && y >= 0 && x < d.sx && y < d.sy && (d.is3D ? (z >= 0 && z < d.sz) : z == 0);
}
template <typename Real> struct PSet {       // the particles a plugin visits: active and not of an excluded type
	const Real* pos; const int* pflag; const int* ptype; int exclude;
	MP_HD bool skip(IndexInt idx) const { return (pflag[idx] & PDELETE) || (ptype && (ptype[idx] & exclude)); }
};

// ---------------------------------------------------------------- markFluidCells
struct ClearFluid {
	static const bool kSplit = false;         // knClearFluidFlags flip.cpp:137-141
	int* flags;
	MP_HD void operator()(const Dims&, int, int, int, IndexInt idx) const {
		const int f = flags[idx];
		if (f & TypeFluid) flags[idx] = (f | TypeEmpty) & ~TypeFluid;
	}
};
// flip.cpp:165-172.  Several particles of one cell race on the same word, but all of them store the same value (the cell's flag with
// Fluid set and Empty cleared), and a particle that reads the already converted flag stores nothing.
template <typename Real> struct MarkFluid {
	Dims d; int* flags; PSet<Real> ps;
	MP_HD void operator()(IndexInt idx) const {
		if (ps.skip(idx)) return;
		const int x = (int)ps.pos[3 * idx], y = (int)ps.pos[3 * idx + 1], z = (int)ps.pos[3 * idx + 2];
		if (!inBounds0(d, x, y, z)) return;
		const IndexInt p = (IndexInt)x + d.Y * y + d.Z * z;
		const int f = flags[p];
		if (f & TypeEmpty) flags[p] = (f | TypeFluid) & ~TypeEmpty;
	}
};
// knSetNbObstacle flip.cpp:142-157: reads `flags`, writes every cell of `out` (the caller swaps the two)
template <typename Real> struct SetNbObstacle {
	static const bool kSplit = false;
	const int* flags; int* out; const Real* phiObs;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt p) const {
		const int f = flags[p];
		int r = f;
		if (liquid::interiorCell(d, i, j, k) && !(phiObs[p] > 0.) && (f & TypeEmpty)) {
			const IndexInt X = 1, Y = d.Y, Z = d.Z;
			bool set = false;
			if ((flags[p - X] & TypeFluid) && (phiObs[p + X] <= 0.)) set = true;
			if ((flags[p + X] & TypeFluid) && (phiObs[p - X] <= 0.)) set = true;
			if ((flags[p - Y] & TypeFluid) && (phiObs[p + Y] <= 0.)) set = true;
			if ((flags[p + Y] & TypeFluid) && (phiObs[p - Y] <= 0.)) set = true;
			if (d.is3D) {
				if ((flags[p - Z] & TypeFluid) && (phiObs[p + Z] <= 0.)) set = true;
				if ((flags[p + Z] & TypeFluid) && (phiObs[p - Z] <= 0.)) set = true;
			}
			if (set) r = (f | TypeFluid) & ~TypeEmpty;
		}
		out[p] = r;
	}
};
// `tmp`: scratch FlagGrid (only used with phiObs); *swapped is set when the result is in tmp (the caller exchanges the buffers)
template <typename Real, typename Exec>
int markFluidCells(Exec& ex, const Dims& d, int* flags, IndexInt np, const PSet<Real>& ps, const Real* phiObs, int* tmp, bool* swapped) {
	*swapped = false;
	ClearFluid cf = { flags }; MP_TRY(ex.cells(d, cf));
	if (np > 0) { MarkFluid<Real> mf = { d, flags, ps }; MP_TRY(ex.parts(np, mf)); }
	if (phiObs) { SetNbObstacle<Real> nb = { flags, tmp, phiObs }; MP_TRY(ex.cells(d, nb)); *swapped = true; }
	return MP_OK;
}

// ---------------------------------------------------------------- bucketing particles by cell
// key[idx] = cell of the particle (or the sentinel d.n for particles that are left out), val[idx] = idx, count[cell]++.
// clampToGrid == false: gridParticleIndex (flip.cpp:275-283: deleted and out-of-bounds particles are left out, ptype is not looked at);
// clampToGrid == true : the private buckets of mapPartsToMAC (skipped particles are left out, positions outside are clamped, as
// BUILD_INDEX clamps them).
template <typename Real> struct KeyCount {
	Dims d; PSet<Real> ps; int* key; int* val; int* count; bool clampToGrid;
	MP_HD void operator()(IndexInt idx) const {
		val[idx] = (int)idx;
		int k = (int)d.n;
		if (!clampToGrid) {
			if (!(ps.pflag[idx] & PDELETE)) {
				const int x = (int)ps.pos[3 * idx], y = (int)ps.pos[3 * idx + 1], z = (int)ps.pos[3 * idx + 2];
				if (inBounds0(d, x, y, z)) k = (int)((IndexInt)x + d.Y * y + d.Z * z);
			}
		} else if (!ps.skip(idx)) {
			k = (int)cellClamped(d, ps.pos + 3 * idx);
		}
		key[idx] = k;
		if (k < (int)d.n) atomicIncr(count + k);
	}
	// the bucket of a particle for the gather of mapPartsToMAC: int(pos) per axis, clamped into the grid.  Every face the particle
	// contributes to lies in the 3 x 3 (x 3) cells around this one (see MapPartsGather).
	static MP_HD int clampAxis(Real p, int s) {
		if (!(p > (Real)0)) return 0;                  // negative or NaN
		if (!(p < (Real)s)) return s - 1;
		return (int)p;
	}
	static MP_HD IndexInt cellClamped(const Dims& d, const Real* pos) {
		const int x = clampAxis(pos[0], d.sx), y = clampAxis(pos[1], d.sy), z = d.is3D ? clampAxis(pos[2], d.sz) : 0;
		return (IndexInt)x + d.Y * y + d.Z * z;
	}
};
MP_HD int bitsFor(IndexInt maxKey) { int b = 1; while (((IndexInt)1 << b) <= maxKey) b++; return b; }

// count -> index (first slot of each cell), sorted particle ids -> sorted.  key/keyTmp/val: scratch int[np]; sorted: int[np].
template <typename Real, typename Exec>
int bucketParticles(Exec& ex, const Dims& d, IndexInt np, const PSet<Real>& ps, bool clampToGrid, int* index, int* key, int* keyTmp, int* val, int* sorted, IndexInt* count) {
	MP_TRY(ex.zero(index, sizeof(int) * (size_t)d.n));
	*count = 0;
	if (np <= 0) return MP_OK;
	KeyCount<Real> kc = { d, ps, key, val, index, clampToGrid };
	MP_TRY(ex.parts(np, kc));
	MP_TRY(ex.exclusiveScan(index, d.n, count));
	MP_TRY(ex.sortPairs(key, keyTmp, val, sorted, np, bitsFor(d.n)));
	return MP_OK;
}

// ---------------------------------------------------------------- unionParticleLevelset
// ComputeUnionLevelsetPindex flip.cpp:308-338 (a gather already in the reference: min over the particles of the cells within `r`),
// fused with the setBound(0.5, 0) that follows it (flip.cpp:349).
template <typename Real> struct UnionLevelset {
	static const bool kSplit = false;
	const Real* pos; const int* index; const int* indexSys; IndexInt count; Real* phi; Real radius; int r; const int* ptype; int exclude;
	const Real* posS;      // positions in index order (posS[3 p + c] = pos[3 indexSys[p] + c]), or NULL: read through indexSys
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		if (i <= 0 || i >= d.sx - 1 || j <= 0 || j >= d.sy - 1 || (d.is3D && (k <= 0 || k >= d.sz - 1))) { phi[idx] = (Real)0.5; return; }
		const Real gx = (Real)i + (Real)0.5, gy = (Real)j + (Real)0.5, gz = (Real)k + (Real)0.5;      // gridPos, flip.cpp:317
		const Real eps2 = sizeof(Real) == 8 ? (Real)(1e-10 * 1e-10) : (Real)(1e-6f * 1e-6f);            // VECTOR_EPSILON^2 vectorbase.h:399-411
		Real phiv = radius;
		const int rZ = d.is3D ? r : 0;
		for (int zj = k - rZ; zj <= k + rZ; zj++) for (int yj = j - r; yj <= j + r; yj++) {
			if (yj < 0 || yj >= d.sy || (d.is3D ? (zj < 0 || zj >= d.sz) : zj != 0)) continue;
			// the cells of one row are neighbours in the index: one particle range per row
			const int x0 = i - r < 0 ? 0 : i - r, x1 = i + r >= d.sx ? d.sx - 1 : i + r;
			const IndexInt c0 = (IndexInt)x0 + d.Y * yj + d.Z * zj, c1 = (IndexInt)x1 + d.Y * yj + d.Z * zj;
			const IndexInt pStart = index[c0], pEnd = (c1 + 1 < d.n) ? (IndexInt)index[c1 + 1] : count;
			for (IndexInt p = pStart; p < pEnd; p++) {
				if (ptype && (ptype[indexSys[p]] & exclude)) continue;
				const Real* q = posS ? posS + 3 * p : pos + 3 * (IndexInt)indexSys[p];
				const Real dx = gx - q[0], dy = gy - q[1], dz = gz - q[2];
				const Real l = dx * dx + dy * dy + dz * dz;
				Real nrm;                                                                        // norm() vectorbase.h:399-404
				if (l <= eps2) nrm = (Real)0;
				else if (fabs((double)l - 1.) < (double)eps2) nrm = (Real)1;
				else nrm = (Real)sqrt(l);
				const Real v = (Real)fabs((double)nrm) - radius;
				phiv = phiv < v ? phiv : v;
			}
		}
		phi[idx] = phiv;
	}
};
template <typename Real> struct GatherIndexed {
	const int* indexSys; const Real* pos; Real* posS;
	MP_HD void operator()(IndexInt s) const { const IndexInt p = indexSys[s]; for (int c = 0; c < 3; c++) posS[3 * s + c] = pos[3 * p + c]; }
};
// posScratch: Real[3 count] or NULL.  With it the positions are first copied into index order (one random read per particle), so that the ~27
// visits of every particle by the cells around it read neighbouring memory instead of following indexSys each time.
template <typename Real, typename Exec>
int unionParticleLevelset(Exec& ex, const Dims& d, const Real* pos, const int* index, const int* indexSys, IndexInt count, Real* phi, double radiusFactor_,
                          const int* ptype, int exclude, Real* posScratch = nullptr) {
	const Real radiusFactor = (Real)radiusFactor_;
	const Real radius = (Real)(0.5 * (double)(Real)((d.is3D ? sqrt(3.) : sqrt(2.)) * ((double)radiusFactor + .01)));      // flip.cpp:186-188, :343
	if (posScratch && count > 0) { GatherIndexed<Real> gi = { indexSys, pos, posScratch }; MP_TRY(ex.parts(count, gi)); }
	UnionLevelset<Real> op = { pos, index, indexSys, count, phi, radius, (int)radius + 1, ptype, exclude, count > 0 ? posScratch : nullptr };
	return ex.cells(d, op);
}

// ---------------------------------------------------------------- MAC interpolation weights
// BUILD_INDEX / BUILD_INDEX_SHIFT util/interpol.h:50-66,:112-125: cell-centred base (xi, yi, zi) with weights s, t, f and the face
// base (sxi, syi, szi) with weights ss, st, sf; the `1 - w` are double subtractions narrowed on assignment.
template <typename Real> struct MacWeights {
	int xi, yi, zi, sxi, syi, szi;
	Real s[2], t[2], f[2], ss[2], st[2], sf[2];
	static MP_HD void axis(Real p, int size, bool clampHigh, int& i, Real (&w)[2]) {
		i = (int)p;
		w[1] = p - (Real)i; w[0] = (Real)(1. - w[1]);
		if (p < 0.) { i = 0; w[0] = (Real)1; w[1] = (Real)0; }
		if (clampHigh && i >= size - 1) { i = size - 2; w[0] = (Real)0; w[1] = (Real)1; }
	}
	MP_HD MacWeights(const Dims& d, const Real* pos) {
		const Real px = pos[0] - (Real)0.5f, py = pos[1] - (Real)0.5f, pz = pos[2] - (Real)0.5f;
		axis(px, d.sx, true, xi, s); axis(py, d.sy, true, yi, t); axis(pz, d.sz, d.is3D, zi, f);
		axis(pos[0], d.sx, true, sxi, ss); axis(pos[1], d.sy, true, syi, st); axis(pos[2], d.sz, d.is3D, szi, sf);
	}
};
// interpolMAC util/interpol.h:127-157
template <typename Real> MP_HD void interpolMAC(const Dims& d, const Real* data, const MacWeights<Real>& m, Real (&out)[3]) {
	const IndexInt X = 1, Y = d.Y, Z = d.Z;
	#define MP_RF(o, c) ref[3 * (o) + (c)]
	{ const Real* ref = data + 3 * (((IndexInt)m.zi * d.sy + m.yi) * d.sx + m.sxi);
	  out[0] = m.f[0] * ((MP_RF(0, 0) * m.t[0] + MP_RF(Y, 0) * m.t[1]) * m.ss[0] + (MP_RF(X, 0) * m.t[0] + MP_RF(X + Y, 0) * m.t[1]) * m.ss[1]) +
	           m.f[1] * ((MP_RF(Z, 0) * m.t[0] + MP_RF(Z + Y, 0) * m.t[1]) * m.ss[0] + (MP_RF(X + Z, 0) * m.t[0] + MP_RF(X + Y + Z, 0) * m.t[1]) * m.ss[1]); }
	{ const Real* ref = data + 3 * (((IndexInt)m.zi * d.sy + m.syi) * d.sx + m.xi);
	  out[1] = m.f[0] * ((MP_RF(0, 1) * m.st[0] + MP_RF(Y, 1) * m.st[1]) * m.s[0] + (MP_RF(X, 1) * m.st[0] + MP_RF(X + Y, 1) * m.st[1]) * m.s[1]) +
	           m.f[1] * ((MP_RF(Z, 1) * m.st[0] + MP_RF(Z + Y, 1) * m.st[1]) * m.s[0] + (MP_RF(X + Z, 1) * m.st[0] + MP_RF(X + Y + Z, 1) * m.st[1]) * m.s[1]); }
	{ const Real* ref = data + 3 * (((IndexInt)m.szi * d.sy + m.yi) * d.sx + m.xi);
	  out[2] = m.sf[0] * ((MP_RF(0, 2) * m.t[0] + MP_RF(Y, 2) * m.t[1]) * m.s[0] + (MP_RF(X, 2) * m.t[0] + MP_RF(X + Y, 2) * m.t[1]) * m.s[1]) +
	           m.sf[1] * ((MP_RF(Z, 2) * m.t[0] + MP_RF(Z + Y, 2) * m.t[1]) * m.s[0] + (MP_RF(X + Z, 2) * m.t[0] + MP_RF(X + Y + Z, 2) * m.t[1]) * m.s[1]); }
	#undef MP_RF
}

// ---------------------------------------------------------------- mapMACToParts / flipVelocityUpdate
// knMapLinearMACGridToVec3_PIC flip.cpp:643-649 (flipRatio < 0) and knMapLinearMACGridToVec3_FLIP :659-667: a gather per particle.
template <typename Real> struct FlipVelocityUpdate {
	Dims d; const Real* vel; const Real* velOld; PSet<Real> ps; Real* pvel; Real flipRatio; bool pic;
	MP_HD void operator()(IndexInt idx) const {
		if (ps.skip(idx)) return;
		const MacWeights<Real> m(d, ps.pos + 3 * idx);
		Real v[3];
		interpolMAC(d, vel, m, v);
		if (pic) { for (int c = 0; c < 3; c++) pvel[3 * idx + c] = v[c]; return; }
		Real o[3];
		interpolMAC(d, velOld, m, o);
		for (int c = 0; c < 3; c++) {
			const Real delta = v[c] - o[c];
			// flipRatio * (v + delta) + (1.0 - flipRatio) * vNew: the second product is a double expression narrowed by the Vec3 operator* (flip.cpp:665)
			pvel[3 * idx + c] = (Real)((double)(flipRatio * (pvel[3 * idx + c] + delta)) + (double)(Real)((1.0 - (double)flipRatio) * (double)v[c]));
		}
	}
};

// ---------------------------------------------------------------- ParticleSystem::advectInGrid
// particle.h:512-536: GridAdvectKernel :446-467 run 1 / 2 / 4 times by integratePointSet (util/integrator.h:26-68), then KnClampPositions
// :494-509 (with bisectBacktracePos :480-490) or KnDeleteInObstacle :471-477.  The reference runs each stage over all particles; no stage
// couples two particles, so one thread takes a particle through all stages (positions and flags are read and written once).
// Mixed precision as in the reference: `0.5 * u` is a double product narrowed by the Vec3 constructor, `oldp * (1. - s)` likewise.
enum : int { PNEW = 1 << 0 };              // particle.h:36
enum : int { IntEuler = 0, IntRK2 = 1, IntRK4 = 2 };      // util/integrator.h:23
template <typename Real> struct AdvectInGrid {
	Dims d; const int* flags; const Real* vel; Real* pos; int* pflag; const int* ptype; int exclude;
	Real dt; int mode; bool deleteInObstacle, stopInObstacle, skipNew;
	MP_HD bool inBounds(const Real* p, int b) const {                                          // GridBase::isInBounds(Vec3, bnd) grid.h:60,:407-415
		const int x = (int)p[0], y = (int)p[1], z = (int)p[2];
		return x >= b && y >= b && x < d.sx - b && y < d.sy - b && (d.is3D ? (z >= b && z < d.sz - b) : z == 0);
	}
	MP_HD bool obstacleAt(const Real* p) const {                                               // FlagGrid::isObstacle(const Vec3&) grid.h:313
		const IndexInt q = (IndexInt)(int)p[0] + d.Y * (int)p[1] + d.Z * (int)p[2];
		if (q < 0 || q >= d.n) return false;                                                    // the reference reads out of bounds here (bisection from a position outside)
		return (flags[q] & TypeObstacle) != 0;
	}
	MP_HD void velKernel(const Real* x, int& fl, int pt, Real (&u)[3]) const {
		if ((fl & PDELETE) || (pt & exclude) || (skipNew && (fl & PNEW))) { u[0] = u[1] = u[2] = 0; return; }
		if (deleteInObstacle || stopInObstacle) {
			if (!inBounds(x, 1) || obstacleAt(x)) {
				if (stopInObstacle) u[0] = u[1] = u[2] = 0;          // otherwise u keeps the value of the previous stage (the result vector persists)
				if (deleteInObstacle) fl |= PDELETE;
				return;
			}
		}
		const MacWeights<Real> m(d, x);
		Real v[3];
		interpolMAC(d, vel, m, v);
		for (int c = 0; c < 3; c++) u[c] = v[c] * dt;
	}
	MP_HD void operator()(IndexInt idx) const {
		Real x[3] = { pos[3 * idx], pos[3 * idx + 1], pos[3 * idx + 2] };
		const Real x0[3] = { x[0], x[1], x[2] };
		int fl = pflag[idx];
		const int fl0 = fl, pt = ptype ? ptype[idx] : 0;
		Real u[3] = { 0, 0, 0 }, uTotal[3];
		velKernel(x, fl, pt, u);
		if (mode == IntEuler) { for (int c = 0; c < 3; c++) x[c] += u[c]; }
		else if (mode == IntRK2) {
			for (int c = 0; c < 3; c++) x[c] = x0[c] + (Real)(0.5 * (double)u[c]);
			velKernel(x, fl, pt, u);
			for (int c = 0; c < 3; c++) x[c] = x0[c] + u[c];
		} else {
			for (int c = 0; c < 3; c++) { uTotal[c] = u[c]; x[c] = x0[c] + (Real)(0.5 * (double)u[c]); }
			velKernel(x, fl, pt, u);
			for (int c = 0; c < 3; c++) { x[c] = x0[c] + (Real)(0.5 * (double)u[c]); uTotal[c] += (Real)(2 * u[c]); }
			velKernel(x, fl, pt, u);
			for (int c = 0; c < 3; c++) { x[c] = x0[c] + u[c]; uTotal[c] += (Real)(2 * u[c]); }
			velKernel(x, fl, pt, u);
			for (int c = 0; c < 3; c++) x[c] = x0[c] + (Real)(1. / 6.) * (uTotal[c] + u[c]);
		}
		if (!deleteInObstacle) {                                     // KnClampPositions
			if (!(fl & PDELETE)) {
				if (pt & exclude) { for (int c = 0; c < 3; c++) x[c] = x0[c]; }
				else {
					if (!inBounds(x, 0)) {
						const Real hi[3] = { (Real)d.sx - (Real)1, (Real)d.sy - (Real)1, (Real)d.sz - (Real)1 };
						for (int c = 0; c < 3; c++) { if (x[c] < (Real)0) x[c] = 0; else if (x[c] > hi[c]) x[c] = hi[c]; }
					}
					if (stopInObstacle && obstacleAt(x)) {              // bisectBacktracePos
						Real s = 0.;
						for (int i = 1; i < 5; ++i) {
							const Real ds = (Real)(1. / (double)(Real)(1 << i));
							const Real sd = s + ds;
							Real q[3];
							for (int c = 0; c < 3; c++) q[c] = (Real)((double)x0[c] * (1. - (double)sd)) + x[c] * sd;
							if (!obstacleAt(q)) s += ds;
						}
						for (int c = 0; c < 3; c++) x[c] = (Real)((double)x0[c] * (1. - (double)s)) + x[c] * s;
					}
				}
			}
		} else if (!(fl & PDELETE)) {                                 // KnDeleteInObstacle
			if (!inBounds(x, 1) || obstacleAt(x)) fl |= PDELETE;
		}
		for (int c = 0; c < 3; c++) pos[3 * idx + c] = x[c];
		if (fl != fl0) pflag[idx] = fl;
	}
};

// ---------------------------------------------------------------- pushOutofObs / projectOutOfBnd
// knPushOutofObs plugin/flip.cpp:528-540: particles whose interpolated obstacle distance is below `thresh` move along the normalised
// central-difference gradient of the cell they are in (getGradient grid.h:520-537, interpol util/interpol.h:68-78, normalize vectorbase.h:415-429).
template <typename Real> struct PushOutOfObs {
	Dims d; Real* pos; PSet<Real> ps; const Real* phiObs; Real shift, thresh;
	MP_HD void operator()(IndexInt idx) const {
		if (ps.skip(idx)) return;
		Real* x = pos + 3 * idx;
		int i = (int)x[0], j = (int)x[1], k = (int)x[2];
		if (!(i >= 0 && j >= 0 && k >= 0 && i < d.sx && j < d.sy && k < d.sz)) return;      // GridBase::isInBounds(Vec3i) grid.h:403-405
		const MacWeights<Real> m(d, x);
		const IndexInt X = 1, Y = d.Y, Z = d.Z;
		const Real* p = phiObs + ((IndexInt)m.xi + Y * m.yi + Z * m.zi);
		const Real v = ((p[0] * m.t[0] + p[Y] * m.t[1]) * m.s[0] + (p[X] * m.t[0] + p[X + Y] * m.t[1]) * m.s[1]) * m.f[0]
		             + ((p[Z] * m.t[0] + p[Y + Z] * m.t[1]) * m.s[0] + (p[X + Z] * m.t[0] + p[X + Y + Z] * m.t[1]) * m.s[1]) * m.f[1];
		if (!(v < thresh)) return;
		if (i > d.sx - 2) i = d.sx - 2;
		if (j > d.sy - 2) j = d.sy - 2;
		if (i < 1) i = 1;
		if (j < 1) j = 1;
		Real g[3];
		g[0] = phiObs[(i + 1) + Y * j + Z * k] - phiObs[(i - 1) + Y * j + Z * k];
		g[1] = phiObs[i + Y * (j + 1) + Z * k] - phiObs[i + Y * (j - 1) + Z * k];
		g[2] = 0;
		if (d.is3D) {
			if (k > d.sz - 2) k = d.sz - 2;
			if (k < 1) k = 1;
			g[2] = phiObs[i + Y * j + Z * (k + 1)] - phiObs[i + Y * j + Z * (k - 1)];
		}
		// normalize(): the comparison against 1. and the reciprocal are double expressions
		const Real l = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
		const Real eps = sizeof(Real) == 8 ? (Real)1e-10 : (Real)1e-6f;
		const Real eps2 = sizeof(Real) == 8 ? (Real)(1e-10 * 1e-10) : (Real)(1e-6f * 1e-6f);
		Real norm;
		if (fabs((double)l - 1.) < (double)eps2) norm = (Real)1.;
		else if (l > eps2) {
			norm = (Real)sqrt(l);
			const Real r = (Real)(1. / (double)norm);
			g[0] *= r; g[1] *= r; g[2] *= r;
		} else return;                                                                     // norm 0 < VECTOR_EPSILON
		if (norm < eps) return;
		const Real a = thresh - v + shift;
		for (int c = 0; c < 3; c++) x[c] = x[c] + g[c] * a;
	}
};
// KnProjectOutOfBnd particle.h:565-576; axis: bit q <-> the q-th letter of "xXyYzZ"
template <typename Real> struct ProjectOutOfBnd {
	Dims d; Real* pos; PSet<Real> ps; Real bnd; int axis;
	MP_HD void operator()(IndexInt idx) const {
		if (ps.skip(idx)) return;
		Real* x = pos + 3 * idx;
		if (axis & 1) x[0] = x[0] < bnd ? bnd : x[0];                                        // std::max(pos.x, bnd)
		if (axis & 2) { const Real hi = (Real)d.sx - bnd; x[0] = hi < x[0] ? hi : x[0]; }    // std::min(pos.x, size - bnd)
		if (axis & 4) x[1] = x[1] < bnd ? bnd : x[1];
		if (axis & 8) { const Real hi = (Real)d.sy - bnd; x[1] = hi < x[1] ? hi : x[1]; }
		if (d.is3D) {
			if (axis & 16) x[2] = x[2] < bnd ? bnd : x[2];
			if (axis & 32) { const Real hi = (Real)d.sz - bnd; x[2] = hi < x[2] ? hi : x[2]; }
		}
	}
};

// ---------------------------------------------------------------- the Lagrangian-particle helpers of scenes/benchmark_dam.py:118-134
// plugin/ptsplugins.cpp:17-70 (KnAddForcePvel, KnUpdateVelocityFromDeltaPos, KnStepEuler, KnSetPartType) and markIsolatedFluidCell
// grid.cpp:866-890.  None of them looks at PDELETE; particles of an excluded type are skipped.
template <typename Real> struct AddForcePvel {
	Real* v; Real da[3]; const int* ptype; int exclude;
	MP_HD void operator()(IndexInt idx) const { if (ptype && (ptype[idx] & exclude)) return; for (int c = 0; c < 3; c++) v[3 * idx + c] += da[c]; }
};
template <typename Real> struct UpdateVelocityFromDeltaPos {
	const Real* pos; Real* v; const Real* xPrev; Real overDt; const int* ptype; int exclude;
	MP_HD void operator()(IndexInt idx) const { if (ptype && (ptype[idx] & exclude)) return; for (int c = 0; c < 3; c++) v[3 * idx + c] = (pos[3 * idx + c] - xPrev[3 * idx + c]) * overDt; }
};
template <typename Real> struct StepEuler {
	Real* pos; const Real* v; Real dt; const int* ptype; int exclude;
	MP_HD void operator()(IndexInt idx) const { if (ptype && (ptype[idx] & exclude)) return; for (int c = 0; c < 3; c++) pos[3 * idx + c] += v[3 * idx + c] * dt; }
};
template <typename Real> struct SetPartType {
	Dims d; const Real* pos; int* ptype; int mark, stype; const int* flags; int cflag;
	MP_HD void operator()(IndexInt idx) const {
		const int x = (int)pos[3 * idx], y = (int)pos[3 * idx + 1], z = (int)pos[3 * idx + 2];
		if (inBounds0(d, x, y, z) && (flags[(IndexInt)x + d.Y * y + d.Z * z] & cflag) && (ptype[idx] & stype)) ptype[idx] = mark;
	}
};
// A fluid cell without fluid neighbours takes `mark`.  In place: a cell that is rewritten has no fluid neighbour, so no cell that reads it can
// change its own decision (the reference's parallel kernel relies on the same argument).  Fluid cells are interior cells.
struct MarkIsolatedFluidCell {
	static const bool kSplit = false;
	int* flags; int mark;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt p) const {
		if (!(flags[p] & TypeFluid) || !liquid::interiorCell(d, i, j, k)) return;
		if ((flags[p - 1] & TypeFluid) || (flags[p + 1] & TypeFluid) || (flags[p - d.Y] & TypeFluid) || (flags[p + d.Y] & TypeFluid)) return;
		if (d.is3D && ((flags[p - d.Z] & TypeFluid) || (flags[p + d.Z] & TypeFluid))) return;
		flags[p] = mark;
	}
};

// ---------------------------------------------------------------- mapPartsToMAC
// One thread per cell: its three faces gather from the particles bucketed (clamped) into the 3 x 3 (x 3) cells around it, visited in
// ascending particle order, which is the order the reference's serial scatter adds them in.  A particle with bucket cell (kx, ky, kz)
// has its face / centre bases in {k-1, k} per axis and spreads to base + {0, 1}, i.e. to faces within one cell of its bucket.
// stomp(VECTOR_EPSILON) grid.cpp:224-226, safeDivide general.h:148-151 and velOld.copyFrom(vel) (flip.cpp:590-594) are fused in.
template <typename Real> struct MapPartsGather {
	static const bool kSplit = false;
	const int* start; const int* sorted; IndexInt count; const Real* pos; const Real* pvel; Real* vel; Real* velOld; Real* weight;
	static MP_HD void add(Real& S, Real& R, Real a, Real b, Real c, Real v) { const Real w = b * (a * c); S += w; R += w * v; }
	// one component of setInterpolMAC (util/interpol.h:159-203) as seen from face (i, j, k): base (bx, by, bz), weights a (x), b (y), c (z)
	static MP_HD void comp(const Dims& d, int i, int j, int k, int bx, int by, int bz, const Real (&a)[2], const Real (&b)[2], const Real (&c)[2], Real v, bool zFirst, Real& S, Real& R) {
		const int di = i - bx, dj = j - by;
		if (di < 0 || di > 1 || dj < 0 || dj > 1) return;
		if (d.is3D) {
			const int dk = k - bz;
			if (dk < 0 || dk > 1) return;
			add(S, R, a[di], b[dj], c[dk], v);
		} else {            // Z stride 0: both z-weights land on the same face, in the order the reference adds them
			if (bz != 0) return;
			if (zFirst) { add(S, R, a[di], b[dj], c[1], v); add(S, R, a[di], b[dj], c[0], v); }
			else        { add(S, R, a[di], b[dj], c[0], v); add(S, R, a[di], b[dj], c[1], v); }
		}
	}
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		// segments of the neighbouring rows: the three cells (i-1, i, i+1) of a row are consecutive buckets, but only each cell's own
		// list is ascending -> one cursor per cell
		int cur[27], end[27], head[27];
		int nseg = 0;
		const int kz0 = d.is3D ? k - 1 : 0, kz1 = d.is3D ? k + 1 : 0;
		for (int zz = kz0; zz <= kz1; zz++) for (int yy = j - 1; yy <= j + 1; yy++) for (int xx = i - 1; xx <= i + 1; xx++) {
			if (xx < 0 || xx >= d.sx || yy < 0 || yy >= d.sy || zz < 0 || zz >= d.sz) continue;
			const IndexInt c = (IndexInt)xx + d.Y * yy + d.Z * zz;
			const int b = start[c], e = (c + 1 < d.n) ? start[c + 1] : (int)count;
			if (b < e) { cur[nseg] = b; end[nseg] = e; head[nseg] = sorted[b]; nseg++; }
		}
		Real S[3] = { 0, 0, 0 }, R[3] = { 0, 0, 0 };
		while (nseg > 0) {
			int best = 0;
			for (int q = 1; q < nseg; q++) if (head[q] < head[best]) best = q;
			const int p = head[best];
			if (++cur[best] < end[best]) head[best] = sorted[cur[best]];
			else { nseg--; cur[best] = cur[nseg]; end[best] = end[nseg]; head[best] = head[nseg]; }
			const MacWeights<Real> m(d, pos + 3 * (IndexInt)p);
			const Real* v = pvel + 3 * (IndexInt)p;
			comp(d, i, j, k, m.sxi, m.yi, m.zi, m.ss, m.t, m.f, v[0], true, S[0], R[0]);
			comp(d, i, j, k, m.xi, m.syi, m.zi, m.s, m.st, m.f, v[1], true, S[1], R[1]);
			comp(d, i, j, k, m.xi, m.yi, m.szi, m.s, m.t, m.sf, v[2], false, S[2], R[2]);
		}
		const Real eps = sizeof(Real) == 8 ? (Real)1e-10 : (Real)1e-6f;
		for (int c = 0; c < 3; c++) {
			Real w = S[c];
			if (w < eps) w = 0;
			const Real r = w ? (R[c] / w) : R[c];
			vel[3 * idx + c] = r; velOld[3 * idx + c] = r;
			if (weight) weight[3 * idx + c] = w;
		}
	}
};
// ---- the same gather through a tree of 3-way merges (default).  The 27-way merge above costs every face ~27 comparisons per visited
// particle, out of thread-private arrays that live in local memory, and reads positions through the particle index: 2.2 ns per particle at
// 256^3.  Here the ascending lists are built level by level -- x-triples (i-1, i, i+1) from the buckets, x-y blocks from the triples, and the
// face's own walk merges the three blocks (k-1, k, k+1) -- each a 3-way merge held in registers, each list built once and read by three
// consumers; entries carry the particle id (the merge key) and its slot in bucket order, where copies of pos / partVel lie next to their
// cell's neighbours.  Same particles in the same (ascending id) order through the same arithmetic: bit-identical to the 27-way walk.
struct Ent { int id, slot; };
struct alignas(16) Ent2 { Ent a, b; };
enum : int { kNoId = 0x7fffffff };
struct MergeAxis {                     // neighbours of cell (i, j, k) along `axis` that exist
	static MP_HD bool lo(const Dims& d, int axis, int i, int j, int k) { return axis == 0 ? i > 0 : (axis == 1 ? j > 0 : (d.is3D && k > 0)); }
	static MP_HD bool hi(const Dims& d, int axis, int i, int j, int k) { return axis == 0 ? i < d.sx - 1 : (axis == 1 ? j < d.sy - 1 : (d.is3D && k < d.sz - 1)); }
	static MP_HD IndexInt stride(const Dims& d, int axis) { return axis == 0 ? 1 : (axis == 1 ? d.Y : d.Z); }
};
// The lists of one level.  Level 0: ranges of `sorted` (start = first slot of a cell's bucket, lengths from the differences, entries are
// (sorted[s], s)).  Levels >= 1: materialised Ent lists; every list starts on a 32-byte boundary (start[] counts entries and is a multiple
// of 4, len[] holds the true lengths), so that a thread -- whose list is private to it and lies ~600 bytes from its neighbour lane's -- moves
// whole 32-byte sectors: read element by element, every sector of a list crossed DRAM four times (measured: 23.7 GB for 5.6 GB of lists).
struct Lists {
	const int* start; const int* len; IndexInt count; const Ent* ent; const int* sorted;
	MP_HD int first(IndexInt c) const { return start[c]; }
	MP_HD int length(const Dims& d, IndexInt c) const { return len ? len[c] : ((c + 1 < d.n) ? start[c + 1] : (int)count) - start[c]; }
};
// reader of one ascending list: the head entry in b0, up to three more of its sector behind it in registers
struct ListCur {
	const Ent* ent; const int* sorted; int pos, end, have; Ent b0, b1, b2, b3;
	MP_HD void fill() {
		have = 0; b0.id = kNoId;
		if (pos >= end) return;
		if (!ent) { b0.id = sorted[pos]; b0.slot = pos; have = 1; return; }
		const Ent2* q = reinterpret_cast<const Ent2*>(ent + pos);        // pos is a multiple of 4 here: two 16-byte loads of one sector
		const Ent2 u = q[0], v = q[1];
		b0 = u.a; b1 = u.b; b2 = v.a; b3 = v.b;
		have = end - pos < 4 ? end - pos : 4;
	}
	MP_HD void open(const Dims& d, const Lists& l, IndexInt c) { ent = l.ent; sorted = l.sorted; pos = l.first(c); end = pos + l.length(d, c); fill(); }
	MP_HD void close() { ent = nullptr; sorted = nullptr; pos = end = have = 0; b0.id = kNoId; }
	MP_HD void pop() { pos++; if (--have > 0) { b0 = b1; b1 = b2; b2 = b3; } else fill(); }
};
// lengths of the merged lists: true length and, rounded up to whole sectors, what the exclusive scan turns into the list's first entry
struct MergeLen {
	static const bool kSplit = false;
	Lists in; int axis; int* lenOut; int* offOut;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		const IndexInt st = MergeAxis::stride(d, axis);
		int n = in.length(d, idx);
		if (MergeAxis::lo(d, axis, i, j, k)) n += in.length(d, idx - st);
		if (MergeAxis::hi(d, axis, i, j, k)) n += in.length(d, idx + st);
		lenOut[idx] = n; offOut[idx] = (n + 3) & ~3;
	}
};
// three ascending lists -> one; `visit(Ent)` is called in ascending id order
struct Merge3 {
	Lists in; int axis;
	template <typename V> MP_HD void walk(const Dims& d, int i, int j, int k, IndexInt idx, V& visit) const {
		const IndexInt st = MergeAxis::stride(d, axis);
		ListCur a, b, c;
		b.open(d, in, idx);
		if (MergeAxis::lo(d, axis, i, j, k)) a.open(d, in, idx - st); else a.close();
		if (MergeAxis::hi(d, axis, i, j, k)) c.open(d, in, idx + st); else c.close();
		for (;;) {
			// ids are distinct: the smallest head is unique until all three lists are exhausted.  One call site of visit(): lanes that take
			// their entry from different lists stay converged through it
			const int mab = a.b0.id < b.b0.id ? a.b0.id : b.b0.id, mn = mab < c.b0.id ? mab : c.b0.id;
			if (mn == kNoId) break;
			const bool ta = a.b0.id == mn, tb = b.b0.id == mn;
			const Ent e = ta ? a.b0 : (tb ? b.b0 : c.b0);
			visit(e);
			if (ta) a.pop(); else if (tb) b.pop(); else c.pop();
		}
	}
};
struct MergeStore {
	static const bool kSplit = false;
	Merge3 m; const int* outStart; Ent* out;
	struct Put {                       // four entries = one sector at a time; the tail of the last sector is padding no reader looks at
		Ent* p; int n; Ent w0, w1, w2, w3;
		MP_HD void flush() { Ent2* q = reinterpret_cast<Ent2*>(p); Ent2 u = { w0, w1 }, v = { w2, w3 }; q[0] = u; q[1] = v; p += 4; n = 0; }
		MP_HD void operator()(const Ent& e) {
			if (n == 0) w0 = e; else if (n == 1) w1 = e; else if (n == 2) w2 = e; else w3 = e;
			if (++n == 4) flush();
		}
	};
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		const Ent z = { 0, 0 };
		Put put = { out + outStart[idx], 0, z, z, z, z };
		m.walk(d, i, j, k, idx, put);
		if (put.n) flush(put);
	}
	static MP_HD void flush(Put& put) { put.flush(); }
};
// copies of pos / partVel in bucket order
template <typename Real> struct GatherSlots {
	const int* sorted; const Real* pos; const Real* pvel; Real* posS; Real* pvelS;
	MP_HD void operator()(IndexInt s) const {
		const IndexInt p = sorted[s];
		for (int c = 0; c < 3; c++) { posS[3 * s + c] = pos[3 * p + c]; pvelS[3 * s + c] = pvel[3 * p + c]; }
	}
};
template <typename Real> MP_HD void finishFace(Real (&S)[3], Real (&R)[3], IndexInt idx, Real* vel, Real* velOld, Real* weight) {
	// stomp(VECTOR_EPSILON), safeDivide, velOld.copyFrom(vel)   flip.cpp:590-594
	const Real eps = sizeof(Real) == 8 ? (Real)1e-10 : (Real)1e-6f;
	for (int c = 0; c < 3; c++) {
		Real w = S[c];
		if (w < eps) w = 0;
		const Real r = w ? (R[c] / w) : R[c];
		vel[3 * idx + c] = r; velOld[3 * idx + c] = r;
		if (weight) weight[3 * idx + c] = w;
	}
}
template <typename Real> struct MapPartsGatherTree {
	static const bool kSplit = false;
	Merge3 m; const Real* posS; const Real* pvelS; Real* vel; Real* velOld; Real* weight;
	struct Acc {
		const Dims* d; int i, j, k; const Real* posS; const Real* pvelS; Real S[3], R[3];
		MP_HD void operator()(const Ent& e) {
			const MacWeights<Real> w(*d, posS + 3 * (IndexInt)e.slot);
			const Real* v = pvelS + 3 * (IndexInt)e.slot;
			MapPartsGather<Real>::comp(*d, i, j, k, w.sxi, w.yi, w.zi, w.ss, w.t, w.f, v[0], true, S[0], R[0]);
			MapPartsGather<Real>::comp(*d, i, j, k, w.xi, w.syi, w.zi, w.s, w.st, w.f, v[1], true, S[1], R[1]);
			MapPartsGather<Real>::comp(*d, i, j, k, w.xi, w.yi, w.szi, w.s, w.t, w.sf, v[2], false, S[2], R[2]);
		}
	};
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		Acc acc = { &d, i, j, k, posS, pvelS, { 0, 0, 0 }, { 0, 0, 0 } };
		m.walk(d, i, j, k, idx, acc);
		finishFace<Real>(acc.S, acc.R, idx, vel, velOld, weight);
	}
};
// ---- 3-D, every axis < 65536 cells: the walk above evaluates the MAC weights of a particle once per visiting cell, 27 times (ncu: 13 G warp
// instructions for 26 M particles, 15 of 32 lanes active).  Instead every particle gets a record, written once in bucket order: its three
// face bases and, per component, the 8 products w = b (a c) and w v that setInterpolMAC adds to the faces base + {0,1}^3 (util/interpol.h:
// 159-203).  A visiting cell looks its own pair up; S += w, R += w v are the reference's additions of the reference's products.
template <typename Real> struct alignas(16) PartRec {
	unsigned bx, by, bz, pad;          // shifted base | centred base << 16, per axis (bases are >= 0 and < 65536)
	Real tab[3][8][2];                  // [component][di + 2 dj + 4 dk] -> { w, w v }
};
template <typename Real> struct BuildRecords {
	Dims d; const int* sorted; const Real* pos; const Real* pvel; PartRec<Real>* rec;
	MP_HD void operator()(IndexInt s) const {
		const IndexInt p = sorted[s];
		const MacWeights<Real> m(d, pos + 3 * p);
		const Real* v = pvel + 3 * p;
		PartRec<Real> r;
		r.bx = (unsigned)m.sxi | ((unsigned)m.xi << 16); r.by = (unsigned)m.syi | ((unsigned)m.yi << 16); r.bz = (unsigned)m.szi | ((unsigned)m.zi << 16); r.pad = 0;
		for (int dk = 0; dk < 2; dk++) for (int dj = 0; dj < 2; dj++) for (int di = 0; di < 2; di++) {
			const int t = di + 2 * dj + 4 * dk;
			Real w;
			w = m.t[dj] * (m.ss[di] * m.f[dk]);  r.tab[0][t][0] = w; r.tab[0][t][1] = w * v[0];      // MapPartsGather::add with (a, b, c) = (x, y, z weights)
			w = m.st[dj] * (m.s[di] * m.f[dk]);  r.tab[1][t][0] = w; r.tab[1][t][1] = w * v[1];
			w = m.t[dj] * (m.s[di] * m.sf[dk]);  r.tab[2][t][0] = w; r.tab[2][t][1] = w * v[2];
		}
		rec[s] = r;
	}
};
template <typename Real> struct MapPartsGatherRec {
	static const bool kSplit = false;
	Merge3 m; const PartRec<Real>* rec; Real* vel; Real* velOld; Real* weight;
	struct Acc {
		int i, j, k; const PartRec<Real>* rec; Real S[3], R[3];
		MP_HD void one(const PartRec<Real>& r, int c, int bi, int bj, int bk) {
			const unsigned di = (unsigned)(i - bi), dj = (unsigned)(j - bj), dk = (unsigned)(k - bk);
			if (di < 2u && dj < 2u && dk < 2u) { const Real* q = r.tab[c][di + 2 * dj + 4 * dk]; S[c] += q[0]; R[c] += q[1]; }
		}
		MP_HD void operator()(const Ent& e) {
			const PartRec<Real>& r = rec[e.slot];
			const unsigned bx = r.bx, by = r.by, bz = r.bz;
			one(r, 0, (int)(bx & 0xffffu), (int)(by >> 16), (int)(bz >> 16));        // x faces: shifted in x, centred in y, z
			one(r, 1, (int)(bx >> 16), (int)(by & 0xffffu), (int)(bz >> 16));
			one(r, 2, (int)(bx >> 16), (int)(by >> 16), (int)(bz & 0xffffu));
		}
	};
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		Acc acc = { i, j, k, rec, { 0, 0, 0 }, { 0, 0, 0 } };
		m.walk(d, i, j, k, idx, acc);
		finishFace<Real>(acc.S, acc.R, idx, vel, velOld, weight);
	}
};

// scratch of the merge tree, provided by the caller.  A particle enters 3 x-triples and 9 x-y blocks, and a non-empty list is padded by up to
// three entries: e1 holds 3 np + 3 min(cells, 3 np) entries at most, e2 9 np + 3 min(cells, 9 np)  (treeEntries below)
template <typename Real> struct MapPartsTreeScratch {
	int* len1; int* off1; int* len2; int* off2;      // int[d.n] each
	Ent* e1; Ent* e2;
	Real* posS; Real* pvelS;                          // Real[3 np] each (used when rec == NULL)
	PartRec<Real>* rec;                               // PartRec[np], or NULL: evaluate the weights in the walk (2-D, huge grids or particle counts)
};
inline IndexInt treeEntries(IndexInt cells, IndexInt np, int copies) { const IndexInt t = (IndexInt)copies * np; return t + 3 * (cells < t ? cells : t) + 4; }
// start: scratch int[d.n]; key / keyTmp / val / sorted: scratch int[np]; tree == NULL: the 27-way walk
template <typename Real, typename Exec>
int mapPartsToMAC(Exec& ex, const Dims& d, Real* vel, Real* velOld, IndexInt np, const PSet<Real>& ps, const Real* pvel, Real* weight,
                  int* start, int* key, int* keyTmp, int* val, int* sorted, const MapPartsTreeScratch<Real>* tree = nullptr) {
	IndexInt count = 0;
	MP_TRY(bucketParticles<Real>(ex, d, np, ps, true, start, key, keyTmp, val, sorted, &count));
	if (!tree) {
		MapPartsGather<Real> op = { start, sorted, count, ps.pos, pvel, vel, velOld, weight };
		return ex.cells(d, op);
	}
	const bool useRec = tree->rec && d.is3D && d.sx < 65536 && d.sy < 65536 && d.sz < 65536;
	if (count > 0) {
		if (useRec) { BuildRecords<Real> br = { d, sorted, ps.pos, pvel, tree->rec }; MP_TRY(ex.parts(count, br)); }
		else { GatherSlots<Real> gs = { sorted, ps.pos, pvel, tree->posS, tree->pvelS }; MP_TRY(ex.parts(count, gs)); }
	}
	IndexInt n1 = 0, n2 = 0;
	const Lists l0 = { start, nullptr, count, nullptr, sorted };
	MergeLen m1 = { l0, 0, tree->len1, tree->off1 }; MP_TRY(ex.cells(d, m1)); MP_TRY(ex.exclusiveScan(tree->off1, d.n, &n1));
	MergeStore s1 = { { l0, 0 }, tree->off1, tree->e1 }; MP_TRY(ex.cells(d, s1));
	const Lists l1 = { tree->off1, tree->len1, n1, tree->e1, nullptr };
	MergeLen m2 = { l1, 1, tree->len2, tree->off2 }; MP_TRY(ex.cells(d, m2)); MP_TRY(ex.exclusiveScan(tree->off2, d.n, &n2));
	MergeStore s2 = { { l1, 1 }, tree->off2, tree->e2 }; MP_TRY(ex.cells(d, s2));
	const Lists l2 = { tree->off2, tree->len2, n2, tree->e2, nullptr };
	if (useRec) { MapPartsGatherRec<Real> op = { { l2, 2 }, tree->rec, vel, velOld, weight }; return ex.cells(d, op); }
	MapPartsGatherTree<Real> op = { { l2, 2 }, tree->posS, tree->pvelS, vel, velOld, weight };
	return ex.cells(d, op);
}

}  // namespace parts

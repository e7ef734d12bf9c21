// Liquid neighbours of the pressure projection (first slice): the plugins that make the level-set free-surface loop of
// scenes/freesurface.py:54-84 device-resident around solvePressure(phi=...):
//   extrapolateMACSimple   fastmarch.cpp:337-375  (knExtrapolateMACSimple :231-258, knUnprojectNormalComp :319-331, knExtrapolateIntoBnd :260-299)
//   extrapolateLsSimple    fastmarch.cpp:470-507  (knExtrapolateLsSimple :439-460, knSetRemaining :463-467)
//   extrapolateVec3Simple  fastmarch.cpp:510-542
//   FlagGrid::updateFromLevelset  grid.cpp:844-854
//   Grid<T>::setBound             grid.cpp:585-593
//
// This header holds the per-cell operations and the pass sequences, written against an executor `Exec` with one member
//     template <class F> int cells(const Dims& d, const F& f);     // run f(d, i, j, k, idx) once for every cell, in any order
// mp_liquid.cu instantiates it with the CUDA launcher (threadCells below: a few cells per thread).  tests/emul/liquid_emul.cpp instantiates the SAME code with a
// host loop, so that the build container (which has no GPU) can check cell arithmetic and pass structure against the reference; that
// shim is test infrastructure and is never part of libmantapress.so.
//
// Why any execution order gives the serial result: a pass `d` only reads marks equal to d and values of cells carrying that mark, and
// only writes cells whose mark is 0 (new mark d+1) -- the sets read and written by one pass are disjoint.
#pragma once
#include <cmath>
#include "mp_common.cuh"

#ifdef __CUDACC__
#define MP_HD __host__ __device__ __forceinline__
#else
#define MP_HD inline
#endif

namespace liquid {

// Launch geometry of the CUDA executor.  One cell per thread makes these passes block-scheduling bound (a million 128-thread blocks of
// a few instructions each at 512^3: 0.55 ms for a pass that touches nothing), so a block of kThreads threads covers kRows rows of up
// to kThreads * kCols cells: thread tx owns the cells i = (bx * kCols + c) * kThreads + tx, c < kCols, of rows by * kRows + r, r < kRows
// -- consecutive lanes stay on consecutive cells.  The host emulation walks the same function (order 3), so the cover is checked.
static const int kThreads = 128, kCols = 4, kRows = 4;
struct LaunchGeom { unsigned gx, gy, gz; };
inline LaunchGeom launchGeomOf(const Dims& d) {
	LaunchGeom g;
	g.gx = (unsigned)((d.sx + kThreads * kCols - 1) / (kThreads * kCols)); g.gy = (unsigned)((d.sy + kRows - 1) / kRows); g.gz = (unsigned)d.sz;
	return g;
}
// Functors with kSplit == true separate their reads from their writes: State load(...) only reads, apply(..., State) computes and
// writes.  A thread first loads for all its kCols cells of a row and then applies -- the loads of several cells are in flight together
// instead of one dependent round trip after the other (the passes are latency bound otherwise: 1 ms per pass at 512^3).  Any order of
// loads and applies of different cells is as good as any other (see the note on execution order above).
struct NoSink { MP_HD void operator()(IndexInt) const {} };
template <typename F, typename Sink = NoSink> MP_HD void threadCells(const Dims& d, const F& f, int bx, int by, int bz, int tx, const Sink& sink = Sink()) {
	const IndexInt plane = (IndexInt)d.sx * d.sy * bz;
	for (int r = 0; r < kRows; r++) {
		const int j = by * kRows + r;
		if (j >= d.sy) break;
		const IndexInt row = d.Y * j + plane;
		if constexpr (F::kSplit) {
			typename F::State st[kCols];
			#pragma unroll
			for (int c = 0; c < kCols; c++) {
				const int i = (bx * kCols + c) * kThreads + tx;
				if (i < d.sx) st[c] = f.load(d, i, j, bz, (IndexInt)i + row);
			}
			#pragma unroll
			for (int c = 0; c < kCols; c++) {
				const int i = (bx * kCols + c) * kThreads + tx;
				if (i < d.sx) { if (f.apply(d, i, j, bz, (IndexInt)i + row, st[c])) sink((IndexInt)i + row); }
			}
		} else {
			for (int c = 0; c < kCols; c++) {
				const int i = (bx * kCols + c) * kThreads + tx;
				if (i >= d.sx) break;
				f(d, i, j, bz, (IndexInt)i + row);
			}
		}
	}
}
// what a cell-by-cell executor (the host emulation's plain walks) calls
template <typename F> MP_HD void oneCell(const Dims& d, const F& f, int i, int j, int k, IndexInt idx) {
	if constexpr (F::kSplit) f.apply(d, i, j, k, idx, f.load(d, i, j, k, idx));
	else f(d, i, j, k, idx);
}

MP_HD bool interiorCell(const Dims& d, int i, int j, int k) {      // the cells of a KERNEL(bnd=1) / FOR_IJK_BND(g, 1)
	return i >= 1 && i <= d.sx - 2 && j >= 1 && j <= d.sy - 2 && (!d.is3D || (k >= 1 && k <= d.sz - 2));
}
MP_HD int outerFaces(const Dims& d, int i, int j, int k) {           // on how many outer faces of the domain the cell lies (per axis at most one)
	return (i == 0 || i == d.sx - 1 ? 1 : 0) + (j == 0 || j == d.sy - 1 ? 1 : 0) + (d.is3D && (k == 0 || k == d.sz - 1) ? 1 : 0);
}
// neighbour q of fastmarch.cpp:233-236: +x -x +y -y +z -z
MP_HD IndexInt nbOffset(const Dims& d, int q) {
	const IndexInt o = (q >> 1) == 0 ? d.X : ((q >> 1) == 1 ? d.Y : d.Z);
	return (q & 1) ? -o : o;
}
MP_HD bool nbInterior(const Dims& d, int q, int i, int j, int k) {
	const int s = (q & 1) ? -1 : 1;
	return interiorCell(d, i + ((q >> 1) == 0 ? s : 0), j + ((q >> 1) == 1 ? s : 0), k + ((q >> 1) == 2 ? s : 0));
}

// ---------------------------------------------------------------- extrapolateMACSimple
// The reference runs the three velocity components one after the other with one Grid<int> of marks each; they are independent, so
// here byte c of ONE int per cell carries the mark of component c and a pass advances all components at once (distance <= 250).
template <typename Real> struct MacMark {
	static const bool kSplit = false;                // fastmarch.cpp:346-358
	const int* flags; int* tmp; int intoObs;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		int w = 0;
		if (interiorCell(d, i, j, k)) {
			const int f = flags[idx];
			const int dim = d.is3D ? 3 : 2;
			for (int c = 0; c < dim; c++) {
				const int fn = flags[idx - (c == 0 ? d.X : (c == 1 ? d.Y : d.Z))];
				bool mark = (f & TypeFluid) || (fn & TypeFluid);
				if (intoObs) mark = mark && !(f & TypeObstacle) && !(fn & TypeObstacle);
				if (mark) w |= 1 << (8 * c);
			}
		}
		tmp[idx] = w;
	}
};
template <typename Real> struct MacExtrapolate {         // knExtrapolateMACSimple fastmarch.cpp:231-258, pass d for every component
	static const bool kSplit = true;
	Real* vel; int* tmp; int pass;
	struct State { int t, tn[6]; bool interior; };
	// a neighbour's word may be rewritten by its own thread meanwhile: its changing bytes go from 0 to pass+1, never through `pass`
	MP_HD State load(const Dims& d, int i, int j, int k, IndexInt idx) const {
		State s;
		s.interior = interiorCell(d, i, j, k);
		if (!s.interior) return s;
		s.t = tmp[idx];
		s.tn[0] = tmp[idx + d.X]; s.tn[1] = tmp[idx - d.X]; s.tn[2] = tmp[idx + d.Y]; s.tn[3] = tmp[idx - d.Y];
		s.tn[4] = tmp[idx + d.Z]; s.tn[5] = tmp[idx - d.Z];            // 2-D: Z == 0, the cell itself; not used
		return s;
	}
	MP_HD bool reached(int w) const { return (w & 255) == pass || ((w >> 8) & 255) == pass || ((w >> 16) & 255) == pass; }      // some component carries the mark of this pass
	// non-zero iff one of the three mark bytes of w equals `pass` (bytes of x = w ^ pass-in-every-byte that are zero; marks are < 128)
	MP_HD unsigned reachedBits(int w) const { const unsigned x = ((unsigned)w ^ (0x010101u * (unsigned)pass)) & 0xffffffu; return (x - 0x010101u) & ~x & 0x808080u; }
	MP_HD bool apply(const Dims& d, int, int, int, IndexInt idx, const State& s) const {      // true: the cell got a new mark
		if (!s.interior) return false;
		const int dim = d.is3D ? 3 : 2;
		const int t = s.t;
		// the bulk of the grid: every component already marked, or no neighbour reached in this pass -- nothing to do (one test instead of
		// eighteen byte comparisons; the pass over all cells was instruction bound)
		if (pass < 128) {
			unsigned any = 0;
			for (int q = 0; q < 2 * dim; q++) any |= reachedBits(s.tn[q]);
			if (!any) return false;
		}
		int tNew = t;
		for (int c = 0; c < dim; c++) {
			if (((t >> (8 * c)) & 255) != 0) continue;
			int nbs = 0; Real avgVel = 0;
			for (int q = 0; q < 2 * dim; q++)
				if (((s.tn[q] >> (8 * c)) & 255) == pass) { avgVel += vel[3 * (idx + nbOffset(d, q)) + c]; nbs++; }
			if (nbs > 0) { tNew |= (pass + 1) << (8 * c); vel[3 * idx + c] = avgVel / (Real)nbs; }
		}
		if (tNew != t) tmp[idx] = tNew;
		return tNew != t;
	}
};
// normalize util/vectorbase.h:415-429: the comparison against 1. and the reciprocal are double expressions
template <typename Real> MP_HD void normalize3(Real& x, Real& y, Real& z) {
	const Real l = x * x + y * y + z * z;
	const Real eps2 = sizeof(Real) == 4 ? (Real)(1e-6f * 1e-6f) : (Real)(1e-10 * 1e-10);      // VECTOR_EPSILON^2 vectorbase.h:52,:55
	const double dl = (double)l - 1.;
	if ((dl < 0 ? -dl : dl) < (double)eps2) return;
	if (l > eps2) {
		const Real norm = sizeof(Real) == 4 ? (Real)sqrtf((float)l) : (Real)sqrt((double)l);
		const Real s = (Real)(1. / (double)norm);
		x *= s; y *= s; z *= s;
	} else { x = 0; y = 0; z = 0; }
}
template <typename Real> struct UnprojectNormal {
	static const bool kSplit = false;        // knUnprojectNormalComp fastmarch.cpp:319-331 with getNormal :302-318
	Real* vel; const Real* phi; Real maxDist;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		if (!interiorCell(d, i, j, k)) return;
		const Real ph = phi[idx];
		if (ph > 0. || ph < -maxDist) return;
		Real nx = phi[idx + d.X] - phi[idx - d.X], ny = phi[idx + d.Y] - phi[idx - d.Y], nz = phi[idx + d.Z] - phi[idx - d.Z];   // 2-D: Z == 0
		const Real vx = vel[3 * idx], vy = vel[3 * idx + 1], vz = vel[3 * idx + 2];
		if (nx * vx + ny * vy + nz * vz < 0.) {
			normalize3<Real>(nx, ny, nz);
			const Real l = nx * vx + ny * vy + nz * vz;
			vel[3 * idx] = vx - nx * l; vel[3 * idx + 1] = vy - ny * l; vel[3 * idx + 2] = vz - nz * l;
		}
	}
};
// knExtrapolateIntoBnd fastmarch.cpp:260-299 reads a copy of the whole velocity grid; only the cells of the outer layer are written
// and they read cells one step inwards, so the new values are first collected in `stage` (outer-layer entries only) and then moved.
template <typename Real> struct IntoBndStage {
	static const bool kSplit = false;
	const int* flags; const Real* vel; Real* stage;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		if (outerFaces(d, i, j, k) == 0) return;
		int c = 0;
		Real v0 = 0, v1 = 0, v2 = 0;
		const bool isObs = flags[idx] & TypeObstacle;
		#define MP_TAKE(off) { const Real* s_ = vel + 3 * (idx + (off)); v0 = s_[0]; v1 = s_[1]; v2 = s_[2]; }
		if (i == 0)              { MP_TAKE(d.X)  if (isObs && v0 < 0.) v0 = 0; c++; }
		else if (i == d.sx - 1)  { MP_TAKE(-d.X) if (isObs && v0 > 0.) v0 = 0; c++; }
		if (j == 0)              { MP_TAKE(d.Y)  if (isObs && v1 < 0.) v1 = 0; c++; }
		else if (j == d.sy - 1)  { MP_TAKE(-d.Y) if (isObs && v1 > 0.) v1 = 0; c++; }
		if (d.is3D) {
			if (k == 0)              { MP_TAKE(d.Z)  if (isObs && v2 < 0.) v2 = 0; c++; }
			else if (k == d.sz - 1)  { MP_TAKE(-d.Z) if (isObs && v2 > 0.) v2 = 0; c++; }
		}
		#undef MP_TAKE
		stage[3 * idx] = v0 / (Real)c; stage[3 * idx + 1] = v1 / (Real)c; stage[3 * idx + 2] = v2 / (Real)c;
	}
};
template <typename Real> struct IntoBndCopy {
	static const bool kSplit = false;
	const Real* stage; Real* vel;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		if (outerFaces(d, i, j, k) == 0) return;
		vel[3 * idx] = stage[3 * idx]; vel[3 * idx + 1] = stage[3 * idx + 1]; vel[3 * idx + 2] = stage[3 * idx + 2];
	}
};

// extrapolateMACFromWeight fastmarch.cpp:410-432: the marks live in a Vec3 weight grid (what mapPartsToMAC leaves behind).  Components are
// independent, so one reset pass and `distance` extrapolation passes serve all three (the reference runs them one component after the other).
template <typename Real> struct WeightReset {            // fastmarch.cpp:419-422
	static const bool kSplit = false;
	Real* weight;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		if (!interiorCell(d, i, j, k)) return;
		const int dim = d.is3D ? 3 : 2;
		for (int c = 0; c < dim; c++) if (weight[3 * idx + c] > 0.) weight[3 * idx + c] = 1.0;
	}
};
template <typename Real> struct WeightExtrapolate {      // knExtrapolateMACFromWeight fastmarch.cpp:378-403
	static const bool kSplit = false;
	Real* vel; Real* weight; int pass;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		if (!interiorCell(d, i, j, k)) return;
		const int dim = d.is3D ? 3 : 2;
		for (int c = 0; c < dim; c++) {
			if (weight[3 * idx + c] != 0) continue;
			int nbs = 0; Real avgVel = 0;
			for (int q = 0; q < 2 * dim; q++) {
				const IndexInt nb = idx + nbOffset(d, q);
				if (weight[3 * nb + c] == (Real)pass) { avgVel += vel[3 * nb + c]; nbs++; }     // a neighbour being marked right now goes 0 -> pass+1, never through `pass`
			}
			if (nbs > 0) { weight[3 * idx + c] = (Real)(pass + 1); vel[3 * idx + c] = avgVel / (Real)nbs; }
		}
	}
};
template <typename Real, typename Exec>
int extrapolateMacFromWeight(Exec& ex, const Dims& d, Real* vel, Real* weight, int distance) {
	{ WeightReset<Real> op = { weight }; MP_TRY(ex.cells(d, op)); }
	for (int pass = 1; pass < 1 + distance; pass++) { WeightExtrapolate<Real> op = { vel, weight, pass }; MP_TRY(ex.cells(d, op)); }
	return MP_OK;
}

// tmp: one int per cell; stage: 3 Reals per cell (only outer-layer entries are touched)
template <typename Real, typename Exec>
int extrapolateMacSimple(Exec& ex, const Dims& d, const int* flags, Real* vel, int distance, const Real* phiObs, bool intoObs, int* tmp, Real* stage) {
	{ MacMark<Real> op = { flags, tmp, intoObs ? 1 : 0 }; MP_TRY(ex.cells(d, op)); }
	for (int pass = 1; pass < 1 + distance; pass++) { MacExtrapolate<Real> op = { vel, tmp, pass }; MP_TRY(ex.cells(d, op)); }
	if (phiObs) { UnprojectNormal<Real> op = { vel, phiObs, (Real)distance }; MP_TRY(ex.cells(d, op)); }
	{ IntoBndStage<Real> op = { flags, vel, stage }; MP_TRY(ex.cells(d, op)); }
	{ IntoBndCopy<Real> op = { stage, vel }; MP_TRY(ex.cells(d, op)); }
	return MP_OK;
}

// ---------------------------------------------------------------- extrapolateLsSimple / extrapolateVec3Simple
template <typename Real> struct LsMark {                 // fastmarch.cpp:475-498: 1 on the chosen side of phi, 2 on the first layer next to it, 0 elsewhere
	static const bool kSplit = true;
	const Real* phi; int* tmp; int inside;
	struct State { Real p, pn[6]; bool interior; };
	MP_HD bool on(Real p) const { return inside ? (p > (Real)0) : (p < (Real)0); }      // the reference compares with the double 0.: same outcome for every Real
	MP_HD State load(const Dims& d, int i, int j, int k, IndexInt idx) const {
		State s;
		s.interior = interiorCell(d, i, j, k);
		if (!s.interior) return s;
		s.p = phi[idx];
		s.pn[0] = phi[idx + d.X]; s.pn[1] = phi[idx - d.X]; s.pn[2] = phi[idx + d.Y]; s.pn[3] = phi[idx - d.Y];
		s.pn[4] = phi[idx + d.Z]; s.pn[5] = phi[idx - d.Z];
		return s;
	}
	MP_HD bool apply(const Dims& d, int i, int j, int k, IndexInt idx, const State& s) const {      // true: first layer next to the chosen side (mark 2)
		int m = 0;
		if (s.interior) {
			if (on(s.p)) m = 1;
			else {
				// cells of the outer layer carry no mark: neighbour q (+x -x +y -y +z -z) of an interior cell is interior iff it is not on the layer
				const bool ok[6] = { i < d.sx - 2, i > 1, j < d.sy - 2, j > 1, k < d.sz - 2, k > 1 };
				const int dim = d.is3D ? 3 : 2;
				for (int q = 0; q < 2 * dim; q++)
					if (ok[q] && on(s.pn[q])) { m = 2; break; }
			}
		}
		tmp[idx] = m;
		return m == 2;
	}
};
// knExtrapolateLsSimple<S> fastmarch.cpp:439-460 (NC = 1: Real, 3: Vec3).  `last`: this is the final pass, so cells that stay unmarked
// get knSetRemaining's value (:463-467) right away -- nobody reads the value of an unmarked cell during the pass.
template <typename Real, int NC> struct LsExtrapolate {
	static const bool kSplit = true;
	Real* val; int* tmp; int pass; Real direction; int last; Real remaining;
	struct State { int t, tn[6]; bool interior; };
	MP_HD State load(const Dims& d, int i, int j, int k, IndexInt idx) const {
		State s;
		s.interior = interiorCell(d, i, j, k);
		if (!s.interior) return s;
		s.t = tmp[idx];
		s.tn[0] = tmp[idx + d.X]; s.tn[1] = tmp[idx - d.X]; s.tn[2] = tmp[idx + d.Y]; s.tn[3] = tmp[idx - d.Y];
		s.tn[4] = tmp[idx + d.Z]; s.tn[5] = tmp[idx - d.Z];
		return s;
	}
	MP_HD bool reached(int w) const { return w == pass; }
	MP_HD bool apply(const Dims& d, int, int, int, IndexInt idx, const State& s) const {      // true: the cell got the mark pass + 1
		if (!s.interior || s.t != 0) return false;
		const int dim = d.is3D ? 3 : 2;
		int nbs = 0;
		Real avg[NC];
		for (int c = 0; c < NC; c++) avg[c] = 0;
		for (int q = 0; q < 2 * dim; q++) {
			if (s.tn[q] != pass) continue;
			const IndexInt nb = idx + nbOffset(d, q);
			for (int c = 0; c < NC; c++) avg[c] += val[NC * nb + c];
			nbs++;
		}
		if (nbs > 0) { tmp[idx] = pass + 1; for (int c = 0; c < NC; c++) val[NC * idx + c] = avg[c] / (Real)nbs + direction; return true; }
		if (last) for (int c = 0; c < NC; c++) val[NC * idx + c] = remaining;
		return false;
	}
};
template <typename Real, int NC> struct LsRemaining {
	static const bool kSplit = false;    // knSetRemaining fastmarch.cpp:463-467 when there was no pass to ride on
	Real* val; const int* tmp; Real remaining;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		if (!interiorCell(d, i, j, k) || tmp[idx] != 0) return;
		for (int c = 0; c < NC; c++) val[NC * idx + c] = remaining;
	}
};
template <typename Real, int NC, typename Exec>
int extrapolateLs(Exec& ex, const Dims& d, Real* val, const Real* phi, int distance, bool inside, Real direction, Real remaining, int* tmp) {
	{ LsMark<Real> op = { phi, tmp, inside ? 1 : 0 }; MP_TRY(ex.cells(d, op)); }
	if (distance < 2) { LsRemaining<Real, NC> op = { val, tmp, remaining }; return ex.cells(d, op); }
	for (int pass = 2; pass < 1 + distance; pass++) {
		LsExtrapolate<Real, NC> op = { val, tmp, pass, direction, pass == distance ? 1 : 0, remaining };
		MP_TRY(ex.cells(d, op));
	}
	return MP_OK;
}

// ---------------------------------------------------------------- FlagGrid::updateFromLevelset, Grid<T>::setBound
template <typename Real> struct UpdateFromLevelset {
	static const bool kSplit = false;     // grid.cpp:844-854; invalidTimeValue = -1000 (levelset.cpp:103 -> fastmarch.h:134)
	int* flags; const Real* phi;
	MP_HD void operator()(const Dims&, int, int, int, IndexInt idx) const {
		const int f = flags[idx];
		if ((f & TypeObstacle) || (f & TypeOutflow)) return;
		const Real p = phi[idx];
		if (p <= (Real)-1000) return;
		flags[idx] = (f & ~(TypeEmpty | TypeFluid)) | ((p <= 0) ? TypeFluid : TypeEmpty);
	}
};
template <typename T, int NC> struct SetBound {
	static const bool kSplit = false;          // knSetBoundary grid.cpp:585-589
	T* g; T value[NC]; int w;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt idx) const {
		const bool bnd = i <= w || i >= d.sx - 1 - w || j <= w || j >= d.sy - 1 - w || (d.is3D && (k <= w || k >= d.sz - 1 - w));
		if (bnd) for (int c = 0; c < NC; c++) g[NC * idx + c] = value[c];
	}
};

// ---------------------------------------------------------------- setWallBcs, second-order variant
// KnSetWallBcsFrac plugin/extforces.cpp:220-303 (taken by setWallBcs when fractions AND phiObs are given, :307-316): on faces next to an
// obstacle the velocity component along the obstacle normal (gradient of phiObs at the face) is removed.  Writes a fresh grid (every cell).
template <typename Real> MP_HD Real halfSum(Real a, Real b) { return (Real)((double)(a + b) * .5); }
template <typename Real> MP_HD void macAtFace(const Dims& d, const Real* v, IndexInt idx, int c, Real& ox, Real& oy, Real& oz) {      // MACGrid::getAtMACX/Y/Z grid.h:437-470
	const IndexInt Y = d.Y, Z = d.Z;
	#define MP_VC(o, cc) v[3 * (idx + (o)) + (cc)]
	if (c == 0) {
		ox = MP_VC(0, 0);
		oy = (Real)(0.25 * (double)(MP_VC(0, 1) + MP_VC(-1, 1) + MP_VC(Y, 1) + MP_VC(Y - 1, 1)));
		oz = 0;
		if (d.is3D) oz = (Real)(0.25 * (double)(MP_VC(0, 2) + MP_VC(-1, 2) + MP_VC(Z, 2) + MP_VC(Z - 1, 2)));
	} else if (c == 1) {
		ox = (Real)(0.25 * (double)(MP_VC(0, 0) + MP_VC(-Y, 0) + MP_VC(1, 0) + MP_VC(1 - Y, 0)));
		oy = MP_VC(0, 1);
		oz = 0;
		if (d.is3D) oz = (Real)(0.25 * (double)(MP_VC(0, 2) + MP_VC(-Y, 2) + MP_VC(Z, 2) + MP_VC(Z - Y, 2)));
	} else {
		ox = (Real)(0.25 * (double)(MP_VC(0, 0) + MP_VC(-Z, 0) + MP_VC(1, 0) + MP_VC(1 - Z, 0)));
		oy = (Real)(0.25 * (double)(MP_VC(0, 1) + MP_VC(-Z, 1) + MP_VC(Y, 1) + MP_VC(Y - Z, 1)));
		oz = MP_VC(0, 2);
	}
	#undef MP_VC
}
template <typename Real> struct WallBcsFrac {
	static const bool kSplit = false;
	const int* flags; const Real* vel; Real* tgt; const Real* phiObs;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt p) const {
		Real out[3] = { vel[3 * p], vel[3 * p + 1], vel[3 * p + 2] };
		const int f = flags[p];
		const bool curFluid = f & TypeFluid, curObs = f & TypeObstacle;
		if ((curFluid || curObs) && interiorCell(d, i, j, k)) {
			const int dim = d.is3D ? 3 : 2;
			const IndexInt S[3] = { d.X, d.Y, d.Z };
			for (int c = 0; c < dim; c++) {
				if (!(curObs || (flags[p - S[c]] & TypeObstacle))) continue;
				Real dphi[3] = { 0, 0, 0 };
				const Real tmp1 = halfSum<Real>(phiObs[p], phiObs[p - S[c]]);
				for (int a = 0; a < dim; a++) {
					if (a == c) { dphi[a] = phiObs[p] - phiObs[p - S[c]]; continue; }
					Real tmp2 = halfSum<Real>(phiObs[p + S[a]], phiObs[p + S[a] - S[c]]);
					const Real phi1 = halfSum<Real>(tmp1, tmp2);
					tmp2 = halfSum<Real>(phiObs[p - S[a]], phiObs[p - S[a] - S[c]]);
					const Real phi2 = halfSum<Real>(tmp1, tmp2);
					dphi[a] = phi1 - phi2;
				}
				normalize3<Real>(dphi[0], dphi[1], dphi[2]);
				Real vx, vy, vz;
				macAtFace<Real>(d, vel, p, c, vx, vy, vz);
				const Real vm = c == 0 ? vx : (c == 1 ? vy : vz);
				out[c] = vm - (dphi[0] * vx + dphi[1] * vy + dphi[2] * vz) * dphi[c];
			}
		}
		tgt[3 * p] = out[0]; tgt[3 * p + 1] = out[1]; tgt[3 * p + 2] = out[2];
	}
};

// ---------------------------------------------------------------- updateFractions / setObstacleFlags
// plugin/initplugins.cpp:437-440 (KnUpdateFractions :371-434, calcFraction :356-369) and :473-475 (KnUpdateFlagsObs :443-470): the producers of
// the `fractions` argument of solvePressure / setWallBcs.  The reference's kernel also writes the +x / +y / +z NEIGHBOUR of a cell next to an
// open / inflow / outflow wall; in the serial order of its loop such a write only survives when the neighbour lies on the outer layer (an
// interior neighbour recomputes its own entry afterwards), so here every cell GATHERS: interior cells compute their own entry and the min-side
// overrides, outer-layer cells take the max-side override of their inner neighbour.  The max-z test reads j, as the reference does (:423).
template <typename Real> MP_HD Real calcFraction(Real phi1, Real phi2, Real fracThreshold) {
	if (phi1 > 0. && phi2 > 0.) return 1.;
	if (phi1 < 0. && phi2 < 0.) return 0.;
	if (phi2 < phi1) { const Real t = phi1; phi1 = phi2; phi2 = t; }
	const Real denom = phi1 - phi2;
	if (denom > -1e-04) return 0.5;
	Real frac = (Real)(1. - (double)(phi1 / denom));
	if (frac < fracThreshold) frac = 0.;
	return (Real)1 < frac ? (Real)1 : frac;           // std::min(Real(1), frac)
}
MP_HD bool openKind(int f) { return (f & TypeInflow) || (f & TypeOutflow) || (f & TypeOpen); }
template <typename Real> struct UpdateFractions {
	static const bool kSplit = false;
	const int* flags; const Real* phiObs; Real* fractions; int w; Real fracThreshold;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt q) const {
		Real fx = 0, fy = 0, fz = 0;
		bool one = false;
		if (interiorCell(d, i, j, k)) {
			const Real ph = phiObs[q];
			fx = calcFraction<Real>(ph, phiObs[q - d.X], fracThreshold);
			fy = calcFraction<Real>(ph, phiObs[q - d.Y], fracThreshold);
			if (d.is3D) fz = calcFraction<Real>(ph, phiObs[q - d.Z], fracThreshold);
			if (!(ph < 0.)) {
				if (i <= w + 1 && openKind(flags[q - d.X])) one = true;
				if (j <= w + 1 && openKind(flags[q - d.Y])) one = true;
				if (d.is3D && k <= w + 1 && openKind(flags[q - d.Z])) one = true;
			}
		} else if (openKind(flags[q])) {
			// the inner neighbour p = q - X / Y / Z of an outer-layer cell, if it is an interior cell outside the obstacle and close enough to its max wall
			if (i == d.sx - 1 && interiorCell(d, i - 1, j, k) && !(phiObs[q - d.X] < 0.) && i - 1 >= d.sx - w - 2) one = true;
			if (j == d.sy - 1 && interiorCell(d, i, j - 1, k) && !(phiObs[q - d.Y] < 0.) && j - 1 >= d.sy - w - 2) one = true;
			if (d.is3D && k == d.sz - 1 && interiorCell(d, i, j, k - 1) && !(phiObs[q - d.Z] < 0.) && j >= d.sz - w - 2) one = true;
		}
		if (one) { fx = 1; fy = 1; if (d.is3D) fz = 1; }
		fractions[3 * q] = fx; fractions[3 * q + 1] = fy; fractions[3 * q + 2] = fz;
	}
};
template <typename Real> struct SetObstacleFlags {       // KnUpdateFlagsObs, KERNEL(bnd = boundaryWidth >= 1)
	static const bool kSplit = false;
	int* flags; const Real* phiObs; const Real* fractions; const Real* phiOut; const Real* phiIn; int bw;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt p) const {
		if (i < bw || i >= d.sx - bw || j < bw || j >= d.sy - bw || (d.is3D && (k < bw || k >= d.sz - bw))) return;
		bool isObs = false;
		if (fractions) {
			Real f = 0;
			f += fractions[3 * p]; f += fractions[3 * (p + d.X)];
			f += fractions[3 * p + 1]; f += fractions[3 * (p + d.Y) + 1];
			if (d.is3D) { f += fractions[3 * p + 2]; f += fractions[3 * (p + d.Z) + 2]; }
			if (f == 0.) isObs = true;
		} else if (phiObs[p] < 0.) isObs = true;
		const bool isOutflow = phiOut && phiOut[p] < 0., isInflow = phiIn && phiIn[p] < 0.;
		flags[p] = isObs ? TypeObstacle : (isInflow ? (TypeFluid | TypeInflow) : (isOutflow ? (TypeEmpty | TypeOutflow) : TypeEmpty));
	}
};

// ---------------------------------------------------------------- getLaplacian / getCurvature (surface tension helpers of solvePressure's `curv` argument)
// LaplaceOp commonkernels.h:75-80, CurvatureOp :83-101, wrapped by plugin/flip.cpp:710-716.  The reference's double literals promote every
// product to double and every named Real narrows; the same evaluation here.  Cells of the outer layer are left alone (KERNEL(bnd=1)).
template <typename Real> struct LaplaceCell {
	static const bool kSplit = false;
	Real* laplace; const Real* grid;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt p) const {
		if (!interiorCell(d, i, j, k)) return;
		Real l = (Real)(((double)grid[p + d.X] - 2.0 * (double)grid[p]) + (double)grid[p - d.X]);
		l = (Real)((double)l + (((double)grid[p + d.Y] - 2.0 * (double)grid[p]) + (double)grid[p - d.Y]));
		if (d.is3D) l = (Real)((double)l + (((double)grid[p + d.Z] - 2.0 * (double)grid[p]) + (double)grid[p - d.Z]));
		laplace[p] = l;
	}
};
template <typename Real> struct CurvatureCell {
	static const bool kSplit = false;
	Real* curv; const Real* grid; Real h;
	MP_HD void operator()(const Dims& d, int i, int j, int k, IndexInt p) const {
		if (!interiorCell(d, i, j, k)) return;
		const IndexInt X = d.X, Y = d.Y, Z = d.Z;
		const Real over_h = (Real)(1.0 / (double)h);
		const double oh = (double)over_h, g2 = 2.0 * (double)grid[p];
		const Real x = (Real)((0.5 * (double)(grid[p + X] - grid[p - X])) * oh);
		const Real y = (Real)((0.5 * (double)(grid[p + Y] - grid[p - Y])) * oh);
		const Real xx = (Real)(((((double)grid[p + X] - g2) + (double)grid[p - X]) * oh) * oh);
		const Real yy = (Real)(((((double)grid[p + Y] - g2) + (double)grid[p - Y]) * oh) * oh);
		const Real xy = (Real)(((0.25 * (double)(((grid[p + X + Y] + grid[p - X - Y]) - grid[p - X + Y]) - grid[p + X - Y])) * oh) * oh);
		Real c = (Real)(((double)(x * x * yy + y * y * xx)) - (2.0 * (double)x) * (double)y * (double)xy);
		Real denom = x * x + y * y;
		if (d.is3D) {
			const Real z = (Real)((0.5 * (double)(grid[p + Z] - grid[p - Z])) * oh);
			const Real zz = (Real)(((((double)grid[p + Z] - g2) + (double)grid[p - Z]) * oh) * oh);
			const Real xz = (Real)(((0.25 * (double)(((grid[p + X + Z] + grid[p - X - Z]) - grid[p - X + Z]) - grid[p + X - Z])) * oh) * oh);
			const Real yz = (Real)(((0.25 * (double)(((grid[p + Y + Z] + grid[p - Y - Z]) - grid[p + Y - Z]) - grid[p - Y + Z])) * oh) * oh);
			c = (Real)((double)c + ((double)(x * x * zz + z * z * xx + y * y * zz + z * z * yy) - 2.0 * (double)(x * z * xz + y * z * yz)));
			denom += z * z;
		}
		const Real eps = sizeof(Real) == 4 ? (Real)1e-6f : (Real)1e-10;           // VECTOR_EPSILON vectorbase.h:52,:55
		const Real dmax = denom > eps ? denom : eps;
		curv[p] = (Real)((double)c / pow((double)dmax, 1.5));                     // the one operation that is not correctly rounded on either side: last-bit differences possible
	}
};

}  // namespace liquid

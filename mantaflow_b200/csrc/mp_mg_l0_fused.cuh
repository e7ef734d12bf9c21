// Level 0 of the GridMg V-cycle, fused (GridMg::doVCycle multigrid.cpp:448-504 with V(1,1), the way ApplyPreconditionMultigrid
// conjugategrad.cpp:162-167 runs it):
//   DOWN    knSet(x, 0) :458 + smoothGS colours 0, 1 (:713-737, knSmoothColor :668-711) + knCalcResidual (:739-771)   -> x, r
//   SMOOTH  smoothGS in either colour order on a given iterate (the post-smoothing runs colours 1, 0 :733)            -> x
// as ONE pass over the level instead of three (DOWN) / two (SMOOTH), with the level-0 operator as 2 bytes per cell.
//
// Why one pass is possible.  A colour sweep only reads the other colour, so the three stages of DOWN form a dependency cone of radius
// 2: r(c) needs x on the 6 neighbours of c, the second-colour x needs first-colour x on its 6 neighbours, and first-colour x on a zero
// iterate is pointwise (b / A0).  A CTA owns an x-y tile T and marches along z.  Per plane it stages b, the operator mask and the
// first-stage iterate on T + 2 cells (x halo widened to a 16-byte vector), completes the second colour on T + 1 one plane behind, and
// forms the residual (SMOOTH: the second colour of the sweep) on T two planes behind -- the halo cells are recomputed by the neighbouring
// CTAs, value for value the same, which costs arithmetic but no extra DRAM traffic (the halo loads hit L2).
//
// The operator mask (k_mg_build_mask0): without face fractions every off-diagonal of the level-0 matrix is exactly 0 or -1 and the
// diagonal a small integer (MakeLaplaceMatrix conjugategrad.h:154-187), so a row is bit 0 = active vertex, bits 1..6 = coupling to
// -x,+x,-y,+y,-z,+z is -1 (otherwise +0, also towards vertices outside the grid), bits 7..9 = the diagonal 1..6, or 7 = read it from
// A (ghost-fluid diagonals, and trivial rows whose diagonal was scaled by 1e-6 :376-381), bit 10 = trivial row (b is scaled :417-424).
// Any other off-diagonal value makes the mask invalid and the unfused kernels run.
//
// Same arithmetic term for term as k_mg_smooth0 / k_mg_residual0: sum = b; sum -= A_nb * x_nb in the order -x,+x,-y,+y,-z,+z (then
// -= A0 * x for the residual); x = sum / A0.  Where those kernels skip a neighbour outside the grid, this one subtracts (+0) * (+0):
// v - (+0) == v for every v, zeros of either sign included.
//
// The phases are plain functions of (thread id, staged planes) so that tests/emul/mg_l0_emul.cpp can walk them on the host (the build
// container has no GPU); the __global__ wrapper in mp_mg.cu calls them with __syncthreads() in between.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>

#ifdef __CUDACC__
#define MGF_HD __host__ __device__ __forceinline__
#else
#define MGF_HD inline
#endif

namespace mgl0 {

enum : unsigned { mActive = 1u, mTrivial = 1u << 10 };
enum : int { MODE_DOWN = 0, MODE_SMOOTH = 1 };

struct Geom { int sx, sy, sz; };

template <typename T, int N> struct alignas(sizeof(T) * N) Vec { T v[N]; };
template <typename T, int N> MGF_HD Vec<T, N> ldVec(const T* p) {
#ifdef __CUDA_ARCH__
	return *reinterpret_cast<const Vec<T, N>*>(p);
#else
	Vec<T, N> r; memcpy(r.v, p, sizeof(T) * N); return r;
#endif
}
template <typename T, int N> MGF_HD void stVec(T* p, const Vec<T, N>& r) {
#ifdef __CUDA_ARCH__
	*reinterpret_cast<Vec<T, N>*>(p) = r;
#else
	memcpy(p, r.v, sizeof(T) * N);
#endif
}

template <typename Real> struct Tile {
	static constexpr int V = 16 / (int)sizeof(Real);        // cells per 16-byte vector
	static constexpr int TX = 32 * V, TY = 16;               // 128 x 16 (float), 64 x 16 (double)
	static constexpr int HX = V, HY = 2;                     // staged halo: x widened to a whole vector
	static constexpr int W2 = TX + 2 * HX, H2 = TY + 2 * HY;
	static constexpr int PLANE = W2 * H2;
	static constexpr int VROW = W2 / V;                      // vectors per staged row
	static constexpr int NVEC = VROW * H2;
	static constexpr int NTHR = 512;
	static constexpr int NSLOT = (NVEC + NTHR - 1) / NTHR;
	static constexpr int XR = 4, BR = 3;                     // ring depths: iterate planes s-1 .. s+2, rhs / mask planes s .. s+2
	static constexpr int W1 = TX + 2, H1 = TY + 2;           // T + 1
	static constexpr int HALF1 = (W1 + 1) / 2;               // cells of one colour in a row of T + 1 (at most)
	static_assert(NTHR == (TX / V) * TY, "the last phase gives every thread one vector of T");
};

template <typename Real> struct Smem {
	Real X[Tile<Real>::XR][Tile<Real>::PLANE];
	Real B[Tile<Real>::BR][Tile<Real>::PLANE];
	unsigned short M[Tile<Real>::BR][Tile<Real>::PLANE];
};
template <typename Real> struct Pre {                       // one plane's share of a thread, in registers between issue() and stage()
	Vec<Real, Tile<Real>::V> b[Tile<Real>::NSLOT], x[Tile<Real>::NSLOT];
	Vec<unsigned short, Tile<Real>::V> m[Tile<Real>::NSLOT];
};

MGF_HD int xr(int q) { return (q + 4) & 3; }                // q >= -2
MGF_HD int br(int q) { return (q + 3) % 3; }
template <typename Real> MGF_HD Real coef(unsigned m, int dir) { return ((m >> (1 + dir)) & 1u) ? (Real)-1 : (Real)0; }
template <typename Real> MGF_HD Real diag(unsigned m, const Real* A0, size_t v) { const unsigned c = (m >> 7) & 7u; return c < 7u ? (Real)(int)c : A0[v]; }

// global -> registers: plane q of b, the mask and (SMOOTH) the iterate over T + 2; everything outside the grid reads as zero
template <typename Real, int MODE>
MGF_HD void issue(const Geom& g, int x0, int y0, int q, int tid, const Real* __restrict__ b, const Real* __restrict__ xin, const unsigned short* __restrict__ mask, Pre<Real>& p)
{
	typedef Tile<Real> T;
	#pragma unroll
	for (int sl = 0; sl < T::NSLOT; sl++) {
		#pragma unroll
		for (int e = 0; e < T::V; e++) { p.b[sl].v[e] = (Real)0; p.x[sl].v[e] = (Real)0; p.m[sl].v[e] = 0; }
		const int vec = tid + sl * T::NTHR;
		if (vec >= T::NVEC) continue;
		const int row = vec / T::VROW, vi = vec - row * T::VROW;
		const int gy = y0 - T::HY + row, gx = x0 - T::HX + vi * T::V;
		if (q < 0 || q >= g.sz || gy < 0 || gy >= g.sy || gx < 0 || gx >= g.sx) continue;      // sx % V == 0: a vector is inside or outside as a whole
		const size_t v = (size_t)gx + (size_t)g.sx * ((size_t)gy + (size_t)g.sy * (size_t)q);
		p.m[sl] = ldVec<unsigned short, T::V>(mask + v);
		p.b[sl] = ldVec<Real, T::V>(b + v);
		if (MODE == MODE_SMOOTH) p.x[sl] = ldVec<Real, T::V>(xin + v);
	}
}

// registers -> staged plane q.  DOWN: the iterate after the first colour `c0` of a sweep over x == 0 (b / A0 on that colour, zero elsewhere)
template <typename Real, int MODE>
MGF_HD void stage(const Geom& g, int x0, int y0, int q, int tid, Real bscale, const Real* __restrict__ A0, int c0, const Pre<Real>& p, Smem<Real>& s)
{
	typedef Tile<Real> T;
	#pragma unroll
	for (int sl = 0; sl < T::NSLOT; sl++) {
		const int vec = tid + sl * T::NTHR;
		if (vec >= T::NVEC) continue;
		const int row = vec / T::VROW, vi = vec - row * T::VROW;
		const int gy = y0 - T::HY + row, gx = x0 - T::HX + vi * T::V;
		const int o = row * T::W2 + vi * T::V;
		Vec<Real, T::V> bv, xv; Vec<unsigned short, T::V> mv = p.m[sl];
		#pragma unroll
		for (int e = 0; e < T::V; e++) {
			const unsigned m = mv.v[e];
			Real bb = p.b[sl].v[e];
			if (bscale != (Real)0 && (m & mTrivial)) bb *= bscale;
			bv.v[e] = bb;
			if (MODE == MODE_SMOOTH) xv.v[e] = p.x[sl].v[e];
			else {
				Real xx = (Real)0;
				if ((m & mActive) && ((gx + e + gy + q + c0) & 1) == 0)
					xx = bb / diag<Real>(m, A0, (size_t)(gx + e) + (size_t)g.sx * ((size_t)gy + (size_t)g.sy * (size_t)q));
				xv.v[e] = xx;
			}
		}
		stVec<Real, T::V>(&s.X[xr(q)][o], xv);
		stVec<Real, T::V>(&s.B[br(q)][o], bv);
		stVec<unsigned short, T::V>(&s.M[br(q)][o], mv);
	}
}

template <typename Real>
MGF_HD Real rowSum(const Smem<Real>& s, int p, int o, unsigned m) {      // b - sum of off-diagonal terms, in the reference's order
	typedef Tile<Real> T;
	const Real* X = s.X[xr(p)];
	Real sum = s.B[br(p)][o];
	sum -= coef<Real>(m, 0) * X[o - 1];
	sum -= coef<Real>(m, 1) * X[o + 1];
	sum -= coef<Real>(m, 2) * X[o - T::W2];
	sum -= coef<Real>(m, 3) * X[o + T::W2];
	sum -= coef<Real>(m, 4) * s.X[xr(p - 1)][o];
	sum -= coef<Real>(m, 5) * s.X[xr(p + 1)][o];
	return sum;
}

// colour `c` of plane p over T + 1, in place in the staged iterate (reads the other colour only)
template <typename Real>
MGF_HD void mid(const Geom& g, int x0, int y0, int p, int tid, int c, const Real* __restrict__ A0, Smem<Real>& s)
{
	typedef Tile<Real> T;
	if (p < 0 || p >= g.sz) return;
	for (int e = tid; e < T::H1 * T::HALF1; e += T::NTHR) {
		const int row = e / T::HALF1, mi = e - row * T::HALF1;
		const int gy = y0 - 1 + row;
		const int first = (x0 - 1 + gy + p + c) & 1;             // the first cell of this colour in the row of T + 1 (x0 - 1 may be -1: & 1 of a negative int is still its parity)
		const int xl = -1 + first + 2 * mi;                      // relative to the tile
		const int gx = x0 + xl;
		if (xl > T::TX || gy < 0 || gy >= g.sy || gx < 0 || gx >= g.sx) continue;
		const int o = (row + T::HY - 1) * T::W2 + (xl + T::HX);
		const unsigned m = s.M[br(p)][o];
		if (!(m & mActive)) continue;
		const Real sum = rowSum<Real>(s, p, o, m);
		s.X[xr(p)][o] = sum / diag<Real>(m, A0, (size_t)gx + (size_t)g.sx * ((size_t)gy + (size_t)g.sy * (size_t)p));
	}
}

// plane p over T, one vector per thread.  DOWN: r = b - A x on every active vertex (0 elsewhere), x and r to global memory.
// SMOOTH: colour `c` of the sweep, x to global memory.
template <typename Real, int MODE>
MGF_HD void last(const Geom& g, int x0, int y0, int p, int tid, int c, const Real* __restrict__ A0, const Smem<Real>& s, Real* __restrict__ xout, Real* __restrict__ rout)
{
	typedef Tile<Real> T;
	const int row = tid / (T::TX / T::V), vi = tid - row * (T::TX / T::V);
	const int gy = y0 + row, gx = x0 + vi * T::V;
	if (gy >= g.sy || gx >= g.sx) return;
	const int o0 = (row + T::HY) * T::W2 + T::HX + vi * T::V;
	const size_t v0 = (size_t)gx + (size_t)g.sx * ((size_t)gy + (size_t)g.sy * (size_t)p);
	const Vec<unsigned short, T::V> mv = ldVec<unsigned short, T::V>(&s.M[br(p)][o0]);
	Vec<Real, T::V> xv = ldVec<Real, T::V>(&s.X[xr(p)][o0]), rv;
	#pragma unroll
	for (int e = 0; e < T::V; e++) {
		const unsigned m = mv.v[e];
		rv.v[e] = (Real)0;
		if (!(m & mActive)) continue;
		if (MODE == MODE_DOWN) {
			Real sum = rowSum<Real>(s, p, o0 + e, m);
			sum -= diag<Real>(m, A0, v0 + e) * xv.v[e];
			rv.v[e] = sum;
		} else if (((gx + e + gy + p + c) & 1) == 0) {
			xv.v[e] = rowSum<Real>(s, p, o0 + e, m) / diag<Real>(m, A0, v0 + e);
		}
	}
	stVec<Real, T::V>(xout + v0, xv);
	if (MODE == MODE_DOWN) stVec<Real, T::V>(rout + v0, rv);
}

// The march of one CTA over planes [k0, k1) of tile (x0, y0), written once for the device (SYNC = __syncthreads, one thread per call)
// and once for the host emulation (which walks tid itself).  See mp_mg.cu k_mg_l0_fused / tests/emul/mg_l0_emul.cpp.
//   prologue   stage k0-2 .. k0+1;  mid k0-1, k0
//   step s     stage s+2 (and issue s+3) | mid s+1 | last s
// SMOOTH(cFirst, cSecond): mid runs cFirst, last runs cSecond.  DOWN(cFirst = 0): stage does colour 0 on x == 0, mid colour 1, last the residual.

// operator mask of one vertex from the level-0 struct-of-arrays operator and the vertex types (k_mg_build_mask0); *bad is raised when the
// row cannot be coded
template <typename Real>
MGF_HD unsigned short maskOf(const Geom& g, int is3D, int x, int y, int z, const Real* __restrict__ A, const signed char* __restrict__ type, int* bad)
{
	const size_t n = (size_t)g.sx * g.sy * g.sz, Y = (size_t)g.sx, Z = (size_t)g.sx * g.sy;
	const size_t v = (size_t)x + Y * y + Z * z;
	const signed char t = type[v];
	if (t == 0) return 0;                                        // vtInactive
	unsigned m = mActive;
	if (t == 2) m |= mTrivial;                                   // vtActiveTrivial
	Real c[6] = { (Real)0, (Real)0, (Real)0, (Real)0, (Real)0, (Real)0 };
	if (x > 0)        c[0] = A[n + v - 1];
	if (x < g.sx - 1) c[1] = A[n + v];
	if (y > 0)        c[2] = A[2 * n + v - Y];
	if (y < g.sy - 1) c[3] = A[2 * n + v];
	if (is3D) {
		if (z > 0)        c[4] = A[3 * n + v - Z];
		if (z < g.sz - 1) c[5] = A[3 * n + v];
	}
	for (int q = 0; q < 6; q++) {
		if (c[q] == (Real)-1) m |= 2u << q;
		else if (!(c[q] == (Real)0) || std::signbit(c[q])) *bad = 1;      // anything but -1 and +0 (face fractions): not codable
	}
	unsigned code = 7;
	const Real a0 = A[v];
	if (t != 2) for (int q = 1; q < 7; q++) if (a0 == (Real)q) code = (unsigned)q;
	return (unsigned short)(m | (code << 7));
}

}  // namespace mgl0

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

// What a thread works on, fixed for the whole march (only the plane changes): its staging vectors of T + 2, its vectors of the
// second-stage region T + 1 (widened to whole vectors; `valid` masks the cells outside T + 1 or outside the grid) and its vector of T.
// Offsets into the staged planes / into a global plane; x0, the halo widths and the vector starts are even, so the colour of cell e of a
// vector at plane p is (par + p + e) & 1 with par = (row's y + colour) & 1.
template <typename Real> struct Ctx {
	static constexpr int NMID = (Tile<Real>::VROW * Tile<Real>::H1 + Tile<Real>::NTHR - 1) / Tile<Real>::NTHR;
	int so[Tile<Real>::NSLOT], sg[Tile<Real>::NSLOT];      // so < 0: no vector; sg < 0: outside the grid in x or y (an x-y plane has < 2^31 cells)
	int mo[NMID], mg[NMID];
	int lo, lg;                                            // lg < 0: outside the grid
	unsigned bits;                                         // y parities: bit sl (staging), bit 4 + k (second stage), bit 8 (T); valid cells of second-stage vector k: bits 12 + 8k ..
	MGF_HD int sy(int sl) const { return (int)((bits >> sl) & 1u); }
	MGF_HD int my(int k) const { return (int)((bits >> (4 + k)) & 1u); }
	MGF_HD int ly() const { return (int)((bits >> 8) & 1u); }
	MGF_HD unsigned mvalid(int k) const { return (bits >> (12 + 8 * k)) & 0xffu; }
};
template <typename Real>
MGF_HD Ctx<Real> makeCtx(const Geom& g, int x0, int y0, int tid)
{
	typedef Tile<Real> T;
	static_assert(T::NSLOT <= 4 && Ctx<Real>::NMID <= 2 && T::V <= 8, "packing of Ctx::bits");
	Ctx<Real> c; c.bits = 0;
	#pragma unroll
	for (int sl = 0; sl < T::NSLOT; sl++) {
		const int vec = tid + sl * T::NTHR;
		c.so[sl] = -1; c.sg[sl] = -1;
		if (vec >= T::NVEC) continue;
		const int row = vec / T::VROW, vi = vec - row * T::VROW;
		const int gy = y0 - T::HY + row, gx = x0 - T::HX + vi * T::V;
		c.so[sl] = row * T::W2 + vi * T::V; c.bits |= (unsigned)(gy & 1) << sl;
		if (gy >= 0 && gy < g.sy && gx >= 0 && gx < g.sx) c.sg[sl] = gx + g.sx * gy;      // sx % V == 0: a vector is inside or outside as a whole
	}
	#pragma unroll
	for (int k = 0; k < Ctx<Real>::NMID; k++) {
		const int it = tid + k * T::NTHR;
		c.mo[k] = -1; c.mg[k] = -1;
		if (it >= T::VROW * T::H1) continue;
		const int row = it / T::VROW, vi = it - row * T::VROW;       // rows -1 .. TY of the tile
		const int gy = y0 - 1 + row, gx = x0 - T::HX + vi * T::V;
		c.mo[k] = (row + T::HY - 1) * T::W2 + vi * T::V; c.bits |= (unsigned)(gy & 1) << (4 + k);
		if (gy < 0 || gy >= g.sy || gx < 0 || gx >= g.sx) continue;
		c.mg[k] = gx + g.sx * gy;
		for (int e = 0; e < T::V; e++) { const int xl = gx + e - x0; if (xl >= -1 && xl <= T::TX) c.bits |= 1u << (12 + 8 * k + e); }
	}
	{
		const int row = tid / (T::TX / T::V), vi = tid - row * (T::TX / T::V);
		const int gy = y0 + row, gx = x0 + vi * T::V;
		c.lo = (row + T::HY) * T::W2 + T::HX + vi * T::V; c.bits |= (unsigned)(gy & 1) << 8;
		c.lg = (gy < g.sy && gx < g.sx) ? gx + g.sx * gy : -1;
	}
	return c;
}

MGF_HD int xr(int q) { return (q + 4) & 3; }                // q >= -2
MGF_HD int br(int q) { return (q + 3) % 3; }
template <typename Real> MGF_HD Real diag(unsigned m, const Real* A0, size_t v) { const unsigned c = (m >> 7) & 7u; return c < 7u ? (Real)(int)c : A0[v]; }

// global -> registers: plane q of b, the mask and (SMOOTH) the iterate over T + 2; everything outside the grid reads as zero
template <typename Real, int MODE>
MGF_HD void issue(const Geom& g, const Ctx<Real>& c, int q, const Real* __restrict__ b, const Real* __restrict__ xin, const unsigned short* __restrict__ mask, Pre<Real>& p)
{
	typedef Tile<Real> T;
	const bool zin = q >= 0 && q < g.sz;
	const size_t pl = (size_t)g.sx * g.sy * (size_t)(zin ? q : 0);
	#pragma unroll
	for (int sl = 0; sl < T::NSLOT; sl++) {
		#pragma unroll
		for (int e = 0; e < T::V; e++) { p.b[sl].v[e] = (Real)0; p.x[sl].v[e] = (Real)0; p.m[sl].v[e] = 0; }
		if (!zin || c.sg[sl] < 0) continue;
		const size_t v = pl + (size_t)c.sg[sl];
		p.m[sl] = ldVec<unsigned short, T::V>(mask + v);
		p.b[sl] = ldVec<Real, T::V>(b + v);
		if (MODE == MODE_SMOOTH) p.x[sl] = ldVec<Real, T::V>(xin + v);
	}
}

// registers -> staged plane q.  DOWN: the iterate after the first colour `c0` of a sweep over x == 0 (b / A0 on that colour, zero elsewhere)
template <typename Real, int MODE>
MGF_HD void stage(const Geom& g, const Ctx<Real>& c, int q, Real bscale, const Real* __restrict__ A0, int c0, const Pre<Real>& p, Smem<Real>& s)
{
	typedef Tile<Real> T;
	const int xs = xr(q), bs = br(q);
	#pragma unroll
	for (int sl = 0; sl < T::NSLOT; sl++) {
		if (c.so[sl] < 0) continue;
		const Vec<unsigned short, T::V> mv = p.m[sl];
		Vec<Real, T::V> bv = p.b[sl], xv;
		unsigned any = 0;
		#pragma unroll
		for (int e = 0; e < T::V; e++) any |= mv.v[e];
		if (bscale != (Real)0 && (any & mTrivial)) {
			#pragma unroll
			for (int e = 0; e < T::V; e++) if (mv.v[e] & mTrivial) bv.v[e] *= bscale;
		}
		if (MODE == MODE_SMOOTH) xv = p.x[sl];
		else {
			#pragma unroll
			for (int e = 0; e < T::V; e++) xv.v[e] = (Real)0;
			if (any & mActive) {
				const size_t v = (size_t)g.sx * g.sy * (size_t)q + (size_t)c.sg[sl];
				if ((c.sy(sl) + q + c0) & 1) {
					#pragma unroll
					for (int e = 1; e < T::V; e += 2) if (mv.v[e] & mActive) xv.v[e] = bv.v[e] / diag<Real>(mv.v[e], A0, v + e);
				} else {
					#pragma unroll
					for (int e = 0; e < T::V; e += 2) if (mv.v[e] & mActive) xv.v[e] = bv.v[e] / diag<Real>(mv.v[e], A0, v + e);
				}
			}
		}
		stVec<Real, T::V>(&s.X[xs][c.so[sl]], xv);
		stVec<Real, T::V>(&s.B[bs][c.so[sl]], bv);
		stVec<unsigned short, T::V>(&s.M[bs][c.so[sl]], mv);
	}
}

// the staged neighbourhood of one vector of plane p, and b - sum of the off-diagonal terms of its cell e in the reference's order.
// A coupling is -1 or +0: sum -= (-1) * x is sum + x, and sum -= (+0) * x leaves every sum but an exact -0 untouched (signs of exact
// zeros aside, the value of the per-colour kernels).
template <typename Real> struct Hood {
	Vec<Real, Tile<Real>::V> b, xc, ym, yp, zm, zp; Real xl, xh;
	Vec<unsigned short, Tile<Real>::V> m;
};
template <typename Real>
MGF_HD void loadHood(const Smem<Real>& s, int p, int o, Hood<Real>& h)
{
	typedef Tile<Real> T;
	const Real* X = s.X[xr(p)];
	h.b = ldVec<Real, T::V>(&s.B[br(p)][o]);
	h.xc = ldVec<Real, T::V>(X + o); h.xl = X[o - 1]; h.xh = X[o + T::V];
	h.ym = ldVec<Real, T::V>(X + o - T::W2); h.yp = ldVec<Real, T::V>(X + o + T::W2);
	h.zm = ldVec<Real, T::V>(&s.X[xr(p - 1)][o]); h.zp = ldVec<Real, T::V>(&s.X[xr(p + 1)][o]);
}
template <typename Real, int E>
MGF_HD Real rowSum(const Hood<Real>& h)
{
	constexpr int V = Tile<Real>::V;
	const unsigned m = h.m.v[E];
	Real sum = h.b.v[E];
	if (m & 2u)  sum += (E == 0 ? h.xl : h.xc.v[E == 0 ? 0 : E - 1]);
	if (m & 4u)  sum += (E == V - 1 ? h.xh : h.xc.v[E == V - 1 ? E : E + 1]);
	if (m & 8u)  sum += h.ym.v[E];
	if (m & 16u) sum += h.yp.v[E];
	if (m & 32u) sum += h.zm.v[E];
	if (m & 64u) sum += h.zp.v[E];
	return sum;
}
template <typename Real, int E0>
MGF_HD void relaxCells(const Hood<Real>& h, unsigned act, const Real* __restrict__ A0, size_t v, Vec<Real, Tile<Real>::V>& x)      // cells E0, E0 + 2, ... of the vector
{
	constexpr int V = Tile<Real>::V;
	if (act & (1u << E0)) x.v[E0] = rowSum<Real, E0>(h) / diag<Real>(h.m.v[E0], A0, v + E0);
	if (V > 2) { constexpr int E2 = V > 2 ? E0 + 2 : E0; if (act & (1u << E2)) x.v[E2] = rowSum<Real, E2>(h) / diag<Real>(h.m.v[E2], A0, v + E2); }
}
template <typename Real>
MGF_HD unsigned activeBits(const Vec<unsigned short, Tile<Real>::V>& m) {
	unsigned a = 0;
	#pragma unroll
	for (int e = 0; e < Tile<Real>::V; e++) a |= (unsigned)(m.v[e] & mActive) << e;
	return a;
}

// colour `col` of plane p over T + 1, in place in the staged iterate (reads the other colour only)
template <typename Real>
MGF_HD void mid(const Geom& g, const Ctx<Real>& c, int p, int col, const Real* __restrict__ A0, Smem<Real>& s)
{
	typedef Tile<Real> T;
	if (p < 0 || p >= g.sz) return;
	#pragma unroll
	for (int k = 0; k < Ctx<Real>::NMID; k++) {
		if (c.mo[k] < 0 || !c.mvalid(k)) continue;
		Hood<Real> h;
		h.m = ldVec<unsigned short, T::V>(&s.M[br(p)][c.mo[k]]);
		const int par = (c.my(k) + p + col) & 1;                    // cells e with (e & 1) == par carry this colour
		const unsigned act = activeBits<Real>(h.m) & c.mvalid(k) & (par ? 0xAAAAAAAAu : 0x55555555u);
		if (!act) continue;
		loadHood<Real>(s, p, c.mo[k], h);
		const size_t v = (size_t)g.sx * g.sy * (size_t)p + (size_t)c.mg[k];
		Vec<Real, T::V> x = h.xc;
		if (par) relaxCells<Real, 1>(h, act, A0, v, x); else relaxCells<Real, 0>(h, act, A0, v, x);
		// the cells of the other colour are written back unchanged (no thread writes them in this phase)
		stVec<Real, T::V>(&s.X[xr(p)][c.mo[k]], x);
	}
}

// plane p over T, one vector per thread.  DOWN: r = b - A x on every active vertex (0 elsewhere), x and r to global memory.
// SMOOTH: colour `col` of the sweep, x to global memory.
template <typename Real, int MODE, int E>
MGF_HD void residCell(const Hood<Real>& h, unsigned act, const Real* __restrict__ A0, size_t v, Vec<Real, Tile<Real>::V>& r)
{
	if (act & (1u << E)) { Real sum = rowSum<Real, E>(h); sum -= diag<Real>(h.m.v[E], A0, v + E) * h.xc.v[E]; r.v[E] = sum; }
}
template <typename Real, int MODE>
MGF_HD void last(const Geom& g, const Ctx<Real>& c, int p, int col, const Real* __restrict__ A0, const Smem<Real>& s, Real* __restrict__ xout, Real* __restrict__ rout)
{
	typedef Tile<Real> T;
	if (c.lg < 0) return;
	Hood<Real> h;
	h.m = ldVec<unsigned short, T::V>(&s.M[br(p)][c.lo]);
	const unsigned actAll = activeBits<Real>(h.m);
	const size_t v = (size_t)g.sx * g.sy * (size_t)p + (size_t)c.lg;
	Vec<Real, T::V> x, r;
	#pragma unroll
	for (int e = 0; e < T::V; e++) r.v[e] = (Real)0;
	if (!actAll) {
		x = ldVec<Real, T::V>(&s.X[xr(p)][c.lo]);
	} else {
		loadHood<Real>(s, p, c.lo, h);
		x = h.xc;
		if (MODE == MODE_DOWN) {
			residCell<Real, MODE, 0>(h, actAll, A0, v, r); residCell<Real, MODE, 1>(h, actAll, A0, v, r);
			if (T::V > 2) { residCell<Real, MODE, (T::V > 2 ? 2 : 0)>(h, actAll, A0, v, r); residCell<Real, MODE, (T::V > 2 ? 3 : 1)>(h, actAll, A0, v, r); }
		} else {
			const int par = (c.ly() + p + col) & 1;
			const unsigned act = actAll & (par ? 0xAAAAAAAAu : 0x55555555u);
			if (par) relaxCells<Real, 1>(h, act, A0, v, x); else relaxCells<Real, 0>(h, act, A0, v, x);
		}
	}
	stVec<Real, T::V>(xout + v, x);
	if (MODE == MODE_DOWN) stVec<Real, T::V>(rout + v, r);
}

// The march of one CTA over planes [k0, k1) of tile (x0, y0), written once for the device (k_mg_l0_fused in mp_mg.cu: one thread per
// call, __syncthreads() between the phases) and once for the host emulation (tests/emul/mg_l0_emul.cpp walks the threads itself):
//   prologue   stage k0-2 .. k0+1;  mid k0-1, k0
//   step s     stage s+2 (and issue s+3) | mid s+1 | last s
// SMOOTH(cFirst, cSecond): mid runs cFirst, last runs cSecond.  DOWN(cFirst): stage does colour cFirst on x == 0, mid the other colour,
// last the residual.

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

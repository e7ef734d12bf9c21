// Per-cell arithmetic of the IC(0) preconditioner (PC_ICP), shared by the CUDA kernels of mp_ic.cu and the host emulation the CPU tests
// run (tests/emul/ic_emul.cpp, test infrastructure): InitPreconditionIncompCholesky conjugategrad.cpp:26-63 written as a gather,
// ApplyPreconditionIncompCholesky conjugategrad.cpp:109-132.  A cell may run once the cells of the previous hyperplane i + j + k = c - 1
// (forward) / the next one (backward) are done; cells of one hyperplane are independent.
#pragma once
#include <cmath>
#include "mp_common.cuh"

#ifndef MP_HD
#ifdef __CUDACC__
#define MP_HD __host__ __device__ __forceinline__
#else
#define MP_HD inline
#endif
#endif

namespace ic {

// factor of one cell with i, j, k >= 1.  The reference scatters `A0(i+1,j,k) -= square(Ai[idx])` etc. from every fluid cell to its
// +x/+y/+z neighbours (also onto non-fluid cells); here the cell subtracts the squares of its -z, -y, -x fluid neighbours in the order
// the serial k / j / i loop applies them.  P* already hold a copy of A* (A0.copyFrom(orgA0) ... :31-34).
template <typename Real>
MP_HD void initCell(const int* flags, Real* P0, Real* Pi, Real* Pj, Real* Pk, const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak, IndexInt idx, IndexInt Y, IndexInt Z) {
	const IndexInt ix = idx - 1, iy = idx - Y, iz = idx - Z;
	Real a = A0[idx];
	if (flags[iz] & TypeFluid) { const Real q = Pk[iz]; a -= q * q; }
	if (flags[iy] & TypeFluid) { const Real q = Pj[iy]; a -= q * q; }
	if (flags[ix] & TypeFluid) { const Real q = Pi[ix]; a -= q * q; }
	if (flags[idx] & TypeFluid) {
		const Real dgl = (Real)sqrt(a);                                       // :39
		const Real invDiagonal = 1.0f / dgl;                                  // :45
		Pi[idx] = Ai[idx] * invDiagonal; Pj[idx] = Aj[idx] * invDiagonal; Pk[idx] = Ak[idx] * invDiagonal;
		P0[idx] = dgl > 0 ? (Real)(1.0 / (double)dgl) : dgl;                  // InvertCheckFluid commonkernels.h:25-29
	} else P0[idx] = a;
}
template <typename Real>
MP_HD void fwdCell(const int* flags, Real* dst, const Real* src, const Real* P0, const Real* Pi, const Real* Pj, const Real* Pk, IndexInt idx, IndexInt Y, IndexInt Z) {
	if (!(flags[idx] & TypeFluid)) return;
	const IndexInt ix = idx - 1, iy = idx - Y, iz = idx - Z;
	dst[idx] = P0[idx] * (src[idx] - dst[ix] * Pi[ix] - dst[iy] * Pj[iy] - dst[iz] * Pk[iz]);
}
template <typename Real>
MP_HD void bwdCell(const int* flags, Real* dst, const Real* P0, const Real* Pi, const Real* Pj, const Real* Pk, IndexInt idx, IndexInt Y, IndexInt Z) {
	if (!(flags[idx] & TypeFluid)) return;
	dst[idx] = P0[idx] * (dst[idx] - dst[idx + 1] * Pi[idx] - dst[idx + Y] * Pj[idx] - dst[idx + Z] * Pk[idx]);
}

// hyperplane geometry: cells i in [1, hx], j in [1, hy], k in [1, hz]; plane c = i + j + k runs from 3 to hx + hy + hz.
// init: h = s - 1 (the scatter of a fluid cell reaches the outer layer), sweeps: h = s - 2 (fluid cells are interior cells)
struct Geom { int sx, sy, sz; IndexInt Y, Z; int hx, hy, hz; };
struct PlaneRange { int klo, khi; };
inline bool planeRange(const Geom& g, int c, PlaneRange& r) {
	r.klo = (c - g.hx - g.hy) > 1 ? (c - g.hx - g.hy) : 1;
	r.khi = (c - 2) < g.hz ? (c - 2) : g.hz;
	return r.khi >= r.klo;
}
// cell (j, k) of plane c -> linear index, false if the plane has no cell there
MP_HD bool planeCell(const Geom& g, int c, int j, int k, IndexInt& idx) {
	const int i = c - j - k;
	if (j > g.hy || i < 1 || i > g.hx) return false;
	idx = (IndexInt)i + g.Y * j + g.Z * k;
	return true;
}

}  // namespace ic

// GridMg on the device: Galerkin geometric multigrid of Dick et al. 2015 without topology awareness
// (multigrid.h:31-137, multigrid.cpp).  Used as the PcMGDynamic / PcMGStatic preconditioner of GridCg
// (conjugategrad.cpp:100-106,:162-167) and exposed 1:1 through mp_mg_*.
//
// Storage: per level l  A (struct-of-arrays: A[s*n + v], s < 4 on level 0 / 14 on levels > 0; the reference
// interleaves A[v*stencil+s], multigrid.cpp:208-218 -- mp_mg_download("a") returns the reference layout),
// x, b, r (Real) and the vertex type (int8).  Everything stays in HBM between setA and the V-cycles.
//
// setA (multigrid.cpp:386-415)
//   k_mg_copy_activate      knCopyA + knActivateVertices (+analyzeStencil)                :321-384
//   coarse vertex selection genCoarseGrid :520-578.  The reference runs a serial bucket-heap.  Its first phase
//                           (vertices with exactly one free interpolation vertex select it; min-key-first) is a
//                           monotone closure and therefore order independent: k_mg_select iterates it in parallel
//                           to the fixed point.  If fine vertices with >= 2 free interpolation vertices remain
//                           (isolated features), the order-dependent second phase is needed: that level is then
//                           redone serially on the host (mp_mg_coarsen.h: per-count stacks with lazy deletion give the
//                           reference's visiting order) -- same result as the reference in every case, fast in the common one.
//   k_mg_galerkin1/k_mg_galerkinN  knGenCoarseGridOperator :580-657 (one thread per stored stencil entry, the
//                           reference's accumulation order per entry)
// V-cycle (doVCycle :448-504)
//   k_mg_smooth0 / k_mg_smoothN     knSmoothColor :668-711 (2 colours on level 0, 8 (4 in 2-D) above, :713-737)
//   k_mg_residual0 / k_mg_residualN knCalcResidual :739-771
//   k_mg_restrict                   knRestrict :904-927          k_mg_interp_add  knInterpolate + knAddAssign :934-954,:445
//   k_mg_coarse_cg                  solveCG :796-902 (double Jacobi-PCG, one CTA; block reductions in fixed order)
#include "mp_common.cuh"
#include "mp_mg_coarsen.h"
#include "mp_mg_l0_fused.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

enum : signed char { vtInactive = 0, vtActive = 1, vtActiveTrivial = 2, vtRemoved = 3, vtZero = 4, vtFree = 5 };   // multigrid.h:87-94
#define MG_MAXLVL 32

struct LvlGeom { int sx, sy, sz, n; };

struct CoarseningPath { int Ux, Uy, Uz, Wx, Wy, Wz, Nx, Ny, Nz, sc, sf, inU; float rw, iw; };

template <typename Real> struct MgLevel {
	LvlGeom g;
	Real *A, *x, *b, *r;
	signed char* type;
};

struct mp_mg {
	mp_context* ctx;
	int prec, is3D, dim, stencil, stencil0, nlev;
	LvlGeom geom[MG_MAXLVL];
	void *A[MG_MAXLVL], *x[MG_MAXLVL], *b[MG_MAXLVL], *r[MG_MAXLVL];
	void* Afull[MG_MAXLVL];          // levels > 0: the whole 27-point (9-point) row of every vertex, colour-major (see k_mg_build_full)
	signed char* type[MG_MAXLVL];
	double* cg;                      // 4 * n_coarsest doubles
	CoarseningPath* dPaths; int npaths; int pathStart[15];
	int numPre, numPost; double coarsestAcc, trivialScale;
	bool isASet, isRhsSet;
	int* dFlags;                     // [0] changed, [1] leftovers, [2] nonZeroStencilSum, [3] trivialFound, [4] coarse iterations
	int* hFlags;                     // pinned
	int hostCoarsenLevels;           // how many levels needed the serial host path in the last setA
	// z-slab mode (mp_dist.cu): the hierarchy is the GLOBAL one and lives on every rank; setA gathers the ranks' slabs of A0/Ai/Aj/Ak, the
	// coarse levels (1/7 of the work) are computed redundantly, and the level-0 work of the V-cycle -- smoothing, residual, restriction,
	// interpolation over this rank's planes [k0,k1) -- is sharded, with a one-plane halo exchange of the iterate after each colour.
	// Same arithmetic on the same values as a single-GPU solve of the global grid: same V-cycle, same iteration count.
	bool slab; int k0, k1, lsz;      // owned global planes, local planes incl. the two ghost planes
	// The large coarse levels are sharded too: levels 0 .. nshard-1 are swept over this rank's planes [K0[l], K1[l]) of the (replicated,
	// global-size) level arrays, with an exchange of the two boundary planes of x_l after every colour; coarse plane K belongs to the
	// owner of fine plane 2K.  Level nshard and below are small and computed redundantly on every rank.
	int nshard; int K0[MG_MAXLVL], K1[MG_MAXLVL];
	// the level-0 operator as 2 bytes per vertex for the fused level-0 kernels (mp_mg_l0_fused.cuh); valid when every off-diagonal is 0 / -1
	unsigned short* mask0; bool mask0Valid;
	// Galerkin products of regular vertices (setA): regular[l][v] = 1 when the whole neighbourhood a coarse row is built from is the unperturbed
	// operator, so the row equals that of every other regular vertex of the level; cst = one such row per level (14 entries), first = its vertex
	unsigned char* regular[MG_MAXLVL]; void* cst; int* first;
	// colour-major sweeps: rowreg[l][c * nc + ci] = 1 when the vertex's whole 27-entry row equals cstFull[l] (32 Reals per level), the full row of
	// the level's first regular vertex -- such vertices take their coefficients from cstFull instead of streaming 108 bytes each
	unsigned char* rowreg[MG_MAXLVL]; void* cstFull; bool rowregOn[MG_MAXLVL];
};

// ---------------------------------------------------------------- index helpers
__host__ __device__ __forceinline__ void vecIdx(const LvlGeom& g, int v, int& x, int& y, int& z) { x = v % g.sx; const int t = v / g.sx; y = t % g.sy; z = t / g.sy; }
__host__ __device__ __forceinline__ int linIdx(const LvlGeom& g, int x, int y, int z) { return x + g.sx * (y + g.sy * z); }
__host__ __device__ __forceinline__ bool inGrid(const LvlGeom& g, int x, int y, int z) { return x >= 0 && y >= 0 && z >= 0 && x < g.sx && y < g.sy && z < g.sz; }

// V-cycle kernels are launched on a 3-D grid (x along threadIdx/blockIdx.x, y = blockIdx.y, z = blockIdx.z) so that no thread
// spends its time on integer divisions to recover (x,y,z) from a linear index
__device__ __forceinline__ bool cell3(int nx, int& x, int& y, int& z) { x = blockIdx.x * blockDim.x + threadIdx.x; y = blockIdx.y; z = blockIdx.z; return x < nx; }
static inline dim3 grid3(int nx, int ny, int nz, int block) { return dim3((unsigned)((nx + block - 1) / block), (unsigned)ny, (unsigned)nz); }

// restriction / interpolation weight 1 / 2^(#odd coordinates) (multigrid.cpp:306-307,:623,:646,:921,:951): exact powers of two,
// built without a division
template <typename Real> __device__ __forceinline__ Real pow2weight(int nOdd) {
	Real w = (Real)1;
	if (nOdd >= 1) w = (Real)0.5;
	if (nOdd >= 2) w = (Real)0.25;
	if (nOdd >= 3) w = (Real)0.125;
	return w;
}

// ---------------------------------------------------------------- setA: level 0
template <typename Real>
__global__ void __launch_bounds__(256) k_mg_copy_activate(LvlGeom g, int is3D, Real trivialScale, const Real* A0, const Real* Ai,
	const Real* Aj, const Real* Ak, Real* A, signed char* __restrict__ type, int* flagsOut)      // slab mode runs it in place (A0 == A, Ai == A + n, ...): no restrict
{
	const int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= g.n) return;
	const size_t n = (size_t)g.n;
	Real a[7];
	a[0] = A0[v]; a[1] = Ai[v]; a[2] = Aj[v]; a[3] = is3D ? Ak[v] : (Real)0;
	signed char t = vtInactive;
	Real diag = a[0];
	if (a[0] != (Real)0) {
		t = vtActive;
		int x, y, z; vecIdx(g, v, x, y, z);
		a[4] = x != 0 ? Ai[v - 1] : (Real)0;
		a[5] = y != 0 ? Aj[v - g.sx] : (Real)0;
		a[6] = (z != 0 && is3D) ? Ak[v - g.sx * g.sy] : (Real)0;
		Real smax = 0, ssum = 0;
		#pragma unroll
		for (int i = 0; i < 7; i++) { ssum += a[i]; smax = fmax(smax, fabs(a[i])); }
		if (fabs(ssum / smax) > (Real)1E-6) flagsOut[2] = 1;
		const bool trivial = a[0] == (Real)1 && a[1] == 0 && a[2] == 0 && a[3] == 0 && a[4] == 0 && a[5] == 0 && a[6] == 0;
		if (trivial) { t = vtActiveTrivial; diag = a[0] * trivialScale; flagsOut[3] = 1; }
	}
	type[v] = t;
	A[v] = diag; A[n + v] = a[1]; A[2 * n + v] = a[2];
	if (is3D) A[3 * n + v] = a[3];
}

// ---------------------------------------------------------------- setA: coarse vertex selection (phase-1 closure of genCoarseGrid)
__global__ void __launch_bounds__(256) k_mg_fill_type(signed char* t, int n, signed char v) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) t[i] = v; }

// one sweep: every active fine vertex with exactly one free interpolation vertex selects it.  The first sweep visits every fine vertex and lists
// the undecided ones (>= 2 free interpolation vertices); the following sweeps visit the list of the sweep before only -- a vertex that is not
// listed had at most one free interpolation vertex, took it, and the number of free ones never grows.  flagsOut: [0] changed, [1] undecided
// vertices remain, [6] a list overflowed (the host then falls back to whole sweeps), [8 + parity] list lengths.
__global__ void __launch_bounds__(256) k_mg_select(LvlGeom gf, LvlGeom gc, const signed char* __restrict__ tf, signed char* tc, int* flagsOut,
	const int* __restrict__ listIn, const int* __restrict__ countIn, int* __restrict__ listOut, int* countOut, int cap)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	int v = t;
	if (listIn) { if (t >= *countIn) return; v = listIn[t]; }
	if (v >= gf.n || tf[v] == vtInactive) return;
	int x, y, z; vecIdx(gf, v, x, y, z);
	int nfree = 0, last = -1;
	for (int iz = z / 2; iz <= (z + 1) / 2; iz++) for (int iy = y / 2; iy <= (y + 1) / 2; iy++) for (int ix = x / 2; ix <= (x + 1) / 2; ix++) {
		const int i = linIdx(gc, ix, iy, iz);
		if (tc[i] == vtFree) { nfree++; last = i; }
	}
	if (nfree == 1) { tc[last] = vtZero; flagsOut[0] = 1; }       // changed
	else if (nfree >= 2) {
		flagsOut[1] = 1;                                          // still undecided (final only in a sweep without changes)
		if (listOut) {                                            // one atomic per warp
			const unsigned m = __activemask();
			const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
			int base = 0;
			if (lane == leader) base = atomicAdd(countOut, __popc(m));
			base = __shfl_sync(m, base, leader);
			const int pos = base + __popc(m & ((1u << lane) - 1u));
			if (pos < cap) listOut[pos] = v; else flagsOut[6] = 1;
		}
	}
}
// before the first sweep: an active fine vertex with even coordinates has one interpolation vertex, the coarse vertex at its place, and selects it
// whatever the order -- done first, the first whole sweep finds the neighbours of those vertices decided instead of racing with them (without this
// most of the fine level ended up on the undecided list of sweep 1)
__global__ void __launch_bounds__(128) k_mg_select_even(LvlGeom gf, LvlGeom gc, const signed char* __restrict__ tf, signed char* __restrict__ tc)
{
	int x, y, z;
	if (!cell3(gc.sx, x, y, z)) return;
	if (2 * x < gf.sx && 2 * y < gf.sy && 2 * z < gf.sz && tf[linIdx(gf, 2 * x, 2 * y, 2 * z)] != vtInactive) tc[linIdx(gc, x, y, z)] = vtZero;
}
__global__ void __launch_bounds__(256) k_mg_activate_coarse(signed char* t, int n) {   // knActivateCoarseVertices :507-516
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) t[i] = (t[i] == vtZero) ? vtActive : vtInactive;
}

// levels where the parallel closure leaves work: the serial, order-dependent selection (mp_mg_coarsen.h)

// ---------------------------------------------------------------- setA: Galerkin operators
// level 1 from the 7-point level 0 along the precomputed paths (V)<-R-(U)<-A-(W)<-I-(N); thread = (coarse vertex, sc)
template <typename Real>
__global__ void __launch_bounds__(256) k_mg_galerkin1(LvlGeom gf, LvlGeom gc, int stencil, const CoarseningPath* __restrict__ paths, const int* __restrict__ pathStart,
	const Real* __restrict__ Af, const signed char* __restrict__ tf, const signed char* __restrict__ tc, Real* __restrict__ A)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (long long)gc.n * stencil) return;
	const int sc = (int)(t / gc.n), v = (int)(t % gc.n);
	if (tc[v] == vtInactive) return;
	int vx, vy, vz; vecIdx(gc, v, vx, vy, vz);
	Real acc = 0;
	for (int q = pathStart[sc]; q < pathStart[sc + 1]; q++) {
		const CoarseningPath p = paths[q];
		const int Nx = vx + p.Nx, Ny = vy + p.Ny, Nz = vz + p.Nz;
		if (!inGrid(gc, Nx, Ny, Nz) || tc[linIdx(gc, Nx, Ny, Nz)] == vtInactive) continue;
		const int Ux = vx * 2 + p.Ux, Uy = vy * 2 + p.Uy, Uz = vz * 2 + p.Uz;
		if (!inGrid(gf, Ux, Uy, Uz)) continue;
		const int u = linIdx(gf, Ux, Uy, Uz); if (tf[u] == vtInactive) continue;
		const int Wx = vx * 2 + p.Wx, Wy = vy * 2 + p.Wy, Wz = vz * 2 + p.Wz;
		if (!inGrid(gf, Wx, Wy, Wz)) continue;
		const int w = linIdx(gf, Wx, Wy, Wz); if (tf[w] == vtInactive) continue;
		const Real a = Af[(size_t)p.sf * gf.n + (p.inU ? u : w)];
		acc += (Real)p.rw * a * (Real)p.iw;
	}
	A[(size_t)sc * gc.n + v] = acc;
}

// level 1, restructured: one thread per coarse vertex walks U (27 restriction vertices, z/y/x order) and W (the 7-point
// stencil of U in the reference's order centre,-x,+x,-y,+y,-z,+z) once, loading each level-0 coefficient a single time,
// and scatters rw*a*iw to the accumulators of the <= 8 coarse vertices N that W interpolates from.  For a fixed
// accumulator the contributions arrive in exactly the order of the reference's sorted path list (U-major, then W), so
// the sums are the same bit for bit as k_mg_galerkin1 / multigrid.cpp:594-614 -- with 189 instead of ~700 coefficient loads.
template <typename Real>
__global__ void __launch_bounds__(128) k_mg_galerkin1_v2(LvlGeom gf, LvlGeom gc, int is3D, const Real* __restrict__ Af,
	const signed char* __restrict__ tf, const signed char* __restrict__ tc, Real* __restrict__ A, const unsigned char* __restrict__ skip, const int* __restrict__ only, size_t outStride,
	const int* __restrict__ list, int listCount)
{
	// list: the vertices to compute (the irregular ones).  skip: regular vertices are filled with the level's constant row afterwards.  only: compute the row of vertex *only alone, into A[e * outStride]
	__shared__ Real acc[14][128];
	int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (only) { if (v != 0 || *only >= gc.n) return; v = *only; }
	else if (list) { if (v >= listCount) return; v = list[v]; }
	if (v >= gc.n || tc[v] == vtInactive || (skip && skip[v])) return;
	const size_t vOut = only ? 0 : (size_t)v;
	const int S = is3D ? 14 : 5, t = threadIdx.x;
	for (int e = 0; e < S; e++) acc[e][t] = (Real)0;
	int vx, vy, vz; vecIdx(gc, v, vx, vy, vz);
	// active coarse neighbours N = V + (nx-1, ny-1, nz-1), local code s = nx + 3 ny + 9 nz (>= 13 are the stored entries)
	unsigned int nmask = 0;
	for (int s = 13; s < 27; s++) {
		const int nx = vx + s % 3 - 1, ny = vy + (s / 3) % 3 - 1, nz = vz + s / 9 - 1;
		if ((is3D || s / 9 == 1) && inGrid(gc, nx, ny, nz) && tc[linIdx(gc, nx, ny, nz)] != vtInactive) nmask |= 1u << s;
	}
	const int p7[7][3] = { {0,0,0}, {-1,0,0}, {1,0,0}, {0,-1,0}, {0,1,0}, {0,0,-1}, {0,0,1} };
	const int uz0 = is3D ? 1 : 2, uz1 = is3D ? 3 : 2, nW = is3D ? 7 : 5;
	for (int uz = uz0; uz <= uz1; uz++) for (int uy = 1; uy <= 3; uy++) for (int ux = 1; ux <= 3; ux++) {     // local frame: V sits at (1,1,1), U = 2V + (u-2)
		const int Ux = 2 * vx + ux - 2, Uy = 2 * vy + uy - 2, Uz = 2 * vz + uz - 2;
		if (!inGrid(gf, Ux, Uy, Uz)) continue;
		const int u = linIdx(gf, Ux, Uy, Uz);
		if (tf[u] == vtInactive) continue;
		const Real rw = pow2weight<Real>((ux & 1) + (uy & 1) + (uz & 1));
		for (int i = 0; i < nW; i++) {
			const int wx = ux + p7[i][0], wy = uy + p7[i][1], wz = uz + p7[i][2];
			const int Wx = Ux + p7[i][0], Wy = Uy + p7[i][1], Wz = Uz + p7[i][2];
			if (!inGrid(gf, Wx, Wy, Wz)) continue;
			const int w = linIdx(gf, Wx, Wy, Wz);
			if (tf[w] == vtInactive) continue;
			const int sf = (i + 1) / 2;
			const Real a = Af[(size_t)sf * gf.n + ((i % 2 == 0) ? u : w)];
			const Real iw = pow2weight<Real>((wx & 1) + (wy & 1) + (wz & 1));
			const Real contrib = rw * a * iw;
			for (int nz = wz / 2; nz <= (wz + 1) / 2; nz++) for (int ny = wy / 2; ny <= (wy + 1) / 2; ny++) for (int nx = wx / 2; nx <= (wx + 1) / 2; nx++) {
				const int sN = nx + 3 * ny + 9 * nz;
				if (sN >= 13 && (nmask >> sN & 1u)) acc[sN - 13][t] += contrib;
			}
		}
	}
	if (is3D) { for (int e = 0; e < 14; e++) A[(size_t)e * outStride + vOut] = acc[e][t]; }
	else {
		// 2-D: stored entries sc = s-13 with nz == 1: s in {13,14,15,16,17} -> e = 0..4
		for (int e = 0; e < 5; e++) A[(size_t)e * outStride + vOut] = acc[e][t];
	}
}

// levels > 1 from a 27-point fine level; thread = (coarse vertex, stored entry e = sc-13)
template <typename Real>
__global__ void __launch_bounds__(256) k_mg_galerkinN(LvlGeom gf, LvlGeom gc, int S, int is3D, const Real* __restrict__ Af,
	const signed char* __restrict__ tf, const signed char* __restrict__ tc, Real* __restrict__ A, const unsigned char* __restrict__ skip, const int* __restrict__ only, size_t outStride,
	const int* __restrict__ list, int listCount)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	int e, v;
	if (only) { if (t >= S || *only >= gc.n) return; e = (int)t; v = *only; }
	else if (list) { if (t >= (long long)listCount * S) return; e = (int)(t / listCount); v = list[(int)(t % listCount)]; }
	else { if (t >= (long long)gc.n * S) return; e = (int)(t / gc.n); v = (int)(t % gc.n); }
	if (tc[v] == vtInactive || (skip && skip[v])) return;
	const size_t vOut = only ? 0 : (size_t)v;
	int V[3]; vecIdx(gc, v, V[0], V[1], V[2]);
	const int smaxz = is3D ? 1 : 0;
	// stencil offset of entry e: sc = e + S - 1 ; SC = N - V + smax
	const int sc = e + S - 1;
	int N[3];
	if (is3D) { N[0] = V[0] + sc % 3 - 1; N[1] = V[1] + (sc / 3) % 3 - 1; N[2] = V[2] + sc / 9 - 1; }
	else      { N[0] = V[0] + sc % 3 - 1; N[1] = V[1] + (sc / 3) % 3 - 1; N[2] = V[2]; }
	Real acc = 0;
	if (inGrid(gc, N[0], N[1], N[2]) && tc[linIdx(gc, N[0], N[1], N[2])] != vtInactive) {
		const int fs[3] = { gf.sx, gf.sy, gf.sz }, cs[3] = { gc.sx, gc.sy, gc.sz };
		int u0[3], u1[3];
		for (int d = 0; d < 3; d++) { u0[d] = max(0, V[d] * 2 - 1); u1[d] = min(fs[d] - 1, V[d] * 2 + 1); }
		for (int Uz = u0[2]; Uz <= u1[2]; Uz++) for (int Uy = u0[1]; Uy <= u1[1]; Uy++) for (int Ux = u0[0]; Ux <= u1[0]; Ux++) {
			const int U[3] = { Ux, Uy, Uz };
			const int u = linIdx(gf, Ux, Uy, Uz);
			if (tf[u] == vtInactive) continue;
			// N must be reachable from U: (U-1)/2 <= N <= min(size-1,(U+2)/2)  (C division truncates like Vec3i '/')
			bool ok = true;
			for (int d = 0; d < 3; d++) { const int lo = (U[d] - 1) / 2, hi = min(cs[d] - 1, (U[d] + 2) / 2); if (N[d] < lo || N[d] > hi) ok = false; }
			if (!ok) continue;
			const Real rw = pow2weight<Real>((Ux & 1) + (Uy & 1) + (Uz & 1));
			int w0[3], w1[3];
			for (int d = 0; d < 3; d++) { w0[d] = max(0, max(U[d] - 1, N[d] * 2 - 1)); w1[d] = min(fs[d] - 1, min(U[d] + 1, N[d] * 2 + 1)); }
			for (int Wz = w0[2]; Wz <= w1[2]; Wz++) for (int Wy = w0[1]; Wy <= w1[1]; Wy++) for (int Wx = w0[0]; Wx <= w1[0]; Wx++) {
				const int w = linIdx(gf, Wx, Wy, Wz);
				if (tf[w] == vtInactive) continue;
				const int sf = (Wx - Ux + 1) + 3 * (Wy - Uy + 1) + 9 * (Wz - Uz + smaxz);
				const Real iw = pow2weight<Real>((Wx & 1) + (Wy & 1) + (Wz & 1));
				const Real a = (sf < S) ? Af[(size_t)(S - 1 - sf) * gf.n + w] : Af[(size_t)(sf - S + 1) * gf.n + u];
				acc += rw * a * iw;
			}
		}
	}
	A[(size_t)e * outStride + vOut] = acc;
}

// ---- regular vertices.  In the interior of the fluid the operator is the same at every vertex, and so is everything a coarse row is built
// from; the Galerkin kernels above spend ~7 k instructions per coarse vertex on re-deriving the same 14 numbers (11 + 10 ms for levels 1 and 2 of
// 512^3).  A coarse vertex is REGULAR when every fine vertex the product reads is in the grid, active and itself regular (level 0: operator mask
// == 6 couplings of -1 and diagonal 6) and its 14 stored coarse neighbours are active.  The row of ONE regular vertex per level is computed by
// the kernel above (`only`), the others copy it: the same arithmetic on the same values gives the same bits.  3-D only.
// the irregular active vertices are listed (one atomic per warp) so that the product kernels run over full warps of them: walls are planes, and a
// vertex-per-thread launch with the regular ones skipped leaves one busy lane per warp along an x wall
__device__ __forceinline__ void listAppend(int* __restrict__ list, int* count, int v) {
	const unsigned m = __activemask();
	const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
	int base = 0;
	if (lane == leader) base = atomicAdd(count, __popc(m));
	base = __shfl_sync(m, base, leader);
	list[base + __popc(m & ((1u << lane) - 1u))] = v;
}
__device__ __forceinline__ bool coarseNbFull(const LvlGeom& gc, const signed char* __restrict__ tc, int vx, int vy, int vz) {
	for (int s = 13; s < 27; s++) {
		const int nx = vx + s % 3 - 1, ny = vy + (s / 3) % 3 - 1, nz = vz + s / 9 - 1;
		if (!inGrid(gc, nx, ny, nz) || tc[linIdx(gc, nx, ny, nz)] == vtInactive) return false;
	}
	return true;
}
__global__ void __launch_bounds__(128) k_mg_classify1(LvlGeom gf, LvlGeom gc, const unsigned short* __restrict__ mask, const signed char* __restrict__ tc,
	unsigned char* __restrict__ reg, int* first, int* __restrict__ irrList, int* irrCount)
{
	int vx, vy, vz;
	if (!cell3(gc.sx, vx, vy, vz)) return;
	const int v = linIdx(gc, vx, vy, vz);
	const unsigned short kRegular = (unsigned short)(1u | 0x7eu | (6u << 7));      // mp_mg_l0_fused.cuh: active, six couplings of -1, diagonal 6
	bool ok = tc[v] != vtInactive && vx > 0 && vy > 0 && vz > 0 && 2 * vx + 1 < gf.sx && 2 * vy + 1 < gf.sy && 2 * vz + 1 < gf.sz;
	if (ok) {
		for (int dz = -1; dz <= 1 && ok; dz++) for (int dy = -1; dy <= 1 && ok; dy++) for (int dx = -1; dx <= 1; dx++)
			if (mask[linIdx(gf, 2 * vx + dx, 2 * vy + dy, 2 * vz + dz)] != kRegular) { ok = false; break; }
	}
	ok = ok && coarseNbFull(gc, tc, vx, vy, vz);
	reg[v] = ok ? 1 : 0;
	if (ok) atomicMin(first, v);
	else if (tc[v] != vtInactive) listAppend(irrList, irrCount, v);
}
__global__ void __launch_bounds__(128) k_mg_classifyN(LvlGeom gf, LvlGeom gc, const unsigned char* __restrict__ regF, const signed char* __restrict__ tf,
	const signed char* __restrict__ tc, unsigned char* __restrict__ reg, int* first, int* __restrict__ irrList, int* irrCount)
{
	int vx, vy, vz;
	if (!cell3(gc.sx, vx, vy, vz)) return;
	const int v = linIdx(gc, vx, vy, vz);
	// the product reads the rows of the fine vertices within 2 of 2V (U within 1, W within 1 of U)
	bool ok = tc[v] != vtInactive && 2 * vx - 2 >= 0 && 2 * vy - 2 >= 0 && 2 * vz - 2 >= 0 && 2 * vx + 2 < gf.sx && 2 * vy + 2 < gf.sy && 2 * vz + 2 < gf.sz;
	if (ok) {
		for (int dz = -2; dz <= 2 && ok; dz++) for (int dy = -2; dy <= 2 && ok; dy++) for (int dx = -2; dx <= 2; dx++) {
			const int u = linIdx(gf, 2 * vx + dx, 2 * vy + dy, 2 * vz + dz);
			if (!regF[u] || tf[u] == vtInactive) { ok = false; break; }
		}
	}
	ok = ok && coarseNbFull(gc, tc, vx, vy, vz);
	reg[v] = ok ? 1 : 0;
	if (ok) atomicMin(first, v);
	else if (tc[v] != vtInactive) listAppend(irrList, irrCount, v);
}
template <typename Real>
__global__ void __launch_bounds__(256) k_mg_fill_regular(int n, int S, const unsigned char* __restrict__ reg, const Real* __restrict__ cst, Real* __restrict__ A)
{
	const int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n || !reg[v]) return;
	for (int e = 0; e < S; e++) A[(size_t)e * n + v] = cst[e];
}

// ---------------------------------------------------------------- V-cycle kernels
template <typename Real>
__global__ void __launch_bounds__(256) k_mg_set_rhs(int n, Real trivialScale, const Real* __restrict__ rhs, const signed char* __restrict__ type, Real* __restrict__ b, const int* doneFlag)
{
	if (doneFlag && *doneFlag) return;
	const int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n) return;
	Real val = rhs[v];
	if (type[v] == vtActiveTrivial) val *= trivialScale;    // knSetRhs :417-424
	b[v] = val;
}

// level-0 colour sweep: colour = (x+y+z) parity ({a0,a3,a5,a6} / {a1,a2,a4,a7}, :721); thread = (x pair, y, z)
template <typename Real, bool ZEROX>
__global__ void __launch_bounds__(128) k_mg_smooth0(LvlGeom g, int is3D, int color, const Real* __restrict__ A, const Real* __restrict__ b, Real bscale,
	const signed char* __restrict__ type, Real* __restrict__ x, const int* doneFlag)
{
	if (doneFlag && *doneFlag) return;
	int m, j, k;
	if (!cell3((g.sx + 1) >> 1, m, j, k)) return;
	const int i = 2 * m + ((j + k + color) & 1);
	if (i >= g.sx) return;
	const int v = linIdx(g, i, j, k);
	const signed char ty = type[v];
	if (ty == vtInactive) return;
	const size_t n = (size_t)g.n; const int Y = g.sx, Z = g.sx * g.sy;
	Real sum = b[v];
	if (bscale != (Real)0 && ty == vtActiveTrivial) sum *= bscale;
	if (!ZEROX) {
		if (i > 0)        sum -= A[n + v - 1] * x[v - 1];
		if (i < g.sx - 1) sum -= A[n + v] * x[v + 1];
		if (j > 0)        sum -= A[2 * n + v - Y] * x[v - Y];
		if (j < g.sy - 1) sum -= A[2 * n + v] * x[v + Y];
		if (is3D) {
			if (k > 0)        sum -= A[3 * n + v - Z] * x[v - Z];
			if (k < g.sz - 1) sum -= A[3 * n + v] * x[v + Z];
		}
	}
	x[v] = sum / A[v];
}

template <typename Real>
__global__ void __launch_bounds__(128) k_mg_residual0(LvlGeom g, int is3D, const Real* __restrict__ A, const Real* __restrict__ b, Real bscale,
	const signed char* __restrict__ type, const Real* __restrict__ x, Real* __restrict__ r, const int* doneFlag)
{
	if (doneFlag && *doneFlag) return;
	int i, j, k;
	if (!cell3(g.sx, i, j, k)) return;
	const int v = linIdx(g, i, j, k);
	const signed char ty = type[v];
	if (ty == vtInactive) return;
	const size_t n = (size_t)g.n; const int Y = g.sx, Z = g.sx * g.sy;
	Real sum = b[v];
	if (bscale != (Real)0 && ty == vtActiveTrivial) sum *= bscale;
	if (i > 0)        sum -= A[n + v - 1] * x[v - 1];
	if (i < g.sx - 1) sum -= A[n + v] * x[v + 1];
	if (j > 0)        sum -= A[2 * n + v - Y] * x[v - Y];
	if (j < g.sy - 1) sum -= A[2 * n + v] * x[v + Y];
	if (is3D) {
		if (k > 0)        sum -= A[3 * n + v - Z] * x[v - Z];
		if (k < g.sz - 1) sum -= A[3 * n + v] * x[v + Z];
	}
	sum -= A[v] * x[v];
	r[v] = sum;
}

// ---------------------------------------------------------------- level 0, 16 bytes of cells per thread
// Same arithmetic as k_mg_smooth0 / k_mg_residual0 / k_mg_interp_add / k_mg_restrict, cell for cell and term for term; what changes is
// the shape of the work: a thread owns V consecutive x-cells (one 16-byte vector), a CTA of 32 x 4 threads an x-y tile, and each CTA
// marches through `kchunk` z-planes, so level 0 of a 512^3 grid is 32 k fat CTAs instead of 1 M one-cell-per-thread ones and every
// array is read with 16-byte loads.  Needs sx % V == 0 (every row starts on a 16-byte boundary).
template <typename Real, int V> struct alignas(sizeof(Real) * V) RVec { Real v[V]; };
template <int V> struct alignas(V) CVec { signed char v[V]; };
template <typename Real, int V> __device__ __forceinline__ RVec<Real, V> ldR(const Real* p) { return *reinterpret_cast<const RVec<Real, V>*>(p); }
template <int V> __device__ __forceinline__ CVec<V> ldC(const signed char* p) { return *reinterpret_cast<const CVec<V>*>(p); }

// MODE 0: one colour of the smoother; MODE 1: the first colour on x == 0 (x = b / A0 on that colour, zeros elsewhere: writes EVERY
// cell, so the caller needs no memset of x); MODE 2: residual r = b - A x
template <typename Real, int V, int MODE>
__global__ void __launch_bounds__(128) k_mg_l0_vec(LvlGeom g, int is3D, int color, int nvx, int kchunk, int kb, int ke, const Real* __restrict__ A, const Real* __restrict__ b, Real bscale,
	const signed char* __restrict__ type, Real* __restrict__ x, Real* __restrict__ r, const int* doneFlag)
{
	if (doneFlag && *doneFlag) return;
	const int m = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y;
	if (m >= nvx || j >= g.sy) return;
	const int k0 = kb + blockIdx.z * kchunk, k1 = min(ke, k0 + kchunk);      // [kb,ke): all planes, or this rank's slab (b, x, r are then shifted views of slab arrays)
	const size_t n = (size_t)g.n; const int Y = g.sx, Z = g.sx * g.sy;
	const int i0 = m * V;
	for (int k = k0; k < k1; k++) {
		const int v = i0 + Y * j + Z * k;
		const CVec<V> ty = ldC<V>(type + v);
		bool mine[V]; bool any = false;
		#pragma unroll
		for (int q = 0; q < V; q++) { mine[q] = ty.v[q] != vtInactive && (MODE == 2 || ((i0 + q + j + k + color) & 1) == 0); any |= mine[q]; }
		if (MODE == 1) {
			RVec<Real, V> out;
			#pragma unroll
			for (int q = 0; q < V; q++) out.v[q] = (Real)0;
			if (any) {
				const RVec<Real, V> bv = ldR<Real, V>(b + v), a0 = ldR<Real, V>(A + v);
				#pragma unroll
				for (int q = 0; q < V; q++) if (mine[q]) {
					Real sum = bv.v[q];
					if (bscale != (Real)0 && ty.v[q] == vtActiveTrivial) sum *= bscale;
					out.v[q] = sum / a0.v[q];
				}
			}
			*reinterpret_cast<RVec<Real, V>*>(x + v) = out;
			continue;
		}
		if (!any) {
			if (MODE == 2) { RVec<Real, V> z; for (int q = 0; q < V; q++) z.v[q] = (Real)0; *reinterpret_cast<RVec<Real, V>*>(r + v) = z; }      // inactive vertices: r stays 0
			continue;
		}
		const RVec<Real, V> bv = ldR<Real, V>(b + v), a0 = ldR<Real, V>(A + v), ai = ldR<Real, V>(A + n + v), aj = ldR<Real, V>(A + 2 * n + v);
		RVec<Real, V> xc = ldR<Real, V>(x + v);
		RVec<Real, V> ajm, xym, xyp, ak, akm, xzm, xzp;
		#pragma unroll
		for (int q = 0; q < V; q++) { ajm.v[q] = xym.v[q] = xyp.v[q] = ak.v[q] = akm.v[q] = xzm.v[q] = xzp.v[q] = (Real)0; }
		if (j > 0) { ajm = ldR<Real, V>(A + 2 * n + v - Y); xym = ldR<Real, V>(x + v - Y); }
		if (j < g.sy - 1) xyp = ldR<Real, V>(x + v + Y);
		if (is3D) {
			ak = ldR<Real, V>(A + 3 * n + v);
			if (k > 0) { akm = ldR<Real, V>(A + 3 * n + v - Z); xzm = ldR<Real, V>(x + v - Z); }
			if (k < g.sz - 1) xzp = ldR<Real, V>(x + v + Z);
		}
		Real aim0 = (Real)0, xm0 = (Real)0, xpL = (Real)0;
		if (i0 > 0) { aim0 = A[n + v - 1]; xm0 = x[v - 1]; }
		if (i0 + V < g.sx) xpL = x[v + V];
		RVec<Real, V> out = xc;
		if (MODE == 2) { for (int q = 0; q < V; q++) out.v[q] = (Real)0; }
		#pragma unroll
		for (int q = 0; q < V; q++) if (mine[q]) {
			const int i = i0 + q;
			Real sum = bv.v[q];
			if (bscale != (Real)0 && ty.v[q] == vtActiveTrivial) sum *= bscale;
			if (i > 0)        sum -= (q == 0 ? aim0 : ai.v[q - 1 < 0 ? 0 : q - 1]) * (q == 0 ? xm0 : xc.v[q - 1 < 0 ? 0 : q - 1]);
			if (i < g.sx - 1) sum -= ai.v[q] * (q == V - 1 ? xpL : xc.v[q + 1 > V - 1 ? V - 1 : q + 1]);
			if (j > 0)        sum -= ajm.v[q] * xym.v[q];
			if (j < g.sy - 1) sum -= aj.v[q] * xyp.v[q];
			if (is3D) {
				if (k > 0)        sum -= akm.v[q] * xzm.v[q];
				if (k < g.sz - 1) sum -= ak.v[q] * xzp.v[q];
			}
			if (MODE == 2) { sum -= a0.v[q] * xc.v[q]; out.v[q] = sum; }
			else out.v[q] = sum / a0.v[q];
		}
		*reinterpret_cast<RVec<Real, V>*>((MODE == 2 ? r : x) + v) = out;
	}
}

// k_mg_l0_vec with the operator as the 2-byte mask of mp_mg_l0_fused.cuh instead of the type byte + four coefficient arrays: a pass
// streams 2 + 3w bytes per cell (MODE 0), 2 + 2w (MODE 1), 2 + 4w (MODE 2) instead of 1 + 7w.  Same terms in the same order; a coupling
// of -1 is added as `sum += x`, a coupling of +0 is skipped (signs of exact zeros aside, the same value).
template <typename Real, int V, int MODE>
__global__ void __launch_bounds__(128) k_mg_l0_vecm(LvlGeom g, int color, int nvx, int kchunk, int kb, int ke, const Real* __restrict__ A0, const unsigned short* __restrict__ mask,
	const Real* __restrict__ b, Real bscale, Real* __restrict__ x, Real* __restrict__ r, const int* doneFlag)
{
	if (doneFlag && *doneFlag) return;
	const int m = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y;
	if (m >= nvx || j >= g.sy) return;
	const int k0 = kb + blockIdx.z * kchunk, k1 = min(ke, k0 + kchunk);
	const int Y = g.sx, Z = g.sx * g.sy;
	const int i0 = m * V;
	typedef mgl0::Vec<unsigned short, V> MVec;
	for (int k = k0; k < k1; k++) {
		const int v = i0 + Y * j + Z * k;
		const MVec mv = *reinterpret_cast<const MVec*>(mask + v);
		unsigned mine = 0, any7 = 0;
		#pragma unroll
		for (int q = 0; q < V; q++) {
			if ((mv.v[q] & mgl0::mActive) && (MODE == 2 || ((i0 + q + j + k + color) & 1) == 0)) mine |= 1u << q;
			any7 |= (unsigned)(((mv.v[q] >> 7) & 7) == 7) << q;
		}
		RVec<Real, V> out;
		#pragma unroll
		for (int q = 0; q < V; q++) out.v[q] = (Real)0;
		if (!mine) {
			if (MODE == 1) *reinterpret_cast<RVec<Real, V>*>(x + v) = out;
			if (MODE == 2) *reinterpret_cast<RVec<Real, V>*>(r + v) = out;      // inactive vertices: r stays 0
			continue;
		}
		RVec<Real, V> bv = ldR<Real, V>(b + v), a0;
		#pragma unroll
		for (int q = 0; q < V; q++) a0.v[q] = (Real)(int)((mv.v[q] >> 7) & 7);
		if (any7 & mine) {      // ghost-fluid diagonals, trivial rows: the stored value
			#pragma unroll
			for (int q = 0; q < V; q++) if ((any7 >> q) & 1) a0.v[q] = A0[v + q];
		}
		if (bscale != (Real)0) {
			#pragma unroll
			for (int q = 0; q < V; q++) if (mv.v[q] & mgl0::mTrivial) bv.v[q] *= bscale;
		}
		if (MODE == 1) {
			#pragma unroll
			for (int q = 0; q < V; q++) if ((mine >> q) & 1) out.v[q] = bv.v[q] / a0.v[q];
			*reinterpret_cast<RVec<Real, V>*>(x + v) = out;
			continue;
		}
		const RVec<Real, V> xc = ldR<Real, V>(x + v);
		RVec<Real, V> xym, xyp, xzm, xzp;
		#pragma unroll
		for (int q = 0; q < V; q++) { xym.v[q] = xyp.v[q] = xzm.v[q] = xzp.v[q] = (Real)0; }
		if (j > 0) xym = ldR<Real, V>(x + v - Y);
		if (j < g.sy - 1) xyp = ldR<Real, V>(x + v + Y);
		if (k > 0) xzm = ldR<Real, V>(x + v - Z);
		if (k < g.sz - 1) xzp = ldR<Real, V>(x + v + Z);
		Real xm0 = (Real)0, xpL = (Real)0;
		if (i0 > 0) xm0 = x[v - 1];
		if (i0 + V < g.sx) xpL = x[v + V];
		if (MODE == 0) out = xc;
		#pragma unroll
		for (int q = 0; q < V; q++) if ((mine >> q) & 1) {
			const unsigned mm = mv.v[q];
			Real sum = bv.v[q];
			if (mm & 2u)  sum += (q == 0 ? xm0 : xc.v[q - 1 < 0 ? 0 : q - 1]);
			if (mm & 4u)  sum += (q == V - 1 ? xpL : xc.v[q + 1 > V - 1 ? V - 1 : q + 1]);
			if (mm & 8u)  sum += xym.v[q];
			if (mm & 16u) sum += xyp.v[q];
			if (mm & 32u) sum += xzm.v[q];
			if (mm & 64u) sum += xzp.v[q];
			if (MODE == 2) { sum -= a0.v[q] * xc.v[q]; out.v[q] = sum; }
			else out.v[q] = sum / a0.v[q];
		}
		*reinterpret_cast<RVec<Real, V>*>((MODE == 2 ? r : x) + v) = out;
	}
}

// knInterpolate + knAddAssign from level 1 into the level-0 iterate, V fine cells per thread
template <typename Real, int V>
__global__ void __launch_bounds__(128) k_mg_interp_add_l0_vec(LvlGeom gf, LvlGeom gc, int nvx, int kchunk, int kb, int ke, const signed char* __restrict__ tf, const signed char* __restrict__ tc,
	const Real* __restrict__ xc, Real* xf, Real* xout, const int* doneFlag)      // xout == xf: in place; else every vertex of xout is written
{
	if (doneFlag && *doneFlag) return;
	const int m = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y;
	if (m >= nvx || y >= gf.sy) return;
	const int k0 = kb + blockIdx.z * kchunk, k1 = min(ke, k0 + kchunk);
	const int x0 = m * V, py = y & 1, cY = gc.sx, cZ = gc.sx * gc.sy;
	for (int z = k0; z < k1; z++) {
		const int v = x0 + gf.sx * (y + gf.sy * z);
		const CVec<V> ty = ldC<V>(tf + v);
		bool any = false;
		#pragma unroll
		for (int q = 0; q < V; q++) any |= ty.v[q] != vtInactive;
		if (!any) { if (xout != xf) *reinterpret_cast<RVec<Real, V>*>(xout + v) = ldR<Real, V>(xf + v); continue; }
		const int pz = z & 1;
		const int base = (x0 >> 1) + cY * (y >> 1) + cZ * (z >> 1);
		// the V fine cells interpolate from the coarse vertices x0/2 .. x0/2 + V/2 of up to 4 coarse rows.  The reference leaves inactive parents
		// out; their iterate is an exact zero throughout the cycle (zeroed by the restriction, never smoothed or corrected), so adding it
		// gives the same sum (signs of exact zeros aside) and the coarse types need not be read
		Real pv[2][2][V / 2 + 1];
		#pragma unroll
		for (int dz = 0; dz < 2; dz++)
			#pragma unroll
			for (int dy = 0; dy < 2; dy++)
				#pragma unroll
				for (int e = 0; e <= V / 2; e++) {
					pv[dz][dy][e] = (Real)0;
					if (dz <= pz && dy <= py && (x0 >> 1) + e < gc.sx) pv[dz][dy][e] = xc[base + e + dy * cY + dz * cZ];
				}
		RVec<Real, V> xv = ldR<Real, V>(xf + v);
		#pragma unroll
		for (int q = 0; q < V; q++) {
			if (ty.v[q] == vtInactive) continue;
			const int px = q & 1, e = q >> 1;                       // x0 is even: parity of x0 + q is that of q
			Real sum = 0;
			#pragma unroll
			for (int dz = 0; dz < 2; dz++)
				#pragma unroll
				for (int dy = 0; dy < 2; dy++) {
					if (dz > pz || dy > py) continue;
					sum += pv[dz][dy][e];
					if (px) sum += pv[dz][dy][e + 1];
				}
			xv.v[q] += pow2weight<Real>(px + py + pz) * sum;
		}
		*reinterpret_cast<RVec<Real, V>*>(xout + v) = xv;
	}
}

// knRestrict of the level-0 residual onto level 1 (and x1 = 0), V / 2 coarse vertices per thread: their 27 fine vertices are
// one scalar + one 16-byte vector in each of 9 fine rows
template <typename Real, int V>
__global__ void __launch_bounds__(128) k_mg_restrict_l0_vec(LvlGeom gf, LvlGeom gc, int ncx, int Kb, const signed char* __restrict__ tf, const signed char* __restrict__ tc,
	const Real* __restrict__ src, Real* __restrict__ dst, Real* __restrict__ xc, const int* doneFlag)
{
	if (doneFlag && *doneFlag) return;
	constexpr int CW = V / 2;
	const int t = blockIdx.x * 32 + threadIdx.x, vy = blockIdx.y * 4 + threadIdx.y, vz = Kb + blockIdx.z;      // coarse planes [Kb, Kb + gridDim.z)
	if (t >= ncx || vy >= gc.sy) return;
	const int vx0 = t * CW;
	bool act[CW]; bool any = false; Real sum[CW];
	#pragma unroll
	for (int c = 0; c < CW; c++) {
		act[c] = false; sum[c] = (Real)0;
		if (vx0 + c < gc.sx) { const int v = linIdx(gc, vx0 + c, vy, vz); xc[v] = (Real)0; act[c] = tc[v] != vtInactive; any |= act[c]; }
	}
	if (!any) return;
	const int fx0 = 2 * vx0;                                         // multiple of V
	const bool vecIn = fx0 < gf.sx;                                  // sx % V == 0: the vector is inside or outside as a whole
	for (int rz = max(0, vz * 2 - 1); rz <= min(gf.sz - 1, vz * 2 + 1); rz++)
	for (int ry = max(0, vy * 2 - 1); ry <= min(gf.sy - 1, vy * 2 + 1); ry++) {
		const int row = gf.sx * (ry + gf.sy * rz);
		const Real wyz = pow2weight<Real>((ry & 1) + (rz & 1));
		// fine x = fx0 - 1 (scalar) and fx0 .. fx0 + V - 1 (vector); coarse vertex c takes fx0 + 2c - 1 .. fx0 + 2c + 1, ascending.  The reference
		// leaves inactive fine vertices out; their residual is an exact zero (written by the residual kernel), so adding it gives the same sum
		// (signs of exact zeros aside) and the fine types need not be read
		Real f[V + 1];
		f[0] = (Real)0;
		if (fx0 > 0) f[0] = src[row + fx0 - 1];
		if (vecIn) {
			const RVec<Real, V> sv = ldR<Real, V>(src + row + fx0);
			#pragma unroll
			for (int q = 0; q < V; q++) f[q + 1] = sv.v[q];
		} else {
			#pragma unroll
			for (int q = 0; q < V; q++) f[q + 1] = (Real)0;
		}
		#pragma unroll
		for (int c = 0; c < CW; c++) {
			if (!act[c]) continue;
			#pragma unroll
			for (int e = 0; e < 3; e++) {                             // rx = fx0 + 2c - 1 + e
				const int q = 2 * c + e;                             // index into f[]
				const int rx = fx0 + 2 * c - 1 + e;
				if (rx < 0 || rx >= gf.sx) continue;
				// weight 1 / 2^(#odd coordinates): rx is odd for e = 0, 2
				const Real rw = (e == 1) ? wyz : wyz * (Real)0.5;
				sum[c] += rw * f[q];
			}
		}
	}
	#pragma unroll
	for (int c = 0; c < CW; c++) if (act[c]) dst[linIdx(gc, vx0 + c, vy, vz)] = sum[c];
}

// ---------------------------------------------------------------- level 0 fused (mp_mg_l0_fused.cuh)
template <typename Real>
__global__ void __launch_bounds__(256) k_mg_build_mask0(LvlGeom g, int is3D, const Real* __restrict__ A, const signed char* __restrict__ type, unsigned short* __restrict__ mask, int* bad)
{
	int x, y, z;
	if (!cell3(g.sx, x, y, z)) return;
	const mgl0::Geom gg = { g.sx, g.sy, g.sz };
	int b = 0;
	mask[linIdx(g, x, y, z)] = mgl0::maskOf<Real>(gg, is3D, x, y, z, A, type, &b);
	if (b) *bad = 1;
}

// MODE_DOWN: x = two colour sweeps (c0, then 1 - c0) over a zero iterate, r = b - A x.  MODE_SMOOTH: x = sweep (c0, then c1) over xin
// (xin != xout: neighbouring CTAs read the halo of xin while this one writes its tile).
template <typename Real, int MODE>
__global__ void __launch_bounds__(mgl0::Tile<Real>::NTHR, 2) k_mg_l0_fused(mgl0::Geom g, int kchunk, int c0, int c1, const Real* __restrict__ A0, const Real* __restrict__ b, Real bscale,
	const unsigned short* __restrict__ mask, const Real* __restrict__ xin, Real* __restrict__ xout, Real* __restrict__ rout, const int* doneFlag)
{
	if (doneFlag && *doneFlag) return;
	typedef mgl0::Tile<Real> T;
	extern __shared__ __align__(16) unsigned char mgl0_smem[];
	mgl0::Smem<Real>& s = *reinterpret_cast<mgl0::Smem<Real>*>(mgl0_smem);
	const int x0 = blockIdx.x * T::TX, y0 = blockIdx.y * T::TY, k0 = blockIdx.z * kchunk, k1 = min(g.sz, k0 + kchunk);
	const int cm = MODE == mgl0::MODE_DOWN ? 1 - c0 : c0;      // the colour completed one plane behind the staging
	const mgl0::Ctx<Real> c = mgl0::makeCtx<Real>(g, x0, y0, threadIdx.x);
	mgl0::Pre<Real> p;
	for (int q = k0 - 2; q <= k0 + 1; q++) {
		mgl0::issue<Real, MODE>(g, c, q, b, xin, mask, p);
		mgl0::stage<Real, MODE>(g, c, q, bscale, A0, c0, p, s);
	}
	__syncthreads();
	mgl0::mid<Real>(g, c, k0 - 1, cm, A0, s);
	mgl0::mid<Real>(g, c, k0, cm, A0, s);
	mgl0::issue<Real, MODE>(g, c, k0 + 2, b, xin, mask, p);
	__syncthreads();
	for (int sp = k0; sp < k1; sp++) {
		mgl0::stage<Real, MODE>(g, c, sp + 2, bscale, A0, c0, p, s);
		if (sp + 1 < k1) mgl0::issue<Real, MODE>(g, c, sp + 3, b, xin, mask, p);      // in flight while this plane is computed
		__syncthreads();
		mgl0::mid<Real>(g, c, sp + 1, cm, A0, s);
		__syncthreads();
		mgl0::last<Real, MODE>(g, c, sp, c1, A0, s, xout, rout);
		__syncthreads();
	}
}

// 27-point (9-point in 2-D) stencil application shared by smoother / residual / coarse CG on levels > 0
template <typename Real, typename VecT, bool SKIPCENTER, bool IS3D>
__device__ __forceinline__ VecT stencilSubT(const LvlGeom& g, const Real* __restrict__ A, const signed char* __restrict__ type,
	const VecT* __restrict__ x, int v, int vx, int vy, int vz, VecT sum)
{
	constexpr int S = IS3D ? 14 : 5;
	// interior vertices (the bulk) need no bounds tests; the loops are fully unrolled (compile-time stencil)
	const bool interior = vx > 0 && vy > 0 && vx < g.sx - 1 && vy < g.sy - 1 && (!IS3D || (vz > 0 && vz < g.sz - 1));
	#pragma unroll
	for (int s = 0; s < (IS3D ? 27 : 9); s++) {
		if (SKIPCENTER && s == S - 1) continue;
		const int dx = s % 3 - 1, dy = (s / 3) % 3 - 1, dz = IS3D ? s / 9 - 1 : 0;
		if (!interior && !inGrid(g, vx + dx, vy + dy, vz + dz)) continue;
		const int nb = v + dx + g.sx * (dy + g.sy * dz);
		if (type[nb] == vtInactive) continue;
		if (s < S) sum -= A[(size_t)(S - 1 - s) * g.n + nb] * x[nb];
		else       sum -= A[(size_t)(s - S + 1) * g.n + v] * x[nb];
	}
	return sum;
}
template <typename Real, typename VecT, bool SKIPCENTER>
__device__ __forceinline__ VecT stencilSub(const LvlGeom& g, int is3D, int S, const Real* __restrict__ A, const signed char* __restrict__ type,
	const VecT* __restrict__ x, int v, int vx, int vy, int vz, VecT sum)
{
	(void)S;
	return is3D ? stencilSubT<Real, VecT, SKIPCENTER, true>(g, A, type, x, v, vx, vy, vz, sum)
	            : stencilSubT<Real, VecT, SKIPCENTER, false>(g, A, type, x, v, vx, vy, vz, sum);
}

// levels > 0: colour = offset inside 2x2x2 blocks (:722); thread = block
template <typename Real>
__global__ void __launch_bounds__(128) k_mg_smoothN(LvlGeom g, int is3D, int S, int ox, int oy, int oz, const Real* __restrict__ A, const Real* __restrict__ b,
	const signed char* __restrict__ type, Real* __restrict__ x, const int* doneFlag)
{
	if (doneFlag && *doneFlag) return;
	int tx, ty_, tz;
	if (!cell3((g.sx + 1) >> 1, tx, ty_, tz)) return;
	const int vx = 2 * tx + ox, vy = 2 * ty_ + oy, vz = 2 * tz + oz;
	if (!inGrid(g, vx, vy, vz)) return;
	const int v = linIdx(g, vx, vy, vz);
	if (type[v] == vtInactive) return;
	const Real sum = stencilSub<Real, Real, true>(g, is3D, S, A, type, x, v, vx, vy, vz, b[v]);
	x[v] = sum / A[v];
}

template <typename Real>
__global__ void __launch_bounds__(128) k_mg_residualN(LvlGeom g, int is3D, int S, const Real* __restrict__ A, const Real* __restrict__ b,
	const signed char* __restrict__ type, const Real* __restrict__ x, Real* __restrict__ r, const int* doneFlag)
{
	if (doneFlag && *doneFlag) return;
	int vx, vy, vz;
	if (!cell3(g.sx, vx, vy, vz)) return;
	const int v = linIdx(g, vx, vy, vz);
	if (type[v] == vtInactive) return;
	r[v] = stencilSub<Real, Real, false>(g, is3D, S, A, type, x, v, vx, vy, vz, b[v]);
}

// ---------------------------------------------------------------- levels > 0, colour-major rows
// The Galerkin operators are stored like the reference stores them: the 14 (5) "upper" entries per vertex, the lower half read from the
// neighbour's row.  A colour sweep over that layout reads x-stride-2 and touches every sector of 27 coefficient arrays 8 times per
// sweep (216 n B for 112 n B of coefficients).  For the sweeps -- and only for them -- setA also writes every vertex's WHOLE row, grouped by
// colour (= parity of x,y,z, the smoother's colours multigrid.cpp:722):
//     Afull[((c * NENT + s) * nc) + ci],   c = ox + 2 oy + 4 oz,  ci = (x>>1) + hbx ((y>>1) + hby (z>>1)),  s = stencil offset index 0..26
// Entries towards vertices outside the grid or inactive ones are 0, so the sweep needs no neighbour tests: it subtracts 0 * x instead of
// skipping (same value; an exact zero at worst changes sign).  One sweep then reads 108 n B of coefficients once, unit stride.
template <typename Real, bool IS3D>
__global__ void __launch_bounds__(128) k_mg_build_full(LvlGeom g, int hbx, int hby, int hbz, const Real* __restrict__ A, const signed char* __restrict__ type, Real* __restrict__ Afull,
	const int* __restrict__ only, Real* cstFull, unsigned char* __restrict__ rowreg)
{
	// only != NULL: one thread assembles the row of vertex *only into cstFull.  Otherwise, with rowreg != NULL, every row is compared with cstFull:
	// rowreg = 1 where all 27 entries are the same numbers (the sweeps then take them from cstFull)
	constexpr int S = IS3D ? 14 : 5, NENT = IS3D ? 27 : 9;
	int tx = blockIdx.x * blockDim.x + threadIdx.x, ty = blockIdx.y, tzc = blockIdx.z;
	int c, tz;
	if (only) {
		if (tx != 0 || ty != 0 || tzc != 0 || *only >= g.n) return;
		int ox, oy, oz; vecIdx(g, *only, ox, oy, oz);
		tx = ox >> 1; ty = oy >> 1; tz = oz >> 1; c = (ox & 1) | ((oy & 1) << 1) | ((oz & 1) << 2);
	} else {
		if (tx >= hbx) return;
		c = tzc / hbz; tz = tzc - c * hbz;
	}
	const int vx = 2 * tx + (c & 1), vy = 2 * ty + ((c >> 1) & 1), vz = 2 * tz + ((c >> 2) & 1);
	const size_t nc = (size_t)hbx * hby * hbz, ci = tx + (size_t)hbx * (ty + (size_t)hby * tz);
	Real* out = Afull + (size_t)c * NENT * nc + ci;
	const bool live = inGrid(g, vx, vy, vz) && type[linIdx(g, vx, vy, vz)] != vtInactive;
	const int v = live ? linIdx(g, vx, vy, vz) : 0;
	bool same = live;
	#pragma unroll
	for (int s = 0; s < NENT; s++) {
		const int dx = s % 3 - 1, dy = (s / 3) % 3 - 1, dz = IS3D ? s / 9 - 1 : 0;
		Real val = (Real)0;
		if (live && inGrid(g, vx + dx, vy + dy, vz + dz)) {
			const int nb = v + dx + g.sx * (dy + g.sy * dz);
			if (type[nb] != vtInactive) val = (s < S) ? A[(size_t)(S - 1 - s) * g.n + nb] : A[(size_t)(s - S + 1) * g.n + v];
		}
		if (only) cstFull[s] = val;
		else {
			out[(size_t)s * nc] = val;
			if (rowreg) same = same && (val == cstFull[s]) && (val != (Real)0 || !signbit(val) == !signbit(cstFull[s]));      // the same number, zeros of the same sign
		}
	}
	if (!only && rowreg) rowreg[(size_t)c * nc + ci] = same ? 1 : 0;
}

// one colour of knSmoothColor (:668-711) / knCalcResidual (:739-771) over the colour-major rows; RESID: all colours in one launch.
// Threads are numbered along the 2x2-block rows of a plane (no CTA is left with a one-vertex tail when a row is 128 k + 1 blocks long); a
// thread first puts its whole row of coefficients and its neighbours' iterate in flight (streaming loads for the coefficients, which are
// read once per sweep, so that the iterate -- re-read by every colour -- stays in L2), then accumulates in the reference's order.
template <typename Real, bool IS3D, bool RESID>
__device__ __forceinline__ void sweepVertex(const LvlGeom& g, int hbx, int hby, int hbz, int c, int tx, int ty, int tz, int Kb, int Ke, const Real* __restrict__ Afull,
	const Real* __restrict__ b, const signed char* __restrict__ type, Real* x, Real* __restrict__ r, const unsigned char* __restrict__ rowreg, const Real* __restrict__ cstFull)
{
	constexpr int S = IS3D ? 14 : 5, NENT = IS3D ? 27 : 9;
	const int vx = 2 * tx + (c & 1), vy = 2 * ty + ((c >> 1) & 1), vz = 2 * tz + ((c >> 2) & 1);
	if (!inGrid(g, vx, vy, vz) || vz < Kb || vz >= Ke) return;
	const int v = linIdx(g, vx, vy, vz);
	if (type[v] == vtInactive) return;
	const size_t nc = (size_t)hbx * hby * hbz;
	const Real* a = Afull + (size_t)c * NENT * nc + (tx + (size_t)hbx * (ty + (size_t)hby * tz));
	const bool interior = vx > 0 && vy > 0 && vx < g.sx - 1 && vy < g.sy - 1 && (!IS3D || (vz > 0 && vz < g.sz - 1));
	Real av[NENT], xv[NENT];
	// a row that equals the level's constant row (setA compared all 27 numbers) is taken from there: the interior of the fluid streams no coefficients
	if (rowreg && rowreg[(size_t)c * nc + (tx + (size_t)hbx * (ty + (size_t)hby * tz))]) {
		#pragma unroll
		for (int s = 0; s < NENT; s++) av[s] = cstFull[s];
	} else {
		#pragma unroll
		for (int s = 0; s < NENT; s++) av[s] = __ldcs(a + (size_t)s * nc);
	}
	#pragma unroll
	for (int s = 0; s < NENT; s++) {
		const int dx = s % 3 - 1, dy = (s / 3) % 3 - 1, dz = IS3D ? s / 9 - 1 : 0;
		int nb = v + dx + g.sx * (dy + g.sy * dz);
		if (!interior && !inGrid(g, vx + dx, vy + dy, vz + dz)) nb = v;      // coefficient is 0 there
		xv[s] = x[nb];
	}
	Real sum = b[v];
	#pragma unroll
	for (int s = 0; s < NENT; s++) {
		if (!RESID && s == S - 1) continue;
		sum -= av[s] * xv[s];
	}
	if (RESID) r[v] = sum; else x[v] = sum / av[S - 1];
}
template <typename Real, bool IS3D, bool RESID>
__global__ void __launch_bounds__(128, 6) k_mg_sweep_full(LvlGeom g, int hbx, int hby, int hbz, int color, int tz0, int ntz, int Kb, int Ke, const Real* __restrict__ Afull, const Real* __restrict__ b,
	const signed char* __restrict__ type, Real* x, Real* __restrict__ r, const int* doneFlag, const unsigned char* __restrict__ rowreg, const Real* __restrict__ cstFull)
{
	if (doneFlag && *doneFlag) return;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= hbx * hby) return;
	const int ty = t / hbx, tx = t - ty * hbx;
	// blockIdx.z walks the 2-plane blocks [tz0, tz0 + ntz) (all of them, or those that hold this rank's planes [Kb,Ke)); RESID: x all colours
	int c = color, tz = tz0 + blockIdx.z;
	if (RESID) { c = blockIdx.z / ntz; tz = tz0 + blockIdx.z - c * ntz; }
	sweepVertex<Real, IS3D, RESID>(g, hbx, hby, hbz, c, tx, ty, tz, Kb, Ke, Afull, b, type, x, r, rowreg, cstFull);
}
// small levels (<= 6000 vertices, 3-D; a 33^3 level is already faster as eight launches over many SMs): all eight colours of a sweep -- and, if asked, the residual after it -- by ONE CTA with a barrier between the
// colours, instead of eight (nine) launches of a few microseconds each; same per-vertex code, same order of the colours
template <typename Real>
__global__ void __launch_bounds__(512, 1) k_mg_sweep_cta(LvlGeom g, int hbx, int hby, int hbz, int reversed, int withResidual, const Real* __restrict__ Afull, const Real* __restrict__ b,
	const signed char* __restrict__ type, Real* x, Real* __restrict__ r, const int* doneFlag, const unsigned char* __restrict__ rowreg, const Real* __restrict__ cstFull)
{
	if (doneFlag && *doneFlag) return;
	const int nblk = hbx * hby * hbz, plane = hbx * hby;
	for (int cc = 0; cc < 8; cc++) {
		const int c = reversed ? 7 - cc : cc;
		for (int t = threadIdx.x; t < nblk; t += blockDim.x) {
			const int tz = t / plane, rem = t - tz * plane, ty = rem / hbx, tx = rem - ty * hbx;
			sweepVertex<Real, true, false>(g, hbx, hby, hbz, c, tx, ty, tz, 0, g.sz, Afull, b, type, x, r, rowreg, cstFull);
		}
		__syncthreads();      // the next colour reads what this one wrote (one CTA: its own global writes are visible after the barrier)
	}
	if (withResidual) {
		for (int t = threadIdx.x; t < 8 * nblk; t += blockDim.x) {
			const int c = t / nblk, q = t - c * nblk, tz = q / plane, rem = q - tz * plane, ty = rem / hbx, tx = rem - ty * hbx;
			sweepVertex<Real, true, true>(g, hbx, hby, hbz, c, tx, ty, tz, 0, g.sz, Afull, b, type, x, r, rowreg, cstFull);
		}
	}
}

// knRestrict :904-927 (dst level = coarse), also zeroes x on the coarse level (knSet :472)
template <typename Real>
__global__ void __launch_bounds__(128) k_mg_restrict(LvlGeom gf, LvlGeom gc, int Kb, const signed char* __restrict__ tf, const signed char* __restrict__ tc,
	const Real* __restrict__ src, Real* __restrict__ dst, Real* __restrict__ xc, const int* doneFlag)
{
	if (doneFlag && *doneFlag) return;
	int vx, vy, vz;
	if (!cell3(gc.sx, vx, vy, vz)) return;
	vz += Kb;                                   // coarse planes [Kb, Kb + gridDim.z)
	const int v = linIdx(gc, vx, vy, vz);
	xc[v] = (Real)0;
	if (tc[v] == vtInactive) return;
	// the residual of an inactive fine vertex is an exact zero (no kernel writes it; setA leaves zeros): adding it is the reference's skip
	Real sum = 0;
	for (int rz = max(0, vz * 2 - 1); rz <= min(gf.sz - 1, vz * 2 + 1); rz++)
	for (int ry = max(0, vy * 2 - 1); ry <= min(gf.sy - 1, vy * 2 + 1); ry++)
	for (int rx = max(0, vx * 2 - 1); rx <= min(gf.sx - 1, vx * 2 + 1); rx++) {
		const int r = linIdx(gf, rx, ry, rz);
		const Real rw = pow2weight<Real>((rx & 1) + (ry & 1) + (rz & 1));
		sum += rw * src[r];
	}
	dst[v] = sum;
}

// knInterpolate :934-954 into r_l, then x_l += r_l (knAddAssign :445-446, over ALL vertices: inactive ones add their stale r)
template <typename Real>
__global__ void __launch_bounds__(128) k_mg_interp_add(LvlGeom gf, LvlGeom gc, int kb, const signed char* __restrict__ tf, const signed char* __restrict__ tc,
	const Real* __restrict__ xc, Real* __restrict__ rf, Real* __restrict__ xf, const int* doneFlag)
{
	if (doneFlag && *doneFlag) return;
	int x, y, z;
	if (!cell3(gf.sx, x, y, z)) return;
	z += kb;                                    // fine planes [kb, kb + gridDim.z)
	const int v = linIdx(gf, x, y, z);
	// inactive vertices: the reference adds their r entry, which is never written and stays 0 -> nothing to do;
	// the interpolated value itself (mr[l] in the reference) is a temporary and is not stored
	if (tf[v] == vtInactive) return;
	// parents (x>>1 .. (x+1)>>1) x (y..) x (z..), summed in the reference's order (x fastest); y/z parity is warp-uniform
	const int px = x & 1, py = y & 1, pz = z & 1;
	const int base = linIdx(gc, x >> 1, y >> 1, z >> 1);
	// the iterate of an inactive coarse vertex is an exact zero throughout the cycle: adding it is the reference's skip
	Real sum = 0;
	for (int dz = 0; dz <= pz; dz++) for (int dy = 0; dy <= py; dy++) {
		const int i0 = base + dy * gc.sx + dz * gc.sx * gc.sy;
		sum += xc[i0];
		if (px) sum += xc[i0 + 1];
	}
	const Real iw = pow2weight<Real>(px + py + pz);
	xf[v] += iw * sum;
}

// solveCG :796-902 on the coarsest level: one CTA, double precision, Jacobi preconditioner
__device__ __forceinline__ double ctaSum(double v, double* sh) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	v = warpSum(v);
	__syncthreads();
	if (lane == 0) sh[warp] = v;
	__syncthreads();
	double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
	if (warp == 0) { t = warpSum(t); if (lane == 0) sh[32] = t; }
	__syncthreads();
	return sh[32];
}
template <typename Real>
__global__ void __launch_bounds__(1024) k_mg_coarse_cg(LvlGeom g, int is3D, int S, int level0, const Real* __restrict__ A, const Real* __restrict__ b,
	const signed char* __restrict__ type, Real* __restrict__ xr, double* __restrict__ cg, double accuracy, int* flagsOut, const int* doneFlag, Real bscale)
{
	if (doneFlag && *doneFlag) return;
	__shared__ double sh[33];
	const int n = g.n;
	double *z = cg, *p = cg + n, *x = cg + 2 * n, *r = cg + 3 * n;
	for (int v = threadIdx.x; v < n; v += blockDim.x) x[v] = (double)xr[v];
	__syncthreads();
	auto applyA = [&](int v, const double* vec) -> double {
		int vx, vy, vz; vecIdx(g, v, vx, vy, vz);
		if (level0) {
			double sum = 0; const int Y = g.sx, Z = g.sx * g.sy;
			if (vx > 0)        sum += (double)A[n + v - 1] * vec[v - 1];
			if (vx < g.sx - 1) sum += (double)A[n + v] * vec[v + 1];
			if (vy > 0)        sum += (double)A[2 * n + v - Y] * vec[v - Y];
			if (vy < g.sy - 1) sum += (double)A[2 * n + v] * vec[v + Y];
			if (is3D) { if (vz > 0) sum += (double)A[3 * n + v - Z] * vec[v - Z]; if (vz < g.sz - 1) sum += (double)A[3 * n + v] * vec[v + Z]; }
			sum += (double)A[v] * vec[v];
			return sum;
		}
		return -stencilSub<Real, double, false>(g, is3D, S, A, type, vec, v, vx, vy, vz, 0.0);
	};
	double aTop = 0, res0 = 0;
	for (int v = threadIdx.x; v < n; v += blockDim.x) {
		if (type[v] == vtInactive) continue;
		Real bv = b[v];
		if (bscale != (Real)0 && type[v] == vtActiveTrivial) bv *= bscale;
		const double rv = (double)bv - applyA(v, x);
		const double zv = rv / (double)A[v];
		r[v] = rv; z[v] = zv; p[v] = zv;
		res0 += rv * rv; aTop += rv * zv;
	}
	aTop = ctaSum(aTop, sh); res0 = sqrt(ctaSum(res0, sh));
	int iter = 0; const int maxIter = 10000;
	for (; iter < maxIter && res0 > 1E-12; iter++) {
		double aBot = 0;
		for (int v = threadIdx.x; v < n; v += blockDim.x) {
			if (type[v] == vtInactive) continue;
			const double zv = applyA(v, p);
			z[v] = zv; aBot += p[v] * zv;
		}
		aBot = ctaSum(aBot, sh);
		const double alpha = aTop / aBot;
		double aTopNew = 0, res = 0;
		for (int v = threadIdx.x; v < n; v += blockDim.x) {
			if (type[v] == vtInactive) continue;
			x[v] += alpha * p[v];
			const double rv = r[v] - alpha * z[v];
			r[v] = rv; res += rv * rv;
			const double zv = rv / (double)A[v];
			z[v] = zv; aTopNew += rv * zv;
		}
		aTopNew = ctaSum(aTopNew, sh); res = sqrt(ctaSum(res, sh));
		if (res / res0 < accuracy) break;
		const double beta = aTopNew / aTop;
		aTop = aTopNew;
		for (int v = threadIdx.x; v < n; v += blockDim.x) p[v] = z[v] + beta * p[v];
		__syncthreads();
	}
	for (int v = threadIdx.x; v < n; v += blockDim.x) xr[v] = (Real)x[v];
	if (threadIdx.x == 0) flagsOut[4] = iter;
}

// The same solve for a coarsest level of <= 1024 vertices with colour-major rows (the usual case): one vertex per thread, the four CG vectors in
// shared memory, the operator applied through the whole-row layout (entries towards inactive or outside vertices are 0: no type tests, 27
// independent coefficient loads), two sums per reduction pass.  Same operations on the same values in the same order per quantity as
// k_mg_coarse_cg (an absent neighbour contributes 0 * 0 instead of being skipped).
__device__ __forceinline__ void ctaSum2(double& a, double& b, double (*sh)[33]) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	a = warpSum(a); b = warpSum(b);
	__syncthreads();
	if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; }
	__syncthreads();
	double ta = (threadIdx.x < (blockDim.x >> 5)) ? sh[0][threadIdx.x] : 0.0, tb = (threadIdx.x < (blockDim.x >> 5)) ? sh[1][threadIdx.x] : 0.0;
	if (warp == 0) { ta = warpSum(ta); tb = warpSum(tb); if (lane == 0) { sh[0][32] = ta; sh[1][32] = tb; } }
	__syncthreads();
	a = sh[0][32]; b = sh[1][32];
}
template <typename Real>
__global__ void __launch_bounds__(1024, 1) k_mg_coarse_cg_rows(LvlGeom g, int hbx, int hby, int hbz, const Real* __restrict__ Afull, const Real* __restrict__ b,
	const signed char* __restrict__ type, Real* __restrict__ xr, double accuracy, int* flagsOut, const int* doneFlag)
{
	if (doneFlag && *doneFlag) return;
	constexpr int S = 14, NENT = 27;
	__shared__ double sh[2][33];
	__shared__ double X[1024], R[1024], P[1024], Z[1024];
	const int n = g.n, v = threadIdx.x;
	const bool in = v < n, act = in && type[v] != vtInactive;
	X[v] = in ? (double)xr[v] : 0.0; R[v] = 0.0; P[v] = 0.0; Z[v] = 0.0;
	// the row is re-read (L1) in every application: 27 registers more would spill at 1024 threads.  Neighbours outside the grid read the vertex
	// itself (their coefficient is 0): bit s of `inside`
	constexpr bool kRegs = false;
	Real a[kRegs ? NENT : 1];
	const Real* row = Afull;
	unsigned inside = 0;
	const size_t nc = (size_t)hbx * hby * hbz;
	const int Y = g.sx, Zs = g.sx * g.sy;
	if (act) {
		int vx, vy, vz; vecIdx(g, v, vx, vy, vz);
		const int c = (vx & 1) | ((vy & 1) << 1) | ((vz & 1) << 2);
		row = Afull + (size_t)c * NENT * nc + ((vx >> 1) + (size_t)hbx * ((vy >> 1) + (size_t)hby * (vz >> 1)));
		#pragma unroll
		for (int s = 0; s < NENT; s++) {
			const int dx = s % 3 - 1, dy = (s / 3) % 3 - 1, dz = s / 9 - 1;
			if (inGrid(g, vx + dx, vy + dy, vz + dz)) inside |= 1u << s;
			if (kRegs) a[s] = row[(size_t)s * nc];
		}
	}
	__syncthreads();
	auto applyA = [&](const double* vec) -> double {      // -(0 - a0 x0 - a1 x1 ...) as stencilSub accumulates it
		double sum = 0.0;
		#pragma unroll
		for (int s = 0; s < NENT; s++) {
			const int dx = s % 3 - 1, dy = (s / 3) % 3 - 1, dz = s / 9 - 1;
			const int nbv = ((inside >> s) & 1u) ? v + dx + Y * dy + Zs * dz : v;
			const Real as = kRegs ? a[kRegs ? s : 0] : row[(size_t)s * nc];
			sum -= (double)as * vec[nbv];
		}
		return -sum;
	};
	const double diag = act ? (double)row[(size_t)(S - 1) * nc] : 1.0;
	double aTop = 0, res0 = 0;
	if (act) {
		const double rv = (double)b[v] - applyA(X);
		const double zv = rv / diag;
		R[v] = rv; Z[v] = zv; P[v] = zv;
		res0 = rv * rv; aTop = rv * zv;
	}
	ctaSum2(aTop, res0, sh); res0 = sqrt(res0);
	int iter = 0; const int maxIter = 10000;
	for (; iter < maxIter && res0 > 1E-12; iter++) {
		double aBot = 0, dummy = 0, zv = 0;
		if (act) { zv = applyA(P); aBot = P[v] * zv; }
		__syncthreads();                                   // every thread has read P / nothing reads Z of this iteration before it is written
		if (act) Z[v] = zv;
		ctaSum2(aBot, dummy, sh);
		const double alpha = aTop / aBot;
		double aTopNew = 0, res = 0;
		if (act) {
			X[v] += alpha * P[v];
			const double rv = R[v] - alpha * Z[v];
			R[v] = rv; res = rv * rv;
			const double zn = rv / diag;
			Z[v] = zn; aTopNew = rv * zn;
		}
		ctaSum2(aTopNew, res, sh); res = sqrt(res);
		if (res / res0 < accuracy) break;
		const double beta = aTopNew / aTop;
		aTop = aTopNew;
		if (act) P[v] = Z[v] + beta * P[v];
		__syncthreads();
	}
	if (in) xr[v] = (Real)X[v];
	if (threadIdx.x == 0) flagsOut[4] = iter;
}

template <typename Real>
__global__ void __launch_bounds__(256) k_mg_copy(int n, const Real* __restrict__ src, Real* __restrict__ dst, const int* doneFlag) {
	if (doneFlag && *doneFlag) return;
	const int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v < n) dst[v] = src[v];
}
template <typename Real>
__global__ void __launch_bounds__(256) k_mg_norm(int n, const Real* __restrict__ r, const signed char* __restrict__ type, double* partials, unsigned int* ticket, double* out) {
	double v[1] = { 0.0 };   // knResidualNormSumSqr :778-784 (Real products, here accumulated in double)
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) if (type[i] != vtInactive) v[0] += (double)(r[i] * r[i]);
	const bool isMax[1] = { false }; double fin[1];
	if (blockReduceFinal<1>(v, isMax, partials, ticket, fin) && threadIdx.x == 0) out[0] = fin[0];
}

// ================================================================ host side
static inline unsigned int nb(long long work, int block) { return (unsigned int)((work + block - 1) / block); }

template <typename Real>
static int mgSetA(mp_mg* m, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak)
{
	mp_context* ctx = m->ctx; cudaStream_t st = ctx->stream;
	MP_CUDA(cudaMemsetAsync(m->dFlags, 0, 8 * sizeof(int), st));
	const LvlGeom g0 = m->geom[0];
	if (m->slab) {
		// the ranks' owned planes of A0/Ai/Aj/Ak straight into the global struct-of-arrays copy, then types and trivial-row scaling in place
		const size_t planeBytes = (size_t)g0.sx * g0.sy * sizeof(Real), n = (size_t)g0.n;
		const mp_grid* src[4] = { A0, Ai, Aj, Ak };
		for (int q = 0; q < 4; q++) MP_TRY(mp_dist_gather_planes(ctx, src[q]->d, planeBytes, (Real*)m->A[0] + q * n));
		Real* A = (Real*)m->A[0];
		k_mg_copy_activate<Real><<<nb(g0.n, 256), 256, 0, st>>>(g0, m->is3D, (Real)m->trivialScale, A, A + n, A + 2 * n, A + 3 * n, A, m->type[0], m->dFlags);
	} else
	k_mg_copy_activate<Real><<<nb(g0.n, 256), 256, 0, st>>>(g0, m->is3D, (Real)m->trivialScale, (const Real*)A0->d, (const Real*)Ai->d, (const Real*)Aj->d,
		(const Real*)Ak->d, (Real*)m->A[0], m->type[0], m->dFlags);
	MP_CHECK_LAUNCH(ctx);
	m->mask0Valid = false;
	if (m->mask0) {
		// the operator as 2 bytes per vertex (mp_mg_l0_fused.cuh); dFlags[5] is raised when a row cannot be coded (face fractions)
		const int bs = g0.sx >= 96 ? 128 : (g0.sx >= 48 ? 64 : 32);
		k_mg_build_mask0<Real><<<grid3(g0.sx, g0.sy, g0.sz, bs), bs, 0, st>>>(g0, m->is3D, (const Real*)m->A[0], m->type[0], m->mask0, m->dFlags + 5);
		MP_CHECK_LAUNCH(ctx);
		MP_CUDA(cudaMemcpyAsync(m->hFlags + 5, m->dFlags + 5, sizeof(int), cudaMemcpyDeviceToHost, st));
		MP_CUDA(cudaStreamSynchronize(st));
		m->mask0Valid = m->hFlags[5] == 0;
	}
	m->hostCoarsenLevels = 0;
	bool regularLevel[MG_MAXLVL] = {};
	for (int l = 1; l < m->nlev; l++) {
		const LvlGeom gf = m->geom[l - 1], gc = m->geom[l];
		k_mg_fill_type<<<nb(gc.n, 256), 256, 0, st>>>(m->type[l], gc.n, vtFree); MP_CHECK_LAUNCH(ctx);
		// phase-1 closure to its fixed point: sweep until a sweep changes nothing; that sweep's "undecided" flag is then final.  Sweeps after the
		// first walk the list of vertices the sweep before left undecided (lists live in the level's b / r arrays, idle during setA)
		{
			int* lists[2] = { (int*)m->b[l - 1], (int*)m->r[l - 1] };
			const long long capLL = (l == 1 && m->slab) ? (long long)gf.sx * gf.sy * m->lsz : (long long)gf.n;
			const int cap = (int)std::min<long long>(capLL, gf.n);
			bool useLists = !(getenv("MP_MG_SELECT_LISTS") && !atoi(getenv("MP_MG_SELECT_LISTS")));
			int listed = 0, maxListed = 0;
			{ const int bs = gc.sx >= 96 ? 128 : (gc.sx >= 48 ? 64 : 32); k_mg_select_even<<<grid3(gc.sx, gc.sy, gc.sz, bs), bs, 0, st>>>(gf, gc, m->type[l - 1], m->type[l]); MP_CHECK_LAUNCH(ctx); }
			for (int sweep = 0; sweep < gf.sx + gf.sy + gf.sz + 8; sweep++) {
				MP_CUDA(cudaMemsetAsync(m->dFlags, 0, 2 * sizeof(int), st));
				const bool fromList = useLists && sweep > 0;
				int* out = useLists ? lists[sweep & 1] : nullptr;
				if (useLists) MP_CUDA(cudaMemsetAsync(m->dFlags + 8 + (sweep & 1), 0, sizeof(int), st));
				const int work = fromList ? listed : gf.n;
				if (work > 0) {
					k_mg_select<<<nb(work, 256), 256, 0, st>>>(gf, gc, m->type[l - 1], m->type[l], m->dFlags, fromList ? lists[(sweep - 1) & 1] : nullptr,
						fromList ? m->dFlags + 8 + ((sweep - 1) & 1) : nullptr, out, m->dFlags + 8 + (sweep & 1), cap);
					MP_CHECK_LAUNCH(ctx);
				}
				MP_CUDA(cudaMemcpyAsync(m->hFlags, m->dFlags, 10 * sizeof(int), cudaMemcpyDeviceToHost, st));
				MP_CUDA(cudaStreamSynchronize(st));
				if (useLists && m->hFlags[6]) {      // a list overflowed (cannot happen with cap == n; slab-sized r0 only): start over with whole sweeps
					useLists = false; MP_CUDA(cudaMemsetAsync(m->dFlags + 6, 0, sizeof(int), st));
					continue;
				}
				listed = m->hFlags[8 + (sweep & 1)];
				maxListed = std::max(maxListed, std::min(listed, cap));
				if (!m->hFlags[0]) break;
			}
			// b / r of the level go back to zeros where the lists were (inactive vertices are never written by the V-cycle)
			if (maxListed > 0) { MP_CUDA(cudaMemsetAsync(lists[0], 0, sizeof(int) * (size_t)maxListed, st)); MP_CUDA(cudaMemsetAsync(lists[1], 0, sizeof(int) * (size_t)maxListed, st)); }
		}
		if (m->hFlags[1]) {
			// order-dependent phase needed: redo this level with the exact serial algorithm
			std::vector<signed char> tf(gf.n), tc(gc.n);
			MP_CUDA(cudaMemcpy(tf.data(), m->type[l - 1], gf.n, cudaMemcpyDeviceToHost));
			mgcoarsen::selectCoarseVertices(mgcoarsen::Dim3i{ gf.sx, gf.sy, gf.sz }, mgcoarsen::Dim3i{ gc.sx, gc.sy, gc.sz }, m->is3D != 0, tf, tc);
			MP_CUDA(cudaMemcpy(m->type[l], tc.data(), gc.n, cudaMemcpyHostToDevice));
			m->hostCoarsenLevels++;
		} else {
			k_mg_activate_coarse<<<nb(gc.n, 256), 256, 0, st>>>(m->type[l], gc.n); MP_CHECK_LAUNCH(ctx);
		}
		const long long work = (long long)gc.n * m->stencil;
		static const int g1variant = getenv("MP_MG_GALERKIN1") ? atoi(getenv("MP_MG_GALERKIN1")) : 2;
		// regular vertices (see k_mg_classify1): classified first, skipped by the product, filled with the row of the level's first regular vertex
		const char* eReg = getenv("MP_MG_REGULAR");      // read per call: the parity tests run both forms in one process
		const bool prevOk = l == 1 ? (m->mask0 && m->mask0Valid) : regularLevel[l - 1];
		regularLevel[l] = (!eReg || atoi(eReg)) && m->is3D && prevOk && m->regular[l] && (l > 1 || g1variant == 2) && gc.n >= 4096;
		const unsigned char* skip = nullptr;
		int* irr = nullptr; int nIrr = 0;
		if (regularLevel[l]) {
			const int bs = gc.sx >= 96 ? 128 : (gc.sx >= 48 ? 64 : 32);
			MP_CUDA(cudaMemsetAsync(m->first, 0x7f, sizeof(int), st));
			MP_CUDA(cudaMemsetAsync(m->dFlags + 10, 0, sizeof(int), st));
			irr = (int*)m->x[l];      // the level's iterate is idle during setA (the restriction zeroes it in every V-cycle)
			if (l == 1) k_mg_classify1<<<grid3(gc.sx, gc.sy, gc.sz, bs), bs, 0, st>>>(gf, gc, m->mask0, m->type[1], m->regular[1], m->first, irr, m->dFlags + 10);
			else        k_mg_classifyN<<<grid3(gc.sx, gc.sy, gc.sz, bs), bs, 0, st>>>(gf, gc, m->regular[l - 1], m->type[l - 1], m->type[l], m->regular[l], m->first, irr, m->dFlags + 10);
			MP_CHECK_LAUNCH(ctx);
			MP_CUDA(cudaMemcpyAsync(m->hFlags + 10, m->dFlags + 10, sizeof(int), cudaMemcpyDeviceToHost, st));
			MP_CUDA(cudaStreamSynchronize(st));
			nIrr = m->hFlags[10];
			skip = m->regular[l];
		}
		if (regularLevel[l]) {      // the products of the listed (irregular) vertices only
			if (nIrr > 0) {
				if (l == 1) k_mg_galerkin1_v2<Real><<<nb(nIrr, 128), 128, 0, st>>>(gf, gc, m->is3D, (const Real*)m->A[0], m->type[0], m->type[1], (Real*)m->A[1], skip, nullptr, (size_t)gc.n, irr, nIrr);
				else        k_mg_galerkinN<Real><<<nb((long long)nIrr * m->stencil, 256), 256, 0, st>>>(gf, gc, m->stencil, m->is3D, (const Real*)m->A[l - 1], m->type[l - 1], m->type[l], (Real*)m->A[l], skip, nullptr, (size_t)gc.n, irr, nIrr);
			}
		}
		else if (l == 1 && g1variant == 2) k_mg_galerkin1_v2<Real><<<nb(gc.n, 128), 128, 0, st>>>(gf, gc, m->is3D, (const Real*)m->A[0], m->type[0], m->type[1], (Real*)m->A[1], nullptr, nullptr, (size_t)gc.n, nullptr, 0);
		else if (l == 1) k_mg_galerkin1<Real><<<nb(work, 256), 256, 0, st>>>(gf, gc, m->stencil, m->dPaths, (const int*)(m->dFlags + 16), (const Real*)m->A[0], m->type[0], m->type[1], (Real*)m->A[1]);
		else        k_mg_galerkinN<Real><<<nb(work, 256), 256, 0, st>>>(gf, gc, m->stencil, m->is3D, (const Real*)m->A[l - 1], m->type[l - 1], m->type[l], (Real*)m->A[l], nullptr, nullptr, (size_t)gc.n, nullptr, 0);
		MP_CHECK_LAUNCH(ctx);
		if (regularLevel[l]) {
			if (l == 1) k_mg_galerkin1_v2<Real><<<1, 128, 0, st>>>(gf, gc, m->is3D, (const Real*)m->A[0], m->type[0], m->type[1], (Real*)m->cst, nullptr, m->first, (size_t)1, nullptr, 0);
			else        k_mg_galerkinN<Real><<<1, 256, 0, st>>>(gf, gc, m->stencil, m->is3D, (const Real*)m->A[l - 1], m->type[l - 1], m->type[l], (Real*)m->cst, nullptr, m->first, (size_t)1, nullptr, 0);
			MP_CHECK_LAUNCH(ctx);
			k_mg_fill_regular<Real><<<nb(gc.n, 256), 256, 0, st>>>(gc.n, m->stencil, m->regular[l], (const Real*)m->cst, (Real*)m->A[l]);
			MP_CHECK_LAUNCH(ctx);
		}
		if (m->Afull[l]) {
			const int hbx = (gc.sx + 1) / 2, hby = (gc.sy + 1) / 2, hbz = m->is3D ? (gc.sz + 1) / 2 : 1, ncol = m->is3D ? 8 : 4;
			const dim3 gr((unsigned)((hbx + 127) / 128), (unsigned)hby, (unsigned)(hbz * ncol));
			// rows equal to the full row of the level's first regular vertex are flagged for the sweeps (MP_MG_ROWREG=0: every row is streamed)
			const char* eRow = getenv("MP_MG_ROWREG");
			m->rowregOn[l] = regularLevel[l] && m->rowreg[l] && (!eRow || atoi(eRow));
			Real* cf = (Real*)m->cstFull + 32 * l;
			if (m->rowregOn[l]) { k_mg_build_full<Real, true><<<1, 1, 0, st>>>(gc, hbx, hby, hbz, (const Real*)m->A[l], m->type[l], (Real*)m->Afull[l], m->first, cf, nullptr); MP_CHECK_LAUNCH(ctx); }
			unsigned char* rr = m->rowregOn[l] ? m->rowreg[l] : nullptr;
			if (m->is3D) k_mg_build_full<Real, true><<<gr, 128, 0, st>>>(gc, hbx, hby, hbz, (const Real*)m->A[l], m->type[l], (Real*)m->Afull[l], nullptr, cf, rr);
			else         k_mg_build_full<Real, false><<<gr, 128, 0, st>>>(gc, hbx, hby, hbz, (const Real*)m->A[l], m->type[l], (Real*)m->Afull[l], nullptr, cf, rr);
			MP_CHECK_LAUNCH(ctx);
		}
	}
	m->isASet = true; m->isRhsSet = false;
	return MP_OK;
}

// level-0 vectors of the running V-cycle: x0 is the caller's dst grid (no copy at the end), b0 either the scaled copy made
// by setRhs or the caller's rhs with trivial rows scaled on the fly (bscale != 0)
template <typename Real> struct L0 { Real* x; const Real* b; Real bscale; Real* xLocal; };      // xLocal: the slab array x is a shifted view of (slab mode)
// slab mode: level-0 vectors are slab arrays (ghost, owned planes, ghost) while A[0] / type[0] are global; the kernels index everything with
// the GLOBAL vertex index, so the slab arrays are passed as views shifted by the planes below the slab (only planes [k0-1, k1] are touched)
template <typename Real> static inline Real* slabView(const mp_mg* m, Real* localBase) {
	return m->slab ? localBase - (ptrdiff_t)(m->k0 - 1) * m->geom[0].sx * m->geom[0].sy : localBase;
}
template <typename Real> static inline Real* l0r(const mp_mg* m) { return slabView<Real>(m, (Real*)m->r[0]); }

// level-0 kernels with 16 bytes of cells per thread: rows must start on 16-byte boundaries
template <typename Real> static inline bool l0vec(const LvlGeom& g) {
	const char* e = getenv("MP_MG_L0VEC");      // read per call: the parity tests run both forms in one process
	return (!e || atoi(e)) && g.sx % (16 / (int)sizeof(Real)) == 0;
}
static inline int l0chunk(const LvlGeom& g) { return g.sz >= 64 ? 8 : (g.sz >= 8 ? 4 : 1); }

// fused level-0 kernels (mp_mg_l0_fused.cuh): single GPU, rows on 16-byte boundaries, operator codable as the mask
template <typename Real> static inline bool l0fused(const mp_mg* m) {
	// read per call: the parity tests run both forms in one process.  Default: on in double, off in float -- measured at 512^3 (profiles/r2_mg_bench.txt):
	// the fused pair is instruction bound (halo recomputation, three barriers per plane) and only beats the masked per-colour kernels in double
	const char* e = getenv("MP_MG_L0FUSED");
	return (e ? atoi(e) != 0 : sizeof(Real) == 8) && m->mask0 && m->mask0Valid && !m->slab && l0vec<Real>(m->geom[0]);
}
// planes per CTA: every chunk pays 4 planes of warm-up, and the CTAs should fill whole waves of 2 CTAs per SM
template <typename Real> static inline int l0fusedChunk(const mp_mg* m) {
	typedef mgl0::Tile<Real> T;
	const LvlGeom g = m->geom[0];
	const long long tiles = (long long)((g.sx + T::TX - 1) / T::TX) * ((g.sy + T::TY - 1) / T::TY), slots = 2ll * m->ctx->smCount;
	int best = g.sz; long long bestCost = -1;
	for (int nchunk = 1; nchunk <= std::max(1, g.sz / 8); nchunk++) {
		const int chunk = (g.sz + nchunk - 1) / nchunk;
		const long long ctas = tiles * ((g.sz + chunk - 1) / chunk), waves = (ctas + slots - 1) / slots;
		const long long cost = waves * (chunk + 4);
		if (bestCost < 0 || cost < bestCost) { bestCost = cost; best = chunk; }
	}
	return best;
}
// the per-colour level-0 kernels with the operator mask (k_mg_l0_vecm)
template <typename Real> static inline bool l0masked(const mp_mg* m) {
	const char* e = getenv("MP_MG_L0MASK");      // read per call
	// z-slabs: A[0] and the mask are those of the global grid on every rank (every rank takes the same decision), the kernels index them globally
	return (!e || atoi(e)) && m->mask0 && m->mask0Valid && l0vec<Real>(m->geom[0]);
}
template <typename Real, int MODE>
static int l0fusedLaunch(mp_mg* m, int c0, int c1, const Real* b, Real bscale, const Real* xin, Real* xout, Real* rout, const int* doneFlag)
{
	typedef mgl0::Tile<Real> T;
	const LvlGeom g = m->geom[0];
	static bool attr = false;      // per instantiation
	if (!attr) { MP_CUDA(cudaFuncSetAttribute(k_mg_l0_fused<Real, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(mgl0::Smem<Real>))); attr = true; }
	const int chunk = l0fusedChunk<Real>(m);
	const dim3 gr((unsigned)((g.sx + T::TX - 1) / T::TX), (unsigned)((g.sy + T::TY - 1) / T::TY), (unsigned)((g.sz + chunk - 1) / chunk));
	const mgl0::Geom gg = { g.sx, g.sy, g.sz };
	k_mg_l0_fused<Real, MODE><<<gr, T::NTHR, sizeof(mgl0::Smem<Real>), m->ctx->stream>>>(gg, chunk, c0, c1, (const Real*)m->A[0], b, bscale, m->mask0, xin, xout, rout, doneFlag);
	MP_CHECK_LAUNCH(m->ctx);
	return MP_OK;
}

// planes of level l this rank sweeps: its share on sharded levels, everything otherwise
static inline void lvlRange(const mp_mg* m, int l, int* kb, int* ke) {
	if (m->slab && l < m->nshard) { *kb = m->K0[l]; *ke = m->K1[l]; } else { *kb = 0; *ke = m->geom[l].sz; }
}
static inline bool lvlSharded(const mp_mg* m, int l) { return m->slab && l >= 1 && l < m->nshard; }
// boundary planes of a sharded level array (levels >= 1: global-size arrays on every rank, each rank keeps its planes current)
template <typename Real> static int lvlHalo(mp_mg* m, int l, void* base) {
	const LvlGeom g = m->geom[l];
	return mp_dist_halo_range(m->ctx, base, (size_t)g.sx * g.sy * sizeof(Real), m->K0[l], m->K1[l], g.sz);
}

template <typename Real>
static int mgSmooth(mp_mg* m, int l, bool reversed, bool zeroX, const int* doneFlag, const L0<Real>& l0, bool skipLastHalo = false)
{
	mp_context* ctx = m->ctx; cudaStream_t st = ctx->stream;
	const LvlGeom g = m->geom[l];
	if (l == 0) {
		const dim3 gr = grid3((g.sx + 1) / 2, g.sy, g.sz, 128);
		for (int c = 0; c < 2; c++) {
			const int color = reversed ? 1 - c : c;
			if (l0vec<Real>(g)) {
				constexpr int V = 16 / (int)sizeof(Real);
				const int nvx = g.sx / V, kchunk = l0chunk(g);
				const int kb = m->slab ? m->k0 : 0, ke = m->slab ? m->k1 : g.sz;
				const dim3 grv((unsigned)((nvx + 31) / 32), (unsigned)((g.sy + 3) / 4), (unsigned)((ke - kb + kchunk - 1) / kchunk)), blk(32, 4, 1);
				if (l0masked<Real>(m)) {
					if (zeroX && c == 0) k_mg_l0_vecm<Real, V, 1><<<grv, blk, 0, st>>>(g, color, nvx, kchunk, kb, ke, (const Real*)m->A[0], m->mask0, l0.b, l0.bscale, l0.x, nullptr, doneFlag);
					else                 k_mg_l0_vecm<Real, V, 0><<<grv, blk, 0, st>>>(g, color, nvx, kchunk, kb, ke, (const Real*)m->A[0], m->mask0, l0.b, l0.bscale, l0.x, nullptr, doneFlag);
				}
				else if (zeroX && c == 0) k_mg_l0_vec<Real, V, 1><<<grv, blk, 0, st>>>(g, m->is3D, color, nvx, kchunk, kb, ke, (const Real*)m->A[0], l0.b, l0.bscale, m->type[0], l0.x, nullptr, doneFlag);
				else                 k_mg_l0_vec<Real, V, 0><<<grv, blk, 0, st>>>(g, m->is3D, color, nvx, kchunk, kb, ke, (const Real*)m->A[0], l0.b, l0.bscale, m->type[0], l0.x, nullptr, doneFlag);
				MP_CHECK_LAUNCH(ctx);
				// slab mode: the neighbours' boundary planes of this colour before the next colour (or the residual) reads them
				if (m->slab && !(skipLastHalo && c == 1)) MP_TRY(mp_dist_halo(ctx, l0.xLocal, (size_t)g.sx * g.sy * sizeof(Real), m->lsz));
				continue;
			}
			// with x == 0 on entry the first colour reduces to x = b / A0 (same arithmetic: the skipped products are exact zeros)
			if (zeroX && c == 0) k_mg_smooth0<Real, true><<<gr, 128, 0, st>>>(g, m->is3D, color, (const Real*)m->A[0], l0.b, l0.bscale, m->type[0], l0.x, doneFlag);
			else                 k_mg_smooth0<Real, false><<<gr, 128, 0, st>>>(g, m->is3D, color, (const Real*)m->A[0], l0.b, l0.bscale, m->type[0], l0.x, doneFlag);
			MP_CHECK_LAUNCH(ctx);
		}
	} else {
		const int ncol = m->is3D ? 8 : 4;
		const int hbx = (g.sx + 1) / 2;
		const int bsz = hbx >= 96 ? 128 : (hbx >= 48 ? 64 : 32);
		const dim3 gr = grid3(hbx, (g.sy + 1) / 2, (g.sz + 1) / 2, bsz);
		// small levels: the whole sweep in one launch of one CTA (MP_MG_SWEEP_CTA=0: one launch per colour)
		static const int sweepCta = getenv("MP_MG_SWEEP_CTA") ? atoi(getenv("MP_MG_SWEEP_CTA")) : 1;
		if (sweepCta && m->Afull[l] && m->is3D && g.n <= 6000 && !lvlSharded(m, l)) {
			const int hby = (g.sy + 1) / 2, hbz = (g.sz + 1) / 2;
			const unsigned char* rr = m->rowregOn[l] ? m->rowreg[l] : nullptr; const Real* cf = (const Real*)m->cstFull + 32 * l;
			k_mg_sweep_cta<Real><<<1, 512, 0, st>>>(g, hbx, hby, hbz, reversed ? 1 : 0, 0, (const Real*)m->Afull[l], (const Real*)m->b[l], m->type[l], (Real*)m->x[l], (Real*)m->r[l], doneFlag, rr, cf);
			MP_CHECK_LAUNCH(ctx);
			return MP_OK;
		}
		for (int c = 0; c < ncol; c++) {
			const int color = reversed ? ncol - 1 - c : c;
			if (m->Afull[l]) {
				const int hby = (g.sy + 1) / 2, hbz = m->is3D ? (g.sz + 1) / 2 : 1;
				int Kb, Ke; lvlRange(m, l, &Kb, &Ke);
				const int tz0 = m->is3D ? Kb / 2 : 0, ntz = m->is3D ? (Ke - 1) / 2 - tz0 + 1 : 1;
				const int bszF = hbx * hby >= 4096 ? 128 : (hbx * hby >= 512 ? 64 : 32);
				const unsigned char* rr = m->rowregOn[l] ? m->rowreg[l] : nullptr; const Real* cf = (const Real*)m->cstFull + 32 * l;
				const dim3 grz((unsigned)((hbx * hby + bszF - 1) / bszF), 1u, (unsigned)ntz);
				if (m->is3D) k_mg_sweep_full<Real, true, false><<<grz, bszF, 0, st>>>(g, hbx, hby, hbz, color, tz0, ntz, Kb, Ke, (const Real*)m->Afull[l], (const Real*)m->b[l], m->type[l], (Real*)m->x[l], nullptr, doneFlag, rr, cf);
				else         k_mg_sweep_full<Real, false, false><<<grz, bszF, 0, st>>>(g, hbx, hby, hbz, color, tz0, ntz, Kb, Ke, (const Real*)m->Afull[l], (const Real*)m->b[l], m->type[l], (Real*)m->x[l], nullptr, doneFlag, rr, cf);
				MP_CHECK_LAUNCH(ctx);
				if (lvlSharded(m, l)) MP_TRY(lvlHalo<Real>(m, l, m->x[l]));      // the next colour (or the residual / the interpolation) reads the neighbours' planes
				continue;
			}
			k_mg_smoothN<Real><<<gr, bsz, 0, st>>>(g, m->is3D, m->stencil, color & 1, (color >> 1) & 1, (color >> 2) & 1,
				(const Real*)m->A[l], (const Real*)m->b[l], m->type[l], (Real*)m->x[l], doneFlag);
			MP_CHECK_LAUNCH(ctx);
		}
	}
	return MP_OK;
}

template <typename Real>
static int mgResidual(mp_mg* m, int l, const int* doneFlag, const L0<Real>& l0)
{
	mp_context* ctx = m->ctx; cudaStream_t st = ctx->stream;
	const LvlGeom g = m->geom[l];
	if (l == 0 && l0vec<Real>(g)) {
		constexpr int V = 16 / (int)sizeof(Real);
		const int nvx = g.sx / V, kchunk = l0chunk(g);
		const int kb = m->slab ? m->k0 : 0, ke = m->slab ? m->k1 : g.sz;
		const dim3 grv((unsigned)((nvx + 31) / 32), (unsigned)((g.sy + 3) / 4), (unsigned)((ke - kb + kchunk - 1) / kchunk)), blk(32, 4, 1);
		if (l0masked<Real>(m)) k_mg_l0_vecm<Real, V, 2><<<grv, blk, 0, st>>>(g, 0, nvx, kchunk, kb, ke, (const Real*)m->A[0], m->mask0, l0.b, l0.bscale, l0.x, l0r<Real>(m), doneFlag);
		else k_mg_l0_vec<Real, V, 2><<<grv, blk, 0, st>>>(g, m->is3D, 0, nvx, kchunk, kb, ke, (const Real*)m->A[0], l0.b, l0.bscale, m->type[0], l0.x, l0r<Real>(m), doneFlag);
	}
	else if (l == 0) k_mg_residual0<Real><<<grid3(g.sx, g.sy, g.sz, 128), 128, 0, st>>>(g, m->is3D, (const Real*)m->A[0], l0.b, l0.bscale, m->type[0], (const Real*)l0.x, (Real*)m->r[0], doneFlag);
	else if (m->Afull[l]) {
		const int hbx = (g.sx + 1) / 2, hby = (g.sy + 1) / 2, hbz = m->is3D ? (g.sz + 1) / 2 : 1, ncol = m->is3D ? 8 : 4;
		const int bsz = hbx >= 96 ? 128 : (hbx >= 48 ? 64 : 32);
		int Kb, Ke; lvlRange(m, l, &Kb, &Ke);
		const int tz0 = m->is3D ? Kb / 2 : 0, ntz = m->is3D ? (Ke - 1) / 2 - tz0 + 1 : 1;
		const int bszF = hbx * hby >= 4096 ? 128 : (hbx * hby >= 512 ? 64 : 32);
		const unsigned char* rr = m->rowregOn[l] ? m->rowreg[l] : nullptr; const Real* cf = (const Real*)m->cstFull + 32 * l;
		const dim3 gr((unsigned)((hbx * hby + bszF - 1) / bszF), 1u, (unsigned)(ntz * ncol));
		if (m->is3D) k_mg_sweep_full<Real, true, true><<<gr, bszF, 0, st>>>(g, hbx, hby, hbz, 0, tz0, ntz, Kb, Ke, (const Real*)m->Afull[l], (const Real*)m->b[l], m->type[l], (Real*)m->x[l], (Real*)m->r[l], doneFlag, rr, cf);
		else         k_mg_sweep_full<Real, false, true><<<gr, bszF, 0, st>>>(g, hbx, hby, hbz, 0, tz0, ntz, Kb, Ke, (const Real*)m->Afull[l], (const Real*)m->b[l], m->type[l], (Real*)m->x[l], (Real*)m->r[l], doneFlag, rr, cf);
	}
	else        k_mg_residualN<Real><<<grid3(g.sx, g.sy, g.sz, g.sx >= 96 ? 128 : (g.sx >= 48 ? 64 : 32)), g.sx >= 96 ? 128 : (g.sx >= 48 ? 64 : 32), 0, st>>>(g, m->is3D, m->stencil, (const Real*)m->A[l], (const Real*)m->b[l], m->type[l], (const Real*)m->x[l], (Real*)m->r[l], doneFlag);
	MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

// doVCycle :448-504.  The level-0 iterate lives directly in `dst` (knCopyToGrid :499 becomes a no-op); xInit: dst already
// holds the initial guess (src) instead of zero.  rhsExt != NULL: use the caller's unscaled rhs (setRhs folded in).
template <typename Real>
static int mgVCycle(mp_mg* m, Real* dst, const Real* rhsExt, bool xInit, bool wantNorm, const int* doneFlag)
{
	mp_context* ctx = m->ctx; cudaStream_t st = ctx->stream;
	const int maxLevel = m->nlev - 1;
	L0<Real> l0; l0.x = slabView<Real>(m, dst); l0.xLocal = dst; l0.b = rhsExt ? slabView<Real>(m, (Real*)rhsExt) : (const Real*)m->b[0]; l0.bscale = rhsExt ? (Real)m->trivialScale : (Real)0;
	if (m->slab && !rhsExt) MP_FAIL(MP_ERR_UNSUPPORTED, "GridMg on z-slabs runs as the preconditioner of GridCg only (rhs folded in)");
	// knSet(x0, 0) :458.  (If a preconditioner call was skipped because the solve is done, dst simply keeps zeros.)
	const size_t n0 = m->slab ? (size_t)m->geom[0].sx * m->geom[0].sy * m->lsz : (size_t)m->geom[0].n;
	const bool fusedDown = maxLevel > 0 && !xInit && m->numPre == 1 && l0fused<Real>(m);      // zero iterate + both colours + residual in one pass (writes every x)
	const bool fusedUp = maxLevel > 0 && m->numPost == 1 && l0fused<Real>(m);
	if (!xInit && !fusedDown) MP_CUDA(cudaMemsetAsync(dst, 0, sizeof(Real) * n0, st));
	for (int l = 0; l < maxLevel; l++) {
		if (l == 0 && fusedDown) MP_TRY((l0fusedLaunch<Real, mgl0::MODE_DOWN>(m, 0, 0, l0.b, l0.bscale, nullptr, l0.x, (Real*)m->r[0], doneFlag)));
		else {
		// with x == 0 on entry the first colour of the first sweep reduces to x = b / A0 (the skipped products are exact zeros)
		for (int i = 0; i < m->numPre; i++) MP_TRY((mgSmooth<Real>(m, l, false, l == 0 && i == 0 && !xInit, doneFlag, l0)));
		MP_TRY((mgResidual<Real>(m, l, doneFlag, l0)));
		}
		const LvlGeom gf = m->geom[l], gc = m->geom[l + 1];
		if (l == 0 && l0vec<Real>(gf)) {
			constexpr int V = 16 / (int)sizeof(Real);
			const int ncx = (gc.sx + V / 2 - 1) / (V / 2);
			if (m->slab) {
				// the residual's ghost planes, then this rank's coarse planes K with fine plane 2K owned (the last rank also takes the planes
				// beyond the fine grid); the other planes of b1 stay zero and the sum over the ranks assembles b1 on every rank
				MP_TRY(mp_dist_halo(ctx, m->r[0], (size_t)gf.sx * gf.sy * sizeof(Real), m->lsz));
				const bool nextSharded = m->nshard > 1;
				if (!nextSharded) MP_CUDA(cudaMemsetAsync(m->b[1], 0, sizeof(Real) * (size_t)gc.n, st));
				MP_CUDA(cudaMemsetAsync(m->x[1], 0, sizeof(Real) * (size_t)gc.n, st));
				const int Kb = m->K0[1], Ke = m->K1[1];
				if (Ke > Kb) {
					const dim3 grs((unsigned)((ncx + 31) / 32), (unsigned)((gc.sy + 3) / 4), (unsigned)(Ke - Kb)), blk(32, 4, 1);
					k_mg_restrict_l0_vec<Real, V><<<grs, blk, 0, st>>>(gf, gc, ncx, Kb, m->type[0], m->type[1], (const Real*)l0r<Real>(m), (Real*)m->b[1], (Real*)m->x[1], doneFlag);
					MP_CHECK_LAUNCH(ctx);
				}
				// level 1 sharded: every rank needs b1 on its own planes only, which it has just computed; else assemble b1 everywhere
				if (!nextSharded) MP_TRY(mp_dist_allreduce_sum(ctx, m->b[1], (size_t)gc.n, (int)sizeof(Real)));
				continue;
			}
			const dim3 grv((unsigned)((ncx + 31) / 32), (unsigned)((gc.sy + 3) / 4), (unsigned)gc.sz), blk(32, 4, 1);
			k_mg_restrict_l0_vec<Real, V><<<grv, blk, 0, st>>>(gf, gc, ncx, 0, m->type[0], m->type[1], (const Real*)m->r[0], (Real*)m->b[1], (Real*)m->x[1], doneFlag);
			MP_CHECK_LAUNCH(ctx);
			continue;
		}
		if (lvlSharded(m, l)) {
			// sharded fine level: its residual's boundary planes, then this rank's coarse planes; a replicated coarse level is assembled by a sum
			MP_TRY(lvlHalo<Real>(m, l, m->r[l]));
			const bool nextSharded = l + 1 < m->nshard;
			if (!nextSharded) MP_CUDA(cudaMemsetAsync(m->b[l + 1], 0, sizeof(Real) * (size_t)gc.n, st));
			MP_CUDA(cudaMemsetAsync(m->x[l + 1], 0, sizeof(Real) * (size_t)gc.n, st));
			const int Kb = (m->K0[l] + 1) / 2, Ke = (m->K1[l] == gf.sz) ? gc.sz : (m->K1[l] + 1) / 2;
			if (Ke > Kb) {
				const int bs = gc.sx >= 96 ? 128 : (gc.sx >= 48 ? 64 : 32);
				k_mg_restrict<Real><<<grid3(gc.sx, gc.sy, Ke - Kb, bs), bs, 0, st>>>(gf, gc, Kb, m->type[l], m->type[l + 1], (const Real*)m->r[l], (Real*)m->b[l + 1], (Real*)m->x[l + 1], doneFlag);
				MP_CHECK_LAUNCH(ctx);
			}
			if (!nextSharded) MP_TRY(mp_dist_allreduce_sum(ctx, m->b[l + 1], (size_t)gc.n, (int)sizeof(Real)));
			continue;
		}
		k_mg_restrict<Real><<<grid3(gc.sx, gc.sy, gc.sz, gc.sx >= 96 ? 128 : (gc.sx >= 48 ? 64 : 32)), gc.sx >= 96 ? 128 : (gc.sx >= 48 ? 64 : 32), 0, st>>>(gf, gc, 0, m->type[l], m->type[l + 1], (const Real*)m->r[l], (Real*)m->b[l + 1], (Real*)m->x[l + 1], doneFlag);
		MP_CHECK_LAUNCH(ctx);
	}
	{
		const LvlGeom g = m->geom[maxLevel];
		const bool lvl0 = maxLevel == 0;
		static const int cgRows = getenv("MP_MG_COARSE_ROWS") ? atoi(getenv("MP_MG_COARSE_ROWS")) : 1;
		if (!lvl0 && cgRows && m->is3D && m->Afull[maxLevel] && g.n <= 1024)
			k_mg_coarse_cg_rows<Real><<<1, 1024, 0, st>>>(g, (g.sx + 1) / 2, (g.sy + 1) / 2, (g.sz + 1) / 2, (const Real*)m->Afull[maxLevel], (const Real*)m->b[maxLevel], m->type[maxLevel],
				(Real*)m->x[maxLevel], m->coarsestAcc, m->dFlags, doneFlag);
		else
		k_mg_coarse_cg<Real><<<1, 1024, 0, st>>>(g, m->is3D, m->stencil, lvl0 ? 1 : 0, (const Real*)m->A[maxLevel], lvl0 ? l0.b : (const Real*)m->b[maxLevel], m->type[maxLevel],
			lvl0 ? l0.x : (Real*)m->x[maxLevel], m->cg, m->coarsestAcc, m->dFlags, doneFlag, lvl0 ? l0.bscale : (Real)0);
		MP_CHECK_LAUNCH(ctx);
	}
	for (int l = maxLevel - 1; l >= 0; l--) {
		const LvlGeom gf = m->geom[l], gc = m->geom[l + 1];
		if (l == 0 && l0vec<Real>(gf)) {
			constexpr int V = 16 / (int)sizeof(Real);
			const int nvx = gf.sx / V, kchunk = l0chunk(gf);
			const int kb = m->slab ? m->k0 : 0, ke = m->slab ? m->k1 : gf.sz;
			const dim3 grv((unsigned)((nvx + 31) / 32), (unsigned)((gf.sy + 3) / 4), (unsigned)((ke - kb + kchunk - 1) / kchunk)), blk(32, 4, 1);
			// fused post-smoothing reads the corrected iterate from r0 (free since the restriction) and writes dst: neighbouring CTAs read each other's halos
			Real* xcorr = fusedUp ? (Real*)m->r[0] : l0.x;
			k_mg_interp_add_l0_vec<Real, V><<<grv, blk, 0, st>>>(gf, gc, nvx, kchunk, kb, ke, m->type[0], m->type[1], (const Real*)m->x[1], l0.x, xcorr, doneFlag);
			MP_CHECK_LAUNCH(ctx);
			if (fusedUp) { MP_TRY((l0fusedLaunch<Real, mgl0::MODE_SMOOTH>(m, 1, 0, l0.b, l0.bscale, xcorr, l0.x, nullptr, doneFlag))); continue; }
			if (m->slab) MP_TRY(mp_dist_halo(ctx, l0.xLocal, (size_t)gf.sx * gf.sy * sizeof(Real), m->lsz));      // the post-smoother reads the neighbours' corrected planes
		} else {
			int kb, ke; lvlRange(m, l, &kb, &ke);
			const int bs = gf.sx >= 96 ? 128 : (gf.sx >= 48 ? 64 : 32);
			k_mg_interp_add<Real><<<grid3(gf.sx, gf.sy, ke - kb, bs), bs, 0, st>>>(gf, gc, kb, m->type[l], m->type[l + 1], (const Real*)m->x[l + 1], (Real*)m->r[l], l == 0 ? l0.x : (Real*)m->x[l], doneFlag);
			MP_CHECK_LAUNCH(ctx);
			if (lvlSharded(m, l)) MP_TRY(lvlHalo<Real>(m, l, m->x[l]));
		}
		for (int i = 0; i < m->numPost; i++) MP_TRY((mgSmooth<Real>(m, l, true, false, doneFlag, l0, l == 0 && i == m->numPost - 1 && !wantNorm)));
	}
	if (wantNorm) MP_TRY((mgResidual<Real>(m, 0, doneFlag, l0)));      // calcResidual(0) only feeds the returned norm (:496-497)
	return MP_OK;
}

template <typename Real>
static int mgSetRhs(mp_mg* m, const Real* rhs, const int* doneFlag)
{
	if (!m->isASet) MP_FAIL(MP_ERR_NOT_SET, "GridMg::setRhs Error: A has not been set.");
	const int n = m->geom[0].n;
	k_mg_set_rhs<Real><<<nb(n, 256), 256, 0, m->ctx->stream>>>(n, (Real)m->trivialScale, rhs, m->type[0], (Real*)m->b[0], doneFlag);
	MP_CHECK_LAUNCH(m->ctx);
	m->isRhsSet = true;
	return MP_OK;
}

void mp_mg_invalidate(mp_mg* mg) { mg->isASet = false; mg->isRhsSet = false; mg->numPre = mg->numPost = 1; mg->coarsestAcc = (mg->prec == 4) ? (double)1E-8f : 1E-8; }
bool mp_mg_matches(const mp_mg* mg, int prec, int sx, int sy, int sz) {      // sz: the size of the caller's grids (the local slab in slab mode)
	return mg->prec == prec && mg->geom[0].sx == sx && mg->geom[0].sy == sy && (mg->slab ? mg->lsz : mg->geom[0].sz) == sz;
}

// InitPreconditionMultigrid conjugategrad.cpp:100-106
int mp_mg_precond_init(mp_mg* mg, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak, double accuracy)
{
	if (!mp_mg_matches(mg, A0->prec, A0->sx, A0->sy, A0->sz)) MP_FAIL(MP_ERR_INVALID, "GridMg preconditioner: the hierarchy was built for another grid size or precision");
	if (!mg->isASet) MP_TRY(mp_mg_set_a(mg, A0, Ai, Aj, Ak));
	// mAccuracy * 1E-4 is evaluated in double from a Real accuracy and narrowed to the Real member (multigrid.h:52)
	mg->coarsestAcc = (mg->prec == 4) ? (double)(float)((double)(float)accuracy * 1E-4) : accuracy * 1E-4;
	mg->numPre = 1; mg->numPost = 1;
	return MP_OK;
}
// ApplyPreconditionMultigrid conjugategrad.cpp:162-167
int mp_mg_precond_apply(mp_mg* mg, mp_grid* dst, const mp_grid* rhs, const int* doneFlag)
{
	if (!mg->isASet) MP_FAIL(MP_ERR_NOT_SET, "GridMg::setRhs Error: A has not been set.");
	if (!dst || !rhs || !mp_mg_matches(mg, dst->prec, dst->sx, dst->sy, dst->sz) || !mp_mg_matches(mg, rhs->prec, rhs->sx, rhs->sy, rhs->sz) || dst->kind != MP_GRID_REAL || rhs->kind != MP_GRID_REAL)
		MP_FAIL(MP_ERR_INVALID, "GridMg preconditioner: dst / rhs do not match the size or precision the hierarchy was built for");
	mg->isRhsSet = false;      // b0 is not materialised on this path (setRhs is folded into the level-0 kernels)
	if (mg->prec == 4) return mgVCycle<float>(mg, (float*)dst->d, (const float*)rhs->d, false, false, doneFlag);
	return mgVCycle<double>(mg, (double*)dst->d, (const double*)rhs->d, false, false, doneFlag);
}

extern "C" {

int mp_mg_create(mp_context* ctx, int prec, int sx, int sy, int sz, mp_mg** out)
{
	if (!ctx || !out) MP_FAIL(MP_ERR_INVALID, "mp_mg_create: NULL argument");
	if (prec != 4 && prec != 8) MP_FAIL(MP_ERR_INVALID, "mp_mg_create: prec must be 4 or 8");
	if ((long long)sx * sy * sz > 2000000000LL) MP_FAIL(MP_ERR_UNSUPPORTED, "mp_mg_create: more than 2^31 cells per GPU are not supported");
	MP_CUDA(cudaSetDevice(ctx->device));
	mp_mg* m = new mp_mg();
	memset(m, 0, sizeof *m);
	m->ctx = ctx; m->prec = prec;
	m->numPre = m->numPost = 1; m->coarsestAcc = (prec == 4) ? (double)1E-8f : 1E-8; m->trivialScale = (prec == 4) ? (double)1E-6f : 1E-6;
	// z-slab mode: `sz` is the local slab (owned planes + 2 ghost planes); the hierarchy is built for the global grid when the level-0
	// kernels can run on 16-byte vectors (rows aligned), else every rank coarsens its own slab (block-Jacobi, more iterations)
	{
		const DistState* ds = ctx->dist;
		const int wantGlobal = getenv("MP_MG_SLAB_GLOBAL") ? atoi(getenv("MP_MG_SLAB_GLOBAL")) : 1;
		if (ds && ds->active && ds->world > 1 && sz > 1 && sz == ds->k1 - ds->k0 + 2 && wantGlobal && sx % (16 / prec) == 0
		    && (long long)sx * sy * ds->gsz <= 2000000000LL) {
			m->slab = true; m->k0 = ds->k0; m->k1 = ds->k1; m->lsz = sz;
			sz = ds->gsz;
		}
	}
	const bool slabMode = m->slab;
	m->is3D = sz > 1; m->dim = m->is3D ? 3 : 2; m->stencil = m->is3D ? 14 : 5; m->stencil0 = m->is3D ? 4 : 3;
	// levels: size_l = (size_{l-1}+2)/2 until all dims <= 5 or n <= 1000 (multigrid.cpp:256-263)
	int l = 0; m->geom[0] = LvlGeom{ sx, sy, sz, sx * sy * sz };
	for (;;) {
		const LvlGeom g = m->geom[l];
		m->nlev = l + 1;
		if (l + 1 > 100 || l + 1 >= MG_MAXLVL) break;
		if (g.sx <= 5 && g.sy <= 5 && g.sz <= 5) break;
		if (g.n <= 1000) break;
		LvlGeom c; c.sx = (g.sx + 2) / 2; c.sy = (g.sy + 2) / 2; c.sz = (g.sz + 2) / 2; c.n = c.sx * c.sy * c.sz;
		m->geom[++l] = c;
	}
	m->nshard = 0;
	if (slabMode) {
		// plane ranges: coarse plane K belongs to the owner of fine plane 2K (the last rank also takes the planes beyond the fine grid).
		// A level is sharded while it is large (>= 4 M vertices) and every rank keeps >= 4 of its planes.
		const DistState* ds = ctx->dist;
		int maxShard = getenv("MP_MG_SHARD_LEVELS") ? atoi(getenv("MP_MG_SHARD_LEVELS")) : MG_MAXLVL;
		const long long minN = getenv("MP_MG_SHARD_MIN_N") ? atoll(getenv("MP_MG_SHARD_MIN_N")) : (4ll << 20);      // the parity check shards small levels too
		if (getenv("MP_MG_FULL") && !atoi(getenv("MP_MG_FULL"))) maxShard = 1;                                     // coarse levels are sharded through the colour-major rows only
		int minPlanes = m->geom[0].sz;
		for (int r = 0; r < ds->world; r++) { int a, b; mp_dist_slab(ds->gsz, r, ds->world, &a, &b); minPlanes = std::min(minPlanes, b - a); }
		m->K0[0] = m->k0; m->K1[0] = m->k1; m->nshard = 1;
		for (int q = 1; q < m->nlev; q++) {
			m->K0[q] = (m->K0[q - 1] + 1) / 2;
			m->K1[q] = (m->K1[q - 1] == m->geom[q - 1].sz) ? m->geom[q].sz : (m->K1[q - 1] + 1) / 2;
			minPlanes /= 2;
			if (m->nshard == q && q < maxShard && q < m->nlev - 1 && (long long)m->geom[q].n >= minN && minPlanes >= 4) m->nshard = q + 1;
		}
	}
	for (l = 0; l < m->nlev; l++) {
		const size_t n = (size_t)m->geom[l].n; const int S = l == 0 ? m->stencil0 : m->stencil;
		MP_CUDA(cudaMalloc(&m->A[l], n * S * prec)); MP_CUDA(cudaMalloc(&m->b[l], n * prec));
		if (l > 0) MP_CUDA(cudaMalloc(&m->x[l], n * prec));        // the level-0 iterate lives in the caller's dst grid
		const size_t nr = (l == 0 && m->slab) ? (size_t)sx * sy * m->lsz : n;      // slab mode: the level-0 residual is a slab array
		MP_CUDA(cudaMalloc(&m->r[l], nr * prec)); MP_CUDA(cudaMalloc((void**)&m->type[l], n));
		const int useFull = getenv("MP_MG_FULL") ? atoi(getenv("MP_MG_FULL")) : 1;
		if (l > 0 && useFull) {
			const LvlGeom g = m->geom[l];
			const size_t nc = (size_t)((g.sx + 1) / 2) * ((g.sy + 1) / 2) * (m->is3D ? (g.sz + 1) / 2 : 1);
			MP_CUDA(cudaMalloc(&m->Afull[l], nc * (m->is3D ? 8 * 27 : 4 * 9) * prec));
			if (m->is3D && g.n >= 4096) MP_CUDA(cudaMalloc((void**)&m->rowreg[l], nc * 8));
		}
		MP_CUDA(cudaMemsetAsync(m->A[l], 0, n * S * prec, ctx->stream)); if (l > 0) MP_CUDA(cudaMemsetAsync(m->x[l], 0, n * prec, ctx->stream));
		MP_CUDA(cudaMemsetAsync(m->b[l], 0, n * prec, ctx->stream)); MP_CUDA(cudaMemsetAsync(m->r[l], 0, nr * prec, ctx->stream));
		MP_CUDA(cudaMemsetAsync(m->type[l], 0, n, ctx->stream));
	}
	for (l = 1; l < m->nlev; l++) if (m->is3D && m->geom[l].n >= 4096) MP_CUDA(cudaMalloc((void**)&m->regular[l], (size_t)m->geom[l].n));
	MP_CUDA(cudaMalloc(&m->cst, 16 * sizeof(double))); MP_CUDA(cudaMalloc((void**)&m->first, sizeof(int)));
	MP_CUDA(cudaMalloc(&m->cstFull, 32 * MG_MAXLVL * sizeof(double))); MP_CUDA(cudaMemsetAsync(m->cstFull, 0, 32 * MG_MAXLVL * sizeof(double), ctx->stream));
	if (sx % (16 / prec) == 0) MP_CUDA(cudaMalloc((void**)&m->mask0, sizeof(unsigned short) * (size_t)m->geom[0].n + 64));
	MP_CUDA(cudaMalloc((void**)&m->cg, sizeof(double) * 4 * (size_t)m->geom[m->nlev - 1].n));
	MP_CUDA(cudaMemsetAsync(m->cg, 0, sizeof(double) * 4 * (size_t)m->geom[m->nlev - 1].n, ctx->stream));
	MP_CUDA(cudaMalloc((void**)&m->dFlags, 64 * sizeof(int)));
	MP_CUDA(cudaMemsetAsync(m->dFlags, 0, 64 * sizeof(int), ctx->stream));
	MP_CUDA(cudaHostAlloc((void**)&m->hFlags, 64 * sizeof(int), cudaHostAllocDefault));

	// coarsening paths for level 1 (multigrid.cpp:286-318), sorted by (sc, U) with generation order among equal keys
	static const int p7[7][3] = { {0,0,0}, {-1,0,0}, {1,0,0}, {0,-1,0}, {0,1,0}, {0,0,-1}, {0,0,1} };
	std::vector<CoarseningPath> paths;
	const int z0 = m->is3D ? 1 : 2, z1 = m->is3D ? 3 : 2;
	for (int uz = z0; uz <= z1; uz++) for (int uy = 1; uy <= 3; uy++) for (int ux = 1; ux <= 3; ux++)
		for (int i = 0; i < 1 + 2 * m->dim; i++) {
			const int wx = ux + p7[i][0], wy = uy + p7[i][1], wz = uz + p7[i][2];
			for (int nz = wz / 2; nz <= (wz + 1) / 2; nz++) for (int ny = wy / 2; ny <= (wy + 1) / 2; ny++) for (int nx = wx / 2; nx <= (wx + 1) / 2; nx++) {
				const int s = nx + 3 * ny + 9 * nz;
				if (s < 13) continue;
				CoarseningPath p;
				p.Nx = nx - 1; p.Ny = ny - 1; p.Nz = nz - 1; p.Ux = ux - 2; p.Uy = uy - 2; p.Uz = uz - 2; p.Wx = wx - 2; p.Wy = wy - 2; p.Wz = wz - 2;
				p.sc = s - 13; p.sf = (i + 1) / 2; p.inU = (i % 2 == 0);
				p.rw = 1.f / (float)(1 << ((ux % 2) + (uy % 2) + (uz % 2)));
				p.iw = 1.f / (float)(1 << ((wx % 2) + (wy % 2) + (wz % 2)));
				paths.push_back(p);
			}
		}
	std::stable_sort(paths.begin(), paths.end(), [](const CoarseningPath& a, const CoarseningPath& b) {
		if (a.sc != b.sc) return a.sc < b.sc;
		return (a.Ux + 1) + 3 * (a.Uy + 1) + 9 * (a.Uz + 1) < (b.Ux + 1) + 3 * (b.Uy + 1) + 9 * (b.Uz + 1); });
	m->npaths = (int)paths.size();
	for (int s = 0; s <= 14; s++) m->pathStart[s] = m->npaths;
	for (int q = m->npaths - 1; q >= 0; q--) m->pathStart[paths[q].sc] = q;
	for (int s = 13; s >= 0; s--) if (m->pathStart[s] > m->pathStart[s + 1]) m->pathStart[s] = m->pathStart[s + 1];
	MP_CUDA(cudaMalloc((void**)&m->dPaths, sizeof(CoarseningPath) * paths.size()));
	MP_CUDA(cudaMemcpy(m->dPaths, paths.data(), sizeof(CoarseningPath) * paths.size(), cudaMemcpyHostToDevice));
	MP_CUDA(cudaMemcpy(m->dFlags + 16, m->pathStart, sizeof(int) * 15, cudaMemcpyHostToDevice));
	*out = m; return MP_OK;
}

int mp_mg_destroy(mp_mg* m)
{
	if (!m) return MP_OK;
	cudaSetDevice(m->ctx->device);
	cudaStreamSynchronize(m->ctx->stream);
	for (int l = 0; l < m->nlev; l++) { cudaFree(m->A[l]); cudaFree(m->x[l]); cudaFree(m->b[l]); cudaFree(m->r[l]); cudaFree(m->type[l]); if (m->Afull[l]) cudaFree(m->Afull[l]); }
	cudaFree(m->cg); cudaFree(m->dPaths); cudaFree(m->dFlags); cudaFreeHost(m->hFlags); if (m->mask0) cudaFree(m->mask0);
	for (int l = 0; l < m->nlev; l++) if (m->regular[l]) cudaFree(m->regular[l]);
	for (int l = 0; l < m->nlev; l++) if (m->rowreg[l]) cudaFree(m->rowreg[l]);
	cudaFree(m->cst); cudaFree(m->first); cudaFree(m->cstFull);
	if (m->ctx->staticMg == m) m->ctx->staticMg = nullptr;
	if (m->ctx->spareMg == m) m->ctx->spareMg = nullptr;
	delete m; return MP_OK;
}

int mp_mg_set_a(mp_mg* m, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak)
{
	if (!m || !A0 || !Ai || !Aj || !Ak) MP_FAIL(MP_ERR_INVALID, "mp_mg_set_a: NULL argument");
	const mp_grid* gs[] = { A0, Ai, Aj, Ak };
	for (const mp_grid* g : gs) {
		if (g->kind != MP_GRID_REAL || g->prec != m->prec || g->sx != m->geom[0].sx || g->sy != m->geom[0].sy || g->sz != (m->slab ? m->lsz : m->geom[0].sz))
			MP_FAIL(MP_ERR_INVALID, "mp_mg_set_a: grid does not match the GridMg size/precision");
	}
	MP_CUDA(cudaSetDevice(m->ctx->device));
	return m->prec == 4 ? mgSetA<float>(m, A0, Ai, Aj, Ak) : mgSetA<double>(m, A0, Ai, Aj, Ak);
}

int mp_mg_set_rhs(mp_mg* m, const mp_grid* rhs)
{
	if (!m || !rhs) MP_FAIL(MP_ERR_INVALID, "mp_mg_set_rhs: NULL argument");
	if (m->slab) MP_FAIL(MP_ERR_UNSUPPORTED, "mp_mg_set_rhs: on z-slabs GridMg runs as the preconditioner of GridCg only");
	if (rhs->kind != MP_GRID_REAL || rhs->prec != m->prec || rhs->n != m->geom[0].n) MP_FAIL(MP_ERR_INVALID, "mp_mg_set_rhs: grid does not match the GridMg size/precision");
	return m->prec == 4 ? mgSetRhs<float>(m, (const float*)rhs->d, nullptr) : mgSetRhs<double>(m, (const double*)rhs->d, nullptr);
}

int mp_mg_is_a_set(const mp_mg* m, int* isSet) { *isSet = m->isASet ? 1 : 0; return MP_OK; }

int mp_mg_do_vcycle(mp_mg* m, mp_grid* dst, const mp_grid* src, double* resNorm)
{
	if (!m || !dst) MP_FAIL(MP_ERR_INVALID, "mp_mg_do_vcycle: NULL argument");
	if (!m->isASet || !m->isRhsSet) MP_FAIL(MP_ERR_NOT_SET, "GridMg::doVCycle Error: A and/or rhs have not been set.");   // :453
	if (m->slab) MP_FAIL(MP_ERR_UNSUPPORTED, "mp_mg_do_vcycle: on z-slabs GridMg runs as the preconditioner of GridCg only");
	if (dst->kind != MP_GRID_REAL || dst->prec != m->prec || dst->n != m->geom[0].n) MP_FAIL(MP_ERR_INVALID, "mp_mg_do_vcycle: dst does not match the GridMg size/precision");
	mp_context* ctx = m->ctx;
	MP_CUDA(cudaSetDevice(ctx->device));
	if (src) {
		if (src->kind != MP_GRID_REAL || src->prec != m->prec || src->n != m->geom[0].n) MP_FAIL(MP_ERR_INVALID, "mp_mg_do_vcycle: src does not match");
		if (src != dst) MP_CUDA(cudaMemcpyAsync(dst->d, src->d, (size_t)m->prec * m->geom[0].n, cudaMemcpyDeviceToDevice, ctx->stream));   // knCopyToVector :457
	}
	if (m->prec == 4) MP_TRY((mgVCycle<float>(m, (float*)dst->d, nullptr, src != nullptr, resNorm != nullptr, nullptr)));
	else              MP_TRY((mgVCycle<double>(m, (double*)dst->d, nullptr, src != nullptr, resNorm != nullptr, nullptr)));
	if (resNorm) {
		const int n = m->geom[0].n;
		unsigned int blocks = nb(n, 256 * 8); if (blocks > 1024) blocks = 1024;
		if (m->prec == 4) k_mg_norm<float><<<blocks, 256, 0, ctx->stream>>>(n, (const float*)m->r[0], m->type[0], ctx->partials, ctx->tickets + 6, ctx->dScal + 16);
		else              k_mg_norm<double><<<blocks, 256, 0, ctx->stream>>>(n, (const double*)m->r[0], m->type[0], ctx->partials, ctx->tickets + 6, ctx->dScal + 16);
		MP_CHECK_LAUNCH(ctx);
		MP_CUDA(cudaMemcpyAsync(ctx->hScal + 16, ctx->dScal + 16, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
		MP_CUDA(cudaStreamSynchronize(ctx->stream));
		*resNorm = std::sqrt(ctx->hScal[16]);
	}
	return MP_OK;
}

int mp_mg_set_coarsest_level_accuracy(mp_mg* m, double accuracy) { m->coarsestAcc = (m->prec == 4) ? (double)(float)accuracy : accuracy; return MP_OK; }
int mp_mg_set_smoothing(mp_mg* m, int numPre, int numPost) { m->numPre = numPre; m->numPost = numPost; return MP_OK; }
int mp_mg_num_levels(const mp_mg* m, int* levels) { *levels = m->nlev; return MP_OK; }
int mp_mg_level_info(const mp_mg* m, int level, int* sx, int* sy, int* sz, int* stencil)
{
	if (level < 0 || level >= m->nlev) MP_FAIL(MP_ERR_INVALID, "mp_mg_level_info: level %d out of range", level);
	if (sx) *sx = m->geom[level].sx; if (sy) *sy = m->geom[level].sy; if (sz) *sz = m->geom[level].sz;
	if (stencil) *stencil = level == 0 ? m->stencil0 : m->stencil;
	return MP_OK;
}

int mp_mg_level0_fused(const mp_mg* m, int* fused)
{
	if (!m || !fused) MP_FAIL(MP_ERR_INVALID, "mp_mg_level0_fused: NULL argument");
	*fused = (m->isASet && (m->prec == 4 ? l0fused<float>(m) : l0fused<double>(m)) && m->nlev > 1) ? 1 : 0;
	return MP_OK;
}

int mp_mg_download(const mp_mg* m, int level, const char* what, void* host)
{
	if (level < 0 || level >= m->nlev) MP_FAIL(MP_ERR_INVALID, "mp_mg_download: level %d out of range", level);
	MP_CUDA(cudaSetDevice(m->ctx->device));
	MP_CUDA(cudaStreamSynchronize(m->ctx->stream));
	const size_t n = (size_t)m->geom[level].n; const int S = level == 0 ? m->stencil0 : m->stencil;
	if (!strcmp(what, "type")) { MP_CUDA(cudaMemcpy(host, m->type[level], n, cudaMemcpyDeviceToHost)); return MP_OK; }
	if (!strcmp(what, "x")) {
		if (level == 0) MP_FAIL(MP_ERR_INVALID, "mp_mg_download: the level-0 iterate lives in the caller's dst grid");
		MP_CUDA(cudaMemcpy(host, m->x[level], n * m->prec, cudaMemcpyDeviceToHost)); return MP_OK; }
	if (!strcmp(what, "b")) { MP_CUDA(cudaMemcpy(host, m->b[level], n * m->prec, cudaMemcpyDeviceToHost)); return MP_OK; }
	if (!strcmp(what, "r")) { MP_CUDA(cudaMemcpy(host, m->r[level], n * m->prec, cudaMemcpyDeviceToHost)); return MP_OK; }
	if (!strcmp(what, "a")) {
		// device SoA [s][v] -> reference interleaved [v][s]
		std::vector<char> tmp(n * S * m->prec);
		MP_CUDA(cudaMemcpy(tmp.data(), m->A[level], n * S * m->prec, cudaMemcpyDeviceToHost));
		for (size_t v = 0; v < n; v++) for (int s = 0; s < S; s++)
			memcpy((char*)host + (v * S + s) * m->prec, tmp.data() + ((size_t)s * n + v) * m->prec, m->prec);
		return MP_OK;
	}
	if (!strcmp(what, "stats")) { ((int*)host)[0] = m->hostCoarsenLevels; MP_CUDA(cudaMemcpy((int*)host + 1, m->dFlags + 4, sizeof(int), cudaMemcpyDeviceToHost)); return MP_OK; }
	MP_FAIL(MP_ERR_INVALID, "mp_mg_download: unknown item '%s'", what);
}

} // extern "C"

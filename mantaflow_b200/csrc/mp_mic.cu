// Modified incomplete Cholesky MIC(0) preconditioner on the device.
//   InitPreconditionModifiedIncompCholesky2   conjugategrad.cpp:66-97   (serial k,j,i sweep in the reference)
//   ApplyPreconditionModifiedIncompCholesky2  conjugategrad.cpp:135-159 (forward + backward substitution, serial)
//
// The lexicographic sweeps are level-scheduled.  A cell depends on its -x/-y/-z (forward) or +x/+y/+z (backward)
// neighbours only, so any schedule that respects that partial order reproduces the serial result; the per-cell arithmetic
// here (operation order, -fmad=false, IEEE div/sqrt, the double-precision "+ 0." promotion of :85-89) is the reference's,
// hence Aprecond and z are BIT-IDENTICAL to the serial sweeps and PcMIC keeps the reference's iteration counts.
//
// v2 schedule (default): two-level wavefront.  The interior is cut into 8x8x8 tiles; tile (bi,bj,bk) depends on its three
// minus-neighbours, so all tiles of a tile-hyperplane bi+bj+bk = c run in one launch ((sx+sy+sz-6)/8 launches per sweep,
// 190 at 512^3, instead of sx+sy+sz-8 = 1528 cell-planes).  Inside a tile one CTA of 64 threads walks the 22 local
// hyperplanes li+lj+lk out of shared memory (thread = (lj,lk) line, skewed march along x, one __syncthreads per step;
// padded so that a warp's 32 accesses fall into 32 different banks).  Halo faces come from the already finished
// neighbour tiles through global memory.
// v1 (MP_MIC=1): one launch per cell hyperplane, kept for A/B timing.
// Both are dependency-/latency-bound rather than bandwidth-bound; DESIGN.md gives stage counts.
#include "mp_common.cuh"
#include <cstdlib>

// ---------------------------------------------------------------- shared per-cell arithmetic
template <typename Real>
__device__ __forceinline__ Real micFactor(Real a0, Real aix, Real ajx, Real akx, Real px, Real aiy, Real ajy, Real aky, Real py,
                                          Real aiz, Real ajz, Real akz, Real pz)
{   // (a?x = coefficient ? of the -x neighbour, p? = Aprecond of that neighbour)   conjugategrad.cpp:76-95
	const Real tau = (Real)0.97, sigma = (Real)0.25;
	const Real tx = aix * px, ty = ajy * py, tz = akz * pz;
	Real e = a0 - tx * tx - ty * ty - tz * tz;
	const Real inner = aix * (ajx + akx) * (px * px) + ajy * (aiy + aky) * (py * py) + akz * (aiz + ajz) * (pz * pz);
	e = (Real)((double)e - (double)tau * ((double)inner + 0.));     // the "+ 0." promotes bracket, product and subtraction to double
	if (e < sigma * a0) e = a0;
	return (Real)(1. / (double)sqrt(e));                              // sqrt in Real (std::sqrt(float)), divide in double (:95)
}

// ================================================================ v2: tiled two-level wavefront
#define T8 8
#define SJ 13               // smem strides: idx = lk*SK + lj*SJ + li ; (SJ-1) = 12 and (SK-1) = 129 == 1 (mod 32) make
#define SK 130              // 12*lj + lk (+ const) a bijection onto the 32 banks for the (lj 0..7, lk 0..3) lanes of a warp
#define SN (9 * SK)
#define SIDX(li, lj, lk) ((lk) * SK + (lj) * SJ + (li))

struct TileGeom { int sx, sy, sz; IndexInt Y, Z; int nbi, nbj, nbk; };

// tile of this CTA on tile-plane c, or false
__device__ __forceinline__ bool tileOf(const TileGeom& g, int c, int bklo, int& bi, int& bj, int& bk) {
	bj = blockIdx.x; bk = bklo + blockIdx.y; bi = c - bj - bk;
	return bi >= 0 && bi < g.nbi;
}

// cooperative load of a 9x9x9 window (8^3 tile + one halo layer) into smem.  Each of the 64 threads owns 12 window
// entries; their global / smem offsets are computed once (WinMap) and reused for every array, and the 12 loads of an
// array are issued back to back (fully unrolled) so that a tile costs ~one DRAM latency per array instead of twelve.
#define WIN_PER_THREAD 12
struct WinMap { IndexInt g[WIN_PER_THREAD]; int s[WIN_PER_THREAD]; };     // g < 0: outside the grid (reads as 0); s < 0: no entry
__device__ __forceinline__ void makeWinMap(const TileGeom& g, int ox, int oy, int oz, WinMap& m) {
	#pragma unroll
	for (int q = 0; q < WIN_PER_THREAD; q++) {
		const int e = threadIdx.x + 64 * q;
		const int li = e % 9, lj = (e / 9) % 9, lk = e / 81;
		const int gi = ox + li, gj = oy + lj, gk = oz + lk;
		m.s[q] = (e < 729) ? SIDX(li, lj, lk) : -1;
		m.g[q] = (e < 729 && gi < g.sx && gj < g.sy && gk < g.sz) ? (IndexInt)gi + g.Y * gj + g.Z * gk : (IndexInt)-1;
	}
}
template <typename Real>
__device__ __forceinline__ void loadWindow(const WinMap& m, const Real* __restrict__ a, Real* s) {
	Real v[WIN_PER_THREAD];
	#pragma unroll
	for (int q = 0; q < WIN_PER_THREAD; q++) v[q] = (m.g[q] >= 0) ? a[m.g[q]] : (Real)0;
	#pragma unroll
	for (int q = 0; q < WIN_PER_THREAD; q++) if (m.s[q] >= 0) s[m.s[q]] = v[q];
}

// MODE 0: init (Aprecond), 1: forward substitution, 2: backward substitution
template <typename Real, int MODE>
__global__ void __launch_bounds__(64) k_mic_tile(TileGeom g, int c, int bklo, const int* __restrict__ flags, Real* __restrict__ dst, const Real* __restrict__ src,
	Real* __restrict__ P, const Real* __restrict__ A0, const Real* __restrict__ Ai, const Real* __restrict__ Aj, const Real* __restrict__ Ak,
	const int* __restrict__ doneFlag)
{
	if (doneFlag && *doneFlag) return;
	int bi, bj, bk;
	if (!tileOf(g, c, bklo, bi, bj, bk)) return;
	__shared__ Real sP[SN], sAi[SN], sAj[SN], sAk[SN], sQ[SN];
	__shared__ unsigned char sF[512];
	// first interior cell of the tile; smem window origin: one cell lower for the minus-halo modes, the tile itself for backward
	const int x0 = 1 + T8 * bi, y0 = 1 + T8 * bj, z0 = 1 + T8 * bk;
	const int h = (MODE == 2) ? 0 : 1;                 // smem offset of tile-local cell 0
	const int ox = x0 - h, oy = y0 - h, oz = z0 - h;
	WinMap wm; makeWinMap(g, ox, oy, oz, wm);
	// interior cells of this thread: e = threadIdx.x + 64 q  ->  (li, lj, lk) = (e & 7, (e >> 3) & 7, e >> 6)
	int fl8[8]; Real r8[8];
	#pragma unroll
	for (int q = 0; q < 8; q++) {
		const int e = threadIdx.x + 64 * q;
		const int gi = x0 + (e & 7), gj = y0 + ((e >> 3) & 7), gk = z0 + (e >> 6);
		const bool in = gi <= g.sx - 2 && gj <= g.sy - 2 && gk <= g.sz - 2;
		const IndexInt idx = (IndexInt)gi + g.Y * gj + g.Z * gk;
		fl8[q] = in ? flags[idx] : 0;
		r8[q] = (MODE == 1 && in) ? src[idx] : (Real)0;
	}
	loadWindow<Real>(wm, Ai, sAi); loadWindow<Real>(wm, Aj, sAj); loadWindow<Real>(wm, Ak, sAk);
	loadWindow<Real>(wm, P, sP);
	if (MODE != 0) loadWindow<Real>(wm, dst, sQ);                 // halo faces = finished neighbours; interior = previous content of dst
	__syncthreads();                                              // the loop below overwrites interior slots of sQ
	#pragma unroll
	for (int q = 0; q < 8; q++) {
		const int e = threadIdx.x + 64 * q;
		const bool fl = (fl8[q] & TypeFluid) != 0;
		if (MODE == 1 && fl) sQ[SIDX((e & 7) + 1, ((e >> 3) & 7) + 1, (e >> 6) + 1)] = r8[q];   // forward: the slot holds r until it is replaced by q
		sF[e] = fl ? 1 : 0;
	}
	__syncthreads();
	const int lj = threadIdx.x & 7, lk = threadIdx.x >> 3;
	for (int step = 0; step < 3 * T8 - 2; step++) {
		// forward/init walk li+lj+lk upwards, backward walks it downwards
		const int li = (MODE == 2) ? (3 * T8 - 3 - step) - lj - lk : step - lj - lk;
		if (li >= 0 && li < T8 && sF[(lk << 6) | (lj << 3) | li]) {
			const int o = SIDX(li + h, lj + h, lk + h);
			if (MODE == 0) {
				const int ox_ = o - 1, oy_ = o - SJ, oz_ = o - SK;
				const Real a0 = A0[(IndexInt)(x0 + li) + g.Y * (y0 + lj) + g.Z * (z0 + lk)];
				sP[o] = micFactor<Real>(a0, sAi[ox_], sAj[ox_], sAk[ox_], sP[ox_], sAi[oy_], sAj[oy_], sAk[oy_], sP[oy_],
				                        sAi[oz_], sAj[oz_], sAk[oz_], sP[oz_]);
			} else if (MODE == 1) {
				const int ox_ = o - 1, oy_ = o - SJ, oz_ = o - SK;
				sQ[o] = sP[o] * (sQ[o] - sQ[ox_] * sAi[ox_] * sP[ox_] - sQ[oy_] * sAj[oy_] * sP[oy_] - sQ[oz_] * sAk[oz_] * sP[oz_]);
			} else {
				const Real p = sP[o];
				sQ[o] = p * (sQ[o] - sQ[o + 1] * sAi[o] * p - sQ[o + SJ] * sAj[o] * p - sQ[o + SK] * sAk[o] * p);
			}
		}
		__syncthreads();
	}
	for (int e = threadIdx.x; e < 512; e += 64) {
		if (!sF[e]) continue;                             // non-fluid cells keep their previous content (reference: `continue`)
		const int li = e & 7, lj2 = (e >> 3) & 7, lk2 = e >> 6;
		const IndexInt idx = (IndexInt)(x0 + li) + g.Y * (y0 + lj2) + g.Z * (z0 + lk2);
		const int o = SIDX(li + h, lj2 + h, lk2 + h);
		if (MODE == 0) P[idx] = sP[o]; else dst[idx] = sQ[o];
	}
}

// ---------------------------------------------------------------- v3: one launch per sweep, columns of tiles with progress flags
// CTA (bj,bk) walks its column of tiles bi = 0..nbi-1 (backward: everything mirrored).  Tile (bi,bj,bk) needs tile bi of the
// columns (bj-1,bk) and (bj,bk-1); each column publishes the number of tiles it has finished in prog[] (release/acquire
// through L2: __threadfence + atomic).  Predecessors always have a smaller linear block index, and blocks are dispatched in
// index order, so a waiting CTA never starves the CTA it waits for.  The coefficient windows of a tile are requested BEFORE
// the wait (they do not depend on the neighbours), only the halo of the running solution is loaded after it, bypassing L1
// (ld.global.cg) because the L1 may hold a stale copy of a sector that a neighbour column has rewritten since.
// A poll budget turns a scheduling surprise into an error flag instead of a hang.
template <typename Real>
__device__ __forceinline__ void loadWindowRegs(const WinMap& m, const Real* __restrict__ a, Real (&v)[WIN_PER_THREAD]) {
	#pragma unroll
	for (int q = 0; q < WIN_PER_THREAD; q++) v[q] = (m.g[q] >= 0) ? a[m.g[q]] : (Real)0;
}
template <typename Real>
__device__ __forceinline__ void loadWindowRegsCg(const WinMap& m, const Real* a, Real (&v)[WIN_PER_THREAD]) {
	#pragma unroll
	for (int q = 0; q < WIN_PER_THREAD; q++) v[q] = (m.g[q] >= 0) ? __ldcg(a + m.g[q]) : (Real)0;
}
template <typename Real>
__device__ __forceinline__ void storeWindow(const WinMap& m, const Real (&v)[WIN_PER_THREAD], Real* s) {
	#pragma unroll
	for (int q = 0; q < WIN_PER_THREAD; q++) if (m.s[q] >= 0) s[m.s[q]] = v[q];
}

template <typename Real, int MODE>
__global__ void __launch_bounds__(64, 8) k_mic_cols(TileGeom g, const int* __restrict__ flags, Real* dst, const Real* __restrict__ src,
	Real* P, const Real* __restrict__ A0, const Real* __restrict__ Ai, const Real* __restrict__ Aj, const Real* __restrict__ Ak,
	int* prog, int* stall, const int* __restrict__ doneFlag)
{
	if (doneFlag && *doneFlag) return;
	__shared__ Real sP[SN], sAi[SN], sAj[SN], sAk[SN], sQ[SN];
	__shared__ unsigned char sF[512];
	const bool bwd = (MODE == 2);
	const int bj = bwd ? g.nbj - 1 - (int)blockIdx.x : (int)blockIdx.x, bk = bwd ? g.nbk - 1 - (int)blockIdx.y : (int)blockIdx.y;
	const int self = bj + g.nbj * bk;
	// predecessor columns in sweep direction (-1: none)
	const int pj = bwd ? (bj + 1 < g.nbj ? self + 1 : -1) : (bj > 0 ? self - 1 : -1);
	const int pk = bwd ? (bk + 1 < g.nbk ? self + g.nbj : -1) : (bk > 0 ? self - g.nbj : -1);
	const int h = bwd ? 0 : 1;
	const int y0 = 1 + T8 * bj, z0 = 1 + T8 * bk;
	const int lj = threadIdx.x & 7, lk = threadIdx.x >> 3;
	for (int t = 0; t < g.nbi; t++) {
		const int bi = bwd ? g.nbi - 1 - t : t;
		const int x0 = 1 + T8 * bi;
		WinMap wm; makeWinMap(g, x0 - h, y0 - h, z0 - h, wm);
		// --- requests that do not depend on the neighbours
		int fl8[8]; Real r8[8];
		#pragma unroll
		for (int q = 0; q < 8; q++) {
			const int e = threadIdx.x + 64 * q;
			const int gi = x0 + (e & 7), gj = y0 + ((e >> 3) & 7), gk = z0 + (e >> 6);
			const bool in = gi <= g.sx - 2 && gj <= g.sy - 2 && gk <= g.sz - 2;
			const IndexInt idx = (IndexInt)gi + g.Y * gj + g.Z * gk;
			fl8[q] = in ? flags[idx] : 0;
			r8[q] = (MODE == 1 && in) ? src[idx] : (Real)0;
		}
		// (the shared arrays are free here: every thread is past the barrier that ended the previous tile)
		loadWindow<Real>(wm, Ai, sAi); loadWindow<Real>(wm, Aj, sAj); loadWindow<Real>(wm, Ak, sAk);
		if (MODE != 0) loadWindow<Real>(wm, (const Real*)P, sP);             // P is read-only during the apply sweeps
		// --- wait for the two neighbour columns to have finished their tile bi
		if (threadIdx.x == 0) {
			long long budget = 1ll << 26;
			if (pj >= 0) while (atomicAdd(prog + pj, 0) <= t && --budget > 0) __nanosleep(64);
			if (pk >= 0) while (atomicAdd(prog + pk, 0) <= t && --budget > 0) __nanosleep(64);
			if (budget <= 0) atomicExch(stall, 1);
			__threadfence();
		}
		__syncthreads();
		{
			Real vH[WIN_PER_THREAD];
			if (MODE == 0) { loadWindowRegsCg<Real>(wm, P, vH); storeWindow<Real>(wm, vH, sP); }
			else           { loadWindowRegsCg<Real>(wm, dst, vH); storeWindow<Real>(wm, vH, sQ); }
		}
		__syncthreads();
		#pragma unroll
		for (int q = 0; q < 8; q++) {
			const int e = threadIdx.x + 64 * q;
			const bool fl = (fl8[q] & TypeFluid) != 0;
			if (MODE == 1 && fl) sQ[SIDX((e & 7) + 1, ((e >> 3) & 7) + 1, (e >> 6) + 1)] = r8[q];
			sF[e] = fl ? 1 : 0;
		}
		__syncthreads();
		for (int step = 0; step < 3 * T8 - 2; step++) {
			const int li = bwd ? (3 * T8 - 3 - step) - lj - lk : step - lj - lk;
			if (li >= 0 && li < T8 && sF[(lk << 6) | (lj << 3) | li]) {
				const int o = SIDX(li + h, lj + h, lk + h);
				if (MODE == 0) {
					const int ox_ = o - 1, oy_ = o - SJ, oz_ = o - SK;
					const Real a0 = A0[(IndexInt)(x0 + li) + g.Y * (y0 + lj) + g.Z * (z0 + lk)];
					sP[o] = micFactor<Real>(a0, sAi[ox_], sAj[ox_], sAk[ox_], sP[ox_], sAi[oy_], sAj[oy_], sAk[oy_], sP[oy_],
					                        sAi[oz_], sAj[oz_], sAk[oz_], sP[oz_]);
				} else if (MODE == 1) {
					const int ox_ = o - 1, oy_ = o - SJ, oz_ = o - SK;
					sQ[o] = sP[o] * (sQ[o] - sQ[ox_] * sAi[ox_] * sP[ox_] - sQ[oy_] * sAj[oy_] * sP[oy_] - sQ[oz_] * sAk[oz_] * sP[oz_]);
				} else {
					const Real p = sP[o];
					sQ[o] = p * (sQ[o] - sQ[o + 1] * sAi[o] * p - sQ[o + SJ] * sAj[o] * p - sQ[o + SK] * sAk[o] * p);
				}
			}
			__syncthreads();
		}
		#pragma unroll
		for (int q = 0; q < 8; q++) {
			const int e = threadIdx.x + 64 * q;
			if (!sF[e]) continue;
			const IndexInt idx = (IndexInt)(x0 + (e & 7)) + g.Y * (y0 + ((e >> 3) & 7)) + g.Z * (z0 + (e >> 6));
			const int o = SIDX((e & 7) + h, ((e >> 3) & 7) + h, (e >> 6) + h);
			if (MODE == 0) P[idx] = sP[o]; else dst[idx] = sQ[o];
		}
		__threadfence();                                   // results visible at L2 before the progress counter moves
		__syncthreads();
		if (threadIdx.x == 0) atomicExch(prog + self, t + 1);
	}
}

template <typename Real, int MODE>
static int micColsSweep(mp_context* ctx, const Dims& d, const mp_grid* flags, Real* dst, const Real* src, Real* P, const Real* A0,
                        const Real* Ai, const Real* Aj, const Real* Ak, const int* doneFlag)
{
	TileGeom g = { d.sx, d.sy, d.sz, d.Y, d.Z, (d.sx - 2 + T8 - 1) / T8, (d.sy - 2 + T8 - 1) / T8, (d.sz - 2 + T8 - 1) / T8 };
	const size_t need = sizeof(int) * ((size_t)g.nbj * g.nbk + 1);
	if (ctx->micProgBytes < need) {
		if (ctx->micProg) { MP_CUDA(cudaStreamSynchronize(ctx->stream)); MP_CUDA(cudaFree(ctx->micProg)); ctx->micProg = nullptr; }
		MP_CUDA(cudaMalloc((void**)&ctx->micProg, need)); ctx->micProgBytes = need;
		MP_CUDA(cudaMemsetAsync(ctx->micProg, 0, need, ctx->stream));
	}
	int* stall = ctx->micProg + (size_t)g.nbj * g.nbk;
	MP_CUDA(cudaMemsetAsync(ctx->micProg, 0, sizeof(int) * (size_t)g.nbj * g.nbk, ctx->stream));      // progress counters, not the stall flag
	const dim3 grid((unsigned)g.nbj, (unsigned)g.nbk, 1);
	k_mic_cols<Real, MODE><<<grid, 64, 0, ctx->stream>>>(g, (const int*)flags->d, dst, src, P, A0, Ai, Aj, Ak, ctx->micProg, stall, doneFlag);
	MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

template <typename Real, int MODE>
static int micTiledSweep(mp_context* ctx, const Dims& d, const mp_grid* flags, Real* dst, const Real* src, Real* P, const Real* A0,
                         const Real* Ai, const Real* Aj, const Real* Ak, const int* doneFlag)
{
	TileGeom g = { d.sx, d.sy, d.sz, d.Y, d.Z, (d.sx - 2 + T8 - 1) / T8, (d.sy - 2 + T8 - 1) / T8, (d.sz - 2 + T8 - 1) / T8 };
	const int cmax = g.nbi + g.nbj + g.nbk - 3;
	for (int q = 0; q <= cmax; q++) {
		const int c = (MODE == 2) ? cmax - q : q;
		const int bklo = (c - (g.nbi - 1) - (g.nbj - 1)) > 0 ? (c - (g.nbi - 1) - (g.nbj - 1)) : 0;
		const int bkhi = c < g.nbk - 1 ? c : g.nbk - 1;
		if (bkhi < bklo) continue;
		const dim3 grid((unsigned)g.nbj, (unsigned)(bkhi - bklo + 1), 1);
		k_mic_tile<Real, MODE><<<grid, 64, 0, ctx->stream>>>(g, c, bklo, (const int*)flags->d, dst, src, P, A0, Ai, Aj, Ak, doneFlag);
		MP_CHECK_LAUNCH(ctx);
	}
	return MP_OK;
}

// ================================================================ v1: one launch per cell hyperplane
struct PlaneGeom { int sx, sy, sz; IndexInt Y, Z; };

// cell of plane c addressed by (j,k) = (1 + blockIdx.x*blockDim.x + threadIdx.x, klo + blockIdx.y); returns false if outside
__device__ __forceinline__ bool planeCell(const PlaneGeom& g, int c, int klo, int& i, int& j, int& k, IndexInt& idx) {
	j = 1 + blockIdx.x * blockDim.x + threadIdx.x; k = klo + blockIdx.y;
	i = c - j - k;
	if (j > g.sy - 2 || i < 1 || i > g.sx - 2) return false;
	idx = (IndexInt)i + g.Y * j + g.Z * k;
	return true;
}

template <typename Real>
__global__ void __launch_bounds__(128) k_mic_init_plane(PlaneGeom g, int c, int klo, const int* __restrict__ flags, Real* __restrict__ P,
	const Real* __restrict__ A0, const Real* __restrict__ Ai, const Real* __restrict__ Aj, const Real* __restrict__ Ak)
{
	int i, j, k; IndexInt idx;
	if (!planeCell(g, c, klo, i, j, k, idx)) return;
	if (!(flags[idx] & TypeFluid)) return;
	const IndexInt ix = idx - 1, iy = idx - g.Y, iz = idx - g.Z;
	P[idx] = micFactor<Real>(A0[idx], Ai[ix], Aj[ix], Ak[ix], P[ix], Ai[iy], Aj[iy], Ak[iy], P[iy], Ai[iz], Aj[iz], Ak[iz], P[iz]);
}

template <typename Real>
__global__ void __launch_bounds__(128) k_mic_fwd_plane(PlaneGeom g, int c, int klo, const int* __restrict__ flags, Real* __restrict__ dst, const Real* __restrict__ src,
	const Real* __restrict__ P, const Real* __restrict__ Ai, const Real* __restrict__ Aj, const Real* __restrict__ Ak, const int* __restrict__ doneFlag)
{
	if (doneFlag && *doneFlag) return;
	int i, j, k; IndexInt idx;
	if (!planeCell(g, c, klo, i, j, k, idx)) return;
	if (!(flags[idx] & TypeFluid)) return;
	const IndexInt ix = idx - 1, iy = idx - g.Y, iz = idx - g.Z;
	const Real p = P[idx];
	dst[idx] = p * (src[idx] - dst[ix] * Ai[ix] * P[ix] - dst[iy] * Aj[iy] * P[iy] - dst[iz] * Ak[iz] * P[iz]);
}

template <typename Real>
__global__ void __launch_bounds__(128) k_mic_bwd_plane(PlaneGeom g, int c, int klo, const int* __restrict__ flags, Real* __restrict__ dst,
	const Real* __restrict__ P, const Real* __restrict__ Ai, const Real* __restrict__ Aj, const Real* __restrict__ Ak, const int* __restrict__ doneFlag)
{
	if (doneFlag && *doneFlag) return;
	int i, j, k; IndexInt idx;
	if (!planeCell(g, c, klo, i, j, k, idx)) return;
	if (!(flags[idx] & TypeFluid)) return;
	const Real p = P[idx];
	dst[idx] = p * (dst[idx] - dst[idx + 1] * Ai[idx] * p - dst[idx + g.Y] * Aj[idx] * p - dst[idx + g.Z] * Ak[idx] * p);
}

struct PlaneLaunch { int c, klo; dim3 grid; };
static inline bool planeLaunch(const PlaneGeom& g, int c, PlaneLaunch& pl) {
	// interior cells only: i,j,k in [1, s-2]
	const int klo = (c - (g.sx - 2) - (g.sy - 2)) > 1 ? (c - (g.sx - 2) - (g.sy - 2)) : 1;
	const int khi = (c - 2) < (g.sz - 2) ? (c - 2) : (g.sz - 2);
	if (khi < klo) return false;
	pl.c = c; pl.klo = klo; pl.grid = dim3((unsigned)((g.sy - 2 + 127) / 128), (unsigned)(khi - klo + 1), 1);
	return true;
}

static inline int micVariant() { static const int v = getenv("MP_MIC") ? atoi(getenv("MP_MIC")) : 3; return v; }

int mp_mic_init_launch(mp_context* ctx, const mp_grid* flags, mp_grid* P, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak)
{
	const Dims d = dimsOf(flags);
	if (!d.is3D) MP_FAIL(MP_ERR_INVALID, "mICP only supports 3D grids so far");
	MP_CUDA(cudaMemsetAsync(P->d, 0, P->bytes, ctx->stream));          // Aprecond.clear() :71
	if (d.sx < 3 || d.sy < 3 || d.sz < 3) return MP_OK;
	if (micVariant() == 3) {
		if (P->prec == 4) return micColsSweep<float, 0>(ctx, d, flags, nullptr, nullptr, (float*)P->d, (const float*)A0->d, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, nullptr);
		return micColsSweep<double, 0>(ctx, d, flags, nullptr, nullptr, (double*)P->d, (const double*)A0->d, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, nullptr);
	}
	if (micVariant() == 2) {
		if (P->prec == 4) return micTiledSweep<float, 0>(ctx, d, flags, nullptr, nullptr, (float*)P->d, (const float*)A0->d, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, nullptr);
		return micTiledSweep<double, 0>(ctx, d, flags, nullptr, nullptr, (double*)P->d, (const double*)A0->d, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, nullptr);
	}
	PlaneGeom g = { d.sx, d.sy, d.sz, d.Y, d.Z };
	const int cmax = (d.sx - 2) + (d.sy - 2) + (d.sz - 2);
	for (int c = 3; c <= cmax; c++) {
		PlaneLaunch pl; if (!planeLaunch(g, c, pl)) continue;
		if (P->prec == 4) k_mic_init_plane<float><<<pl.grid, 128, 0, ctx->stream>>>(g, c, pl.klo, (const int*)flags->d, (float*)P->d, (const float*)A0->d, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d);
		else              k_mic_init_plane<double><<<pl.grid, 128, 0, ctx->stream>>>(g, c, pl.klo, (const int*)flags->d, (double*)P->d, (const double*)A0->d, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d);
		MP_CHECK_LAUNCH(ctx);
	}
	return MP_OK;
}

int mp_mic_apply_launch(mp_context* ctx, mp_grid* dst, const mp_grid* var1, const mp_grid* flags, const mp_grid* P,
                        const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak, const int* doneFlag)
{
	const Dims d = dimsOf(flags);
	if (!d.is3D) MP_FAIL(MP_ERR_INVALID, "mICP only supports 3D grids so far");
	if (d.sx < 3 || d.sy < 3 || d.sz < 3) return MP_OK;
	if (micVariant() == 3) {
		if (dst->prec == 4) {
			MP_TRY((micColsSweep<float, 1>(ctx, d, flags, (float*)dst->d, (const float*)var1->d, (float*)P->d, nullptr, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, doneFlag)));
			return micColsSweep<float, 2>(ctx, d, flags, (float*)dst->d, nullptr, (float*)P->d, nullptr, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, doneFlag);
		}
		MP_TRY((micColsSweep<double, 1>(ctx, d, flags, (double*)dst->d, (const double*)var1->d, (double*)P->d, nullptr, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, doneFlag)));
		return micColsSweep<double, 2>(ctx, d, flags, (double*)dst->d, nullptr, (double*)P->d, nullptr, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, doneFlag);
	}
	if (micVariant() == 2) {
		if (dst->prec == 4) {
			MP_TRY((micTiledSweep<float, 1>(ctx, d, flags, (float*)dst->d, (const float*)var1->d, (float*)P->d, nullptr, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, doneFlag)));
			return micTiledSweep<float, 2>(ctx, d, flags, (float*)dst->d, nullptr, (float*)P->d, nullptr, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, doneFlag);
		}
		MP_TRY((micTiledSweep<double, 1>(ctx, d, flags, (double*)dst->d, (const double*)var1->d, (double*)P->d, nullptr, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, doneFlag)));
		return micTiledSweep<double, 2>(ctx, d, flags, (double*)dst->d, nullptr, (double*)P->d, nullptr, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, doneFlag);
	}
	PlaneGeom g = { d.sx, d.sy, d.sz, d.Y, d.Z };
	const int cmax = (d.sx - 2) + (d.sy - 2) + (d.sz - 2);
	for (int c = 3; c <= cmax; c++) {
		PlaneLaunch pl; if (!planeLaunch(g, c, pl)) continue;
		if (dst->prec == 4) k_mic_fwd_plane<float><<<pl.grid, 128, 0, ctx->stream>>>(g, c, pl.klo, (const int*)flags->d, (float*)dst->d, (const float*)var1->d, (const float*)P->d, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, doneFlag);
		else                k_mic_fwd_plane<double><<<pl.grid, 128, 0, ctx->stream>>>(g, c, pl.klo, (const int*)flags->d, (double*)dst->d, (const double*)var1->d, (const double*)P->d, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, doneFlag);
		MP_CHECK_LAUNCH(ctx);
	}
	for (int c = cmax; c >= 3; c--) {
		PlaneLaunch pl; if (!planeLaunch(g, c, pl)) continue;
		if (dst->prec == 4) k_mic_bwd_plane<float><<<pl.grid, 128, 0, ctx->stream>>>(g, c, pl.klo, (const int*)flags->d, (float*)dst->d, (const float*)P->d, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, doneFlag);
		else                k_mic_bwd_plane<double><<<pl.grid, 128, 0, ctx->stream>>>(g, c, pl.klo, (const int*)flags->d, (double*)dst->d, (const double*)P->d, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, doneFlag);
		MP_CHECK_LAUNCH(ctx);
	}
	return MP_OK;
}

extern "C" {

int mp_mic_init(mp_context* ctx, const mp_grid* flags, mp_grid* Aprecond, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak)
{
	if (!ctx || !flags || !Aprecond || !A0 || !Ai || !Aj || !Ak) MP_FAIL(MP_ERR_INVALID, "mp_mic_init: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_mic_init: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, Aprecond, MP_GRID_REAL, "Aprecond", false));
	MP_TRY(mp_check_same(Aprecond, A0, MP_GRID_REAL, "A0", false)); MP_TRY(mp_check_same(Aprecond, Ai, MP_GRID_REAL, "Ai", false));
	MP_TRY(mp_check_same(Aprecond, Aj, MP_GRID_REAL, "Aj", false)); MP_TRY(mp_check_same(Aprecond, Ak, MP_GRID_REAL, "Ak", false));
	MP_CUDA(cudaSetDevice(ctx->device));
	MP_TRY(mp_check_flags_interior(ctx, flags));
	return mp_mic_init_launch(ctx, flags, Aprecond, A0, Ai, Aj, Ak);
}

int mp_mic_apply(mp_context* ctx, mp_grid* dst, const mp_grid* var1, const mp_grid* flags, const mp_grid* Aprecond,
                 const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak)
{
	if (!ctx || !flags || !Aprecond || !dst || !var1 || !Ai || !Aj || !Ak) MP_FAIL(MP_ERR_INVALID, "mp_mic_apply: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_mic_apply: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, dst, MP_GRID_REAL, "dst", false)); MP_TRY(mp_check_same(dst, var1, MP_GRID_REAL, "var1", false));
	MP_TRY(mp_check_same(dst, Aprecond, MP_GRID_REAL, "Aprecond", false)); MP_TRY(mp_check_same(dst, Ai, MP_GRID_REAL, "Ai", false));
	MP_TRY(mp_check_same(dst, Aj, MP_GRID_REAL, "Aj", false)); MP_TRY(mp_check_same(dst, Ak, MP_GRID_REAL, "Ak", false));
	(void)A0;
	if (dst == var1) MP_FAIL(MP_ERR_INVALID, "mp_mic_apply: dst must not alias var1");
	MP_CUDA(cudaSetDevice(ctx->device));
	MP_TRY(mp_check_flags_interior(ctx, flags));
	return mp_mic_apply_launch(ctx, dst, var1, flags, Aprecond, Ai, Aj, Ak, nullptr);
}

}

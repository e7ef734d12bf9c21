// Modified incomplete Cholesky MIC(0) preconditioner on the device.
//   InitPreconditionModifiedIncompCholesky2   conjugategrad.cpp:66-97   (serial k,j,i sweep in the reference)
//   ApplyPreconditionModifiedIncompCholesky2  conjugategrad.cpp:135-159 (forward + backward substitution, serial)
//
// The lexicographic sweeps are level-scheduled.  A cell depends on its -x/-y/-z (forward) or +x/+y/+z (backward)
// neighbours only, so any schedule that respects that partial order reproduces the serial result; the per-cell arithmetic
// here (operation order, -fmad=false, IEEE div/sqrt, the double-precision "+ 0." promotion of :85-89) is the reference's,
// hence Aprecond and z are BIT-IDENTICAL to the serial sweeps and PcMIC keeps the reference's iteration counts.
//
// Schedules (MP_MIC forces one; all are tested bit for bit against the serial sweep):
//  v1  one launch per cell hyperplane (kept for A/B timing);
//  v2  two-level wavefront: 8x8x8 tiles, all tiles of a tile-hyperplane in one launch ((sx+sy+sz-6)/8 launches per sweep); inside a
//      tile one CTA of 64 threads walks the 22 local hyperplanes out of padded, bank-conflict-free shared memory;
//  v3  ONE launch per sweep: CTA (bj,bk) walks its column of tiles and waits on progress counters of its predecessor columns
//      (the factor always uses this one; the sweeps on grids below ~100 MB);
//  v4  one WARP per 8x4 column of rows, cell-level wavefront by warp shuffles, cp.async ring, tagged mailboxes between columns
//      (the sweeps on large grids).
// All are dependency-/latency-bound rather than bandwidth-bound; DESIGN.md section 5 has the chain analysis.
#include "mp_common.cuh"
#include <cstdlib>
#include <vector>
#include <algorithm>

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

// A wait that ran out of its budget (a scheduling surprise) raises this flag instead of hanging the GPU; the host turns it into an error.
static int micStallFlag(mp_context* ctx) {
	if (!ctx->micStall) { MP_CUDA(cudaMalloc((void**)&ctx->micStall, sizeof(int))); MP_CUDA(cudaMemsetAsync(ctx->micStall, 0, sizeof(int), ctx->stream)); }
	return MP_OK;
}
int mp_mic_check_stall(mp_context* ctx) {     // synchronises the stream
	if (!ctx->micStall) return MP_OK;
	int v = 0;
	MP_CUDA(cudaMemcpyAsync(&v, ctx->micStall, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	if (v) { MP_CUDA(cudaMemsetAsync(ctx->micStall, 0, sizeof(int), ctx->stream)); MP_FAIL(MP_ERR_CUDA, "MIC sweep: a dependency wait timed out (results invalid)"); }
	return MP_OK;
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
	MP_TRY(micStallFlag(ctx)); int* stall = ctx->micStall;
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

// ================================================================ v4: one WARP per column, cell-level wavefront
// Lane (lj,lk) of a warp owns the grid row (j,k) = (8J+lj, 4K+lk).  At step t it works on cell i = t - (lj+lk) of its
// row: it trails its -y / -z neighbour lane by exactly ONE CELL, whose output it receives with a warp shuffle, and the
// x-recurrence stays in a register.  The forward sweep passes the PRODUCTS q*Aj*P, q*Ak*P (formed by the cell that owns the
// coefficients, in the reference's operation order), the backward sweep the raw q.  The dependent chain of the whole sweep
// is therefore ~sx + hops x (hand-off lag) steps of one shuffle and six flops -- the true depth of the lexicographic sweep --
// instead of one launch or one barrier per hyperplane.
//  * Memory is touched by chunks of CH cells (one 32-byte sector per lane and array: 8 floats / 4 doubles).  A ROUND is CH
//    steps; once per round, for all lanes at the same time, the chunk a lane will enter D rounds later is requested with
//    cp.async (16 B, L2 only) into the lane's ring in shared memory, and the chunk it finished is written back with
//    128-bit stores.  Inside a step the operands are single LDS from the ring ([array][lane][cell]: conflict-free because
//    lanes that share banks are on different cells).
//  * Warps hand their edge rows to the next column through MAILBOXES in global memory whose 16-byte pieces each carry the
//    sweep's sequence number next to the data (one 128-bit store / load, the "LL" idea of NCCL): no fence, no progress
//    counter.  The consumer prefetches the entry like any other operand and only if a piece still carries an old tag
//    re-reads it until the producer has written it.  Predecessor warps come earlier in the dispatch order (`order`, a
//    wavefront over the columns), so a waiting warp never starves the one it waits for.
//  * A per-chunk fluid bit mask (built once per solve from flags, interior cells only) replaces the 4 B/cell flag read;
//    chunks that contain non-fluid cells also load the old dst, whose non-fluid cells take part in the products exactly
//    as in the reference (they are never written).
// Per cell and sweep: 5 Real read + 1 written (24 B float / 48 B double) + ~4.5 B of mailbox traffic.
template <typename Real> struct ChunkOf { static constexpr int CH = 32 / (int)sizeof(Real), LOG = sizeof(Real) == 4 ? 3 : 2, NP = sizeof(Real) == 4 ? 3 : 4,
                                                                S = sizeof(Real) == 4 ? 5 : 6; };     // S ring slots: prefetch distance S-2 rounds
struct ColGeom { int sx, sy, sz; IndexInt Y, Z; int nch, nJ, nK, nRounds; };

// ring slot of one warp: 7 arrays [array][lane] x 32 B (0 P, 1 Ai, 2 Aj, 3 Ak, 4 RQ = rhs of the fluid cells / old dst of the others,
// 5 IY, 6 IZ = what the predecessor columns hand to the edge lanes) and the chunk's fluid mask per lane
#define MW_ARR(a) ((a) * 1024)
#define MW_FM 7168
#define MW_SLOT_BYTES 7296

__device__ __forceinline__ void cpAsync16(unsigned int sdst, const void* gsrc) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sdst), "l"(gsrc) : "memory"); }
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
// one 128-bit access (STG.E.128.STRONG.GPU / LDG.E.128.STRONG.GPU): data and tag of a mailbox piece travel together
__device__ __forceinline__ void stPiece(uint4* p, uint4 v) { asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ uint4 ldPiece(const uint4* p) { uint4 v; asm volatile("ld.relaxed.gpu.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint4 ldsVec(unsigned int sa) { uint4 v; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa) : "memory"); return v; }
__device__ __forceinline__ void stsVec(unsigned int sa, uint4 v) { asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(sa), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ unsigned int ldsU32(unsigned int sa) { unsigned int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(sa) : "memory"); return v; }
__device__ __forceinline__ void stsU32(unsigned int sa, unsigned int v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(sa), "r"(v) : "memory"); }
__device__ __forceinline__ float ldsReal(unsigned int sa, float) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sa) : "memory"); return v; }
__device__ __forceinline__ double ldsReal(unsigned int sa, double) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa) : "memory"); return v; }
__device__ __forceinline__ void stsReal(unsigned int sa, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(sa), "f"(v) : "memory"); }
__device__ __forceinline__ void stsReal(unsigned int sa, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(sa), "d"(v) : "memory"); }

// every word of a piece is used (data or tag): a dead word would let the compiler reuse a register that a prefetch in flight still writes
__device__ __forceinline__ bool pieceOk(const uint4 (&pc)[3], unsigned int tag) {
	return pc[0].w == tag && pc[1].w == tag && pc[2].z == tag && pc[2].w == tag;
}
__device__ __forceinline__ bool pieceOk(const uint4 (&pc)[4], unsigned int tag) {
	bool ok = true;
	#pragma unroll
	for (int q = 0; q < 4; q++) ok = ok && pc[q].z == tag && pc[q].w == tag;
	return ok;
}
// the 32 bytes of a chunk (two 16-byte halves h0 h1) <-> tagged mailbox pieces
__device__ __forceinline__ void packPieces(uint4 h0, uint4 h1, unsigned int tag, uint4 (&pc)[3]) {    // float: 3+3+2 values
	pc[0] = make_uint4(h0.x, h0.y, h0.z, tag); pc[1] = make_uint4(h0.w, h1.x, h1.y, tag); pc[2] = make_uint4(h1.z, h1.w, tag, tag);
}
__device__ __forceinline__ void packPieces(uint4 h0, uint4 h1, unsigned int tag, uint4 (&pc)[4]) {    // double: one value per piece
	pc[0] = make_uint4(h0.x, h0.y, tag, tag); pc[1] = make_uint4(h0.z, h0.w, tag, tag); pc[2] = make_uint4(h1.x, h1.y, tag, tag); pc[3] = make_uint4(h1.z, h1.w, tag, tag);
}
__device__ __forceinline__ void unpackPieces(const uint4 (&pc)[3], uint4& h0, uint4& h1) {
	h0 = make_uint4(pc[0].x, pc[0].y, pc[0].z, pc[1].x); h1 = make_uint4(pc[1].y, pc[1].z, pc[2].x, pc[2].y);
}
__device__ __forceinline__ void unpackPieces(const uint4 (&pc)[4], uint4& h0, uint4& h1) {
	h0 = make_uint4(pc[0].x, pc[0].y, pc[1].x, pc[1].y); h1 = make_uint4(pc[2].x, pc[2].y, pc[3].x, pc[3].y);
}

// request chunk [i0, i0+CH) of the row a points to into the 32 bytes at shared address sa (VEC: rows are whole, aligned chunks)
template <typename Real, int CH, bool VEC>
__device__ __forceinline__ void fetchChunk(const Real* a, int i0, int sx, unsigned int sa) {
	if (VEC) { cpAsync16(sa, a + i0); cpAsync16(sa + 16, a + i0 + CH / 2); }
	else {   // ragged or unaligned rows: plain loads (through L2), zero beyond the row
		#pragma unroll
		for (int u = 0; u < CH; u++) stsReal(sa + u * (unsigned int)sizeof(Real), (i0 + u < sx) ? __ldcg(a + i0 + u) : (Real)0);
	}
}
// write the cells of `mask` of a finished chunk (32 bytes at shared address sa) to the row
template <typename Real, int CH, bool VEC>
__device__ __forceinline__ void storeChunk(Real* a, int i0, unsigned int mask, unsigned int sa) {
	if (VEC && mask == (1u << CH) - 1u) {
		uint4* o = (uint4*)(a + i0);
		o[0] = ldsVec(sa); o[1] = ldsVec(sa + 16);
	} else {
		#pragma unroll
		for (int u = 0; u < CH; u++) if ((mask >> u) & 1u) a[i0 + u] = ldsReal(sa + u * (unsigned int)sizeof(Real), Real());
	}
}

// one thread per (chunk, row): bit u = cell CH*c+u is a fluid cell of the interior
template <int CH>
__global__ void __launch_bounds__(256) k_mic_fluid_mask(ColGeom g, const int* __restrict__ flags, unsigned char* __restrict__ fmask) {
	const IndexInt n = (IndexInt)g.nch * g.sy * g.sz;
	for (IndexInt t = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (IndexInt)gridDim.x * blockDim.x) {
		const int c = (int)(t % g.nch); const IndexInt rw = t / g.nch; const int j = (int)(rw % g.sy), k = (int)(rw / g.sy);
		unsigned int m = 0;
		if (j >= 1 && j <= g.sy - 2 && k >= 1 && k <= g.sz - 2) {
			#pragma unroll
			for (int u = 0; u < CH; u++) { const int i = CH * c + u; if (i >= 1 && i <= g.sx - 2 && (flags[(IndexInt)i + g.Y * j + g.Z * k] & TypeFluid)) m |= 1u << u; }
		}
		fmask[t] = (unsigned char)m;
	}
}

template <typename Real, int MODE, bool VEC>    // MODE 1: forward substitution, 2: backward substitution
__global__ void __launch_bounds__(32) k_mic_warp(ColGeom g, const unsigned char* __restrict__ fmask, Real* dst, const Real* __restrict__ src,
	const Real* __restrict__ P, const Real* __restrict__ Ai, const Real* __restrict__ Aj, const Real* __restrict__ Ak,
	uint4* mailY, uint4* mailZ, unsigned int tag, int* stall, const int* __restrict__ doneFlag, const int* __restrict__ order)
{
	constexpr int CH = ChunkOf<Real>::CH, LOG = ChunkOf<Real>::LOG, NP = ChunkOf<Real>::NP, S = ChunkOf<Real>::S, D = S - 2;
	constexpr unsigned int FULL = 0xffffffffu, FULLM = (1u << CH) - 1u, RB = (unsigned int)sizeof(Real);
	constexpr bool bwd = (MODE == 2);
	extern __shared__ __align__(16) unsigned char ring[];
	if (doneFlag && *doneFlag) return;
	const int lane = threadIdx.x, lj = lane & 7, lk = lane >> 3;
	const int col = order[blockIdx.x];
	const int J = bwd ? g.nJ - 1 - col % g.nJ : col % g.nJ, K = bwd ? g.nK - 1 - col / g.nJ : col / g.nJ;
	const int j = 8 * J + lj, k = 4 * K + lk;
	const bool rowIn = j < g.sy && k < g.sz;
	const IndexInt row = rowIn ? g.Y * j + g.Z * k : 0;
	const unsigned char* fmRow = fmask + (IndexInt)g.nch * (rowIn ? (IndexInt)j + (IndexInt)g.sy * k : 0);
	const Real *rP = P + row, *rAi = Ai + row, *rAj = Aj + row, *rAk = Ak + row, *rSrc = bwd ? nullptr : src + row;
	Real* rDst = dst + row;
	const int skew = bwd ? (7 - lj) + (3 - lk) : lj + lk;
	// A lane walks its row in processing order: chunk c (0..nch-1) is the grid chunk c (forward) or nch-1-c (backward).
	// In round T it finishes chunk A = T - kA and, unless skew is a multiple of CH, starts chunk A+1; position u0 of chunk A at step 0.
	const int kA = (skew + CH - 1) >> LOG, kE = skew >> LOG, u0 = (-skew) & (CH - 1);
	// mailboxes: entry (column, row-in-plane, grid chunk), 4 pieces of 16 B.  This lane reads the predecessor column's entry of its
	// row (lanes on the warp's leading edge) and writes its own column's entry (lanes on the trailing edge, if a column follows).
	const int Jp = bwd ? J + 1 : J - 1, Kp = bwd ? K + 1 : K - 1;
	const bool edgeY = bwd ? lj == 7 : lj == 0, edgeZ = bwd ? lk == 3 : lk == 0;
	const bool rdY = rowIn && edgeY && Jp >= 0 && Jp < g.nJ, rdZ = rowIn && edgeZ && Kp >= 0 && Kp < g.nK;
	const bool wrY = rowIn && (bwd ? (lj == 0 && J > 0) : (lj == 7 && J + 1 < g.nJ)), wrZ = rowIn && (bwd ? (lk == 0 && K > 0) : (lk == 3 && K + 1 < g.nK));
	const uint4* inY = mailY + ((IndexInt)(rdY ? Jp : 0) * g.sz + (rowIn ? k : 0)) * g.nch * 4;
	const uint4* inZ = mailZ + ((IndexInt)(rdZ ? Kp : 0) * g.sy + (rowIn ? j : 0)) * g.nch * 4;
	uint4* outY = mailY + ((IndexInt)J * g.sz + (rowIn ? k : 0)) * g.nch * 4;
	uint4* outZ = mailZ + ((IndexInt)K * g.sy + (rowIn ? j : 0)) * g.nch * 4;
	const unsigned int ringS = (unsigned int)__cvta_generic_to_shared(ring);
	const unsigned int laneOff = lane * 32;

	// virtual chunks (before the row starts, rows outside the grid) read zeros: clear the ring once
	for (int q = lane; q < S * MW_SLOT_BYTES / 16; q += 32) stsVec(ringS + q * 16, make_uint4(0u, 0u, 0u, 0u));
	__syncwarp();

	// ---- loop state, kept incremental so that the once-per-round bookkeeping stays short
	const unsigned int nchL = rowIn ? (unsigned int)g.nch : 0u;       // chunk c of this lane exists iff (unsigned)c < nchL
	const int sx = g.sx, nchm1 = g.nch - 1;
	const int dAE = kA - kE;                                          // 0: the chunk entered in a round is chunk A, 1: it is chunk A+1
	constexpr unsigned int RING = (unsigned int)S * MW_SLOT_BYTES;
	auto nextSlot = [&](unsigned int o) -> unsigned int { o += MW_SLOT_BYTES; return o == RING ? 0u : o; };
	unsigned int oA = (unsigned int)((-D - kA + 4 * S) % S) * MW_SLOT_BYTES;   // ring offset of chunk A of round T = -D
	unsigned int oReq = (unsigned int)((-kE + 4 * S) % S) * MW_SLOT_BYTES;     // ... of the chunk requested in that round
	const unsigned int laneS = ringS + laneOff;

	// mailbox entry of a chunk: requested into registers one round before the chunk is entered (pc); valid if the producer was done by
	// then, else re-read until it is; then laid out like an operand array (IY / IZ) for the steps
	auto receive = [&](uint4 (&pc)[NP], const uint4* entry, unsigned int dstArr) {
		bool ok = pieceOk(pc, tag);
		if (!ok) {
			int budget = 1 << 22;
			do {
				#pragma unroll
				for (int q = 0; q < NP; q++) pc[q] = ldPiece(entry + q);
				ok = pieceOk(pc, tag);
				if (!ok) __nanosleep(128);                 // 12 edge lanes per warp poll: back off instead of flooding L2
			} while (!ok && --budget > 0);
			if (!ok) atomicExch(stall, 1);
		}
		uint4 h0, h1; unpackPieces(pc, h0, h1);
		stsVec(dstArr, h0); stsVec(dstArr + 16, h1);
	};

	uint4 pmY[NP], pmZ[NP];
	#pragma unroll
	for (int q = 0; q < NP; q++) { pmY[q] = make_uint4(0u, 0u, 0u, 0u); pmZ[q] = make_uint4(0u, 0u, 0u, 0u); }
	// step s of a round works on position u0+s (processing order): in chunk A below CH, in chunk A+1 from CH on -- both fixed per lane
	unsigned int offS[CH]; unsigned int crossM = 0u;
	#pragma unroll
	for (int s = 0; s < CH; s++) {
		const int pos = u0 + s, um = pos & (CH - 1);
		offS[s] = (unsigned int)(bwd ? CH - 1 - um : um) * RB;
		if (pos >= CH) crossM |= 1u << s;
	}
	Real tx = (Real)0, oy = (Real)0, oz = (Real)0;     // carried: x-recurrence (own previous cell), outputs for the neighbour lanes
	unsigned int fmReq = ((unsigned int)(-kE) < nchL) ? (unsigned int)fmRow[bwd ? nchm1 + kE : -kE] : 0u;   // fluid mask of the chunk the next request is for
	#pragma unroll 1
	for (int T = -D; T < g.nRounds; T++) {
		const int cE = T - kE, cA = cE - dAE, cReq = cE + D;
		// ---- once per round, all lanes together: request the chunk entered D rounds from now, look one more ahead for its mask
		if ((unsigned int)cReq < nchL) {
			const int i0 = CH * (bwd ? nchm1 - cReq : cReq);
			const unsigned int sl = laneS + oReq;
			fetchChunk<Real, CH, VEC>(rP, i0, sx, sl + MW_ARR(0));
			fetchChunk<Real, CH, VEC>(rAi, i0, sx, sl + MW_ARR(1));
			fetchChunk<Real, CH, VEC>(rAj, i0, sx, sl + MW_ARR(2));
			fetchChunk<Real, CH, VEC>(rAk, i0, sx, sl + MW_ARR(3));
			fetchChunk<Real, CH, VEC>((bwd || fmReq == 0u) ? (const Real*)rDst : rSrc, i0, sx, sl + MW_ARR(4));   // own row of dst: nobody else writes it
			stsU32(ringS + oReq + MW_FM + lane * 4, fmReq);
		}
		cpAsyncCommit();                                   // one group per round, empty or not
		fmReq = ((unsigned int)(cReq + 1) < nchL) ? (unsigned int)fmRow[bwd ? nchm1 - (cReq + 1) : cReq + 1] : 0u;
		oReq = nextSlot(oReq);
		if (T < 0) { oA = nextSlot(oA); continue; }        // (prologue: fills the pipeline)
		cpAsyncWait<D>();                                  // the chunk entered in this round has landed
		const unsigned int oB = nextSlot(oA), oE = dAE ? oB : oA;
		const bool inE = (unsigned int)cE < nchL, inA = (unsigned int)cA < nchL, inB = (unsigned int)(cA + 1) < nchL;
		const int gE = bwd ? nchm1 - cE : cE;
		if (inE) {
			if (rdY) receive(pmY, inY + 4 * gE, laneS + oE + MW_ARR(5));
			if (rdZ) receive(pmZ, inZ + 4 * gE, laneS + oE + MW_ARR(6));
			if (!bwd) {    // a chunk with fluid and non-fluid cells: the latter take part with their old dst value (rare: fetched here)
				const unsigned int fmE = ldsU32(ringS + oE + MW_FM + lane * 4);
				if (fmE != 0u && fmE != FULLM) {
					#pragma unroll
					for (int u = 0; u < CH; u++)
						if (!((fmE >> u) & 1u)) stsReal(laneS + oE + MW_ARR(4) + u * RB, (CH * gE + u < sx) ? __ldcg(rDst + CH * gE + u) : (Real)0);
				}
			}
		}
		{	// next round's mailbox entries: in flight during this round's steps (always loaded, from a valid address, straight into pm*)
			const bool inN = (unsigned int)(cE + 1) < nchL;
			const int gN = inN ? (bwd ? nchm1 - (cE + 1) : cE + 1) : 0;
			const uint4* aY = (rdY ? inY : mailY) + 4 * gN; const uint4* aZ = (rdZ ? inZ : mailZ) + 4 * gN;
			#pragma unroll
			for (int q = 0; q < NP; q++) { pmY[q] = __ldcg(aY + q); pmZ[q] = __ldcg(aZ + q); }     // weak, through L2; the tags say whether it was in time
		}
		const unsigned int slA = laneS + oA, slB = laneS + oB;
		const unsigned int fmA = inA ? ldsU32(ringS + oA + MW_FM + lane * 4) : 0u, fmB = inB ? ldsU32(ringS + oB + MW_FM + lane * 4) : 0u;
		// fluid bits of this round's CH steps, in processing order
		const unsigned int pA = bwd ? (__brev(fmA) >> (32 - CH)) : fmA, pB = bwd ? (__brev(fmB) >> (32 - CH)) : fmB;
		const unsigned int stepFluid = ((pA | (pB << CH)) >> u0) & FULLM;
		// operands of this round's CH cells -> registers (position u0+s in processing order: chunk A below CH, chunk A+1 from CH on),
		// then the dependent chain runs on registers and shuffles only, then the results go back to the ring
		unsigned int ad[CH];
		Real p_[CH], ai_[CH], aj_[CH], ak_[CH], q_[CH], ey_[CH], ez_[CH];
		#pragma unroll
		for (int s = 0; s < CH; s++) {
			ad[s] = (((crossM >> s) & 1u) ? slB : slA) + offS[s];
			p_[s] = ldsReal(ad[s] + MW_ARR(0), Real()); ai_[s] = ldsReal(ad[s] + MW_ARR(1), Real());
			aj_[s] = ldsReal(ad[s] + MW_ARR(2), Real()); ak_[s] = ldsReal(ad[s] + MW_ARR(3), Real());
			q_[s] = ldsReal(ad[s] + MW_ARR(4), Real());        // fluid cell: the rhs (forward) / the forward result (backward); else the old content, which stays
			ey_[s] = edgeY ? ldsReal(ad[s] + MW_ARR(5), Real()) : (Real)0;
			ez_[s] = edgeZ ? ldsReal(ad[s] + MW_ARR(6), Real()) : (Real)0;
		}
		Real oy_[CH], oz_[CH];
		#pragma unroll
		for (int s = 0; s < CH; s++) {
			Real iy = bwd ? __shfl_down_sync(FULL, oy, 1) : __shfl_up_sync(FULL, oy, 1);
			Real iz = bwd ? __shfl_down_sync(FULL, oz, 8) : __shfl_up_sync(FULL, oz, 8);
			if (edgeY) iy = ey_[s];
			if (edgeZ) iz = ez_[s];
			const bool fl = (stepFluid >> s) & 1u;
			const Real p = p_[s];
			if (!bwd) {
				const Real qn = p * (q_[s] - tx - iy - iz);
				const Real q = fl ? qn : q_[s];
				tx = q * ai_[s] * p; oy = q * aj_[s] * p; oz = q * ak_[s] * p;
				q_[s] = q; oy_[s] = oy; oz_[s] = oz;
			} else {
				const Real qn = p * (q_[s] - tx * ai_[s] * p - iy * aj_[s] * p - iz * ak_[s] * p);
				const Real q = fl ? qn : q_[s];
				tx = q; oy = q; oz = q;
				q_[s] = q;
			}
		}
		#pragma unroll
		for (int s = 0; s < CH; s++) {
			stsReal(ad[s] + MW_ARR(4), q_[s]);                 // the finished chunk is written back from here
			if (!bwd) {                                        // the products of the edge rows are handed on from the dead coefficient cells
				if (wrY) stsReal(ad[s] + MW_ARR(2), oy_[s]);
				if (wrZ) stsReal(ad[s] + MW_ARR(3), oz_[s]);
			}
		}
		// ---- chunk A is finished: hand the edge rows to the next columns first (critical path), then write the result back
		if (inA) {
			const int cb = bwd ? nchm1 - cA : cA;
			if (wrY) { uint4 pc[NP]; packPieces(ldsVec(slA + MW_ARR(bwd ? 4 : 2)), ldsVec(slA + MW_ARR(bwd ? 4 : 2) + 16), tag, pc);
				#pragma unroll
				for (int qq = 0; qq < NP; qq++) stPiece(outY + 4 * cb + qq, pc[qq]); }
			if (wrZ) { uint4 pc[NP]; packPieces(ldsVec(slA + MW_ARR(bwd ? 4 : 3)), ldsVec(slA + MW_ARR(bwd ? 4 : 3) + 16), tag, pc);
				#pragma unroll
				for (int qq = 0; qq < NP; qq++) stPiece(outZ + 4 * cb + qq, pc[qq]); }
			if (fmA) storeChunk<Real, CH, VEC>(rDst, CH * cb, fmA, slA + MW_ARR(4));
		}
		oA = oB;
	}
}

template <typename Real, int MODE>
static int micWarpSweep(mp_context* ctx, const Dims& d, Real* dst, const Real* src, const Real* P, const Real* Ai, const Real* Aj, const Real* Ak, const int* doneFlag)
{
	constexpr int CH = ChunkOf<Real>::CH, S = ChunkOf<Real>::S;
	ColGeom g = { d.sx, d.sy, d.sz, d.Y, d.Z, (d.sx + CH - 1) / CH, (d.sy + 7) / 8, (d.sz + 3) / 4, 0 };
	g.nRounds = g.nch + (10 + CH - 1) / CH;          // the last lane trails the first by 10 cells
	const bool vec = (d.sx % CH) == 0;               // every chunk is a whole, 32-byte aligned sector
	uint4* mailY = (uint4*)ctx->micMail; uint4* mailZ = mailY + (size_t)g.nJ * g.sz * g.nch * 4;
	MP_TRY(micStallFlag(ctx)); int* stall = ctx->micStall;
	if (++ctx->micTag == 0) ctx->micTag = 1;         // mailboxes are zero-initialised: 0 is never a valid tag
	// shared memory doubles as the occupancy knob: the sweep is a chain of dependent rounds, and a warp that shares its scheduler
	// with others runs its rounds slower -- MP_MIC_WARPS_PER_SM caps the resident warps per SM by padding the allocation
	static const int capWarps = getenv("MP_MIC_WARPS_PER_SM") ? atoi(getenv("MP_MIC_WARPS_PER_SM")) : 0;
	size_t smem = (size_t)S * MW_SLOT_BYTES;
	if (capWarps > 0) { const size_t want = (size_t)(227 * 1024) / capWarps - 1024; if (want > smem) smem = want; }
	const unsigned int grid = (unsigned)(g.nJ * g.nK);
	static bool attr = false;
	if (!attr) {
		MP_CUDA(cudaFuncSetAttribute(k_mic_warp<float, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		MP_CUDA(cudaFuncSetAttribute(k_mic_warp<float, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		MP_CUDA(cudaFuncSetAttribute(k_mic_warp<float, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		MP_CUDA(cudaFuncSetAttribute(k_mic_warp<float, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		MP_CUDA(cudaFuncSetAttribute(k_mic_warp<double, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		MP_CUDA(cudaFuncSetAttribute(k_mic_warp<double, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		MP_CUDA(cudaFuncSetAttribute(k_mic_warp<double, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		MP_CUDA(cudaFuncSetAttribute(k_mic_warp<double, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		attr = true;
	}
	if (vec) k_mic_warp<Real, MODE, true><<<grid, 32, smem, ctx->stream>>>(g, ctx->micMask, dst, src, P, Ai, Aj, Ak, mailY, mailZ, ctx->micTag, stall, doneFlag, ctx->micOrder);
	else     k_mic_warp<Real, MODE, false><<<grid, 32, smem, ctx->stream>>>(g, ctx->micMask, dst, src, P, Ai, Aj, Ak, mailY, mailZ, ctx->micTag, stall, doneFlag, ctx->micOrder);
	MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

// (re)build the fluid mask and size the progress counters for the warp sweeps
static int micWarpPrepare(mp_context* ctx, const Dims& d, const mp_grid* flags, int prec) {
	const int CH = 32 / prec;
	ColGeom g = { d.sx, d.sy, d.sz, d.Y, d.Z, (d.sx + CH - 1) / CH, (d.sy + 7) / 8, (d.sz + 3) / 4, 0 };
	const size_t needP = sizeof(int) * ((size_t)g.nJ * g.nK + 1), needM = (size_t)g.nch * d.sy * d.sz;
	if (ctx->micProgBytes < needP) {
		if (ctx->micProg) { MP_CUDA(cudaStreamSynchronize(ctx->stream)); MP_CUDA(cudaFree(ctx->micProg)); ctx->micProg = nullptr; }
		MP_CUDA(cudaMalloc((void**)&ctx->micProg, needP)); ctx->micProgBytes = needP;
	}
	MP_CUDA(cudaMemsetAsync(ctx->micProg, 0, needP, ctx->stream));
	const size_t needMail = 64 * ((size_t)g.nJ * d.sz + (size_t)g.nK * d.sy) * g.nch;
	if (ctx->micMailBytes < needMail) {
		if (ctx->micMail) { MP_CUDA(cudaStreamSynchronize(ctx->stream)); MP_CUDA(cudaFree(ctx->micMail)); ctx->micMail = nullptr; }
		MP_CUDA(cudaMalloc((void**)&ctx->micMail, needMail)); ctx->micMailBytes = needMail;
	}
	MP_CUDA(cudaMemsetAsync(ctx->micMail, 0, needMail, ctx->stream)); ctx->micTag = 0;     // a different geometry re-indexes the entries: start the tags over
	if (ctx->micMaskBytes < needM) {
		if (ctx->micMask) { MP_CUDA(cudaStreamSynchronize(ctx->stream)); MP_CUDA(cudaFree(ctx->micMask)); ctx->micMask = nullptr; }
		MP_CUDA(cudaMalloc((void**)&ctx->micMask, needM)); ctx->micMaskBytes = needM;
	}
	{	// dispatch order of the columns: by earliest start
		const int n = g.nJ * g.nK, lagJ = 9, lagK = 8;     // hand-off lags are about equal in both directions
		if (ctx->micOrderCount < n) {
			if (ctx->micOrder) { MP_CUDA(cudaStreamSynchronize(ctx->stream)); MP_CUDA(cudaFree(ctx->micOrder)); ctx->micOrder = nullptr; }
			MP_CUDA(cudaMalloc((void**)&ctx->micOrder, sizeof(int) * n)); ctx->micOrderCount = n;
		}
		std::vector<int> ord(n);
		for (int q = 0; q < n; q++) ord[q] = q;
		const int nJ = g.nJ;
		std::stable_sort(ord.begin(), ord.end(), [nJ, lagJ, lagK](int a, int b) { return lagJ * (a % nJ) + lagK * (a / nJ) < lagJ * (b % nJ) + lagK * (b / nJ); });
		MP_CUDA(cudaMemcpyAsync(ctx->micOrder, ord.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
		MP_CUDA(cudaStreamSynchronize(ctx->stream));      // ord is a local
	}
	const unsigned int blocks = gridFor((IndexInt)needM, 256) < (unsigned)ctx->smCount * 16 ? gridFor((IndexInt)needM, 256) : (unsigned)ctx->smCount * 16;
	if (prec == 4) k_mic_fluid_mask<8><<<blocks, 256, 0, ctx->stream>>>(g, (const int*)flags->d, ctx->micMask);
	else           k_mic_fluid_mask<4><<<blocks, 256, 0, ctx->stream>>>(g, (const int*)flags->d, ctx->micMask);
	MP_CHECK_LAUNCH(ctx);
	ctx->micMaskFor = flags; ctx->micMaskPrec = prec;
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

// MP_MIC = 1 | 2 | 3 | 4 forces a schedule; default: the warp columns (v4) for large grids, the tile columns (v3) below ~100 MB per
// Real grid, where both are bound by the dependency chain and the tiles' shorter chain wins (profiles/r1_mic_variants.txt)
static inline int micVariantFor(const mp_grid* g) {
	const char* e = getenv("MP_MIC");            // read per call: the parity tests switch schedules inside one process
	const int forced = e ? atoi(e) : 0;
	if (forced) return forced;
	return (double)g->n * g->prec >= 1.0e8 ? 4 : 3;
}

int mp_mic_init_launch(mp_context* ctx, const mp_grid* flags, mp_grid* P, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak)
{
	const Dims d = dimsOf(flags);
	if (!d.is3D) MP_FAIL(MP_ERR_INVALID, "mICP only supports 3D grids so far");
	{	// mp_set_mic_ordering(ctx, 1, ...): MIC(0) of the block red-black ordering (mp_micrb.cu) when the matrix allows it
		bool rb = false;
		MP_TRY(mp_micrb_prepare(ctx, flags, P, Ai, Aj, Ak, &rb));
		if (rb) return mp_micrb_init_launch(ctx, P, A0);
	}
	MP_CUDA(cudaMemsetAsync(P->d, 0, P->bytes, ctx->stream));          // Aprecond.clear() :71
	if (d.sx < 3 || d.sy < 3 || d.sz < 3) return MP_OK;
	const int variant = micVariantFor(P);
	if (variant == 4) MP_TRY(micWarpPrepare(ctx, d, flags, P->prec));       // fluid mask of the warp sweeps; the factor itself uses the tile columns
	if (variant >= 3) {
		if (P->prec == 4) return micColsSweep<float, 0>(ctx, d, flags, nullptr, nullptr, (float*)P->d, (const float*)A0->d, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, nullptr);
		return micColsSweep<double, 0>(ctx, d, flags, nullptr, nullptr, (double*)P->d, (const double*)A0->d, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, nullptr);
	}
	if (variant == 2) {
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
	if (mp_micrb_active(ctx, flags, P)) return mp_micrb_apply_launch(ctx, dst, var1, P, doneFlag);     // the factor in P is the block red-black one
	const int variant = micVariantFor(dst);
	if (variant == 4) {
		if (ctx->micMaskFor != flags || ctx->micMaskPrec != dst->prec) MP_TRY(micWarpPrepare(ctx, d, flags, dst->prec));
		if (dst->prec == 4) {
			MP_TRY((micWarpSweep<float, 1>(ctx, d, (float*)dst->d, (const float*)var1->d, (const float*)P->d, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, doneFlag)));
			return micWarpSweep<float, 2>(ctx, d, (float*)dst->d, nullptr, (const float*)P->d, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, doneFlag);
		}
		MP_TRY((micWarpSweep<double, 1>(ctx, d, (double*)dst->d, (const double*)var1->d, (const double*)P->d, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, doneFlag)));
		return micWarpSweep<double, 2>(ctx, d, (double*)dst->d, nullptr, (const double*)P->d, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, doneFlag);
	}
	if (variant == 3) {
		if (dst->prec == 4) {
			MP_TRY((micColsSweep<float, 1>(ctx, d, flags, (float*)dst->d, (const float*)var1->d, (float*)P->d, nullptr, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, doneFlag)));
			return micColsSweep<float, 2>(ctx, d, flags, (float*)dst->d, nullptr, (float*)P->d, nullptr, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, doneFlag);
		}
		MP_TRY((micColsSweep<double, 1>(ctx, d, flags, (double*)dst->d, (const double*)var1->d, (double*)P->d, nullptr, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, doneFlag)));
		return micColsSweep<double, 2>(ctx, d, flags, (double*)dst->d, nullptr, (double*)P->d, nullptr, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, doneFlag);
	}
	if (variant == 2) {
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
	ctx->micMaskFor = nullptr;                       // standalone call: the flags may have changed since the mask was built
	MP_TRY(mp_mic_apply_launch(ctx, dst, var1, flags, Aprecond, Ai, Aj, Ak, nullptr));
	return mp_mic_check_stall(ctx);
}

}

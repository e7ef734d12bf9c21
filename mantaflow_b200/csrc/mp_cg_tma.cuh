// The fused PcNone matvec of GridCg as a TMA-staged, warp-specialised, persistent kernel (sm_100a).
//
// Same arithmetic as k_matvec_fused (mp_cg.cu): s = r + beta s_old (UpdateSearchVec conjugategrad.cpp:193-196), x += alpha_prev s_old
// (gridScaledAdd :254), t = A s (ApplyMatrix conjugategrad.h:118-133, left-to-right), dp = t.s (GridDotProduct :175-178) -- value for value.
// What changes is how the bytes travel:
//   * one CTA (or two) per SM, resident for the whole launch; CTA b takes the work items b, b + G, b + 2G, ... where an item is an x-y tile
//     of 32 16-byte vectors x 8 rows marched through a chunk of z-planes.  The number of chunks is chosen so the items divide evenly
//     over the G resident CTAs (no partial last wave), and consecutive items are x-y neighbours, so tiles that share a halo row are
//     in flight together and the halo is an L2 hit.
//   * a producer warp streams every plane of the tile through a ring of NSTAGE shared-memory stages with cp.async.bulk.tensor (TMA,
//     3-D tensor maps over the grids, out-of-range halo cells zero-filled by the hardware) signalling an mbarrier per stage:
//       r box and s_old box (tile + 1 row of halo in y, + 16 bytes in x), x tile, coupling-mask tile.
//     The 8 consumer warps never issue a global load: the search vector of the +-y neighbours is formed from the staged r / s_old rows,
//     the +-x neighbours come from the neighbouring lane (shuffle) or the staged halo column, +-z is carried in registers along the march.
//     Loads are therefore NSTAGE - 2 planes ahead of the arithmetic instead of one dependent round trip per plane.
//   * the matrix is two bytes per cell: bit 0 fluid row, bits 1..6 coupling to -x,+x,-y,+y,-z,+z present (every off-diagonal of the
//     pressure matrix is 0 or -1 without face fractions, MakeLaplaceMatrix conjugategrad.h:169-171), bits 7..9 the diagonal when it is a
//     small integer 0..6 (it counts the non-obstacle neighbours), 7 = read A0 (ghost-fluid diagonals, the pinned cell) -- verified
//     against A0/Ai/Aj/Ak by k_build_cmask on every solve.
// DRAM traffic per cell: R r, s_old, x (3w) + mask 2, W t, s, x (3w) = 2 + 6w -> 26 B float / 50 B double.
#pragma once
#include <cuda.h>
#include "mp_common.cuh"
#include "mp_cg.cuh"

template <typename Real, int TY_> struct FusedTmaGeom {
	static constexpr int V = 16 / (int)sizeof(Real);       // cells per 16-byte vector
	static constexpr int TX = 32 * V, TY = TY_;            // tile: one warp row of vectors x TY rows (one consumer warp per row)
	static constexpr int HX = V;                           // x halo in cells (TMA boxes are 16-byte granular)
	static constexpr int BX = TX + 2 * HX, BY = TY + 2;    // staged box of r / s_old
	static constexpr int boxBytes = BX * BY * (int)sizeof(Real);
	static constexpr int boxPad = (boxBytes + 127) / 128 * 128;      // TMA destinations are 128-byte aligned
	static constexpr int xBytes = TX * TY * (int)sizeof(Real);
	static constexpr int mBytes = TX * TY * 2;
	static constexpr int offS = boxPad, offX = 2 * boxPad, offM = 2 * boxPad + xBytes;
	static constexpr int stageBytes = 2 * boxPad + xBytes + (mBytes + 127) / 128 * 128;
	static constexpr int txEdge = 2 * boxBytes, txInterior = 2 * boxBytes + xBytes + mBytes;
	static constexpr int consumers = 32 * TY, threads = consumers + 32;      // TY consumer warps + 1 producer warp
	static constexpr int ctasPerSm = TY <= 8 ? 2 : 1;
};

// ---------------------------------------------------------------- PTX wrappers (mbarrier, TMA)
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbarExpectTx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbarArrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity) {
	uint32_t ok;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	} while (!ok);
}
__device__ __forceinline__ void tmaLoad3D(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
		:: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

template <typename Real, int V> struct alignas(sizeof(Real) * V) TVec { Real v[V]; };
template <int V> struct alignas(2 * V) MVec { unsigned short v[V]; };

// ---------------------------------------------------------------- the kernel
template <typename Real, int TY, int NSTAGE>
__global__ void __launch_bounds__((FusedTmaGeom<Real, TY>::threads), (FusedTmaGeom<Real, TY>::ctasPerSm)) k_matvec_fused_tma(
	const __grid_constant__ CUtensorMap mapR, const __grid_constant__ CUtensorMap mapS, const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapM,
	Dims d, int tilesX, int tiles, int chunk, int nitems,
	Real* __restrict__ dst, Real* __restrict__ sNew, Real* __restrict__ x, const Real* __restrict__ A0,
	CgScal<Real>* sc, double* partials, unsigned int* ticket, double* distLocal)
{
	typedef FusedTmaGeom<Real, TY> G;
	constexpr int V = G::V;
	typedef TVec<Real, V> Vec;
	if (sc->done) return;
	extern __shared__ __align__(128) unsigned char smem[];      // TMA destinations are 128-byte aligned
	uint64_t* const bars = (uint64_t*)(smem + (size_t)NSTAGE * G::stageBytes);      // full[NSTAGE], empty[NSTAGE]
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	if (tid == 0) {
		for (int q = 0; q < NSTAGE; q++) { mbarInit(smemAddr(bars + q), 1); mbarInit(smemAddr(bars + NSTAGE + q), G::consumers / 32); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	const Real beta = sc->beta, alphaP = sc->xPending ? sc->alpha : (Real)0;
	const IndexInt Y = d.Y, Z = d.Z;
	double acc = 0.0;

	if (warp == G::consumers / 32) {
		// ------------------------------------------------ producer: one lane walks the same item / plane sequence as the consumers
		if (lane == 0) {
			int slot = 0; uint32_t phase = 0;
			for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
				const int zc = item / tiles, tile = item - zc * tiles;
				const int x0 = (tile % tilesX) * G::TX, y0 = (tile / tilesX) * G::TY;
				const int k0 = d.kb + zc * chunk, k1 = min(d.ke, k0 + chunk);
				for (int p = k0 - 1; p <= k1; p++) {
					if (p < 0 || p >= d.sz) continue;                    // outside the grid: the consumers use zeros
					const bool interior = p >= k0 && p < k1;
					const uint32_t full = smemAddr(bars + slot), stage = smemAddr(smem + (size_t)slot * G::stageBytes);
					mbarWait(smemAddr(bars + NSTAGE + slot), phase ^ 1);  // the consumers have left this stage
					mbarExpectTx(full, interior ? G::txInterior : G::txEdge);
					tmaLoad3D(stage, &mapR, x0 - G::HX, y0 - 1, p, full);
					tmaLoad3D(stage + G::offS, &mapS, x0 - G::HX, y0 - 1, p, full);
					if (interior) {
						tmaLoad3D(stage + G::offX, &mapX, x0, y0, p, full);
						tmaLoad3D(stage + G::offM, &mapM, x0, y0, p, full);
					}
					if (++slot == NSTAGE) { slot = 0; phase ^= 1; }
				}
			}
		}
	} else {
		// ------------------------------------------------ consumers: 32 x 8 threads, one 16-byte vector of cells each
		const int tx = lane, ty = warp;
		const int cOff = (ty + 1) * G::BX + G::HX + tx * V;             // this thread's vector inside a staged box
		int slot = 0; uint32_t phase = 0;
		#define STAGE_PTR(s_) (smem + (size_t)(s_) * G::stageBytes)
		#define ADVANCE() do { if (++slot == NSTAGE) { slot = 0; phase ^= 1; } } while (0)
		#define RELEASE(s_) do { __syncwarp(); if (lane == 0) mbarArrive(smemAddr(bars + NSTAGE + (s_))); } while (0)
		for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
			const int zc = item / tiles, tile = item - zc * tiles;
			const int x0 = (tile % tilesX) * G::TX, y0 = (tile / tilesX) * G::TY;
			const int k0 = d.kb + zc * chunk, k1 = min(d.ke, k0 + chunk);
			const int gx = x0 + tx * V, gy = y0 + ty;
			const bool inb = gx < d.sx && gy < d.sy;                      // sx % V == 0: a vector is inside or outside as a whole
			IndexInt idx = (IndexInt)gx + Y * gy + Z * k0;
			Vec sm, s0, sp, so0, sop;                                     // new search vector at k-1, k, k+1; old one at k, k+1
			#pragma unroll
			for (int q = 0; q < V; q++) { sm.v[q] = (Real)0; sp.v[q] = (Real)0; sop.v[q] = (Real)0; }
			auto centre = [&](int s_, Vec& snew, Vec& sold) {
				const Real* bR = (const Real*)STAGE_PTR(s_); const Real* bS = (const Real*)(STAGE_PTR(s_) + G::offS);
				const Vec rv = *(const Vec*)(bR + cOff); sold = *(const Vec*)(bS + cOff);
				#pragma unroll
				for (int q = 0; q < V; q++) snew.v[q] = rv.v[q] + beta * sold.v[q];
			};
			if (k0 - 1 >= 0) {                                           // plane k0-1: only its centre values are needed
				mbarWait(smemAddr(bars + slot), phase);
				Vec dummy; centre(slot, sm, dummy);
				if (d.world > 1 && k0 == d.kb && inb) *(Vec*)(sNew + idx - Z) = sm;      // slab mode: the lower ghost plane of the new vector
				RELEASE(slot); ADVANCE();
			}
			mbarWait(smemAddr(bars + slot), phase);
			int cur = slot; ADVANCE();
			centre(cur, s0, so0);
			for (int k = k0; k < k1; k++, idx += Z) {
				int nxt = cur;
				if (k + 1 < d.sz) {
					mbarWait(smemAddr(bars + slot), phase);
					nxt = slot; ADVANCE();
					centre(nxt, sp, sop);
				}
				const unsigned char* st = STAGE_PTR(cur);
				const Real* bR = (const Real*)st; const Real* bS = (const Real*)(st + G::offS);
				const MVec<V> f = *(const MVec<V>*)(st + G::offM + (size_t)(ty * G::TX + tx * V) * 2);
				Vec xv = *(const Vec*)(st + G::offX + (size_t)(ty * G::TX + tx * V) * sizeof(Real));
				#pragma unroll
				for (int q = 0; q < V; q++) xv.v[q] += alphaP * so0.v[q];
				// +-y neighbours of the new search vector from the staged rows; +-x from the neighbouring lanes, the halo column at the tile edge
				Vec sym, syp;
				{
					const Vec r0 = *(const Vec*)(bR + cOff - G::BX), o0 = *(const Vec*)(bS + cOff - G::BX);
					const Vec r1 = *(const Vec*)(bR + cOff + G::BX), o1 = *(const Vec*)(bS + cOff + G::BX);
					#pragma unroll
					for (int q = 0; q < V; q++) { sym.v[q] = r0.v[q] + beta * o0.v[q]; syp.v[q] = r1.v[q] + beta * o1.v[q]; }
				}
				Real sxm0 = __shfl_up_sync(0xffffffffu, s0.v[V - 1], 1), sxpL = __shfl_down_sync(0xffffffffu, s0.v[0], 1);
				if (lane == 0) sxm0 = bR[cOff - 1] + beta * bS[cOff - 1];
				if (lane == 31) sxpL = bR[cOff + V] + beta * bS[cOff + V];
				int any = 0;
				#pragma unroll
				for (int q = 0; q < V; q++) any |= f.v[q];
				Vec out = s0;
				if (any & 1) {
					#pragma unroll
					for (int q = 0; q < V; q++) {
						const int m = f.v[q];
						if (m & 1) {
							const int code = (m >> 7) & 7;
							const Real a0 = code < 7 ? (Real)code : A0[idx + q];
							const Real xm = (q == 0) ? sxm0 : s0.v[q - 1 < 0 ? 0 : q - 1];
							const Real xp = (q == V - 1) ? sxpL : s0.v[q + 1 > V - 1 ? V - 1 : q + 1];
							// same left-to-right order as ApplyMatrix; a present coupling contributes s_nb * (-1) = -s_nb
							Real t = s0.v[q] * a0;
							t = t + ((m & 2) ? -xm : (Real)0);
							t = t + ((m & 4) ? -xp : (Real)0);
							t = t + ((m & 8) ? -sym.v[q] : (Real)0);
							t = t + ((m & 16) ? -syp.v[q] : (Real)0);
							t = t + ((m & 32) ? -sm.v[q] : (Real)0);
							t = t + ((m & 64) ? -sp.v[q] : (Real)0);
							out.v[q] = t;
						}
					}
				}
				RELEASE(cur);                                              // everything of this stage is in registers now
				if (inb) {
					*(Vec*)(x + idx) = xv;
					*(Vec*)(dst + idx) = out;
					*(Vec*)(sNew + idx) = s0;
				}
				#pragma unroll
				for (int q = 0; q < V; q++) acc += (double)(out.v[q] * s0.v[q]);
				cur = nxt; sm = s0; s0 = sp; so0 = sop;
			}
			if (k1 < d.sz) {                                               // the stage of plane k1 is still held
				if (d.world > 1 && k1 == d.ke && inb) *(Vec*)(sNew + idx) = s0;            // slab mode: the upper ghost plane (idx is at plane k1 now)
				RELEASE(cur);
			}
		}
		#undef STAGE_PTR
		#undef ADVANCE
		#undef RELEASE
	}
	double v[1] = { acc }; const bool isMax[1] = { false }; double fin[1];
	if (blockReduceFinalL<1>(v, isMax, partials, ticket, fin, (unsigned)tid, G::threads, blockIdx.x, gridDim.x) && tid == 0) {
		if (distLocal) distLocal[0] = fin[0]; else cgFinA<Real>(sc, fin[0]);
	}
}

// ---------------------------------------------------------------- host side: tensor maps and the work decomposition
struct FusedTma {
	bool on = false;
	CUtensorMap mapR, mapS[2], mapX, mapM;      // mapS[0]: the caller's search grid, mapS[1]: search2
	mp_grid* mask16 = nullptr;                  // 2 bytes per cell, row pitch rounded up to 8 cells (backed by a pooled 4-byte grid)
	int pitch = 0, ty = 16, tilesX = 0, tiles = 0, chunk = 0, nitems = 0, ctas = 0, smemBytes = 0;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encodeTiledFn() {
	static EncodeTiledFn fn = nullptr; static bool tried = false;
	if (!tried) {
		tried = true;
		void* p = nullptr; cudaDriverEntryPointQueryResult qr;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
		else cudaGetLastError();
	}
	return fn;
}
// 3-D map over an x-fastest array of `es`-byte elements: dims (nx, ny, nz), row pitch `pitch` elements, box (bx, by, 1)
static int encodeMap3D(CUtensorMap* m, void* base, int es, CUtensorMapDataType dt, int nx, int pitch, int ny, int nz, int bx, int by) {
	EncodeTiledFn fn = encodeTiledFn();
	if (!fn) MP_FAIL(MP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
	const cuuint64_t dims[3] = { (cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz };
	const cuuint64_t strides[2] = { (cuuint64_t)pitch * es, (cuuint64_t)pitch * es * ny };
	const cuuint32_t box[3] = { (cuuint32_t)bx, (cuuint32_t)by, 1 }, estr[3] = { 1, 1, 1 };
	const CUresult r = fn(m, dt, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) MP_FAIL(MP_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (dims %d x %d x %d, pitch %d, box %d x %d, %d-byte elements)", (int)r, nx, ny, nz, pitch, bx, by, es);
	return MP_OK;
}
// How many z-chunks: CTA b takes items b, b+G, ...; the launch lasts ceil(items / G) rounds of (chunk + 2) staged planes each.
static void fusedTmaDecompose(int tiles, int planes, int G, int* chunkOut, int* nitemsOut) {
	long long best = -1; int bestChunk = planes;
	for (int nchunk = 1; nchunk <= planes; nchunk++) {
		const int chunk = (planes + nchunk - 1) / nchunk;
		if (chunk < 24 && nchunk > 1) break;
		const int nc = (planes + chunk - 1) / chunk;
		const long long rounds = ((long long)tiles * nc + G - 1) / G;
		const long long cost = rounds * (chunk + 2);
		if (best < 0 || cost < best) { best = cost; bestChunk = chunk; }
	}
	*chunkOut = bestChunk; *nitemsOut = tiles * ((planes + bestChunk - 1) / bestChunk);
}

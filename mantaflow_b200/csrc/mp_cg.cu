// GridCg on the device: conjugategrad.h:65-151, conjugategrad.cpp:170-339.
//
// One PCG iteration of the reference is 9 full-grid passes.  Here it is
//   k_matvec_dot    t = A s  fused with  dp = t.s           -> alpha          (R flags,s,A0,Ai,Aj,Ak  W t : 4+6w B/cell)
//   k_axpy2_norm    x += alpha s ; r -= alpha t  fused with |r|_inf (or sum r^2) and, for PcNone, r.r -> beta
//                                                                             (R x,s,r,t W x,r : 6w B/cell)
//   [preconditioner z = M^-1 r : MIC sweeps (mp_mic.cu) or a GridMg V-cycle (mp_mg.cu), then k_dot_zr -> beta]
//   k_update_search s = z + beta s                                            (R z,s W s : 3w B/cell)
// PcNone on a matrix whose off-diagonals are all 0 / -1 (every case without face fractions) runs two kernels instead:
//   k_matvec_fused  s = r + beta s_old, x += alpha_prev s_old, t = A s, dp = t.s  ->  k_axpy1_norm  r -= alpha t, norms   (44 B/cell float)
// All scalars (sigma, alpha, beta, dp, resNorm are `Real`, accumulators double, conjugategrad.h:108-113,
// conjugategrad.cpp:250-252,279-280) live on the device; the host only polls a pinned copy every few
// iterations, so the loop of solvePressureSystem (pressure.cpp:436-439) never stalls the stream.
// Kernels early-out once `done` is set, which makes the overshoot of the lagged poll free of side effects.
#include "mp_common.cuh"
#include "mp_cg.cuh"
#include <cstdlib>

// ---------------------------------------------------------------- vector access helpers
template <typename T, int V> struct alignas(sizeof(T) * V) VecT { T v[V]; };
template <typename T, int V> __device__ __forceinline__ VecT<T, V> ldv(const T* p) { return *reinterpret_cast<const VecT<T, V>*>(p); }
template <typename T, int V> __device__ __forceinline__ void stv(T* p, const VecT<T, V>& x) { *reinterpret_cast<VecT<T, V>*>(p) = x; }

// ---------------------------------------------------------------- scalar updates of one iteration (conjugategrad.cpp:241-295)
// Executed by the last block of the reducing kernel on a single GPU, or by k_cg_combine after the all-gather of the
// ranks' partial sums in slab mode (same code, so both paths produce the same scalars).
template <typename Real> __device__ __forceinline__ void cgFinA(CgScal<Real>* sc, double dpSum) {
	const Real dp = (Real)dpSum;                       // iterate(): mIterations++ ; dp ; alpha (:241,:250-252)
	Real alpha = (Real)0.;
	if (fabs((double)dp) > 0.) alpha = sc->sigma / dp;
	sc->alpha = alpha; sc->dp = dp;
	sc->iterations += 1;
	sc->xPending = 1;
}
template <typename Real> __device__ __forceinline__ void cgFinB(CgScal<Real>* sc, double nrm, double rr, int mode) {
	const Real resNorm = (Real)nrm;
	sc->resNorm = resNorm;
	if (resNorm < sc->accuracy) { sc->sigma = resNorm; sc->done = 1; }              // :274-277
	else {
		if (mode == 0) {
			const Real sigmaNew = (Real)rr;
			sc->beta = sigmaNew / sc->sigma;                                          // :279-280
			sc->sigma = sigmaNew;                                                     // :286
		}
		if (!(resNorm < (Real)1e35)) { sc->diverged = 1; sc->done = 1; }              // :288-295
	}
}
template <typename Real> __device__ __forceinline__ void cgFinZR(CgScal<Real>* sc, double zr, int isInit) {
	const Real sigmaNew = (Real)zr;
	if (isInit) sc->sigma = sigmaNew;                                                 // doInit :234
	else { sc->beta = sigmaNew / sc->sigma; sc->sigma = sigmaNew; }                   // :279-286
}
// slab mode: combine the all-gathered partials (gathered[r*8 + q]) in rank order and apply the same update
template <typename Real>
__global__ void k_cg_combine(const double* __restrict__ gathered, int world, CgScal<Real>* sc, int stage, int mode) {
	if (sc->done && stage != 3) return;
	double a = 0.0, b = 0.0;
	const bool useMax = (stage == 1) && !sc->useL2;
	if (useMax) a = -1.0;
	for (int r = 0; r < world; r++) {
		const double g0 = gathered[8 * r], g1 = gathered[8 * r + 1];
		a = useMax ? fmax(a, g0) : a + g0;
		b += g1;
	}
	if (stage == 0) cgFinA<Real>(sc, a);
	else if (stage == 1) cgFinB<Real>(sc, a, b, mode);
	else cgFinZR<Real>(sc, a, stage == 3);
}

#include "mp_cg_tma.cuh"

// ---------------------------------------------------------------- k_matvec_dot
// ApplyMatrix / ApplyMatrix2D (conjugategrad.h:118-151) over ALL cells: identity rows on non-fluid cells,
// same left-to-right summation order as the reference (bit-identical t with -fmad=false), fused with
// GridDotProduct(t, s) (product in Real, double accumulation, conjugategrad.cpp:175-178).
// Each thread owns V consecutive x-cells (V=4 float / 2 double when sx % V == 0 -> 16-byte loads along x);
// +-Y / +-Z neighbours are re-read through L1/L2 (a z-plane of the whole chip's working window stays in the
// 126 MB L2, so DRAM traffic is the compulsory 4+6w B/cell).
template <typename Real, int V, bool IS3D>
__global__ void __launch_bounds__(256) k_matvec_dot(Dims d, const int* __restrict__ flags, Real* __restrict__ dst, const Real* __restrict__ src,
	const Real* __restrict__ A0, const Real* __restrict__ Ai, const Real* __restrict__ Aj, const Real* __restrict__ Ak,
	CgScal<Real>* sc, double* partials, unsigned int* ticket, int finalize, double* distLocal)
{
	if (sc && sc->done) return;
	const IndexInt Y = d.Y, Z = d.Z;
	const IndexInt nv = d.i1 / V;                // owned range [i0,i1); both are multiples of V (launcher)
	double acc = 0.0;
	for (IndexInt vi = d.i0 / V + (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; vi < nv; vi += (IndexInt)gridDim.x * blockDim.x) {
		const IndexInt idx = vi * V;
		const VecT<int, V> f = ldv<int, V>(flags + idx);
		const VecT<Real, V> s = ldv<Real, V>(src + idx);
		bool any = false;
		#pragma unroll
		for (int q = 0; q < V; q++) any |= (f.v[q] & TypeFluid) != 0;
		VecT<Real, V> out = s;
		if (any) {
			// a vector holding a fluid cell lies in an interior row/plane, so the +-Y/+-Z vectors are in range
			const VecT<Real, V> a0 = ldv<Real, V>(A0 + idx), ai = ldv<Real, V>(Ai + idx), aj = ldv<Real, V>(Aj + idx);
			const VecT<Real, V> ajm = ldv<Real, V>(Aj + idx - Y), sym = ldv<Real, V>(src + idx - Y), syp = ldv<Real, V>(src + idx + Y);
			VecT<Real, V> ak, akm, szm, szp;
			if (IS3D) { ak = ldv<Real, V>(Ak + idx); akm = ldv<Real, V>(Ak + idx - Z); szm = ldv<Real, V>(src + idx - Z); szp = ldv<Real, V>(src + idx + Z); }
			// x neighbours: inside the vector, plus one scalar on each side (only read when that lane is fluid)
			Real sxm0 = 0, aim0 = 0, sxpL = 0;
			if (f.v[0] & TypeFluid) { sxm0 = src[idx - 1]; aim0 = Ai[idx - 1]; }
			if (f.v[V - 1] & TypeFluid) sxpL = src[idx + V];
			#pragma unroll
			for (int q = 0; q < V; q++) {
				if (f.v[q] & TypeFluid) {
					const Real sm = (q == 0) ? sxm0 : s.v[q - 1 < 0 ? 0 : q - 1];
					const Real am = (q == 0) ? aim0 : ai.v[q - 1 < 0 ? 0 : q - 1];
					const Real sp = (q == V - 1) ? sxpL : s.v[q + 1 > V - 1 ? V - 1 : q + 1];
					Real t = s.v[q] * a0.v[q] + sm * am + sp * ai.v[q] + sym.v[q] * ajm.v[q] + syp.v[q] * aj.v[q];
					if (IS3D) t = t + szm.v[q] * akm.v[q] + szp.v[q] * ak.v[q];
					out.v[q] = t;
				}
			}
		}
		stv<Real, V>(dst + idx, out);
		#pragma unroll
		for (int q = 0; q < V; q++) acc += (double)(out.v[q] * s.v[q]);
	}
	if (!finalize) return;
	double v[1] = { acc }; const bool isMax[1] = { false }; double fin[1];
	if (blockReduceFinal<1>(v, isMax, partials, ticket, fin) && threadIdx.x == 0) {
		if (distLocal) distLocal[0] = fin[0]; else cgFinA<Real>(sc, fin[0]);
	}
}

// ---------------------------------------------------------------- k_matvec_zmarch (3-D, vectorised sizes)
// Same arithmetic as k_matvec_dot, restructured so that NOTHING relies on the L2 for reuse along z: a CTA of 32 x 8
// threads owns an x-y tile (32 vectors x 8 rows) and marches along z through a chunk of planes, carrying s(k-1), s(k),
// s(k+1) and Ak(k-1) in registers; every plane of s, flags, A0, Ai, Aj, Ak is therefore requested once per tile.  The
// +-Y rows were loaded by the neighbouring warps of the same CTA one step earlier (L1 hits); only tile edges go to L2.
template <typename Real, int V>
__global__ void __launch_bounds__(256) k_matvec_zmarch(Dims d, int nvx, int chunk, const int* __restrict__ flags, Real* __restrict__ dst, const Real* __restrict__ src,
	const Real* __restrict__ A0, const Real* __restrict__ Ai, const Real* __restrict__ Aj, const Real* __restrict__ Ak,
	CgScal<Real>* sc, double* partials, unsigned int* ticket, int finalize, double* distLocal)
{
	if (sc && sc->done) return;
	const IndexInt Y = d.Y, Z = d.Z;
	const int vx = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
	const int k0 = d.kb + blockIdx.z * chunk, k1 = min(d.ke, k0 + chunk);
	double acc = 0.0;
	if (vx < nvx && j < d.sy && k0 < k1) {
		IndexInt idx = (IndexInt)vx * V + Y * j + Z * k0;
		VecT<Real, V> sm, s0, sp, akm;
		#pragma unroll
		for (int q = 0; q < V; q++) { sm.v[q] = (Real)0; akm.v[q] = (Real)0; }
		if (k0 > 0) { sm = ldv<Real, V>(src + idx - Z); akm = ldv<Real, V>(Ak + idx - Z); }
		s0 = ldv<Real, V>(src + idx);
		for (int k = k0; k < k1; k++, idx += Z) {
			const VecT<int, V> f = ldv<int, V>(flags + idx);
			const VecT<Real, V> ak = ldv<Real, V>(Ak + idx);
			if (k + 1 < d.sz) sp = ldv<Real, V>(src + idx + Z);
			bool any = false;
			#pragma unroll
			for (int q = 0; q < V; q++) any |= (f.v[q] & TypeFluid) != 0;
			VecT<Real, V> out = s0;
			if (any) {
				const VecT<Real, V> a0 = ldv<Real, V>(A0 + idx), ai = ldv<Real, V>(Ai + idx), aj = ldv<Real, V>(Aj + idx);
				const VecT<Real, V> ajm = ldv<Real, V>(Aj + idx - Y), sym = ldv<Real, V>(src + idx - Y), syp = ldv<Real, V>(src + idx + Y);
				Real sxm0 = 0, aim0 = 0, sxpL = 0;
				if (f.v[0] & TypeFluid) { sxm0 = src[idx - 1]; aim0 = Ai[idx - 1]; }
				if (f.v[V - 1] & TypeFluid) sxpL = src[idx + V];
				#pragma unroll
				for (int q = 0; q < V; q++) {
					if (f.v[q] & TypeFluid) {
						const Real xm = (q == 0) ? sxm0 : s0.v[q - 1 < 0 ? 0 : q - 1];
						const Real am = (q == 0) ? aim0 : ai.v[q - 1 < 0 ? 0 : q - 1];
						const Real xp = (q == V - 1) ? sxpL : s0.v[q + 1 > V - 1 ? V - 1 : q + 1];
						Real t = s0.v[q] * a0.v[q] + xm * am + xp * ai.v[q] + sym.v[q] * ajm.v[q] + syp.v[q] * aj.v[q];
						t = t + sm.v[q] * akm.v[q] + sp.v[q] * ak.v[q];
						out.v[q] = t;
					}
				}
			}
			stv<Real, V>(dst + idx, out);
			#pragma unroll
			for (int q = 0; q < V; q++) acc += (double)(out.v[q] * s0.v[q]);
			sm = s0; s0 = sp; akm = ak;
		}
	}
	if (!finalize) return;
	double v[1] = { acc }; const bool isMax[1] = { false }; double fin[1];
	const unsigned int tid = threadIdx.y * 32 + threadIdx.x;
	const unsigned int blockLinear = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z), numBlocks = gridDim.x * gridDim.y * gridDim.z;
	if (blockReduceFinalL<1>(v, isMax, partials, ticket, fin, tid, 256, blockLinear, numBlocks) && tid == 0) {
		if (distLocal) distLocal[0] = fin[0]; else cgFinA<Real>(sc, fin[0]);
	}
}

// ---------------------------------------------------------------- coupling-mask fast path
// Without face fractions every off-diagonal of the pressure matrix is exactly 0 or -1 (MakeLaplaceMatrix
// conjugategrad.h:169-171; fixPressure only zeroes entries, pressure.cpp:241-244), so Ai/Aj/Ak carry one bit of
// information per face.  k_build_cmask verifies that on the actual arrays and packs, per cell, bit0 = fluid row and
// bits 1..6 = coupling to -x,+x,-y,+y,-z,+z present.  k_matvec_zmarch_masked then streams cmask, A0 and s only:
// 4+3w B/cell (16 / 28) instead of 4+6w (28 / 52).  A product s_nb * (-1) is exactly -s_nb and s_nb * 0 is a zero, so
// the result equals ApplyMatrix's value for value (signs of exact zeros aside).  Any other off-diagonal value (face
// fractions) makes the mask invalid and the general kernel is used.
template <typename Real>
__global__ void __launch_bounds__(256) k_build_cmask(Dims d, const int* __restrict__ flags, const Real* __restrict__ A0, const Real* __restrict__ Ai, const Real* __restrict__ Aj,
	const Real* __restrict__ Ak, int* __restrict__ cmask, unsigned short* __restrict__ mask16, int pitch, int* bad)
{
	const IndexInt idx = d.i0 + (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= d.i1) return;
	int m = 0;
	if (flags[idx] & TypeFluid) {
		m = 1;
		const Real c[6] = { Ai[idx - 1], Ai[idx], Aj[idx - d.Y], Aj[idx], d.is3D ? Ak[idx - d.Z] : (Real)0, d.is3D ? Ak[idx] : (Real)0 };
		#pragma unroll
		for (int q = 0; q < 6; q++) {
			if (c[q] == (Real)-1) m |= 2 << q;
			else if (c[q] != (Real)0) *bad = 1;
		}
	}
	cmask[idx] = m;
	if (mask16) {      // + the diagonal as a 3-bit code: 0..6 = that integer (the count of non-obstacle neighbours), 7 = read A0
		int code = 0;
		if (m) {
			const Real a0 = A0[idx];
			code = 7;
			#pragma unroll
			for (int q = 0; q < 7; q++) if (a0 == (Real)q) code = q;
		}
		int i, j, k; cellOf(d, idx, i, j, k);
		mask16[((size_t)k * d.sy + j) * pitch + i] = (unsigned short)(m | (code << 7));
	}
}

template <typename Real, int V>
__global__ void __launch_bounds__(256) k_matvec_zmarch_masked(Dims d, int nvx, int chunk, const int* __restrict__ cmask, Real* __restrict__ dst, const Real* __restrict__ src,
	const Real* __restrict__ A0, CgScal<Real>* sc, double* partials, unsigned int* ticket, int finalize, double* distLocal)
{
	if (sc && sc->done) return;
	const IndexInt Y = d.Y, Z = d.Z;
	const int vx = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
	const int k0 = d.kb + blockIdx.z * chunk, k1 = min(d.ke, k0 + chunk);
	double acc = 0.0;
	if (vx < nvx && j < d.sy && k0 < k1) {
		IndexInt idx = (IndexInt)vx * V + Y * j + Z * k0;
		VecT<Real, V> sm, s0, sp;
		#pragma unroll
		for (int q = 0; q < V; q++) sm.v[q] = (Real)0;
		if (k0 > 0) sm = ldv<Real, V>(src + idx - Z);
		s0 = ldv<Real, V>(src + idx);
		for (int k = k0; k < k1; k++, idx += Z) {
			const VecT<int, V> f = ldv<int, V>(cmask + idx);
			if (k + 1 < d.sz) sp = ldv<Real, V>(src + idx + Z);
			int any = 0;
			#pragma unroll
			for (int q = 0; q < V; q++) any |= f.v[q];
			VecT<Real, V> out = s0;
			if (any & 1) {
				const VecT<Real, V> a0 = ldv<Real, V>(A0 + idx), sym = ldv<Real, V>(src + idx - Y), syp = ldv<Real, V>(src + idx + Y);
				Real sxm0 = 0, sxpL = 0;
				if (f.v[0] & 1) sxm0 = src[idx - 1];
				if (f.v[V - 1] & 1) sxpL = src[idx + V];
				#pragma unroll
				for (int q = 0; q < V; q++) {
					const int m = f.v[q];
					if (m & 1) {
						const Real xm = (q == 0) ? sxm0 : s0.v[q - 1 < 0 ? 0 : q - 1];
						const Real xp = (q == V - 1) ? sxpL : s0.v[q + 1 > V - 1 ? V - 1 : q + 1];
						// same left-to-right order as ApplyMatrix; a present coupling contributes s_nb * (-1) = -s_nb
						Real t = s0.v[q] * a0.v[q];
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
			stv<Real, V>(dst + idx, out);
			#pragma unroll
			for (int q = 0; q < V; q++) acc += (double)(out.v[q] * s0.v[q]);
			sm = s0; s0 = sp;
		}
	}
	if (!finalize) return;
	double v[1] = { acc }; const bool isMax[1] = { false }; double fin[1];
	const unsigned int tid = threadIdx.y * 32 + threadIdx.x;
	const unsigned int blockLinear = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z), numBlocks = gridDim.x * gridDim.y * gridDim.z;
	if (blockReduceFinalL<1>(v, isMax, partials, ticket, fin, tid, 256, blockLinear, numBlocks) && tid == 0) {
		if (distLocal) distLocal[0] = fin[0]; else cgFinA<Real>(sc, fin[0]);
	}
}

// ---------------------------------------------------------------- k_axpy2_norm
// gridScaledAdd x2 (conjugategrad.cpp:254-255) + residual norm (:267-271) [+ for PcNone z==r: sigmaNew = r.r, :279]
// MODE 0: PcNone (z = r, finalises beta/sigma here), MODE 1: preconditioned (only the norm / stop test here)
// ================================================================ fused PcNone iteration (coupling-mask matrices, 3-D, vector rows)
// Two kernels per iteration instead of three, 44 instead of 52 B/cell (float): the search-vector update of iteration k-1 and its
// x-update ride on the matvec of iteration k, which reads r and the old search vector anyway:
//   k_matvec_fused   s = r + beta s_old  (UpdateSearchVec :193-196, recomputed for the stencil neighbours: same inputs, same bits),
//                    x += alpha_prev s_old (gridScaledAdd :257), t = A s, dp = t.s -> alpha      (R r,s_old,x,cmask,A0  W s,t,x : 4+7w)
//   k_axpy1_norm     r -= alpha t, |r| (or sum r^2), r.r -> beta                                  (R r,t  W r : 3w)
// s ping-pongs between the caller's search grid and one more grid (no thread reads what another writes in the same launch);
// iteration k (0-based) writes buffer (k even ? B : A).  After the loop k_flush_x applies the last pending x-update.
template <typename Real, int V>
__global__ void __launch_bounds__(256, 5) k_matvec_fused(Dims d, int nvx, int chunk, const int* __restrict__ cmask, Real* __restrict__ dst,
	const Real* __restrict__ sOld, Real* __restrict__ sNew, const Real* __restrict__ r, Real* __restrict__ x,
	const Real* __restrict__ A0, CgScal<Real>* sc, double* partials, unsigned int* ticket, double* distLocal)
{
	if (sc->done) return;
	// (a shared-memory tile of the new vector with one barrier per plane was measured slower than re-forming the +-y / +-x neighbours
	//  from r and s_old through L1: 0.98 vs 0.87 ms at 512^3)
	const Real beta = sc->beta, alphaP = sc->xPending ? sc->alpha : (Real)0;
	const IndexInt Y = d.Y, Z = d.Z;
	const int vx = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
	const int k0 = d.kb + blockIdx.z * chunk, k1 = min(d.ke, k0 + chunk);
	double acc = 0.0;
	#define SNEW(off) snewOf(ldv<Real, V>(r + (off)), ldv<Real, V>(sOld + (off)), beta)
	auto snewOf = [](const VecT<Real, V>& rv, const VecT<Real, V>& sv, Real b) { VecT<Real, V> o;
		#pragma unroll
		for (int q = 0; q < V; q++) o.v[q] = rv.v[q] + b * sv.v[q];
		return o; };
	if (vx < nvx && j < d.sy && k0 < k1) {
		IndexInt idx = (IndexInt)vx * V + Y * j + Z * k0;
		VecT<Real, V> sm, s0, sp, so0, sop;      // new search vector at k-1, k, k+1; old one at k, k+1 (for the x-update)
		#pragma unroll
		for (int q = 0; q < V; q++) { sm.v[q] = (Real)0; sp.v[q] = (Real)0; sop.v[q] = (Real)0; }
		if (k0 > 0) {
			sm = SNEW(idx - Z);
			if (d.world > 1 && k0 == d.kb) stv<Real, V>(sNew + idx - Z, sm);       // slab mode: the lower ghost plane of the new vector
		}
		so0 = ldv<Real, V>(sOld + idx);
		{ const VecT<Real, V> rv = ldv<Real, V>(r + idx);
			#pragma unroll
			for (int q = 0; q < V; q++) s0.v[q] = rv.v[q] + beta * so0.v[q]; }
		for (int k = k0; k < k1; k++, idx += Z) {
			const VecT<int, V> f = ldv<int, V>(cmask + idx);
			if (k + 1 < d.sz) {
				sop = ldv<Real, V>(sOld + idx + Z);
				const VecT<Real, V> rv = ldv<Real, V>(r + idx + Z);
				#pragma unroll
				for (int q = 0; q < V; q++) sp.v[q] = rv.v[q] + beta * sop.v[q];
			}
			VecT<Real, V> xv = ldv<Real, V>(x + idx);
			#pragma unroll
			for (int q = 0; q < V; q++) xv.v[q] += alphaP * so0.v[q];
			stv<Real, V>(x + idx, xv);
			int any = 0;
			#pragma unroll
			for (int q = 0; q < V; q++) any |= f.v[q];
			VecT<Real, V> out = s0;
			if (any & 1) {
				const VecT<Real, V> a0 = ldv<Real, V>(A0 + idx), sym = SNEW(idx - Y), syp = SNEW(idx + Y);
				Real sxm0 = 0, sxpL = 0;
				if (f.v[0] & 1) sxm0 = r[idx - 1] + beta * sOld[idx - 1];
				if (f.v[V - 1] & 1) sxpL = r[idx + V] + beta * sOld[idx + V];
				#pragma unroll
				for (int q = 0; q < V; q++) {
					const int m = f.v[q];
					if (m & 1) {
						const Real xm = (q == 0) ? sxm0 : s0.v[q - 1 < 0 ? 0 : q - 1];
						const Real xp = (q == V - 1) ? sxpL : s0.v[q + 1 > V - 1 ? V - 1 : q + 1];
						Real t = s0.v[q] * a0.v[q];
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
			stv<Real, V>(dst + idx, out);
			stv<Real, V>(sNew + idx, s0);
			#pragma unroll
			for (int q = 0; q < V; q++) acc += (double)(out.v[q] * s0.v[q]);
			sm = s0; s0 = sp; so0 = sop;
		}
		if (d.world > 1 && k1 == d.ke && k1 < d.sz) stv<Real, V>(sNew + idx, s0);      // slab mode: the upper ghost plane (idx is at plane k1 now)
	}
	#undef SNEW
	double v[1] = { acc }; const bool isMax[1] = { false }; double fin[1];
	const unsigned int tid = threadIdx.y * 32 + threadIdx.x;
	const unsigned int blockLinear = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z), numBlocks = gridDim.x * gridDim.y * gridDim.z;
	if (blockReduceFinalL<1>(v, isMax, partials, ticket, fin, tid, 256, blockLinear, numBlocks) && tid == 0) {
		if (distLocal) distLocal[0] = fin[0]; else cgFinA<Real>(sc, fin[0]);
	}
}

// r -= alpha t with the residual norms of the stop test and r.r (the next sigma); in slab mode it also hands the first / last owned plane
// of the new residual to the neighbours (see k_update_search)
template <typename Real, int V>
__global__ void __launch_bounds__(256) k_axpy1_norm(IndexInt n, Real* __restrict__ r, const Real* __restrict__ t,
	CgScal<Real>* sc, double* partials, unsigned int* ticket, double* distLocal, HaloOut ho, IndexInt plane)
{
	if (sc->done) return;
	const Real nalpha = -sc->alpha;
	const bool useL2 = sc->useL2 != 0;
	const IndexInt nv = n / V;
	Real* const outLo = (Real*)ho.lo; Real* const outHi = (Real*)ho.hi;
	double nrm = useL2 ? 0.0 : -1.0, rr = 0.0;
	for (IndexInt vi = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; vi < nv; vi += (IndexInt)gridDim.x * blockDim.x) {
		const IndexInt idx = vi * V;
		VecT<Real, V> rv = ldv<Real, V>(r + idx);
		const VecT<Real, V> tv = ldv<Real, V>(t + idx);
		#pragma unroll
		for (int q = 0; q < V; q++) {
			rv.v[q] += nalpha * tv.v[q];
			const double rd = (double)rv.v[q];
			if (useL2) nrm += rd * rd; else nrm = fmax(nrm, fabs(rd));
			rr += (double)(rv.v[q] * rv.v[q]);
		}
		stv<Real, V>(r + idx, rv);
		if (outLo && idx < plane) stv<Real, V>(outLo + idx, rv);
		if (outHi && idx >= n - plane) stv<Real, V>(outHi + (idx - (n - plane)), rv);
	}
	if (ho.ticket) __threadfence_system();
	double v[2] = { nrm, rr }; const bool isMax[2] = { !useL2, false }; double fin[2];
	if (blockReduceFinal<2>(v, isMax, partials, ticket, fin) && threadIdx.x == 0) {
		if (distLocal) { distLocal[0] = fin[0]; distLocal[1] = fin[1]; } else cgFinB<Real>(sc, fin[0], fin[1], 0);
		if (ho.ticket) {      // this thread runs after every block has finished (and fenced) its stores
			__threadfence_system();
			if (ho.flagLo) st_release_sys(ho.flagLo, ho.seq);
			if (ho.flagHi) st_release_sys(ho.flagHi, ho.seq);
		}
	}
}

// after the loop: the x-update of the last executed iteration.  That iteration (index iterations-1) wrote its search vector to
// buffer B when its index is even, to A when odd.
template <typename Real, int V>
__global__ void __launch_bounds__(256) k_flush_x(IndexInt n, Real* __restrict__ x, const Real* __restrict__ sA, const Real* __restrict__ sB, CgScal<Real>* sc, unsigned int* ticket)
{
	if (!sc->xPending) return;
	const Real alpha = sc->alpha;
	const Real* s = (sc->iterations & 1) ? sB : sA;
	const IndexInt nv = n / V;
	for (IndexInt vi = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; vi < nv; vi += (IndexInt)gridDim.x * blockDim.x) {
		VecT<Real, V> xv = ldv<Real, V>(x + vi * V); const VecT<Real, V> sv = ldv<Real, V>(s + vi * V);
		#pragma unroll
		for (int q = 0; q < V; q++) xv.v[q] += alpha * sv.v[q];
		stv<Real, V>(x + vi * V, xv);
	}
	__syncthreads();
	if (threadIdx.x == 0) { __threadfence(); if (atomicAdd(ticket, 1u) == gridDim.x - 1) { *ticket = 0; sc->xPending = 0; } }     // the last block clears the flag
}

template <typename Real, int V, int MODE>
__global__ void __launch_bounds__(256) k_axpy2_norm(IndexInt n, Real* __restrict__ x, const Real* __restrict__ s, Real* __restrict__ r, const Real* __restrict__ t,
	CgScal<Real>* sc, double* partials, unsigned int* ticket, double* distLocal)
{
	if (sc->done) return;
	const Real alpha = sc->alpha, nalpha = -alpha;
	const bool useL2 = sc->useL2 != 0;
	const IndexInt nv = n / V;
	double nrm = useL2 ? 0.0 : -1.0, rr = 0.0;
	for (IndexInt vi = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; vi < nv; vi += (IndexInt)gridDim.x * blockDim.x) {
		const IndexInt idx = vi * V;
		VecT<Real, V> xv = ldv<Real, V>(x + idx), rv = ldv<Real, V>(r + idx);
		const VecT<Real, V> sv = ldv<Real, V>(s + idx), tv = ldv<Real, V>(t + idx);
		#pragma unroll
		for (int q = 0; q < V; q++) {
			xv.v[q] += alpha * sv.v[q];
			rv.v[q] += nalpha * tv.v[q];
			const double rd = (double)rv.v[q];
			if (useL2) nrm += rd * rd; else nrm = fmax(nrm, fabs(rd));
			if (MODE == 0) rr += (double)(rv.v[q] * rv.v[q]);
		}
		stv<Real, V>(x + idx, xv); stv<Real, V>(r + idx, rv);
	}
	double v[2] = { nrm, rr }; const bool isMax[2] = { !useL2, false }; double fin[2];
	if (blockReduceFinal<2>(v, isMax, partials, ticket, fin) && threadIdx.x == 0) {
		if (distLocal) { distLocal[0] = fin[0]; distLocal[1] = fin[1]; } else cgFinB<Real>(sc, fin[0], fin[1], MODE);
	}
}

// sigmaNew = z.r after a preconditioner application -> beta, sigma (conjugategrad.cpp:279-286); also used by doInit (:234)
template <typename Real, int V>
__global__ void __launch_bounds__(256) k_dot_zr(IndexInt n, const Real* __restrict__ z, const Real* __restrict__ r,
	CgScal<Real>* sc, double* partials, unsigned int* ticket, int isInit, double* distLocal)
{
	if (sc->done) return;
	const IndexInt nv = n / V;
	double acc = 0.0;
	for (IndexInt vi = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; vi < nv; vi += (IndexInt)gridDim.x * blockDim.x) {
		const VecT<Real, V> zv = ldv<Real, V>(z + vi * V), rv = ldv<Real, V>(r + vi * V);
		#pragma unroll
		for (int q = 0; q < V; q++) acc += (double)(zv.v[q] * rv.v[q]);
	}
	double v[1] = { acc }; const bool isMax[1] = { false }; double fin[1];
	if (blockReduceFinal<1>(v, isMax, partials, ticket, fin) && threadIdx.x == 0) {
		if (distLocal) distLocal[0] = fin[0]; else cgFinZR<Real>(sc, fin[0], isInit);
	}
}

// UpdateSearchVec conjugategrad.cpp:193-196: s = z + beta s
// Slab mode with peer memory: the threads that produce the first / last owned plane also store it into the neighbours'
// receive buffers (NVLink peer stores); the last block to finish publishes the planes with a system-scope release flag.
template <typename Real, int V>
__global__ void __launch_bounds__(256) k_update_search(IndexInt n, Real* __restrict__ s, const Real* __restrict__ z, const CgScal<Real>* sc,
	HaloOut ho, IndexInt plane)
{
	if (sc->done) return;
	const Real beta = sc->beta;
	const IndexInt nv = n / V;
	Real* const outLo = (Real*)ho.lo; Real* const outHi = (Real*)ho.hi;
	for (IndexInt vi = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; vi < nv; vi += (IndexInt)gridDim.x * blockDim.x) {
		VecT<Real, V> sv = ldv<Real, V>(s + vi * V); const VecT<Real, V> zv = ldv<Real, V>(z + vi * V);
		#pragma unroll
		for (int q = 0; q < V; q++) sv.v[q] = zv.v[q] + beta * sv.v[q];
		stv<Real, V>(s + vi * V, sv);
		const IndexInt idx = vi * V;
		if (outLo && idx < plane) stv<Real, V>(outLo + idx, sv);
		if (outHi && idx >= n - plane) stv<Real, V>(outHi + (idx - (n - plane)), sv);
	}
	if (ho.ticket) {
		__threadfence_system();
		__syncthreads();
		if (threadIdx.x == 0) {
			const unsigned int t = atomicAdd(ho.ticket, 1u);
			if (t == gridDim.x - 1) {
				*ho.ticket = 0;
				__threadfence_system();
				if (ho.flagLo) st_release_sys(ho.flagLo, ho.seq);
				if (ho.flagHi) st_release_sys(ho.flagHi, ho.seq);
			}
		}
	}
}

// doInit (conjugategrad.cpp:209-235) for PcNone: x = 0, r = b, s = b, sigma = b.b  (one pass)
template <typename Real, int V, int MODE>   // MODE 0: PcNone; MODE 1: only x = 0, r = b (preconditioner follows)
__global__ void __launch_bounds__(256) k_cg_init(IndexInt n, Real* __restrict__ x, const Real* __restrict__ b, Real* __restrict__ r, Real* __restrict__ s,
	CgScal<Real>* sc, double* partials, unsigned int* ticket, double* distLocal)
{
	const IndexInt nv = n / V;
	double acc = 0.0;
	VecT<Real, V> zero;
	#pragma unroll
	for (int q = 0; q < V; q++) zero.v[q] = (Real)0;
	for (IndexInt vi = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; vi < nv; vi += (IndexInt)gridDim.x * blockDim.x) {
		const VecT<Real, V> bv = ldv<Real, V>(b + vi * V);
		stv<Real, V>(x + vi * V, zero); stv<Real, V>(r + vi * V, bv);
		if (MODE == 0) {
			stv<Real, V>(s + vi * V, bv);
			#pragma unroll
			for (int q = 0; q < V; q++) acc += (double)(bv.v[q] * bv.v[q]);
		}
	}
	if (MODE != 0) return;
	double v[1] = { acc }; const bool isMax[1] = { false }; double fin[1];
	if (blockReduceFinal<1>(v, isMax, partials, ticket, fin) && threadIdx.x == 0) {
		if (distLocal) distLocal[0] = fin[0]; else cgFinZR<Real>(sc, fin[0], 1);
	}
}

template <typename Real>
__global__ void k_scal_reset(CgScal<Real>* sc, Real accuracy, int useL2) {
	sc->sigma = (Real)0; sc->alpha = (Real)0; sc->beta = (Real)0; sc->dp = (Real)0; sc->resNorm = (Real)1e20; sc->accuracy = accuracy;
	sc->iterations = 0; sc->done = 0; sc->diverged = 0; sc->useL2 = useL2; sc->xPending = 0;
}

// ================================================================ launch helpers
static inline int vecWidth(const mp_grid* g) {   // 16-byte vectors along x when every row start stays aligned
	const int V = (g->prec == 4) ? 4 : 2;
	return (g->sx % V == 0) ? V : 1;
}
static inline unsigned int streamBlocks(mp_context* ctx, IndexInt work) {
	unsigned int b = gridFor(work, 256);
	const unsigned int cap = (unsigned int)ctx->smCount * 8;      // 148 SMs x 8 CTAs of 256 threads = one full wave
	return b < cap ? b : cap;
}
static inline double* distLocalOf(mp_context* ctx, const Dims& d) { return d.world > 1 ? ctx->dist->dLocal : nullptr; }

#define DISPATCH_RV(g, ...) do { \
	const int V_ = vecWidth(g); \
	if ((g)->prec == 4) { typedef float Real; if (V_ == 4) { constexpr int V = 4; __VA_ARGS__; } else { constexpr int V = 1; __VA_ARGS__; } } \
	else { typedef double Real; if (V_ == 2) { constexpr int V = 2; __VA_ARGS__; } else { constexpr int V = 1; __VA_ARGS__; } } } while (0)

// slab mode: all-gather the ranks' partials and apply the scalar update of `stage` (0 alpha, 1 norm/beta, 2 z.r, 3 init sigma)
static int cgCombine(mp_context* ctx, const mp_grid* g, void* sc, int stage, int mode) {
	if (ctx->dist->p2p) MP_TRY(mp_dist_p2p_scalars(ctx));      // peer stores + flags instead of an NCCL all-gather
	else MP_TRY(mp_dist_allgather(ctx, 2));
	if (g->prec == 4) k_cg_combine<float><<<1, 1, 0, ctx->stream>>>(ctx->dist->dGather, ctx->dist->world, (CgScal<float>*)sc, stage, mode);
	else              k_cg_combine<double><<<1, 1, 0, ctx->stream>>>(ctx->dist->dGather, ctx->dist->world, (CgScal<double>*)sc, stage, mode);
	MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

int mp_launch_matvec(mp_context* ctx, const mp_grid* flags, mp_grid* dst, const mp_grid* src,
	const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak, void* sc, int finalize, const mp_grid* cmask = nullptr)
{
	const Dims d = dimsOf(flags);
	double* dl = (sc && finalize) ? distLocalOf(ctx, d) : nullptr;
	static const int variant = getenv("MP_MATVEC") ? atoi(getenv("MP_MATVEC")) : 1;    // 0: L2-reuse kernel, 1: z-marching kernel
	const int Vw = vecWidth(dst);
	if (variant == 1 && d.is3D && Vw > 1) {
		const int nvx = d.sx / Vw, planes = d.ke - d.kb;
		// z-chunks: enough CTAs for >= 4 waves of 148 SMs x 3 resident CTAs, chunks of >= 16 planes
		const int tiles = ((nvx + 31) / 32) * ((d.sy + 7) / 8);
		int nchunk = (4 * 3 * ctx->smCount + tiles - 1) / tiles; if (nchunk < 1) nchunk = 1;
		int chunk = (planes + nchunk - 1) / nchunk; if (chunk < 16) chunk = 16; if (chunk > planes) chunk = planes;
		nchunk = (planes + chunk - 1) / chunk;
		const dim3 grid((nvx + 31) / 32, (d.sy + 7) / 8, nchunk), block(32, 8, 1);
		if (cmask && (long long)grid.x * grid.y * grid.z <= kMaxPartials) {
			if (dst->prec == 4) k_matvec_zmarch_masked<float, 4><<<grid, block, 0, ctx->stream>>>(d, nvx, chunk, (const int*)cmask->d, (float*)dst->d, (const float*)src->d,
				(const float*)A0->d, (CgScal<float>*)sc, ctx->partials, ctx->tickets + 2, finalize, dl);
			else k_matvec_zmarch_masked<double, 2><<<grid, block, 0, ctx->stream>>>(d, nvx, chunk, (const int*)cmask->d, (double*)dst->d, (const double*)src->d,
				(const double*)A0->d, (CgScal<double>*)sc, ctx->partials, ctx->tickets + 2, finalize, dl);
			MP_CHECK_LAUNCH(ctx);
			ctx->lastMatvecKernel = 2;
			if (dl) MP_TRY(cgCombine(ctx, dst, sc, 0, 0));
			return MP_OK;
		}
		if ((long long)grid.x * grid.y * grid.z <= kMaxPartials) {
			if (dst->prec == 4) k_matvec_zmarch<float, 4><<<grid, block, 0, ctx->stream>>>(d, nvx, chunk, (const int*)flags->d, (float*)dst->d, (const float*)src->d,
				(const float*)A0->d, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, (CgScal<float>*)sc, ctx->partials, ctx->tickets + 2, finalize, dl);
			else k_matvec_zmarch<double, 2><<<grid, block, 0, ctx->stream>>>(d, nvx, chunk, (const int*)flags->d, (double*)dst->d, (const double*)src->d,
				(const double*)A0->d, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, (CgScal<double>*)sc, ctx->partials, ctx->tickets + 2, finalize, dl);
			MP_CHECK_LAUNCH(ctx);
			ctx->lastMatvecKernel = 1;
			if (dl) MP_TRY(cgCombine(ctx, dst, sc, 0, 0));
			return MP_OK;
		}
	}
	ctx->lastMatvecKernel = 0;
	DISPATCH_RV(dst, {
		const unsigned int blocks = streamBlocks(ctx, (d.i1 - d.i0) / V);
		if (d.is3D) k_matvec_dot<Real, V, true><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (Real*)dst->d, (const Real*)src->d,
			(const Real*)A0->d, (const Real*)Ai->d, (const Real*)Aj->d, (const Real*)Ak->d, (CgScal<Real>*)sc, ctx->partials, ctx->tickets + 2, finalize, dl);
		else k_matvec_dot<Real, V, false><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (Real*)dst->d, (const Real*)src->d,
			(const Real*)A0->d, (const Real*)Ai->d, (const Real*)Aj->d, (const Real*)Ak->d, (CgScal<Real>*)sc, ctx->partials, ctx->tickets + 2, finalize, dl);
	});
	MP_CHECK_LAUNCH(ctx);
	if (dl) MP_TRY(cgCombine(ctx, dst, sc, 0, 0));
	return MP_OK;
}

extern "C" int mp_apply_matrix(mp_context* ctx, const mp_grid* flags, mp_grid* dst, const mp_grid* src,
                               const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak)
{
	if (!ctx || !flags || !dst || !src || !A0 || !Ai || !Aj || !Ak) MP_FAIL(MP_ERR_INVALID, "mp_apply_matrix: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_apply_matrix: flags is not a FlagGrid");
	if (dst == src) MP_FAIL(MP_ERR_INVALID, "mp_apply_matrix: dst must not alias src");
	MP_TRY(mp_check_same(flags, dst, MP_GRID_REAL, "dst", false)); MP_TRY(mp_check_same(dst, src, MP_GRID_REAL, "src", false));
	MP_TRY(mp_check_same(dst, A0, MP_GRID_REAL, "A0", false)); MP_TRY(mp_check_same(dst, Ai, MP_GRID_REAL, "Ai", false));
	MP_TRY(mp_check_same(dst, Aj, MP_GRID_REAL, "Aj", false)); MP_TRY(mp_check_same(dst, Ak, MP_GRID_REAL, "Ak", false));
	MP_CUDA(cudaSetDevice(ctx->device));
	MP_TRY(mp_check_flags_interior(ctx, flags));
	// slab mode: the caller's src / Ak must have current ghost planes (mp_dist_exchange_halo)
	return mp_launch_matvec(ctx, flags, dst, src, A0, Ai, Aj, Ak, nullptr, 0);
}

// ================================================================ GridCg object
struct mp_cg {
	mp_context* ctx;
	mp_grid *dst, *rhs, *residual, *search, *tmp, *A0, *Ai, *Aj, *Ak; const mp_grid* flags;
	int pcMethod;                 // mp_cg_pc_type
	mp_grid* pcA0; mp_grid *pcAi = nullptr, *pcAj = nullptr, *pcAk = nullptr; mp_mg* mg;
	bool inited; bool useL2; double accuracy;
	void* dSc;                    // CgScal<Real> on the device
	CgScalHost* hSc;              // pinned mirror (2 slots for the lagged poll)
	cudaEvent_t pollEv[2];
	int iterations; double resNorm, sigma; bool diverged, finished;
	bool flagsChecked;
	mp_grid* cmask;               // coupling mask of the fast matvec path, or NULL (general kernel)
	bool fused;                   // PcNone on a masked matrix: two fused kernels per iteration, search vector ping-pongs between search / search2
	mp_grid* search2; long long fusedEnq;
	bool stepwise = false;           // driven through iterate(): x / residual / search are caller-visible after every call -> no fused loop
	int fusedNvx, fusedChunk; dim3 fusedGrid;
	FusedTma tma;                 // the TMA-staged persistent variant of the fused matvec (default when the grid qualifies)
};

template <typename Real>
__global__ void k_scal_export(const CgScal<Real>* sc, CgScalHost* out) {
	out->sigma = (double)sc->sigma; out->resNorm = (double)sc->resNorm; out->iterations = sc->iterations; out->done = sc->done; out->diverged = sc->diverged;
}

static int cgPollAsync(mp_cg* cg, int slot) {
	mp_context* ctx = cg->ctx;
	// the scalar block is exported by a 1-thread kernel straight into pinned (mapped) host memory, then an event marks it
	if (cg->dst->prec == 4) k_scal_export<float><<<1, 1, 0, ctx->stream>>>((const CgScal<float>*)cg->dSc, cg->hSc + slot);
	else                    k_scal_export<double><<<1, 1, 0, ctx->stream>>>((const CgScal<double>*)cg->dSc, cg->hSc + slot);
	MP_CHECK_LAUNCH(ctx);
	MP_CUDA(cudaEventRecord(cg->pollEv[slot], ctx->stream));
	return MP_OK;
}
static int cgPollWait(mp_cg* cg, int slot) {
	MP_CUDA(cudaEventSynchronize(cg->pollEv[slot]));
	const CgScalHost& h = cg->hSc[slot];
	cg->iterations = h.iterations; cg->resNorm = h.resNorm; cg->sigma = h.sigma; cg->diverged = h.diverged != 0; cg->finished = h.done != 0;
	return MP_OK;
}

int mp_mic_init_launch(mp_context* ctx, const mp_grid* flags, mp_grid* P, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak);
int mp_mic_check_stall(mp_context* ctx);
int mp_mic_apply_launch(mp_context* ctx, mp_grid* dst, const mp_grid* var1, const mp_grid* flags, const mp_grid* P,
                        const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak, const int* doneFlag);
int mp_ic_init_launch(mp_context* ctx, const mp_grid* flags, mp_grid* P0, mp_grid* Pi, mp_grid* Pj, mp_grid* Pk, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak);
int mp_ic_apply_launch(mp_context* ctx, mp_grid* dst, const mp_grid* var1, const mp_grid* flags, const mp_grid* P0, const mp_grid* Pi, const mp_grid* Pj, const mp_grid* Pk, const int* doneFlag);
int mp_mg_precond_init(mp_mg* mg, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak, double accuracy);
int mp_mg_precond_apply(mp_mg* mg, mp_grid* dst, const mp_grid* rhs, const int* doneFlag);

static int cgApplyPrecond(mp_cg* cg, const int* doneFlag) {
	mp_context* ctx = cg->ctx;
	if (cg->pcMethod == MP_CG_PC_MICP) return mp_mic_apply_launch(ctx, cg->tmp, cg->residual, cg->flags, cg->pcA0, cg->Ai, cg->Aj, cg->Ak, doneFlag);
	if (cg->pcMethod == MP_CG_PC_MGP) return mp_mg_precond_apply(cg->mg, cg->tmp, cg->residual, doneFlag);
	if (cg->pcMethod == MP_CG_PC_ICP) return mp_ic_apply_launch(ctx, cg->tmp, cg->residual, cg->flags, cg->pcA0, cg->pcAi, cg->pcAj, cg->pcAk, doneFlag);   // :257-258
	MP_FAIL(MP_ERR_UNSUPPORTED, "GridCg: preconditioner %d not implemented on the device", cg->pcMethod);
}

static int cgHalo(mp_cg* cg, mp_grid* g) {       // one-plane ghost exchange of a Real slab grid (no-op on a single GPU)
	return mp_dist_halo(cg->ctx, g->d, (size_t)g->sx * g->sy * g->prec, g->sz);
}

// tensor maps over residual / search / search2 / x / mask16 and the item decomposition of k_matvec_fused_tma
static const int kFusedStages = 6;
template <typename Real, int TY> static int cgFusedTmaSetupT(mp_cg* cg, const Dims& d) {
	typedef FusedTmaGeom<Real, TY> G;
	mp_context* ctx = cg->ctx; FusedTma& t = cg->tma;
	const CUtensorMapDataType dt = sizeof(Real) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
	const int es = (int)sizeof(Real);
	MP_TRY(encodeMap3D(&t.mapR, cg->residual->d, es, dt, d.sx, d.sx, d.sy, d.sz, G::BX, G::BY));
	MP_TRY(encodeMap3D(&t.mapS[0], cg->search->d, es, dt, d.sx, d.sx, d.sy, d.sz, G::BX, G::BY));
	MP_TRY(encodeMap3D(&t.mapS[1], cg->search2->d, es, dt, d.sx, d.sx, d.sy, d.sz, G::BX, G::BY));
	MP_TRY(encodeMap3D(&t.mapX, cg->dst->d, es, dt, d.sx, d.sx, d.sy, d.sz, G::TX, G::TY));
	MP_TRY(encodeMap3D(&t.mapM, t.mask16->d, 2, CU_TENSOR_MAP_DATA_TYPE_UINT16, d.sx, t.pitch, d.sy, d.sz, G::TX, G::TY));
	t.ty = TY;
	t.tilesX = (d.sx + G::TX - 1) / G::TX; t.tiles = t.tilesX * ((d.sy + G::TY - 1) / G::TY);
	t.smemBytes = kFusedStages * G::stageBytes + 2 * kFusedStages * 8;
	MP_CUDA(cudaFuncSetAttribute(k_matvec_fused_tma<Real, TY, kFusedStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, t.smemBytes));
	int perSm = 0;
	MP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_matvec_fused_tma<Real, TY, kFusedStages>, G::threads, t.smemBytes));
	if (perSm < 1) { t.on = false; return MP_OK; }
	if (perSm > G::ctasPerSm) perSm = G::ctasPerSm;
	const int G_ = ctx->smCount * perSm;
	fusedTmaDecompose(t.tiles, d.ke - d.kb, G_, &t.chunk, &t.nitems);
	t.ctas = t.nitems < G_ ? t.nitems : G_;
	if (t.ctas > kMaxPartials) t.on = false;
	return MP_OK;
}
static int cgFusedTmaSetup(mp_cg* cg, const Dims& d) {
	const int ty = getenv("MP_TMA_TY") ? atoi(getenv("MP_TMA_TY")) : 16;
	if (cg->dst->prec == 4) return ty == 8 ? cgFusedTmaSetupT<float, 8>(cg, d) : cgFusedTmaSetupT<float, 16>(cg, d);
	return ty == 8 ? cgFusedTmaSetupT<double, 8>(cg, d) : cgFusedTmaSetupT<double, 16>(cg, d);
}
template <typename Real, int TY> static void cgFusedTmaLaunch(mp_cg* cg, const Dims& d, bool even, Real* sNew, double* dl) {
	mp_context* ctx = cg->ctx; const FusedTma& t = cg->tma;
	k_matvec_fused_tma<Real, TY, kFusedStages><<<t.ctas, FusedTmaGeom<Real, TY>::threads, t.smemBytes, ctx->stream>>>(t.mapR, t.mapS[even ? 0 : 1], t.mapX, t.mapM, d,
		t.tilesX, t.tiles, t.chunk, t.nitems, (Real*)cg->tmp->d, sNew, (Real*)cg->dst->d, (const Real*)cg->A0->d, (CgScal<Real>*)cg->dSc, ctx->partials, ctx->tickets + 2, dl);
}

static int cgDoInit(mp_cg* cg) {    // doInit conjugategrad.cpp:209-235
	mp_context* ctx = cg->ctx;
	const Dims d = dimsOf(cg->flags);
	MP_TRY(mp_dist_check_grid(cg->flags));
	if (!cg->flagsChecked) { MP_TRY(mp_check_flags_interior(ctx, cg->flags)); cg->flagsChecked = true; }
	if (cg->pcMethod == MP_CG_PC_MICP && !d.is3D) MP_FAIL(MP_ERR_INVALID, "mICP only supports 3D grids so far");   // :222
	if (cg->pcMethod == MP_CG_PC_ICP && !d.is3D) MP_FAIL(MP_ERR_INVALID, "ICP only supports 3D grids so far");    // :218
	if (cg->pcMethod == MP_CG_PC_ICP && d.world > 1) MP_FAIL(MP_ERR_UNSUPPORTED, "GridCg: PC_ICP is not available on z-slab sharded grids");
	// Slab mode with a preconditioner: block-Jacobi over the slabs.  MIC(0) / GridMg are built and applied on
	// this rank's slab only -- the ghost planes are the local grid's outer layer (A0 == 0 there, so they are non-fluid for the
	// MIC sweeps and inactive vertices for GridMg), i.e. the preconditioner is the block diagonal of the global one.  It is
	// still symmetric positive definite, CG converges to the same solution; iteration counts differ and are reported.
	cg->inited = true; cg->iterations = 0; cg->finished = false; cg->diverged = false; cg->resNorm = 1e20;
	const bool none = cg->pcMethod == MP_CG_PC_NONE;
	const IndexInt nOwn = d.i1 - d.i0;
	double* dl = distLocalOf(ctx, d);
	if (d.world > 1) MP_TRY(mp_dist_p2p_prepare(ctx, 4 * (((size_t)d.sx * d.sy * cg->dst->prec + 15) / 16 * 16)));   // collective; no-op once built
	if (d.world > 1) MP_TRY(cgHalo(cg, cg->Ak));      // Ak[idx-Z] of the first owned plane lives on the lower neighbour
	{	// coupling-mask fast path of the matvec (3-D, vector-aligned rows, every off-diagonal in {0,-1})
		static const int useMask = getenv("MP_MATVEC_MASK") ? atoi(getenv("MP_MATVEC_MASK")) : 1;
		if (cg->cmask) { mp_grid_destroy(cg->cmask); cg->cmask = nullptr; }
		if (useMask && d.is3D && vecWidth(cg->dst) > 1) {
			MP_TRY(mp_grid_create(ctx, MP_GRID_FLAGS, 4, cg->flags->sx, cg->flags->sy, cg->flags->sz, &cg->cmask));
			int* bad = (int*)(ctx->dScal + 20);
			MP_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
			const unsigned int blocks = gridFor(d.i1 - d.i0, 256);
			// MP_CG_FUSED: 0 three-kernel loop, 1 fused matvec with plain loads, 2 (default) fused matvec staged by TMA
			const int fusedMode = getenv("MP_CG_FUSED") ? atoi(getenv("MP_CG_FUSED")) : 2;     // read per solve: the parity tests run all loops in one process
			unsigned short* m16 = nullptr; cg->tma.on = false; cg->tma.pitch = (cg->flags->sx + 7) / 8 * 8;
			if (fusedMode >= 2 && cg->pcMethod == MP_CG_PC_NONE && !cg->stepwise && encodeTiledFn() && cg->tma.pitch <= 2 * cg->flags->sx) {
				if (!cg->tma.mask16) MP_TRY(mp_grid_create(ctx, MP_GRID_FLAGS, 4, cg->flags->sx, cg->flags->sy, cg->flags->sz, &cg->tma.mask16));
				else MP_CUDA(cudaMemsetAsync(cg->tma.mask16->d, 0, cg->tma.mask16->bytes, ctx->stream));
				m16 = (unsigned short*)cg->tma.mask16->d; cg->tma.on = true;
			}
			if (cg->dst->prec == 4) k_build_cmask<float><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)cg->flags->d, (const float*)cg->A0->d, (const float*)cg->Ai->d, (const float*)cg->Aj->d, (const float*)cg->Ak->d, (int*)cg->cmask->d, m16, cg->tma.pitch, bad);
			else                    k_build_cmask<double><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)cg->flags->d, (const double*)cg->A0->d, (const double*)cg->Ai->d, (const double*)cg->Aj->d, (const double*)cg->Ak->d, (int*)cg->cmask->d, m16, cg->tma.pitch, bad);
			MP_CHECK_LAUNCH(ctx);
			MP_CUDA(cudaMemcpyAsync(ctx->hScal + 20, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
			MP_CUDA(cudaStreamSynchronize(ctx->stream));
			if (*(int*)(ctx->hScal + 20)) { mp_grid_destroy(cg->cmask); cg->cmask = nullptr; }     // face fractions: general kernel
		}
	}
	{	// fused PcNone loop: same launch geometry as the masked z-marching matvec
		const int useFused = getenv("MP_CG_FUSED") ? atoi(getenv("MP_CG_FUSED")) : 1;     // read per solve: the parity tests run both loops in one process
		cg->fused = false; cg->fusedEnq = 0;
		if (useFused && none && cg->cmask && !cg->stepwise) {
			const int Vw = vecWidth(cg->dst), nvx = d.sx / Vw, planes = d.ke - d.kb;
			const int tiles = ((nvx + 31) / 32) * ((d.sy + 7) / 8);
			int nchunk = (4 * 3 * ctx->smCount + tiles - 1) / tiles; if (nchunk < 1) nchunk = 1;
			int chunk = (planes + nchunk - 1) / nchunk; if (chunk < 16) chunk = 16; if (chunk > planes) chunk = planes;
			nchunk = (planes + chunk - 1) / chunk;
			const dim3 grid((nvx + 31) / 32, (d.sy + 7) / 8, nchunk);
			if ((long long)grid.x * grid.y * grid.z <= kMaxPartials) {
				cg->fused = true; cg->fusedNvx = nvx; cg->fusedChunk = chunk; cg->fusedGrid = grid;
				if (!cg->search2) MP_TRY(mp_grid_create(ctx, MP_GRID_REAL, cg->dst->prec, cg->dst->sx, cg->dst->sy, cg->dst->sz, &cg->search2));
				if (cg->tma.on) MP_TRY(cgFusedTmaSetup(cg, d));
			}
		}
	}
	DISPATCH_RV(cg->dst, {
		k_scal_reset<Real><<<1, 1, 0, ctx->stream>>>((CgScal<Real>*)cg->dSc, (Real)cg->accuracy, cg->useL2 ? 1 : 0);
		MP_CHECK_LAUNCH(ctx);
		const unsigned int blocks = streamBlocks(ctx, nOwn / V);
		Real *x = (Real*)cg->dst->d + d.i0, *r = (Real*)cg->residual->d + d.i0, *s = (Real*)cg->search->d + d.i0; const Real* b = (const Real*)cg->rhs->d + d.i0;
		if (none) k_cg_init<Real, V, 0><<<blocks, 256, 0, ctx->stream>>>(nOwn, x, b, r, s, (CgScal<Real>*)cg->dSc, ctx->partials, ctx->tickets + 3, dl);
		else      k_cg_init<Real, V, 1><<<blocks, 256, 0, ctx->stream>>>(nOwn, x, b, r, s, (CgScal<Real>*)cg->dSc, ctx->partials, ctx->tickets + 3, dl);
		MP_CHECK_LAUNCH(ctx);
	});
	if (none && dl) MP_TRY(cgCombine(ctx, cg->dst, cg->dSc, 3, 0));
	if (!none) {
		if (cg->pcMethod == MP_CG_PC_MICP) MP_TRY(mp_mic_init_launch(ctx, cg->flags, cg->pcA0, cg->A0, cg->Ai, cg->Aj, cg->Ak));
		else if (cg->pcMethod == MP_CG_PC_ICP) MP_TRY(mp_ic_init_launch(ctx, cg->flags, cg->pcA0, cg->pcAi, cg->pcAj, cg->pcAk, cg->A0, cg->Ai, cg->Aj, cg->Ak));   // :219
		else MP_TRY(mp_mg_precond_init(cg->mg, cg->A0, cg->Ai, cg->Aj, cg->Ak, cg->accuracy));
		MP_TRY(cgApplyPrecond(cg, nullptr));
		MP_TRY(mp_grid_copy_from(cg->search, cg->tmp));                     // mSearch.copyFrom(mTmp) :232
		DISPATCH_RV(cg->dst, {
			const unsigned int blocks = streamBlocks(ctx, nOwn / V);
			k_dot_zr<Real, V><<<blocks, 256, 0, ctx->stream>>>(nOwn, (const Real*)cg->tmp->d + d.i0, (const Real*)cg->residual->d + d.i0, (CgScal<Real>*)cg->dSc, ctx->partials, ctx->tickets + 4, 1, dl);
			MP_CHECK_LAUNCH(ctx);
		});
		if (dl) MP_TRY(cgCombine(ctx, cg->dst, cg->dSc, 3, 0));
	}
	if (d.world > 1) MP_TRY(cgHalo(cg, cg->search));  // the first matvec reads the neighbours' planes of s
	if (d.world > 1 && cg->fused) MP_TRY(cgHalo(cg, cg->residual));   // ... and, fused, forms s = r + beta s_old on them
	return MP_OK;
}

static int cgEnqueueIteration(mp_cg* cg, int iterIndex = -1) {   // iterate conjugategrad.cpp:237-299, no host synchronisation
	mp_context* ctx = cg->ctx;
	const Dims d = dimsOf(cg->flags);
	const bool none = cg->pcMethod == MP_CG_PC_NONE;
	const IndexInt nOwn = d.i1 - d.i0;
	double* dl = distLocalOf(ctx, d);
	const size_t planeBytes = (size_t)d.sx * d.sy * cg->dst->prec;
	const bool p2pHalo = dl && ctx->dist->p2p && planeBytes % 16 == 0;
	// sampled kernel timing: 5 events around the 4 stages of this iteration
	cudaEvent_t* pe = nullptr;
	if (ctx->profPeriod > 0 && iterIndex >= 0 && iterIndex % ctx->profPeriod == 0 && ctx->profCount < 128) pe = &ctx->profEv[5 * ctx->profCount++];
	#define PROF(k) do { if (pe) MP_CUDA(cudaEventRecord(pe[k], ctx->stream)); } while (0)
	PROF(0);
	if (cg->fused) {
		const bool even = (cg->fusedEnq & 1) == 0; cg->fusedEnq++;
		mp_grid* sOld = even ? cg->search : cg->search2; mp_grid* sNew = even ? cg->search2 : cg->search;
		const dim3 block(32, 8, 1);
		HaloOut ho;
		if (p2pHalo) MP_TRY(mp_dist_p2p_halo_out(ctx, planeBytes, &ho));
		DISPATCH_RV(cg->dst, {
			CgScal<Real>* sc = (CgScal<Real>*)cg->dSc;
			if (cg->tma.on) {
				if (cg->tma.ty == 8) cgFusedTmaLaunch<Real, 8>(cg, d, even, (Real*)sNew->d, dl); else cgFusedTmaLaunch<Real, 16>(cg, d, even, (Real*)sNew->d, dl);
				ctx->lastMatvecKernel = 4;
			} else {
				k_matvec_fused<Real, V><<<cg->fusedGrid, block, 0, ctx->stream>>>(d, cg->fusedNvx, cg->fusedChunk, (const int*)cg->cmask->d, (Real*)cg->tmp->d,
					(const Real*)sOld->d, (Real*)sNew->d, (const Real*)cg->residual->d, (Real*)cg->dst->d, (const Real*)cg->A0->d, sc, ctx->partials, ctx->tickets + 2, dl);
				ctx->lastMatvecKernel = 3;
			}
			MP_CHECK_LAUNCH(ctx);
			if (dl) MP_TRY(cgCombine(ctx, cg->dst, cg->dSc, 0, 0));
			PROF(1);
			const unsigned int blocks = streamBlocks(ctx, nOwn / V);
			k_axpy1_norm<Real, V><<<blocks, 256, 0, ctx->stream>>>(nOwn, (Real*)cg->residual->d + d.i0, (const Real*)cg->tmp->d + d.i0, sc, ctx->partials, ctx->tickets + 5, dl,
				ho, (IndexInt)d.sx * d.sy);
			MP_CHECK_LAUNCH(ctx);
			if (dl) MP_TRY(cgCombine(ctx, cg->dst, cg->dSc, 1, 0));
			if (p2pHalo) MP_TRY(mp_dist_p2p_halo_in(ctx, cg->residual->d, planeBytes, cg->residual->sz, &sc->done));
			else if (dl) MP_TRY(cgHalo(cg, cg->residual));
		});
		PROF(2); PROF(3); PROF(4);
		return MP_OK;
	}
	MP_TRY(mp_launch_matvec(ctx, cg->flags, cg->tmp, cg->search, cg->A0, cg->Ai, cg->Aj, cg->Ak, cg->dSc, 1, cg->cmask));
	PROF(1);
	DISPATCH_RV(cg->dst, {
		const unsigned int blocks = streamBlocks(ctx, nOwn / V);
		CgScal<Real>* sc = (CgScal<Real>*)cg->dSc;
		Real *x = (Real*)cg->dst->d + d.i0, *r = (Real*)cg->residual->d + d.i0, *s = (Real*)cg->search->d + d.i0, *t = (Real*)cg->tmp->d + d.i0;
		if (none) {
			k_axpy2_norm<Real, V, 0><<<blocks, 256, 0, ctx->stream>>>(nOwn, x, s, r, t, sc, ctx->partials, ctx->tickets + 5, dl);
			MP_CHECK_LAUNCH(ctx);
			if (dl) MP_TRY(cgCombine(ctx, cg->dst, cg->dSc, 1, 0));
			PROF(2); PROF(3);
			HaloOut ho;
			if (p2pHalo) MP_TRY(mp_dist_p2p_halo_out(ctx, planeBytes, &ho));
			k_update_search<Real, V><<<blocks, 256, 0, ctx->stream>>>(nOwn, s, r, sc, ho, (IndexInt)d.sx * d.sy);
			MP_CHECK_LAUNCH(ctx);
			if (p2pHalo) MP_TRY(mp_dist_p2p_halo_in(ctx, cg->search->d, planeBytes, cg->search->sz, &sc->done));
			else if (dl) MP_TRY(cgHalo(cg, cg->search));
		} else {
			k_axpy2_norm<Real, V, 1><<<blocks, 256, 0, ctx->stream>>>(nOwn, x, s, r, t, sc, ctx->partials, ctx->tickets + 5, dl);
			MP_CHECK_LAUNCH(ctx);
			if (dl) MP_TRY(cgCombine(ctx, cg->dst, cg->dSc, 1, 1));
			PROF(2);
			MP_TRY(cgApplyPrecond(cg, &sc->done));
			k_dot_zr<Real, V><<<blocks, 256, 0, ctx->stream>>>(nOwn, t, r, sc, ctx->partials, ctx->tickets + 4, 0, dl);
			MP_CHECK_LAUNCH(ctx);
			if (dl) MP_TRY(cgCombine(ctx, cg->dst, cg->dSc, 2, 0));
			PROF(3);
			HaloOut ho;
			if (p2pHalo) MP_TRY(mp_dist_p2p_halo_out(ctx, planeBytes, &ho));
			k_update_search<Real, V><<<blocks, 256, 0, ctx->stream>>>(nOwn, s, t, sc, ho, (IndexInt)d.sx * d.sy);
			MP_CHECK_LAUNCH(ctx);
			if (p2pHalo) MP_TRY(mp_dist_p2p_halo_in(ctx, cg->search->d, planeBytes, cg->search->sz, &sc->done));
			else if (dl) MP_TRY(cgHalo(cg, cg->search));
		}
	});
	PROF(4);
	#undef PROF
	return MP_OK;
}

// Fused loop: x lags one update behind (k_matvec_fused applies alpha_{k-1} s_{k-1} while forming s_k); k_flush_x applies the pending
// update and clears xPending, so a following fused matvec adds alpha_prev = 0.
static int cgFlushFused(mp_cg* cg) {
	mp_context* ctx = cg->ctx;
	const Dims d = dimsOf(cg->flags);
	const IndexInt nOwn = d.i1 - d.i0;
	DISPATCH_RV(cg->dst, {
		k_flush_x<Real, V><<<streamBlocks(ctx, nOwn / V), 256, 0, ctx->stream>>>(nOwn, (Real*)cg->dst->d + d.i0, (const Real*)cg->search->d + d.i0,
			(const Real*)cg->search2->d + d.i0, (CgScal<Real>*)cg->dSc, ctx->tickets + 8);
		MP_CHECK_LAUNCH(ctx);
	});
	return MP_OK;
}
// Stepwise callers (GridCg::iterate driven by solvePressureSystem pressure.cpp:436-439, cgSolveWE, the VIC solve) see x, residual and
// the search grid after EVERY call.  The fused loop is one phase behind the reference there: it forms s_k = r_k + beta_k s_{k-1} at the
// START of iteration k, the reference at the end of iteration k-1 (UpdateSearchVec conjugategrad.cpp:283).  Leaving the fused loop
// therefore applies the pending x-update and that search-vector update into the caller's search grid (same operands, same bits); the
// three-kernel loop continues from there.
static int cgUnfuse(mp_cg* cg) {
	mp_context* ctx = cg->ctx;
	const Dims d = dimsOf(cg->flags);
	const IndexInt nOwn = d.i1 - d.i0;
	MP_TRY(cgFlushFused(cg));
	// enqueued fused iteration q (0-based) wrote its search vector to search2 when q is even
	if (cg->fusedEnq & 1) MP_CUDA(cudaMemcpyAsync(cg->search->d, cg->search2->d, cg->search->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	DISPATCH_RV(cg->dst, {
		k_update_search<Real, V><<<streamBlocks(ctx, nOwn / V), 256, 0, ctx->stream>>>(nOwn, (Real*)cg->search->d + d.i0, (const Real*)cg->residual->d + d.i0,
			(const CgScal<Real>*)cg->dSc, HaloOut(), (IndexInt)d.sx * d.sy);
		MP_CHECK_LAUNCH(ctx);
	});
	if (d.world > 1) MP_TRY(cgHalo(cg, cg->search));
	cg->fused = false;
	return MP_OK;
}

static int cgFinishCheck(mp_cg* cg) {
	if (cg->diverged) MP_FAIL(MP_ERR_DIVERGED, "GridCg::iterate: The CG solver diverged, residual norm > 1e30, stopping.");
	return MP_OK;
}

int mp_cg_run(mp_cg* cg, int maxIter) {
	mp_context* ctx = cg->ctx;
	MP_CUDA(cudaSetDevice(ctx->device));
	if (!cg->inited) MP_TRY(cgDoInit(cg));
	if (cg->finished) return cgFinishCheck(cg);
	ctx->profCount = 0;
	// batches of iterations; the poll of batch b is awaited only after batch b+1 has been enqueued.  In slab mode every
	// rank sees the same scalars at the same polls, so all ranks enqueue the same sequence of kernels and NCCL calls.
	const IndexInt n = cg->dst->n;
	int batch = n >= (IndexInt)1 << 24 ? 4 : (n >= (IndexInt)1 << 21 ? 8 : 16);
	if (cg->pcMethod != MP_CG_PC_NONE) batch = (batch + 3) / 4;
	int enq = 0, slot = 0; bool pending[2] = { false, false };
	while (enq < maxIter) {
		const int nb = (maxIter - enq) < batch ? (maxIter - enq) : batch;
		for (int q = 0; q < nb; q++) MP_TRY(cgEnqueueIteration(cg, enq + q));
		enq += nb;
		MP_TRY(cgPollAsync(cg, slot)); pending[slot] = true;
		const int other = slot ^ 1;
		if (pending[other]) { MP_TRY(cgPollWait(cg, other)); pending[other] = false; if (cg->finished) break; }
		slot = other;
	}
	for (int q = 0; q < 2; q++) if (pending[q]) { MP_TRY(cgPollWait(cg, q)); pending[q] = false; }
	if (cg->fused) MP_TRY(cgFlushFused(cg));      // the x-update of the last executed iteration is still pending
	MP_TRY(cgPollAsync(cg, 0)); MP_TRY(cgPollWait(cg, 0));      // the state after everything that was enqueued
	MP_TRY(mp_dist_p2p_check(ctx));
	if (cg->pcMethod == MP_CG_PC_MICP) MP_TRY(mp_mic_check_stall(ctx));
	if (ctx->profPeriod > 0) {
		// only samples of iterations that really ran (before `done`) count
		int used = 0; double acc[4] = {0, 0, 0, 0};
		for (int q = 0; q < ctx->profCount; q++) {
			if (q * ctx->profPeriod >= cg->iterations) break;
			for (int k = 0; k < 4; k++) { float ms = 0; cudaEventElapsedTime(&ms, ctx->profEv[5 * q + k], ctx->profEv[5 * q + k + 1]); acc[k] += ms; }
			used++;
		}
		for (int k = 0; k < 4; k++) ctx->profMs[k] = used ? (float)(acc[k] / used) : 0.f;
		ctx->profCount = used;
	}
	return cgFinishCheck(cg);
}

extern "C" {

int mp_cg_create(mp_context* ctx, mp_grid* dst, mp_grid* rhs, mp_grid* residual, mp_grid* search, const mp_grid* flags, mp_grid* tmp,
                 mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak, mp_cg** out)
{
	if (!ctx || !dst || !rhs || !residual || !search || !flags || !tmp || !A0 || !Ai || !Aj || !Ak || !out) MP_FAIL(MP_ERR_INVALID, "mp_cg_create: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_cg_create: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, dst, MP_GRID_REAL, "dst", false));
	const mp_grid* gs[] = { rhs, residual, search, tmp, A0, Ai, Aj, Ak }; const char* nm[] = { "rhs", "residual", "search", "tmp", "A0", "Ai", "Aj", "Ak" };
	for (int q = 0; q < 8; q++) MP_TRY(mp_check_same(dst, gs[q], MP_GRID_REAL, nm[q], false));
	MP_CUDA(cudaSetDevice(ctx->device));
	mp_cg* cg = new mp_cg();
	cg->ctx = ctx; cg->dst = dst; cg->rhs = rhs; cg->residual = residual; cg->search = search; cg->flags = flags; cg->tmp = tmp;
	cg->A0 = A0; cg->Ai = Ai; cg->Aj = Aj; cg->Ak = Ak;
	cg->pcMethod = MP_CG_PC_NONE; cg->pcA0 = nullptr; cg->mg = nullptr;
	cg->inited = false; cg->useL2 = true;            // GridCgInterface() : mUseL2Norm(true), conjugategrad.h:31
	cg->accuracy = dst->prec == 4 ? (double)1e-6f : 1e-10;   // mAccuracy(VECTOR_EPSILON) conjugategrad.cpp:206, vectorbase.h:52,:55
	cg->iterations = 0; cg->resNorm = 1e20; cg->sigma = 0; cg->diverged = false; cg->finished = false; cg->flagsChecked = false; cg->cmask = nullptr; cg->fused = false; cg->search2 = nullptr; cg->fusedEnq = 0;
	MP_CUDA(cudaMalloc(&cg->dSc, 256));
	MP_CUDA(cudaHostAlloc((void**)&cg->hSc, sizeof(CgScalHost) * 2, cudaHostAllocMapped));
	memset(cg->hSc, 0, sizeof(CgScalHost) * 2);
	MP_CUDA(cudaEventCreateWithFlags(&cg->pollEv[0], cudaEventDisableTiming));
	MP_CUDA(cudaEventCreateWithFlags(&cg->pollEv[1], cudaEventDisableTiming));
	*out = cg; return MP_OK;
}
int mp_cg_destroy(mp_cg* cg) {
	if (!cg) return MP_OK;
	cudaSetDevice(cg->ctx->device);
	cudaStreamSynchronize(cg->ctx->stream);
	cudaFree(cg->dSc); cudaFreeHost(cg->hSc); cudaEventDestroy(cg->pollEv[0]); cudaEventDestroy(cg->pollEv[1]);
	if (cg->cmask) mp_grid_destroy(cg->cmask);
	if (cg->search2) mp_grid_destroy(cg->search2);
	if (cg->tma.mask16) mp_grid_destroy(cg->tma.mask16);
	delete cg; return MP_OK;
}
int mp_cg_set_accuracy(mp_cg* cg, double accuracy) { cg->accuracy = accuracy; return MP_OK; }
int mp_cg_set_use_l2_norm(mp_cg* cg, int useL2) { cg->useL2 = useL2 != 0; return MP_OK; }
int mp_cg_set_ic_preconditioner(mp_cg* cg, int method, mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak) {
	// conjugategrad.cpp:310-326.  The reference asserts method is PC_ICP|PC_mICP, which makes PcNone unreachable from
	// solvePressure in 3-D; north_star requires PcNone, so PC_None is accepted here and documented.
	if (method != MP_CG_PC_NONE && method != MP_CG_PC_ICP && method != MP_CG_PC_MICP)
		MP_FAIL(MP_ERR_INVALID, "GridCg<APPLYMAT>::setICPreconditioner: Invalid method specified.");
	if (method != MP_CG_PC_NONE) {
		if (!A0) MP_FAIL(MP_ERR_INVALID, "setICPreconditioner: preconditioner grid A0 is NULL");
		MP_TRY(mp_check_same(cg->dst, A0, MP_GRID_REAL, "pcA0", false));
	}
	if (method == MP_CG_PC_ICP) {       // IC(0) keeps a scaled copy of the whole matrix (":25 needs 4 add. grids")
		if (!Ai || !Aj || !Ak) MP_FAIL(MP_ERR_INVALID, "setICPreconditioner: PC_ICP needs the four preconditioner grids A0, Ai, Aj, Ak");
		MP_TRY(mp_check_same(cg->dst, Ai, MP_GRID_REAL, "pcAi", false)); MP_TRY(mp_check_same(cg->dst, Aj, MP_GRID_REAL, "pcAj", false)); MP_TRY(mp_check_same(cg->dst, Ak, MP_GRID_REAL, "pcAk", false));
	}
	cg->pcMethod = method;
	if (method != MP_CG_PC_NONE && cg->dst->sz == 1) cg->pcMethod = MP_CG_PC_NONE;   // "only supported in 3D for now, disabling it" :315-321
	cg->pcA0 = A0; cg->pcAi = Ai; cg->pcAj = Aj; cg->pcAk = Ak;
	return MP_OK;
}
int mp_cg_set_mg_preconditioner(mp_cg* cg, int method, mp_mg* mg) {
	if (method != MP_CG_PC_MGP) MP_FAIL(MP_ERR_INVALID, "GridCg<APPLYMAT>::setMGPreconditioner: Invalid method specified.");   // :330
	if (!mg) MP_FAIL(MP_ERR_INVALID, "setMGPreconditioner: mg is NULL");
	cg->pcMethod = method; cg->mg = mg; return MP_OK;
}
int mp_cg_force_reinit(mp_cg* cg) { cg->inited = false; cg->finished = false; cg->stepwise = false; return MP_OK; }

int mp_cg_iterate(mp_cg* cg, int* keepGoing) {
	mp_context* ctx = cg->ctx;
	MP_CUDA(cudaSetDevice(ctx->device));
	if (!cg->inited) { cg->stepwise = true; MP_TRY(cgDoInit(cg)); }      // stepwise from the start: three-kernel loop
	if (cg->finished) { if (keepGoing) *keepGoing = 0; return cgFinishCheck(cg); }     // converged: nothing is enqueued any more
	if (cg->fused) MP_TRY(cgUnfuse(cg));                                   // a solve() ran before: leave the fused loop's lagged state
	MP_TRY(cgEnqueueIteration(cg));
	MP_TRY(cgPollAsync(cg, 0)); MP_TRY(cgPollWait(cg, 0));
	if (keepGoing) *keepGoing = cg->finished ? 0 : 1;      // once converged, further calls are no-ops returning false
	return cgFinishCheck(cg);
}
int mp_cg_solve(mp_cg* cg, int maxIter) { return mp_cg_run(cg, maxIter); }
int mp_cg_get(mp_cg* cg, int* iterations, double* resNorm, double* sigma) {
	if (iterations) *iterations = cg->iterations; if (resNorm) *resNorm = cg->resNorm; if (sigma) *sigma = cg->sigma; return MP_OK;
}

} // extern "C"

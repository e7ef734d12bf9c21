// Modified incomplete Cholesky MIC(0) preconditioner on the device.
//   InitPreconditionModifiedIncompCholesky2   conjugategrad.cpp:66-97   (serial k,j,i sweep in the reference)
//   ApplyPreconditionModifiedIncompCholesky2  conjugategrad.cpp:135-159 (forward + backward substitution, serial)
//
// The lexicographic sweeps are re-scheduled on hyperplanes i+j+k = c: every cell of plane c depends only on
// plane c-1 (forward) / c+1 (backward), so each plane is one data-parallel step and the per-cell arithmetic
// (operation order, -fmad=false, IEEE div/sqrt) is IDENTICAL to the serial reference -> bit-identical
// Aprecond / z and therefore the reference's iteration counts.
// v1 schedule: one launch per plane (sx+sy+sz-8 launches per sweep); the kernels are dependency-/launch-bound,
// not bandwidth-bound -- DESIGN.md states the stage counts.
#include "mp_common.cuh"

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
	const Real tau = (Real)0.97, sigma = (Real)0.25;
	const IndexInt ix = idx - 1, iy = idx - g.Y, iz = idx - g.Z;
	const Real px = P[ix], py = P[iy], pz = P[iz];
	const Real aix = Ai[ix], ajy = Aj[iy], akz = Ak[iz];
	const Real tx = aix * px, ty = ajy * py, tz = akz * pz;
	Real e = A0[idx] - tx * tx - ty * ty - tz * tz;
	const Real inner = aix * (Aj[ix] + Ak[ix]) * (px * px) + ajy * (Ai[iy] + Ak[iy]) * (py * py) + akz * (Ai[iz] + Aj[iz]) * (pz * pz);
	// conjugategrad.cpp:85-89: the "+ 0." promotes bracket, product and subtraction to double
	e = (Real)((double)e - (double)tau * ((double)inner + 0.));
	if (e < sigma * A0[idx]) e = A0[idx];
	P[idx] = (Real)(1. / (double)sqrt(e));       // sqrt in Real (std::sqrt(float) overload), divide in double (:95)
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

int mp_mic_init_launch(mp_context* ctx, const mp_grid* flags, mp_grid* P, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak)
{
	const Dims d = dimsOf(flags);
	if (!d.is3D) MP_FAIL(MP_ERR_INVALID, "mICP only supports 3D grids so far");
	MP_CUDA(cudaMemsetAsync(P->d, 0, P->bytes, ctx->stream));          // Aprecond.clear() :71
	if (d.sx < 3 || d.sy < 3 || d.sz < 3) return MP_OK;
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
	MP_CUDA(cudaSetDevice(ctx->device));
	MP_TRY(mp_check_flags_interior(ctx, flags));
	return mp_mic_apply_launch(ctx, dst, var1, flags, Aprecond, Ai, Aj, Ak, nullptr);
}

}

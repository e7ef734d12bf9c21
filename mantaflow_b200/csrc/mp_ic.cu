// IC(0) "a la Wavelet Turbulence" (GridCgInterface::PC_ICP): InitPreconditionIncompCholesky conjugategrad.cpp:26-63 and
// ApplyPreconditionIncompCholesky conjugategrad.cpp:109-132 -- the preconditioner of the VIC Poisson solve (vortexplugins.cpp:266-285).
//
// The reference loops run serially in k / j / i order; cell (i,j,k) depends on (i-1,j,k), (i,j-1,k), (i,j,k-1) only, so all cells of a
// hyperplane i + j + k = c are independent: one launch per hyperplane (the schedule of the first MIC version, mp_mic.cu "v1"; the
// pipelined column schedules of MIC(0) are the next step for this preconditioner, the VIC scenes are small).
// The factorisation is written as a gather: the reference scatters `A0(i+1,j,k) -= square(Ai[idx])` etc. to the +x/+y/+z neighbours
// (also onto non-fluid cells); here a cell subtracts the squares of its -z, -y, -x fluid neighbours in the order the serial loop
// applies them, which gives the same bits in all four factor grids (checked against the reference's goldens, `icp_*`).
#include "mp_ic_cells.cuh"

namespace {

typedef ic::Geom IcGeom;
// cell of plane c addressed by (j, k) = (1 + blockIdx.x * blockDim.x + threadIdx.x, klo + blockIdx.y)
__device__ __forceinline__ bool icCell(const IcGeom& g, int c, int klo, IndexInt& idx) {
	return ic::planeCell(g, c, 1 + (int)(blockIdx.x * blockDim.x + threadIdx.x), klo + (int)blockIdx.y, idx);
}
struct IcLaunch { int klo; dim3 grid; };
inline bool icLaunch(const IcGeom& g, int c, IcLaunch& pl) {
	ic::PlaneRange r;
	if (!ic::planeRange(g, c, r)) return false;
	pl.klo = r.klo; pl.grid = dim3((unsigned)((g.hy + 127) / 128), (unsigned)(r.khi - r.klo + 1), 1);
	return true;
}

template <typename Real>
__global__ void __launch_bounds__(128) k_ic_init_plane(IcGeom g, int c, int klo, const int* __restrict__ flags, Real* P0, Real* Pi, Real* Pj, Real* Pk,
	const Real* __restrict__ A0, const Real* __restrict__ Ai, const Real* __restrict__ Aj, const Real* __restrict__ Ak)
{
	IndexInt idx;
	if (icCell(g, c, klo, idx)) ic::initCell<Real>(flags, P0, Pi, Pj, Pk, A0, Ai, Aj, Ak, idx, g.Y, g.Z);
}

template <typename Real>
__global__ void __launch_bounds__(128) k_ic_fwd_plane(IcGeom g, int c, int klo, const int* __restrict__ flags, Real* dst, const Real* __restrict__ src,
	const Real* __restrict__ P0, const Real* __restrict__ Pi, const Real* __restrict__ Pj, const Real* __restrict__ Pk, const int* __restrict__ doneFlag)
{
	if (doneFlag && *doneFlag) return;
	IndexInt idx;
	if (icCell(g, c, klo, idx)) ic::fwdCell<Real>(flags, dst, src, P0, Pi, Pj, Pk, idx, g.Y, g.Z);
}

template <typename Real>
__global__ void __launch_bounds__(128) k_ic_bwd_plane(IcGeom g, int c, int klo, const int* __restrict__ flags, Real* dst,
	const Real* __restrict__ P0, const Real* __restrict__ Pi, const Real* __restrict__ Pj, const Real* __restrict__ Pk, const int* __restrict__ doneFlag)
{
	if (doneFlag && *doneFlag) return;
	IndexInt idx;
	if (icCell(g, c, klo, idx)) ic::bwdCell<Real>(flags, dst, P0, Pi, Pj, Pk, idx, g.Y, g.Z);
}

template <typename Real>
int icInit(mp_context* ctx, const Dims& d, const int* flags, Real* P0, Real* Pi, Real* Pj, Real* Pk, const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak) {
	const IcGeom g = { d.sx, d.sy, d.sz, d.Y, d.Z, d.sx - 1, d.sy - 1, d.sz - 1 };      // the scatter of a fluid cell reaches the outer layer
	for (int c = 3; c <= g.hx + g.hy + g.hz; c++) {
		IcLaunch pl; if (!icLaunch(g, c, pl)) continue;
		k_ic_init_plane<Real><<<pl.grid, 128, 0, ctx->stream>>>(g, c, pl.klo, flags, P0, Pi, Pj, Pk, A0, Ai, Aj, Ak);
		MP_CHECK_LAUNCH(ctx);
	}
	return MP_OK;
}
template <typename Real>
int icApply(mp_context* ctx, const Dims& d, const int* flags, Real* dst, const Real* src, const Real* P0, const Real* Pi, const Real* Pj, const Real* Pk, const int* doneFlag) {
	const IcGeom g = { d.sx, d.sy, d.sz, d.Y, d.Z, d.sx - 2, d.sy - 2, d.sz - 2 };      // fluid cells are interior cells
	const int cmax = g.hx + g.hy + g.hz;
	for (int c = 3; c <= cmax; c++) {
		IcLaunch pl; if (!icLaunch(g, c, pl)) continue;
		k_ic_fwd_plane<Real><<<pl.grid, 128, 0, ctx->stream>>>(g, c, pl.klo, flags, dst, src, P0, Pi, Pj, Pk, doneFlag);
		MP_CHECK_LAUNCH(ctx);
	}
	for (int c = cmax; c >= 3; c--) {
		IcLaunch pl; if (!icLaunch(g, c, pl)) continue;
		k_ic_bwd_plane<Real><<<pl.grid, 128, 0, ctx->stream>>>(g, c, pl.klo, flags, dst, P0, Pi, Pj, Pk, doneFlag);
		MP_CHECK_LAUNCH(ctx);
	}
	return MP_OK;
}

}  // namespace

// called by GridCg (mp_cg.cu) and by the entry points below; arguments are checked by the callers
int mp_ic_init_launch(mp_context* ctx, const mp_grid* flags, mp_grid* P0, mp_grid* Pi, mp_grid* Pj, mp_grid* Pk, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak)
{
	const Dims d = dimsOf(flags);
	if (!d.is3D) MP_FAIL(MP_ERR_INVALID, "ICP only supports 3D grids so far");                      // conjugategrad.cpp:218
	if (ctx->dist && ctx->dist->active) MP_FAIL(MP_ERR_UNSUPPORTED, "GridCg: PC_ICP is not available on z-slab sharded grids");
	const mp_grid* src[4] = { A0, Ai, Aj, Ak }; mp_grid* dst[4] = { P0, Pi, Pj, Pk };
	for (int q = 0; q < 4; q++) MP_CUDA(cudaMemcpyAsync(dst[q]->d, src[q]->d, src[q]->bytes, cudaMemcpyDeviceToDevice, ctx->stream));      // A0.copyFrom(orgA0) ... :31-34
	if (d.sx < 3 || d.sy < 3 || d.sz < 3) return MP_OK;
	if (P0->prec == 4) return icInit<float>(ctx, d, (const int*)flags->d, (float*)P0->d, (float*)Pi->d, (float*)Pj->d, (float*)Pk->d, (const float*)A0->d, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d);
	return icInit<double>(ctx, d, (const int*)flags->d, (double*)P0->d, (double*)Pi->d, (double*)Pj->d, (double*)Pk->d, (const double*)A0->d, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d);
}

int mp_ic_apply_launch(mp_context* ctx, mp_grid* dst, const mp_grid* var1, const mp_grid* flags, const mp_grid* P0, const mp_grid* Pi, const mp_grid* Pj, const mp_grid* Pk, const int* doneFlag)
{
	const Dims d = dimsOf(flags);
	if (!d.is3D) MP_FAIL(MP_ERR_INVALID, "ICP only supports 3D grids so far");
	if (d.sx < 3 || d.sy < 3 || d.sz < 3) return MP_OK;
	if (dst->prec == 4) return icApply<float>(ctx, d, (const int*)flags->d, (float*)dst->d, (const float*)var1->d, (const float*)P0->d, (const float*)Pi->d, (const float*)Pj->d, (const float*)Pk->d, doneFlag);
	return icApply<double>(ctx, d, (const int*)flags->d, (double*)dst->d, (const double*)var1->d, (const double*)P0->d, (const double*)Pi->d, (const double*)Pj->d, (const double*)Pk->d, doneFlag);
}

extern "C" {

int mp_ic_init(mp_context* ctx, const mp_grid* flags, mp_grid* P0, mp_grid* Pi, mp_grid* Pj, mp_grid* Pk, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak)
{
	if (!ctx || !flags || !P0 || !Pi || !Pj || !Pk || !A0 || !Ai || !Aj || !Ak) MP_FAIL(MP_ERR_INVALID, "mp_ic_init: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_ic_init: flags is not a FlagGrid");
	const mp_grid* all[8] = { P0, Pi, Pj, Pk, A0, Ai, Aj, Ak };
	MP_TRY(mp_check_same(flags, P0, MP_GRID_REAL, "A0 (preconditioner)", false));
	for (int q = 1; q < 8; q++) MP_TRY(mp_check_same(P0, all[q], MP_GRID_REAL, "matrix / preconditioner grid", false));
	for (int q = 0; q < 4; q++) for (int r = 4; r < 8; r++) if (all[q]->d == all[r]->d) MP_FAIL(MP_ERR_INVALID, "mp_ic_init: the preconditioner grids must not alias the matrix");
	MP_CUDA(cudaSetDevice(ctx->device));
	MP_TRY(mp_check_flags_interior(ctx, flags));
	return mp_ic_init_launch(ctx, flags, P0, Pi, Pj, Pk, A0, Ai, Aj, Ak);
}

int mp_ic_apply(mp_context* ctx, mp_grid* dst, const mp_grid* var1, const mp_grid* flags, const mp_grid* P0, const mp_grid* Pi, const mp_grid* Pj, const mp_grid* Pk)
{
	if (!ctx || !flags || !dst || !var1 || !P0 || !Pi || !Pj || !Pk) MP_FAIL(MP_ERR_INVALID, "mp_ic_apply: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_ic_apply: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, dst, MP_GRID_REAL, "dst", false)); MP_TRY(mp_check_same(dst, var1, MP_GRID_REAL, "var1", false));
	MP_TRY(mp_check_same(dst, P0, MP_GRID_REAL, "A0", false)); MP_TRY(mp_check_same(dst, Pi, MP_GRID_REAL, "Ai", false));
	MP_TRY(mp_check_same(dst, Pj, MP_GRID_REAL, "Aj", false)); MP_TRY(mp_check_same(dst, Pk, MP_GRID_REAL, "Ak", false));
	if (dst == var1 || dst->d == var1->d) MP_FAIL(MP_ERR_INVALID, "mp_ic_apply: dst must not alias var1");
	MP_CUDA(cudaSetDevice(ctx->device));
	MP_TRY(mp_check_flags_interior(ctx, flags));
	return mp_ic_apply_launch(ctx, dst, var1, flags, P0, Pi, Pj, Pk, nullptr);
}

}

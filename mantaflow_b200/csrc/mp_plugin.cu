// The pressure plugins on device-resident grids: computePressureRhs, solvePressureSystem, correctVelocity,
// solvePressure, releaseMG (plugin/pressure.cpp:252-521) and the host-buffer entry point that a maintainer's
// pressure.cpp calls while grids have no device mirror.
#include "mp_common.cuh"
#include <cmath>

int mp_make_matrix_fused(mp_context* ctx, const mp_grid* flags, mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak, const mp_grid* fractions, const mp_grid* phi, double gfClamp);
int mp_fix_pressure_auto(mp_context* ctx, const mp_grid* flags, mp_grid* rhs, mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak);
int mp_cg_run(mp_cg* cg, int maxIter);
void mp_mg_invalidate(mp_mg* mg);
bool mp_mg_matches(const mp_mg* mg, int prec, int sx, int sy, int sz);

template <typename Real>
__global__ void __launch_bounds__(256) k_add_mean_corr(Real* __restrict__ rhs, IndexInt n, const double* sumcnt) {
	// rhs += (Real)(-sum / (Real)cnt)  on ALL cells (pressure.cpp:297-298); sum,cnt come from k_make_rhs on the device
	const Real corr = (Real)(-sumcnt[0] / (double)(Real)(int)sumcnt[1]);
	for (IndexInt i = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (IndexInt)gridDim.x * blockDim.x) rhs[i] += corr;
}

// cgSolveDiffusion matrix, conjugategrad.cpp:360-375: MakeLaplaceMatrix on an all-fluid dummy FlagGrid (interior cells get
// A0 = 2*dim, Ai = Aj = Ak = -1), then obstacle rows -> identity, all other rows -> I + alpha * L
template <typename Real>
__global__ void __launch_bounds__(256) k_diffusion_matrix(Dims d, const int* __restrict__ flags, Real alpha, Real* __restrict__ A0, Real* __restrict__ Ai,
	Real* __restrict__ Aj, Real* __restrict__ Ak)
{
	const IndexInt idx = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= d.n) return;
	int i, j, k; cellOf(d, idx, i, j, k);
	const bool interior = i >= 1 && i < d.sx - 1 && j >= 1 && j < d.sy - 1 && (!d.is3D || (k >= 1 && k < d.sz - 1));
	Real a0 = 0, ai = 0, aj = 0, ak = 0;
	if (interior) { a0 = d.is3D ? (Real)6 : (Real)4; ai = (Real)-1; aj = (Real)-1; if (d.is3D) ak = (Real)-1; }
	if (flags[idx] & TypeObstacle) { ai = aj = ak = (Real)0.0; a0 = (Real)1.0; }
	else { ai *= alpha; aj *= alpha; ak *= alpha; a0 *= alpha; a0 = (Real)((double)a0 + 1.); }
	A0[idx] = a0; Ai[idx] = ai; Aj[idx] = aj; Ak[idx] = ak;
}
// cgSolveWE waves.cpp:112-118: A *= s, A0 += 1 on every cell;  MakeRhsWE :72-80 on the interior (the stencil of the Crank-Nicolson term is
// the reference's: x and y neighbours only, also in 3-D)
template <typename Real>
__global__ void __launch_bounds__(256) k_we_setup(Dims d, Real s, Real* __restrict__ A0, Real* __restrict__ Ai, Real* __restrict__ Aj, Real* __restrict__ Ak,
	Real* __restrict__ rhs, const Real* __restrict__ ut, const Real* __restrict__ utm1, int crankNic)
{
	const IndexInt idx = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= d.n) return;
	int i, j, k; cellOf(d, idx, i, j, k);
	Ai[idx] = Ai[idx] * s; Aj[idx] = Aj[idx] * s; Ak[idx] = Ak[idx] * s;
	A0[idx] = (Real)((double)(A0[idx] * s) + 1.);
	const bool interior = i >= 1 && i < d.sx - 1 && j >= 1 && j < d.sy - 1 && (!d.is3D || (k >= 1 && k < d.sz - 1));
	Real r = 0;
	if (interior) {
		r = (Real)(2. * (double)ut[idx] - (double)utm1[idx]);
		if (crankNic) r = (Real)((double)r + (double)s * (-4. * (double)ut[idx] + 1. * (double)ut[idx - d.X] + 1. * (double)ut[idx + d.X] + 1. * (double)ut[idx - d.Y] + 1. * (double)ut[idx + d.Y]));
	}
	rhs[idx] = r;
}
// knGetComponent / knSetComponent grid.cpp:676-685
template <typename Real>
__global__ void __launch_bounds__(256) k_component(IndexInt n, Real* __restrict__ vec, Real* __restrict__ comp, int c, int set) {
	for (IndexInt i = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (IndexInt)gridDim.x * blockDim.x) {
		if (set) vec[3 * i + c] = comp[i]; else comp[i] = vec[3 * i + c];
	}
}

// VICintegration's grid half (plugin/vortexplugins.cpp:253-299): CurlOp commonkernels.h:38-47 (bnd = 1, 3-D) and GetShiftedComponent :104-108 /
// GetComponent :111-113.  The factors 0.5 are exact in either precision, so `0.5 * Real` evaluated in double and narrowed equals the Real product.
template <typename Real>
__global__ void __launch_bounds__(256) k_vic_curl(Dims d, const Real* __restrict__ w, Real* __restrict__ curl)
{
	const IndexInt idx = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= d.n) return;
	int i, j, k; cellOf(d, idx, i, j, k);
	Real v0 = 0, v1 = 0, v2 = 0;
	if (i >= 1 && i < d.sx - 1 && j >= 1 && j < d.sy - 1 && k >= 1 && k < d.sz - 1) {
		const IndexInt X = 3, Y = 3 * d.Y, Z = 3 * d.Z; const Real* c = w + 3 * idx;
		v2 = (Real)0.5 * ((c[X + 1] - c[-X + 1]) - (c[Y + 0] - c[-Y + 0]));
		v0 = (Real)0.5 * ((c[Y + 2] - c[-Y + 2]) - (c[Z + 1] - c[-Z + 1]));
		v1 = (Real)0.5 * ((c[Z + 0] - c[-Z + 0]) - (c[X + 2] - c[-X + 2]));
	}
	curl[3 * idx] = v0; curl[3 * idx + 1] = v1; curl[3 * idx + 2] = v2;        // outside bnd the fresh grid keeps its zeros
}
template <typename Real>
__global__ void __launch_bounds__(256) k_vic_rhs(Dims d, const Real* __restrict__ curl, Real* __restrict__ rhs, int c, int shifted)
{
	const IndexInt idx = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= d.n) return;
	if (!shifted) { rhs[idx] = curl[3 * idx + c]; return; }
	int i, j, k; cellOf(d, idx, i, j, k);
	Real r = 0;                                                               // bnd = 1: the border of the (fresh) rhs grid stays zero
	if (i >= 1 && i < d.sx - 1 && j >= 1 && j < d.sy - 1 && k >= 1 && k < d.sz - 1) {
		const IndexInt sh = c == 0 ? 1 : (c == 1 ? d.Y : d.Z);
		r = (Real)0.5 * (curl[3 * idx + c] + curl[3 * (idx - sh) + c]);
	}
	rhs[idx] = r;
}
// solution *= scale; SetComponent(vel, solution, c) :297-298
template <typename Real>
__global__ void __launch_bounds__(256) k_vic_store(IndexInt n, Real* __restrict__ sol, Real* __restrict__ vel, Real scale, int c)
{
	const IndexInt idx = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= n) return;
	const Real v = sol[idx] * scale;
	sol[idx] = v; vel[3 * idx + c] = v;
}

struct GridHolder {   // RAII for the temp grids of one call (the reference takes them from the FluidSolver pool, pressure.cpp:331-338)
	std::vector<mp_grid*> gs;
	~GridHolder() { for (mp_grid* g : gs) mp_grid_destroy(g); }
	int make(mp_context* ctx, int kind, int prec, const mp_grid* like, mp_grid** out) {
		int rc = mp_grid_create(ctx, kind, prec, like->sx, like->sy, like->sz, out);
		if (rc == MP_OK) gs.push_back(*out);
		return rc;
	}
};

extern "C" {

int mp_release_mg(mp_context* ctx) {
	// releaseMG pressure.cpp:252-266.  The hierarchy is invalidated (isASet = false); its device allocations are parked in
	// the context so that the next `new GridMg(size)` of the same size costs no cudaMalloc/cudaFree (PcMGDynamic does this
	// every solve, pressure.cpp:423-429,:451).
	if (!ctx || !ctx->staticMg) return MP_OK;
	if (ctx->spareMg) mp_mg_destroy(ctx->spareMg);
	ctx->spareMg = ctx->staticMg; ctx->staticMg = nullptr;
	mp_mg_invalidate(ctx->spareMg);
	return MP_OK;
}

int mp_compute_pressure_rhs(mp_context* ctx, mp_grid* rhs, const mp_grid* vel, const mp_grid* pressure, const mp_grid* flags,
                            const mp_grid* phi, const mp_grid* perCellCorr, const mp_grid* fractions, const mp_grid* obvel,
                            const mp_grid* curv, const mp_pressure_params* params)
{
	(void)pressure;
	mp_pressure_params def; if (!params) { mp_pressure_params_default(&def); params = &def; }
	MP_TRY(mp_make_rhs(ctx, flags, rhs, vel, perCellCorr, fractions, obvel, phi, curv, params->surfTens, params->gfClamp, nullptr, nullptr));
	if (params->enforceCompatibility) {
		unsigned int blocks = gridFor(rhs->n, 256 * 4);
		if (rhs->prec == 4) k_add_mean_corr<float><<<blocks, 256, 0, ctx->stream>>>((float*)rhs->d, rhs->n, ctx->dScal + 2);
		else                k_add_mean_corr<double><<<blocks, 256, 0, ctx->stream>>>((double*)rhs->d, rhs->n, ctx->dScal + 2);
		MP_CHECK_LAUNCH(ctx);
	}
	return MP_OK;
}

int mp_solve_pressure_system(mp_context* ctx, mp_grid* rhs, mp_grid* vel, mp_grid* pressure, const mp_grid* flags,
                             const mp_grid* phi, const mp_grid* perCellCorr, const mp_grid* fractions,
                             const mp_grid* curv, const mp_pressure_params* params, mp_solve_info* info)
{
	(void)vel; (void)perCellCorr; (void)curv;
	if (!ctx || !rhs || !pressure || !flags) MP_FAIL(MP_ERR_INVALID, "mp_solve_pressure_system: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_solve_pressure_system: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, rhs, MP_GRID_REAL, "rhs", false));
	MP_TRY(mp_check_same(rhs, pressure, MP_GRID_REAL, "pressure", false));
	MP_TRY(mp_check_same(rhs, phi, MP_GRID_REAL, "phi", true));
	MP_TRY(mp_check_same(rhs, fractions, MP_GRID_MAC, "fractions", true));
	mp_pressure_params def; if (!params) { mp_pressure_params_default(&def); params = &def; }
	MP_CUDA(cudaSetDevice(ctx->device));
	int preconditioner = params->preconditioner;
	if (!params->precondition) preconditioner = MP_PC_NONE;                       // pressure.cpp:328
	if (preconditioner < MP_PC_NONE || preconditioner > MP_PC_MG_STATIC) MP_FAIL(MP_ERR_INVALID, "solvePressureSystem: invalid preconditioner %d", preconditioner);
	const int prec = rhs->prec;
	const bool is3D = flags->sz > 1;

	GridHolder tmp;
	mp_grid *residual, *search, *A0, *Ai, *Aj, *Ak, *t;                           // :331-338
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, rhs, &residual)); MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, rhs, &search));
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, rhs, &A0)); MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, rhs, &Ai));
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, rhs, &Aj)); MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, rhs, &Ak));
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, rhs, &t));

	MP_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
	MP_TRY(mp_make_matrix_fused(ctx, flags, A0, Ai, Aj, Ak, fractions, phi, params->gfClamp));   // :341-345

	long long fixed = -1;
	const bool wantFix = params->zeroPressureFixing || (prec == 4 ? (double)(float)params->cgAccuracy : params->cgAccuracy) < 1e-07;   // :349
	if (wantFix) MP_TRY(mp_fix_pressure_auto(ctx, flags, rhs, A0, Ai, Aj, Ak));
	MP_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));

	mp_cg* cg = nullptr;
	MP_TRY(mp_cg_create(ctx, pressure, rhs, residual, search, flags, t, A0, Ai, Aj, Ak, &cg));   // :392-396
	struct CgGuard { mp_cg* c; ~CgGuard() { mp_cg_destroy(c); } } cgGuard{ cg };
	mp_cg_set_accuracy(cg, prec == 4 ? (double)(float)params->cgAccuracy : params->cgAccuracy);
	mp_cg_set_use_l2_norm(cg, params->useL2Norm);

	int maxIter = 0;
	mp_grid* pca0 = nullptr; mp_mg* pmg = nullptr;
	if (preconditioner == MP_PC_NONE || preconditioner == MP_PC_MIC) {
		const int maxDim = std::max(flags->sx, std::max(flags->sy, flags->sz));
		// (int)(cgMaxIterFac * size.max()) with Real arithmetic (:408)
		if (prec == 4) maxIter = (int)((float)params->cgMaxIterFac * (float)maxDim) * (is3D ? 1 : 4);
		else           maxIter = (int)(params->cgMaxIterFac * (double)maxDim) * (is3D ? 1 : 4);
		if (preconditioner == MP_PC_MIC) MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, rhs, &pca0));
		MP_TRY(mp_cg_set_ic_preconditioner(cg, preconditioner == MP_PC_MIC ? MP_CG_PC_MICP : MP_CG_PC_NONE, pca0, nullptr, nullptr, nullptr));
	} else {
		maxIter = 100;                                                            // :419
		pmg = ctx->staticMg;
		if (pmg && preconditioner == MP_PC_MG_DYNAMIC) { mp_release_mg(ctx); pmg = nullptr; }   // :423-426
		// a kept hierarchy belongs to one grid size and precision (gMapMG is per FluidSolver, and a FluidSolver has one grid size): a caller
		// that projects other grids in the same context gets a fresh hierarchy instead of V-cycles with the old geometry
		if (pmg && !mp_mg_matches(pmg, prec, flags->sx, flags->sy, flags->sz)) { mp_release_mg(ctx); pmg = nullptr; }
		if (!pmg) {
			if (ctx->spareMg && mp_mg_matches(ctx->spareMg, prec, flags->sx, flags->sy, flags->sz)) { pmg = ctx->spareMg; ctx->spareMg = nullptr; }
			else MP_TRY(mp_mg_create(ctx, prec, flags->sx, flags->sy, flags->sz, &pmg));
			ctx->staticMg = pmg;
		}
		MP_TRY(mp_cg_set_mg_preconditioner(cg, MP_CG_PC_MGP, pmg));
	}

	int rc = mp_cg_run(cg, maxIter);                                              // :436-439
	MP_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
	if (info) {
		mp_cg_get(cg, &info->iterations, &info->resNorm, nullptr);
		info->maxIter = maxIter;
		if (wantFix) {
			cudaMemcpyAsync(ctx->hScal + 12, ctx->dScal + 12, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
			cudaStreamSynchronize(ctx->stream);
			fixed = *(long long*)(ctx->hScal + 12);
		}
		info->fixedCell = fixed;
		info->mgLevels = 0;
		if (pmg) mp_mg_num_levels(pmg, &info->mgLevels);
		cudaEventSynchronize(ctx->ev[3]);
		cudaEventElapsedTime(&info->msMatrix, ctx->ev[1], ctx->ev[2]);
		cudaEventElapsedTime(&info->msSolve, ctx->ev[2], ctx->ev[3]);
		if (ctx->profPeriod > 0) {
			info->msMatvecAvg = ctx->profMs[0]; info->msAxpyAvg = ctx->profMs[1]; info->msPrecondAvg = ctx->profMs[2]; info->msUpdateAvg = ctx->profMs[3];
			info->profSamples = ctx->profCount;
		}
		info->matvecKernel = ctx->lastMatvecKernel;
	}
	if (pmg && preconditioner == MP_PC_MG_DYNAMIC) mp_release_mg(ctx);            // :451
	return rc;
}

int mp_solve_pressure(mp_context* ctx, mp_grid* vel, mp_grid* pressure, const mp_grid* flags,
                      const mp_grid* phi, const mp_grid* perCellCorr, const mp_grid* fractions, const mp_grid* obvel,
                      const mp_grid* curv, mp_grid* retRhs, const mp_pressure_params* params, mp_solve_info* info)
{
	if (!ctx || !vel || !pressure || !flags) MP_FAIL(MP_ERR_INVALID, "mp_solve_pressure: NULL argument");
	mp_pressure_params def; if (!params) { mp_pressure_params_default(&def); params = &def; }
	MP_CUDA(cudaSetDevice(ctx->device));
	if (info) { memset(info, 0, sizeof *info); info->fixedCell = -1; }
	GridHolder tmp; mp_grid* rhs;
	MP_TRY(mp_check_same(flags, pressure, MP_GRID_REAL, "pressure", false));
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, pressure->prec, pressure, &rhs));          // Grid<Real> rhs(parent) :497
	MP_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
	MP_TRY(mp_compute_pressure_rhs(ctx, rhs, vel, pressure, flags, phi, perCellCorr, fractions, obvel, curv, params));
	MP_TRY(mp_solve_pressure_system(ctx, rhs, vel, pressure, flags, phi, perCellCorr, fractions, curv, params, info));
	MP_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
	MP_TRY(mp_correct_velocity(ctx, vel, pressure, flags, phi, curv, params));
	if (retRhs) { MP_TRY(mp_check_same(pressure, retRhs, MP_GRID_REAL, "retRhs", false)); MP_TRY(mp_grid_copy_from(retRhs, rhs)); }   // :518-520
	MP_CUDA(cudaEventRecord(ctx->ev[4], ctx->stream));
	MP_CUDA(cudaEventSynchronize(ctx->ev[4]));
	if (info) {
		cudaEventElapsedTime(&info->msRhs, ctx->ev[0], ctx->ev[1]);
		cudaEventElapsedTime(&info->msCorrect, ctx->ev[3], ctx->ev[4]);
		cudaEventElapsedTime(&info->msTotal, ctx->ev[0], ctx->ev[4]);
	}
	return MP_OK;
}

int mp_cg_solve_diffusion(mp_context* ctx, const mp_grid* flags, mp_grid* grid, double alpha, double cgMaxIterFac, double cgAccuracy, mp_solve_info* info)
{
	if (!ctx || !flags || !grid) MP_FAIL(MP_ERR_INVALID, "mp_cg_solve_diffusion: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_cg_solve_diffusion: flags is not a FlagGrid");
	if (grid->kind != MP_GRID_REAL && grid->kind != MP_GRID_MAC)
		MP_FAIL(MP_ERR_INVALID, "cgSolveDiffusion: Grid Type is not supported (only Real, Vec3, MAC, or Levelset)");       // conjugategrad.cpp:419
	if (grid->sx != flags->sx || grid->sy != flags->sy || grid->sz != flags->sz) MP_FAIL(MP_ERR_INVALID, "mp_cg_solve_diffusion: grid and flags differ in size");
	if (ctx->dist && ctx->dist->active) MP_FAIL(MP_ERR_UNSUPPORTED, "mp_cg_solve_diffusion: not sharded across GPUs");
	MP_CUDA(cudaSetDevice(ctx->device));
	const int prec = grid->prec;
	const Dims d = dimsOf(flags);
	GridHolder tmp;
	mp_grid *rhs, *residual, *search, *t, *A0, *Ai, *Aj, *Ak, *u = nullptr;
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &rhs)); MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &residual));
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &search)); MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &t));
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &A0)); MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &Ai));
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &Aj)); MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &Ak));
	const unsigned int blocks = gridFor(d.n, 256);
	if (prec == 4) k_diffusion_matrix<float><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (float)alpha, (float*)A0->d, (float*)Ai->d, (float*)Aj->d, (float*)Ak->d);
	else           k_diffusion_matrix<double><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, alpha, (double*)A0->d, (double*)Ai->d, (double*)Aj->d, (double*)Ak->d);
	MP_CHECK_LAUNCH(ctx);
	const int maxDim = std::max(flags->sx, std::max(flags->sy, flags->sz));
	const int maxIter = (prec == 4 ? (int)((float)cgMaxIterFac * (float)maxDim) : (int)(cgMaxIterFac * (double)maxDim)) * (d.is3D ? 1 : 4);   // :379
	const bool isVec = grid->kind == MP_GRID_MAC;
	if (isVec) MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &u)); else u = grid;
	mp_cg* cg = nullptr;
	MP_TRY(mp_cg_create(ctx, u, rhs, residual, search, flags, t, A0, Ai, Aj, Ak, &cg));           // no preconditioner, L2 norm: GridCg defaults
	struct CgGuard { mp_cg* c; ~CgGuard() { mp_cg_destroy(c); } } guard{ cg };
	mp_cg_set_accuracy(cg, prec == 4 ? (double)(float)cgAccuracy : cgAccuracy);
	const int ncomp = isVec ? (d.is3D ? 3 : 2) : 1;
	const unsigned int cb = gridFor(d.n, 256 * 4);
	for (int c = 0; c < ncomp; c++) {
		if (isVec) {
			if (prec == 4) k_component<float><<<cb, 256, 0, ctx->stream>>>(d.n, (float*)grid->d, (float*)u->d, c, 0);
			else           k_component<double><<<cb, 256, 0, ctx->stream>>>(d.n, (double*)grid->d, (double*)u->d, c, 0);
			MP_CHECK_LAUNCH(ctx);
			mp_cg_force_reinit(cg);
		}
		MP_TRY(mp_grid_copy_from(rhs, u));                                                     // rhs.copyFrom(u) :383,:412
		MP_TRY(mp_cg_run(cg, maxIter));
		if (isVec) {
			if (prec == 4) k_component<float><<<cb, 256, 0, ctx->stream>>>(d.n, (float*)grid->d, (float*)u->d, c, 1);
			else           k_component<double><<<cb, 256, 0, ctx->stream>>>(d.n, (double*)grid->d, (double*)u->d, c, 1);
			MP_CHECK_LAUNCH(ctx);
		}
	}
	if (info) { memset(info, 0, sizeof *info); info->fixedCell = -1; mp_cg_get(cg, &info->iterations, &info->resNorm, nullptr); info->maxIter = maxIter; info->matvecKernel = ctx->lastMatvecKernel; }
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	return MP_OK;
}

// The grid half of VICintegration plugin/vortexplugins.cpp:253-299 -- from the vorticity grid the Peskin kernel (:203-250, mesh code) leaves to
// the velocity: MakeLaplaceMatrix, CurlOp, then per component rhs = (shifted) component of the curl, GridCg<ApplyMatrix> with the L2 stop test
// and PreconditionType(precondition), solution *= scale, SetComponent.  As in the reference, setICPreconditioner accepts PC_ICP (1) and
// PC_mICP (2) only (conjugategrad.cpp:312): the plugin's default precondition = 0 is an error there and here.
int mp_vic_poisson(mp_context* ctx, const mp_grid* flags, const mp_grid* vorticity, mp_grid* vel, int velIsMac, double cgMaxIterFac, double cgAccuracy,
                   double scale, int precondition, int* iterations)
{
	if (!ctx || !flags || !vorticity || !vel) MP_FAIL(MP_ERR_INVALID, "mp_vic_poisson: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_vic_poisson: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, vorticity, MP_GRID_MAC, "vorticity", false)); MP_TRY(mp_check_same(flags, vel, MP_GRID_MAC, "vel", false));
	if (vorticity->prec != vel->prec) MP_FAIL(MP_ERR_INVALID, "mp_vic_poisson: vorticity and vel differ in precision");
	if (vorticity == vel) MP_FAIL(MP_ERR_INVALID, "mp_vic_poisson: vorticity and vel must be two grids");
	if (flags->sz <= 1) MP_FAIL(MP_ERR_UNSUPPORTED, "VICintegration: 3-D grids only (GridCg<ApplyMatrix>)");
	if (precondition != MP_CG_PC_ICP && precondition != MP_CG_PC_MICP) MP_FAIL(MP_ERR_INVALID, "GridCg<APPLYMAT>::setICPreconditioner: Invalid method specified.");
	if (ctx->dist && ctx->dist->active) MP_FAIL(MP_ERR_UNSUPPORTED, "mp_vic_poisson: not sharded across GPUs");
	MP_CUDA(cudaSetDevice(ctx->device));
	MP_TRY(mp_check_flags_interior(ctx, flags));
	const int prec = vel->prec;
	const Dims d = dimsOf(flags);
	GridHolder tmp;
	mp_grid *curl, *rhs, *sol, *residual, *search, *t, *A0, *Ai, *Aj, *Ak, *P0, *P1, *P2, *P3;
	MP_TRY(tmp.make(ctx, MP_GRID_MAC, prec, flags, &curl));
	mp_grid** reals[] = { &rhs, &sol, &residual, &search, &t, &A0, &Ai, &Aj, &Ak, &P0, &P1, &P2, &P3 };
	for (mp_grid** g : reals) MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, g));
	MP_TRY(mp_make_laplace_matrix(ctx, flags, A0, Ai, Aj, Ak, nullptr));
	const unsigned int blocks = gridFor(d.n, 256);
	if (prec == 4) k_vic_curl<float><<<blocks, 256, 0, ctx->stream>>>(d, (const float*)vorticity->d, (float*)curl->d);
	else           k_vic_curl<double><<<blocks, 256, 0, ctx->stream>>>(d, (const double*)vorticity->d, (double*)curl->d);
	MP_CHECK_LAUNCH(ctx);
	const int maxDim = std::max(flags->sx, std::max(flags->sy, flags->sz));
	const int maxIter = prec == 4 ? (int)((float)cgMaxIterFac * (float)maxDim) : (int)(cgMaxIterFac * (double)maxDim);      // :273
	for (int c = 0; c < 3; c++) {
		if (prec == 4) k_vic_rhs<float><<<blocks, 256, 0, ctx->stream>>>(d, (const float*)curl->d, (float*)rhs->d, c, velIsMac ? 1 : 0);
		else           k_vic_rhs<double><<<blocks, 256, 0, ctx->stream>>>(d, (const double*)curl->d, (double*)rhs->d, c, velIsMac ? 1 : 0);
		MP_CHECK_LAUNCH(ctx);
		mp_cg* cg = nullptr;                                                                       // a new GridCg per component :274
		MP_TRY(mp_cg_create(ctx, sol, rhs, residual, search, flags, t, A0, Ai, Aj, Ak, &cg));
		struct CgGuard { mp_cg* c; ~CgGuard() { mp_cg_destroy(c); } } guard{ cg };
		mp_cg_set_accuracy(cg, prec == 4 ? (double)(float)cgAccuracy : cgAccuracy);
		mp_cg_set_use_l2_norm(cg, 1);
		MP_TRY(mp_cg_set_ic_preconditioner(cg, precondition, P0, P1, P2, P3));
		MP_TRY(mp_cg_run(cg, maxIter));
		if (iterations) mp_cg_get(cg, &iterations[c], nullptr, nullptr);
		if (prec == 4) k_vic_store<float><<<blocks, 256, 0, ctx->stream>>>(d.n, (float*)sol->d, (float*)vel->d, (float)scale, c);
		else           k_vic_store<double><<<blocks, 256, 0, ctx->stream>>>(d.n, (double*)sol->d, (double*)vel->d, scale, c);
		MP_CHECK_LAUNCH(ctx);
	}
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	return MP_OK;
}

// cgSolveWE plugin/waves.cpp:86-147: one implicit step of the wave equation on the device GridCg (no preconditioner, GridCg's default stop test)
int mp_cg_solve_we(mp_context* ctx, const mp_grid* flags, mp_grid* ut, mp_grid* utm1, mp_grid* out, int crankNic, double cSqr, double cgMaxIterFac,
                   double cgAccuracy, double dt, mp_solve_info* info)
{
	if (!ctx || !flags || !ut || !utm1 || !out) MP_FAIL(MP_ERR_INVALID, "mp_cg_solve_we: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_cg_solve_we: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, ut, MP_GRID_REAL, "ut", false)); MP_TRY(mp_check_same(ut, utm1, MP_GRID_REAL, "utm1", false)); MP_TRY(mp_check_same(ut, out, MP_GRID_REAL, "out", false));
	if (ut == utm1 || ut == out || utm1 == out) MP_FAIL(MP_ERR_INVALID, "mp_cg_solve_we: ut, utm1 and out must be three grids");
	if (ctx->dist && ctx->dist->active) MP_FAIL(MP_ERR_UNSUPPORTED, "mp_cg_solve_we: not sharded across GPUs");
	MP_CUDA(cudaSetDevice(ctx->device));
	const int prec = ut->prec;
	const Dims d = dimsOf(flags);
	GridHolder tmp;
	mp_grid *rhs, *residual, *search, *t, *A0, *Ai, *Aj, *Ak;
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &rhs)); MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &residual));
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &search)); MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &t));
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &A0)); MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &Ai));
	MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &Aj)); MP_TRY(tmp.make(ctx, MP_GRID_REAL, prec, flags, &Ak));
	MP_CUDA(cudaMemsetAsync(out->d, 0, out->bytes, ctx->stream));                                   // out.clear() :106
	MP_TRY(mp_make_laplace_matrix(ctx, flags, A0, Ai, Aj, Ak, nullptr));
	const unsigned int blocks = gridFor(d.n, 256);
	if (prec == 4) {
		const float fdt = (float)dt, s = (float)((double)(fdt * fdt * (float)cSqr) * 0.5);           // :111
		k_we_setup<float><<<blocks, 256, 0, ctx->stream>>>(d, s, (float*)A0->d, (float*)Ai->d, (float*)Aj->d, (float*)Ak->d, (float*)rhs->d, (const float*)ut->d, (const float*)utm1->d, crankNic);
	} else {
		const double s = (dt * dt * cSqr) * 0.5;
		k_we_setup<double><<<blocks, 256, 0, ctx->stream>>>(d, s, (double*)A0->d, (double*)Ai->d, (double*)Aj->d, (double*)Ak->d, (double*)rhs->d, (const double*)ut->d, (const double*)utm1->d, crankNic);
	}
	MP_CHECK_LAUNCH(ctx);
	const int maxDim = std::max(flags->sx, std::max(flags->sy, flags->sz));
	const int maxIter = (prec == 4 ? (int)((float)cgMaxIterFac * (float)maxDim) : (int)(cgMaxIterFac * (double)maxDim)) * (d.is3D ? 1 : 4);   // :129
	mp_cg* cg = nullptr;
	MP_TRY(mp_cg_create(ctx, out, rhs, residual, search, flags, t, A0, Ai, Aj, Ak, &cg));
	struct CgGuard { mp_cg* c; ~CgGuard() { mp_cg_destroy(c); } } guard{ cg };
	mp_cg_set_accuracy(cg, prec == 4 ? (double)(float)cgAccuracy : cgAccuracy);
	MP_TRY(mp_cg_run(cg, maxIter));
	if (info) { memset(info, 0, sizeof *info); info->fixedCell = -1; mp_cg_get(cg, &info->iterations, &info->resNorm, nullptr); info->maxIter = maxIter; info->matvecKernel = ctx->lastMatvecKernel; }
	// utm1.swap(ut); ut.copyFrom(out) :145-146
	if (ut->owns && utm1->owns) { void* p = ut->d; ut->d = utm1->d; utm1->d = p; }
	else MP_CUDA(cudaMemcpyAsync(utm1->d, ut->d, ut->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	MP_CUDA(cudaMemcpyAsync(ut->d, out->d, ut->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	return MP_OK;
}

int mp_solve_pressure_host(mp_context* ctx, int prec, int sx, int sy, int sz,
                           void* vel, void* pressure, const int* flags,
                           const void* phi, const void* perCellCorr, const void* fractions, const void* obvel,
                           const void* curv, void* retRhs, const mp_pressure_params* params, mp_solve_info* info)
{
	if (!ctx || !vel || !pressure || !flags) MP_FAIL(MP_ERR_INVALID, "mp_solve_pressure_host: NULL argument");
	MP_CUDA(cudaSetDevice(ctx->device));
	GridHolder h;
	mp_grid *gFlags, *gVel, *gP, *gPhi = nullptr, *gCorr = nullptr, *gFrac = nullptr, *gOb = nullptr, *gCurv = nullptr, *gRet = nullptr;
	mp_grid like; like.sx = sx; like.sy = sy; like.sz = sz;
	MP_TRY(h.make(ctx, MP_GRID_FLAGS, 4, &like, &gFlags)); MP_TRY(h.make(ctx, MP_GRID_MAC, prec, &like, &gVel)); MP_TRY(h.make(ctx, MP_GRID_REAL, prec, &like, &gP));
	if (phi) MP_TRY(h.make(ctx, MP_GRID_REAL, prec, &like, &gPhi));
	if (perCellCorr) MP_TRY(h.make(ctx, MP_GRID_REAL, prec, &like, &gCorr));
	if (fractions) MP_TRY(h.make(ctx, MP_GRID_MAC, prec, &like, &gFrac));
	if (obvel) MP_TRY(h.make(ctx, MP_GRID_MAC, prec, &like, &gOb));
	if (curv) MP_TRY(h.make(ctx, MP_GRID_REAL, prec, &like, &gCurv));
	if (retRhs) MP_TRY(h.make(ctx, MP_GRID_REAL, prec, &like, &gRet));
	MP_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));
	MP_TRY(mp_grid_upload_async(gFlags, flags)); MP_TRY(mp_grid_upload_async(gVel, vel));
	if (phi) MP_TRY(mp_grid_upload_async(gPhi, phi));
	if (perCellCorr) MP_TRY(mp_grid_upload_async(gCorr, perCellCorr));
	if (fractions) MP_TRY(mp_grid_upload_async(gFrac, fractions));
	if (obvel) MP_TRY(mp_grid_upload_async(gOb, obvel));
	if (curv) MP_TRY(mp_grid_upload_async(gCurv, curv));
	MP_CUDA(cudaEventRecord(ctx->ev[6], ctx->stream));
	// the pressure grid is an output only: GridCg::doInit clears it (conjugategrad.cpp:214)
	MP_TRY(mp_solve_pressure(ctx, gVel, gP, gFlags, gPhi, gCorr, gFrac, gOb, gCurv, gRet, params, info));
	MP_CUDA(cudaEventRecord(ctx->ev[6 + 1], ctx->stream));
	MP_TRY(mp_grid_download_async(gVel, vel)); MP_TRY(mp_grid_download_async(gP, pressure));
	if (retRhs) MP_TRY(mp_grid_download_async(gRet, retRhs));
	MP_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	if (info) {
		cudaEventElapsedTime(&info->msH2D, ctx->ev[5], ctx->ev[6]);
		cudaEventElapsedTime(&info->msD2H, ctx->ev[7], ctx->ev[0]);
	}
	return MP_OK;
}

} // extern "C"

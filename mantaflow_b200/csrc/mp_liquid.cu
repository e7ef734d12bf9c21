// Liquid neighbours of the pressure projection on the device (first slice).  The per-cell operations and the pass
// sequences are in mp_liquid_cells.cuh (shared with the host emulation the CPU tests run); this file is the CUDA executor (launch geometry:
// liquid::threadCells, 128 threads x 4 x 4 cells per block) and the C-ABI entry points.
// Every pass is a streaming sweep: 4 B/cell of marks plus the few values that change; nothing returns to the host.
#include "mp_liquid_cells.cuh"

namespace {

template <typename F>
__global__ void __launch_bounds__(liquid::kThreads, 4) k_liquid_cells(Dims d, F f) {
	liquid::threadCells(d, f, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, (int)threadIdx.x);
}

// ---------------------------------------------------------------- frontier schedule of the extrapolation passes
// A pass only changes unmarked cells next to a cell that carries the mark of that pass, and those cells were all marked by the pass
// before.  So instead of sweeping the grid `distance` times (4 B/cell of marks through the L2 per sweep, seven loads per cell), the first
// sweep also lists the cells it marks, and every further pass walks the six neighbours of the listed cells only -- the SAME per-cell
// code (F::load / F::apply of mp_liquid_cells.cuh) on the few cells it can change, in any order (see the note on execution order there).
// A candidate with several listed neighbours is taken by the one that comes first in the neighbour order +x -x +y -y +z -z; the
// others leave it alone, so a cell is processed and listed once per pass and the lists hold at most one entry per cell.
struct ListSink {
	int* list; int* count;
	// the lanes of a warp that list a cell in the same instruction share ONE atomic on the counter (a million single-address atomics made the
	// listing pass of a 512^3 extrapolation 0.8 - 1.7 ms); the order of a list's entries is free (see above)
	__device__ __forceinline__ void operator()(IndexInt idx) const {
		const unsigned m = __activemask();
		const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
		int base = 0;
		if (lane == leader) base = atomicAdd(count, __popc(m));
		base = __shfl_sync(m, base, leader);
		list[base + __popc(m & ((1u << lane) - 1u))] = (int)idx;
	}
};
template <typename F>
__global__ void __launch_bounds__(liquid::kThreads, 4) k_liquid_cells_list(Dims d, F f, ListSink sink) {
	liquid::threadCells(d, f, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, (int)threadIdx.x, sink);
}
template <typename F>
__global__ void __launch_bounds__(256) k_liquid_frontier(Dims d, F f, const int* __restrict__ list, const int* __restrict__ countIn, ListSink out) {
	const int nq = d.is3D ? 6 : 4;
	const long long work = (long long)(*countIn) * nq;
	for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < work; t += (long long)gridDim.x * blockDim.x) {
		const int e = (int)(t / nq), q = (int)(t - (long long)e * nq);
		const IndexInt nb = (IndexInt)list[e] + liquid::nbOffset(d, q);
		int i, j, k; cellOf(d, nb, i, j, k);
		const typename F::State s = f.load(d, i, j, k, nb);
		if (!s.interior) continue;
		// from the candidate's side the listed cell is neighbour q ^ 1: take it only if no earlier neighbour is listed too
		const int mine = q ^ 1;
		bool owner = f.reached(s.tn[mine]);
		for (int qq = 0; qq < mine; qq++) owner = owner && !f.reached(s.tn[qq]);
		if (!owner) continue;
		if (f.apply(d, i, j, k, nb, s)) out(nb);
	}
}

struct CudaExec {
	mp_context* ctx;
	template <typename F> int cells(const Dims& d, const F& f) {
		const liquid::LaunchGeom g = liquid::launchGeomOf(d);
		k_liquid_cells<F><<<dim3(g.gx, g.gy, g.gz), liquid::kThreads, 0, ctx->stream>>>(d, f);
		MP_CHECK_LAUNCH(ctx);
		return MP_OK;
	}
};

struct Tmp {      // scratch grid of the context's pool, released on scope exit
	mp_grid* g = nullptr;
	~Tmp() { if (g) mp_grid_destroy(g); }
};

int checkLiquid(const char* who, mp_context* ctx, const mp_grid* g) {
	if (!ctx || !g) MP_FAIL(MP_ERR_INVALID, "%s: NULL argument", who);
	if (ctx->dist && ctx->dist->active) MP_FAIL(MP_ERR_UNSUPPORTED, "%s: not available on z-slab sharded grids yet", who);
	if (g->sy > 65535 || g->sz > 65535) MP_FAIL(MP_ERR_UNSUPPORTED, "%s: grids beyond 65535 cells in y or z are not supported", who);
	MP_CUDA(cudaSetDevice(ctx->device));
	return MP_OK;
}
bool noInterior(const mp_grid* g) { return g->sx < 3 || g->sy < 3 || (g->sz > 1 && g->sz < 3); }

static bool useFrontier(const mp_grid* g) {
	const char* e = getenv("MP_LIQUID_FRONTIER");      // read per call: the tests run both schedules in one process
	return (!e || atoi(e)) && g->n < ((IndexInt)1 << 31);
}
// two cell lists (one int per cell each, ping-pong) and their counters (ctx->dScal + 32 .. 35 as ints)
struct Frontier {
	Tmp a, b; int* cnt;
	int init(mp_context* ctx, const mp_grid* like) {
		MP_TRY(mp_grid_create_scratch(ctx, MP_GRID_FLAGS, 4, like->sx, like->sy, like->sz, &a.g));
		MP_TRY(mp_grid_create_scratch(ctx, MP_GRID_FLAGS, 4, like->sx, like->sy, like->sz, &b.g));
		cnt = (int*)(ctx->dScal + 32);
		MP_CUDA(cudaMemsetAsync(cnt, 0, 4 * sizeof(int), ctx->stream));
		return MP_OK;
	}
	int* list(int p) const { return (int*)((p & 1) ? b.g->d : a.g->d); }
};
template <typename F> static int cellsListed(mp_context* ctx, const Dims& d, const F& f, int* list, int* count) {
	const liquid::LaunchGeom g = liquid::launchGeomOf(d);
	ListSink sink = { list, count };
	k_liquid_cells_list<F><<<dim3(g.gx, g.gy, g.gz), liquid::kThreads, 0, ctx->stream>>>(d, f, sink);
	MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}
template <typename F> static int frontierPass(mp_context* ctx, const Dims& d, const F& f, const int* listIn, const int* countIn, int* listOut, int* countOut) {
	MP_CUDA(cudaMemsetAsync(countOut, 0, sizeof(int), ctx->stream));
	ListSink sink = { listOut, countOut };
	k_liquid_frontier<F><<<ctx->smCount * 4, 256, 0, ctx->stream>>>(d, f, listIn, countIn, sink);
	MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

template <typename Real>
int macSimple(mp_context* ctx, const mp_grid* flags, mp_grid* vel, int distance, const mp_grid* phiObs, int intoObs) {
	Tmp tmp, stage;
	MP_TRY(mp_grid_create_scratch(ctx, MP_GRID_FLAGS, vel->prec, vel->sx, vel->sy, vel->sz, &tmp.g));      // every cell written by MacMark
	MP_TRY(mp_grid_create_scratch(ctx, MP_GRID_MAC, vel->prec, vel->sx, vel->sy, vel->sz, &stage.g));       // only its outer layer is written and read
	CudaExec ex = { ctx };
	if (useFrontier(vel) && distance >= 2) {
		// extrapolateMacSimple (mp_liquid_cells.cuh) with passes 2 .. distance on the cell lists
		const Dims d = dimsOf(flags);
		Frontier fr; MP_TRY(fr.init(ctx, vel));
		Real* v = (Real*)vel->d; int* t = (int*)tmp.g->d;
		{ liquid::MacMark<Real> op = { (const int*)flags->d, t, intoObs ? 1 : 0 }; MP_TRY(ex.cells(d, op)); }
		{ liquid::MacExtrapolate<Real> op = { v, t, 1 }; MP_TRY(cellsListed(ctx, d, op, fr.list(2), fr.cnt + 0)); }      // pass 1 lists the cells it marks (mark 2)
		for (int pass = 2; pass < 1 + distance; pass++) {
			liquid::MacExtrapolate<Real> op = { v, t, pass };
			MP_TRY(frontierPass(ctx, d, op, fr.list(pass), fr.cnt + (pass & 1), fr.list(pass + 1), fr.cnt + ((pass + 1) & 1)));
		}
		if (phiObs) { liquid::UnprojectNormal<Real> op = { v, (const Real*)phiObs->d, (Real)distance }; MP_TRY(ex.cells(d, op)); }
		{ liquid::IntoBndStage<Real> op = { (const int*)flags->d, v, (Real*)stage.g->d }; MP_TRY(ex.cells(d, op)); }
		{ liquid::IntoBndCopy<Real> op = { (const Real*)stage.g->d, v }; MP_TRY(ex.cells(d, op)); }
		return MP_OK;
	}
	return liquid::extrapolateMacSimple<Real>(ex, dimsOf(flags), (const int*)flags->d, (Real*)vel->d, distance, phiObs ? (const Real*)phiObs->d : nullptr,
	                                          intoObs != 0, (int*)tmp.g->d, (Real*)stage.g->d);
}

template <typename Real>
int lsSimple(mp_context* ctx, mp_grid* val, const mp_grid* phi, int distance, int inside, bool vec3) {
	Tmp tmp;
	MP_TRY(mp_grid_create_scratch(ctx, MP_GRID_FLAGS, val->prec, val->sx, val->sy, val->sz, &tmp.g));
	CudaExec ex = { ctx };
	const Dims d = dimsOf(val);
	if (useFrontier(val) && distance >= 3) {
		// extrapolateLs (mp_liquid_cells.cuh): the mark pass lists the first layer (mark 2), passes 2 .. distance run on the lists, the cells
		// no pass reached get knSetRemaining's value in one sweep at the end
		const Real direction = vec3 ? (Real)0 : (inside ? (Real)-1. : (Real)1.);
		const Real remaining = vec3 ? (Real)0 : (Real)(direction * (distance + 2));
		Frontier fr; MP_TRY(fr.init(ctx, val));
		int* t = (int*)tmp.g->d;
		{ liquid::LsMark<Real> op = { (const Real*)phi->d, t, inside ? 1 : 0 }; MP_TRY(cellsListed(ctx, d, op, fr.list(2), fr.cnt + 0)); }
		for (int pass = 2; pass < 1 + distance; pass++) {
			int* lin = fr.list(pass); int* cin = fr.cnt + (pass & 1); int* lout = fr.list(pass + 1); int* cout = fr.cnt + ((pass + 1) & 1);
			if (vec3) { liquid::LsExtrapolate<Real, 3> op = { (Real*)val->d, t, pass, direction, 0, remaining }; MP_TRY(frontierPass(ctx, d, op, lin, cin, lout, cout)); }
			else      { liquid::LsExtrapolate<Real, 1> op = { (Real*)val->d, t, pass, direction, 0, remaining }; MP_TRY(frontierPass(ctx, d, op, lin, cin, lout, cout)); }
		}
		if (vec3) { liquid::LsRemaining<Real, 3> op = { (Real*)val->d, t, remaining }; return ex.cells(d, op); }
		liquid::LsRemaining<Real, 1> op = { (Real*)val->d, t, remaining }; return ex.cells(d, op);
	}
	if (vec3) return liquid::extrapolateLs<Real, 3>(ex, d, (Real*)val->d, (const Real*)phi->d, distance, inside != 0, (Real)0, (Real)0, (int*)tmp.g->d);
	const Real direction = inside ? (Real)-1. : (Real)1.;
	return liquid::extrapolateLs<Real, 1>(ex, d, (Real*)val->d, (const Real*)phi->d, distance, inside != 0, direction, (Real)(direction * (distance + 2)), (int*)tmp.g->d);
}

}  // namespace

// setWallBcs(fractions, phiObs): called by mp_set_wall_bcs (mp_step.cu); the result grid is swapped in like MACGrid::swap (extforces.cpp:313)
int mp_set_wall_bcs_frac_impl(mp_context* ctx, const mp_grid* flags, mp_grid* vel, const mp_grid* phiObs)
{
	MP_TRY(checkLiquid("mp_set_wall_bcs", ctx, flags));
	MP_TRY(mp_check_same(flags, phiObs, MP_GRID_REAL, "phiObs", false));
	if (phiObs->prec != vel->prec) MP_FAIL(MP_ERR_INVALID, "setWallBcs: phiObs and vel differ in precision");
	Tmp tgt; MP_TRY(mp_grid_create_scratch(ctx, MP_GRID_MAC, vel->prec, vel->sx, vel->sy, vel->sz, &tgt.g));      // every cell is written
	CudaExec ex = { ctx };
	const Dims d = dimsOf(flags);
	if (vel->prec == 4) { liquid::WallBcsFrac<float> op = { (const int*)flags->d, (const float*)vel->d, (float*)tgt.g->d, (const float*)phiObs->d }; MP_TRY(ex.cells(d, op)); }
	else { liquid::WallBcsFrac<double> op = { (const int*)flags->d, (const double*)vel->d, (double*)tgt.g->d, (const double*)phiObs->d }; MP_TRY(ex.cells(d, op)); }
	if (vel->owns && tgt.g->owns) { void* q = vel->d; vel->d = tgt.g->d; tgt.g->d = q; return MP_OK; }
	MP_CUDA(cudaMemcpyAsync(vel->d, tgt.g->d, vel->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	return MP_OK;
}

extern "C" {

int mp_extrapolate_mac_simple(mp_context* ctx, const mp_grid* flags, mp_grid* vel, int distance, const mp_grid* phiObs, int intoObs)
{
	MP_TRY(checkLiquid("mp_extrapolate_mac_simple", ctx, flags));
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "extrapolateMACSimple: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, vel, MP_GRID_MAC, "vel", false));
	if (phiObs) MP_TRY(mp_check_same(flags, phiObs, MP_GRID_REAL, "phiObs", false));
	if (phiObs && phiObs->prec != vel->prec) MP_FAIL(MP_ERR_INVALID, "extrapolateMACSimple: phiObs and vel differ in precision");
	if (distance > 250) MP_FAIL(MP_ERR_UNSUPPORTED, "extrapolateMACSimple: distance %d > 250 is not supported (marks are bytes)", distance);
	if (noInterior(flags)) MP_FAIL(MP_ERR_INVALID, "extrapolateMACSimple: grid without interior cells");
	if (vel->prec == 4) return macSimple<float>(ctx, flags, vel, distance, phiObs, intoObs);
	return macSimple<double>(ctx, flags, vel, distance, phiObs, intoObs);
}

int mp_extrapolate_mac_from_weight(mp_context* ctx, mp_grid* vel, mp_grid* weight, int distance)
{
	MP_TRY(checkLiquid("mp_extrapolate_mac_from_weight", ctx, vel));
	if (vel->kind != MP_GRID_MAC) MP_FAIL(MP_ERR_INVALID, "extrapolateMACFromWeight: vel is not a MAC grid");
	MP_TRY(mp_check_same(vel, weight, MP_GRID_MAC, "weight", false));
	if (vel == weight || vel->d == weight->d) MP_FAIL(MP_ERR_INVALID, "extrapolateMACFromWeight: vel and weight must be different grids");
	if (noInterior(vel)) MP_FAIL(MP_ERR_INVALID, "extrapolateMACFromWeight: grid without interior cells");
	CudaExec ex = { ctx };
	if (vel->prec == 4) return liquid::extrapolateMacFromWeight<float>(ex, dimsOf(vel), (float*)vel->d, (float*)weight->d, distance);
	return liquid::extrapolateMacFromWeight<double>(ex, dimsOf(vel), (double*)vel->d, (double*)weight->d, distance);
}

int mp_extrapolate_ls_simple(mp_context* ctx, mp_grid* phi, int distance, int inside)
{
	MP_TRY(checkLiquid("mp_extrapolate_ls_simple", ctx, phi));
	if (phi->kind != MP_GRID_REAL) MP_FAIL(MP_ERR_INVALID, "extrapolateLsSimple: phi is not a real grid");
	if (noInterior(phi)) MP_FAIL(MP_ERR_INVALID, "extrapolateLsSimple: grid without interior cells");
	if (phi->prec == 4) return lsSimple<float>(ctx, phi, phi, distance, inside, false);
	return lsSimple<double>(ctx, phi, phi, distance, inside, false);
}

int mp_extrapolate_vec3_simple(mp_context* ctx, mp_grid* vel, const mp_grid* phi, int distance, int inside)
{
	MP_TRY(checkLiquid("mp_extrapolate_vec3_simple", ctx, vel));
	if (!phi) MP_FAIL(MP_ERR_INVALID, "mp_extrapolate_vec3_simple: NULL phi");
	if (vel->kind != MP_GRID_MAC) MP_FAIL(MP_ERR_INVALID, "extrapolateVec3Simple: vel is not a Vec3 grid");
	if (phi->kind != MP_GRID_REAL || phi->sx != vel->sx || phi->sy != vel->sy || phi->sz != vel->sz || phi->prec != vel->prec)
		MP_FAIL(MP_ERR_INVALID, "extrapolateVec3Simple: phi does not match vel (kind, size or precision)");
	if (noInterior(vel)) MP_FAIL(MP_ERR_INVALID, "extrapolateVec3Simple: grid without interior cells");
	if (vel->prec == 4) return lsSimple<float>(ctx, vel, phi, distance, inside, true);
	return lsSimple<double>(ctx, vel, phi, distance, inside, true);
}

int mp_flags_update_from_levelset(mp_context* ctx, mp_grid* flags, const mp_grid* levelset)
{
	MP_TRY(checkLiquid("mp_flags_update_from_levelset", ctx, flags));
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "updateFromLevelset: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, levelset, MP_GRID_REAL, "levelset", false));
	CudaExec ex = { ctx };
	const Dims d = dimsOf(flags);
	if (levelset->prec == 4) { liquid::UpdateFromLevelset<float> op = { (int*)flags->d, (const float*)levelset->d }; return ex.cells(d, op); }
	liquid::UpdateFromLevelset<double> op = { (int*)flags->d, (const double*)levelset->d };
	return ex.cells(d, op);
}

int mp_update_fractions(mp_context* ctx, const mp_grid* flags, const mp_grid* phiObs, mp_grid* fractions, int boundaryWidth, double fracThreshold)
{
	MP_TRY(checkLiquid("mp_update_fractions", ctx, flags));
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "updateFractions: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, phiObs, MP_GRID_REAL, "phiObs", false));
	MP_TRY(mp_check_same(flags, fractions, MP_GRID_MAC, "fractions", false));
	if (phiObs->prec != fractions->prec) MP_FAIL(MP_ERR_INVALID, "updateFractions: phiObs and fractions differ in precision");
	CudaExec ex = { ctx };
	const Dims d = dimsOf(flags);
	if (fractions->prec == 4) { liquid::UpdateFractions<float> op = { (const int*)flags->d, (const float*)phiObs->d, (float*)fractions->d, boundaryWidth, (float)fracThreshold }; return ex.cells(d, op); }
	liquid::UpdateFractions<double> op = { (const int*)flags->d, (const double*)phiObs->d, (double*)fractions->d, boundaryWidth, fracThreshold };
	return ex.cells(d, op);
}

int mp_set_obstacle_flags(mp_context* ctx, mp_grid* flags, const mp_grid* phiObs, const mp_grid* fractions, const mp_grid* phiOut, const mp_grid* phiIn, int boundaryWidth)
{
	MP_TRY(checkLiquid("mp_set_obstacle_flags", ctx, flags));
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "setObstacleFlags: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, phiObs, MP_GRID_REAL, "phiObs", false));
	MP_TRY(mp_check_same(flags, fractions, MP_GRID_MAC, "fractions", true));
	MP_TRY(mp_check_same(flags, phiOut, MP_GRID_REAL, "phiOut", true));
	MP_TRY(mp_check_same(flags, phiIn, MP_GRID_REAL, "phiIn", true));
	const int prec = phiObs->prec;
	if ((fractions && fractions->prec != prec) || (phiOut && phiOut->prec != prec) || (phiIn && phiIn->prec != prec)) MP_FAIL(MP_ERR_INVALID, "setObstacleFlags: grids differ in precision");
	if (boundaryWidth < 1) MP_FAIL(MP_ERR_INVALID, "setObstacleFlags: boundaryWidth must be >= 1 (the kernel reads the +x / +y / +z neighbour of every cell it visits)");
	CudaExec ex = { ctx };
	const Dims d = dimsOf(flags);
	if (prec == 4) { liquid::SetObstacleFlags<float> op = { (int*)flags->d, (const float*)phiObs->d, fractions ? (const float*)fractions->d : nullptr, phiOut ? (const float*)phiOut->d : nullptr,
	                                                        phiIn ? (const float*)phiIn->d : nullptr, boundaryWidth }; return ex.cells(d, op); }
	liquid::SetObstacleFlags<double> op = { (int*)flags->d, (const double*)phiObs->d, fractions ? (const double*)fractions->d : nullptr, phiOut ? (const double*)phiOut->d : nullptr,
	                                        phiIn ? (const double*)phiIn->d : nullptr, boundaryWidth };
	return ex.cells(d, op);
}

static int stencilOp(const char* who, mp_context* ctx, mp_grid* out, const mp_grid* grid, double h, bool curvature)
{
	MP_TRY(checkLiquid(who, ctx, out));
	if (out->kind != MP_GRID_REAL) MP_FAIL(MP_ERR_INVALID, "%s: the result is not a real grid", who);
	MP_TRY(mp_check_same(out, grid, MP_GRID_REAL, "grid", false));
	if (out == grid || out->d == grid->d) MP_FAIL(MP_ERR_INVALID, "%s: result and input must be different grids", who);
	if (noInterior(out)) return MP_OK;
	CudaExec ex = { ctx };
	const Dims d = dimsOf(out);
	if (out->prec == 4) {
		if (curvature) { liquid::CurvatureCell<float> op = { (float*)out->d, (const float*)grid->d, (float)h }; return ex.cells(d, op); }
		liquid::LaplaceCell<float> op = { (float*)out->d, (const float*)grid->d }; return ex.cells(d, op);
	}
	if (curvature) { liquid::CurvatureCell<double> op = { (double*)out->d, (const double*)grid->d, h }; return ex.cells(d, op); }
	liquid::LaplaceCell<double> op = { (double*)out->d, (const double*)grid->d }; return ex.cells(d, op);
}
int mp_get_laplacian(mp_context* ctx, mp_grid* laplacian, const mp_grid* grid) { return stencilOp("mp_get_laplacian", ctx, laplacian, grid, 1.0, false); }
int mp_get_curvature(mp_context* ctx, mp_grid* curv, const mp_grid* grid, double h) { return stencilOp("mp_get_curvature", ctx, curv, grid, h, true); }

int mp_grid_set_bound(mp_context* ctx, mp_grid* g, double vx, double vy, double vz, int boundaryWidth)
{
	MP_TRY(checkLiquid("mp_grid_set_bound", ctx, g));
	CudaExec ex = { ctx };
	const Dims d = dimsOf(g);
	if (g->kind == MP_GRID_FLAGS) { liquid::SetBound<int, 1> op = { (int*)g->d, { (int)vx }, boundaryWidth }; return ex.cells(d, op); }
	if (g->kind == MP_GRID_REAL) {
		if (g->prec == 4) { liquid::SetBound<float, 1> op = { (float*)g->d, { (float)vx }, boundaryWidth }; return ex.cells(d, op); }
		liquid::SetBound<double, 1> op = { (double*)g->d, { vx }, boundaryWidth }; return ex.cells(d, op);
	}
	if (g->prec == 4) { liquid::SetBound<float, 3> op = { (float*)g->d, { (float)vx, (float)vy, (float)vz }, boundaryWidth }; return ex.cells(d, op); }
	liquid::SetBound<double, 3> op = { (double*)g->d, { vx, vy, vz }, boundaryWidth };
	return ex.cells(d, op);
}

}

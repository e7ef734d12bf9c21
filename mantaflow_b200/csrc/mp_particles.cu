// FLIP particle <-> grid plugins on the device (second slice).  The per-cell / per-particle operations and the pass
// sequences are in mp_particles_cells.cuh (shared with the host emulation the CPU tests run); this file is the CUDA executor and the
// C-ABI entry points.  Particle arrays are mp_grids of size (N, 1, 1): Vec3 data as MP_GRID_MAC, int data as MP_GRID_FLAGS.
// Scan and stable radix sort come from cub (bucketing particles by cell); everything else is hand-written.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "mp_particles_cells.cuh"

namespace {

template <typename F>
__global__ void __launch_bounds__(liquid::kThreads, 4) k_parts_cells(Dims d, F f) {
	liquid::threadCells(d, f, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, (int)threadIdx.x);
}
// one particle per thread, grid-stride (the grid is capped at a few waves of the SMs)
template <typename F>
__global__ void __launch_bounds__(256) k_parts(IndexInt np, F f) {
	for (IndexInt idx = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; idx < np; idx += (IndexInt)gridDim.x * blockDim.x) f(idx);
}

struct Tmp {      // scratch array of the context's pool, released on scope exit
	mp_grid* g = nullptr;
	~Tmp() { if (g) mp_grid_destroy(g); }
	int* ints() const { return (int*)g->d; }
};
int scratchInts(mp_context* ctx, IndexInt n, Tmp& t) {
	if (n < 1) n = 1;
	if (n > 0x7fffffffLL) MP_FAIL(MP_ERR_UNSUPPORTED, "particle plugins: more than 2^31 - 1 entries");
	return mp_grid_create_scratch(ctx, MP_GRID_FLAGS, 4, (int)n, 1, 1, &t.g);
}

struct CudaExec {
	mp_context* ctx;
	template <typename F> int cells(const Dims& d, const F& f) {
		const liquid::LaunchGeom g = liquid::launchGeomOf(d);
		k_parts_cells<F><<<dim3(g.gx, g.gy, g.gz), liquid::kThreads, 0, ctx->stream>>>(d, f);
		MP_CHECK_LAUNCH(ctx);
		return MP_OK;
	}
	template <typename F> int parts(IndexInt np, const F& f) {
		if (np <= 0) return MP_OK;
		IndexInt blocks = (np + 255) / 256;
		const IndexInt cap = (IndexInt)ctx->smCount * 32;
		if (blocks > cap) blocks = cap;
		k_parts<F><<<(unsigned)blocks, 256, 0, ctx->stream>>>(np, f);
		MP_CHECK_LAUNCH(ctx);
		return MP_OK;
	}
	int zero(void* p, size_t bytes) { MP_CUDA(cudaMemsetAsync(p, 0, bytes, ctx->stream)); return MP_OK; }
	// in place; the total (last exclusive value + last count) goes to the host: one synchronisation per bucketing
	int exclusiveScan(int* data, IndexInt n, IndexInt* total) {
		int* h = (int*)ctx->hScal;
		MP_CUDA(cudaMemcpyAsync(h, data + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
		size_t bytes = 0;
		MP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, data, data, (int)n, ctx->stream));
		Tmp tmp; MP_TRY(scratchInts(ctx, (IndexInt)((bytes + 3) / 4), tmp));
		MP_CUDA(cub::DeviceScan::ExclusiveSum(tmp.g->d, bytes, data, data, (int)n, ctx->stream));
		ctx->launches++;
		MP_CUDA(cudaMemcpyAsync(h + 1, data + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
		MP_CUDA(cudaStreamSynchronize(ctx->stream));
		*total = (IndexInt)h[0] + (IndexInt)h[1];
		return MP_OK;
	}
	int sortPairs(int* keys, int* keysTmp, int* vals, int* valsOut, IndexInt n, int keyBits) {
		size_t bytes = 0;
		MP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, keysTmp, vals, valsOut, (int)n, 0, keyBits, ctx->stream));
		Tmp tmp; MP_TRY(scratchInts(ctx, (IndexInt)((bytes + 3) / 4), tmp));
		MP_CUDA(cub::DeviceRadixSort::SortPairs(tmp.g->d, bytes, keys, keysTmp, vals, valsOut, (int)n, 0, keyBits, ctx->stream));
		ctx->launches++;
		return MP_OK;
	}
};

// ---- argument checks shared by the entry points
int checkCtx(const char* who, mp_context* ctx, const mp_grid* g) {
	if (!ctx || !g) MP_FAIL(MP_ERR_INVALID, "%s: NULL argument", who);
	if (ctx->dist && ctx->dist->active) MP_FAIL(MP_ERR_UNSUPPORTED, "%s: not available on z-slab sharded grids yet", who);
	if (g->sy > 65535 || g->sz > 65535) MP_FAIL(MP_ERR_UNSUPPORTED, "%s: grids beyond 65535 cells in y or z are not supported", who);
	if (g->n >= 0x7fffffffLL) MP_FAIL(MP_ERR_UNSUPPORTED, "%s: grids of 2^31 - 1 cells or more are not supported (cell keys are ints)", who);
	MP_CUDA(cudaSetDevice(ctx->device));
	return MP_OK;
}
// a per-particle array with room for np entries of the given kind (and precision, for Vec3 data)
int checkArray(const char* who, const char* name, const mp_grid* a, int kind, int prec, long long np, bool optional) {
	if (!a) { if (optional || np == 0) return MP_OK; MP_FAIL(MP_ERR_INVALID, "%s: NULL %s", who, name); }
	if (a->kind != kind) MP_FAIL(MP_ERR_INVALID, "%s: %s has the wrong element type", who, name);
	if (kind != MP_GRID_FLAGS && a->prec != prec) MP_FAIL(MP_ERR_INVALID, "%s: %s differs in precision", who, name);
	if (a->n < np) MP_FAIL(MP_ERR_INVALID, "%s: %s holds %lld entries, %lld particles given", who, name, a->n, np);
	return MP_OK;
}
int checkParts(const char* who, long long np, const mp_grid* pos, const mp_grid* pflag, const mp_grid* ptype, int prec) {
	if (np < 0 || np > 0x7fffffffLL) MP_FAIL(MP_ERR_INVALID, "%s: bad particle count %lld", who, np);
	MP_TRY(checkArray(who, "pos", pos, MP_GRID_MAC, prec, np, false));
	MP_TRY(checkArray(who, "pflag", pflag, MP_GRID_FLAGS, prec, np, false));
	MP_TRY(checkArray(who, "ptype", ptype, MP_GRID_FLAGS, prec, np, true));
	return MP_OK;
}
template <typename Real> parts::PSet<Real> psetOf(const mp_grid* pos, const mp_grid* pflag, const mp_grid* ptype, int exclude) {
	parts::PSet<Real> ps = { pos ? (const Real*)pos->d : nullptr, pflag ? (const int*)pflag->d : nullptr, ptype ? (const int*)ptype->d : nullptr, exclude };
	return ps;
}

template <typename Real>
int markFluid(mp_context* ctx, long long np, const mp_grid* pos, const mp_grid* pflag, mp_grid* flags, const mp_grid* phiObs, const mp_grid* ptype, int exclude) {
	Tmp tmp;
	if (phiObs) MP_TRY(mp_grid_create_scratch(ctx, MP_GRID_FLAGS, 4, flags->sx, flags->sy, flags->sz, &tmp.g));      // every cell written by SetNbObstacle
	CudaExec ex = { ctx };
	bool swapped = false;
	MP_TRY(parts::markFluidCells<Real>(ex, dimsOf(flags), (int*)flags->d, np, psetOf<Real>(pos, pflag, ptype, exclude), phiObs ? (const Real*)phiObs->d : nullptr,
	                                   tmp.g ? tmp.ints() : nullptr, &swapped));
	if (swapped) {
		if (flags->owns && tmp.g->owns) { void* q = flags->d; flags->d = tmp.g->d; tmp.g->d = q; }
		else MP_CUDA(cudaMemcpyAsync(flags->d, tmp.g->d, flags->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	}
	return MP_OK;
}

template <typename Real>
int mapParts(mp_context* ctx, mp_grid* vel, mp_grid* velOld, long long np, const mp_grid* pos, const mp_grid* pflag, const mp_grid* partVel, mp_grid* weight,
             const mp_grid* ptype, int exclude) {
	Tmp start, key, keyTmp, val, sorted;
	MP_TRY(scratchInts(ctx, vel->n, start)); MP_TRY(scratchInts(ctx, np, key)); MP_TRY(scratchInts(ctx, np, keyTmp)); MP_TRY(scratchInts(ctx, np, val)); MP_TRY(scratchInts(ctx, np, sorted));
	CudaExec ex = { ctx };
	// the tree of 3-way merges (default; MP_MAPPARTS=0: the 27-way walk).  Its lists hold a particle 3 + 9 times: 96 + 24 bytes per particle
	const char* e = getenv("MP_MAPPARTS");
	const bool useTree = (!e || atoi(e)) && np > 0 && 2 * parts::treeEntries(vel->n, np, 9) <= 0x7fffffffLL;
	if (!useTree)
		return parts::mapPartsToMAC<Real>(ex, dimsOf(vel), (Real*)vel->d, (Real*)velOld->d, np, psetOf<Real>(pos, pflag, ptype, exclude), partVel ? (const Real*)partVel->d : nullptr,
		                                  weight ? (Real*)weight->d : nullptr, start.ints(), key.ints(), keyTmp.ints(), val.ints(), sorted.ints());
	Tmp len1, off1, len2, off2, e1, e2, posS, pvelS, rec;
	const IndexInt realInts = 3 * np * (IndexInt)(sizeof(Real) / 4);
	// per-particle records (3-D; MP_MAPPARTS=2: evaluate the weights in the walk): 208 / 400 bytes per particle
	const IndexInt recInts = np * (IndexInt)(sizeof(parts::PartRec<Real>) / 4);
	const bool useRec = !(e && atoi(e) == 2) && vel->sz > 1 && vel->sx < 65536 && vel->sy < 65536 && vel->sz < 65536 && recInts <= 0x7fffffffLL;
	if (useRec) MP_TRY(scratchInts(ctx, recInts, rec));
	MP_TRY(scratchInts(ctx, vel->n, len1)); MP_TRY(scratchInts(ctx, vel->n, off1)); MP_TRY(scratchInts(ctx, vel->n, len2)); MP_TRY(scratchInts(ctx, vel->n, off2));
	MP_TRY(scratchInts(ctx, 2 * parts::treeEntries(vel->n, np, 3), e1)); MP_TRY(scratchInts(ctx, 2 * parts::treeEntries(vel->n, np, 9), e2));
	if (!useRec) { MP_TRY(scratchInts(ctx, realInts, posS)); MP_TRY(scratchInts(ctx, realInts, pvelS)); }
	const parts::MapPartsTreeScratch<Real> tree = { len1.ints(), off1.ints(), len2.ints(), off2.ints(), (parts::Ent*)e1.g->d, (parts::Ent*)e2.g->d,
	                                                posS.g ? (Real*)posS.g->d : nullptr, pvelS.g ? (Real*)pvelS.g->d : nullptr, rec.g ? (parts::PartRec<Real>*)rec.g->d : nullptr };
	return parts::mapPartsToMAC<Real>(ex, dimsOf(vel), (Real*)vel->d, (Real*)velOld->d, np, psetOf<Real>(pos, pflag, ptype, exclude), partVel ? (const Real*)partVel->d : nullptr,
	                                  weight ? (Real*)weight->d : nullptr, start.ints(), key.ints(), keyTmp.ints(), val.ints(), sorted.ints(), &tree);
}

int flipUpdate(const char* who, mp_context* ctx, const mp_grid* vel, const mp_grid* velOld, long long np, const mp_grid* pos, const mp_grid* pflag, mp_grid* partVel,
               double flipRatio, bool pic, const mp_grid* ptype, int exclude) {
	MP_TRY(checkCtx(who, ctx, vel));
	if (vel->kind != MP_GRID_MAC) MP_FAIL(MP_ERR_INVALID, "%s: vel is not a MAC grid", who);
	if (!pic) MP_TRY(mp_check_same(vel, velOld, MP_GRID_MAC, "velOld", false));
	MP_TRY(checkParts(who, np, pos, pflag, ptype, vel->prec));
	MP_TRY(checkArray(who, "partVel", partVel, MP_GRID_MAC, vel->prec, np, false));
	if (vel->sx < 2 || vel->sy < 2 || (vel->sz > 1 && vel->sz < 2)) MP_FAIL(MP_ERR_INVALID, "%s: the interpolation needs at least two cells per axis", who);
	if (np == 0) return MP_OK;
	CudaExec ex = { ctx };
	const Dims d = dimsOf(vel);
	if (vel->prec == 4) {
		parts::FlipVelocityUpdate<float> op = { d, (const float*)vel->d, pic ? nullptr : (const float*)velOld->d, psetOf<float>(pos, pflag, ptype, exclude), (float*)partVel->d, (float)flipRatio, pic };
		return ex.parts(np, op);
	}
	parts::FlipVelocityUpdate<double> op = { d, (const double*)vel->d, pic ? nullptr : (const double*)velOld->d, psetOf<double>(pos, pflag, ptype, exclude), (double*)partVel->d, flipRatio, pic };
	return ex.parts(np, op);
}

}  // namespace

extern "C" {

int mp_mark_fluid_cells(mp_context* ctx, long long np, const mp_grid* pos, const mp_grid* pflag, mp_grid* flags, const mp_grid* phiObs, const mp_grid* ptype, int exclude)
{
	MP_TRY(checkCtx("mp_mark_fluid_cells", ctx, flags));
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "markFluidCells: flags is not a FlagGrid");
	if (phiObs) MP_TRY(mp_check_same(flags, phiObs, MP_GRID_REAL, "phiObs", false));
	const int prec = pos ? pos->prec : (phiObs ? phiObs->prec : 4);
	if (phiObs && phiObs->prec != prec) MP_FAIL(MP_ERR_INVALID, "markFluidCells: phiObs and the particle positions differ in precision");
	MP_TRY(checkParts("mp_mark_fluid_cells", np, pos, pflag, ptype, prec));
	if (prec == 4) return markFluid<float>(ctx, np, pos, pflag, flags, phiObs, ptype, exclude);
	return markFluid<double>(ctx, np, pos, pflag, flags, phiObs, ptype, exclude);
}

int mp_grid_particle_index(mp_context* ctx, long long np, const mp_grid* pos, const mp_grid* pflag, mp_grid* indexSys, const mp_grid* flags, mp_grid* index, long long* count)
{
	MP_TRY(checkCtx("mp_grid_particle_index", ctx, index));
	if (!count) MP_FAIL(MP_ERR_INVALID, "mp_grid_particle_index: NULL count");
	if (index->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "gridParticleIndex: index is not a Grid<int>");
	if (flags) MP_TRY(mp_check_same(index, flags, MP_GRID_FLAGS, "flags", false));
	const int prec = pos ? pos->prec : 4;
	MP_TRY(checkParts("mp_grid_particle_index", np, pos, pflag, nullptr, prec));
	MP_TRY(checkArray("mp_grid_particle_index", "indexSys", indexSys, MP_GRID_FLAGS, prec, np, false));
	Tmp key, keyTmp, val;
	MP_TRY(scratchInts(ctx, np, key)); MP_TRY(scratchInts(ctx, np, keyTmp)); MP_TRY(scratchInts(ctx, np, val));
	Tmp sortedTmp;      // np == 0 with no indexSys array: a one-entry stand-in that is never written
	if (!indexSys) MP_TRY(scratchInts(ctx, 1, sortedTmp));
	int* sorted = indexSys ? (int*)indexSys->d : sortedTmp.ints();
	CudaExec ex = { ctx };
	IndexInt c = 0;
	const Dims d = dimsOf(index);
	if (prec == 4) MP_TRY(parts::bucketParticles<float>(ex, d, np, psetOf<float>(pos, pflag, nullptr, 0), false, (int*)index->d, key.ints(), keyTmp.ints(), val.ints(), sorted, &c));
	else MP_TRY(parts::bucketParticles<double>(ex, d, np, psetOf<double>(pos, pflag, nullptr, 0), false, (int*)index->d, key.ints(), keyTmp.ints(), val.ints(), sorted, &c));
	*count = c;
	return MP_OK;
}

int mp_union_particle_levelset(mp_context* ctx, long long np, const mp_grid* pos, const mp_grid* indexSys, long long count, const mp_grid* flags, const mp_grid* index,
                               mp_grid* phi, double radiusFactor, const mp_grid* ptype, int exclude)
{
	MP_TRY(checkCtx("mp_union_particle_levelset", ctx, phi));
	if (!index) MP_FAIL(MP_ERR_INVALID, "mp_union_particle_levelset: NULL index");
	if (phi->kind != MP_GRID_REAL) MP_FAIL(MP_ERR_INVALID, "unionParticleLevelset: phi is not a real grid");
	MP_TRY(mp_check_same(phi, index, MP_GRID_FLAGS, "index", false));
	if (flags) MP_TRY(mp_check_same(phi, flags, MP_GRID_FLAGS, "flags", false));
	if (np < 0 || count < 0 || count > np) MP_FAIL(MP_ERR_INVALID, "unionParticleLevelset: bad counts (%lld particles, %lld indexed)", np, count);
	MP_TRY(checkArray("mp_union_particle_levelset", "pos", pos, MP_GRID_MAC, phi->prec, np, false));
	MP_TRY(checkArray("mp_union_particle_levelset", "indexSys", indexSys, MP_GRID_FLAGS, 4, count, false));
	MP_TRY(checkArray("mp_union_particle_levelset", "ptype", ptype, MP_GRID_FLAGS, 4, np, true));
	CudaExec ex = { ctx };
	const Dims d = dimsOf(phi);
	const int* isys = indexSys ? (const int*)indexSys->d : nullptr;
	const int* pt = ptype ? (const int*)ptype->d : nullptr;
	Tmp posS;      // positions in index order (MP_UNION_SORTED=0: read through indexSys)
	const char* e = getenv("MP_UNION_SORTED");
	if ((!e || atoi(e)) && count > 0) MP_TRY(scratchInts(ctx, 3 * count * (phi->prec / 4), posS));
	if (phi->prec == 4) return parts::unionParticleLevelset<float>(ex, d, pos ? (const float*)pos->d : nullptr, (const int*)index->d, isys, count, (float*)phi->d, radiusFactor, pt, exclude,
	                                                               posS.g ? (float*)posS.g->d : nullptr);
	return parts::unionParticleLevelset<double>(ex, d, pos ? (const double*)pos->d : nullptr, (const int*)index->d, isys, count, (double*)phi->d, radiusFactor, pt, exclude,
	                                            posS.g ? (double*)posS.g->d : nullptr);
}

int mp_map_parts_to_mac(mp_context* ctx, const mp_grid* flags, mp_grid* vel, mp_grid* velOld, long long np, const mp_grid* pos, const mp_grid* pflag, const mp_grid* partVel,
                        mp_grid* weight, const mp_grid* ptype, int exclude)
{
	MP_TRY(checkCtx("mp_map_parts_to_mac", ctx, vel));
	if (vel->kind != MP_GRID_MAC) MP_FAIL(MP_ERR_INVALID, "mapPartsToMAC: vel is not a MAC grid");
	MP_TRY(mp_check_same(vel, velOld, MP_GRID_MAC, "velOld", false));
	MP_TRY(mp_check_same(vel, weight, MP_GRID_MAC, "weight", true));
	if (flags && (flags->kind != MP_GRID_FLAGS || flags->sx != vel->sx || flags->sy != vel->sy || flags->sz != vel->sz)) MP_FAIL(MP_ERR_INVALID, "mapPartsToMAC: flags does not match vel");
	if (vel->d == velOld->d || (weight && (weight->d == vel->d || weight->d == velOld->d))) MP_FAIL(MP_ERR_INVALID, "mapPartsToMAC: vel, velOld and weight must be different grids");
	MP_TRY(checkParts("mp_map_parts_to_mac", np, pos, pflag, ptype, vel->prec));
	MP_TRY(checkArray("mp_map_parts_to_mac", "partVel", partVel, MP_GRID_MAC, vel->prec, np, false));
	if (vel->sx < 2 || vel->sy < 2 || (vel->sz > 1 && vel->sz < 2)) MP_FAIL(MP_ERR_INVALID, "mapPartsToMAC: the interpolation needs at least two cells per axis");
	if (vel->prec == 4) return mapParts<float>(ctx, vel, velOld, np, pos, pflag, partVel, weight, ptype, exclude);
	return mapParts<double>(ctx, vel, velOld, np, pos, pflag, partVel, weight, ptype, exclude);
}

int mp_parts_advect_in_grid(mp_context* ctx, const mp_grid* flags, const mp_grid* vel, long long np, mp_grid* pos, mp_grid* pflag, double dt, int integrationMode,
                            int deleteInObstacle, int stopInObstacle, int skipNew, const mp_grid* ptype, int exclude)
{
	MP_TRY(checkCtx("mp_parts_advect_in_grid", ctx, flags));
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "advectInGrid: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, vel, MP_GRID_MAC, "vel", false));
	if (integrationMode < 0 || integrationMode > 2) MP_FAIL(MP_ERR_INVALID, "unknown integration type");      // util/integrator.h:63
	MP_TRY(checkParts("mp_parts_advect_in_grid", np, pos, pflag, ptype, vel->prec));
	if (vel->sx < 2 || vel->sy < 2 || (vel->sz > 1 && vel->sz < 2)) MP_FAIL(MP_ERR_INVALID, "advectInGrid: the interpolation needs at least two cells per axis");
	if (np == 0) return MP_OK;
	CudaExec ex = { ctx };
	const Dims d = dimsOf(flags);
	const int* pt = ptype ? (const int*)ptype->d : nullptr;
	if (vel->prec == 4) {
		parts::AdvectInGrid<float> op = { d, (const int*)flags->d, (const float*)vel->d, (float*)pos->d, (int*)pflag->d, pt, exclude, (float)dt, integrationMode,
		                                  deleteInObstacle != 0, stopInObstacle != 0, skipNew != 0 };
		return ex.parts(np, op);
	}
	parts::AdvectInGrid<double> op = { d, (const int*)flags->d, (const double*)vel->d, (double*)pos->d, (int*)pflag->d, pt, exclude, dt, integrationMode,
	                                   deleteInObstacle != 0, stopInObstacle != 0, skipNew != 0 };
	return ex.parts(np, op);
}

int mp_push_out_of_obs(mp_context* ctx, long long np, mp_grid* pos, const mp_grid* pflag, const mp_grid* flags, const mp_grid* phiObs, double shift, double thresh,
                       const mp_grid* ptype, int exclude)
{
	MP_TRY(checkCtx("mp_push_out_of_obs", ctx, phiObs));
	if (phiObs->kind != MP_GRID_REAL) MP_FAIL(MP_ERR_INVALID, "pushOutofObs: phiObs is not a real grid");
	if (flags) MP_TRY(mp_check_same(phiObs, flags, MP_GRID_FLAGS, "flags", false));
	MP_TRY(checkParts("mp_push_out_of_obs", np, pos, pflag, ptype, phiObs->prec));
	if (phiObs->sx < 3 || phiObs->sy < 3 || (phiObs->sz > 1 && phiObs->sz < 3)) MP_FAIL(MP_ERR_INVALID, "pushOutofObs: the gradient needs at least three cells per axis");
	if (np == 0) return MP_OK;
	CudaExec ex = { ctx };
	const Dims d = dimsOf(phiObs);
	if (phiObs->prec == 4) { parts::PushOutOfObs<float> op = { d, (float*)pos->d, psetOf<float>(pos, pflag, ptype, exclude), (const float*)phiObs->d, (float)shift, (float)thresh }; return ex.parts(np, op); }
	parts::PushOutOfObs<double> op = { d, (double*)pos->d, psetOf<double>(pos, pflag, ptype, exclude), (const double*)phiObs->d, shift, thresh };
	return ex.parts(np, op);
}

int mp_parts_project_out_of_bnd(mp_context* ctx, const mp_grid* flags, long long np, mp_grid* pos, const mp_grid* pflag, double bnd, const char* plane,
                                const mp_grid* ptype, int exclude)
{
	MP_TRY(checkCtx("mp_parts_project_out_of_bnd", ctx, flags));
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "projectOutOfBnd: flags is not a FlagGrid");
	if (!plane) MP_FAIL(MP_ERR_INVALID, "mp_parts_project_out_of_bnd: NULL plane");
	if (!pos && np > 0) MP_FAIL(MP_ERR_INVALID, "mp_parts_project_out_of_bnd: NULL pos");
	MP_TRY(checkParts("mp_parts_project_out_of_bnd", np, pos, pflag, ptype, pos ? pos->prec : 4));
	if (np == 0) return MP_OK;
	int axis = 0;
	for (const char* c = plane; *c; c++) for (int q = 0; q < 6; q++) if (*c == "xXyYzZ"[q]) axis |= 1 << q;      // particle.h:580-588
	CudaExec ex = { ctx };
	const Dims d = dimsOf(flags);
	if (pos->prec == 4) { parts::ProjectOutOfBnd<float> op = { d, (float*)pos->d, psetOf<float>(pos, pflag, ptype, exclude), (float)bnd, axis }; return ex.parts(np, op); }
	parts::ProjectOutOfBnd<double> op = { d, (double*)pos->d, psetOf<double>(pos, pflag, ptype, exclude), bnd, axis };
	return ex.parts(np, op);
}

// ---- plugin/ptsplugins.cpp:17-70, ParticleSystem::getPosPdata particle.h:422-427, markIsolatedFluidCell grid.cpp:866-890
static int checkPdata(const char* who, mp_context* ctx, long long np, const mp_grid* a, const char* name, int kind, int prec) {
	if (!ctx) MP_FAIL(MP_ERR_INVALID, "%s: NULL context", who);
	if (np < 0 || np > 0x7fffffffLL) MP_FAIL(MP_ERR_INVALID, "%s: bad particle count %lld", who, np);
	MP_TRY(checkArray(who, name, a, kind, prec, np, false));
	MP_CUDA(cudaSetDevice(ctx->device));
	return MP_OK;
}
int mp_add_force_pvel(mp_context* ctx, long long np, mp_grid* vel, double ax, double ay, double az, double dt, const mp_grid* ptype, int exclude)
{
	const int prec = vel ? vel->prec : 4;
	MP_TRY(checkPdata("mp_add_force_pvel", ctx, np, vel, "vel", MP_GRID_MAC, prec)); MP_TRY(checkArray("mp_add_force_pvel", "ptype", ptype, MP_GRID_FLAGS, 4, np, true));
	if (np == 0) return MP_OK;
	CudaExec ex = { ctx };
	const int* pt = ptype ? (const int*)ptype->d : nullptr;
	if (prec == 4) { const float d = (float)dt; parts::AddForcePvel<float> op = { (float*)vel->d, { (float)ax * d, (float)ay * d, (float)az * d }, pt, exclude }; return ex.parts(np, op); }
	parts::AddForcePvel<double> op = { (double*)vel->d, { ax * dt, ay * dt, az * dt }, pt, exclude };
	return ex.parts(np, op);
}
int mp_update_velocity_from_delta_pos(mp_context* ctx, long long np, const mp_grid* pos, mp_grid* vel, const mp_grid* xPrev, double dt, const mp_grid* ptype, int exclude)
{
	const int prec = vel ? vel->prec : 4;
	MP_TRY(checkPdata("mp_update_velocity_from_delta_pos", ctx, np, vel, "vel", MP_GRID_MAC, prec));
	MP_TRY(checkArray("mp_update_velocity_from_delta_pos", "pos", pos, MP_GRID_MAC, prec, np, false)); MP_TRY(checkArray("mp_update_velocity_from_delta_pos", "x_prev", xPrev, MP_GRID_MAC, prec, np, false));
	MP_TRY(checkArray("mp_update_velocity_from_delta_pos", "ptype", ptype, MP_GRID_FLAGS, 4, np, true));
	if (np == 0) return MP_OK;
	CudaExec ex = { ctx };
	const int* pt = ptype ? (const int*)ptype->d : nullptr;
	if (prec == 4) { parts::UpdateVelocityFromDeltaPos<float> op = { (const float*)pos->d, (float*)vel->d, (const float*)xPrev->d, (float)(1.0 / (double)(float)dt), pt, exclude }; return ex.parts(np, op); }
	parts::UpdateVelocityFromDeltaPos<double> op = { (const double*)pos->d, (double*)vel->d, (const double*)xPrev->d, 1.0 / dt, pt, exclude };
	return ex.parts(np, op);
}
int mp_euler_step(mp_context* ctx, long long np, mp_grid* pos, const mp_grid* vel, double dt, const mp_grid* ptype, int exclude)
{
	const int prec = pos ? pos->prec : 4;
	MP_TRY(checkPdata("mp_euler_step", ctx, np, pos, "pos", MP_GRID_MAC, prec)); MP_TRY(checkArray("mp_euler_step", "vel", vel, MP_GRID_MAC, prec, np, false));
	MP_TRY(checkArray("mp_euler_step", "ptype", ptype, MP_GRID_FLAGS, 4, np, true));
	if (np == 0) return MP_OK;
	CudaExec ex = { ctx };
	const int* pt = ptype ? (const int*)ptype->d : nullptr;
	if (prec == 4) { parts::StepEuler<float> op = { (float*)pos->d, (const float*)vel->d, (float)dt, pt, exclude }; return ex.parts(np, op); }
	parts::StepEuler<double> op = { (double*)pos->d, (const double*)vel->d, dt, pt, exclude };
	return ex.parts(np, op);
}
int mp_set_part_type(mp_context* ctx, long long np, const mp_grid* pos, mp_grid* ptype, int mark, int stype, const mp_grid* flags, int cflag)
{
	MP_TRY(checkCtx("mp_set_part_type", ctx, flags));
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "setPartType: flags is not a FlagGrid");
	const int prec = pos ? pos->prec : 4;
	MP_TRY(checkPdata("mp_set_part_type", ctx, np, pos, "pos", MP_GRID_MAC, prec)); MP_TRY(checkArray("mp_set_part_type", "ptype", ptype, MP_GRID_FLAGS, 4, np, false));
	if (np == 0) return MP_OK;
	CudaExec ex = { ctx };
	const Dims d = dimsOf(flags);
	if (prec == 4) { parts::SetPartType<float> op = { d, (const float*)pos->d, (int*)ptype->d, mark, stype, (const int*)flags->d, cflag }; return ex.parts(np, op); }
	parts::SetPartType<double> op = { d, (const double*)pos->d, (int*)ptype->d, mark, stype, (const int*)flags->d, cflag };
	return ex.parts(np, op);
}
int mp_parts_get_pos_pdata(mp_context* ctx, long long np, const mp_grid* pos, mp_grid* target)
{
	const int prec = pos ? pos->prec : 4;
	MP_TRY(checkPdata("mp_parts_get_pos_pdata", ctx, np, pos, "pos", MP_GRID_MAC, prec)); MP_TRY(checkArray("mp_parts_get_pos_pdata", "target", target, MP_GRID_MAC, prec, np, false));
	if (np == 0) return MP_OK;
	MP_CUDA(cudaMemcpyAsync(target->d, pos->d, (size_t)np * 3 * prec, cudaMemcpyDeviceToDevice, ctx->stream));
	return MP_OK;
}
int mp_mark_isolated_fluid_cell(mp_context* ctx, mp_grid* flags, int mark)
{
	MP_TRY(checkCtx("mp_mark_isolated_fluid_cell", ctx, flags));
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "markIsolatedFluidCell: flags is not a FlagGrid");
	CudaExec ex = { ctx };
	parts::MarkIsolatedFluidCell op = { (int*)flags->d, mark };
	return ex.cells(dimsOf(flags), op);
}

int mp_map_mac_to_parts(mp_context* ctx, const mp_grid* flags, const mp_grid* vel, long long np, const mp_grid* pos, const mp_grid* pflag, mp_grid* partVel,
                        const mp_grid* ptype, int exclude)
{
	(void)flags;
	return flipUpdate("mp_map_mac_to_parts", ctx, vel, nullptr, np, pos, pflag, partVel, -1.0, true, ptype, exclude);
}

int mp_flip_velocity_update(mp_context* ctx, const mp_grid* flags, const mp_grid* vel, const mp_grid* velOld, long long np, const mp_grid* pos, const mp_grid* pflag,
                            mp_grid* partVel, double flipRatio, const mp_grid* ptype, int exclude)
{
	(void)flags;
	if (!velOld) MP_FAIL(MP_ERR_INVALID, "mp_flip_velocity_update: NULL velOld");
	return flipUpdate("mp_flip_velocity_update", ctx, vel, velOld, np, pos, pflag, partVel, flipRatio, false, ptype, exclude);
}

}

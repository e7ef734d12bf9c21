// Multi-GPU: z-slab sharding of one pressure solve over the GPUs of one box, one process per GPU.
//
// The reference has no distributed mode at all; this is new design.  z is the slowest index
// (grid.h:70), so rank r owns the contiguous planes [k0,k1) of the global grid and keeps one ghost plane on each
// side.  Data-path communication, all on the context's stream (NCCL over NVLink 5 / NVSwitch):
//   * one-plane halo exchange (ncclSend/ncclRecv grouped, <= 2 neighbours) of the CG search vector per iteration,
//     of Ak/A0/Ai/Aj once per solve, of the pressure before correctVelocity;
//   * per iteration two tiny all-gathers of the ranks' partial reductions (p.Ap ; {|r|, r.r}); every rank then
//     combines them in rank order with the same kernel, so alpha/beta are bit-identical on all ranks.
// NCCL is not linked: it is dlopen'ed (the copy torch already loaded, or nccl_library_path), so the library still
// loads -- and the single-GPU path still runs -- without NCCL.
#include "mp_common.cuh"
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi {
	ncclResult_t (*GetUniqueId)(ncclUniqueId*);
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)();
	ncclResult_t (*GroupEnd)();
	const char* (*GetErrorString)(ncclResult_t);
	void* handle = nullptr;
} g_nccl;

int loadNccl(const char* path) {
	if (g_nccl.handle) return MP_OK;
	void* h = nullptr;
	if (path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);    // the copy the host process (torch) already loaded
	if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) MP_FAIL(MP_ERR_COMM, "mp_dist: cannot load NCCL (%s)", dlerror());
	#define SYM(name) do { *(void**)(&g_nccl.name) = dlsym(h, "nccl" #name); if (!g_nccl.name) MP_FAIL(MP_ERR_COMM, "mp_dist: NCCL symbol nccl" #name " missing"); } while (0)
	SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(AllGather); SYM(Send); SYM(Recv); SYM(GroupStart); SYM(GroupEnd); SYM(GetErrorString);
	#undef SYM
	g_nccl.handle = h;
	return MP_OK;
}
}
#define MP_NCCL(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
	mp_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); return MP_ERR_COMM; } } while (0)

int mp_dist_check_grid(const mp_grid* g) {
	const DistState* ds = g->ctx->dist;
	if (!ds || !ds->active) return MP_OK;
	if (g->sz != ds->k1 - ds->k0 + 2) MP_FAIL(MP_ERR_INVALID, "slab grid must have sz = owned planes + 2 ghost planes = %d, got %d", ds->k1 - ds->k0 + 2, g->sz);
	return MP_OK;
}

int mp_dist_halo(mp_context* ctx, void* base, size_t planeBytes, int szLocal) {
	DistState* ds = ctx->dist;
	if (!ds || !ds->active || ds->world == 1) return MP_OK;
	char* b = (char*)base;
	const int nzl = szLocal - 2;
	MP_NCCL(g_nccl.GroupStart());
	if (ds->rank > 0) {
		MP_NCCL(g_nccl.Send(b + planeBytes * 1, planeBytes, ncclChar, ds->rank - 1, (ncclComm_t)ds->comm, ctx->stream));
		MP_NCCL(g_nccl.Recv(b, planeBytes, ncclChar, ds->rank - 1, (ncclComm_t)ds->comm, ctx->stream));
	}
	if (ds->rank < ds->world - 1) {
		MP_NCCL(g_nccl.Send(b + planeBytes * nzl, planeBytes, ncclChar, ds->rank + 1, (ncclComm_t)ds->comm, ctx->stream));
		MP_NCCL(g_nccl.Recv(b + planeBytes * (nzl + 1), planeBytes, ncclChar, ds->rank + 1, (ncclComm_t)ds->comm, ctx->stream));
	}
	MP_NCCL(g_nccl.GroupEnd());
	return MP_OK;
}

int mp_dist_allgather(mp_context* ctx, int nvals) {
	DistState* ds = ctx->dist;
	(void)nvals;
	MP_NCCL(g_nccl.AllGather(ds->dLocal, ds->dGather, 8, ncclDouble, (ncclComm_t)ds->comm, ctx->stream));
	return MP_OK;
}

__global__ void k_dist_pack(const double* src, int n, double* dLocal) { for (int q = 0; q < n; q++) dLocal[q] = src[q]; }
__global__ void k_dist_sum(const double* gathered, int world, int n, double* out) {
	for (int q = 0; q < n; q++) { double s = 0; for (int r = 0; r < world; r++) s += gathered[8 * r + q]; out[q] = s; }   // fixed rank order
}
// in-place global sum of n <= 8 doubles living at device pointer vals (same result, bit for bit, on every rank)
int mp_dist_sum(mp_context* ctx, double* vals, int n) {
	DistState* ds = ctx->dist;
	if (!ds || !ds->active || ds->world == 1) return MP_OK;
	k_dist_pack<<<1, 1, 0, ctx->stream>>>(vals, n, ds->dLocal); MP_CHECK_LAUNCH(ctx);
	MP_TRY(mp_dist_allgather(ctx, n));
	k_dist_sum<<<1, 1, 0, ctx->stream>>>(ds->dGather, ds->world, n, vals); MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

extern "C" {

int mp_dist_unique_id(void* out128) {
	MP_TRY(loadNccl(nullptr));
	ncclUniqueId id;
	MP_NCCL(g_nccl.GetUniqueId(&id));
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	memcpy(out128, &id, 128);
	return MP_OK;
}

int mp_dist_init(mp_context* ctx, int rank, int world, const void* id128, const char* nccl_library_path) {
	if (!ctx || !id128) MP_FAIL(MP_ERR_INVALID, "mp_dist_init: NULL argument");
	if (world < 1 || rank < 0 || rank >= world) MP_FAIL(MP_ERR_INVALID, "mp_dist_init: bad rank %d / world %d", rank, world);
	if (ctx->dist) MP_FAIL(MP_ERR_INVALID, "mp_dist_init: context is already distributed");
	MP_TRY(loadNccl(nccl_library_path));
	MP_CUDA(cudaSetDevice(ctx->device));
	DistState* ds = new DistState();
	ds->rank = rank; ds->world = world;
	ncclUniqueId id; memcpy(&id, id128, 128);
	ncclComm_t comm;
	MP_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
	ds->comm = (void*)comm;
	MP_CUDA(cudaMalloc((void**)&ds->dGather, sizeof(double) * 8 * world));
	MP_CUDA(cudaMalloc((void**)&ds->dLocal, sizeof(double) * 8));
	MP_CUDA(cudaMemset(ds->dGather, 0, sizeof(double) * 8 * world));
	MP_CUDA(cudaMemset(ds->dLocal, 0, sizeof(double) * 8));
	ctx->dist = ds;
	return MP_OK;
}

int mp_dist_set_domain(mp_context* ctx, int sz_global) {
	if (!ctx || !ctx->dist) MP_FAIL(MP_ERR_INVALID, "mp_dist_set_domain: call mp_dist_init first");
	DistState* ds = ctx->dist;
	if (sz_global < 3 * ds->world) MP_FAIL(MP_ERR_INVALID, "mp_dist_set_domain: %d planes are too few for %d ranks", sz_global, ds->world);
	ds->gsz = sz_global;
	mp_dist_slab(sz_global, ds->rank, ds->world, &ds->k0, &ds->k1);
	ds->active = true;
	return MP_OK;
}

int mp_dist_exchange_halo(mp_context* ctx, mp_grid* g) {
	if (!ctx || !g) MP_FAIL(MP_ERR_INVALID, "mp_dist_exchange_halo: NULL argument");
	MP_TRY(mp_dist_check_grid(g));
	return mp_dist_halo(ctx, g->d, (size_t)g->sx * g->sy * g->comps() * g->elemSize(), g->sz);
}

int mp_dist_shutdown(mp_context* ctx) {
	if (!ctx || !ctx->dist) return MP_OK;
	DistState* ds = ctx->dist;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if (ds->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ds->comm);
	cudaFree(ds->dGather); cudaFree(ds->dLocal);
	delete ds; ctx->dist = nullptr;
	return MP_OK;
}

int mp_dist_slab(int sz, int rank, int world, int* k0, int* k1) {
	if (world < 1 || rank < 0 || rank >= world) MP_FAIL(MP_ERR_INVALID, "mp_dist_slab: bad rank %d / world %d", rank, world);
	const int base = sz / world, rem = sz % world;
	*k0 = rank * base + (rank < rem ? rank : rem);
	*k1 = *k0 + base + (rank < rem ? 1 : 0);
	return MP_OK;
}

}

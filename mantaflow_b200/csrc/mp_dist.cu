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
#include <algorithm>
#include <cstdlib>

namespace {
struct NcclApi {
	ncclResult_t (*GetUniqueId)(ncclUniqueId*);
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
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
	SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(AllGather); SYM(Send); SYM(Recv); SYM(Broadcast); SYM(AllReduce); SYM(GroupStart); SYM(GroupEnd); SYM(GetErrorString);
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

// boundary planes of a GLOBAL-size array of which every rank keeps its planes [K0,K1) current: plane K0 goes to the rank below (which stores
// it at the same place in its copy), plane K1-1 to the rank above; planes K0-1 and K1 arrive
int mp_dist_halo_range(mp_context* ctx, void* base, size_t planeBytes, int K0, int K1, int nplanes) {
	DistState* ds = ctx->dist;
	if (!ds || !ds->active || ds->world == 1) return MP_OK;
	char* b = (char*)base;
	MP_NCCL(g_nccl.GroupStart());
	if (ds->rank > 0 && K0 > 0 && K1 > K0) {
		MP_NCCL(g_nccl.Send(b + planeBytes * (size_t)K0, planeBytes, ncclChar, ds->rank - 1, (ncclComm_t)ds->comm, ctx->stream));
		MP_NCCL(g_nccl.Recv(b + planeBytes * (size_t)(K0 - 1), planeBytes, ncclChar, ds->rank - 1, (ncclComm_t)ds->comm, ctx->stream));
	}
	if (ds->rank < ds->world - 1 && K1 < nplanes && K1 > K0) {
		MP_NCCL(g_nccl.Send(b + planeBytes * (size_t)(K1 - 1), planeBytes, ncclChar, ds->rank + 1, (ncclComm_t)ds->comm, ctx->stream));
		MP_NCCL(g_nccl.Recv(b + planeBytes * (size_t)K1, planeBytes, ncclChar, ds->rank + 1, (ncclComm_t)ds->comm, ctx->stream));
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

// ================================================================ peer-memory path (NVLink P2P over CUDA IPC)
// The per-iteration exchanges of the CG loop do not go through NCCL: every rank exports one arena (cudaIpcGetMemHandle),
// maps its peers' arenas, and
//   * k_update_search writes its first / last owned plane of the new search vector straight into the neighbours' ghost
//     planes (peer stores over NVLink) and, once all its blocks are done, bumps a flag in their arenas;
//   * the ranks' partial reductions are scattered into every peer's gather slots with a flag per source rank, and
//     k_cg_combine_p2p spins on the flags before combining in rank order (slots are double-buffered by sequence parity).
// NCCL stays for the once-per-solve exchanges and for carrying the IPC handles.
struct PeerPtrs { char* p[16]; };

// one thread per peer: my partials -> peer's slot [parity][myRank], then the flag (release at system scope)
__global__ void k_p2p_scatter(PeerPtrs peers, int world, int rank, const double* __restrict__ local, unsigned int seq) {
	const int r = threadIdx.x;
	if (r >= world) return;
	char* base = peers.p[r] + kArenaGather;
	double* slot = (double*)base + ((seq & 1u) * 16 + rank) * kGatherStride;
	unsigned int* flag = (unsigned int*)(base + 2 * 16 * kGatherStride * sizeof(double)) + (seq & 1u) * 16 + rank;
	slot[0] = local[0]; slot[1] = local[1]; slot[2] = local[2]; slot[3] = local[3];
	__threadfence_system();
	st_release_sys(flag, seq);
}
static const unsigned long long kWaitNs = 30ull * 1000000000ull;   // a peer that is 30 s late is gone: flag the stall instead of hanging the GPU
__device__ __forceinline__ bool waitFlag(const unsigned int* flag, unsigned int seq) {
	if ((int)(ld_acquire_sys(flag) - seq) >= 0) return true;
	const unsigned long long t0 = globalTimerNs();
	while ((int)(ld_acquire_sys(flag) - seq) < 0) { if (globalTimerNs() - t0 > kWaitNs) return false; }
	return true;
}
// gather my own slots (all ranks) into dGather[r*8+q] once every flag has arrived
__global__ void k_p2p_collect(char* arena, int world, unsigned int seq, double* __restrict__ gathered, int* stall) {
	const int r = threadIdx.x;
	if (r >= world) return;
	char* base = arena + kArenaGather;
	const unsigned int* flag = (const unsigned int*)(base + 2 * 16 * kGatherStride * sizeof(double)) + (seq & 1u) * 16 + r;
	if (!waitFlag(flag, seq)) atomicExch(stall, 1);
	const double* slot = (const double*)base + ((seq & 1u) * 16 + r) * kGatherStride;
	for (int q = 0; q < 4; q++) gathered[8 * r + q] = __ldcv(slot + q);
}
// every block waits for the neighbours' flags, then the grid copies the two received planes into the ghost planes
__global__ void __launch_bounds__(256) k_p2p_halo_in(char* arena, const uint4* __restrict__ recvLo, const uint4* __restrict__ recvHi,
	uint4* __restrict__ ghostLo, uint4* __restrict__ ghostHi, size_t nvec, unsigned int seq, const int* done, int* stall)
{
	if (done && *done) return;
	if (threadIdx.x == 0) {
		const unsigned int* f = (const unsigned int*)(arena + kArenaFlags);
		bool o = true;
		if (recvLo) o = waitFlag(f + 0, seq) && o;
		if (recvHi) o = waitFlag(f + 1, seq) && o;
		if (!o) atomicExch(stall, 1);
	}
	__syncthreads();
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
		if (recvLo) ghostLo[i] = __ldcv(recvLo + i);
		if (recvHi) ghostHi[i] = __ldcv(recvHi + i);
	}
}

int mp_dist_p2p_prepare(mp_context* ctx, size_t searchBytes) {
	DistState* ds = ctx->dist;
	static const int enable = getenv("MP_P2P") ? atoi(getenv("MP_P2P")) : 1;
	if (!ds || !ds->active || ds->world == 1 || ds->world > 16 || !enable || ds->p2pUnavailable) { if (ds) ds->p2p = false; return MP_OK; }
	if (ds->arena && ds->searchBytes >= searchBytes) { ds->p2p = true; return MP_OK; }
	// (re)build: every rank takes this branch in the same call because all slabs change size together
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	for (int r = 0; r < (int)ds->peer.size(); r++) if (ds->peer[r] && r != ds->rank) cudaIpcCloseMemHandle(ds->peer[r]);
	ds->peer.clear();
	if (ds->arena) { MP_CUDA(cudaFree(ds->arena)); ds->arena = nullptr; }
	// all ranks allocate the same size: the largest slab's search vector (slabs differ by at most one plane)
	size_t maxBytes = searchBytes;
	{
		double* tmp = ds->dLocal; const double mine = (double)searchBytes;
		MP_CUDA(cudaMemcpyAsync(tmp, &mine, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		MP_TRY(mp_dist_allgather(ctx, 1));
		std::vector<double> all(8 * ds->world);
		MP_CUDA(cudaMemcpyAsync(all.data(), ds->dGather, sizeof(double) * 8 * ds->world, cudaMemcpyDeviceToHost, ctx->stream));
		MP_CUDA(cudaStreamSynchronize(ctx->stream));
		for (int r = 0; r < ds->world; r++) maxBytes = std::max(maxBytes, (size_t)all[8 * r]);
	}
	ds->searchBytes = maxBytes; ds->arenaBytes = kArenaSearch + maxBytes + 512;
	MP_CUDA(cudaMalloc((void**)&ds->arena, ds->arenaBytes));
	MP_CUDA(cudaMemset(ds->arena, 0, ds->arenaBytes));
	cudaIpcMemHandle_t h; MP_CUDA(cudaIpcGetMemHandle(&h, ds->arena));
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
	char *dH = nullptr, *dAll = nullptr;
	MP_CUDA(cudaMalloc((void**)&dH, 64)); MP_CUDA(cudaMalloc((void**)&dAll, 64 * ds->world));
	MP_CUDA(cudaMemcpy(dH, &h, 64, cudaMemcpyHostToDevice));
	MP_NCCL(g_nccl.AllGather(dH, dAll, 64, ncclChar, (ncclComm_t)ds->comm, ctx->stream));
	std::vector<cudaIpcMemHandle_t> hs(ds->world);
	MP_CUDA(cudaMemcpyAsync(hs.data(), dAll, 64 * ds->world, cudaMemcpyDeviceToHost, ctx->stream));
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	cudaFree(dH); cudaFree(dAll);
	ds->peer.assign(ds->world, nullptr);
	bool okLocal = true;
	for (int r = 0; r < ds->world; r++) {
		if (r == ds->rank) { ds->peer[r] = ds->arena; continue; }
		void* p = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&p, hs[r], cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) { cudaGetLastError(); okLocal = false; continue; }     // (no peer access / separate IPC namespace)
		ds->peer[r] = (char*)p;
	}
	if (getenv("MP_P2P_FAIL_RANK") && atoi(getenv("MP_P2P_FAIL_RANK")) == ds->rank) okLocal = false;      // test hook: pretend this rank could not map its peers
	ds->haloSeq = 0; ds->scalSeq = 0;
	// One all-gather does two jobs: nobody may write into a peer's arena before that peer has finished clearing it, and the ranks
	// agree on the outcome -- if ANY rank could not map a peer, ALL of them keep the NCCL exchanges (a mixed choice would deadlock).
	{
		const double mine = okLocal ? 1.0 : 0.0;
		MP_CUDA(cudaMemcpyAsync(ds->dLocal, &mine, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		MP_TRY(mp_dist_allgather(ctx, 1));
		std::vector<double> all(8 * ds->world);
		MP_CUDA(cudaMemcpyAsync(all.data(), ds->dGather, sizeof(double) * 8 * ds->world, cudaMemcpyDeviceToHost, ctx->stream));
		MP_CUDA(cudaStreamSynchronize(ctx->stream));
		bool allOk = true;
		for (int r = 0; r < ds->world; r++) allOk = allOk && all[8 * r] > 0.5;
		if (!allOk) {
			for (int r = 0; r < ds->world; r++) if (ds->peer[r] && r != ds->rank) cudaIpcCloseMemHandle(ds->peer[r]);
			ds->peer.clear();
			ds->p2p = false; ds->p2pUnavailable = true;
			return MP_OK;
		}
	}
	ds->p2p = true;
	return MP_OK;
}

int mp_dist_p2p_scalars(mp_context* ctx) {
	DistState* ds = ctx->dist;
	PeerPtrs pp; for (int r = 0; r < 16; r++) pp.p[r] = r < ds->world ? ds->peer[r] : nullptr;
	const unsigned int seq = ++ds->scalSeq;
	k_p2p_scatter<<<1, 32, 0, ctx->stream>>>(pp, ds->world, ds->rank, ds->dLocal, seq); MP_CHECK_LAUNCH(ctx);
	k_p2p_collect<<<1, 32, 0, ctx->stream>>>(ds->arena, ds->world, seq, ds->dGather, (int*)(ds->arena + kArenaFlags) + 8); MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

static inline size_t recvOffset(const DistState* ds, unsigned int seq, int side) { return kArenaSearch + ((seq & 1u) * 2 + side) * (ds->searchBytes / 4); }

int mp_dist_p2p_halo_out(mp_context* ctx, size_t planeBytes, HaloOut* ho) {
	DistState* ds = ctx->dist;
	*ho = HaloOut();
	if (!ds || !ds->p2p) return MP_OK;
	if (4 * planeBytes > ds->searchBytes) MP_FAIL(MP_ERR_INVALID, "mp_dist: peer arena smaller than the halo planes");
	const unsigned int seq = ++ds->haloSeq;
	ho->seq = seq; ho->ticket = ctx->tickets + 7;
	if (ds->rank > 0)             { ho->lo = ds->peer[ds->rank - 1] + recvOffset(ds, seq, 1); ho->flagLo = (unsigned int*)(ds->peer[ds->rank - 1] + kArenaFlags) + 1; }
	if (ds->rank < ds->world - 1) { ho->hi = ds->peer[ds->rank + 1] + recvOffset(ds, seq, 0); ho->flagHi = (unsigned int*)(ds->peer[ds->rank + 1] + kArenaFlags) + 0; }
	return MP_OK;
}

int mp_dist_p2p_halo_in(mp_context* ctx, void* base, size_t planeBytes, int szLocal, const int* done) {
	DistState* ds = ctx->dist;
	const unsigned int seq = ds->haloSeq;
	char* b = (char*)base;
	if (planeBytes % 16) MP_FAIL(MP_ERR_INVALID, "mp_dist: halo plane is not a multiple of 16 bytes");
	const bool lo = ds->rank > 0, hi = ds->rank < ds->world - 1;
	const size_t nvec = planeBytes / 16;
	unsigned int blocks = (unsigned int)std::min<size_t>((nvec + 255) / 256, 64);
	k_p2p_halo_in<<<blocks, 256, 0, ctx->stream>>>(ds->arena, lo ? (const uint4*)(ds->arena + recvOffset(ds, seq, 0)) : nullptr,
		hi ? (const uint4*)(ds->arena + recvOffset(ds, seq, 1)) : nullptr, (uint4*)b, (uint4*)(b + planeBytes * (size_t)(szLocal - 1)), nvec, seq, done,
		(int*)(ds->arena + kArenaFlags) + 8);
	MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

int mp_dist_p2p_check(mp_context* ctx) {
	DistState* ds = ctx->dist;
	if (!ds || !ds->p2p) return MP_OK;
	int stall = 0;
	MP_CUDA(cudaMemcpyAsync(&stall, ds->arena + kArenaFlags + 8 * sizeof(int), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	if (stall) MP_FAIL(MP_ERR_COMM, "mp_dist: a peer-memory wait timed out (a neighbouring rank stopped making progress)");
	return MP_OK;
}

__global__ void k_dist_pack(const double* src, int n, double* dLocal) { for (int q = 0; q < n; q++) dLocal[q] = src[q]; }
__global__ void k_dist_sum(const double* gathered, int world, int n, double* out) {
	for (int q = 0; q < n; q++) { double s = 0; for (int r = 0; r < world; r++) s += gathered[8 * r + q]; out[q] = s; }   // fixed rank order
}
// every rank's owned planes of a slab array (local layout: ghost, owned..., ghost) into one global array on every rank: one broadcast per
// rank, grouped (the slabs of mp_dist_slab may differ by one plane, so this is not a fixed-count all-gather)
int mp_dist_gather_planes(mp_context* ctx, const void* localBase, size_t planeBytes, void* globalBase) {
	DistState* ds = ctx->dist;
	if (!ds || !ds->active || ds->world == 1) MP_FAIL(MP_ERR_INVALID, "mp_dist_gather_planes: context is not in slab mode");
	MP_NCCL(g_nccl.GroupStart());
	for (int r = 0; r < ds->world; r++) {
		int k0, k1; mp_dist_slab(ds->gsz, r, ds->world, &k0, &k1);
		MP_NCCL(g_nccl.Broadcast((const char*)localBase + planeBytes, (char*)globalBase + planeBytes * (size_t)k0, planeBytes * (size_t)(k1 - k0), ncclChar, r,
			(ncclComm_t)ds->comm, ctx->stream));
	}
	MP_NCCL(g_nccl.GroupEnd());
	return MP_OK;
}
// in-place sum over the ranks of an array of Reals (every rank contributes zeros outside the part it computed, so the sum is exact)
int mp_dist_allreduce_sum(mp_context* ctx, void* data, size_t count, int prec) {
	DistState* ds = ctx->dist;
	if (!ds || !ds->active || ds->world == 1) return MP_OK;
	MP_NCCL(g_nccl.AllReduce(data, data, count, prec == 4 ? ncclFloat : ncclDouble, ncclSum, (ncclComm_t)ds->comm, ctx->stream));
	return MP_OK;
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
	for (int r = 0; r < (int)ds->peer.size(); r++) if (ds->peer[r] && r != ds->rank) cudaIpcCloseMemHandle(ds->peer[r]);
	if (ds->arena) cudaFree(ds->arena);
	if (ds->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ds->comm);
	cudaFree(ds->dGather); cudaFree(ds->dLocal);
	delete ds; ctx->dist = nullptr;
	return MP_OK;
}

int mp_dist_exchange_mode(const mp_context* ctx, int* mode) {
	if (!ctx || !mode) MP_FAIL(MP_ERR_INVALID, "mp_dist_exchange_mode: NULL argument");
	const DistState* ds = ctx->dist;
	*mode = (!ds || ds->world <= 1) ? 0 : (ds->p2p ? 2 : 1);
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

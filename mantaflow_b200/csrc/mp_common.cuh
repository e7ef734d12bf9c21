// Shared internals of libmantapress (sm_100a).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/mantapress.h"

typedef long long IndexInt;   // general.h:88

// FlagGrid::CellType grid.h:292-304
enum : int { TypeFluid = 1, TypeObstacle = 2, TypeEmpty = 4, TypeInflow = 8, TypeOutflow = 16, TypeOpen = 32, TypeStick = 64 };

// ---------------------------------------------------------------- errors
void mp_set_error(const char* fmt, ...);
#define MP_FAIL(code, ...) do { mp_set_error(__VA_ARGS__); return (code); } while (0)
#define MP_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	mp_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); return MP_ERR_CUDA; } } while (0)
#define MP_TRY(call) do { int rc_ = (call); if (rc_ != MP_OK) return rc_; } while (0)
#define MP_CHECK_LAUNCH(ctx) do { (ctx)->launches++; MP_CUDA(cudaGetLastError()); } while (0)

// ---------------------------------------------------------------- context / grid
struct DistState;   // below
struct mp_micrb_state;

struct mp_context {
	int device = 0;
	int smCount = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t copyStream = nullptr;
	// reduction scratch: per-block partials (double) for up to kMaxPartials blocks x kSlots values, ticket counters
	double* partials = nullptr;
	unsigned int* tickets = nullptr;
	// generic device result slots + pinned host mirror
	double* dScal = nullptr;      // 64 doubles
	double* hScal = nullptr;      // pinned, 64 doubles
	long long launches = 0;
	mp_mg* staticMg = nullptr;    // gMapMG[parent] pressure.cpp:250
	mp_mg* spareMg = nullptr;     // a released GridMg kept for reuse of its allocations (PcMGDynamic rebuilds every solve)
	DistState* dist = nullptr;
	cudaEvent_t ev[8] = {};
	// FluidSolver::GridStorage analogue (fluidsolver.cpp:33-50): freed device blocks are kept for reuse so the
	// ~12 temporary grids of a solve cost no cudaMalloc/cudaFree after the first call
	std::vector<std::pair<void*, size_t>> pool;
	size_t poolBytes = 0, poolMaxBlock = 0;
	// sampled per-kernel timing of the CG loop
	int profPeriod = 0;
	std::vector<cudaEvent_t> profEv;      // 5 events per sample: t0 | matvec | axpy | precond+dot | update
	int profCount = 0;
	float profMs[4] = {0, 0, 0, 0};
	unsigned char* micMask = nullptr; size_t micMaskBytes = 0; const void* micMaskFor = nullptr; int micMaskPrec = 0;   // per-chunk fluid bits of the MIC warp sweeps
	void* micMail = nullptr; size_t micMailBytes = 0; unsigned int micTag = 0;   // edge-row mailboxes of the warp sweeps + sweep sequence number
	void* stagePin[2] = { nullptr, nullptr }; cudaEvent_t stageEv[2]; int stageThreads = 1;     // pinned bounce buffers of the pageable-memory copies
	int* micStall = nullptr;                              // raised by a MIC sweep whose dependency wait ran out of budget
	int* micOrder = nullptr; int micOrderCount = 0;       // dispatch order of the warp columns
	int* micProg = nullptr; size_t micProgBytes = 0;     // per-column progress counters (+ stall flag) of the pipelined MIC sweeps
	int micRb = 0, micRbTY = 0, micRbTZ = 0;             // mp_set_mic_ordering: 1 = block red-black ordering (mp_micrb.cu) with tiles of TY x TZ rows
	struct mp_micrb_state* micRbState = nullptr;         // its byte mask and geometry
	int lastMatvecKernel = 0;     // which matvec instantiation the last launch used (reported in mp_solve_info)
};
static const int kMaxPartials = 1 << 16;   // max blocks of a reducing kernel
static const int kSlots = 4;               // values reduced per kernel

struct mp_grid {
	mp_context* ctx;
	int kind, prec, sx, sy, sz;
	IndexInt n;          // cells
	size_t bytes;
	void* d;
	bool owns;
	int comps() const { return kind == MP_GRID_MAC ? 3 : 1; }
	int elemSize() const { return kind == MP_GRID_FLAGS ? 4 : prec; }
	bool is3D() const { return sz > 1; }
};

// z-slab sharding state of a context (mp_dist.cu).  A slab grid holds the owned planes [k0,k1) of the global grid
// plus one ghost plane on each side: local plane kl <-> global plane k0 - 1 + kl, local sz = k1 - k0 + 2.
struct DistState {
	int rank = 0, world = 1;
	void* comm = nullptr;        // ncclComm_t
	void* nccl = nullptr;        // dlopen handle
	bool active = false;         // mp_dist_set_domain called: 3-D grids of this context are slabs
	int gsz = 0, k0 = 0, k1 = 0;
	double* dGather = nullptr;   // world * 8 doubles: all-gathered partial results
	double* dLocal = nullptr;    // 8 doubles: this rank's partial results
	// ---- peer-memory path of the per-iteration exchanges (NVLink P2P through CUDA IPC), see mp_dist.cu ----
	// arena layout (identical on every rank): [0,4K) halo flags | [4K,64K) scalar gather slots | [64K, ...) search vector slab
	bool p2p = false, p2pUnavailable = false;     // unavailable: some rank could not map a peer's arena -> every rank stays on NCCL
	char* arena = nullptr; size_t arenaBytes = 0; size_t searchBytes = 0;
	std::vector<char*> peer;     // peers' arenas mapped into this process (peer[rank] == arena)
	unsigned int haloSeq = 0, scalSeq = 0;
};
static const size_t kStageBytes = (size_t)32 << 20;
static const size_t kArenaFlags = 0, kArenaGather = 4096, kArenaSearch = 65536;
// what the kernel producing the new search vector needs to push its boundary planes to the neighbours (all null: no push)
struct HaloOut {
	char* lo = nullptr; char* hi = nullptr;                     // lower / upper neighbour's receive buffer of this parity
	unsigned int* flagLo = nullptr; unsigned int* flagHi = nullptr;
	unsigned int* ticket = nullptr; unsigned int seq = 0;
};
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) { unsigned int v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long globalTimerNs() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
static const int kGatherStride = 8;          // doubles per rank per slot

struct Dims {
	int sx, sy, sz;
	IndexInt X, Y, Z, n;     // strides (Z == 0 in 2-D, grid.cpp:55) and cell count
	bool is3D;
	// sharding: owned local planes [kb,ke), global plane of local plane 0, global sz, ranks (single GPU: 0,sz,0,sz,1)
	int kb, ke, kOff, gsz, world;
	IndexInt i0, i1;         // owned linear range [kb*Z, ke*Z) (whole grid on a single GPU / in 2-D)
};
// linear index -> (i, j, k).  A 64-bit division costs ~100 instructions; every grid of one GPU has < 2^31 cells, so the common case takes
// two 32-bit unsigned divisions (a warp-uniform branch).
__host__ __device__ __forceinline__ void cellOf(const Dims& d, IndexInt idx, int& i, int& j, int& k) {
	if (d.n <= 0x7fffffffLL) {
		const unsigned u = (unsigned)idx, t = u / (unsigned)d.sx;
		i = (int)(u - t * (unsigned)d.sx); k = (int)(t / (unsigned)d.sy); j = (int)(t - (unsigned)k * (unsigned)d.sy);
	} else {
		i = (int)(idx % d.sx); const IndexInt t = idx / d.sx; j = (int)(t % d.sy); k = (int)(t / d.sy);
	}
}
static inline Dims dimsOf(const mp_grid* g) {
	Dims d; d.sx = g->sx; d.sy = g->sy; d.sz = g->sz; d.is3D = g->sz > 1;
	d.X = 1; d.Y = g->sx; d.Z = d.is3D ? (IndexInt)g->sx * g->sy : 0; d.n = (IndexInt)g->sx * g->sy * g->sz;
	d.kb = 0; d.ke = g->sz; d.kOff = 0; d.gsz = g->sz; d.world = 1; d.i0 = 0; d.i1 = d.n;
	const DistState* ds = g->ctx ? g->ctx->dist : nullptr;
	if (ds && ds->active && d.is3D) {
		d.kb = 1; d.ke = g->sz - 1; d.kOff = ds->k0 - 1; d.gsz = ds->gsz; d.world = ds->world;
		d.i0 = d.kb * d.Z; d.i1 = d.ke * d.Z;
	}
	return d;
}
int mp_dist_check_grid(const mp_grid* g);                                  // slab grids must have sz == k1-k0+2
int mp_dist_halo(mp_context* ctx, void* base, size_t planeBytes, int szLocal);   // exchange the two ghost planes (no-op when world == 1)
int mp_dist_allgather(mp_context* ctx, int nvals);
int mp_dist_halo_range(mp_context* ctx, void* base, size_t planeBytes, int K0, int K1, int nplanes);          // boundary planes of a global-size array kept current on [K0,K1)
int mp_dist_gather_planes(mp_context* ctx, const void* localBase, size_t planeBytes, void* globalBase);   // owned planes of every rank -> a global array on every rank
int mp_dist_allreduce_sum(mp_context* ctx, void* data, size_t count, int prec);                          // in-place sum of a Real array over the ranks
int mp_dist_p2p_prepare(mp_context* ctx, size_t searchBytes);                  // collective: (re)build the arena, exchange IPC handles
int mp_dist_p2p_scalars(mp_context* ctx);                                      // dLocal -> every peer's gather slot + flag; returns the slot parity via ds->scalSeq
int mp_dist_p2p_halo_out(mp_context* ctx, size_t planeBytes, HaloOut* ho);        // next sequence number + the neighbours' receive buffers
int mp_dist_p2p_halo_in(mp_context* ctx, void* base, size_t planeBytes, int szLocal, const int* done);   // wait for the neighbours, fill the ghost planes
int mp_dist_p2p_check(mp_context* ctx);                                        // after a sync: did any wait time out?
int mp_dist_sum(mp_context* ctx, double* deviceVals, int n);                   // in-place global sum of n <= 8 doubles (no-op when world == 1)                         // dLocal[0..nvals) of every rank -> dGather[rank*8 + q]

int mp_check_same(const mp_grid* ref, const mp_grid* g, int kind, const char* name, bool optional);
extern "C" int mp_grid_create_scratch(mp_context* ctx, int kind, int prec, int sx, int sy, int sz, mp_grid** out);   // no clear pass: the caller writes every cell
int mp_check_flags_interior(mp_context* ctx, const mp_grid* flags);
int mp_set_wall_bcs_frac_impl(mp_context* ctx, const mp_grid* flags, mp_grid* vel, const mp_grid* phiObs);   // KnSetWallBcsFrac (mp_liquid.cu)   // fluid cells must not touch the outer layer

template <typename T> static inline T* dptr(const mp_grid* g) { return g ? (T*)g->d : (T*)nullptr; }

static inline unsigned int gridFor(IndexInt work, int block) {
	IndexInt b = (work + block - 1) / block;
	if (b < 1) b = 1;
	return (unsigned int)b;
}

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ double warpSum(double v) {
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ double warpMax(double v) {
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

// Block-level reduction of NV values (sum or max per slot), written to partials[slot*kMaxPartials + block];
// the last block to arrive (ticket) reduces the partials in a fixed order -> deterministic result.
// Returns true in ALL threads of the last block, with final[] valid in thread 0 only.
template <int NV>
__device__ __forceinline__ bool blockReduceFinalL(double (&v)[NV], const bool (&isMax)[NV], double* partials, unsigned int* ticket,
                                                  double (&fin)[NV], const unsigned int tid, const unsigned int nthreads,
                                                  const unsigned int blockLinear, const unsigned int numBlocks) {
	__shared__ double sh[NV][32];
	__shared__ bool amLast;
	const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
	#pragma unroll
	for (int q = 0; q < NV; q++) {
		double w = isMax[q] ? warpMax(v[q]) : warpSum(v[q]);
		if (lane == 0) sh[q][warp] = w;
	}
	__syncthreads();
	if (warp == 0) {
		#pragma unroll
		for (int q = 0; q < NV; q++) {
			double w = (lane < nwarps) ? sh[q][lane] : (isMax[q] ? -1.0 : 0.0);
			w = isMax[q] ? warpMax(w) : warpSum(w);
			if (lane == 0) partials[(size_t)q * kMaxPartials + blockLinear] = w;
		}
	}
	if (tid == 0) {
		__threadfence();
		unsigned int t = atomicAdd(ticket, 1u);
		amLast = (t == numBlocks - 1);
	}
	__syncthreads();
	if (!amLast) return false;
	__threadfence();
	#pragma unroll
	for (int q = 0; q < NV; q++) {
		double acc = isMax[q] ? -1.0 : 0.0;
		for (unsigned int b = tid; b < numBlocks; b += nthreads) {
			double p = __ldcg(&partials[(size_t)q * kMaxPartials + b]);
			acc = isMax[q] ? fmax(acc, p) : acc + p;
		}
		double w = isMax[q] ? warpMax(acc) : warpSum(acc);
		__syncthreads();
		if (lane == 0) sh[q][warp] = w;
		__syncthreads();
		if (warp == 0) {
			double u = (lane < nwarps) ? sh[q][lane] : (isMax[q] ? -1.0 : 0.0);
			u = isMax[q] ? warpMax(u) : warpSum(u);
			if (lane == 0) fin[q] = u;
		}
	}
	if (tid == 0) *ticket = 0;   // re-arm for the next launch on this stream
	return true;
}
// 1-D launch convenience: result valid in threadIdx.x == 0 of the last block
template <int NV>
__device__ __forceinline__ bool blockReduceFinal(double (&v)[NV], const bool (&isMax)[NV], double* partials, unsigned int* ticket, double (&fin)[NV]) {
	return blockReduceFinalL<NV>(v, isMax, partials, ticket, fin, threadIdx.x, blockDim.x, blockIdx.x, gridDim.x);
}
#endif

// MIC(0) in block red-black ordering (mp_micrb.cu)
int  mp_micrb_prepare(mp_context* ctx, const mp_grid* flags, const mp_grid* P, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak, bool* use);
bool mp_micrb_active(const mp_context* ctx, const mp_grid* flags, const mp_grid* P);
int  mp_micrb_init_launch(mp_context* ctx, mp_grid* P, const mp_grid* A0);
int  mp_micrb_apply_launch(mp_context* ctx, mp_grid* dst, const mp_grid* var1, const mp_grid* P, const int* doneFlag);
void mp_micrb_release(mp_context* ctx);
void mp_micrb_tiles(const mp_context* ctx, int* ty, int* tz);

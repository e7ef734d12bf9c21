// Context, Grid storage mirror and BLAS-1 / reduction entry points of libmantapress.
// Reference: Grid<T> grid.cpp:47-96,:205-210; FluidSolver::GridStorage fluidsolver.cpp:33-50;
// GridDotProduct conjugategrad.cpp:175-178; getMaxAbs grid.cpp:319-323; GridSumSqr commonkernels.h:32-35.
#include "mp_common.cuh"
#include <algorithm>
#include "mp_gridops.cuh"
#include <thread>
#include <vector>
#include <cstring>
#include <cstdlib>
#include <cstdarg>

static thread_local char g_err[1024] = "";
void mp_set_error(const char* fmt, ...) {
	va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}

extern "C" {

int mp_version(void) { return MP_VERSION; }
const char* mp_last_error(void) { return g_err; }
const char* mp_status_string(int s) {
	switch (s) {
	case MP_OK: return "MP_OK"; case MP_ERR_INVALID: return "MP_ERR_INVALID"; case MP_ERR_CUDA: return "MP_ERR_CUDA";
	case MP_ERR_DIVERGED: return "MP_ERR_DIVERGED"; case MP_ERR_NOT_SET: return "MP_ERR_NOT_SET";
	case MP_ERR_UNSUPPORTED: return "MP_ERR_UNSUPPORTED"; case MP_ERR_COMM: return "MP_ERR_COMM"; }
	return "MP_ERR_?";
}
int mp_device_count(int* count) {
	int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) { n = 0; cudaGetLastError(); }
	*count = n; return MP_OK;
}
void mp_pressure_params_default(mp_pressure_params* p) {   // pressure.cpp:481-494
	p->cgAccuracy = 1e-3; p->gfClamp = 1e-04; p->cgMaxIterFac = 1.5; p->precondition = 1; p->preconditioner = MP_PC_MIC;
	p->enforceCompatibility = 0; p->useL2Norm = 0; p->zeroPressureFixing = 0; p->surfTens = 0.;
}

int mp_context_create(int device, mp_context** out) {
	if (!out) MP_FAIL(MP_ERR_INVALID, "mp_context_create: out is NULL");
	int n = 0; mp_device_count(&n);
	if (n == 0) MP_FAIL(MP_ERR_CUDA, "mp_context_create: no CUDA device available (this library has no CPU fallback)");
	if (device < 0 || device >= n) MP_FAIL(MP_ERR_INVALID, "mp_context_create: device %d out of range (%d devices)", device, n);
	MP_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop; MP_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) MP_FAIL(MP_ERR_UNSUPPORTED, "mp_context_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
	mp_context* c = new mp_context();
	c->device = device; c->smCount = prop.multiProcessorCount;
	MP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	MP_CUDA(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
	MP_CUDA(cudaMalloc(&c->partials, sizeof(double) * kMaxPartials * kSlots));
	MP_CUDA(cudaMalloc(&c->tickets, sizeof(unsigned int) * 64));
	MP_CUDA(cudaMemset(c->tickets, 0, sizeof(unsigned int) * 64));
	MP_CUDA(cudaMalloc(&c->dScal, sizeof(double) * 64));
	MP_CUDA(cudaMemset(c->dScal, 0, sizeof(double) * 64));
	MP_CUDA(cudaHostAlloc(&c->hScal, sizeof(double) * 64, cudaHostAllocDefault));
	for (int i = 0; i < 8; i++) MP_CUDA(cudaEventCreate(&c->ev[i]));
	*out = c; return MP_OK;
}
int mp_context_destroy(mp_context* c) {
	if (!c) return MP_OK;
	cudaSetDevice(c->device);
	mp_release_mg(c);
	if (c->spareMg) { mp_mg_destroy(c->spareMg); c->spareMg = nullptr; }
	mp_dist_shutdown(c);
	cudaStreamSynchronize(c->stream);
	for (auto& pb : c->pool) cudaFree(pb.first);
	c->pool.clear();
	for (auto& e : c->profEv) cudaEventDestroy(e);
	for (int i = 0; i < 8; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
	mp_micrb_release(c);
	if (c->micProg) cudaFree(c->micProg);
	if (c->micMask) cudaFree(c->micMask);
	if (c->micOrder) cudaFree(c->micOrder);
	if (c->micStall) cudaFree(c->micStall);
	if (c->micMail) cudaFree(c->micMail);
	for (int b = 0; b < 2; b++) if (c->stagePin[b]) { cudaFreeHost(c->stagePin[b]); cudaEventDestroy(c->stageEv[b]); }
	cudaFree(c->partials); cudaFree(c->tickets); cudaFree(c->dScal); cudaFreeHost(c->hScal);
	cudaStreamDestroy(c->stream); cudaStreamDestroy(c->copyStream);
	delete c; return MP_OK;
}
// give every pooled (currently unused) device block back to the driver, e.g. before another library on the same GPU needs the memory
int mp_context_trim(mp_context* c) {
	if (!c) MP_FAIL(MP_ERR_INVALID, "mp_context_trim: NULL context");
	MP_CUDA(cudaSetDevice(c->device));
	MP_CUDA(cudaStreamSynchronize(c->stream));
	for (auto& pb : c->pool) cudaFree(pb.first);
	c->pool.clear(); c->poolBytes = 0;
	return MP_OK;
}
int mp_context_synchronize(mp_context* c) { MP_CUDA(cudaSetDevice(c->device)); MP_CUDA(cudaStreamSynchronize(c->stream)); return MP_OK; }
void* mp_context_stream(mp_context* c) { return (void*)c->stream; }
int mp_context_device(const mp_context* c) { return c->device; }
int mp_context_sm_count(const mp_context* c) { return c->smCount; }
int mp_context_kernel_launches(const mp_context* c, long long* count) { *count = c->launches; return MP_OK; }
int mp_context_set_profiling(mp_context* c, int period) {
	c->profPeriod = period < 0 ? 0 : period;
	if (period > 0 && c->profEv.empty()) { c->profEv.resize(5 * 128); for (auto& e : c->profEv) MP_CUDA(cudaEventCreate(&e)); }
	return MP_OK;
}

static int gridCreate(mp_context* ctx, int kind, int prec, int sx, int sy, int sz, mp_grid** out, bool clear) {
	if (!ctx || !out) MP_FAIL(MP_ERR_INVALID, "mp_grid_create: NULL argument");
	if (kind != MP_GRID_REAL && kind != MP_GRID_FLAGS && kind != MP_GRID_MAC) MP_FAIL(MP_ERR_INVALID, "mp_grid_create: bad kind %d", kind);
	if (kind != MP_GRID_FLAGS && prec != 4 && prec != 8) MP_FAIL(MP_ERR_INVALID, "mp_grid_create: prec must be 4 or 8, got %d", prec);
	if (sx < 1 || sy < 1 || sz < 1) MP_FAIL(MP_ERR_INVALID, "mp_grid_create: bad size %dx%dx%d", sx, sy, sz);
	MP_CUDA(cudaSetDevice(ctx->device));
	mp_grid* g = new mp_grid();
	g->ctx = ctx; g->kind = kind; g->prec = (kind == MP_GRID_FLAGS) ? 4 : prec; g->sx = sx; g->sy = sy; g->sz = sz;
	g->n = (IndexInt)sx * sy * sz; g->bytes = (size_t)g->n * g->comps() * g->elemSize(); g->owns = true;
	// +256 bytes of slack so vector loads of the last (partial) vector never leave the allocation
	g->d = nullptr;
	for (size_t q = 0; q < ctx->pool.size(); q++) if (ctx->pool[q].second == g->bytes) {
		g->d = ctx->pool[q].first; ctx->poolBytes -= g->bytes; ctx->pool.erase(ctx->pool.begin() + q); break; }
	if (!g->d) {
		cudaError_t e = cudaMalloc(&g->d, g->bytes + 256);
		if (e != cudaSuccess && !ctx->pool.empty()) {      // give pooled blocks back and retry once
			for (auto& pb : ctx->pool) cudaFree(pb.first);
			ctx->pool.clear(); ctx->poolBytes = 0; cudaGetLastError();
			e = cudaMalloc(&g->d, g->bytes + 256);
		}
		if (e != cudaSuccess) { const size_t b = g->bytes; delete g; cudaGetLastError(); mp_set_error("mp_grid_create: cudaMalloc(%zu) failed: %s", b, cudaGetErrorString(e)); return MP_ERR_CUDA; }
	}
	if (clear) MP_CUDA(cudaMemsetAsync(g->d, 0, g->bytes + 256, ctx->stream));     // Grid<T>(parent) clears, grid.cpp:57
	*out = g; return MP_OK;
}
int mp_grid_create(mp_context* ctx, int kind, int prec, int sx, int sy, int sz, mp_grid** out) { return gridCreate(ctx, kind, prec, sx, sy, sz, out, true); }
// internal: scratch grid whose every cell the caller writes before reading (no clear pass)
int mp_grid_create_scratch(mp_context* ctx, int kind, int prec, int sx, int sy, int sz, mp_grid** out) { return gridCreate(ctx, kind, prec, sx, sy, sz, out, false); }
int mp_grid_destroy(mp_grid* g) {
	if (!g) return MP_OK;
	cudaSetDevice(g->ctx->device);
	if (g->owns) {
		// stream-ordered reuse: every consumer of a pooled block runs on ctx->stream, so no synchronisation is needed
		mp_context* ctx = g->ctx;
		// the pool is bounded by BYTES: a solve parks ~12 grids, so room for 16 blocks of the largest size seen (at least 1 GiB) is kept and
		// the oldest blocks go first -- a process that changes grid sizes (or grows particle arrays) does not pile up dead blocks
		if (g->bytes > ctx->poolMaxBlock) ctx->poolMaxBlock = g->bytes;
		const size_t cap = std::max((size_t)1 << 30, 16 * ctx->poolMaxBlock);
		ctx->pool.push_back(std::make_pair(g->d, g->bytes)); ctx->poolBytes += g->bytes;
		if (ctx->poolBytes > cap || ctx->pool.size() > 256) {
			cudaStreamSynchronize(ctx->stream);
			while (!ctx->pool.empty() && (ctx->poolBytes > cap || ctx->pool.size() > 256)) {
				cudaFree(ctx->pool.front().first); ctx->poolBytes -= ctx->pool.front().second; ctx->pool.erase(ctx->pool.begin());
			}
		}
	}
	delete g; return MP_OK;
}
// Host <-> device copies.  Pinned (cudaHostAlloc'ed / registered) host memory goes to the DMA engine directly.  Pageable memory -- what
// the reference's Grid<T>::mData is (new T[], fluidsolver.cpp:37) -- would be staged by the driver through one bounce buffer at ~10 GB/s;
// here it is pipelined through two pinned 32 MiB buffers filled / drained by a few host threads, so that the CPU copy of one chunk
// overlaps the DMA of the other.  Small grids (< 16 MiB) keep the plain copy: the bounce buffers would cost more than they save.
static const size_t kStageMinBytes = (size_t)16 << 20;
static bool hostIsPinned(const void* p) {
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeHost;
}
static void parallelCopy(void* dst, const void* src, size_t bytes, int nthreads) {
	if (bytes < ((size_t)4 << 20) || nthreads <= 1) { memcpy(dst, src, bytes); return; }
	std::vector<std::thread> th;
	const size_t per = ((bytes / nthreads) + 4095) & ~(size_t)4095;
	for (int t = 0; t < nthreads; t++) {
		const size_t off = (size_t)t * per;
		if (off >= bytes) break;
		const size_t len = off + per > bytes ? bytes - off : per;
		th.emplace_back([=]() { memcpy((char*)dst + off, (const char*)src + off, len); });
	}
	for (auto& t : th) t.join();
}
static int stagerInit(mp_context* ctx) {
	if (ctx->stagePin[0]) return MP_OK;
	for (int b = 0; b < 2; b++) { MP_CUDA(cudaHostAlloc(&ctx->stagePin[b], kStageBytes, cudaHostAllocDefault)); MP_CUDA(cudaEventCreateWithFlags(&ctx->stageEv[b], cudaEventDisableTiming)); }
	const unsigned hc = std::thread::hardware_concurrency();
	ctx->stageThreads = getenv("MP_COPY_THREADS") ? atoi(getenv("MP_COPY_THREADS")) : (int)(hc >= 16 ? 8 : (hc >= 4 ? hc / 2 : 1));
	return MP_OK;
}
static int stagedUpload(mp_context* ctx, void* dev, const void* host, size_t bytes) {
	MP_TRY(stagerInit(ctx));
	int b = 0;
	for (size_t off = 0; off < bytes; off += kStageBytes, b ^= 1) {
		const size_t len = bytes - off < kStageBytes ? bytes - off : kStageBytes;
		MP_CUDA(cudaEventSynchronize(ctx->stageEv[b]));                    // the DMA that last read this buffer is done
		parallelCopy(ctx->stagePin[b], (const char*)host + off, len, ctx->stageThreads);
		MP_CUDA(cudaMemcpyAsync((char*)dev + off, ctx->stagePin[b], len, cudaMemcpyHostToDevice, ctx->stream));
		MP_CUDA(cudaEventRecord(ctx->stageEv[b], ctx->stream));
	}
	return MP_OK;
}
static int stagedDownload(mp_context* ctx, void* host, const void* dev, size_t bytes) {       // returns with the host copy complete
	MP_TRY(stagerInit(ctx));
	int b = 0; size_t prevOff = 0, prevLen = 0; bool havePrev = false;
	for (size_t off = 0; off < bytes; off += kStageBytes, b ^= 1) {
		const size_t len = bytes - off < kStageBytes ? bytes - off : kStageBytes;
		MP_CUDA(cudaMemcpyAsync(ctx->stagePin[b], (const char*)dev + off, len, cudaMemcpyDeviceToHost, ctx->stream));
		MP_CUDA(cudaEventRecord(ctx->stageEv[b], ctx->stream));
		if (havePrev) { MP_CUDA(cudaEventSynchronize(ctx->stageEv[b ^ 1])); parallelCopy((char*)host + prevOff, ctx->stagePin[b ^ 1], prevLen, ctx->stageThreads); }
		prevOff = off; prevLen = len; havePrev = true;
	}
	if (havePrev) { MP_CUDA(cudaEventSynchronize(ctx->stageEv[b ^ 1])); parallelCopy((char*)host + prevOff, ctx->stagePin[b ^ 1], prevLen, ctx->stageThreads); }
	return MP_OK;
}
int mp_grid_upload_async(mp_grid* g, const void* host) {
	if (g->bytes < kStageMinBytes || hostIsPinned(host)) { MP_CUDA(cudaMemcpyAsync(g->d, host, g->bytes, cudaMemcpyHostToDevice, g->ctx->stream)); return MP_OK; }
	return stagedUpload(g->ctx, g->d, host, g->bytes);
}
int mp_grid_download_async(const mp_grid* g, void* host) {
	if (g->bytes < kStageMinBytes || hostIsPinned(host)) { MP_CUDA(cudaMemcpyAsync(host, g->d, g->bytes, cudaMemcpyDeviceToHost, g->ctx->stream)); return MP_OK; }
	return stagedDownload(g->ctx, host, g->d, g->bytes);
}
int mp_grid_upload(mp_grid* g, const void* host) {
	MP_CUDA(cudaSetDevice(g->ctx->device));
	MP_TRY(mp_grid_upload_async(g, host));
	MP_CUDA(cudaStreamSynchronize(g->ctx->stream)); return MP_OK;
}
int mp_grid_download(const mp_grid* g, void* host) {
	MP_CUDA(cudaSetDevice(g->ctx->device));
	MP_TRY(mp_grid_download_async(g, host));
	MP_CUDA(cudaStreamSynchronize(g->ctx->stream)); return MP_OK;
}
int mp_grid_clear(mp_grid* g) { MP_CUDA(cudaMemsetAsync(g->d, 0, g->bytes, g->ctx->stream)); return MP_OK; }
int mp_grid_copy_from(mp_grid* dst, const mp_grid* src) {
	if (dst->bytes != src->bytes || dst->kind != src->kind) MP_FAIL(MP_ERR_INVALID, "mp_grid_copy_from: grids differ in size or kind");
	MP_CUDA(cudaMemcpyAsync(dst->d, src->d, src->bytes, cudaMemcpyDeviceToDevice, dst->ctx->stream)); return MP_OK;
}
void* mp_grid_device_ptr(mp_grid* g) { return g->d; }
int mp_grid_info(const mp_grid* g, int* kind, int* prec, int* sx, int* sy, int* sz) {
	if (kind) *kind = g->kind; if (prec) *prec = g->prec; if (sx) *sx = g->sx; if (sy) *sy = g->sy; if (sz) *sz = g->sz; return MP_OK;
}
int mp_host_alloc(void** out, unsigned long long bytes) { MP_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault)); return MP_OK; }
int mp_host_free(void* p) { MP_CUDA(cudaFreeHost(p)); return MP_OK; }

} // extern "C"

int mp_check_same(const mp_grid* ref, const mp_grid* g, int kind, const char* name, bool optional) {
	if (!g) { if (optional) return MP_OK; MP_FAIL(MP_ERR_INVALID, "grid '%s' is NULL", name); }
	if (g->kind != kind) MP_FAIL(MP_ERR_INVALID, "grid '%s' has kind %d, expected %d", name, g->kind, kind);
	if (g->sx != ref->sx || g->sy != ref->sy || g->sz != ref->sz) MP_FAIL(MP_ERR_INVALID, "grid '%s' is %dx%dx%d, expected %dx%dx%d", name, g->sx, g->sy, g->sz, ref->sx, ref->sy, ref->sz);
	if (kind != MP_GRID_FLAGS && ref->kind != MP_GRID_FLAGS && g->prec != ref->prec) MP_FAIL(MP_ERR_INVALID, "grid '%s' has precision %d, expected %d", name, g->prec, ref->prec);
	if (g->ctx != ref->ctx) MP_FAIL(MP_ERR_INVALID, "grid '%s' belongs to another context", name);
	return MP_OK;
}

// ---------------------------------------------------------------- BLAS-1 / reductions
template <typename Real, int MODE>   // MODE 0: dot(a,b)  1: max|a|  2: sum (double)a^2
__global__ void __launch_bounds__(256) k_reduce(const Real* __restrict__ a, const Real* __restrict__ b, IndexInt n,
                                               double* partials, unsigned int* ticket, double* out) {
	double v[1] = { MODE == 1 ? -1.0 : 0.0 };
	for (IndexInt i = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (IndexInt)gridDim.x * blockDim.x) {
		if (MODE == 0) v[0] += (double)(a[i] * b[i]);          // product in Real, accumulation in double
		else if (MODE == 1) v[0] = fmax(v[0], fabs((double)a[i]));
		else { double x = (double)a[i]; v[0] += x * x; }
	}
	const bool isMax[1] = { MODE == 1 };
	double fin[1];
	if (blockReduceFinal<1>(v, isMax, partials, ticket, fin) && threadIdx.x == 0) out[0] = fin[0];
}
template <typename Real>
__global__ void __launch_bounds__(256) k_scaled_add(Real* __restrict__ me, const Real* __restrict__ other, Real f, IndexInt n) {
	for (IndexInt i = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (IndexInt)gridDim.x * blockDim.x) me[i] += f * other[i];
}
template <typename Real>
__global__ void __launch_bounds__(256) k_add_const(Real* __restrict__ me, Real c, IndexInt n) {
	for (IndexInt i = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (IndexInt)gridDim.x * blockDim.x) me[i] += c;
}

// Grid<Vec3>::getMaxAbs grid.cpp:330-332: sqrt(CompMaxVec) = the largest |v|; normSquare in Real, the maximum is order independent
template <typename Real>
__global__ void __launch_bounds__(256) k_reduce_vec_max(const Real* __restrict__ a, IndexInt n, double* partials, unsigned int* ticket, double* out) {
	double v[1] = { -1.0 };
	for (IndexInt i = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (IndexInt)gridDim.x * blockDim.x) {
		const Real x = a[3 * i], y = a[3 * i + 1], z = a[3 * i + 2];
		const Real s = x * x + y * y + z * z;
		v[0] = fmax(v[0], (double)s);
	}
	const bool isMax[1] = { true };
	double fin[1];
	if (blockReduceFinal<1>(v, isMax, partials, ticket, fin) && threadIdx.x == 0) out[0] = fin[0];
}
static int vecMaxAbs(mp_context* ctx, const mp_grid* a, double* out) {
	MP_CUDA(cudaSetDevice(ctx->device));
	unsigned int blocks = gridFor(a->n, 256 * 8); if (blocks > (unsigned)ctx->smCount * 8) blocks = ctx->smCount * 8;
	if (a->prec == 4) k_reduce_vec_max<float><<<blocks, 256, 0, ctx->stream>>>((const float*)a->d, a->n, ctx->partials, ctx->tickets + 0, ctx->dScal);
	else              k_reduce_vec_max<double><<<blocks, 256, 0, ctx->stream>>>((const double*)a->d, a->n, ctx->partials, ctx->tickets + 0, ctx->dScal);
	MP_CHECK_LAUNCH(ctx);
	MP_CUDA(cudaMemcpyAsync(ctx->hScal, ctx->dScal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	*out = a->prec == 4 ? (double)sqrtf((float)ctx->hScal[0]) : sqrt(ctx->hScal[0]);       // the maximum is a Real value: the cast back is exact
	return MP_OK;
}

template <int MODE>
static int reduceLaunch(mp_context* ctx, const mp_grid* a, const mp_grid* b, double* out) {
	if (!a || a->kind != MP_GRID_REAL) MP_FAIL(MP_ERR_INVALID, "reduction: grid must be a Real grid");
	if (b) MP_TRY(mp_check_same(a, b, MP_GRID_REAL, "b", false));
	MP_CUDA(cudaSetDevice(ctx->device));
	unsigned int blocks = gridFor(a->n, 256 * 8); if (blocks > (unsigned)ctx->smCount * 8) blocks = ctx->smCount * 8;
	if (a->prec == 4) k_reduce<float, MODE><<<blocks, 256, 0, ctx->stream>>>((const float*)a->d, b ? (const float*)b->d : nullptr, a->n, ctx->partials, ctx->tickets + 0, ctx->dScal);
	else              k_reduce<double, MODE><<<blocks, 256, 0, ctx->stream>>>((const double*)a->d, b ? (const double*)b->d : nullptr, a->n, ctx->partials, ctx->tickets + 0, ctx->dScal);
	MP_CHECK_LAUNCH(ctx);
	MP_CUDA(cudaMemcpyAsync(ctx->hScal, ctx->dScal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	*out = ctx->hScal[0]; return MP_OK;
}

template <typename T>
__global__ void __launch_bounds__(256) k_grid_arith(gridops::Op<T> f, IndexInt n) {
	for (IndexInt e = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (IndexInt)gridDim.x * blockDim.x) f(e);
}
template <typename T>
static int gridArith(mp_context* ctx, mp_grid* me, int op, const mp_grid* other, double x, double y, double z) {
	gridops::Op<T> f;
	f.me = (T*)me->d; f.other = other ? (const T*)other->d : nullptr; f.op = op; f.comps = me->comps();
	const double c[3] = { x, me->comps() == 3 ? y : x, me->comps() == 3 ? z : x };
	for (int q = 0; q < 3; q++) { f.c0[q] = (T)(op == MP_OP_CLAMP ? x : c[q]); f.c1[q] = (T)y; }
	const IndexInt n = me->n * me->comps();
	unsigned int blocks = gridFor(n, 256 * 4); if (blocks > (unsigned)ctx->smCount * 16) blocks = ctx->smCount * 16;
	k_grid_arith<T><<<blocks, 256, 0, ctx->stream>>>(f, n);
	MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

extern "C" {
int mp_grid_arith(mp_context* ctx, mp_grid* me, int op, const mp_grid* other, double x, double y, double z) {
	if (!ctx || !me) MP_FAIL(MP_ERR_INVALID, "mp_grid_arith: NULL argument");
	if (op < MP_OP_SET_CONST || op > MP_OP_SAFE_DIVIDE) MP_FAIL(MP_ERR_INVALID, "mp_grid_arith: unknown operation %d", op);
	const bool binary = op == MP_OP_ADD || op == MP_OP_SUB || op == MP_OP_MULT || op == MP_OP_ADD_SCALED || op == MP_OP_SAFE_DIVIDE;
	if (binary) { MP_TRY(mp_check_same(me, other, me->kind, "other", false)); if (other->d == me->d && op != MP_OP_ADD && op != MP_OP_MULT) MP_FAIL(MP_ERR_INVALID, "mp_grid_arith: other must be another grid"); }
	else if (other) MP_FAIL(MP_ERR_INVALID, "mp_grid_arith: this operation takes no second grid");
	if (ctx->dist && ctx->dist->active && me->sz > 1) MP_TRY(mp_dist_check_grid(me));       // element-wise: ghost planes included, nothing to exchange
	MP_CUDA(cudaSetDevice(ctx->device));
	if (me->kind == MP_GRID_FLAGS) return gridArith<int>(ctx, me, op, other, x, y, z);
	if (me->prec == 4) return gridArith<float>(ctx, me, op, other, x, y, z);
	return gridArith<double>(ctx, me, op, other, x, y, z);
}
int mp_grid_dot(mp_context* ctx, const mp_grid* a, const mp_grid* b, double* out) { return reduceLaunch<0>(ctx, a, b, out); }
int mp_grid_max_abs(mp_context* ctx, const mp_grid* a, double* out) {
	if (a && ctx && out && a->kind == MP_GRID_MAC) return vecMaxAbs(ctx, a, out);        // Grid<Vec3>::getMaxAbs (adaptTimestep(vel.getMaxAbs()) in the liquid scenes)
	return reduceLaunch<1>(ctx, a, nullptr, out);
}
int mp_grid_sum_sqr(mp_context* ctx, const mp_grid* a, double* out) { return reduceLaunch<2>(ctx, a, nullptr, out); }
int mp_grid_scaled_add(mp_context* ctx, mp_grid* me, const mp_grid* other, double factor) {
	MP_TRY(mp_check_same(me, other, MP_GRID_REAL, "other", false));
	unsigned int blocks = gridFor(me->n, 256 * 4);
	if (me->prec == 4) k_scaled_add<float><<<blocks, 256, 0, ctx->stream>>>((float*)me->d, (const float*)other->d, (float)factor, me->n);
	else               k_scaled_add<double><<<blocks, 256, 0, ctx->stream>>>((double*)me->d, (const double*)other->d, factor, me->n);
	MP_CHECK_LAUNCH(ctx); return MP_OK;
}
int mp_grid_add_const(mp_context* ctx, mp_grid* me, double value) {
	if (me->kind != MP_GRID_REAL) MP_FAIL(MP_ERR_INVALID, "mp_grid_add_const: Real grid expected");
	unsigned int blocks = gridFor(me->n, 256 * 4);
	if (me->prec == 4) k_add_const<float><<<blocks, 256, 0, ctx->stream>>>((float*)me->d, (float)value, me->n);
	else               k_add_const<double><<<blocks, 256, 0, ctx->stream>>>((double*)me->d, value, me->n);
	MP_CHECK_LAUNCH(ctx); return MP_OK;
}
}

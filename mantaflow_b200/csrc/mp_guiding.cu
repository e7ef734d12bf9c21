// PD_fluid_guiding plugin/fluidguiding.cpp:294-353: the primal-dual guiding loop, up to maxIters solvePressure calls per
// step on copies of the velocity.  Everything stays on the device: the eleven MAC temporaries, the two separable Gaussian blurs per
// iteration (apply1DKernelDirX/Y/Z :49-82, applySeparableKernel2D/3D :85-130), the MACGrid algebra (fused per stage, every operation
// rounded on its own as in grid.h:478-486) and the stop test (getRNorm :140-145, getEpsDual :165-168); two scalars per iteration go to the host.
// With PcMGStatic the hierarchy is built once and reused by all solves (pressure.cpp:421-431).
// The blur coefficients are computed on the host exactly as get1DGaussianBlurKernel :30-45 does through the sparse Matrix class
// (util/rcmatrix.h: increments <= VECTOR_EPSILON are dropped, :186-187).  Float results are bit-identical to the reference's.
#include "mp_common.cuh"
#include <cmath>
#include <vector>

namespace {

#define PD_MAXK 129
template <typename Real> struct BlurK { int kn; Real w[PD_MAXK]; };

template <typename Real> static inline Real hostExp(Real v);
template <> inline float hostExp<float>(float v) { return expf(v); }
template <> inline double hostExp<double>(double v) { return exp(v); }

template <typename Real>
void blurKernel(int n, BlurK<Real>& K) {
	const Real eps = sizeof(Real) == 4 ? (Real)1e-6f : (Real)1e-10;      // VECTOR_EPSILON vectorbase.h
	const int sigma = n;
	K.kn = n;
	Real sumG = 0;
	for (int j = 0; j < n; j++) {
		Real xv = (Real)(-(n - 1) * 0.5), yv = (Real)(j - (n - 1) * 0.5);
		if (!(std::abs(xv) > eps)) xv = 0;
		if (!(std::abs(yv) > eps)) yv = 0;
		Real g = (Real)(1 / (2 * M_PI * sigma * sigma) * hostExp<Real>(-(xv * xv + yv * yv) / (2 * sigma * sigma)));
		if (!(std::abs(g) > eps)) g = 0;
		K.w[j] = g;
		sumG += K.w[j];
	}
	const double k = 1.0 / sumG;
	for (int j = 0; j < n; j++) { Real v = (Real)(K.w[j] * k); if (!(std::abs(v) > eps)) v = 0; K.w[j] = v; }
}

__device__ __forceinline__ bool cellOfG(const Dims& d, int& i, int& j, int& k, IndexInt& idx) {
	i = blockIdx.x * blockDim.x + threadIdx.x; j = blockIdx.y; k = blockIdx.z;
	idx = (IndexInt)i + d.Y * j + (IndexInt)d.sx * d.sy * k;
	return i < d.sx;
}
static inline dim3 cellGridG(const Dims& d) { return dim3((unsigned)((d.sx + 127) / 128), (unsigned)d.sy, (unsigned)d.sz); }

// out = in convolved with the 1-D kernel along DIR, truncated at the grid border, accumulated in the reference's order (m = 0 .. kn-1)
template <typename Real, int DIR>
__global__ void __launch_bounds__(128) k_pd_conv1d(Dims d, const Real* __restrict__ in, Real* __restrict__ out, BlurK<Real> K) {
	int i, j, k; IndexInt idx;
	if (!cellOfG(d, i, j, k, idx)) return;
	const int pos = DIR == 0 ? i : (DIR == 1 ? j : k), size = DIR == 0 ? d.sx : (DIR == 1 ? d.sy : d.sz);
	const IndexInt stride = DIR == 0 ? d.X : (DIR == 1 ? d.Y : d.Z);
	const int kCentre = K.kn / 2;
	Real a0 = 0, a1 = 0, a2 = 0;
	for (int m = 0, ind = K.kn - 1, q = pos - kCentre; m < K.kn; m++, ind--, q++) {
		if (q < 0) continue;
		else if (q >= size) break;
		const Real* v = in + 3 * (idx + (IndexInt)(q - pos) * stride);
		const Real w = K.w[ind];
		a0 += v[0] * w; a1 += v[1] * w; a2 += v[2] * w;
	}
	out[3 * idx] = a0; out[3 * idx + 1] = a1; out[3 * idx + 2] = a2;
}
// grid = blurred, except on faces that touch an obstacle cell, which keep their value (:98-104, :121-127)
template <typename Real>
__global__ void __launch_bounds__(128) k_pd_blur_finish(Dims d, const int* __restrict__ flags, Real* __restrict__ grid, const Real* __restrict__ blurred, const Real* __restrict__ orig) {
	int i, j, k; IndexInt idx;
	if (!cellOfG(d, i, j, k, idx)) return;
	const bool keep = (i > 0 && (flags[idx - d.X] & TypeObstacle)) || (j > 0 && (flags[idx - d.Y] & TypeObstacle)) ||
	                  (d.is3D && k > 0 && (flags[idx - d.Z] & TypeObstacle)) || (flags[idx] & TypeObstacle);
	const Real* s = keep ? orig : blurred;
	grid[3 * idx] = s[3 * idx]; grid[3 * idx + 1] = s[3 * idx + 1]; grid[3 * idx + 2] = s[3 * idx + 2];
}

// ---- fused stages of the MACGrid algebra; q runs over the 3N scalars, c = q / 3 is the cell (invA has one value per cell, :254-264)
template <typename Real>
__global__ void __launch_bounds__(256) k_pd_inv_a(IndexInt n, const Real* __restrict__ weight, Real sigma, Real* __restrict__ invA) {
	const IndexInt c = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n) return;
	Real val = 2 * weight[c] * weight[c] + sigma;
	if (val < 0.01) val = (Real)0.01;
	invA[c] = (Real)(1.0 / val);
}
template <typename Real>      // Q = velT - velC (before the blurs)
__global__ void __launch_bounds__(256) k_pd_q0(IndexInt n3, const Real* __restrict__ velT, const Real* __restrict__ velC, Real* __restrict__ Q) {
	const IndexInt q = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	if (q < n3) Q[q] = velT[q] - velC[q];
}
template <typename Real>      // Q = Q*2 + (-sigma)*velC (after the blurs)
__global__ void __launch_bounds__(256) k_pd_q1(IndexInt n3, Real sigma, const Real* __restrict__ velC, Real* __restrict__ Q) {
	const IndexInt q = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n3) return;
	Real v = Q[q] * (Real)2.0;
	v = v + (-sigma) * velC[q];
	Q[q] = v;
}
// x-update up to the blurs (:316-319, prox_f :267-270, applyApproxInvM :229-232): x0 = x; x = (x/sigma + y)*sigma + Q; vn = x*invA
template <typename Real>
__global__ void __launch_bounds__(256) k_pd_stage_a(IndexInt n3, Real invSigma, Real sigma, Real* __restrict__ x, Real* __restrict__ x0, const Real* __restrict__ y,
	const Real* __restrict__ Q, const Real* __restrict__ invA, Real* __restrict__ vn) {
	const IndexInt q = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n3) return;
	Real v = x[q];
	x0[q] = v;
	v = v * invSigma; v = v + y[q];
	v = v * sigma; v = v + Q[q];
	x[q] = v;
	vn[q] = v * invA[q / 3];
}
// rest of the x-update and the z-update before the solve (:233-238, :271, :320, :323-324)
template <typename Real>
__global__ void __launch_bounds__(256) k_pd_stage_b(IndexInt n3, Real sigma, Real tau, Real* __restrict__ x, const Real* __restrict__ x0, const Real* __restrict__ y,
	const Real* __restrict__ velC, const Real* __restrict__ invA, const Real* __restrict__ vn, Real* __restrict__ z, Real* __restrict__ z0) {
	const IndexInt q = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n3) return;
	const Real ia = invA[q / 3];
	Real w = vn[q] * (Real)2.0; w = w * ia;
	Real v = x[q] * ia; v = v - w;
	v = v + velC[q];
	v = v * (-sigma); v = v + sigma * y[q]; v = v + x0[q];
	x[q] = v;
	const Real zo = z[q];
	z0[q] = zo;
	z[q] = zo + (-tau) * v;
}
__device__ __forceinline__ void atomicMaxNonNeg(float* a, float v) { atomicMax((int*)a, __float_as_int(v)); }
__device__ __forceinline__ void atomicMaxNonNeg(double* a, double v) { atomicMax((unsigned long long*)a, (unsigned long long)__double_as_longlong(v)); }
// y-update (:330-333) and the two max-norms of the stop test: out[0] = max |z - z0|^2, out[1] = max |z|^2 over the cells
template <typename Real>
__global__ void __launch_bounds__(256) k_pd_stage_c(IndexInt n, Real theta, const Real* __restrict__ z, const Real* __restrict__ z0, Real* __restrict__ y, Real* out) {
	const IndexInt c = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	Real r2 = 0, z2 = 0;
	if (c < n) {
		Real rr[3], zz[3];
		#pragma unroll
		for (int e = 0; e < 3; e++) {
			const Real zv = z[3 * c + e], zo = z0[3 * c + e];
			Real v = zv - zo;
			rr[e] = v; zz[e] = zv;
			v = v * theta; v = v + zv;
			y[3 * c + e] = v;
		}
		r2 = rr[0] * rr[0] + rr[1] * rr[1] + rr[2] * rr[2];
		z2 = zz[0] * zz[0] + zz[1] * zz[1] + zz[2] * zz[2];
	}
	// max is exact in any order: warp, then one atomic per warp
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, o)); z2 = fmax(z2, __shfl_xor_sync(0xffffffffu, z2, o)); }
	if ((threadIdx.x & 31) == 0) { atomicMaxNonNeg(out, r2); atomicMaxNonNeg(out + 1, z2); }
}

struct Scratch { std::vector<mp_grid*> gs; ~Scratch() { for (mp_grid* g : gs) mp_grid_destroy(g); }
	int mac(mp_context* ctx, const mp_grid* like, mp_grid** out, bool clear) {
		const int rc = clear ? mp_grid_create(ctx, MP_GRID_MAC, like->prec, like->sx, like->sy, like->sz, out) : mp_grid_create_scratch(ctx, MP_GRID_MAC, like->prec, like->sx, like->sy, like->sz, out);
		if (rc == MP_OK) gs.push_back(*out);
		return rc;
	} };

template <typename Real>
int blur(mp_context* ctx, const Dims& d, const mp_grid* flags, mp_grid* grid, mp_grid* orig, mp_grid* t1, mp_grid* t2, const BlurK<Real>& K) {
	const dim3 cg = cellGridG(d);
	MP_CUDA(cudaMemcpyAsync(orig->d, grid->d, grid->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	k_pd_conv1d<Real, 0><<<cg, 128, 0, ctx->stream>>>(d, (const Real*)grid->d, (Real*)t1->d, K); MP_CHECK_LAUNCH(ctx);
	k_pd_conv1d<Real, 1><<<cg, 128, 0, ctx->stream>>>(d, (const Real*)t1->d, (Real*)t2->d, K); MP_CHECK_LAUNCH(ctx);
	const mp_grid* res = t2;
	if (d.is3D) { k_pd_conv1d<Real, 2><<<cg, 128, 0, ctx->stream>>>(d, (const Real*)t2->d, (Real*)t1->d, K); MP_CHECK_LAUNCH(ctx); res = t1; }
	k_pd_blur_finish<Real><<<cg, 128, 0, ctx->stream>>>(d, (const int*)flags->d, (Real*)grid->d, (const Real*)res->d, (const Real*)orig->d); MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

template <typename Real>
int guide(mp_context* ctx, mp_grid* vel, const mp_grid* velT, mp_grid* pressure, const mp_grid* flags, const mp_grid* weight, int blurRadius,
	double theta_, double tau_, double sigma_, double epsRel_, double epsAbs_, int maxIters, const mp_grid* phi, const mp_grid* perCellCorr,
	const mp_grid* fractions, const mp_grid* obvel, const mp_pressure_params& pp, const mp_grid* curv, int* iterations)
{
	const Dims d = dimsOf(flags);
	const IndexInt n = d.n, n3 = 3 * d.n;
	const Real theta = (Real)theta_, tau = (Real)tau_, sigma = (Real)sigma_, epsRel = (Real)epsRel_, epsAbs = (Real)epsAbs_;
	BlurK<Real> K; blurKernel<Real>(2 * blurRadius + 1, K);
	Scratch sc;
	mp_grid *velC, *x, *y, *z, *x0, *z0, *Q, *vn, *t1, *t2, *orig, *invA;
	MP_TRY(sc.mac(ctx, vel, &velC, false)); MP_TRY(sc.mac(ctx, vel, &x, true)); MP_TRY(sc.mac(ctx, vel, &y, true)); MP_TRY(sc.mac(ctx, vel, &z, true));
	MP_TRY(sc.mac(ctx, vel, &x0, false)); MP_TRY(sc.mac(ctx, vel, &z0, false)); MP_TRY(sc.mac(ctx, vel, &Q, false)); MP_TRY(sc.mac(ctx, vel, &vn, false));
	MP_TRY(sc.mac(ctx, vel, &t1, false)); MP_TRY(sc.mac(ctx, vel, &t2, false)); MP_TRY(sc.mac(ctx, vel, &orig, false));
	MP_TRY(mp_grid_create_scratch(ctx, MP_GRID_REAL, vel->prec, vel->sx, vel->sy, vel->sz, &invA)); sc.gs.push_back(invA);
	const unsigned int b3 = gridFor(n3, 256), b1 = gridFor(n, 256);
	MP_CUDA(cudaMemcpyAsync(velC->d, vel->d, vel->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	// precomputeQ :243-250, precomputeInvA :254-264
	k_pd_q0<Real><<<b3, 256, 0, ctx->stream>>>(n3, (const Real*)velT->d, (const Real*)velC->d, (Real*)Q->d); MP_CHECK_LAUNCH(ctx);
	MP_TRY(blur<Real>(ctx, d, flags, Q, orig, t1, t2, K)); MP_TRY(blur<Real>(ctx, d, flags, Q, orig, t1, t2, K));
	k_pd_q1<Real><<<b3, 256, 0, ctx->stream>>>(n3, sigma, (const Real*)velC->d, (Real*)Q->d); MP_CHECK_LAUNCH(ctx);
	k_pd_inv_a<Real><<<b1, 256, 0, ctx->stream>>>(n, (const Real*)weight->d, sigma, (Real*)invA->d); MP_CHECK_LAUNCH(ctx);
	const Real invSigma = (Real)(1.0 / sigma);
	Real* dNorm = (Real*)(ctx->dScal + 28);          // two scalars of the stop test
	int iter = 0;
	for (iter = 0; iter < maxIters; iter++) {
		k_pd_stage_a<Real><<<b3, 256, 0, ctx->stream>>>(n3, invSigma, sigma, (Real*)x->d, (Real*)x0->d, (const Real*)y->d, (const Real*)Q->d, (const Real*)invA->d, (Real*)vn->d); MP_CHECK_LAUNCH(ctx);
		MP_TRY(blur<Real>(ctx, d, flags, vn, orig, t1, t2, K)); MP_TRY(blur<Real>(ctx, d, flags, vn, orig, t1, t2, K));
		k_pd_stage_b<Real><<<b3, 256, 0, ctx->stream>>>(n3, sigma, tau, (Real*)x->d, (const Real*)x0->d, (const Real*)y->d, (const Real*)velC->d, (const Real*)invA->d,
			(const Real*)vn->d, (Real*)z->d, (Real*)z0->d); MP_CHECK_LAUNCH(ctx);
		MP_TRY(mp_solve_pressure(ctx, z, pressure, flags, phi, perCellCorr, fractions, obvel, curv, nullptr, &pp, nullptr));      // :328-329
		MP_CUDA(cudaMemsetAsync(dNorm, 0, 2 * sizeof(Real), ctx->stream));
		k_pd_stage_c<Real><<<b1, 256, 0, ctx->stream>>>(n, theta, (const Real*)z->d, (const Real*)z0->d, (Real*)y->d, dNorm); MP_CHECK_LAUNCH(ctx);
		bool stop = false;
		if (iter > 0) {
			Real h[2];
			MP_CUDA(cudaMemcpyAsync(h, dNorm, 2 * sizeof(Real), cudaMemcpyDeviceToHost, ctx->stream));
			MP_CUDA(cudaStreamSynchronize(ctx->stream));
			const Real rnorm = std::sqrt(h[0]), zmax = std::sqrt(h[1]);
			const Real epsDual = (Real)(std::sqrt(d.is3D ? 3.0 : 2.0) * (double)epsAbs + (double)(epsRel * zmax));
			stop = rnorm < epsDual;
		}
		if (stop || (iter == maxIters - 1)) break;
	}
	MP_CUDA(cudaMemcpyAsync(vel->d, z->d, vel->bytes, cudaMemcpyDeviceToDevice, ctx->stream));       // vel.copyFrom(z) :348
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	if (iterations) *iterations = iter;
	return MP_OK;
}

}  // namespace

extern "C" int mp_pd_fluid_guiding(mp_context* ctx, mp_grid* vel, const mp_grid* velT, mp_grid* pressure, const mp_grid* flags, const mp_grid* weight,
	int blurRadius, double theta, double tau, double sigma, double epsRel, double epsAbs, int maxIters,
	const mp_grid* phi, const mp_grid* perCellCorr, const mp_grid* fractions, const mp_grid* obvel, double gfClamp, double cgMaxIterFac, double cgAccuracy,
	int preconditioner, int zeroPressureFixing, const mp_grid* curv, double surfTens, int* iterations)
{
	if (!ctx || !vel || !velT || !pressure || !flags || !weight) MP_FAIL(MP_ERR_INVALID, "mp_pd_fluid_guiding: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_pd_fluid_guiding: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, vel, MP_GRID_MAC, "vel", false)); MP_TRY(mp_check_same(vel, velT, MP_GRID_MAC, "velT", false));
	MP_TRY(mp_check_same(flags, pressure, MP_GRID_REAL, "pressure", false)); MP_TRY(mp_check_same(pressure, weight, MP_GRID_REAL, "weight", false));
	if (blurRadius < 0 || 2 * blurRadius + 1 > PD_MAXK) MP_FAIL(MP_ERR_INVALID, "PD_fluid_guiding: blurRadius must be in [0, %d]", (PD_MAXK - 1) / 2);
	if (ctx->dist && ctx->dist->active) MP_FAIL(MP_ERR_UNSUPPORTED, "mp_pd_fluid_guiding: not available on z-slab sharded grids yet");
	MP_CUDA(cudaSetDevice(ctx->device));
	mp_pressure_params pp;
	pp.cgAccuracy = cgAccuracy; pp.gfClamp = gfClamp; pp.cgMaxIterFac = cgMaxIterFac; pp.precondition = 1; pp.preconditioner = preconditioner;
	pp.enforceCompatibility = 0; pp.useL2Norm = 0; pp.zeroPressureFixing = zeroPressureFixing; pp.surfTens = surfTens;
	if (vel->prec == 4) return guide<float>(ctx, vel, velT, pressure, flags, weight, blurRadius, theta, tau, sigma, epsRel, epsAbs, maxIters, phi, perCellCorr, fractions, obvel, pp, curv, iterations);
	return guide<double>(ctx, vel, velT, pressure, flags, weight, blurRadius, theta, tau, sigma, epsRel, epsAbs, maxIters, phi, perCellCorr, fractions, obvel, pp, curv, iterations);
}

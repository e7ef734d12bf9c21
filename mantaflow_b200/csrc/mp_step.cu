// The steps either side of the pressure projection, so that a whole smoke step keeps its fields in HBM:
//   setWallBcs          plugin/extforces.cpp:186-218, :307-316   (KnSetWallBcs; the phiObs + fractions variant KnSetWallBcsFrac is in mp_liquid.cu)
//   addGravity          plugin/extforces.cpp:45-65               (KnApplyForce, additive)
//   addBuoyancy         plugin/extforces.cpp:75-90               (KnAddBuoyancy)
//   advectSemiLagrange  plugin/advection.cpp:25-58, :81-316, :323-461 (semi-Lagrange + MacCormack for Real and MAC grids,
//                       orderSpace 1 / 2 (cubic lookups, util/interpolHigh.h), orderTrace 1 / 2 (explicit-midpoint back-tracing),
//                       clamp modes 1 and 2, convective outflow boundary)
// One thread per cell, gathers only, every cell written by exactly one thread: any order of execution gives the serial result.
// The per-cell arithmetic keeps the reference's evaluation (double literals promote, results narrow on assignment; -fmad=false),
// so the fields are BIT-IDENTICAL to the reference's in both precisions (tests/test_gpu_step.py).
#include "mp_common.cuh"
#include <cfloat>

namespace {

template <typename Real> struct V3 { Real x, y, z; };

// launch: x over i (128 threads), blockIdx.y = j, blockIdx.z = k -- no 64-bit division per cell
__device__ __forceinline__ bool cellOf(const Dims& d, int& i, int& j, int& k, IndexInt& idx) {
	i = blockIdx.x * blockDim.x + threadIdx.x; j = blockIdx.y; k = blockIdx.z;
	idx = (IndexInt)i + d.Y * j + (IndexInt)d.sx * d.sy * k;
	return i < d.sx;
}
__device__ __forceinline__ bool isInterior(const Dims& d, int i, int j, int k) {   // the cells of a KERNEL(bnd=1)
	return i >= 1 && i <= d.sx - 2 && j >= 1 && j <= d.sy - 2 && (!d.is3D || (k >= 1 && k <= d.sz - 2));
}
static inline dim3 cellGrid(const Dims& d) { return dim3((unsigned)((d.sx + 127) / 128), (unsigned)d.sy, (unsigned)d.sz); }

// ---------------------------------------------------------------- setWallBcs / forces
// The three one-pass plugins share one kernel shell: a thread owns V consecutive cells of a row (V = 4 float / 2 double when sx % V == 0,
// i.e. three 16-byte accesses to the AoS velocity instead of 3V strided ones; V = 1 otherwise), applies the per-cell operation and writes
// the vectors back only if a cell changed.
template <typename Real, int V> struct alignas(sizeof(Real) * V) VecS { Real v[V]; };

template <typename Real> struct WallBcsOp {          // KnSetWallBcs extforces.cpp:186-218
	const int* flags; const Real* obvel;
	__device__ __forceinline__ bool operator()(const Dims& d, int i, int j, int k, IndexInt idx, Real& vx, Real& vy, Real& vz) const {
		const int fl = flags[idx];
		const bool curFluid = fl & TypeFluid, curObs = fl & TypeObstacle;
		if (!curFluid && !curObs) return false;
		Real bx = 0, by = 0, bz = 0;
		if (obvel) { bx = obvel[3 * idx]; by = obvel[3 * idx + 1]; if (d.is3D) bz = obvel[3 * idx + 2]; }
		const Real ox = vx, oy = vy, oz = vz;
		if (i > 0 && (flags[idx - d.X] & TypeObstacle)) vx = bx;
		if (i > 0 && curObs && (flags[idx - d.X] & TypeFluid)) vx = bx;
		if (j > 0 && (flags[idx - d.Y] & TypeObstacle)) vy = by;
		if (j > 0 && curObs && (flags[idx - d.Y] & TypeFluid)) vy = by;
		if (!d.is3D) vz = 0;
		else {
			if (k > 0 && (flags[idx - d.Z] & TypeObstacle)) vz = bz;
			if (k > 0 && curObs && (flags[idx - d.Z] & TypeFluid)) vz = bz;
		}
		if (curFluid) {
			if ((i > 0 && (flags[idx - d.X] & TypeStick)) || (i < d.sx - 1 && (flags[idx + d.X] & TypeStick))) vy = vz = 0;
			if ((j > 0 && (flags[idx - d.Y] & TypeStick)) || (j < d.sy - 1 && (flags[idx + d.Y] & TypeStick))) vx = vz = 0;
			if (d.is3D && ((k > 0 && (flags[idx - d.Z] & TypeStick)) || (k < d.sz - 1 && (flags[idx + d.Z] & TypeStick)))) vx = vy = 0;
		}
		return !(vx == ox && vy == oy && vz == oz);      // (NaN compares unequal: rewritten, harmless)
	}
};
template <typename Real> struct ApplyForceOp {       // KnApplyForce extforces.cpp:45-58, additive
	const int* flags; const Real* exclude; Real fx, fy, fz;
	__device__ __forceinline__ bool operator()(const Dims& d, int i, int j, int k, IndexInt idx, Real& vx, Real& vy, Real& vz) const {
		if (!isInterior(d, i, j, k)) return false;
		const bool curFluid = flags[idx] & TypeFluid, curEmpty = flags[idx] & TypeEmpty;
		if (!curFluid && !curEmpty) return false;
		if (exclude && (exclude[idx] < 0.)) return false;
		if ((flags[idx - d.X] & TypeFluid) || (curFluid && (flags[idx - d.X] & TypeEmpty))) vx = vx + fx;
		if ((flags[idx - d.Y] & TypeFluid) || (curFluid && (flags[idx - d.Y] & TypeEmpty))) vy = vy + fy;
		if (d.is3D && ((flags[idx - d.Z] & TypeFluid) || (curFluid && (flags[idx - d.Z] & TypeEmpty)))) vz = vz + fz;
		return true;
	}
};
template <typename Real> struct BuoyancyOp {         // KnAddBuoyancy extforces.cpp:75-83
	const int* flags; const Real* factor; Real sx_, sy_, sz_;
	__device__ __forceinline__ bool operator()(const Dims& d, int i, int j, int k, IndexInt idx, Real& vx, Real& vy, Real& vz) const {
		if (!isInterior(d, i, j, k)) return false;
		if (!(flags[idx] & TypeFluid)) return false;
		const Real f0 = factor[idx];
		if (flags[idx - d.X] & TypeFluid) vx = (Real)((double)vx + (0.5 * (double)sx_) * (double)(f0 + factor[idx - d.X]));
		if (flags[idx - d.Y] & TypeFluid) vy = (Real)((double)vy + (0.5 * (double)sy_) * (double)(f0 + factor[idx - d.Y]));
		if (d.is3D && (flags[idx - d.Z] & TypeFluid)) vz = (Real)((double)vz + (0.5 * (double)sz_) * (double)(f0 + factor[idx - d.Z]));
		return true;
	}
};
template <typename Real, int V, typename Op>
__global__ void __launch_bounds__(128) k_cells(Dims d, Real* vel, Op op) {
	const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * V, j = blockIdx.y, k = blockIdx.z;
	if (i0 >= d.sx) return;
	const IndexInt idx0 = (IndexInt)i0 + d.Y * j + (IndexInt)d.sx * d.sy * k;
	Real* p = vel + 3 * idx0;
	Real v[3 * V];
	if (V == 1) { v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; }
	else {
		#pragma unroll
		for (int q = 0; q < 3; q++) { const VecS<Real, V> t = ((const VecS<Real, V>*)p)[q];
			#pragma unroll
			for (int e = 0; e < V; e++) v[q * V + e] = t.v[e]; }
	}
	bool changed = false;
	#pragma unroll
	for (int c = 0; c < V; c++) changed |= op(d, i0 + c, j, k, idx0 + c, v[3 * c], v[3 * c + 1], v[3 * c + 2]);
	if (!changed) return;
	if (V == 1) { p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; }
	else {
		#pragma unroll
		for (int q = 0; q < 3; q++) { VecS<Real, V> t;
			#pragma unroll
			for (int e = 0; e < V; e++) t.v[e] = v[q * V + e];
			((VecS<Real, V>*)p)[q] = t; }
	}
}
template <typename Real, typename Op>
int launchCells(mp_context* ctx, const Dims& d, Real* vel, const Op& op) {
	constexpr int VV = 16 / (int)sizeof(Real);
	if (d.sx % VV == 0) k_cells<Real, VV, Op><<<dim3((unsigned)((d.sx / VV + 127) / 128), (unsigned)d.sy, (unsigned)d.sz), 128, 0, ctx->stream>>>(d, vel, op);
	else                k_cells<Real, 1, Op><<<cellGrid(d), 128, 0, ctx->stream>>>(d, vel, op);
	MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}
// ---------------------------------------------------------------- interpolation (util/interpol.h:50-91) and MAC accessors (grid.h:424-466)
// STRIDE 1: Grid<Real>; STRIDE 3: one component of a Vec3 grid (a points at that component of cell 0)
template <typename Real, int STRIDE>
__device__ __forceinline__ Real interpol(const Dims& d, const Real* __restrict__ a, Real posx, Real posy, Real posz) {
	const Real px = posx - 0.5f, py = posy - 0.5f, pz = posz - 0.5f;
	int xi = (int)px, yi = (int)py, zi = (int)pz;
	Real s1 = px - (Real)xi, s0 = (Real)(1. - s1);
	Real t1 = py - (Real)yi, t0 = (Real)(1. - t1);
	Real f1 = pz - (Real)zi, f0 = (Real)(1. - f1);
	if (px < 0.) { xi = 0; s0 = 1.0; s1 = 0.0; }
	if (py < 0.) { yi = 0; t0 = 1.0; t1 = 0.0; }
	if (pz < 0.) { zi = 0; f0 = 1.0; f1 = 0.0; }
	if (xi >= d.sx - 1) { xi = d.sx - 2; s0 = 0.0; s1 = 1.0; }
	if (yi >= d.sy - 1) { yi = d.sy - 2; t0 = 0.0; t1 = 1.0; }
	if (d.sz > 1) { if (zi >= d.sz - 1) { zi = d.sz - 2; f0 = 0.0; f1 = 1.0; } }
	const IndexInt X = STRIDE, Y = d.Y * STRIDE, Z = d.Z * STRIDE;
	const Real* p = a + ((IndexInt)xi + d.Y * yi + d.Z * zi) * STRIDE;
	return ((p[0] * t0 + p[Y] * t1) * s0 + (p[X] * t0 + p[X + Y] * t1) * s1) * f0
	     + ((p[Z] * t0 + p[Y + Z] * t1) * s0 + (p[X + Z] * t0 + p[X + Y + Z] * t1) * s1) * f1;
}
template <typename Real>
__device__ __forceinline__ V3<Real> macCentered(const Dims& d, const Real* __restrict__ v, IndexInt idx) {
	V3<Real> o;
	o.x = (Real)(0.5 * (double)(v[3 * idx] + v[3 * (idx + 1)]));
	o.y = (Real)(0.5 * (double)(v[3 * idx + 1] + v[3 * (idx + d.Y) + 1]));
	o.z = 0;
	if (d.is3D) o.z = (Real)(0.5 * (double)(v[3 * idx + 2] + v[3 * (idx + d.Z) + 2]));
	return o;
}
#define VC(o, c) v[3 * (idx + (o)) + (c)]
template <typename Real, int C>
__device__ __forceinline__ V3<Real> macAt(const Dims& d, const Real* __restrict__ v, IndexInt idx) {   // getAtMACX / Y / Z
	const IndexInt Y = d.Y, Z = d.Z;
	V3<Real> o;
	if (C == 0) {
		o.x = VC(0, 0);
		o.y = (Real)(0.25 * (double)(VC(0, 1) + VC(-1, 1) + VC(Y, 1) + VC(Y - 1, 1)));
		o.z = 0;
		if (d.is3D) o.z = (Real)(0.25 * (double)(VC(0, 2) + VC(-1, 2) + VC(Z, 2) + VC(Z - 1, 2)));
	} else if (C == 1) {
		o.x = (Real)(0.25 * (double)(VC(0, 0) + VC(-Y, 0) + VC(1, 0) + VC(1 - Y, 0)));
		o.y = VC(0, 1);
		o.z = 0;
		if (d.is3D) o.z = (Real)(0.25 * (double)(VC(0, 2) + VC(-Y, 2) + VC(Z, 2) + VC(Z - Y, 2)));
	} else {
		o.x = (Real)(0.25 * (double)(VC(0, 0) + VC(-Z, 0) + VC(1, 0) + VC(1 - Z, 0)));
		o.y = (Real)(0.25 * (double)(VC(0, 1) + VC(-Z, 1) + VC(Y, 1) + VC(Y - Z, 1)));
		o.z = VC(0, 2);
	}
	return o;
}
#undef VC

// ---------------------------------------------------------------- higher-order lookups: orderSpace 2 (util/interpolHigh.h), orderTrace 2 (interpolMAC)
// cubicInterp util/interpolHigh.h:22-39.  The reference instantiates it for Real and for Vec3: the Vec3 one narrows after every scalar * vector
// product (vectorbase.h:272-279), the Real one forms the polynomial coefficients in double and narrows once.  In double both are the same arithmetic.
template <typename Real, bool VEC>
__device__ __forceinline__ Real cubicInterp(Real t, Real p0, Real p1, Real p2, Real p3) {
	const Real d0 = (Real)((double)(p2 - p0) * 0.5), d1 = (Real)((double)(p3 - p1) * 0.5), dk = p2 - p1;
	Real a2, a3;
	if (VEC) { a2 = ((Real)(3.0 * (double)dk) - (Real)(2.0 * (double)d0)) - d1; a3 = ((Real)(-2.0 * (double)dk) + d0) + d1; }
	else     { a2 = (Real)((3.0 * (double)dk - 2.0 * (double)d0) - (double)d1); a3 = (Real)((-2.0 * (double)dk + (double)d0) + (double)d1); }
	const Real sq = t * t, cu = sq * t;
	return ((a3 * cu + a2 * sq) + d0 * t) + p1;
}
// interpolCubic / interpolCubic2D util/interpolHigh.h:42-172: 4 x 4 (x 4) support, x first, then y, then z; within one cell of the border the
// reference falls back to the linear lookup.  STRIDE 3 = one component of a Vec3 grid (the Vec3 instantiation of cubicInterp).
template <typename Real, int STRIDE>
__device__ Real interpolCubic(const Dims& d, const Real* __restrict__ a, Real posx, Real posy, Real posz) {
	const Real px = posx - 0.5f, py = posy - 0.5f, pz = posz - 0.5f;
	const int x1 = (int)px, y1 = (int)py, z1 = (int)pz;
	const bool out = x1 < 1 || y1 < 1 || x1 + 2 >= d.sx || y1 + 2 >= d.sy || (d.is3D && (z1 < 1 || z1 + 2 >= d.sz));
	if (out) return interpol<Real, STRIDE>(d, a, posx, posy, posz);
	const Real tx = px - (Real)x1, ty = py - (Real)y1;
	constexpr bool VEC = STRIDE == 3;
	const int nz = d.is3D ? 4 : 1;
	Real zp[4];
	for (int c = 0; c < nz; c++) {
		const Real* pl = a + ((IndexInt)(x1 - 1) + d.Y * (y1 - 1) + (d.is3D ? d.Z * (z1 - 1 + c) : (IndexInt)0)) * STRIDE;
		Real yp[4];
		#pragma unroll
		for (int b = 0; b < 4; b++) {
			const Real* r = pl + d.Y * b * STRIDE;
			yp[b] = cubicInterp<Real, VEC>(tx, r[0], r[STRIDE], r[2 * STRIDE], r[3 * STRIDE]);
		}
		zp[c] = cubicInterp<Real, VEC>(ty, yp[0], yp[1], yp[2], yp[3]);
	}
	if (!d.is3D) return zp[0];
	return cubicInterp<Real, VEC>(pz - (Real)z1, zp[0], zp[1], zp[2], zp[3]);
}
// one axis of BUILD_INDEX / BUILD_INDEX_SHIFT util/interpol.h:50-66, :112-125
template <typename Real>
__device__ __forceinline__ void axisWeights(Real p, int size, bool clampHigh, int& i, Real& w0, Real& w1) {
	i = (int)p; w1 = p - (Real)i; w0 = (Real)(1. - w1);
	if (p < 0.) { i = 0; w0 = 1; w1 = 0; }
	if (clampHigh && i >= size - 1) { i = size - 2; w0 = 0; w1 = 1; }
}
// MACGrid::getInterpolated = interpolMAC util/interpol.h:127-157: each component looked up on its own faces
template <typename Real>
__device__ V3<Real> interpolMACAt(const Dims& d, const Real* __restrict__ v, Real posx, Real posy, Real posz) {
	int xi, yi, zi, sxi, syi, szi; Real s0, s1, t0, t1, f0, f1, ss0, ss1, st0, st1, sf0, sf1;
	axisWeights<Real>(posx - 0.5f, d.sx, true, xi, s0, s1); axisWeights<Real>(posy - 0.5f, d.sy, true, yi, t0, t1); axisWeights<Real>(posz - 0.5f, d.sz, d.sz > 1, zi, f0, f1);
	axisWeights<Real>(posx, d.sx, true, sxi, ss0, ss1); axisWeights<Real>(posy, d.sy, true, syi, st0, st1); axisWeights<Real>(posz, d.sz, d.sz > 1, szi, sf0, sf1);
	const IndexInt X = 3, Y = 3 * d.Y, Z = 3 * d.Z;
	V3<Real> o;
	{ const Real* p = v + 3 * (((IndexInt)zi * d.sy + yi) * d.sx + sxi);
	  o.x = f0 * ((p[0] * t0 + p[Y] * t1) * ss0 + (p[X] * t0 + p[X + Y] * t1) * ss1) + f1 * ((p[Z] * t0 + p[Z + Y] * t1) * ss0 + (p[X + Z] * t0 + p[X + Y + Z] * t1) * ss1); }
	{ const Real* p = v + 3 * (((IndexInt)zi * d.sy + syi) * d.sx + xi) + 1;
	  o.y = f0 * ((p[0] * st0 + p[Y] * st1) * s0 + (p[X] * st0 + p[X + Y] * st1) * s1) + f1 * ((p[Z] * st0 + p[Z + Y] * st1) * s0 + (p[X + Z] * st0 + p[X + Y + Z] * st1) * s1); }
	{ const Real* p = v + 3 * (((IndexInt)szi * d.sy + yi) * d.sx + xi) + 2;
	  o.z = sf0 * ((p[0] * t0 + p[Y] * t1) * s0 + (p[X] * t0 + p[X + Y] * t1) * s1) + sf1 * ((p[Z] * t0 + p[Z + Y] * t1) * s0 + (p[X + Z] * t0 + p[X + Y + Z] * t1) * s1); }
	return o;
}
// Grid<Real>::getInterpolatedHi (grid.h:146-151) and MACGrid::getInterpolatedComponentHi<C> (grid.h:268-273; the cubic one is
// interpolCubicMAC(pos)[C], util/interpolHigh.h:174-181: the lookup position moved by half a cell along C, 0 for z in 2-D)
template <typename Real, int OS>
__device__ __forceinline__ Real lookupReal(const Dims& d, const Real* __restrict__ src, Real x, Real y, Real z) {
	return OS == 1 ? interpol<Real, 1>(d, src, x, y, z) : interpolCubic<Real, 1>(d, src, x, y, z);
}
template <typename Real, int OS, int C>
__device__ __forceinline__ Real lookupMAC(const Dims& d, const Real* __restrict__ src, Real x, Real y, Real z) {
	if (OS == 1) return interpol<Real, 3>(d, src + C, x, y, z);
	if (C == 2 && !d.is3D) return (Real)0;
	return interpolCubic<Real, 3>(d, src + C, C == 0 ? x + (Real)0.5 : x + (Real)0, C == 1 ? y + (Real)0.5 : y + (Real)0, C == 2 ? z + (Real)0.5 : z + (Real)0);
}

// ---------------------------------------------------------------- SemiLagrange / SemiLagrangeMAC (advection.cpp:25-77)
// Pass 1 of an advection: dst = src traced back along vel in the interior, 0 on the outer layer (the reference's fresh grid).
// OT 2: explicit midpoint (:32-37); the MAC kernel takes its velocities from SRC there (:58-73), as the reference does.
template <typename Real, int OS, int OT>
__device__ __forceinline__ Real slReal(const Dims& d, const Real* __restrict__ vel, const Real* __restrict__ src, Real dt, int i, int j, int k, IndexInt idx) {
	const V3<Real> c = macCentered<Real>(d, vel, idx);
	if (OT == 1) return lookupReal<Real, OS>(d, src, (i + 0.5f) - c.x * dt, (j + 0.5f) - c.y * dt, (k + 0.5f) - c.z * dt);
	const Real p0x = i + 0.5f, p0y = j + 0.5f, p0z = k + 0.5f;
	const V3<Real> u = interpolMACAt<Real>(d, vel, p0x - (Real)((double)(c.x * dt) * 0.5), p0y - (Real)((double)(c.y * dt) * 0.5), p0z - (Real)((double)(c.z * dt) * 0.5));
	return lookupReal<Real, OS>(d, src, p0x - u.x * dt, p0y - u.y * dt, p0z - u.z * dt);
}
template <typename Real, int OS, int OT, int C>
__device__ __forceinline__ Real slMACc(const Dims& d, const Real* __restrict__ vel, const Real* __restrict__ src, Real dt, int i, int j, int k, IndexInt idx) {
	if (OT == 1) {
		const V3<Real> m = macAt<Real, C>(d, vel, idx);
		return lookupMAC<Real, OS, C>(d, src, (i + 0.5f) - m.x * dt, (j + 0.5f) - m.y * dt, (k + 0.5f) - m.z * dt);
	}
	const V3<Real> m = macAt<Real, C>(d, src, idx);
	const Real p0x = i + 0.5f, p0y = j + 0.5f, p0z = k + 0.5f;
	const Real fx = C == 0 ? (Real)i : p0x, fy = C == 1 ? (Real)j : p0y, fz = C == 2 ? (Real)k : p0z;      // the face centre
	const V3<Real> u = interpolMACAt<Real>(d, src, fx - (Real)((double)(m.x * dt) * 0.5), fy - (Real)((double)(m.y * dt) * 0.5), fz - (Real)((double)(m.z * dt) * 0.5));
	return lookupMAC<Real, OS, C>(d, src, p0x - u.x * dt, p0y - u.y * dt, p0z - u.z * dt);
}
template <typename Real, int OS, int OT>
__device__ __forceinline__ V3<Real> slMAC(const Dims& d, const Real* __restrict__ vel, const Real* __restrict__ src, Real dt, int i, int j, int k, IndexInt idx) {
	V3<Real> o;
	o.x = slMACc<Real, OS, OT, 0>(d, vel, src, dt, i, j, k, idx);
	o.y = slMACc<Real, OS, OT, 1>(d, vel, src, dt, i, j, k, idx);
	o.z = slMACc<Real, OS, OT, 2>(d, vel, src, dt, i, j, k, idx);
	return o;
}
template <typename Real, int OS, int OT, bool D3>
__global__ void __launch_bounds__(128) k_semi_lagrange(Dims d_, const Real* __restrict__ vel, Real* __restrict__ dst, const Real* __restrict__ src, Real dt) {
	Dims d = d_; if (D3) d.is3D = true;          // 3-D instantiation: the dimension tests fold, the corner loops unroll
	int i, j, k; IndexInt idx;
	if (!cellOf(d, i, j, k, idx)) return;
	dst[idx] = isInterior(d, i, j, k) ? slReal<Real, OS, OT>(d, vel, src, dt, i, j, k, idx) : (Real)0;
}
template <typename Real, int OS, int OT, bool D3>
__global__ void __launch_bounds__(128) k_semi_lagrange_mac(Dims d_, const Real* __restrict__ vel, Real* __restrict__ dst, const Real* __restrict__ src, Real dt) {
	Dims d = d_; if (D3) d.is3D = true;          // 3-D instantiation: the dimension tests fold, the corner loops unroll
	int i, j, k; IndexInt idx;
	if (!cellOf(d, i, j, k, idx)) return;
	V3<Real> o; o.x = o.y = o.z = 0;
	if (isInterior(d, i, j, k)) o = slMAC<Real, OS, OT>(d, vel, src, dt, i, j, k, idx);
	dst[3 * idx] = o.x; dst[3 * idx + 1] = o.y; dst[3 * idx + 2] = o.z;
}

// SemiLagrange<Vec3> (a cell-centred Grid<Vec3>): the position of slReal, one lookup per component (interpol<Vec3> / interpolCubic<Vec3> share the weights)
template <typename Real, int OS, int OT>
__device__ __forceinline__ V3<Real> slVec3(const Dims& d, const Real* __restrict__ vel, const Real* __restrict__ src, Real dt, int i, int j, int k, IndexInt idx) {
	const V3<Real> c = macCentered<Real>(d, vel, idx);
	Real px, py, pz;
	if (OT == 1) { px = (i + 0.5f) - c.x * dt; py = (j + 0.5f) - c.y * dt; pz = (k + 0.5f) - c.z * dt; }
	else {
		const Real p0x = i + 0.5f, p0y = j + 0.5f, p0z = k + 0.5f;
		const V3<Real> u = interpolMACAt<Real>(d, vel, p0x - (Real)((double)(c.x * dt) * 0.5), p0y - (Real)((double)(c.y * dt) * 0.5), p0z - (Real)((double)(c.z * dt) * 0.5));
		px = p0x - u.x * dt; py = p0y - u.y * dt; pz = p0z - u.z * dt;
	}
	V3<Real> o;
	if (OS == 1) { o.x = interpol<Real, 3>(d, src, px, py, pz); o.y = interpol<Real, 3>(d, src + 1, px, py, pz); o.z = interpol<Real, 3>(d, src + 2, px, py, pz); }
	else { o.x = interpolCubic<Real, 3>(d, src, px, py, pz); o.y = interpolCubic<Real, 3>(d, src + 1, px, py, pz); o.z = interpolCubic<Real, 3>(d, src + 2, px, py, pz); }
	return o;
}
template <typename Real, int OS, int OT, bool D3>
__global__ void __launch_bounds__(128) k_semi_lagrange_vec3(Dims d_, const Real* __restrict__ vel, Real* __restrict__ dst, const Real* __restrict__ src, Real dt) {
	Dims d = d_; if (D3) d.is3D = true;          // 3-D instantiation: the dimension tests fold, the corner loops unroll
	int i, j, k; IndexInt idx;
	if (!cellOf(d, i, j, k, idx)) return;
	V3<Real> o; o.x = o.y = o.z = 0;
	if (isInterior(d, i, j, k)) o = slVec3<Real, OS, OT>(d, vel, src, dt, i, j, k, idx);
	dst[3 * idx] = o.x; dst[3 * idx + 1] = o.y; dst[3 * idx + 2] = o.z;
}

// ---------------------------------------------------------------- MacCormack (advection.cpp:81-117 correct, :141-287 clamp)
template <typename Real> __device__ __forceinline__ Real realMax();
template <> __device__ __forceinline__ float realMax<float>() { return FLT_MAX; }
template <> __device__ __forceinline__ double realMax<double>() { return DBL_MAX; }
__device__ __forceinline__ int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ bool checkFlag(const int* flags, IndexInt q) { return flags[q] & (TypeFluid | TypeEmpty); }

// Pass 2 of a MacCormack advection of a Real grid: the reference's backward trace (SemiLagrange with -dt on fwd), MacCormackCorrect and
// MacCormackClamp only ever combine values of ONE cell (plus read-only neighbourhoods of fwd / orig), so they are one kernel here:
// three full-grid round trips of bwd and the corrected grid never reach HBM.
template <typename Real, int OS, int OT, bool D3>
__global__ void __launch_bounds__(128) k_mc_rest(Dims d_, const int* __restrict__ flags, const Real* __restrict__ vel, Real* __restrict__ dst, const Real* __restrict__ orig,
	const Real* __restrict__ fwd, Real dt, Real strength, int clampMode) {
	Dims d = d_; if (D3) d.is3D = true;          // 3-D instantiation: the dimension tests fold, the corner loops unroll
	int i, j, k; IndexInt idx;
	if (!cellOf(d, i, j, k, idx)) return;
	const bool in = isInterior(d, i, j, k);
	const Real f = fwd[idx];
	const Real bwd = in ? slReal<Real, OS, OT>(d, vel, fwd, -dt, i, j, k, idx) : (Real)0;
	Real dval = f;                                                   // MacCormackCorrect :81-91 (all cells)
	if (flags[idx] & TypeFluid) dval = (Real)((double)dval + ((double)strength * 0.5) * (double)(orig[idx] - bwd));
	if (in) {                                                        // MacCormackClamp :241-267
		const V3<Real> c = macCentered<Real>(d, vel, idx);
		const Real v[3] = { c.x * dt, c.y * dt, c.z * dt }, pos[3] = { (Real)i, (Real)j, (Real)k };
		{	// doClampComponent :141-186
			Real minv = realMax<Real>(), maxv = -realMax<Real>();
			bool haveFl = false;
			const int numPos = clampMode == 1 ? 2 : 1;
			for (int l = 0; l < numPos; l++) {
				int cp[3];
				#pragma unroll
				for (int a = 0; a < 3; a++) cp[a] = (int)(l == 0 ? pos[a] - v[a] : pos[a] + v[a]);
				const int i0 = iclamp(cp[0], 0, d.sx - 2), j0 = iclamp(cp[1], 0, d.sy - 2), k0 = iclamp(cp[2], 0, d.is3D ? d.sz - 2 : 1);
				const int k1 = d.is3D ? k0 + 1 : k0;
				#pragma unroll
				for (int cc = 0; cc < (d.is3D ? 2 : 1); cc++)
				#pragma unroll
				for (int b = 0; b < 2; b++)
				#pragma unroll
				for (int a = 0; a < 2; a++) {
					const IndexInt q = (IndexInt)(i0 + a) + d.Y * (j0 + b) + d.Z * (cc ? k1 : k0);
					if (checkFlag(flags, q)) { const Real o = orig[q]; if (o < minv) minv = o; if (o > maxv) maxv = o; haveFl = true; }
				}
			}
			if (!haveFl) dval = f;
			else if (clampMode == 1) dval = dval < minv ? minv : (dval > maxv ? maxv : dval);
			else if (dval < minv || dval > maxv) dval = f;
		}
		if (clampMode == 1) {   // lookups that leave the grid or end in an obstacle fall back to first order (:252-264)
			int pf[3], pb[3];
			#pragma unroll
			for (int a = 0; a < 3; a++) { pf[a] = (int)((pos[a] + (Real)0.5) - v[a]); pb[a] = (int)((pos[a] + (Real)0.5) + v[a]); }
			const int ux = d.sx - 1, uy = d.sy - 1, uz = d.sz - 1;
			bool bad = pf[0] < 0 || pf[1] < 0 || pf[2] < 0 || pb[0] < 0 || pb[1] < 0 || pb[2] < 0 ||
			           pf[0] > ux || pf[1] > uy || ((pf[2] > uz) && d.is3D) || pb[0] > ux || pb[1] > uy || ((pb[2] > uz) && d.is3D);
			if (!bad) bad = (flags[(IndexInt)pf[0] + d.Y * pf[1] + d.Z * pf[2]] & TypeObstacle) || (flags[(IndexInt)pb[0] + d.Y * pb[1] + d.Z * pb[2]] & TypeObstacle);
			if (bad) dval = f;
		}
	}
	dst[idx] = dval;
}

// Pass 2 for a cell-centred Grid<Vec3>: MacCormackCorrect<Vec3> (the correction is narrowed before it is added: double * Vec3, then Vec3 += Vec3) and
// doClampComponent<Vec3> (:141-186): minimum / maximum per component over the flagged corners; clampMode 2 resets the WHOLE vector when one
// component leaves its range (cmpMinMax<Vec3> :134-136), clampMode 1 clamps component by component (clamp<Vec3> vectorbase.h:605-609)
template <typename Real, int OS, int OT, bool D3>
__global__ void __launch_bounds__(128) k_mc_rest_vec3(Dims d_, const int* __restrict__ flags, const Real* __restrict__ vel, Real* __restrict__ dst, const Real* __restrict__ orig,
	const Real* __restrict__ fwd, Real dt, Real strength, int clampMode) {
	Dims d = d_; if (D3) d.is3D = true;          // 3-D instantiation: the dimension tests fold, the corner loops unroll
	int i, j, k; IndexInt idx;
	if (!cellOf(d, i, j, k, idx)) return;
	const bool in = isInterior(d, i, j, k);
	const Real f[3] = { fwd[3 * idx], fwd[3 * idx + 1], fwd[3 * idx + 2] };
	V3<Real> bw; bw.x = bw.y = bw.z = 0;
	if (in) bw = slVec3<Real, OS, OT>(d, vel, fwd, -dt, i, j, k, idx);
	const Real b[3] = { bw.x, bw.y, bw.z };
	Real dv[3] = { f[0], f[1], f[2] };
	if (flags[idx] & TypeFluid) {
		#pragma unroll
		for (int c = 0; c < 3; c++) dv[c] = dv[c] + (Real)(((double)strength * 0.5) * (double)(orig[3 * idx + c] - b[c]));
	}
	if (in) {
		const V3<Real> cv = macCentered<Real>(d, vel, idx);
		const Real v[3] = { cv.x * dt, cv.y * dt, cv.z * dt }, pos[3] = { (Real)i, (Real)j, (Real)k };
		Real minv[3] = { realMax<Real>(), realMax<Real>(), realMax<Real>() }, maxv[3] = { -realMax<Real>(), -realMax<Real>(), -realMax<Real>() };
		bool haveFl = false;
		const int numPos = clampMode == 1 ? 2 : 1;
		for (int l = 0; l < numPos; l++) {
			int cp[3];
			#pragma unroll
			for (int a = 0; a < 3; a++) cp[a] = (int)(l == 0 ? pos[a] - v[a] : pos[a] + v[a]);
			const int i0 = iclamp(cp[0], 0, d.sx - 2), j0 = iclamp(cp[1], 0, d.sy - 2), k0 = iclamp(cp[2], 0, d.is3D ? d.sz - 2 : 1);
			const int k1 = d.is3D ? k0 + 1 : k0;
			#pragma unroll
			for (int cc = 0; cc < (d.is3D ? 2 : 1); cc++)
			#pragma unroll
			for (int bb = 0; bb < 2; bb++)
			#pragma unroll
			for (int a = 0; a < 2; a++) {
				const IndexInt q = (IndexInt)(i0 + a) + d.Y * (j0 + bb) + d.Z * (cc ? k1 : k0);
				if (checkFlag(flags, q)) {
					#pragma unroll
					for (int c = 0; c < 3; c++) { const Real o = orig[3 * q + c]; if (o < minv[c]) minv[c] = o; if (o > maxv[c]) maxv[c] = o; }
					haveFl = true;
				}
			}
		}
		if (!haveFl) { dv[0] = f[0]; dv[1] = f[1]; dv[2] = f[2]; }
		else if (clampMode == 1) {
			#pragma unroll
			for (int c = 0; c < 3; c++) dv[c] = dv[c] < minv[c] ? minv[c] : (dv[c] > maxv[c] ? maxv[c] : dv[c]);
		} else if (dv[0] < minv[0] || dv[0] > maxv[0] || dv[1] < minv[1] || dv[1] > maxv[1] || dv[2] < minv[2] || dv[2] > maxv[2]) { dv[0] = f[0]; dv[1] = f[1]; dv[2] = f[2]; }
		if (clampMode == 1) {
			int pf[3], pb[3];
			#pragma unroll
			for (int a = 0; a < 3; a++) { pf[a] = (int)((pos[a] + (Real)0.5) - v[a]); pb[a] = (int)((pos[a] + (Real)0.5) + v[a]); }
			const int ux = d.sx - 1, uy = d.sy - 1, uz = d.sz - 1;
			bool bad = pf[0] < 0 || pf[1] < 0 || pf[2] < 0 || pb[0] < 0 || pb[1] < 0 || pb[2] < 0 ||
			           pf[0] > ux || pf[1] > uy || ((pf[2] > uz) && d.is3D) || pb[0] > ux || pb[1] > uy || ((pb[2] > uz) && d.is3D);
			if (!bad) bad = (flags[(IndexInt)pf[0] + d.Y * pf[1] + d.Z * pf[2]] & TypeObstacle) || (flags[(IndexInt)pb[0] + d.Y * pb[1] + d.Z * pb[2]] & TypeObstacle);
			if (bad) { dv[0] = f[0]; dv[1] = f[1]; dv[2] = f[2]; }
		}
	}
	dst[3 * idx] = dv[0]; dst[3 * idx + 1] = dv[1]; dst[3 * idx + 2] = dv[2];
}

template <typename Real, int C>     // doClampComponentMAC<c> :191-235
__device__ __forceinline__ Real clampComponentMAC(const Dims& d, const int* __restrict__ flags, Real dst, const Real* __restrict__ orig, Real fwd,
	int i, int j, int k, IndexInt idx, V3<Real> vel, int clampMode) {
	Real minv = realMax<Real>(), maxv = -realMax<Real>();
	const Real pos[3] = { (Real)i, (Real)j, (Real)k }, v[3] = { vel.x, vel.y, vel.z };
	if (clampMode == 2) {
		const IndexInt nb = idx - (C == 0 ? d.X : (C == 1 ? d.Y : d.Z));
		if (!(checkFlag(flags, idx) && checkFlag(flags, nb))) return fwd;
	}
	const int numPos = clampMode == 1 ? 2 : 1;
	for (int l = 0; l < numPos; l++) {
		int cp[3];
		#pragma unroll
		for (int a = 0; a < 3; a++) cp[a] = (int)(l == 0 ? pos[a] - v[a] : pos[a] + v[a]);
		const int i0 = iclamp(cp[0], 0, d.sx - 2), j0 = iclamp(cp[1], 0, d.sy - 2), k0 = iclamp(cp[2], 0, d.is3D ? d.sz - 2 : 0);
		const int k1 = d.is3D ? k0 + 1 : k0;
		const Real* p = orig + 3 * ((IndexInt)i0 + d.Y * j0 + d.Z * k0) + C;       // the corner (i0, j0, k0); the others at fixed offsets from it
		const IndexInt oY = 3 * d.Y, oZ = 3 * d.Z * (k1 - k0);
		#pragma unroll
		for (int cc = 0; cc < (d.is3D ? 2 : 1); cc++)
		#pragma unroll
		for (int b = 0; b < 2; b++)
		#pragma unroll
		for (int a = 0; a < 2; a++) {
			const Real o = p[3 * a + (b ? oY : 0) + (cc ? oZ : 0)];
			if (o < minv) minv = o;
			if (o > maxv) maxv = o;
		}
	}
	if (clampMode == 1) return dst < minv ? minv : (dst > maxv ? maxv : dst);
	if (dst < minv || dst > maxv) dst = fwd;
	return dst;
}
// Pass 2 for a MAC grid: backward trace + MacCormackCorrectMAC (:94-117, all cells) + MacCormackClampMAC (:270-287, interior);
// *anyOutflow is raised when the flags hold an outflow cell, so that the boundary pass can be skipped otherwise
template <typename Real, int OS, int OT, bool D3>
__global__ void __launch_bounds__(128) k_mc_rest_mac(Dims d_, const int* __restrict__ flags, const Real* __restrict__ vel, Real* __restrict__ dst, const Real* __restrict__ orig,
	const Real* __restrict__ fwd, Real dt, Real strength, int clampMode, int* anyOutflow) {
	Dims d = d_; if (D3) d.is3D = true;          // 3-D instantiation: the dimension tests fold, the corner loops unroll
	int i, j, k; IndexInt idx;
	if (!cellOf(d, i, j, k, idx)) return;
	const bool in = isInterior(d, i, j, k);
	const int fl = flags[idx];
	if (fl & TypeOutflow) *anyOutflow = 1;
	const Real f[3] = { fwd[3 * idx], fwd[3 * idx + 1], fwd[3 * idx + 2] };
	V3<Real> bw; bw.x = bw.y = bw.z = 0;
	if (in) bw = slMAC<Real, OS, OT>(d, vel, fwd, -dt, i, j, k, idx);
	const Real b[3] = { bw.x, bw.y, bw.z };
	bool skip[3] = { false, false, false };
	if (!(fl & TypeFluid)) skip[0] = skip[1] = skip[2] = true;
	if (i > 0 && !(flags[idx - d.X] & TypeFluid)) skip[0] = true;
	if (j > 0 && !(flags[idx - d.Y] & TypeFluid)) skip[1] = true;
	if (k > 0 && !(flags[idx - d.Z] & TypeFluid)) skip[2] = true;
	Real o[3];
	#pragma unroll
	for (int c = 0; c < 3; c++) o[c] = skip[c] ? f[c] : (Real)((double)f[c] + ((double)strength * 0.5) * (double)(orig[3 * idx + c] - b[c]));
	if (in) {
		V3<Real> m = macAt<Real, 0>(d, vel, idx); m.x = m.x * dt; m.y = m.y * dt; m.z = m.z * dt;
		o[0] = clampComponentMAC<Real, 0>(d, flags, o[0], orig, f[0], i, j, k, idx, m, clampMode);
		m = macAt<Real, 1>(d, vel, idx); m.x = m.x * dt; m.y = m.y * dt; m.z = m.z * dt;
		o[1] = clampComponentMAC<Real, 1>(d, flags, o[1], orig, f[1], i, j, k, idx, m, clampMode);
		if (d.is3D) {
			m = macAt<Real, 2>(d, vel, idx); m.x = m.x * dt; m.y = m.y * dt; m.z = m.z * dt;
			o[2] = clampComponentMAC<Real, 2>(d, flags, o[2], orig, f[2], i, j, k, idx, m, clampMode);
		}
	}
	dst[3 * idx] = o[0]; dst[3 * idx + 1] = o[1]; dst[3 * idx + 2] = o[2];
}

// ---------------------------------------------------------------- convective outflow boundary (advection.cpp:323-392)
// writes the extrapolated velocity of the outflow cells into velDst (zero elsewhere); k_outflow_copy then moves it into vel
template <typename Real>
__global__ void __launch_bounds__(128) k_outflow_extrapolate(Dims d, const int* __restrict__ flags, const Real* __restrict__ vel, Real* __restrict__ velDst,
	const Real* __restrict__ velPrev, Real timeStep, const int* __restrict__ anyOutflow) {
	if (anyOutflow && !*anyOutflow) return;
	int i, j, k; IndexInt idx;
	if (!cellOf(d, i, j, k, idx) || !(flags[idx] & TypeOutflow)) return;
	Real avg[3] = { 0, 0, 0 }; int count = 0;
	const int nmax = d.is3D ? 1 : 0;
	for (int nn = -nmax; nn <= nmax; nn++) for (int m = -1; m <= 1; m++) for (int l = -1; l <= 1; l++) {
		const int a = i + l, b = j + m, c = k + nn;
		if (a < 0 || b < 0 || c < 0 || a >= d.sx || b >= d.sy || c >= d.sz) continue;
		const IndexInt q = (IndexInt)a + d.Y * b + d.Z * c;
		if (flags[q] & (TypeFluid | TypeOutflow)) { avg[0] += vel[3 * q]; avg[1] += vel[3 * q + 1]; avg[2] += vel[3 * q + 2]; count++; }
	}
	if (count > 0) { avg[0] = avg[0] / (Real)count; avg[1] = avg[1] / (Real)count; avg[2] = avg[2] / (Real)count; }
	const int cur[3] = { i, j, k }, size[3] = { d.sx, d.sy, d.sz };
	const IndexInt stride[3] = { d.X, d.Y, d.Z };
	const Real v0 = vel[3 * idx], v1 = vel[3 * idx + 1], v2 = vel[3 * idx + 2];
	const Real p0 = velPrev[3 * idx], p1 = velPrev[3 * idx + 1], p2 = velPrev[3 * idx + 2];
	Real o0 = 0, o1 = 0, o2 = 0;
	int cnt = 0;
	const int dim = d.is3D ? 3 : 2;
	for (int c = 0; c < dim; c++) {
		const Real factor = timeStep * ((Real)1.0 > avg[c] ? (Real)1.0 : avg[c]);
		int lo = cur[c] - 1, up = cur[c] + 1;
		for (int dd = 0; dd < 2; dd++) {
			const bool fromLower = lo >= 0 && lo < size[c] && (flags[idx + (IndexInt)(lo - cur[c]) * stride[c]] & TypeFluid);
			const bool fromUpper = up >= 0 && up < size[c] && (flags[idx + (IndexInt)(up - cur[c]) * stride[c]] & TypeFluid);
			if (fromLower || fromUpper) {
				if (fromLower) { const IndexInt q = idx - stride[c];      // the value is always taken from the DIRECT neighbour (:367 uses `low`, not `flLow`)
					o0 += ((v0 - p0) / factor) + vel[3 * q]; o1 += ((v1 - p1) / factor) + vel[3 * q + 1]; o2 += ((v2 - p2) / factor) + vel[3 * q + 2]; cnt++; }
				if (fromUpper) { const IndexInt q = idx + stride[c];
					o0 += ((v0 - p0) / factor) + vel[3 * q]; o1 += ((v1 - p1) / factor) + vel[3 * q + 1]; o2 += ((v2 - p2) / factor) + vel[3 * q + 2]; cnt++; }
				break;
			}
			lo--; up++;
		}
	}
	if (cnt > 0) { o0 /= (Real)cnt; o1 /= (Real)cnt; o2 /= (Real)cnt; }
	velDst[3 * idx] = o0; velDst[3 * idx + 1] = o1; velDst[3 * idx + 2] = o2;
}
template <typename Real>
__global__ void __launch_bounds__(128) k_outflow_copy(Dims d, const int* __restrict__ flags, const Real* __restrict__ velDst, Real* __restrict__ vel, const int* __restrict__ anyOutflow) {
	if (anyOutflow && !*anyOutflow) return;
	int i, j, k; IndexInt idx;
	if (!cellOf(d, i, j, k, idx) || !(flags[idx] & TypeOutflow)) return;
	vel[3 * idx] = velDst[3 * idx]; vel[3 * idx + 1] = velDst[3 * idx + 1]; vel[3 * idx + 2] = velDst[3 * idx + 2];
}

struct Tmp {      // scratch grid of the context's pool, released on scope exit
	mp_grid* g = nullptr;
	~Tmp() { if (g) mp_grid_destroy(g); }
};

template <typename Real>
int applyOutflowBC(mp_context* ctx, const Dims& d, const mp_grid* flags, mp_grid* vel, const mp_grid* velPrev, double dt, const int* anyOutflow) {
	Tmp t; MP_TRY(mp_grid_create_scratch(ctx, MP_GRID_MAC, vel->prec, vel->sx, vel->sy, vel->sz, &t.g));     // only its outflow cells are written and read
	const double ts = 1.0 > dt * 4 ? 1.0 : dt * 4;
	k_outflow_extrapolate<Real><<<cellGrid(d), 128, 0, ctx->stream>>>(d, (const int*)flags->d, (const Real*)vel->d, (Real*)t.g->d, (const Real*)velPrev->d, (Real)ts, anyOutflow); MP_CHECK_LAUNCH(ctx);
	k_outflow_copy<Real><<<cellGrid(d), 128, 0, ctx->stream>>>(d, (const int*)flags->d, (const Real*)t.g->d, (Real*)vel->d, anyOutflow); MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

// result grid `neu` becomes the content of `grid` (the reference swaps the data pointers, grid.cpp:99-110)
int adopt(mp_context* ctx, mp_grid* grid, mp_grid* neu) {
	if (grid->owns && neu->owns) { void* p = grid->d; grid->d = neu->d; neu->d = p; return MP_OK; }
	MP_CUDA(cudaMemcpyAsync(grid->d, neu->d, grid->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	return MP_OK;
}

// Two passes per advection: (1) the forward trace, (2, MacCormack only) backward trace + correction + clamping; the reference's
// four kernels and three temporaries (fwd, bwd, newGrid) become two kernels and two temporaries, every cell written exactly once.
template <typename Real, int OS, int OT, bool D3>
int advect(mp_context* ctx, const mp_grid* flags, const mp_grid* vel, mp_grid* grid, int order, double strength, int clampMode, double dt_, bool vec3) {
	const Dims d = dimsOf(flags);
	const bool mac = grid->kind == MP_GRID_MAC && !vec3;
	const Real dt = (Real)dt_;
	const dim3 cg = cellGrid(d);
	const int* F = (const int*)flags->d; const Real* V = (const Real*)vel->d;
	Tmp fwd; MP_TRY(mp_grid_create_scratch(ctx, grid->kind, grid->prec, grid->sx, grid->sy, grid->sz, &fwd.g));
	if (vec3)     k_semi_lagrange_vec3<Real, OS, OT, D3><<<cg, 128, 0, ctx->stream>>>(d, V, (Real*)fwd.g->d, (const Real*)grid->d, dt);
	else if (mac) k_semi_lagrange_mac<Real, OS, OT, D3><<<cg, 128, 0, ctx->stream>>>(d, V, (Real*)fwd.g->d, (const Real*)grid->d, dt);
	else          k_semi_lagrange<Real, OS, OT, D3><<<cg, 128, 0, ctx->stream>>>(d, V, (Real*)fwd.g->d, (const Real*)grid->d, dt);
	MP_CHECK_LAUNCH(ctx);
	if (order == 1) {
		if (mac) MP_TRY(applyOutflowBC<Real>(ctx, d, flags, fwd.g, grid, (double)dt, nullptr));
		return adopt(ctx, grid, fwd.g);
	}
	Tmp neu; MP_TRY(mp_grid_create_scratch(ctx, grid->kind, grid->prec, grid->sx, grid->sy, grid->sz, &neu.g));
	if (vec3) {
		k_mc_rest_vec3<Real, OS, OT, D3><<<cg, 128, 0, ctx->stream>>>(d, F, V, (Real*)neu.g->d, (const Real*)grid->d, (const Real*)fwd.g->d, dt, (Real)strength, clampMode); MP_CHECK_LAUNCH(ctx);
	} else if (mac) {
		int* any = (int*)(ctx->dScal + 24);
		MP_CUDA(cudaMemsetAsync(any, 0, sizeof(int), ctx->stream));
		k_mc_rest_mac<Real, OS, OT, D3><<<cg, 128, 0, ctx->stream>>>(d, F, V, (Real*)neu.g->d, (const Real*)grid->d, (const Real*)fwd.g->d, dt, (Real)strength, clampMode, any); MP_CHECK_LAUNCH(ctx);
		MP_TRY(applyOutflowBC<Real>(ctx, d, flags, neu.g, grid, (double)dt, any));
	} else {
		k_mc_rest<Real, OS, OT, D3><<<cg, 128, 0, ctx->stream>>>(d, F, V, (Real*)neu.g->d, (const Real*)grid->d, (const Real*)fwd.g->d, dt, (Real)strength, clampMode); MP_CHECK_LAUNCH(ctx);
	}
	return adopt(ctx, grid, neu.g);
}

int checkStep(const char* who, mp_context* ctx, const mp_grid* flags, const mp_grid* vel) {
	if (!ctx || !flags || !vel) MP_FAIL(MP_ERR_INVALID, "%s: NULL argument", who);
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "%s: flags is not a FlagGrid", who);
	MP_TRY(mp_check_same(flags, vel, MP_GRID_MAC, "vel", false));
	if (ctx->dist && ctx->dist->active) MP_FAIL(MP_ERR_UNSUPPORTED, "%s: not available on z-slab sharded grids yet", who);
	MP_CUDA(cudaSetDevice(ctx->device));
	return MP_OK;
}
static inline int imax3(int a, int b, int c) { const int m = a > b ? a : b; return m > c ? m : c; }

}  // namespace

extern "C" {

int mp_set_wall_bcs(mp_context* ctx, const mp_grid* flags, mp_grid* vel, const mp_grid* obvel, const mp_grid* fractions, const mp_grid* phiObs, int boundaryWidth)
{
	MP_TRY(checkStep("mp_set_wall_bcs", ctx, flags, vel));
	(void)boundaryWidth;
	if (phiObs && fractions) {       // the second-order variant KnSetWallBcsFrac (extforces.cpp:220-303); it reads neither fractions nor obvel
		MP_TRY(mp_check_same(vel, fractions, MP_GRID_MAC, "fractions", false));
		return mp_set_wall_bcs_frac_impl(ctx, flags, vel, phiObs);
	}
	if (obvel) MP_TRY(mp_check_same(vel, obvel, MP_GRID_MAC, "obvel", false));
	const Dims d = dimsOf(flags);
	if (vel->prec == 4) { WallBcsOp<float> op = { (const int*)flags->d, obvel ? (const float*)obvel->d : nullptr }; return launchCells<float>(ctx, d, (float*)vel->d, op); }
	WallBcsOp<double> op = { (const int*)flags->d, obvel ? (const double*)obvel->d : nullptr };
	return launchCells<double>(ctx, d, (double*)vel->d, op);
}

int mp_add_gravity(mp_context* ctx, const mp_grid* flags, mp_grid* vel, double gx, double gy, double gz, const mp_grid* exclude, int scale, double dt)
{
	MP_TRY(checkStep("mp_add_gravity", ctx, flags, vel));
	if (exclude) MP_TRY(mp_check_same(flags, exclude, MP_GRID_REAL, "exclude", false));
	const Dims d = dimsOf(flags);
	const double g[3] = { gx, gy, gz };
	if (vel->prec == 4) {
		const float gridScale = scale ? (float)(float)(1.0 / imax3(d.sx, d.sy, d.sz)) : 1.f; float f[3];
		for (int c = 0; c < 3; c++) f[c] = ((float)g[c] * (float)dt) / gridScale;
		ApplyForceOp<float> op = { (const int*)flags->d, exclude ? (const float*)exclude->d : nullptr, f[0], f[1], f[2] };
		return launchCells<float>(ctx, d, (float*)vel->d, op);
	} else {
		const float gridScale = scale ? (float)(1.0 / imax3(d.sx, d.sy, d.sz)) : 1.f; double f[3];
		for (int c = 0; c < 3; c++) f[c] = (g[c] * dt) / gridScale;
		ApplyForceOp<double> op = { (const int*)flags->d, exclude ? (const double*)exclude->d : nullptr, f[0], f[1], f[2] };
		return launchCells<double>(ctx, d, (double*)vel->d, op);
	}
}

int mp_add_buoyancy(mp_context* ctx, const mp_grid* flags, const mp_grid* density, mp_grid* vel, double gx, double gy, double gz, double coefficient, int scale, double dt)
{
	MP_TRY(checkStep("mp_add_buoyancy", ctx, flags, vel));
	if (!density) MP_FAIL(MP_ERR_INVALID, "mp_add_buoyancy: NULL density");
	MP_TRY(mp_check_same(flags, density, MP_GRID_REAL, "density", false));
	const Dims d = dimsOf(flags);
	const double g[3] = { gx, gy, gz };
	if (vel->prec == 4) {
		const float gridScale = scale ? (float)(float)(1.0 / imax3(d.sx, d.sy, d.sz)) : 1.f; float f[3];
		for (int c = 0; c < 3; c++) f[c] = (((-(float)g[c]) * (float)dt) / gridScale) * (float)coefficient;
		BuoyancyOp<float> op = { (const int*)flags->d, (const float*)density->d, f[0], f[1], f[2] };
		return launchCells<float>(ctx, d, (float*)vel->d, op);
	} else {
		const float gridScale = scale ? (float)(1.0 / imax3(d.sx, d.sy, d.sz)) : 1.f; double f[3];
		for (int c = 0; c < 3; c++) f[c] = (((-g[c]) * dt) / gridScale) * coefficient;
		BuoyancyOp<double> op = { (const int*)flags->d, (const double*)density->d, f[0], f[1], f[2] };
		return launchCells<double>(ctx, d, (double*)vel->d, op);
	}
}

static int advectEntry(mp_context* ctx, const mp_grid* flags, const mp_grid* vel, mp_grid* grid, int order, double strength, int orderSpace,
                       int clampMode, int orderTrace, double dt, bool vec3)
{
	MP_TRY(checkStep("mp_advect_semi_lagrange", ctx, flags, vel));
	if (!grid) MP_FAIL(MP_ERR_INVALID, "mp_advect_semi_lagrange: NULL grid");
	if (order != 1 && order != 2) MP_FAIL(MP_ERR_INVALID, "AdvectSemiLagrange: Only order 1 (regular SL) and 2 (MacCormack) supported");
	if (orderSpace != 1 && orderSpace != 2) MP_FAIL(MP_ERR_INVALID, "Unknown interpolation order %d", orderSpace);      // grid.h:150
	if (orderTrace != 1 && orderTrace != 2) MP_FAIL(MP_ERR_INVALID, "Unknown backtracing order %d", orderTrace);        // advection.cpp:39
	if (grid->kind != MP_GRID_REAL && grid->kind != MP_GRID_MAC) MP_FAIL(MP_ERR_INVALID, "AdvectSemiLagrange: Grid Type is not supported (only Real, Vec3, MAC, Levelset)");
	if (vec3 && grid->kind != MP_GRID_MAC) MP_FAIL(MP_ERR_INVALID, "mp_advect_semi_lagrange_vec3: grid does not hold Vec3 cells");
	MP_TRY(mp_check_same(flags, grid, grid->kind, "grid", false));
	if (flags->sx < 3 || flags->sy < 3 || (flags->sz > 1 && flags->sz < 3)) return MP_OK;       // no interior cells
#define MP_ADV(R, D3) (orderSpace == 1 ? (orderTrace == 1 ? advect<R, 1, 1, D3> : advect<R, 1, 2, D3>) : (orderTrace == 1 ? advect<R, 2, 1, D3> : advect<R, 2, 2, D3>))(ctx, flags, vel, grid, order, strength, clampMode, dt, vec3)
	const bool is3D = flags->sz > 1;
	if (grid->prec == 4) return is3D ? MP_ADV(float, true) : MP_ADV(float, false);
	return is3D ? MP_ADV(double, true) : MP_ADV(double, false);
#undef MP_ADV
}
int mp_advect_semi_lagrange(mp_context* ctx, const mp_grid* flags, const mp_grid* vel, mp_grid* grid, int order, double strength, int orderSpace,
                            int clampMode, int orderTrace, double dt)
{
	return advectEntry(ctx, flags, vel, grid, order, strength, orderSpace, clampMode, orderTrace, dt, false);
}
int mp_advect_semi_lagrange_vec3(mp_context* ctx, const mp_grid* flags, const mp_grid* vel, mp_grid* grid, int order, double strength, int orderSpace,
                                 int clampMode, int orderTrace, double dt)
{
	return advectEntry(ctx, flags, vel, grid, order, strength, orderSpace, clampMode, orderTrace, dt, true);
}

}

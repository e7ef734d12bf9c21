// TEST INFRASTRUCTURE ONLY -- never linked into libmantapress.so and never used by the product path.
//
// Host emulation of the fused level-0 V-cycle kernels: walks the phase functions of mantaflow_b200/csrc/mp_mg_l0_fused.cuh (the code
// k_mg_l0_fused in mp_mg.cu runs, one call per thread with __syncthreads() between the phases) CTA by CTA, thread by thread, in the
// kernel's own phase order.  The build container has no GPU; this lets `pytest -m "not gpu"` check the tile / halo / ring logic and the
// arithmetic against a plain numpy statement of knSmoothColor / knCalcResidual (multigrid.cpp:668-771).
// Built by tests/test_mg_l0_fused_emul.py:  g++ -O2 -ffp-contract=off -shared -fPIC
#include "../../mantaflow_b200/csrc/mp_mg_l0_fused.cuh"
#include <vector>
#include <algorithm>

using namespace mgl0;

template <typename Real>
static int buildMask(int sx, int sy, int sz, int is3D, const Real* A, const signed char* type, unsigned short* mask) {
	Geom g = { sx, sy, sz }; int bad = 0;
	for (int z = 0; z < sz; z++) for (int y = 0; y < sy; y++) for (int x = 0; x < sx; x++)
		mask[(size_t)x + (size_t)sx * (y + (size_t)sy * z)] = maskOf<Real>(g, is3D, x, y, z, A, type, &bad);
	return bad;
}

// order: 0 threads ascending, 1 descending (the phases must not depend on the order of the threads inside a phase)
template <typename Real, int MODE>
static void run(int sx, int sy, int sz, int kchunk, int cFirst, int cSecond, const Real* A0, const Real* b, Real bscale, const unsigned short* mask,
                const Real* xin, Real* xout, Real* rout, int order) {
	typedef Tile<Real> T;
	const Geom g = { sx, sy, sz };
	std::vector<Smem<Real>> smv(1); Smem<Real>& s = smv[0];
	std::vector<Pre<Real>> pre(T::NTHR);
	std::vector<Ctx<Real>> ctx(T::NTHR);
	const int tilesX = (sx + T::TX - 1) / T::TX, tilesY = (sy + T::TY - 1) / T::TY, nchunk = (sz + kchunk - 1) / kchunk;
	auto tidOf = [&](int t) { return order ? T::NTHR - 1 - t : t; };
	const int cm = MODE == MODE_DOWN ? 1 - cFirst : cFirst;
	for (int bz = 0; bz < nchunk; bz++) for (int by = 0; by < tilesY; by++) for (int bx = 0; bx < tilesX; bx++) {
		// shared memory comes uninitialised
		memset(&s, 0x7f, sizeof(s));
		const int x0 = bx * T::TX, y0 = by * T::TY, k0 = bz * kchunk, k1 = std::min(sz, k0 + kchunk);
		for (int t = 0; t < T::NTHR; t++) ctx[t] = makeCtx<Real>(g, x0, y0, t);
		for (int q = k0 - 2; q <= k0 + 1; q++)
			for (int t = 0; t < T::NTHR; t++) { const int tid = tidOf(t); issue<Real, MODE>(g, ctx[tid], q, b, xin, mask, pre[tid]); stage<Real, MODE>(g, ctx[tid], q, bscale, A0, cFirst, pre[tid], s); }
		for (int t = 0; t < T::NTHR; t++) { const int tid = tidOf(t); mid<Real>(g, ctx[tid], k0 - 1, cm, A0, s); mid<Real>(g, ctx[tid], k0, cm, A0, s); }
		for (int t = 0; t < T::NTHR; t++) { const int tid = tidOf(t); issue<Real, MODE>(g, ctx[tid], k0 + 2, b, xin, mask, pre[tid]); }
		for (int sp = k0; sp < k1; sp++) {
			for (int t = 0; t < T::NTHR; t++) {
				const int tid = tidOf(t);
				stage<Real, MODE>(g, ctx[tid], sp + 2, bscale, A0, cFirst, pre[tid], s);
				if (sp + 1 < k1) issue<Real, MODE>(g, ctx[tid], sp + 3, b, xin, mask, pre[tid]);
			}
			for (int t = 0; t < T::NTHR; t++) mid<Real>(g, ctx[tidOf(t)], sp + 1, cm, A0, s);
			for (int t = 0; t < T::NTHR; t++) last<Real, MODE>(g, ctx[tidOf(t)], sp, cSecond, A0, s, xout, rout);
		}
	}
}

extern "C" {
int mgl0_build_mask_f32(int sx, int sy, int sz, int is3D, const float* A, const signed char* type, unsigned short* mask) { return buildMask<float>(sx, sy, sz, is3D, A, type, mask); }
int mgl0_build_mask_f64(int sx, int sy, int sz, int is3D, const double* A, const signed char* type, unsigned short* mask) { return buildMask<double>(sx, sy, sz, is3D, A, type, mask); }
void mgl0_run_f32(int mode, int sx, int sy, int sz, int kchunk, int cFirst, int cSecond, const float* A0, const float* b, float bscale, const unsigned short* mask, const float* xin, float* xout, float* rout, int order) {
	if (mode == MODE_DOWN) run<float, MODE_DOWN>(sx, sy, sz, kchunk, cFirst, cSecond, A0, b, bscale, mask, xin, xout, rout, order);
	else run<float, MODE_SMOOTH>(sx, sy, sz, kchunk, cFirst, cSecond, A0, b, bscale, mask, xin, xout, rout, order);
}
void mgl0_run_f64(int mode, int sx, int sy, int sz, int kchunk, int cFirst, int cSecond, const double* A0, const double* b, double bscale, const unsigned short* mask, const double* xin, double* xout, double* rout, int order) {
	if (mode == MODE_DOWN) run<double, MODE_DOWN>(sx, sy, sz, kchunk, cFirst, cSecond, A0, b, bscale, mask, xin, xout, rout, order);
	else run<double, MODE_SMOOTH>(sx, sy, sz, kchunk, cFirst, cSecond, A0, b, bscale, mask, xin, xout, rout, order);
}
int mgl0_tile(int prec, int* tx, int* ty) { if (prec == 4) { *tx = Tile<float>::TX; *ty = Tile<float>::TY; } else { *tx = Tile<double>::TX; *ty = Tile<double>::TY; } return 0; }
}

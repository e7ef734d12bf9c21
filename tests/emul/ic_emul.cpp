// TEST INFRASTRUCTURE ONLY -- never linked into libmantapress.so and never used by the product path.
//
// Host emulation of the IC(0) kernels (mantaflow_b200/csrc/mp_ic.cu): the per-cell code and the hyperplane geometry of mp_ic_cells.cuh,
// walked plane by plane on the host -- inside a plane in ascending (order 0) or descending (order 1) (k, j) order: cells of a plane are
// independent, which the tests assert.  Built by tests/test_oracle_icp.py:  g++ -O2 -ffp-contract=off -I/usr/local/cuda/include -shared -fPIC
#include "../../mantaflow_b200/csrc/mp_ic_cells.cuh"
#include <cstdarg>

void mp_set_error(const char*, ...) {}

namespace {
template <typename F> void walkPlane(const ic::Geom& g, int c, int order, const F& f) {
	ic::PlaneRange r;
	if (!ic::planeRange(g, c, r)) return;
	const int nk = r.khi - r.klo + 1, nj = ((g.hy + 127) / 128) * 128;      // the launch covers j in [1, nj]: whole blocks of 128 threads
	for (int t = 0; t < nk * nj; t++) {
		const int u = order ? nk * nj - 1 - t : t;
		IndexInt idx;
		if (ic::planeCell(g, c, 1 + u % nj, r.klo + u / nj, idx)) f(idx);
	}
}
template <typename Real> int icInit(int order, int sx, int sy, int sz, const int* flags, Real* P0, Real* Pi, Real* Pj, Real* Pk, const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak) {
	const IndexInt n = (IndexInt)sx * sy * sz, Y = sx, Z = (IndexInt)sx * sy;
	memcpy(P0, A0, sizeof(Real) * n); memcpy(Pi, Ai, sizeof(Real) * n); memcpy(Pj, Aj, sizeof(Real) * n); memcpy(Pk, Ak, sizeof(Real) * n);
	if (sx < 3 || sy < 3 || sz < 3) return MP_OK;
	const ic::Geom g = { sx, sy, sz, Y, Z, sx - 1, sy - 1, sz - 1 };
	for (int c = 3; c <= g.hx + g.hy + g.hz; c++) walkPlane(g, c, order, [&](IndexInt idx) { ic::initCell<Real>(flags, P0, Pi, Pj, Pk, A0, Ai, Aj, Ak, idx, Y, Z); });
	return MP_OK;
}
template <typename Real> int icApply(int order, int sx, int sy, int sz, const int* flags, Real* dst, const Real* src, const Real* P0, const Real* Pi, const Real* Pj, const Real* Pk) {
	const IndexInt Y = sx, Z = (IndexInt)sx * sy;
	if (sx < 3 || sy < 3 || sz < 3) return MP_OK;
	const ic::Geom g = { sx, sy, sz, Y, Z, sx - 2, sy - 2, sz - 2 };
	const int cmax = g.hx + g.hy + g.hz;
	for (int c = 3; c <= cmax; c++) walkPlane(g, c, order, [&](IndexInt idx) { ic::fwdCell<Real>(flags, dst, src, P0, Pi, Pj, Pk, idx, Y, Z); });
	for (int c = cmax; c >= 3; c--) walkPlane(g, c, order, [&](IndexInt idx) { ic::bwdCell<Real>(flags, dst, P0, Pi, Pj, Pk, idx, Y, Z); });
	return MP_OK;
}
}  // namespace

extern "C" {
int emu_ic_init(int prec, int order, int sx, int sy, int sz, const int* flags, void* P0, void* Pi, void* Pj, void* Pk, const void* A0, const void* Ai, const void* Aj, const void* Ak) {
	return prec == 4 ? icInit<float>(order, sx, sy, sz, flags, (float*)P0, (float*)Pi, (float*)Pj, (float*)Pk, (const float*)A0, (const float*)Ai, (const float*)Aj, (const float*)Ak)
	                 : icInit<double>(order, sx, sy, sz, flags, (double*)P0, (double*)Pi, (double*)Pj, (double*)Pk, (const double*)A0, (const double*)Ai, (const double*)Aj, (const double*)Ak);
}
int emu_ic_apply(int prec, int order, int sx, int sy, int sz, const int* flags, void* dst, const void* src, const void* P0, const void* Pi, const void* Pj, const void* Pk) {
	return prec == 4 ? icApply<float>(order, sx, sy, sz, flags, (float*)dst, (const float*)src, (const float*)P0, (const float*)Pi, (const float*)Pj, (const float*)Pk)
	                 : icApply<double>(order, sx, sy, sz, flags, (double*)dst, (const double*)src, (const double*)P0, (const double*)Pi, (const double*)Pj, (const double*)Pk);
}
}

// TEST INFRASTRUCTURE ONLY -- never linked into libmantapress.so and never used by the product path.
//
// Host emulation of the liquid-neighbour kernels: instantiates the per-cell operations and pass sequences of
// mantaflow_b200/csrc/mp_liquid_cells.cuh (the code the CUDA kernels of mp_liquid.cu run, one thread per cell) with an executor that
// walks the cells in a host loop.  The build container has no GPU; this lets `pytest -m "not gpu"` check the arithmetic and the pass
// structure of those kernels against the unmodified reference.  `order` selects the walk (0 lexicographic, 1 reverse, 2 a strided
// permutation, 3 the block / thread geometry of the CUDA launch): the passes are written so that any order gives the same result, which the tests assert.
// Built by tests/test_liquid_emulation.py:  g++ -O2 -ffp-contract=off -I/usr/local/cuda/include -shared -fPIC
#include "../../mantaflow_b200/csrc/mp_liquid_cells.cuh"
#include <cstdarg>
#include <cstdlib>

void mp_set_error(const char*, ...) {}

namespace {
struct HostExec {
	int order;
	static IndexInt gcd(IndexInt a, IndexInt b) { while (b) { const IndexInt t = a % b; a = b; b = t; } return a; }
	template <typename F> int cells(const Dims& d, const F& f) {
		if (order == 3) {      // the CUDA launch: every thread of every block of liquid::launchGeomOf(d), through the kernel's own cell mapping
			const liquid::LaunchGeom g = liquid::launchGeomOf(d);
			for (unsigned bz = 0; bz < g.gz; bz++) for (unsigned by = 0; by < g.gy; by++) for (unsigned bx = 0; bx < g.gx; bx++)
				for (int tx = liquid::kThreads - 1; tx >= 0; tx--) liquid::threadCells(d, f, (int)bx, (int)by, (int)bz, tx);
			return MP_OK;
		}
		const IndexInt n = d.n;
		IndexInt stride = 1;
		if (order == 2) { stride = 7919; while (gcd(stride, n) != 1) stride += 2; }       // a stride coprime to n visits every cell exactly once
		for (IndexInt t = 0; t < n; t++) {
			IndexInt idx = t;
			if (order == 1) idx = n - 1 - t;
			else if (order == 2) idx = (t * stride) % n;
			const int i = (int)(idx % d.sx), j = (int)((idx / d.sx) % d.sy), k = (int)(idx / ((IndexInt)d.sx * d.sy));
			liquid::oneCell(d, f, i, j, k, idx);
		}
		return MP_OK;
	}
};
Dims mkDims(int sx, int sy, int sz) {
	mp_grid g; g.ctx = nullptr; g.kind = MP_GRID_REAL; g.prec = 4; g.sx = sx; g.sy = sy; g.sz = sz; g.n = (IndexInt)sx * sy * sz; g.bytes = 0; g.d = nullptr; g.owns = false;
	return dimsOf(&g);
}
template <typename Real> int macSimple(int order, int sx, int sy, int sz, const int* flags, Real* vel, int distance, const Real* phiObs, int intoObs) {
	const Dims d = mkDims(sx, sy, sz);
	int* tmp = (int*)malloc(sizeof(int) * (size_t)d.n);
	Real* stage = (Real*)malloc(sizeof(Real) * 3 * (size_t)d.n);
	for (IndexInt q = 0; q < d.n; q++) tmp[q] = 0x5a5a5a5a;          // scratch grids come uninitialised from the pool
	for (IndexInt q = 0; q < 3 * d.n; q++) stage[q] = (Real)1e30;
	HostExec ex = { order };
	const int rc = liquid::extrapolateMacSimple<Real>(ex, d, flags, vel, distance, phiObs, intoObs != 0, tmp, stage);
	free(tmp); free(stage);
	return rc;
}
template <typename Real> int lsSimple(int order, int sx, int sy, int sz, Real* val, const Real* phi, int distance, int inside, bool vec3) {
	const Dims d = mkDims(sx, sy, sz);
	int* tmp = (int*)malloc(sizeof(int) * (size_t)d.n);
	for (IndexInt q = 0; q < d.n; q++) tmp[q] = 0x5a5a5a5a;
	HostExec ex = { order };
	int rc;
	if (vec3) rc = liquid::extrapolateLs<Real, 3>(ex, d, val, phi, distance, inside != 0, (Real)0, (Real)0, tmp);
	else { const Real direction = inside ? (Real)-1. : (Real)1.;
	       rc = liquid::extrapolateLs<Real, 1>(ex, d, val, phi, distance, inside != 0, direction, (Real)(direction * (distance + 2)), tmp); }
	free(tmp);
	return rc;
}
}  // namespace

extern "C" {
int emu_extrapolate_mac_simple(int prec, int order, int sx, int sy, int sz, const int* flags, void* vel, int distance, const void* phiObs, int intoObs) {
	return prec == 4 ? macSimple<float>(order, sx, sy, sz, flags, (float*)vel, distance, (const float*)phiObs, intoObs)
	                 : macSimple<double>(order, sx, sy, sz, flags, (double*)vel, distance, (const double*)phiObs, intoObs);
}
int emu_extrapolate_mac_from_weight(int prec, int order, int sx, int sy, int sz, void* vel, void* weight, int distance) {
	const Dims d = mkDims(sx, sy, sz); HostExec ex = { order };
	return prec == 4 ? liquid::extrapolateMacFromWeight<float>(ex, d, (float*)vel, (float*)weight, distance)
	                 : liquid::extrapolateMacFromWeight<double>(ex, d, (double*)vel, (double*)weight, distance);
}
int emu_extrapolate_ls_simple(int prec, int order, int sx, int sy, int sz, void* phi, int distance, int inside) {
	return prec == 4 ? lsSimple<float>(order, sx, sy, sz, (float*)phi, (const float*)phi, distance, inside, false)
	                 : lsSimple<double>(order, sx, sy, sz, (double*)phi, (const double*)phi, distance, inside, false);
}
int emu_extrapolate_vec3_simple(int prec, int order, int sx, int sy, int sz, void* vel, const void* phi, int distance, int inside) {
	return prec == 4 ? lsSimple<float>(order, sx, sy, sz, (float*)vel, (const float*)phi, distance, inside, true)
	                 : lsSimple<double>(order, sx, sy, sz, (double*)vel, (const double*)phi, distance, inside, true);
}
int emu_update_from_levelset(int prec, int order, int sx, int sy, int sz, int* flags, const void* phi) {
	const Dims d = mkDims(sx, sy, sz); HostExec ex = { order };
	if (prec == 4) { liquid::UpdateFromLevelset<float> op = { flags, (const float*)phi }; return ex.cells(d, op); }
	liquid::UpdateFromLevelset<double> op = { flags, (const double*)phi }; return ex.cells(d, op);
}
int emu_set_wall_bcs_frac(int prec, int order, int sx, int sy, int sz, const int* flags, void* vel, const void* phiObs) {
	const Dims d = mkDims(sx, sy, sz); HostExec ex = { order };
	const size_t bytes = (size_t)prec * 3 * (size_t)d.n;
	void* tgt = malloc(bytes); memset(tgt, 0x5a, bytes);
	int rc;
	if (prec == 4) { liquid::WallBcsFrac<float> op = { flags, (const float*)vel, (float*)tgt, (const float*)phiObs }; rc = ex.cells(d, op); }
	else { liquid::WallBcsFrac<double> op = { flags, (const double*)vel, (double*)tgt, (const double*)phiObs }; rc = ex.cells(d, op); }
	memcpy(vel, tgt, bytes); free(tgt);
	return rc;
}
int emu_update_fractions(int prec, int order, int sx, int sy, int sz, const int* flags, const void* phiObs, void* fractions, int w, double thr) {
	const Dims d = mkDims(sx, sy, sz); HostExec ex = { order };
	if (prec == 4) { liquid::UpdateFractions<float> op = { flags, (const float*)phiObs, (float*)fractions, w, (float)thr }; return ex.cells(d, op); }
	liquid::UpdateFractions<double> op = { flags, (const double*)phiObs, (double*)fractions, w, thr }; return ex.cells(d, op);
}
int emu_set_obstacle_flags(int prec, int order, int sx, int sy, int sz, int* flags, const void* phiObs, const void* fractions, const void* phiOut, const void* phiIn, int bw) {
	const Dims d = mkDims(sx, sy, sz); HostExec ex = { order };
	if (prec == 4) { liquid::SetObstacleFlags<float> op = { flags, (const float*)phiObs, (const float*)fractions, (const float*)phiOut, (const float*)phiIn, bw }; return ex.cells(d, op); }
	liquid::SetObstacleFlags<double> op = { flags, (const double*)phiObs, (const double*)fractions, (const double*)phiOut, (const double*)phiIn, bw }; return ex.cells(d, op);
}
int emu_stencil(int prec, int order, int sx, int sy, int sz, void* out, const void* grid, double h, int curvature) {
	const Dims d = mkDims(sx, sy, sz); HostExec ex = { order };
	if (prec == 4) {
		if (curvature) { liquid::CurvatureCell<float> op = { (float*)out, (const float*)grid, (float)h }; return ex.cells(d, op); }
		liquid::LaplaceCell<float> op = { (float*)out, (const float*)grid }; return ex.cells(d, op);
	}
	if (curvature) { liquid::CurvatureCell<double> op = { (double*)out, (const double*)grid, h }; return ex.cells(d, op); }
	liquid::LaplaceCell<double> op = { (double*)out, (const double*)grid }; return ex.cells(d, op);
}
int emu_set_bound(int prec, int order, int sx, int sy, int sz, void* grid, int ncomp, double value, int w) {
	const Dims d = mkDims(sx, sy, sz); HostExec ex = { order };
	if (prec == 4) {
		if (ncomp == 1) { liquid::SetBound<float, 1> op = { (float*)grid, { (float)value }, w }; return ex.cells(d, op); }
		liquid::SetBound<float, 3> op = { (float*)grid, { (float)value, (float)value, (float)value }, w }; return ex.cells(d, op);
	}
	if (ncomp == 1) { liquid::SetBound<double, 1> op = { (double*)grid, { value }, w }; return ex.cells(d, op); }
	liquid::SetBound<double, 3> op = { (double*)grid, { value, value, value }, w }; return ex.cells(d, op);
}
}

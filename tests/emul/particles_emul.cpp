// TEST INFRASTRUCTURE ONLY -- never linked into libmantapress.so and never used by the product path.
//
// Host emulation of the FLIP particle <-> grid kernels: instantiates the per-cell / per-particle operations and the pass sequences of
// mantaflow_b200/csrc/mp_particles_cells.cuh (the code the CUDA kernels of mp_particles.cu run) with an executor that walks cells and
// particles in host loops, scans serially and sorts with std::stable_sort (the CUDA executor uses cub's scan and stable radix sort).
// `order`: 0 lexicographic, 1 reverse, 2 a strided permutation, 3 the block / thread geometry of the CUDA launch (cells) / reverse
// (particles) -- the passes are written so that any order gives the same result, which the tests assert.
// Built by tests/test_oracle_flip.py:  g++ -O2 -ffp-contract=off -I/usr/local/cuda/include -shared -fPIC
#include "../../mantaflow_b200/csrc/mp_particles_cells.cuh"
#include "../../mantaflow_b200/csrc/mp_gridops.cuh"
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <vector>

void mp_set_error(const char*, ...) {}

namespace {
struct HostExec {
	int order;
	static IndexInt gcd(IndexInt a, IndexInt b) { while (b) { const IndexInt t = a % b; a = b; b = t; } return a; }
	static IndexInt permuted(int order, IndexInt t, IndexInt n, IndexInt stride) {
		if (order == 1 || order == 3) return n - 1 - t;
		if (order == 2) return (t * stride) % n;
		return t;
	}
	static IndexInt strideFor(int order, IndexInt n) {
		IndexInt stride = 1;
		if (order == 2) { stride = 7919; while (gcd(stride, n) != 1) stride += 2; }
		return stride;
	}
	template <typename F> int cells(const Dims& d, const F& f) {
		if (order == 3) {      // the CUDA launch: every thread of every block of liquid::launchGeomOf(d), through the kernel's own cell mapping
			const liquid::LaunchGeom g = liquid::launchGeomOf(d);
			for (unsigned bz = 0; bz < g.gz; bz++) for (unsigned by = 0; by < g.gy; by++) for (unsigned bx = 0; bx < g.gx; bx++)
				for (int tx = liquid::kThreads - 1; tx >= 0; tx--) liquid::threadCells(d, f, (int)bx, (int)by, (int)bz, tx);
			return MP_OK;
		}
		const IndexInt n = d.n, stride = strideFor(order, n);
		for (IndexInt t = 0; t < n; t++) {
			const IndexInt idx = permuted(order, t, n, stride);
			const int i = (int)(idx % d.sx), j = (int)((idx / d.sx) % d.sy), k = (int)(idx / ((IndexInt)d.sx * d.sy));
			liquid::oneCell(d, f, i, j, k, idx);
		}
		return MP_OK;
	}
	template <typename F> int parts(IndexInt np, const F& f) {
		const IndexInt stride = strideFor(order, np);
		for (IndexInt t = 0; t < np; t++) f(permuted(order, t, np, stride));
		return MP_OK;
	}
	int zero(void* p, size_t bytes) { memset(p, 0, bytes); return MP_OK; }
	int exclusiveScan(int* data, IndexInt n, IndexInt* total) {
		IndexInt run = 0;
		for (IndexInt q = 0; q < n; q++) { const int c = data[q]; data[q] = (int)run; run += c; }
		*total = run;
		return MP_OK;
	}
	int sortPairs(int* keys, int* keysTmp, int* vals, int* valsOut, IndexInt n, int keyBits) {
		for (IndexInt q = 0; q < n; q++) if (keys[q] < 0 || ((IndexInt)keys[q] >> keyBits) != 0) return MP_ERR_INVALID;      // the radix sort only looks at keyBits bits
		std::vector<IndexInt> perm((size_t)n);
		for (IndexInt q = 0; q < n; q++) perm[(size_t)q] = q;
		std::stable_sort(perm.begin(), perm.end(), [&](IndexInt a, IndexInt b) { return keys[a] < keys[b]; });
		for (IndexInt q = 0; q < n; q++) { keysTmp[q] = keys[perm[(size_t)q]]; valsOut[q] = vals[perm[(size_t)q]]; }
		return MP_OK;
	}
};
Dims mkDims(int sx, int sy, int sz) {
	mp_grid g; g.ctx = nullptr; g.kind = MP_GRID_REAL; g.prec = 4; g.sx = sx; g.sy = sy; g.sz = sz; g.n = (IndexInt)sx * sy * sz; g.bytes = 0; g.d = nullptr; g.owns = false;
	return dimsOf(&g);
}
int* scratchInts(IndexInt n) {       // scratch arrays come uninitialised from the pool
	if (n < 1) n = 1;
	int* p = (int*)malloc(sizeof(int) * (size_t)n);
	for (IndexInt q = 0; q < n; q++) p[q] = 0x5a5a5a5a;
	return p;
}

template <typename Real> int markFluid(int order, const Dims& d, int* flags, long long np, const Real* pos, const int* pflag, const Real* phiObs, const int* ptype, int exclude) {
	HostExec ex = { order };
	int* tmp = phiObs ? scratchInts(d.n) : nullptr;
	bool swapped = false;
	parts::PSet<Real> ps = { pos, pflag, ptype, exclude };
	const int rc = parts::markFluidCells<Real>(ex, d, flags, np, ps, phiObs, tmp, &swapped);
	if (swapped) memcpy(flags, tmp, sizeof(int) * (size_t)d.n);
	free(tmp);
	return rc;
}
template <typename Real> int particleIndex(int order, const Dims& d, long long np, const Real* pos, const int* pflag, int* index, int* indexSys, long long* count) {
	HostExec ex = { order };
	int *key = scratchInts(np), *keyTmp = scratchInts(np), *val = scratchInts(np);
	parts::PSet<Real> ps = { pos, pflag, nullptr, 0 };
	IndexInt c = 0;
	const int rc = parts::bucketParticles<Real>(ex, d, np, ps, false, index, key, keyTmp, val, indexSys, &c);
	*count = c;
	free(key); free(keyTmp); free(val);
	return rc;
}
template <typename Real> int mapParts(int order, const Dims& d, Real* vel, Real* velOld, long long np, const Real* pos, const int* pflag, const Real* pvel, Real* weight,
                                      const int* ptype, int exclude) {
	HostExec ex = { order };
	int *start = scratchInts(d.n), *key = scratchInts(np), *keyTmp = scratchInts(np), *val = scratchInts(np), *sorted = scratchInts(np);
	parts::PSet<Real> ps = { pos, pflag, ptype, exclude };
	// MP_MAPPARTS=0: the 27-way walk; default: the tree of 3-way merges (what mp_particles.cu launches) -- scratch comes uninitialised
	const char* e = getenv("MP_MAPPARTS");
	int rc;
	if (e && !atoi(e)) rc = parts::mapPartsToMAC<Real>(ex, d, vel, velOld, np, ps, pvel, weight, start, key, keyTmp, val, sorted);
	else {
		const IndexInt n = np < 1 ? 1 : np;
		parts::MapPartsTreeScratch<Real> t;
		t.len1 = scratchInts(d.n); t.off1 = scratchInts(d.n); t.len2 = scratchInts(d.n); t.off2 = scratchInts(d.n);
		void *raw1 = nullptr, *raw2 = nullptr;      // 32-byte aligned like a device allocation; filled with garbage
		const size_t b1 = sizeof(parts::Ent) * (size_t)parts::treeEntries(d.n, n, 3), b2 = sizeof(parts::Ent) * (size_t)parts::treeEntries(d.n, n, 9);
		if (posix_memalign(&raw1, 64, b1) || posix_memalign(&raw2, 64, b2)) return MP_ERR_CUDA;
		memset(raw1, 0x5a, b1); memset(raw2, 0x5a, b2);
		t.e1 = (parts::Ent*)raw1; t.e2 = (parts::Ent*)raw2;
		t.posS = (Real*)malloc(sizeof(Real) * 3 * (size_t)n); t.pvelS = (Real*)malloc(sizeof(Real) * 3 * (size_t)n);
		memset(t.posS, 0x7f, sizeof(Real) * 3 * (size_t)n); memset(t.pvelS, 0x7f, sizeof(Real) * 3 * (size_t)n);
		// MP_MAPPARTS=2: the weights evaluated in the walk (what 2-D grids take anyway); default: per-particle records
		void* rawRec = nullptr;
		if (!(e && atoi(e) == 2)) { if (posix_memalign(&rawRec, 64, sizeof(parts::PartRec<Real>) * (size_t)n)) return MP_ERR_CUDA; memset(rawRec, 0x5a, sizeof(parts::PartRec<Real>) * (size_t)n); }
		t.rec = (parts::PartRec<Real>*)rawRec;
		rc = parts::mapPartsToMAC<Real>(ex, d, vel, velOld, np, ps, pvel, weight, start, key, keyTmp, val, sorted, &t);
		free(rawRec);
		free(t.len1); free(t.off1); free(t.len2); free(t.off2); free(t.e1); free(t.e2); free(t.posS); free(t.pvelS);
	}
	free(start); free(key); free(keyTmp); free(val); free(sorted);
	return rc;
}
template <typename Real> int flipUpdate(int order, const Dims& d, const Real* vel, const Real* velOld, long long np, const Real* pos, const int* pflag, Real* pvel, double flipRatio,
                                        const int* ptype, int exclude) {
	HostExec ex = { order };
	parts::PSet<Real> ps = { pos, pflag, ptype, exclude };
	parts::FlipVelocityUpdate<Real> op = { d, vel, velOld, ps, pvel, (Real)flipRatio, flipRatio < 0 };
	return ex.parts(np, op);
}
template <typename Real> int advect(int order, const Dims& d, const int* flags, const Real* vel, long long np, Real* pos, int* pflag, double dt, int mode, int del, int stop, int skipNew,
                                    const int* ptype, int exclude) {
	HostExec ex = { order };
	parts::AdvectInGrid<Real> op = { d, flags, vel, pos, pflag, ptype, exclude, (Real)dt, mode, del != 0, stop != 0, skipNew != 0 };
	return ex.parts(np, op);
}
}  // namespace

template <typename T> int gridArith(int order, IndexInt n, int comps, void* me, int op, const void* other, double x, double y, double z) {
	// the constants exactly as mp_api.cu::gridArith prepares them
	gridops::Op<T> f;
	f.me = (T*)me; f.other = (const T*)other; f.op = op; f.comps = comps;
	const double c[3] = { x, comps == 3 ? y : x, comps == 3 ? z : x };
	for (int q = 0; q < 3; q++) { f.c0[q] = (T)(op == MP_OP_CLAMP ? x : c[q]); f.c1[q] = (T)y; }
	HostExec ex = { order };
	return ex.parts(n * comps, f);
}
extern "C" {
// elem: 0 int, 4 float, 8 double
int emu_grid_arith(int elem, int order, long long n, int comps, void* me, int op, const void* other, double x, double y, double z) {
	if (elem == 0) return gridArith<int>(order, n, comps, me, op, other, x, y, z);
	if (elem == 4) return gridArith<float>(order, n, comps, me, op, other, x, y, z);
	return gridArith<double>(order, n, comps, me, op, other, x, y, z);
}
int emu_add_force_pvel(int prec, int order, long long np, void* pvel, double ax, double ay, double az, double dt, const int* ptype, int exclude) {
	HostExec ex = { order };
	if (prec == 4) { const float d = (float)dt; parts::AddForcePvel<float> op = { (float*)pvel, { (float)ax * d, (float)ay * d, (float)az * d }, ptype, exclude }; return ex.parts(np, op); }
	parts::AddForcePvel<double> op = { (double*)pvel, { ax * dt, ay * dt, az * dt }, ptype, exclude }; return ex.parts(np, op);
}
int emu_update_velocity_from_delta_pos(int prec, int order, long long np, const void* pos, void* pvel, const void* xPrev, double dt, const int* ptype, int exclude) {
	HostExec ex = { order };
	if (prec == 4) { parts::UpdateVelocityFromDeltaPos<float> op = { (const float*)pos, (float*)pvel, (const float*)xPrev, (float)(1.0 / (double)(float)dt), ptype, exclude }; return ex.parts(np, op); }
	parts::UpdateVelocityFromDeltaPos<double> op = { (const double*)pos, (double*)pvel, (const double*)xPrev, 1.0 / dt, ptype, exclude }; return ex.parts(np, op);
}
int emu_euler_step(int prec, int order, long long np, void* pos, const void* pvel, double dt, const int* ptype, int exclude) {
	HostExec ex = { order };
	if (prec == 4) { parts::StepEuler<float> op = { (float*)pos, (const float*)pvel, (float)dt, ptype, exclude }; return ex.parts(np, op); }
	parts::StepEuler<double> op = { (double*)pos, (const double*)pvel, dt, ptype, exclude }; return ex.parts(np, op);
}
int emu_set_part_type(int prec, int order, int sx, int sy, int sz, long long np, const void* pos, int* ptype, int mark, int stype, const int* flags, int cflag) {
	const Dims d = mkDims(sx, sy, sz); HostExec ex = { order };
	if (prec == 4) { parts::SetPartType<float> op = { d, (const float*)pos, ptype, mark, stype, flags, cflag }; return ex.parts(np, op); }
	parts::SetPartType<double> op = { d, (const double*)pos, ptype, mark, stype, flags, cflag }; return ex.parts(np, op);
}
int emu_mark_isolated_fluid_cell(int prec, int order, int sx, int sy, int sz, int* flags, int mark) {
	(void)prec;
	const Dims d = mkDims(sx, sy, sz); HostExec ex = { order };
	parts::MarkIsolatedFluidCell op = { flags, mark }; return ex.cells(d, op);
}
int emu_push_out_of_obs(int prec, int order, int sx, int sy, int sz, long long np, void* pos, const int* pflag, const void* phiObs, double shift, double thresh, const int* ptype, int exclude) {
	const Dims d = mkDims(sx, sy, sz); HostExec ex = { order };
	if (prec == 4) { parts::PSet<float> ps = { (const float*)pos, pflag, ptype, exclude }; parts::PushOutOfObs<float> op = { d, (float*)pos, ps, (const float*)phiObs, (float)shift, (float)thresh }; return ex.parts(np, op); }
	parts::PSet<double> ps = { (const double*)pos, pflag, ptype, exclude }; parts::PushOutOfObs<double> op = { d, (double*)pos, ps, (const double*)phiObs, shift, thresh }; return ex.parts(np, op);
}
int emu_project_out_of_bnd(int prec, int order, int sx, int sy, int sz, long long np, void* pos, const int* pflag, double bnd, int axis, const int* ptype, int exclude) {
	const Dims d = mkDims(sx, sy, sz); HostExec ex = { order };
	if (prec == 4) { parts::PSet<float> ps = { (const float*)pos, pflag, ptype, exclude }; parts::ProjectOutOfBnd<float> op = { d, (float*)pos, ps, (float)bnd, axis }; return ex.parts(np, op); }
	parts::PSet<double> ps = { (const double*)pos, pflag, ptype, exclude }; parts::ProjectOutOfBnd<double> op = { d, (double*)pos, ps, bnd, axis }; return ex.parts(np, op);
}
int emu_advect_in_grid(int prec, int order, int sx, int sy, int sz, const int* flags, const void* vel, long long np, void* pos, int* pflag, double dt, int mode, int del, int stop,
                       int skipNew, const int* ptype, int exclude) {
	const Dims d = mkDims(sx, sy, sz);
	return prec == 4 ? advect<float>(order, d, flags, (const float*)vel, np, (float*)pos, pflag, dt, mode, del, stop, skipNew, ptype, exclude)
	                 : advect<double>(order, d, flags, (const double*)vel, np, (double*)pos, pflag, dt, mode, del, stop, skipNew, ptype, exclude);
}
// signatures follow the oracle's mfo_* entry points (oracle/mf_oracle.c), with (prec, order) in front
int emu_mark_fluid_cells(int prec, int order, int sx, int sy, int sz, int* flags, long long np, const void* pos, const int* pflag, const void* phiObs, const int* ptype, int exclude) {
	const Dims d = mkDims(sx, sy, sz);
	return prec == 4 ? markFluid<float>(order, d, flags, np, (const float*)pos, pflag, (const float*)phiObs, ptype, exclude)
	                 : markFluid<double>(order, d, flags, np, (const double*)pos, pflag, (const double*)phiObs, ptype, exclude);
}
int emu_grid_particle_index(int prec, int order, int sx, int sy, int sz, long long np, const void* pos, const int* pflag, int* index, int* indexSys, long long* count) {
	const Dims d = mkDims(sx, sy, sz);
	return prec == 4 ? particleIndex<float>(order, d, np, (const float*)pos, pflag, index, indexSys, count)
	                 : particleIndex<double>(order, d, np, (const double*)pos, pflag, index, indexSys, count);
}
int emu_union_particle_levelset(int prec, int order, int sx, int sy, int sz, long long np, const void* pos, const int* index, const int* indexSys, long long count, void* phi,
                                double radiusFactor, const int* ptype, int exclude) {
	(void)np;
	const Dims d = mkDims(sx, sy, sz); HostExec ex = { order };
	// positions in index order, as mp_particles.cu runs it (MP_UNION_SORTED=0: through indexSys); the scratch comes uninitialised
	const char* e = getenv("MP_UNION_SORTED");
	void* scratch = ((!e || atoi(e)) && count > 0) ? malloc((size_t)prec * 3 * (size_t)count) : nullptr;
	if (scratch) memset(scratch, 0x7f, (size_t)prec * 3 * (size_t)count);
	const int rc = prec == 4 ? parts::unionParticleLevelset<float>(ex, d, (const float*)pos, index, indexSys, count, (float*)phi, radiusFactor, ptype, exclude, (float*)scratch)
	                         : parts::unionParticleLevelset<double>(ex, d, (const double*)pos, index, indexSys, count, (double*)phi, radiusFactor, ptype, exclude, (double*)scratch);
	free(scratch);
	return rc;
}
int emu_map_parts_to_mac(int prec, int order, int sx, int sy, int sz, void* vel, void* velOld, long long np, const void* pos, const int* pflag, const void* pvel, void* weight,
                         const int* ptype, int exclude) {
	const Dims d = mkDims(sx, sy, sz);
	return prec == 4 ? mapParts<float>(order, d, (float*)vel, (float*)velOld, np, (const float*)pos, pflag, (const float*)pvel, (float*)weight, ptype, exclude)
	                 : mapParts<double>(order, d, (double*)vel, (double*)velOld, np, (const double*)pos, pflag, (const double*)pvel, (double*)weight, ptype, exclude);
}
int emu_flip_velocity_update(int prec, int order, int sx, int sy, int sz, const void* vel, const void* velOld, long long np, const void* pos, const int* pflag, void* pvel,
                             double flipRatio, const int* ptype, int exclude) {
	const Dims d = mkDims(sx, sy, sz);
	return prec == 4 ? flipUpdate<float>(order, d, (const float*)vel, (const float*)velOld, np, (const float*)pos, pflag, (float*)pvel, flipRatio, ptype, exclude)
	                 : flipUpdate<double>(order, d, (const double*)vel, (const double*)velOld, np, (const double*)pos, pflag, (double*)pvel, flipRatio, ptype, exclude);
}
}

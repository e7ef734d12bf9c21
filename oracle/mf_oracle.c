/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of mantaflow's pressure-projection path.
 *
 * Nothing here is linked, imported or executed by the product path (mantaflow_b200/); only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker.  It is a from-scratch plain-C restatement of the reference algorithm,
 * each function citing the reference file:line (relative to the mantaflow tree) it follows.
 *
 * PARITY PIN: validated against the UNMODIFIED reference compiled from /root/reference
 * (oracle/_ref/libmanta_ref_f{32,64}.so, see oracle/Makefile + oracle/ref_harness.cpp) by
 * tests/test_oracle_vs_reference.py, and against golden vectors generated from that reference
 * (tests/golden/, generator tests/golden/make_golden.py).  The reference ships no golden data of
 * its own (tools/testdata/readme.txt:1).
 *
 * Precision: compiled twice, Real=float (reference default build "fp1") and Real=double ("fp2",
 * -DDOUBLEPRECISION=ON).  Compiled with -ffp-contract=off: the reference build has no FMA.
 *
 * Layout (grid.h:70): idx = i + sx*(j + sy*k), x fastest; flags int32; MAC velocity AoS {x,y,z}.
 * 2-D grids have sz==1 and Z-stride 0 (grid.cpp:55).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdint.h>

#if MF_REAL_IS_DOUBLE
typedef double Real;
#define R_SQRT sqrt
#define R_FABS fabs
#else
typedef float Real;
#define R_SQRT sqrtf
#define R_FABS fabsf
#endif

typedef long long IndexInt;

enum { TypeFluid = 1, TypeObstacle = 2, TypeEmpty = 4, TypeInflow = 8, TypeOutflow = 16, TypeOpen = 32, TypeStick = 64 }; /* grid.h:292-304 */

static char g_err[512];
const char* mfo_last_error(void) { return g_err; }
int mfo_real_size(void) { return (int)sizeof(Real); }
int mfo_is_reference(void) { return 0; }
int mfo_set_debug_level(int l) { (void)l; return 0; }

#define IS3D (sz > 1)
#define IDX(i, j, k) ((IndexInt)(i) + (IndexInt)sx * (j) + (IndexInt)SZ * (k))
#define STRIDES const IndexInt X = 1, Y = sx, SZ = (sz > 1) ? (IndexInt)sx * sy : 0, Z = SZ; (void)X; (void)Y; (void)Z;

static inline Real rmin(Real a, Real b) { return a < b ? a : b; }
#define REAL_MAX_ (MF_REAL_IS_DOUBLE ? (Real)1.7976931348623157e308 : (Real)3.402823466e38f)
static inline Real rmax(Real a, Real b) { return a > b ? a : b; }

/* ---------------------------------------------------------------------------------------------
 * setWallBcs without obvel / fractions: plugin/extforces.cpp:186-218 (KnSetWallBcs), :307-316   */
int mfo_set_wall_bcs(int sx, int sy, int sz, const int* flags, Real* vel)
{
	STRIDES
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		const IndexInt idx = IDX(i, j, k);
		const int curFluid = flags[idx] & TypeFluid, curObs = flags[idx] & TypeObstacle;
		if (!curFluid && !curObs) continue;
		Real* v = vel + 3 * idx;
		if (i > 0 && (flags[idx - X] & TypeObstacle)) v[0] = 0;
		if (i > 0 && curObs && (flags[idx - X] & TypeFluid)) v[0] = 0;
		if (j > 0 && (flags[idx - Y] & TypeObstacle)) v[1] = 0;
		if (j > 0 && curObs && (flags[idx - Y] & TypeFluid)) v[1] = 0;
		if (!IS3D) { v[2] = 0; } else {
			if (k > 0 && (flags[idx - Z] & TypeObstacle)) v[2] = 0;
			if (k > 0 && curObs && (flags[idx - Z] & TypeFluid)) v[2] = 0;
		}
		if (curFluid) {
			if ((i > 0 && (flags[idx - X] & TypeStick)) || (i < sx - 1 && (flags[idx + X] & TypeStick))) v[1] = v[2] = 0;
			if ((j > 0 && (flags[idx - Y] & TypeStick)) || (j < sy - 1 && (flags[idx + Y] & TypeStick))) v[0] = v[2] = 0;
			if (IS3D && ((k > 0 && (flags[idx - Z] & TypeStick)) || (k < sz - 1 && (flags[idx + Z] & TypeStick)))) v[0] = v[1] = 0;
		}
	}
	return 0;
}

/* ---------------------------------------------------------------------------------------------
 * ghost-fluid helpers: plugin/pressure.cpp:115-133                                              */
static inline Real thetaHelper(Real inside, Real outside)
{
	const Real denom = inside - outside;
	if ((double)denom > -1e-04) return (Real)0.5;           /* Real compared with a double literal (:118) */
	return rmax((Real)0, rmin((Real)1, inside / denom));
}
static inline Real ghostFluidHelper(IndexInt idx, IndexInt offset, const Real* phi, Real gfClamp)
{
	Real alpha = thetaHelper(phi[idx], phi[idx + offset]);
	if (alpha < gfClamp) return gfClamp;
	return (Real)(1. - (1. / (double)alpha));                /* evaluated in double, then narrowed (:127) */
}
static inline Real surfTensHelper(IndexInt idx, IndexInt offset, const Real* phi, const Real* curv, Real surfTens, Real gfClamp)
{
	return surfTens * (curv[idx + offset] - ghostFluidHelper(idx, offset, phi, gfClamp) * curv[idx]);
}
static inline int ghostFluidWasClamped(IndexInt idx, IndexInt offset, const Real* phi, Real gfClamp)
{
	return thetaHelper(phi[idx], phi[idx + offset]) < gfClamp;   /* :191-196 */
}

/* ---------------------------------------------------------------------------------------------
 * computePressureRhs: plugin/pressure.cpp:277-299, kernel MakeRhs :32-84 (bnd=1)                */
int mfo_compute_rhs(int sx, int sy, int sz, const int* flags, const Real* vel, Real* rhs,
	const Real* phi, const Real* perCellCorr, const Real* fractions, const Real* obvel, const Real* curv,
	double gfClamp_, double surfTens_, int enforceCompatibility, double* sum_out, int* cnt_out)
{
	STRIDES
	const Real gfClamp = (Real)gfClamp_, surfTens = (Real)surfTens_;
	double sum = 0; int cnt = 0;
	const int k0 = IS3D ? 1 : 0, k1 = IS3D ? sz - 1 : 1;
	for (int k = k0; k < k1; k++) for (int j = 1; j < sy - 1; j++) for (int i = 1; i < sx - 1; i++) {
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeFluid)) { rhs[idx] = 0; continue; }
		const Real* v = vel + 3 * idx; const Real* vx = vel + 3 * (idx + X); const Real* vy = vel + 3 * (idx + Y); const Real* vz = vel + 3 * (idx + Z);
		Real set = 0;
		if (!fractions) {
			set = v[0] - vx[0] + v[1] - vy[1];
			if (IS3D) set += v[2] - vz[2];
		} else {
			const Real* f = fractions + 3 * idx; const Real* fx = fractions + 3 * (idx + X); const Real* fy = fractions + 3 * (idx + Y); const Real* fz = fractions + 3 * (idx + Z);
			set = f[0] * v[0] - fx[0] * vx[0] + f[1] * v[1] - fy[1] * vy[1];
			if (IS3D) set += f[2] * v[2] - fz[2] * vz[2];
			if (obvel) {
				const Real* o = obvel + 3 * idx; const Real* ox = obvel + 3 * (idx + X); const Real* oy = obvel + 3 * (idx + Y); const Real* oz = obvel + 3 * (idx + Z);
				set += (1 - f[0]) * o[0] - (1 - fx[0]) * ox[0] + (1 - f[1]) * o[1] - (1 - fy[1]) * oy[1];
				if (IS3D) set += (1 - f[2]) * o[2] - (1 - fz[2]) * oz[2];
			}
		}
		if (phi && curv) {
			if (flags[idx - X] & TypeEmpty) set += surfTensHelper(idx, -X, phi, curv, surfTens, gfClamp);
			if (flags[idx + X] & TypeEmpty) set += surfTensHelper(idx, +X, phi, curv, surfTens, gfClamp);
			if (flags[idx - Y] & TypeEmpty) set += surfTensHelper(idx, -Y, phi, curv, surfTens, gfClamp);
			if (flags[idx + Y] & TypeEmpty) set += surfTensHelper(idx, +Y, phi, curv, surfTens, gfClamp);
			if (IS3D) {
				if (flags[idx - Z] & TypeEmpty) set += surfTensHelper(idx, -Z, phi, curv, surfTens, gfClamp);
				if (flags[idx + Z] & TypeEmpty) set += surfTensHelper(idx, +Z, phi, curv, surfTens, gfClamp);
			}
		}
		if (perCellCorr) set += perCellCorr[idx];
		sum += set; cnt++;
		rhs[idx] = set;
	}
	if (enforceCompatibility) {                               /* :297-298, applied to ALL cells */
		const Real corr = (Real)(-sum / (Real)cnt);
		const IndexInt n = (IndexInt)sx * sy * sz;
		for (IndexInt q = 0; q < n; q++) rhs[q] += corr;
	}
	if (sum_out) *sum_out = sum;
	if (cnt_out) *cnt_out = cnt;
	return 0;
}

/* ---------------------------------------------------------------------------------------------
 * MakeLaplaceMatrix conjugategrad.h:154-187 (+ ApplyGhostFluidDiagonal pressure.cpp:136-151
 * when phi != NULL).  A* must be zero on entry (the reference allocates cleared grids).          */
int mfo_make_matrix(int sx, int sy, int sz, const int* flags, const Real* fractions, const Real* phi, double gfClamp_,
	Real* A0, Real* Ai, Real* Aj, Real* Ak)
{
	STRIDES
	const Real gfClamp = (Real)gfClamp_;
	const int k0 = IS3D ? 1 : 0, k1 = IS3D ? sz - 1 : 1;
	#pragma omp parallel for schedule(static)
	for (int k = k0; k < k1; k++) for (int j = 1; j < sy - 1; j++) for (int i = 1; i < sx - 1; i++) {
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeFluid)) continue;
		if (!fractions) {
			if (!(flags[idx - X] & TypeObstacle)) A0[idx] += 1.;
			if (!(flags[idx + X] & TypeObstacle)) A0[idx] += 1.;
			if (!(flags[idx - Y] & TypeObstacle)) A0[idx] += 1.;
			if (!(flags[idx + Y] & TypeObstacle)) A0[idx] += 1.;
			if (IS3D && !(flags[idx - Z] & TypeObstacle)) A0[idx] += 1.;
			if (IS3D && !(flags[idx + Z] & TypeObstacle)) A0[idx] += 1.;
			if (flags[idx + X] & TypeFluid) Ai[idx] = -1.;
			if (flags[idx + Y] & TypeFluid) Aj[idx] = -1.;
			if (IS3D && (flags[idx + Z] & TypeFluid)) Ak[idx] = -1.;
		} else {
			A0[idx] += fractions[3 * idx + 0];
			A0[idx] += fractions[3 * (idx + X) + 0];
			A0[idx] += fractions[3 * idx + 1];
			A0[idx] += fractions[3 * (idx + Y) + 1];
			if (IS3D) A0[idx] += fractions[3 * idx + 2];
			if (IS3D) A0[idx] += fractions[3 * (idx + Z) + 2];
			if (flags[idx + X] & TypeFluid) Ai[idx] = -fractions[3 * (idx + X) + 0];
			if (flags[idx + Y] & TypeFluid) Aj[idx] = -fractions[3 * (idx + Y) + 1];
			if (IS3D && (flags[idx + Z] & TypeFluid)) Ak[idx] = -fractions[3 * (idx + Z) + 2];
		}
		if (phi) {
			if (flags[idx - X] & TypeEmpty) A0[idx] -= ghostFluidHelper(idx, -X, phi, gfClamp);
			if (flags[idx + X] & TypeEmpty) A0[idx] -= ghostFluidHelper(idx, +X, phi, gfClamp);
			if (flags[idx - Y] & TypeEmpty) A0[idx] -= ghostFluidHelper(idx, -Y, phi, gfClamp);
			if (flags[idx + Y] & TypeEmpty) A0[idx] -= ghostFluidHelper(idx, +Y, phi, gfClamp);
			if (IS3D) {
				if (flags[idx - Z] & TypeEmpty) A0[idx] -= ghostFluidHelper(idx, -Z, phi, gfClamp);
				if (flags[idx + Z] & TypeEmpty) A0[idx] -= ghostFluidHelper(idx, +Z, phi, gfClamp);
			}
		}
	}
	return 0;
}

/* ---------------------------------------------------------------------------------------------
 * pressure pinning: cell choice plugin/pressure.cpp:349-382 (CountEmptyCells :217-220)          */
long long mfo_choose_fix_cell(int sx, int sy, int sz, const int* flags)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	for (IndexInt q = 0; q < n; q++) if (flags[q] & TypeEmpty) return -1;
	const int cx = sx / 2, cz = IS3D ? sz / 2 : 0;
	for (int d = 0; d < 3; d++) {                          /* top centre, then one and two cells below */
		const int cy = sy - 1 - d;
		if (cy < 0) continue;
		const IndexInt idx = IDX(cx, cy, cz);
		if (flags[idx] & TypeFluid) return idx;
	}
	const int k0 = IS3D ? 1 : 0, k1 = IS3D ? sz - 1 : 1;
	for (int k = k0; k < k1; k++) for (int j = 1; j < sy - 1; j++) for (int i = 1; i < sx - 1; i++)
		if (flags[IDX(i, j, k)] & TypeFluid) return IDX(i, j, k);
	return -1;
}

/* fixPressure plugin/pressure.cpp:226-245 */
int mfo_fix_pressure(int sx, int sy, int sz, long long p, double value_, Real* rhs, Real* A0, Real* Ai, Real* Aj, Real* Ak)
{
	STRIDES
	const Real value = (Real)value_;
	rhs[p + X] -= Ai[p] * value;
	rhs[p + Y] -= Aj[p] * value;
	rhs[p - X] -= Ai[p - X] * value;
	rhs[p - Y] -= Aj[p - Y] * value;
	if (IS3D) { rhs[p + Z] -= Ak[p] * value; rhs[p - Z] -= Ak[p - Z] * value; }
	rhs[p] = value;
	A0[p] = (Real)1;
	Ai[p] = Aj[p] = Ak[p] = (Real)0;
	Ai[p - X] = (Real)0;
	Aj[p - Y] = (Real)0;
	if (IS3D) Ak[p - Z] = (Real)0;
	return 0;
}

/* ---------------------------------------------------------------------------------------------
 * ApplyMatrix / ApplyMatrix2D conjugategrad.h:118-151 (idx mode over ALL cells)                 */
static void apply_matrix(int sx, int sy, int sz, const int* flags, Real* dst, const Real* src,
	const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	if (IS3D) {
		#pragma omp parallel for schedule(static)
		for (IndexInt idx = 0; idx < n; idx++) {
			if (!(flags[idx] & TypeFluid)) { dst[idx] = src[idx]; continue; }
			dst[idx] = src[idx] * A0[idx]
				+ src[idx - X] * Ai[idx - X] + src[idx + X] * Ai[idx]
				+ src[idx - Y] * Aj[idx - Y] + src[idx + Y] * Aj[idx]
				+ src[idx - Z] * Ak[idx - Z] + src[idx + Z] * Ak[idx];
		}
	} else {
		#pragma omp parallel for schedule(static)
		for (IndexInt idx = 0; idx < n; idx++) {
			if (!(flags[idx] & TypeFluid)) { dst[idx] = src[idx]; continue; }
			dst[idx] = src[idx] * A0[idx]
				+ src[idx - X] * Ai[idx - X] + src[idx + X] * Ai[idx]
				+ src[idx - Y] * Aj[idx - Y] + src[idx + Y] * Aj[idx];
		}
	}
}
int mfo_apply_matrix(int sx, int sy, int sz, const int* flags, Real* dst, const Real* src,
	const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{ apply_matrix(sx, sy, sz, flags, dst, src, A0, Ai, Aj, Ak); return 0; }

/* deterministic blocked reductions (the reference's OpenMP/TBB join order is unspecified,
 * codegen_kernel.cpp:204-210,:462; accumulators are double as in the reference)              */
#define RBLK 8192
static double dot_product(const Real* a, const Real* b, IndexInt n)   /* GridDotProduct conjugategrad.cpp:175-178 */
{
	const IndexInt nb = (n + RBLK - 1) / RBLK;
	double* part = (double*)malloc(sizeof(double) * (size_t)nb);
	#pragma omp parallel for schedule(static)
	for (IndexInt b_ = 0; b_ < nb; b_++) {
		double s = 0; const IndexInt e = (b_ + 1) * RBLK < n ? (b_ + 1) * RBLK : n;
		for (IndexInt q = b_ * RBLK; q < e; q++) s += (a[q] * b[q]);     /* product in Real, sum in double */
		part[b_] = s;
	}
	double s = 0; for (IndexInt b_ = 0; b_ < nb; b_++) s += part[b_];
	free(part); return s;
}
static double sum_sqr(const Real* a, IndexInt n)                       /* GridSumSqr commonkernels.h:32-35 */
{
	const IndexInt nb = (n + RBLK - 1) / RBLK;
	double* part = (double*)malloc(sizeof(double) * (size_t)nb);
	#pragma omp parallel for schedule(static)
	for (IndexInt b_ = 0; b_ < nb; b_++) {
		double s = 0; const IndexInt e = (b_ + 1) * RBLK < n ? (b_ + 1) * RBLK : n;
		for (IndexInt q = b_ * RBLK; q < e; q++) s += (double)a[q] * (double)a[q];
		part[b_] = s;
	}
	double s = 0; for (IndexInt b_ = 0; b_ < nb; b_++) s += part[b_];
	free(part); return s;
}
static Real max_abs(const Real* a, IndexInt n)                         /* Grid<Real>::getMaxAbs grid.cpp:319-323 */
{
	Real m = 0;
	#pragma omp parallel for reduction(max:m) schedule(static)
	for (IndexInt q = 0; q < n; q++) { Real v = R_FABS(a[q]); if (v > m) m = v; }
	return m;
}
static void scaled_add(Real* me, const Real* other, Real f, IndexInt n) /* gridScaledAdd grid.h:478 */
{
	#pragma omp parallel for schedule(static)
	for (IndexInt q = 0; q < n; q++) me[q] += f * other[q];
}

/* ---------------------------------------------------------------------------------------------
 * MIC(0): InitPreconditionModifiedIncompCholesky2 conjugategrad.cpp:66-97 (serial, k/j/i order) */
static void mic_init(int sx, int sy, int sz, const int* flags, Real* P, const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	memset(P, 0, sizeof(Real) * (size_t)n);
	const Real tau = (Real)0.97, sigma = (Real)0.25;
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeFluid)) continue;
		const IndexInt ix = idx - X, iy = idx - Y, iz = idx - Z;
		const Real tx = Ai[ix] * P[ix], ty = Aj[iy] * P[iy], tz = Ak[iz] * P[iz];
		Real e = A0[idx] - tx * tx - ty * ty - tz * tz;
		/* the "+ 0." in the reference promotes the bracket, the product with tau and the
		   subtraction to double before narrowing back to Real (:85-89) */
		const Real inner = Ai[ix] * (Aj[ix] + Ak[ix]) * (P[ix] * P[ix])
		                 + Aj[iy] * (Ai[iy] + Ak[iy]) * (P[iy] * P[iy])
		                 + Ak[iz] * (Ai[iz] + Aj[iz]) * (P[iz] * P[iz]);
		e = (Real)((double)e - (double)tau * ((double)inner + 0.));
		if (e < sigma * A0[idx]) e = A0[idx];
		P[idx] = (Real)(1. / (double)R_SQRT(e));
	}
}
int mfo_mic_init(int sx, int sy, int sz, const int* flags, Real* P, const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{ mic_init(sx, sy, sz, flags, P, A0, Ai, Aj, Ak); return 0; }

/* ApplyPreconditionModifiedIncompCholesky2 conjugategrad.cpp:135-159 */
static void mic_apply(int sx, int sy, int sz, const int* flags, Real* dst, const Real* src, const Real* P,
	const Real* Ai, const Real* Aj, const Real* Ak)
{
	STRIDES
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeFluid)) continue;
		const Real p = P[idx];
		dst[idx] = p * (src[idx]
			- dst[idx - X] * Ai[idx - X] * P[idx - X]
			- dst[idx - Y] * Aj[idx - Y] * P[idx - Y]
			- dst[idx - Z] * Ak[idx - Z] * P[idx - Z]);
	}
	for (int k = sz - 1; k >= 0; k--) for (int j = sy - 1; j >= 0; j--) for (int i = sx - 1; i >= 0; i--) {
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeFluid)) continue;
		const Real p = P[idx];
		dst[idx] = p * (dst[idx]
			- dst[idx + X] * Ai[idx] * p
			- dst[idx + Y] * Aj[idx] * p
			- dst[idx + Z] * Ak[idx] * p);
	}
}
int mfo_mic_apply(int sx, int sy, int sz, const int* flags, Real* dst, const Real* src, const Real* P,
	const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{ (void)A0; mic_apply(sx, sy, sz, flags, dst, src, P, Ai, Aj, Ak); return 0; }

/* ---------------------------------------------------------------------------------------------
 * MIC(0) in BLOCK RED-BLACK ordering -- the reformulated preconditioner north_star (4) allows ("level-scheduled or colored triangular
 * solves", iteration counts reported).  NO REFERENCE COUNTERPART: this is the specification of the device kernels k_micrb<MODE>
 * (csrc/mp_micrb.cu), pinned in two ways: (a) with one tile covering the grid the ordering is the reference's and factor and sweeps equal
 * mic_init / mic_apply (conjugategrad.cpp:66-97, :135-159) bit for bit; (b) the factor of a small grid is checked against a dense
 * incomplete Cholesky factorization of the permuted matrix (tests/test_micrb.py).
 * Ordering: the grid is cut into tiles of T0 x T1 x T2 cells; tiles with even (tx+ty+tz) come first ("red"), then the odd ones ("black"),
 * cells in lexicographic order inside a tile.  A 7-point stencil couples tiles through faces only, so tiles of one colour are independent.
 * For a cell c and a face neighbour n: n precedes c iff (same tile and n is lexicographically earlier) or (other tile and n is red).
 *   e_c = A_cc - sum_{n prec c} (A_cn P_n)^2 - tau * sum_{n prec c} A_cn * (sum_{s succ n, s != c} A_ns) * P_n^2,  P_c = 1/sqrt(e_c)
 *   forward  z_c = P_c (r_c - sum_{n prec c} z_n A_cn P_n),  backward  z_c = P_c (z_c - sum_{s succ c} z_s A_cs P_c)
 * with the reference's safety rule (e < sigma A_cc -> e = A_cc) and its mixed Real / double roundings.  Order of the terms: first the
 * neighbours in OTHER tiles (direction order -x,-y,-z,+x,+y,+z), then the neighbours in the cell's own tile in the reference's order
 * (-x,-y,-z forward and in the factor, +x,+y,+z backward) -- with one tile that is the reference's expression.                        */
static int g_mic_tile[3] = { 1 << 30, 1 << 30, 1 << 30 };
int mfo_set_mic_tiles(int t0, int t1, int t2) { g_mic_tile[0] = t0 > 0 ? t0 : 1 << 30; g_mic_tile[1] = t1 > 0 ? t1 : 1 << 30; g_mic_tile[2] = t2 > 0 ? t2 : 1 << 30; return 0; }
static inline int rb_colour(int i, int j, int k) { return (i / g_mic_tile[0] + j / g_mic_tile[1] + k / g_mic_tile[2]) & 1; }
/* direction q (0..5 = -x,-y,-z,+x,+y,+z) from cell (i,j,k): does it leave the cell's tile? */
static inline int rb_leaves(int i, int j, int k, int q)
{
	const int c[3] = { i, j, k }; const int a = q % 3;
	return c[a] / g_mic_tile[a] != (c[a] + (q >= 3 ? 1 : -1)) / g_mic_tile[a];
}
/* does the neighbour of cell (i,j,k) in direction q precede it?  (the neighbour must lie in the grid) */
static inline int rb_prec(int i, int j, int k, int q)
{
	if (!rb_leaves(i, j, k, q)) return q < 3;
	return rb_colour(i, j, k) == 1;                             /* the neighbour's tile has the other colour; red comes first */
}
static const int rb_di[6] = { -1, 0, 0, 1, 0, 0 }, rb_dj[6] = { 0, -1, 0, 0, 1, 0 }, rb_dk[6] = { 0, 0, -1, 0, 0, 1 };
static void micrb_init(int sx, int sy, int sz, const int* flags, Real* P, const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	memset(P, 0, sizeof(Real) * (size_t)n);
	const Real tau = (Real)0.97, sigma = (Real)0.25;
	const IndexInt off[6] = { -X, -Y, -Z, X, Y, Z };
	const Real* Aq[3] = { Ai, Aj, Ak };
	for (int colour = 0; colour < 2; colour++)
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		if (rb_colour(i, j, k) != colour) continue;
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeFluid)) continue;
		/* fluid cells are interior cells (mp_check_flags_interior / the reference's own assumption): all six neighbours exist */
		Real e = A0[idx], inner = 0; int first = 1;
		for (int pass = 0; pass < 2; pass++) for (int q = 0; q < 6; q++) {          /* pass 0: other tiles, pass 1: own tile */
			if (rb_leaves(i, j, k, q) != (pass == 0) || !rb_prec(i, j, k, q)) continue;
			const IndexInt nb = idx + off[q];
			const Real a = Aq[q % 3][q < 3 ? nb : idx];                         /* the coupling of c and nb is stored at the lower cell */
			const Real t = a * P[nb];
			e = e - t * t;
			/* couplings of nb to its successors other than c (sums of the entries in direction order) */
			const int ni = i + rb_di[q], nj = j + rb_dj[q], nk = k + rb_dk[q];
			Real ssum = 0; int sfirst = 1;
			for (int r = 0; r < 6; r++) {
				if (r == (q + 3) % 6) continue;                                  /* that is c */
				const int si = ni + rb_di[r], sj = nj + rb_dj[r], sk = nk + rb_dk[r];
				if (si < 0 || sj < 0 || sk < 0 || si >= sx || sj >= sy || sk >= sz) continue;
				if (rb_prec(ni, nj, nk, r)) continue;                            /* a predecessor of nb */
				const IndexInt sidx = nb + off[r];
				const Real an = Aq[r % 3][r < 3 ? sidx : nb];
				if (sfirst) { ssum = an; sfirst = 0; } else ssum = ssum + an;
			}
			const Real term = a * ssum * (P[nb] * P[nb]);
			if (first) { inner = term; first = 0; } else inner = inner + term;
		}
		e = (Real)((double)e - (double)tau * ((double)inner + 0.));
		if (e < sigma * A0[idx]) e = A0[idx];
		P[idx] = (Real)(1. / (double)R_SQRT(e));
	}
}
static void micrb_apply(int sx, int sy, int sz, const int* flags, Real* dst, const Real* src, const Real* P,
	const Real* Ai, const Real* Aj, const Real* Ak)
{
	STRIDES
	const IndexInt off[6] = { -X, -Y, -Z, X, Y, Z };
	const Real* Aq[3] = { Ai, Aj, Ak };
	for (int colour = 0; colour < 2; colour++)
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		if (rb_colour(i, j, k) != colour) continue;
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeFluid)) continue;
		Real acc = src[idx];
		for (int pass = 0; pass < 2; pass++) for (int q = 0; q < 6; q++) {
			if (rb_leaves(i, j, k, q) != (pass == 0) || !rb_prec(i, j, k, q)) continue;
			const IndexInt nb = idx + off[q];
			acc = acc - dst[nb] * Aq[q % 3][q < 3 ? nb : idx] * P[nb];
		}
		dst[idx] = P[idx] * acc;
	}
	for (int colour = 1; colour >= 0; colour--)
	for (int k = sz - 1; k >= 0; k--) for (int j = sy - 1; j >= 0; j--) for (int i = sx - 1; i >= 0; i--) {
		if (rb_colour(i, j, k) != colour) continue;
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeFluid)) continue;
		const Real p = P[idx];
		Real acc = dst[idx];
		for (int pass = 0; pass < 2; pass++) for (int qq = 0; qq < 6; qq++) {
			const int q = pass == 0 ? qq : (qq + 3) % 6;                         /* own tile: +x, +y, +z (the reference's order) */
			if (rb_leaves(i, j, k, q) != (pass == 0) || rb_prec(i, j, k, q)) continue;
			const IndexInt nb = idx + off[q];
			acc = acc - dst[nb] * Aq[q % 3][q < 3 ? nb : idx] * p;
		}
		dst[idx] = p * acc;
	}
}
int mfo_micrb_init(int sx, int sy, int sz, const int* flags, Real* P, const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{ micrb_init(sx, sy, sz, flags, P, A0, Ai, Aj, Ak); return 0; }
int mfo_micrb_apply(int sx, int sy, int sz, const int* flags, Real* dst, const Real* src, const Real* P,
	const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{ (void)A0; micrb_apply(sx, sy, sz, flags, dst, src, P, Ai, Aj, Ak); return 0; }

/* =============================================================================================
 * GridMg  (multigrid.h:31-137, multigrid.cpp)                                                   */
enum { vtInactive = 0, vtActive = 1, vtActiveTrivial = 2, vtRemoved = 3, vtZero = 4, vtFree = 5 };  /* multigrid.h:87-94 */
#define MG_MAXLVL 32

typedef struct { int Ux, Uy, Uz, Wx, Wy, Wz, Nx, Ny, Nz, sc, sf, inU; Real rw, iw; } CoarseningPath;

typedef struct MgState {
	int is3D, dim, stencil, stencil0, nlev;
	int smin[3], smax[3];
	int size[MG_MAXLVL][3]; int pitch[MG_MAXLVL][3]; int n[MG_MAXLVL];
	Real* A[MG_MAXLVL]; Real* x[MG_MAXLVL]; Real* b[MG_MAXLVL]; Real* r[MG_MAXLVL]; signed char* type[MG_MAXLVL];
	double *cg1, *cg2, *cg3, *cg4;
	CoarseningPath* paths; int npaths;
	int numPre, numPost; Real coarsestAcc, trivialScale;
	int isASet, isRhsSet;
	int lastCoarseIters;
} MgState;

static int path_less(const void* a_, const void* b_)   /* multigrid.cpp:314-318 (key only; ties keep generation order) */
{
	const CoarseningPath* a = (const CoarseningPath*)a_; const CoarseningPath* b = (const CoarseningPath*)b_;
	if (a->sc != b->sc) return a->sc < b->sc ? -1 : 1;
	const int ka = (a->Ux + 1) + 3 * (a->Uy + 1) + 9 * (a->Uz + 1), kb = (b->Ux + 1) + 3 * (b->Uy + 1) + 9 * (b->Uz + 1);
	if (ka != kb) return ka < kb ? -1 : 1;
	return 0;
}

static MgState* mg_create(int sx, int sy, int sz)       /* GridMg::GridMg multigrid.cpp:220-319 */
{
	MgState* m = (MgState*)calloc(1, sizeof(MgState));
	m->numPre = m->numPost = 1; m->coarsestAcc = (Real)1E-8; m->trivialScale = (Real)1E-6;
	m->is3D = sz > 1; m->dim = m->is3D ? 3 : 2;
	m->stencil = m->is3D ? 14 : 5; m->stencil0 = m->is3D ? 4 : 3;
	m->smin[0] = m->smin[1] = -1; m->smin[2] = m->is3D ? -1 : 0;
	m->smax[0] = m->smax[1] = 1;  m->smax[2] = m->is3D ? 1 : 0;
	int l = 0;
	m->size[0][0] = sx; m->size[0][1] = sy; m->size[0][2] = sz;
	for (;;) {
		const int* s = m->size[l];
		m->pitch[l][0] = 1; m->pitch[l][1] = s[0]; m->pitch[l][2] = s[0] * s[1];
		m->n[l] = s[0] * s[1] * s[2];
		const int st = (l == 0) ? m->stencil0 : m->stencil;
		m->A[l] = (Real*)calloc((size_t)m->n[l] * st, sizeof(Real));
		m->x[l] = (Real*)calloc((size_t)m->n[l], sizeof(Real));
		m->b[l] = (Real*)calloc((size_t)m->n[l], sizeof(Real));
		m->r[l] = (Real*)calloc((size_t)m->n[l], sizeof(Real));
		m->type[l] = (signed char*)calloc((size_t)m->n[l], 1);
		m->nlev = l + 1;
		if (l + 1 > 100) break;
		if (s[0] <= 5 && s[1] <= 5 && s[2] <= 5) break;
		if (m->n[l] <= 1000) break;
		for (int d = 0; d < 3; d++) m->size[l + 1][d] = (s[d] + 2) / 2;
		l++;
	}
	const int nc = m->n[m->nlev - 1];
	m->cg1 = (double*)calloc((size_t)nc, sizeof(double)); m->cg2 = (double*)calloc((size_t)nc, sizeof(double));
	m->cg3 = (double*)calloc((size_t)nc, sizeof(double)); m->cg4 = (double*)calloc((size_t)nc, sizeof(double));

	/* coarsening paths (V)<-R-(U)<-A-(W)<-I-(N) for level 1, multigrid.cpp:286-318 */
	static const int p7[7][3] = { {0,0,0}, {-1,0,0}, {1,0,0}, {0,-1,0}, {0,1,0}, {0,0,-1}, {0,0,1} };
	m->paths = (CoarseningPath*)malloc(sizeof(CoarseningPath) * 27 * 7 * 8);
	m->npaths = 0;
	for (int uz = 2 + m->smin[2]; uz <= 2 + m->smax[2]; uz++) for (int uy = 1; uy <= 3; uy++) for (int ux = 1; ux <= 3; ux++) {
		for (int i = 0; i < 1 + 2 * m->dim; i++) {
			const int wx = ux + p7[i][0], wy = uy + p7[i][1], wz = uz + p7[i][2];
			for (int nz = wz / 2; nz <= (wz + 1) / 2; nz++) for (int ny = wy / 2; ny <= (wy + 1) / 2; ny++) for (int nx = wx / 2; nx <= (wx + 1) / 2; nx++) {
				const int s = nx + 3 * ny + 9 * nz;
				if (s >= 13) {
					CoarseningPath* p = &m->paths[m->npaths++];
					p->Nx = nx - 1; p->Ny = ny - 1; p->Nz = nz - 1;
					p->Ux = ux - 2; p->Uy = uy - 2; p->Uz = uz - 2;
					p->Wx = wx - 2; p->Wy = wy - 2; p->Wz = wz - 2;
					p->sc = s - 13; p->sf = (i + 1) / 2; p->inU = (i % 2 == 0);
					p->rw = (Real)1 / (Real)(1 << ((ux % 2) + (uy % 2) + (uz % 2)));
					p->iw = (Real)1 / (Real)(1 << ((wx % 2) + (wy % 2) + (wz % 2)));
				}
			}
		}
	}
	/* stable insertion sort on the reference's key (std::sort's order among equal keys is unspecified) */
	for (int a = 1; a < m->npaths; a++) {
		CoarseningPath t = m->paths[a]; int b = a - 1;
		while (b >= 0 && path_less(&t, &m->paths[b]) < 0) { m->paths[b + 1] = m->paths[b]; b--; }
		m->paths[b + 1] = t;
	}
	return m;
}

static void mg_destroy(MgState* m)
{
	if (!m) return;
	for (int l = 0; l < m->nlev; l++) { free(m->A[l]); free(m->x[l]); free(m->b[l]); free(m->r[l]); free(m->type[l]); }
	free(m->cg1); free(m->cg2); free(m->cg3); free(m->cg4); free(m->paths); free(m);
}

#define VEC(v, l, V) int V[3]; { const int* s_ = m->size[l]; V[0] = (v) % s_[0]; V[1] = ((v) % (s_[0] * s_[1])) / s_[0]; V[2] = (v) / (s_[0] * s_[1]); }
#define LIN(V0, V1, V2, l) ((V0) + (V1) * m->pitch[l][1] + (V2) * m->pitch[l][2])
#define INGRID(V0, V1, V2, l) ((V0) >= 0 && (V1) >= 0 && (V2) >= 0 && (V0) < m->size[l][0] && (V1) < m->size[l][1] && (V2) < m->size[l][2])

/* ---- bucket min-heap, multigrid.cpp:59-203 (pop order = LIFO within a key) ---- */
typedef struct { int key, prev, next; } HEntry;
typedef struct { int N, K, size, minKey; HEntry* e; } NKMinHeap;
static void heap_init(NKMinHeap* h, int N, int K) { h->N = N; h->K = K; h->size = 0; h->minKey = -1; h->e = (HEntry*)malloc(sizeof(HEntry) * ((size_t)N + K)); for (int i = 0; i < N + K; i++) { h->e[i].key = -1; h->e[i].prev = -1; h->e[i].next = -1; } }
static inline int heap_get(NKMinHeap* h, int id) { return h->e[h->K + id].key; }
static void heap_set(NKMinHeap* h, int id, int key)
{
	const int kid = h->K + id; HEntry* e = h->e;
	if (e[kid].key == key) return;
	if (e[kid].key != -1) {
		const int pred = e[kid].prev, succ = e[kid].next;
		e[pred].next = succ; if (succ != -1) e[succ].prev = pred;
		const int removed = e[kid].key;
		if (removed == h->minKey) {
			if (h->size == 1) h->minKey = -1;
			else for (; h->minKey < h->K; h->minKey++) if (e[h->minKey].next != -1) break;
		}
		h->size--;
	}
	e[kid].key = key;
	if (key == -1) { e[kid].next = e[kid].prev = -1; return; }
	h->size++;
	if (h->minKey == -1) h->minKey = key; else if (key < h->minKey) h->minKey = key;
	const int tmp = e[key].next;
	e[key].next = kid; e[kid].prev = key; e[kid].next = tmp; if (tmp != -1) e[tmp].prev = kid;
}
static int heap_pop(NKMinHeap* h)
{
	if (h->size == 0) return -1;
	HEntry* e = h->e;
	const int kid = e[h->minKey].next, id = kid - h->K;
	const int pred = e[kid].prev, succ = e[kid].next;
	e[pred].next = succ; if (succ != -1) e[succ].prev = pred;
	e[kid].key = -1; e[kid].prev = -1; e[kid].next = -1;
	h->size--;
	if (h->size == 0) h->minKey = -1;
	else for (; h->minKey < h->K; h->minKey++) if (e[h->minKey].next != -1) break;
	return id;
}

/* genCoarseGrid multigrid.cpp:520-578 */
static void mg_gen_coarse_grid(MgState* m, int l)
{
	signed char* tc = m->type[l]; const signed char* tf = m->type[l - 1];
	memset(tc, vtFree, (size_t)m->n[l]);
	NKMinHeap h; heap_init(&h, m->n[l - 1], m->is3D ? 9 : 5);
	for (int v = 0; v < m->n[l - 1]; v++) if (tf[v] != vtInactive) {
		VEC(v, l - 1, V)
		heap_set(&h, v, 1 << ((V[0] % 2) + (V[1] % 2) + (V[2] % 2)));
	}
	while (h.size > 0) {
		const int v = heap_pop(&h);
		VEC(v, l - 1, V)
		int vdone = 0;
		for (int iz = V[2] / 2; iz <= (V[2] + 1) / 2; iz++) for (int iy = V[1] / 2; iy <= (V[1] + 1) / 2; iy++) for (int ix = V[0] / 2; ix <= (V[0] + 1) / 2; ix++) {
			const int i = LIN(ix, iy, iz, l);
			if (tc[i] == vtFree) {
				if (vdone) tc[i] = vtRemoved; else { tc[i] = vtZero; vdone = 1; }
				const int* sf = m->size[l - 1];
				const int r0x = ix * 2 - 1 > 0 ? ix * 2 - 1 : 0, r0y = iy * 2 - 1 > 0 ? iy * 2 - 1 : 0, r0z = iz * 2 - 1 > 0 ? iz * 2 - 1 : 0;
				const int r1x = ix * 2 + 1 < sf[0] - 1 ? ix * 2 + 1 : sf[0] - 1, r1y = iy * 2 + 1 < sf[1] - 1 ? iy * 2 + 1 : sf[1] - 1, r1z = iz * 2 + 1 < sf[2] - 1 ? iz * 2 + 1 : sf[2] - 1;
				for (int rz = r0z; rz <= r1z; rz++) for (int ry = r0y; ry <= r1y; ry++) for (int rx = r0x; rx <= r1x; rx++) {
					const int r = LIN(rx, ry, rz, l - 1);
					const int key = heap_get(&h, r);
					if (key > 1) heap_set(&h, r, key - 1);
					else if (key > -1) heap_set(&h, r, -1);
				}
			}
		}
	}
	free(h.e);
	for (int i = 0; i < m->n[l]; i++) {                     /* knActivateCoarseVertices :507-516 */
		if (tc[i] == vtFree) tc[i] = vtRemoved;
		if (tc[i] == vtZero) tc[i] = vtActive;
		if (tc[i] == vtRemoved) tc[i] = vtInactive;
	}
}

/* knGenCoarseGridOperator multigrid.cpp:580-657: A_l = R A_{l-1} I */
static void mg_gen_coarse_operator(MgState* m, int l)
{
	const int S = m->stencil, S0 = m->stencil0;
	Real* A = m->A[l]; const Real* Af = m->A[l - 1];
	const signed char* tc = m->type[l]; const signed char* tf = m->type[l - 1];
	#pragma omp parallel for schedule(dynamic, 256)
	for (int idx = 0; idx < m->n[l]; idx++) {
		if (tc[idx] == vtInactive) continue;
		for (int i = 0; i < S; i++) A[(size_t)idx * S + i] = (Real)0;
		VEC(idx, l, V)
		if (l == 1) {
			for (int q = 0; q < m->npaths; q++) {
				const CoarseningPath* p = &m->paths[q];
				const int Nx = V[0] + p->Nx, Ny = V[1] + p->Ny, Nz = V[2] + p->Nz;
				if (!INGRID(Nx, Ny, Nz, l) || tc[LIN(Nx, Ny, Nz, l)] == vtInactive) continue;
				const int Ux = V[0] * 2 + p->Ux, Uy = V[1] * 2 + p->Uy, Uz = V[2] * 2 + p->Uz;
				if (!INGRID(Ux, Uy, Uz, 0)) continue;
				const int u = LIN(Ux, Uy, Uz, 0); if (tf[u] == vtInactive) continue;
				const int Wx = V[0] * 2 + p->Wx, Wy = V[1] * 2 + p->Wy, Wz = V[2] * 2 + p->Wz;
				if (!INGRID(Wx, Wy, Wz, 0)) continue;
				const int w = LIN(Wx, Wy, Wz, 0); if (tf[w] == vtInactive) continue;
				if (p->inU) A[(size_t)idx * S + p->sc] += p->rw * Af[(size_t)u * S0 + p->sf] * p->iw;
				else        A[(size_t)idx * S + p->sc] += p->rw * Af[(size_t)w * S0 + p->sf] * p->iw;
			}
		} else {
			const int* sf = m->size[l - 1]; const int* sc_ = m->size[l];
			int u0[3], u1[3];
			for (int d = 0; d < 3; d++) { u0[d] = V[d] * 2 - 1 > 0 ? V[d] * 2 - 1 : 0; u1[d] = V[d] * 2 + 1 < sf[d] - 1 ? V[d] * 2 + 1 : sf[d] - 1; }
			for (int Uz = u0[2]; Uz <= u1[2]; Uz++) for (int Uy = u0[1]; Uy <= u1[1]; Uy++) for (int Ux = u0[0]; Ux <= u1[0]; Ux++) {
				const int U[3] = { Ux, Uy, Uz };
				const int u = LIN(Ux, Uy, Uz, l - 1);
				if (tf[u] == vtInactive) continue;
				const Real rw = (Real)1 / (Real)(1 << ((Ux % 2) + (Uy % 2) + (Uz % 2)));
				int n0[3], n1[3];
				/* C integer division truncates toward zero exactly like the reference's Vec3i '/' */
				for (int d = 0; d < 3; d++) { n0[d] = (U[d] - 1) / 2; n1[d] = (U[d] + 2) / 2 < sc_[d] - 1 ? (U[d] + 2) / 2 : sc_[d] - 1; }
				for (int Nz = n0[2]; Nz <= n1[2]; Nz++) for (int Ny = n0[1]; Ny <= n1[1]; Ny++) for (int Nx = n0[0]; Nx <= n1[0]; Nx++) {
					const int N[3] = { Nx, Ny, Nz };
					const int nn = LIN(Nx, Ny, Nz, l);
					if (tc[nn] == vtInactive) continue;
					const int sc = (Nx - V[0] + m->smax[0]) + 3 * (Ny - V[1] + m->smax[1]) + 9 * (Nz - V[2] + m->smax[2]);
					if (sc < S - 1) continue;
					int w0[3], w1[3];
					for (int d = 0; d < 3; d++) {
						int a = U[d] - 1 > N[d] * 2 - 1 ? U[d] - 1 : N[d] * 2 - 1; if (a < 0) a = 0;
						int b = U[d] + 1 < N[d] * 2 + 1 ? U[d] + 1 : N[d] * 2 + 1; if (b > sf[d] - 1) b = sf[d] - 1;
						w0[d] = a; w1[d] = b;
					}
					for (int Wz = w0[2]; Wz <= w1[2]; Wz++) for (int Wy = w0[1]; Wy <= w1[1]; Wy++) for (int Wx = w0[0]; Wx <= w1[0]; Wx++) {
						const int w = LIN(Wx, Wy, Wz, l - 1);
						if (tf[w] == vtInactive) continue;
						const int sfi = (Wx - Ux + m->smax[0]) + 3 * (Wy - Uy + m->smax[1]) + 9 * (Wz - Uz + m->smax[2]);
						const Real iw = (Real)1 / (Real)(1 << ((Wx % 2) + (Wy % 2) + (Wz % 2)));
						if (sfi < S) A[(size_t)idx * S + sc - S + 1] += rw * Af[(size_t)w * S + S - 1 - sfi] * iw;
						else         A[(size_t)idx * S + sc - S + 1] += rw * Af[(size_t)u * S + sfi - S + 1] * iw;
					}
				}
			}
		}
	}
}

/* GridMg::setA multigrid.cpp:386-415 (knCopyA :350-358, knActivateVertices :360-384, analyzeStencil :321-348) */
static void mg_set_a(MgState* m, const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{
	const int S0 = m->stencil0; Real* A = m->A[0]; const int n = m->n[0];
	for (int idx = 0; idx < n; idx++) {
		A[(size_t)idx * S0 + 0] = A0[idx]; A[(size_t)idx * S0 + 1] = Ai[idx]; A[(size_t)idx * S0 + 2] = Aj[idx];
		if (m->is3D) A[(size_t)idx * S0 + 3] = Ak[idx];
	}
	for (int v = 0; v < n; v++) {
		m->type[0][v] = vtInactive;
		if (A[(size_t)v * S0 + 0] != (Real)0) {
			m->type[0][v] = vtActive;
			VEC(v, 0, V)
			Real a[7];
			a[0] = A[(size_t)v * S0 + 0]; a[1] = A[(size_t)v * S0 + 1]; a[2] = A[(size_t)v * S0 + 2];
			a[3] = m->is3D ? A[(size_t)v * S0 + 3] : (Real)0;
			a[4] = V[0] != 0 ? A[(size_t)(v - m->pitch[0][0]) * S0 + 1] : (Real)0;
			a[5] = V[1] != 0 ? A[(size_t)(v - m->pitch[0][1]) * S0 + 2] : (Real)0;
			a[6] = (V[2] != 0 && m->is3D) ? A[(size_t)(v - m->pitch[0][2]) * S0 + 3] : (Real)0;
			const int trivial = a[0] == (Real)1 && a[1] == 0 && a[2] == 0 && a[3] == 0 && a[4] == 0 && a[5] == 0 && a[6] == 0;
			if (trivial) { m->type[0][v] = vtActiveTrivial; A[(size_t)v * S0 + 0] *= m->trivialScale; }
		}
	}
	for (int l = 1; l < m->nlev; l++) { mg_gen_coarse_grid(m, l); mg_gen_coarse_operator(m, l); }
	m->isASet = 1; m->isRhsSet = 0;
}

/* NOTE on knActivateVertices' read-after-scale: the reference's kernel scales A0 of trivial rows
 * in place while other threads may analyse neighbours; analyzeStencil only reads the *diagonal of
 * its own row* and the *off-diagonals* of neighbours, and trivial rows have zero off-diagonals, so
 * the result does not depend on the order.  Same here.                                            */

static void mg_set_rhs(MgState* m, const Real* rhs)          /* multigrid.cpp:417-433 */
{
	for (int i = 0; i < m->n[0]; i++) { Real v = rhs[i]; if (m->type[0][i] == vtActiveTrivial) v *= m->trivialScale; m->b[0][i] = v; }
	m->isRhsSet = 1;
}

/* knSmoothColor multigrid.cpp:668-711 + smoothGS :713-737 */
static void mg_smooth(MgState* m, int l, int reversed)
{
	static const int a8[8][3] = { {0,0,0},{1,0,0},{0,1,0},{1,1,0},{0,0,1},{1,0,1},{0,1,1},{1,1,1} };
	int colors[8][4]; int ncol, percol;
	if (m->is3D) {
		if (l == 0) { ncol = 2; percol = 4; int c0[4] = {0,3,5,6}, c1[4] = {1,2,4,7}; memcpy(colors[0], c0, sizeof c0); memcpy(colors[1], c1, sizeof c1); }
		else { ncol = 8; percol = 1; for (int c = 0; c < 8; c++) colors[c][0] = c; }
	} else {
		if (l == 0) { ncol = 2; percol = 2; colors[0][0] = 0; colors[0][1] = 3; colors[1][0] = 1; colors[1][1] = 2; }
		else { ncol = 4; percol = 1; for (int c = 0; c < 4; c++) colors[c][0] = c; }
	}
	const int* s = m->size[l];
	const int bs[3] = { (s[0] + 1) / 2, (s[1] + 1) / 2, (s[2] + 1) / 2 };
	const int nblocks = bs[0] * bs[1] * bs[2];
	const int S = m->stencil, S0 = m->stencil0;
	Real* x = m->x[l]; const Real* A = m->A[l]; const Real* b = m->b[l]; const signed char* type = m->type[l];
	for (int c = 0; c < ncol; c++) {
		const int color = reversed ? ncol - 1 - c : c;
		#pragma omp parallel for schedule(static)
		for (int blk = 0; blk < nblocks; blk++) {
			const int bo[3] = { blk % bs[0], (blk % (bs[0] * bs[1])) / bs[0], blk / (bs[0] * bs[1]) };
			for (int off = 0; off < percol; off++) {
				const int* co = a8[colors[color][off]];
				const int V[3] = { bo[0] * 2 + co[0], bo[1] * 2 + co[1], bo[2] * 2 + co[2] };
				if (!INGRID(V[0], V[1], V[2], l)) continue;
				const int v = LIN(V[0], V[1], V[2], l);
				if (type[v] == vtInactive) continue;
				Real sum = b[v];
				if (l == 0) {
					for (int d = 0; d < m->dim; d++) {
						if (V[d] > 0)        { const int n = v - m->pitch[0][d]; sum -= A[(size_t)n * S0 + d + 1] * x[n]; }
						if (V[d] < s[d] - 1) { const int n = v + m->pitch[0][d]; sum -= A[(size_t)v * S0 + d + 1] * x[n]; }
					}
					x[v] = sum / A[(size_t)v * S0 + 0];
				} else {
					int sidx = 0;
					for (int dz = m->smin[2]; dz <= m->smax[2]; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++, sidx++) {
						if (sidx == S - 1) continue;
						const int N0 = V[0] + dx, N1 = V[1] + dy, N2 = V[2] + dz;
						if (!INGRID(N0, N1, N2, l)) continue;
						const int n = LIN(N0, N1, N2, l);
						if (type[n] == vtInactive) continue;
						if (sidx < S) sum -= A[(size_t)n * S + S - 1 - sidx] * x[n];
						else          sum -= A[(size_t)v * S + sidx - S + 1] * x[n];
					}
					x[v] = sum / A[(size_t)v * S + 0];
				}
			}
		}
	}
}

/* knCalcResidual multigrid.cpp:739-771 */
static void mg_residual(MgState* m, int l)
{
	const int* s = m->size[l]; const int S = m->stencil, S0 = m->stencil0;
	const Real* x = m->x[l]; const Real* A = m->A[l]; const Real* b = m->b[l]; const signed char* type = m->type[l]; Real* r = m->r[l];
	#pragma omp parallel for schedule(static)
	for (int idx = 0; idx < m->n[l]; idx++) {
		if (type[idx] == vtInactive) continue;
		VEC(idx, l, V)
		Real sum = b[idx];
		if (l == 0) {
			for (int d = 0; d < m->dim; d++) {
				if (V[d] > 0)        { const int n = idx - m->pitch[0][d]; sum -= A[(size_t)n * S0 + d + 1] * x[n]; }
				if (V[d] < s[d] - 1) { const int n = idx + m->pitch[0][d]; sum -= A[(size_t)idx * S0 + d + 1] * x[n]; }
			}
			sum -= A[(size_t)idx * S0 + 0] * x[idx];
		} else {
			int sidx = 0;
			for (int dz = m->smin[2]; dz <= m->smax[2]; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++, sidx++) {
				const int N0 = V[0] + dx, N1 = V[1] + dy, N2 = V[2] + dz;
				if (!INGRID(N0, N1, N2, l)) continue;
				const int n = LIN(N0, N1, N2, l);
				if (type[n] == vtInactive) continue;
				if (sidx < S) sum -= A[(size_t)n * S + S - 1 - sidx] * x[n];
				else          sum -= A[(size_t)idx * S + sidx - S + 1] * x[n];
			}
		}
		r[idx] = sum;
	}
}

/* knRestrict multigrid.cpp:904-927 */
static void mg_restrict(MgState* m, int ld, const Real* src, Real* dst)
{
	const int ls = ld - 1; const int* sf = m->size[ls];
	#pragma omp parallel for schedule(static)
	for (int idx = 0; idx < m->n[ld]; idx++) {
		if (m->type[ld][idx] == vtInactive) continue;
		VEC(idx, ld, V)
		Real sum = (Real)0;
		int r0[3], r1[3];
		for (int d = 0; d < 3; d++) { r0[d] = V[d] * 2 - 1 > 0 ? V[d] * 2 - 1 : 0; r1[d] = V[d] * 2 + 1 < sf[d] - 1 ? V[d] * 2 + 1 : sf[d] - 1; }
		for (int rz = r0[2]; rz <= r1[2]; rz++) for (int ry = r0[1]; ry <= r1[1]; ry++) for (int rx = r0[0]; rx <= r1[0]; rx++) {
			const int r = LIN(rx, ry, rz, ls);
			if (m->type[ls][r] == vtInactive) continue;
			const Real rw = (Real)1 / (Real)(1 << ((rx % 2) + (ry % 2) + (rz % 2)));
			sum += rw * src[r];
		}
		dst[idx] = sum;
	}
}

/* knInterpolate multigrid.cpp:934-954 */
static void mg_interpolate(MgState* m, int ld, const Real* src, Real* dst)
{
	const int ls = ld + 1;
	#pragma omp parallel for schedule(static)
	for (int idx = 0; idx < m->n[ld]; idx++) {
		if (m->type[ld][idx] == vtInactive) continue;
		VEC(idx, ld, V)
		Real sum = (Real)0;
		for (int iz = V[2] / 2; iz <= (V[2] + 1) / 2; iz++) for (int iy = V[1] / 2; iy <= (V[1] + 1) / 2; iy++) for (int ix = V[0] / 2; ix <= (V[0] + 1) / 2; ix++) {
			const int i = LIN(ix, iy, iz, ls);
			if (m->type[ls][i] != vtInactive) sum += src[i];
		}
		const Real iw = (Real)1 / (Real)(1 << ((V[0] % 2) + (V[1] % 2) + (V[2] % 2)));
		dst[idx] = iw * sum;
	}
}

/* GridMg::solveCG multigrid.cpp:796-902: serial double-precision Jacobi-PCG on the coarsest level */
static double mg_apply_stencil(const MgState* m, int v, int l, const double* vec)
{
	const int* s = m->size[l]; const int S = m->stencil, S0 = m->stencil0; const Real* A = m->A[l];
	VEC(v, l, V)
	double sum = 0;
	if (l == 0) {
		for (int d = 0; d < m->dim; d++) {
			if (V[d] > 0)        { const int n = v - m->pitch[0][d]; sum += A[(size_t)n * S0 + d + 1] * vec[n]; }
			if (V[d] < s[d] - 1) { const int n = v + m->pitch[0][d]; sum += A[(size_t)v * S0 + d + 1] * vec[n]; }
		}
		sum += A[(size_t)v * S0 + 0] * vec[v];
	} else {
		int sidx = 0;
		for (int dz = m->smin[2]; dz <= m->smax[2]; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++, sidx++) {
			const int N0 = V[0] + dx, N1 = V[1] + dy, N2 = V[2] + dz;
			if (!INGRID(N0, N1, N2, l)) continue;
			const int n = LIN(N0, N1, N2, l);
			if (m->type[l][n] == vtInactive) continue;
			if (sidx < S) sum += A[(size_t)n * S + S - 1 - sidx] * vec[n];
			else          sum += A[(size_t)v * S + sidx - S + 1] * vec[n];
		}
	}
	return sum;
}
static void mg_solve_cg(MgState* m, int l)
{
	double *z = m->cg1, *p = m->cg2, *x = m->cg3, *r = m->cg4;
	const int n = m->n[l]; const signed char* type = m->type[l];
	const int S = (l == 0) ? m->stencil0 : m->stencil; const Real* A = m->A[l];
	double alphaTop = 0, initialResidual = 0;
	for (int v = 0; v < n; v++) x[v] = m->x[l][v];
	for (int v = 0; v < n; v++) {
		if (type[v] == vtInactive) continue;
		r[v] = m->b[l][v] - mg_apply_stencil(m, v, l, x);
		z[v] = r[v] / A[(size_t)v * S + 0];
		initialResidual += r[v] * r[v];
		p[v] = z[v];
		alphaTop += r[v] * z[v];
	}
	initialResidual = sqrt(initialResidual);
	int iter = 0; const int maxIter = 10000; double residual = -1;
	for (; iter < maxIter && initialResidual > 1E-12; iter++) {
		double alphaBot = 0;
		for (int v = 0; v < n; v++) { if (type[v] == vtInactive) continue; z[v] = mg_apply_stencil(m, v, l, p); alphaBot += p[v] * z[v]; }
		const double alpha = alphaTop / alphaBot;
		double alphaTopNew = 0; residual = 0;
		for (int v = 0; v < n; v++) {
			if (type[v] == vtInactive) continue;
			x[v] += alpha * p[v];
			r[v] -= alpha * z[v];
			residual += r[v] * r[v];
			z[v] = r[v] / A[(size_t)v * S + 0];
			alphaTopNew += r[v] * z[v];
		}
		residual = sqrt(residual);
		if (residual / initialResidual < m->coarsestAcc) break;
		const double beta = alphaTopNew / alphaTop;
		alphaTop = alphaTopNew;
		for (int v = 0; v < n; v++) p[v] = z[v] + beta * p[v];
	}
	m->lastCoarseIters = iter;
	for (int v = 0; v < n; v++) m->x[l][v] = (Real)x[v];
}

/* GridMg::doVCycle multigrid.cpp:448-504 with src == NULL (as the preconditioner calls it) */
static void mg_vcycle(MgState* m, Real* dst)
{
	const int maxLevel = m->nlev - 1;
	memset(m->x[0], 0, sizeof(Real) * (size_t)m->n[0]);
	for (int l = 0; l < maxLevel; l++) {
		for (int i = 0; i < m->numPre; i++) mg_smooth(m, l, 0);
		mg_residual(m, l);
		mg_restrict(m, l + 1, m->r[l], m->b[l + 1]);
		memset(m->x[l + 1], 0, sizeof(Real) * (size_t)m->n[l + 1]);
	}
	mg_solve_cg(m, maxLevel);
	for (int l = maxLevel - 1; l >= 0; l--) {
		mg_interpolate(m, l, m->x[l + 1], m->r[l]);
		for (int i = 0; i < m->n[l]; i++) m->x[l][i] += m->r[l][i];
		for (int i = 0; i < m->numPost; i++) mg_smooth(m, l, 1);
	}
	mg_residual(m, 0);                                        /* result (norm) unused by the caller */
	memcpy(dst, m->x[0], sizeof(Real) * (size_t)m->n[0]);
}

/* ---- exported GridMg probes (same shape as oracle/ref_harness.cpp) ---- */
static MgState* g_mg = 0;
int mfo_mg_create(int sx, int sy, int sz) { mg_destroy(g_mg); g_mg = mg_create(sx, sy, sz); return 0; }
int mfo_mg_destroy(void) { mg_destroy(g_mg); g_mg = 0; return 0; }
int mfo_mg_set_a(const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak) { mg_set_a(g_mg, A0, Ai, Aj, Ak); return 0; }
int mfo_mg_num_levels(void) { return g_mg ? g_mg->nlev : 0; }
int mfo_mg_level_size(int l, int* out3) { for (int d = 0; d < 3; d++) out3[d] = g_mg->size[l][d]; return 0; }
int mfo_mg_stencil_size(int l) { return l == 0 ? g_mg->stencil0 : g_mg->stencil; }
int mfo_mg_get_type(int l, signed char* out) { memcpy(out, g_mg->type[l], (size_t)g_mg->n[l]); return 0; }
int mfo_mg_get_a(int l, Real* out) { memcpy(out, g_mg->A[l], sizeof(Real) * (size_t)g_mg->n[l] * (l == 0 ? g_mg->stencil0 : g_mg->stencil)); return 0; }
int mfo_mg_get_x(int l, Real* out) { memcpy(out, g_mg->x[l], sizeof(Real) * (size_t)g_mg->n[l]); return 0; }
int mfo_mg_get_b(int l, Real* out) { memcpy(out, g_mg->b[l], sizeof(Real) * (size_t)g_mg->n[l]); return 0; }
int mfo_mg_get_r(int l, Real* out) { memcpy(out, g_mg->r[l], sizeof(Real) * (size_t)g_mg->n[l]); return 0; }
int mfo_mg_vcycle(const Real* rhs, Real* dst, double coarsestAccuracy, int pre, int post)
{
	if (!g_mg || !g_mg->isASet) { snprintf(g_err, sizeof g_err, "GridMg::setRhs Error: A has not been set."); return 1; }
	g_mg->coarsestAcc = (Real)coarsestAccuracy; g_mg->numPre = pre; g_mg->numPost = post;
	mg_set_rhs(g_mg, rhs); mg_vcycle(g_mg, dst); return 0;
}

/* ---------------------------------------------------------------------------------------------
 * IC(0) "a la Wavelet Turbulence": InitPreconditionIncompCholesky conjugategrad.cpp:26-63 (serial, k/j/i order; the scatter to the
 * +x/+y/+z neighbours also lands on non-fluid cells), InvertCheckFluid commonkernels.h:25-29.  P0..Pk receive the factor. */
static void ic_init(int sx, int sy, int sz, const int* flags, Real* P0, Real* Pi, Real* Pj, Real* Pk, const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	memcpy(P0, A0, sizeof(Real) * (size_t)n); memcpy(Pi, Ai, sizeof(Real) * (size_t)n);
	memcpy(Pj, Aj, sizeof(Real) * (size_t)n); memcpy(Pk, Ak, sizeof(Real) * (size_t)n);
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeFluid)) continue;
		P0[idx] = R_SQRT(P0[idx]);
		const Real invDiagonal = 1.0f / P0[idx];
		Pi[idx] *= invDiagonal; Pj[idx] *= invDiagonal; Pk[idx] *= invDiagonal;
		P0[idx + X] -= Pi[idx] * Pi[idx];
		P0[idx + Y] -= Pj[idx] * Pj[idx];
		P0[idx + Z] -= Pk[idx] * Pk[idx];
	}
	for (IndexInt q = 0; q < n; q++) if ((flags[q] & TypeFluid) && P0[q] > 0) P0[q] = (Real)(1.0 / (double)P0[q]);
}
int mfo_ic_init(int sx, int sy, int sz, const int* flags, Real* P0, Real* Pi, Real* Pj, Real* Pk, const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{
	if (sz < 2) { snprintf(g_err, sizeof g_err, "ICP only supports 3D grids so far"); return 1; }
	ic_init(sx, sy, sz, flags, P0, Pi, Pj, Pk, A0, Ai, Aj, Ak); return 0;
}

/* ApplyPreconditionIncompCholesky conjugategrad.cpp:109-132 */
static void ic_apply(int sx, int sy, int sz, const int* flags, Real* dst, const Real* src, const Real* P0, const Real* Pi, const Real* Pj, const Real* Pk)
{
	STRIDES
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeFluid)) continue;
		dst[idx] = P0[idx] * (src[idx] - dst[idx - X] * Pi[idx - X] - dst[idx - Y] * Pj[idx - Y] - dst[idx - Z] * Pk[idx - Z]);
	}
	for (int k = sz - 1; k >= 0; k--) for (int j = sy - 1; j >= 0; j--) for (int i = sx - 1; i >= 0; i--) {
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeFluid)) continue;
		dst[idx] = P0[idx] * (dst[idx] - dst[idx + X] * Pi[idx] - dst[idx + Y] * Pj[idx] - dst[idx + Z] * Pk[idx]);
	}
}
int mfo_ic_apply(int sx, int sy, int sz, const int* flags, Real* dst, const Real* src, const Real* P0, const Real* Pi, const Real* Pj, const Real* Pk)
{ ic_apply(sx, sy, sz, flags, dst, src, P0, Pi, Pj, Pk); return 0; }

/* =============================================================================================
 * GridCg<APPLYMAT>: conjugategrad.cpp:201-307 (doInit :209-235, iterate :237-299)
 * pc: 0 PC_None, 1 PC_mICP, 2 PC_MGP, 3 PC_ICP.  `mg` may carry an already-set hierarchy (PcMGStatic).     */
static int cg_run(int sx, int sy, int sz, const int* flags, const Real* rhs, Real* x,
	const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak,
	int pc, Real accuracy, int useL2, int maxIter, MgState* mg, int* iterations, double* resNormOut)
{
	const IndexInt n = (IndexInt)sx * sy * sz;
	Real* residual = (Real*)calloc((size_t)n, sizeof(Real));
	Real* search = (Real*)calloc((size_t)n, sizeof(Real));
	Real* tmp = (Real*)calloc((size_t)n, sizeof(Real));
	Real* P = 0; Real* Pc[3] = { 0, 0, 0 };
	int rc = 0;
	if ((pc == 1 || pc == 3 || pc == 4) && !IS3D) pc = 0;      /* setICPreconditioner :315-321 */
	/* doInit */
	memset(x, 0, sizeof(Real) * (size_t)n);
	memcpy(residual, rhs, sizeof(Real) * (size_t)n);
	if (pc == 1) {
		P = (Real*)calloc((size_t)n, sizeof(Real));
		mic_init(sx, sy, sz, flags, P, A0, Ai, Aj, Ak);
		mic_apply(sx, sy, sz, flags, tmp, residual, P, Ai, Aj, Ak);
	} else if (pc == 4) {                                       /* MIC(0) in block red-black ordering (no reference counterpart, see micrb_init) */
		P = (Real*)calloc((size_t)n, sizeof(Real));
		micrb_init(sx, sy, sz, flags, P, A0, Ai, Aj, Ak);
		micrb_apply(sx, sy, sz, flags, tmp, residual, P, Ai, Aj, Ak);
	} else if (pc == 3) {
		P = (Real*)calloc((size_t)n, sizeof(Real));
		for (int c = 0; c < 3; c++) Pc[c] = (Real*)calloc((size_t)n, sizeof(Real));
		ic_init(sx, sy, sz, flags, P, Pc[0], Pc[1], Pc[2], A0, Ai, Aj, Ak);
		ic_apply(sx, sy, sz, flags, tmp, residual, P, Pc[0], Pc[1], Pc[2]);
	} else if (pc == 2) {
		if (!mg->isASet) mg_set_a(mg, A0, Ai, Aj, Ak);         /* InitPreconditionMultigrid :100-106 */
		mg->coarsestAcc = (Real)(accuracy * 1E-4); mg->numPre = 1; mg->numPost = 1;
		mg_set_rhs(mg, residual); mg_vcycle(mg, tmp);
	} else memcpy(tmp, residual, sizeof(Real) * (size_t)n);
	memcpy(search, tmp, sizeof(Real) * (size_t)n);
	Real sigma = (Real)dot_product(tmp, residual, n);
	Real resNorm = (Real)1e20;
	int its = 0;
	for (int iter = 0; iter < maxIter; iter++) {
		its++;
		apply_matrix(sx, sy, sz, flags, tmp, search, A0, Ai, Aj, Ak);
		const Real dp = (Real)dot_product(tmp, search, n);
		Real alpha = 0.;
		if (R_FABS(dp) > 0.) alpha = sigma / dp;
		scaled_add(x, search, alpha, n);
		scaled_add(residual, tmp, -alpha, n);
		if (pc == 1) mic_apply(sx, sy, sz, flags, tmp, residual, P, Ai, Aj, Ak);
		else if (pc == 4) micrb_apply(sx, sy, sz, flags, tmp, residual, P, Ai, Aj, Ak);
		else if (pc == 3) ic_apply(sx, sy, sz, flags, tmp, residual, P, Pc[0], Pc[1], Pc[2]);
		else if (pc == 2) { mg_set_rhs(mg, residual); mg_vcycle(mg, tmp); }
		else memcpy(tmp, residual, sizeof(Real) * (size_t)n);
		if (useL2) resNorm = (Real)sum_sqr(residual, n); else resNorm = max_abs(residual, n);
		if (resNorm < accuracy) { sigma = resNorm; break; }
		const Real sigmaNew = (Real)dot_product(tmp, residual, n);
		const Real beta = sigmaNew / sigma;
		#pragma omp parallel for schedule(static)
		for (IndexInt q = 0; q < n; q++) search[q] = tmp[q] + beta * search[q];   /* UpdateSearchVec :193-196 */
		sigma = sigmaNew;
		if (!(resNorm < 1e35)) { snprintf(g_err, sizeof g_err, "GridCg::iterate: The CG solver diverged, residual norm > 1e30, stopping."); rc = 1; break; }
	}
	if (iterations) *iterations = its;
	if (resNormOut) *resNormOut = resNorm;
	free(residual); free(search); free(tmp); free(P); free(Pc[0]); free(Pc[1]); free(Pc[2]);
	return rc;
}

int mfo_cg_solve(int sx, int sy, int sz, const int* flags, const Real* rhs, Real* x,
	const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak,
	int pc, double accuracy, int useL2, int maxIter, int* iterations, double* resNorm)
{
	MgState* mg = (pc == 2) ? mg_create(sx, sy, sz) : 0;
	const int rc = cg_run(sx, sy, sz, flags, rhs, x, A0, Ai, Aj, Ak, pc, (Real)accuracy, useL2, maxIter, mg, iterations, resNorm);
	mg_destroy(mg);
	return rc;
}

/* ---------------------------------------------------------------------------------------------
 * cgSolveDiffusion conjugategrad.cpp:350-423 (another GridCg caller).  ncomp 1: Real grid, 3: Vec3/MAC grid
 * (each component solved separately through getComponent/setComponent, grid.cpp:676-685).            */
int mfo_cg_solve_diffusion(int sx, int sy, int sz, const int* flags, Real* data, int ncomp, double alpha_, double cgMaxIterFac, double cgAccuracy)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	const Real alpha = (Real)alpha_;
	Real *A0 = (Real*)calloc((size_t)n, sizeof(Real)), *Ai = (Real*)calloc((size_t)n, sizeof(Real)), *Aj = (Real*)calloc((size_t)n, sizeof(Real)), *Ak = (Real*)calloc((size_t)n, sizeof(Real));
	int* dummy = (int*)malloc(sizeof(int) * (size_t)n);
	for (IndexInt q = 0; q < n; q++) dummy[q] = TypeFluid;                      /* flagsDummy.setConst(TypeFluid) :361 */
	mfo_make_matrix(sx, sy, sz, dummy, 0, 0, 0., A0, Ai, Aj, Ak);
	free(dummy);
	for (IndexInt q = 0; q < n; q++) {                                          /* :364-375 */
		if (flags[q] & TypeObstacle) { Ai[q] = Aj[q] = Ak[q] = 0.0; A0[q] = 1.0; }
		else { Ai[q] *= alpha; Aj[q] *= alpha; Ak[q] *= alpha; A0[q] *= alpha; A0[q] += 1.; }
	}
	const int maxDim = sx > sy ? (sx > sz ? sx : sz) : (sy > sz ? sy : sz);
	const int maxIter = (int)((Real)cgMaxIterFac * maxDim) * (IS3D ? 1 : 4);
	Real* u = (Real*)malloc(sizeof(Real) * (size_t)n); Real* rhs = (Real*)malloc(sizeof(Real) * (size_t)n);
	int rc = 0;
	for (int c = 0; c < (ncomp == 1 ? 1 : (IS3D ? 3 : 2)) && rc == 0; c++) {
		for (IndexInt q = 0; q < n; q++) rhs[q] = data[(size_t)q * ncomp + c];
		/* GridCg defaults: no preconditioner, L2 norm (sum r^2 un-square-rooted) -- cgSolveDiffusion never calls setUseL2Norm */
		rc = cg_run(sx, sy, sz, flags, rhs, u, A0, Ai, Aj, Ak, 0, (Real)cgAccuracy, 1, maxIter, 0, 0, 0);
		for (IndexInt q = 0; q < n; q++) data[(size_t)q * ncomp + c] = u[q];
	}
	free(u); free(rhs); free(A0); free(Ai); free(Aj); free(Ak);
	return rc;
}

/* ---------------------------------------------------------------------------------------------
 * The grid half of VICintegration plugin/vortexplugins.cpp:253-299 (another GridCg caller): from the vorticity grid the Peskin
 * kernel left (:203-250, mesh code, not restated) to the velocity -- MakeLaplaceMatrix, CurlOp commonkernels.h:38-47, per component
 * GetShiftedComponent :104-108 (MAC target) / GetComponent :111-113, GridCg<ApplyMatrix> with the L2 stop test and the preconditioner
 * PreconditionType(precondition) (1 PC_ICP, 2 PC_mICP; conjugategrad.h:30), solution *= scale, SetComponent :121-123.            */
int mfo_vic_poisson(int sx, int sy, int sz, const int* flags, const Real* vort, Real* vel, int velIsMac, double cgMaxIterFac, double cgAccuracy,
	double scale_, int precondition, int* iters)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	if (!IS3D) { snprintf(g_err, sizeof g_err, "VICintegration: 3-D grids only (GridCg<ApplyMatrix>)"); return 1; }
	/* setICPreconditioner accepts PC_ICP and PC_mICP only (conjugategrad.cpp:312): the plugin's default precondition = 0 throws in the reference */
	if (precondition != 1 && precondition != 2) { snprintf(g_err, sizeof g_err, "GridCg<APPLYMAT>::setICPreconditioner: Invalid method specified."); return 1; }
	Real *A0 = (Real*)calloc((size_t)n, sizeof(Real)), *Ai = (Real*)calloc((size_t)n, sizeof(Real)), *Aj = (Real*)calloc((size_t)n, sizeof(Real)), *Ak = (Real*)calloc((size_t)n, sizeof(Real));
	Real *curl = (Real*)calloc((size_t)n * 3, sizeof(Real)), *rhs = (Real*)calloc((size_t)n, sizeof(Real)), *sol = (Real*)calloc((size_t)n, sizeof(Real));
	mfo_make_matrix(sx, sy, sz, flags, 0, 0, 0., A0, Ai, Aj, Ak);
	#define W(i_, j_, k_, c_) vort[3 * ((IndexInt)(i_) + Y * (j_) + Z * (k_)) + (c_)]
	for (int k = 1; k < sz - 1; k++) for (int j = 1; j < sy - 1; j++) for (int i = 1; i < sx - 1; i++) {                /* CurlOp, bnd = 1 */
		Real* v = curl + 3 * ((IndexInt)i + Y * j + Z * k);
		v[2] = (Real)(0.5 * (double)((W(i + 1, j, k, 1) - W(i - 1, j, k, 1)) - (W(i, j + 1, k, 0) - W(i, j - 1, k, 0))));
		v[0] = (Real)(0.5 * (double)((W(i, j + 1, k, 2) - W(i, j - 1, k, 2)) - (W(i, j, k + 1, 1) - W(i, j, k - 1, 1))));
		v[1] = (Real)(0.5 * (double)((W(i, j, k + 1, 0) - W(i, j, k - 1, 0)) - (W(i + 1, j, k, 2) - W(i - 1, j, k, 2))));
	}
	#undef W
	const int maxDim = sx > sy ? (sx > sz ? sx : sz) : (sy > sz ? sy : sz);
	const int maxIter = (int)((Real)cgMaxIterFac * maxDim);                                                            /* :273 */
	const int pc = precondition == 1 ? 3 : 1;                                                                          /* cg_run: 3 = IC(0), 1 = MIC(0) */
	const Real scale = (Real)scale_;
	int rc = 0;
	for (int c = 0; c < 3 && rc == 0; c++) {
		if (velIsMac) {                                                                                                /* GetShiftedComponent, bnd = 1: the border of rhs keeps its zeros */
			const IndexInt sh = c == 0 ? X : (c == 1 ? Y : Z);
			for (int k = 1; k < sz - 1; k++) for (int j = 1; j < sy - 1; j++) for (int i = 1; i < sx - 1; i++) {
				const IndexInt q = (IndexInt)i + Y * j + Z * k;
				rhs[q] = (Real)(0.5 * (double)(curl[3 * q + c] + curl[3 * (q - sh) + c]));
			}
		} else for (IndexInt q = 0; q < n; q++) rhs[q] = curl[3 * q + c];
		int its = 0;
		rc = cg_run(sx, sy, sz, flags, rhs, sol, A0, Ai, Aj, Ak, pc, (Real)cgAccuracy, 1, maxIter, 0, &its, 0);
		if (iters) iters[c] = its;
		for (IndexInt q = 0; q < n; q++) { sol[q] *= scale; vel[3 * q + c] = sol[q]; }                                   /* solution *= scale; SetComponent */
	}
	free(A0); free(Ai); free(Aj); free(Ak); free(curl); free(rhs); free(sol);
	return rc;
}

/* ---------------------------------------------------------------------------------------------
 * correctVelocity plugin/pressure.cpp:455-476: knCorrectVelocity :87-109,
 * knCorrectVelocityGhostFluid :154-187, knReplaceClampedGhostFluidVels :198-214                 */
int mfo_correct_velocity(int sx, int sy, int sz, const int* flags, Real* vel, const Real* pressure,
	const Real* phi, const Real* curv, double gfClamp_, double surfTens_)
{
	STRIDES
	const Real gfClamp = (Real)gfClamp_, surfTens = (Real)surfTens_;
	const int k0 = IS3D ? 1 : 0, k1 = IS3D ? sz - 1 : 1;
	#pragma omp parallel for schedule(static)
	for (int k = k0; k < k1; k++) for (int j = 1; j < sy - 1; j++) for (int i = 1; i < sx - 1; i++) {
		const IndexInt idx = IDX(i, j, k); Real* v = vel + 3 * idx;
		if (flags[idx] & TypeFluid) {
			if (flags[idx - X] & TypeFluid) v[0] -= (pressure[idx] - pressure[idx - X]);
			if (flags[idx - Y] & TypeFluid) v[1] -= (pressure[idx] - pressure[idx - Y]);
			if (IS3D && (flags[idx - Z] & TypeFluid)) v[2] -= (pressure[idx] - pressure[idx - Z]);
			if (flags[idx - X] & TypeEmpty) v[0] -= pressure[idx];
			if (flags[idx - Y] & TypeEmpty) v[1] -= pressure[idx];
			if (IS3D && (flags[idx - Z] & TypeEmpty)) v[2] -= pressure[idx];
		} else if ((flags[idx] & TypeEmpty) && !(flags[idx] & TypeOutflow)) {
			if (flags[idx - X] & TypeFluid) v[0] += pressure[idx - X]; else v[0] = 0.f;
			if (flags[idx - Y] & TypeFluid) v[1] += pressure[idx - Y]; else v[1] = 0.f;
			if (IS3D) { if (flags[idx - Z] & TypeFluid) v[2] += pressure[idx - Z]; else v[2] = 0.f; }
		}
	}
	if (!phi) return 0;
	#pragma omp parallel for schedule(static)
	for (int k = k0; k < k1; k++) for (int j = 1; j < sy - 1; j++) for (int i = 1; i < sx - 1; i++) {
		const IndexInt idx = IDX(i, j, k); Real* v = vel + 3 * idx;
		const int fl = flags[idx] & TypeFluid, em = (flags[idx] & TypeEmpty) && !(flags[idx] & TypeOutflow);
		if (fl) {
			if (flags[idx - X] & TypeEmpty) v[0] += pressure[idx] * ghostFluidHelper(idx, -X, phi, gfClamp);
			if (flags[idx - Y] & TypeEmpty) v[1] += pressure[idx] * ghostFluidHelper(idx, -Y, phi, gfClamp);
			if (IS3D && (flags[idx - Z] & TypeEmpty)) v[2] += pressure[idx] * ghostFluidHelper(idx, -Z, phi, gfClamp);
		} else if (em) {
			if (flags[idx - X] & TypeFluid) v[0] -= pressure[idx - X] * ghostFluidHelper(idx - X, +X, phi, gfClamp); else v[0] = 0.f;
			if (flags[idx - Y] & TypeFluid) v[1] -= pressure[idx - Y] * ghostFluidHelper(idx - Y, +Y, phi, gfClamp); else v[1] = 0.f;
			if (IS3D) { if (flags[idx - Z] & TypeFluid) v[2] -= pressure[idx - Z] * ghostFluidHelper(idx - Z, +Z, phi, gfClamp); else v[2] = 0.f; }
		}
		if (curv) {
			if (fl) {
				if (flags[idx - X] & TypeEmpty) v[0] += surfTensHelper(idx, -X, phi, curv, surfTens, gfClamp);
				if (flags[idx - Y] & TypeEmpty) v[1] += surfTensHelper(idx, -Y, phi, curv, surfTens, gfClamp);
				if (IS3D && (flags[idx - Z] & TypeEmpty)) v[2] += surfTensHelper(idx, -Z, phi, curv, surfTens, gfClamp);
			} else if (em) {
				v[0] -= (flags[idx - X] & TypeFluid) ? surfTensHelper(idx - X, +X, phi, curv, surfTens, gfClamp) : 0.f;
				v[1] -= (flags[idx - Y] & TypeFluid) ? surfTensHelper(idx - Y, +Y, phi, curv, surfTens, gfClamp) : 0.f;
				if (IS3D) v[2] -= (flags[idx - Z] & TypeFluid) ? surfTensHelper(idx - Z, +Z, phi, curv, surfTens, gfClamp) : 0.f;
			}
		}
	}
	/* knReplaceClampedGhostFluidVels: reads neighbours' *updated* velocity; an empty cell only
	   reads components of fluid cells, which this kernel never writes, so the order is immaterial */
	#pragma omp parallel for schedule(static)
	for (int k = k0; k < k1; k++) for (int j = 1; j < sy - 1; j++) for (int i = 1; i < sx - 1; i++) {
		const IndexInt idx = IDX(i, j, k); Real* v = vel + 3 * idx;
		if (!(flags[idx] & TypeEmpty)) continue;
		if ((flags[idx - X] & TypeFluid) && ghostFluidWasClamped(idx - X, +X, phi, gfClamp)) v[0] = vel[3 * (idx - X) + 0];
		if ((flags[idx - Y] & TypeFluid) && ghostFluidWasClamped(idx - Y, +Y, phi, gfClamp)) v[1] = vel[3 * (idx - Y) + 1];
		if (IS3D && (flags[idx - Z] & TypeFluid) && ghostFluidWasClamped(idx - Z, +Z, phi, gfClamp)) v[2] = vel[3 * (idx - Z) + 2];
		if ((flags[idx + X] & TypeFluid) && ghostFluidWasClamped(idx + X, -X, phi, gfClamp)) v[0] = vel[3 * (idx + X) + 0];
		if ((flags[idx + Y] & TypeFluid) && ghostFluidWasClamped(idx + Y, -Y, phi, gfClamp)) v[1] = vel[3 * (idx + Y) + 1];
		if (IS3D && (flags[idx + Z] & TypeFluid) && ghostFluidWasClamped(idx + Z, -Z, phi, gfClamp)) v[2] = vel[3 * (idx + Z) + 2];
	}
	return 0;
}

/* ---------------------------------------------------------------------------------------------
 * solvePressure plugin/pressure.cpp:480-521 = computePressureRhs + solvePressureSystem (:312-452)
 * + correctVelocity.  `solver_key` != 0 keeps the PcMGStatic hierarchy across calls like gMapMG. */
#define MAXKEYS 16
static long long g_keys[MAXKEYS]; static MgState* g_static_mg[MAXKEYS];
static MgState** static_slot(long long key)
{
	for (int i = 0; i < MAXKEYS; i++) if (g_keys[i] == key) return &g_static_mg[i];
	for (int i = 0; i < MAXKEYS; i++) if (g_keys[i] == 0) { g_keys[i] = key; return &g_static_mg[i]; }
	return 0;
}
int mfo_release_solver(long long key)
{
	for (int i = 0; i < MAXKEYS; i++) if (g_keys[i] == key) { mg_destroy(g_static_mg[i]); g_static_mg[i] = 0; g_keys[i] = 0; }
	return 0;
}

int mfo_solve_pressure(long long solver_key, int sx, int sy, int sz, const int* flags, Real* vel, Real* pressure,
	const Real* phi, const Real* perCellCorr, const Real* fractions, const Real* obvel, const Real* curv, Real* retRhs,
	double cgAccuracy, double gfClamp, double cgMaxIterFac, int precondition, int preconditioner,
	int enforceCompatibility, int useL2Norm, int zeroPressureFixing, double surfTens,
	int* iterations, double* resNorm)
{
	const IndexInt n = (IndexInt)sx * sy * sz;
	Real* rhs = (Real*)calloc((size_t)n, sizeof(Real));
	Real *A0 = (Real*)calloc((size_t)n, sizeof(Real)), *Ai = (Real*)calloc((size_t)n, sizeof(Real)), *Aj = (Real*)calloc((size_t)n, sizeof(Real)), *Ak = (Real*)calloc((size_t)n, sizeof(Real));
	mfo_compute_rhs(sx, sy, sz, flags, vel, rhs, phi, perCellCorr, fractions, obvel, curv, gfClamp, surfTens, enforceCompatibility, 0, 0);
	if (!precondition) preconditioner = 0;
	mfo_make_matrix(sx, sy, sz, flags, fractions, phi, gfClamp, A0, Ai, Aj, Ak);
	if (zeroPressureFixing || (Real)cgAccuracy < 1e-07) {
		const long long fix = mfo_choose_fix_cell(sx, sy, sz, flags);
		if (fix >= 0) mfo_fix_pressure(sx, sy, sz, fix, 0., rhs, A0, Ai, Aj, Ak);
	}
	int maxIter, pc, rc;
	const int maxDim = sx > sy ? (sx > sz ? sx : sz) : (sy > sz ? sy : sz);
	MgState* mg = 0; MgState** slot = 0;
	if (preconditioner == 0 || preconditioner == 1) {
		/* the reference asserts on PcNone here; the oracle (like the GPU build) runs plain CG */
		maxIter = (int)((Real)cgMaxIterFac * maxDim) * (IS3D ? 1 : 4);
		pc = preconditioner;
	} else if (preconditioner == 2 || preconditioner == 3) {
		maxIter = 100; pc = 2;
		if (solver_key) { slot = static_slot(solver_key); mg = slot ? *slot : 0; }
		if (mg && preconditioner == 2) { mg_destroy(mg); mg = 0; *slot = 0; }
		if (!mg) { mg = mg_create(sx, sy, sz); if (slot) *slot = mg; }
	} else { snprintf(g_err, sizeof g_err, "invalid preconditioner"); free(rhs); free(A0); free(Ai); free(Aj); free(Ak); return 1; }
	rc = cg_run(sx, sy, sz, flags, rhs, pressure, A0, Ai, Aj, Ak, pc, (Real)cgAccuracy, useL2Norm, maxIter, mg, iterations, resNorm);
	if (mg && (preconditioner == 2 || !slot)) { mg_destroy(mg); if (slot) *slot = 0; }
	if (rc == 0) mfo_correct_velocity(sx, sy, sz, flags, vel, pressure, phi, curv, gfClamp, surfTens);
	if (retRhs) memcpy(retRhs, rhs, sizeof(Real) * (size_t)n);
	free(rhs); free(A0); free(Ai); free(Aj); free(Ak);
	return rc;
}

/* =============================================================================================
 * The steps either side of the projection: wall boundary conditions with obstacle velocity,
 * gravity / buoyancy, semi-Lagrangian and MacCormack advection.  Same mixed float/double evaluation as the
 * reference (double literals promote, results narrow on assignment).
 * ============================================================================================= */

/* setWallBcs with obvel: plugin/extforces.cpp:186-218 */
int mfo_set_wall_bcs_obvel(int sx, int sy, int sz, const int* flags, Real* vel, const Real* obvel)
{
	STRIDES
	if (!obvel) return mfo_set_wall_bcs(sx, sy, sz, flags, vel);
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		const IndexInt idx = IDX(i, j, k);
		const int curFluid = flags[idx] & TypeFluid, curObs = flags[idx] & TypeObstacle;
		if (!curFluid && !curObs) continue;
		Real* v = vel + 3 * idx;
		const Real bx = obvel[3 * idx], by = obvel[3 * idx + 1], bz = IS3D ? obvel[3 * idx + 2] : (Real)0;
		if (i > 0 && (flags[idx - X] & TypeObstacle)) v[0] = bx;
		if (i > 0 && curObs && (flags[idx - X] & TypeFluid)) v[0] = bx;
		if (j > 0 && (flags[idx - Y] & TypeObstacle)) v[1] = by;
		if (j > 0 && curObs && (flags[idx - Y] & TypeFluid)) v[1] = by;
		if (!IS3D) { v[2] = 0; } else {
			if (k > 0 && (flags[idx - Z] & TypeObstacle)) v[2] = bz;
			if (k > 0 && curObs && (flags[idx - Z] & TypeFluid)) v[2] = bz;
		}
		if (curFluid) {
			if ((i > 0 && (flags[idx - X] & TypeStick)) || (i < sx - 1 && (flags[idx + X] & TypeStick))) v[1] = v[2] = 0;
			if ((j > 0 && (flags[idx - Y] & TypeStick)) || (j < sy - 1 && (flags[idx + Y] & TypeStick))) v[0] = v[2] = 0;
			if (IS3D && ((k > 0 && (flags[idx - Z] & TypeStick)) || (k < sz - 1 && (flags[idx + Z] & TypeStick)))) v[0] = v[1] = 0;
		}
	}
	return 0;
}

/* interior loop of a KERNEL(bnd=1): 2-D grids run k = 0 only */
#define FOR_BND1 const int k0_ = IS3D ? 1 : 0, k1_ = IS3D ? sz - 1 : 1; \
	for (int k = k0_; k < k1_; k++) for (int j = 1; j < sy - 1; j++) for (int i = 1; i < sx - 1; i++)

static inline int imax3(int a, int b, int c) { int m = a > b ? a : b; return m > c ? m : c; }

/* addGravity: extforces.cpp:61-65 (force vector), KnApplyForce :45-58 with additive = true */
int mfo_add_gravity(int sx, int sy, int sz, const int* flags, Real* vel, double gx, double gy, double gz, const Real* exclude, int scale, double dt_)
{
	STRIDES
	const Real dt = (Real)dt_, dx = (Real)(1.0 / imax3(sx, sy, sz));       /* grid.cpp:56 mDx */
	const float gridScale = scale ? (float)dx : 1;
	const Real g[3] = { (Real)gx, (Real)gy, (Real)gz };
	Real f[3];
	for (int c = 0; c < 3; c++) f[c] = (g[c] * dt) / gridScale;
	FOR_BND1 {
		const IndexInt idx = IDX(i, j, k);
		const int curFluid = flags[idx] & TypeFluid, curEmpty = flags[idx] & TypeEmpty;
		if (!curFluid && !curEmpty) continue;
		if (exclude && (exclude[idx] < 0.)) continue;
		Real* v = vel + 3 * idx;
		if ((flags[idx - X] & TypeFluid) || (curFluid && (flags[idx - X] & TypeEmpty))) v[0] = v[0] + f[0];
		if ((flags[idx - Y] & TypeFluid) || (curFluid && (flags[idx - Y] & TypeEmpty))) v[1] = v[1] + f[1];
		if (IS3D && ((flags[idx - Z] & TypeFluid) || (curFluid && (flags[idx - Z] & TypeEmpty)))) v[2] = v[2] + f[2];
	}
	return 0;
}

/* addBuoyancy: extforces.cpp:86-90, KnAddBuoyancy :75-83 */
int mfo_add_buoyancy(int sx, int sy, int sz, const int* flags, const Real* density, Real* vel, double gx, double gy, double gz,
                     double coefficient, int scale, double dt_)
{
	STRIDES
	const Real dt = (Real)dt_, dx = (Real)(1.0 / imax3(sx, sy, sz)), coef = (Real)coefficient;
	const float gridScale = scale ? (float)dx : 1;
	const Real g[3] = { (Real)gx, (Real)gy, (Real)gz };
	Real f[3];
	for (int c = 0; c < 3; c++) f[c] = (((-g[c]) * dt) / gridScale) * coef;
	FOR_BND1 {
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeFluid)) continue;
		Real* v = vel + 3 * idx;
		if (flags[idx - X] & TypeFluid) v[0] = (Real)((double)v[0] + (0.5 * (double)f[0]) * (double)(density[idx] + density[idx - X]));
		if (flags[idx - Y] & TypeFluid) v[1] = (Real)((double)v[1] + (0.5 * (double)f[1]) * (double)(density[idx] + density[idx - Y]));
		if (IS3D && (flags[idx - Z] & TypeFluid)) v[2] = (Real)((double)v[2] + (0.5 * (double)f[2]) * (double)(density[idx] + density[idx - Z]));
	}
	return 0;
}

/* ---- interpolation: util/interpol.h:50-91 (BUILD_INDEX, interpol, interpolComponent) */
typedef struct { IndexInt idx; Real s0, s1, t0, t1, f0, f1; } Interp;
static inline Interp build_index(int sx, int sy, int sz, const Real pos[3])
{
	const IndexInt SZ = (sz > 1) ? (IndexInt)sx * sy : 0;
	Interp o;
	const Real px = pos[0] - 0.5f, py = pos[1] - 0.5f, pz = pos[2] - 0.5f;
	int xi = (int)px, yi = (int)py, zi = (int)pz;
	o.s1 = px - (Real)xi; o.s0 = (Real)(1. - o.s1);
	o.t1 = py - (Real)yi; o.t0 = (Real)(1. - o.t1);
	o.f1 = pz - (Real)zi; o.f0 = (Real)(1. - o.f1);
	if (px < 0.) { xi = 0; o.s0 = 1.0; o.s1 = 0.0; }
	if (py < 0.) { yi = 0; o.t0 = 1.0; o.t1 = 0.0; }
	if (pz < 0.) { zi = 0; o.f0 = 1.0; o.f1 = 0.0; }
	if (xi >= sx - 1) { xi = sx - 2; o.s0 = 0.0; o.s1 = 1.0; }
	if (yi >= sy - 1) { yi = sy - 2; o.t0 = 0.0; o.t1 = 1.0; }
	if (sz > 1) { if (zi >= sz - 1) { zi = sz - 2; o.f0 = 0.0; o.f1 = 1.0; } }
	o.idx = (IndexInt)xi + (IndexInt)sx * yi + SZ * zi;
	return o;
}
/* stride = 1 for Grid<Real>, 3 (+ component offset in d) for one component of a Vec3 grid */
static inline Real interpol_s(const Real* d, int stride, int sx, int sy, int sz, const Real pos[3])
{
	const IndexInt X = stride, Y = (IndexInt)sx * stride, Z = (sz > 1) ? (IndexInt)sx * sy * stride : 0;
	const Interp q = build_index(sx, sy, sz, pos);
	const Real* p = d + q.idx * stride;
	return ((p[0] * q.t0 + p[Y] * q.t1) * q.s0 + (p[X] * q.t0 + p[X + Y] * q.t1) * q.s1) * q.f0
	     + ((p[Z] * q.t0 + p[Y + Z] * q.t1) * q.s0 + (p[X + Z] * q.t0 + p[X + Y + Z] * q.t1) * q.s1) * q.f1;
}

/* ---- MAC accessors: grid.h:424-466 */
static inline void mac_centered(const Real* v, int sx, int sy, int sz, IndexInt idx, Real out[3])
{
	const IndexInt Y = sx, Z = (sz > 1) ? (IndexInt)sx * sy : 0;
	out[0] = (Real)(0.5 * (double)(v[3 * idx] + v[3 * (idx + 1)]));
	out[1] = (Real)(0.5 * (double)(v[3 * idx + 1] + v[3 * (idx + Y) + 1]));
	out[2] = 0;
	if (sz > 1) out[2] = (Real)(0.5 * (double)(v[3 * idx + 2] + v[3 * (idx + Z) + 2]));
}
static inline void mac_at(const Real* v, int sx, int sy, int sz, IndexInt idx, int c, Real out[3])
{
	const IndexInt Y = sx, Z = (sz > 1) ? (IndexInt)sx * sy : 0;
#define Vc(o, cc) v[3 * (idx + (o)) + (cc)]
	if (c == 0) {
		out[0] = Vc(0, 0);
		out[1] = (Real)(0.25 * (double)(Vc(0, 1) + Vc(-1, 1) + Vc(Y, 1) + Vc(Y - 1, 1)));
		out[2] = 0;
		if (sz > 1) out[2] = (Real)(0.25 * (double)(Vc(0, 2) + Vc(-1, 2) + Vc(Z, 2) + Vc(Z - 1, 2)));
	} else if (c == 1) {
		out[0] = (Real)(0.25 * (double)(Vc(0, 0) + Vc(-Y, 0) + Vc(1, 0) + Vc(1 - Y, 0)));
		out[1] = Vc(0, 1);
		out[2] = 0;
		if (sz > 1) out[2] = (Real)(0.25 * (double)(Vc(0, 2) + Vc(-Y, 2) + Vc(Z, 2) + Vc(Z - Y, 2)));
	} else {
		out[0] = (Real)(0.25 * (double)(Vc(0, 0) + Vc(-Z, 0) + Vc(1, 0) + Vc(1 - Z, 0)));
		out[1] = (Real)(0.25 * (double)(Vc(0, 1) + Vc(-Z, 1) + Vc(Y, 1) + Vc(Y - Z, 1)));
		out[2] = Vc(0, 2);
	}
#undef Vc
}

/* ---- higher-order lookups: util/interpolHigh.h.  cubicInterp :22-39 is instantiated for Real and for Vec3; the Vec3 one narrows after every
 * scalar * vector product (vectorbase.h:272-279), the Real one forms the coefficients in double and narrows once. */
static inline Real cubic_interp(Real t, const Real* p, int vec)
{
	const Real d0 = (Real)((double)(p[2] - p[0]) * 0.5), d1 = (Real)((double)(p[3] - p[1]) * 0.5), dk = p[2] - p[1];
	Real a2, a3;
	if (vec) { const Real q = (Real)(3.0 * (double)dk), r = (Real)(2.0 * (double)d0), u = (Real)(-2.0 * (double)dk); a2 = (q - r) - d1; a3 = (u + d0) + d1; }
	else     { a2 = (Real)((3.0 * (double)dk - 2.0 * (double)d0) - (double)d1); a3 = (Real)((-2.0 * (double)dk + (double)d0) + (double)d1); }
	const Real sq = t * t, cu = sq * t;
	return ((a3 * cu + a2 * sq) + d0 * t) + p[1];
}
/* interpolCubic :78-172 / interpolCubic2D :42-76; stride 3 = one component of a Vec3 grid */
static Real interpol_cubic_s(const Real* d, int stride, int sx, int sy, int sz, const Real pos[3])
{
	const Real px = pos[0] - 0.5f, py = pos[1] - 0.5f, pz = pos[2] - 0.5f;
	const int x1 = (int)px, y1 = (int)py, z1 = (int)pz;
	const int is3 = sz > 1, vec = stride == 3;
	if (x1 - 1 < 0 || y1 - 1 < 0 || x1 + 2 >= sx || y1 + 2 >= sy || (is3 && (z1 - 1 < 0 || z1 + 2 >= sz))) return interpol_s(d, stride, sx, sy, sz, pos);
	const Real tx = px - x1, ty = py - y1, tz = pz - z1;
	Real zp[4];
	for (int c = 0; c < (is3 ? 4 : 1); c++) {
		Real yp[4];
		for (int b = 0; b < 4; b++) {
			Real row[4];
			for (int a = 0; a < 4; a++) row[a] = d[((IndexInt)(x1 - 1 + a) + (IndexInt)sx * (y1 - 1 + b) + (is3 ? (IndexInt)sx * sy * (z1 - 1 + c) : 0)) * stride];
			yp[b] = cubic_interp(tx, row, vec);
		}
		zp[c] = cubic_interp(ty, yp, vec);
	}
	return is3 ? cubic_interp(tz, zp, vec) : zp[0];
}
/* Grid<Real>::getInterpolatedHi grid.h:146-151 */
static inline Real lookup_real(const Real* src, int sx, int sy, int sz, const Real pos[3], int orderSpace)
{
	return orderSpace == 1 ? interpol_s(src, 1, sx, sy, sz, pos) : interpol_cubic_s(src, 1, sx, sy, sz, pos);
}
/* MACGrid::getInterpolatedComponentHi<c> grid.h:268-273; order 2 = interpolCubicMAC(pos)[c] util/interpolHigh.h:174-181 (position moved by half a cell along c) */
static inline Real lookup_mac(const Real* src, int sx, int sy, int sz, const Real pos[3], int c, int orderSpace)
{
	if (orderSpace == 1) return interpol_s(src + c, 3, sx, sy, sz, pos);
	if (c == 2 && sz <= 1) return 0;
	Real q[3] = { pos[0] + (Real)0, pos[1] + (Real)0, pos[2] + (Real)0 };
	q[c] = pos[c] + (Real)0.5;
	return interpol_cubic_s(src + c, 3, sx, sy, sz, q);
}
static void interpol_mac(int sx, int sy, int sz, const Real* data, const Real* pos, Real out[3]);

/* SemiLagrange (advection.cpp:25-41) and SemiLagrangeMAC (:44-77): dst interior only, rest stays as it was.  orderTrace 2 = explicit midpoint;
 * the MAC kernel then takes its velocities from src (:60-71), as the reference does. */
static void semi_lagrange_real(int sx, int sy, int sz, const Real* vel, Real* dst, const Real* src, Real dt, int orderSpace, int orderTrace)
{
	STRIDES
	FOR_BND1 {
		const IndexInt idx = IDX(i, j, k);
		Real c[3]; mac_centered(vel, sx, sy, sz, idx, c);
		if (orderTrace == 1) {
			const Real pos[3] = { (i + 0.5f) - c[0] * dt, (j + 0.5f) - c[1] * dt, (k + 0.5f) - c[2] * dt };
			dst[idx] = lookup_real(src, sx, sy, sz, pos, orderSpace);
		} else {
			const Real p0[3] = { i + 0.5f, j + 0.5f, k + 0.5f };
			Real p1[3], u[3], p2[3];
			for (int a = 0; a < 3; a++) p1[a] = p0[a] - (Real)((double)(c[a] * dt) * 0.5);
			interpol_mac(sx, sy, sz, vel, p1, u);
			for (int a = 0; a < 3; a++) p2[a] = p0[a] - u[a] * dt;
			dst[idx] = lookup_real(src, sx, sy, sz, p2, orderSpace);
		}
	}
}
static void semi_lagrange_mac(int sx, int sy, int sz, const Real* vel, Real* dst, const Real* src, Real dt, int orderSpace, int orderTrace)
{
	STRIDES
	FOR_BND1 {
		const IndexInt idx = IDX(i, j, k);
		for (int c = 0; c < 3; c++) {
			Real m[3];
			if (orderTrace == 1) {
				mac_at(vel, sx, sy, sz, idx, c, m);      /* getAtMACZ is evaluated in 2-D too (:53); its z stride is 0 there */
				const Real pos[3] = { (i + 0.5f) - m[0] * dt, (j + 0.5f) - m[1] * dt, (k + 0.5f) - m[2] * dt };
				dst[3 * idx + c] = lookup_mac(src, sx, sy, sz, pos, c, orderSpace);
			} else {
				mac_at(src, sx, sy, sz, idx, c, m);
				const Real p0[3] = { i + 0.5f, j + 0.5f, k + 0.5f };
				Real f0[3] = { p0[0], p0[1], p0[2] }, p1[3], u[3], p2[3];
				f0[c] = (Real)(c == 0 ? i : (c == 1 ? j : k));
				for (int a = 0; a < 3; a++) p1[a] = f0[a] - (Real)((double)(m[a] * dt) * 0.5);
				interpol_mac(sx, sy, sz, src, p1, u);
				for (int a = 0; a < 3; a++) p2[a] = p0[a] - u[a] * dt;
				dst[3 * idx + c] = lookup_mac(src, sx, sy, sz, p2, c, orderSpace);
			}
		}
	}
}

/* SemiLagrange<Vec3>: a cell-centred Grid<Vec3>; the position of the Real kernel, one lookup per component */
static void semi_lagrange_vec3(int sx, int sy, int sz, const Real* vel, Real* dst, const Real* src, Real dt, int orderSpace, int orderTrace)
{
	STRIDES
	FOR_BND1 {
		const IndexInt idx = IDX(i, j, k);
		Real c[3], pos[3]; mac_centered(vel, sx, sy, sz, idx, c);
		if (orderTrace == 1) { pos[0] = (i + 0.5f) - c[0] * dt; pos[1] = (j + 0.5f) - c[1] * dt; pos[2] = (k + 0.5f) - c[2] * dt; }
		else {
			const Real p0[3] = { i + 0.5f, j + 0.5f, k + 0.5f };
			Real p1[3], u[3];
			for (int a = 0; a < 3; a++) p1[a] = p0[a] - (Real)((double)(c[a] * dt) * 0.5);
			interpol_mac(sx, sy, sz, vel, p1, u);
			for (int a = 0; a < 3; a++) pos[a] = p0[a] - u[a] * dt;
		}
		for (int a = 0; a < 3; a++)
			dst[3 * idx + a] = orderSpace == 1 ? interpol_s(src + a, 3, sx, sy, sz, pos) : interpol_cubic_s(src + a, 3, sx, sy, sz, pos);
	}
}

static inline int check_flag(const int* flags, IndexInt idx) { return flags[idx] & (TypeFluid | TypeEmpty); }
static inline int iclamp(int v, int lo, int hi) { if (v < lo) return lo; if (v > hi) return hi; return v; }
#define REAL_MAX_ (sizeof(Real) == 4 ? (Real)3.402823466e+38F : (Real)1.7976931348623157e+308)

/* doClampComponent: advection.cpp:141-186 */
static Real clamp_component(int sx, int sy, int sz, const int* flags, Real dst, const Real* orig, Real fwd, const Real pos[3], const Real vel[3], int clampMode)
{
	STRIDES
	Real minv = REAL_MAX_, maxv = -REAL_MAX_;
	int haveFl = 0;
	const int numPos = clampMode == 1 ? 2 : 1;
	for (int l = 0; l < numPos; l++) {
		int cp[3];
		for (int c = 0; c < 3; c++) cp[c] = (int)(l == 0 ? pos[c] - vel[c] : pos[c] + vel[c]);
		const int i0 = iclamp(cp[0], 0, sx - 2), j0 = iclamp(cp[1], 0, sy - 2), k0 = iclamp(cp[2], 0, IS3D ? sz - 2 : 1);
		const int i1 = i0 + 1, j1 = j0 + 1, k1 = IS3D ? k0 + 1 : k0;
		const int ii[2] = { i0, i1 }, jj[2] = { j0, j1 }, kk[2] = { k0, k1 };
		for (int c = 0; c < (IS3D ? 2 : 1); c++) for (int b = 0; b < 2; b++) for (int a = 0; a < 2; a++) {
			const IndexInt q = IDX(ii[a], jj[b], kk[c]);
			if (check_flag(flags, q)) { if (orig[q] < minv) minv = orig[q]; if (orig[q] > maxv) maxv = orig[q]; haveFl = 1; }
		}
	}
	if (!haveFl) return fwd;
	if (clampMode == 1) { if (dst < minv) return minv; if (dst > maxv) return maxv; return dst; }
	if (dst < minv || dst > maxv) dst = fwd;
	return dst;
}
/* doClampComponentMAC<c>: advection.cpp:191-235 */
static Real clamp_component_mac(int sx, int sy, int sz, const int* flags, int c, Real dst, const Real* orig, Real fwd, int i, int j, int k, const Real vel[3], int clampMode)
{
	STRIDES
	Real minv = REAL_MAX_, maxv = -REAL_MAX_;
	const Real pos[3] = { (Real)i, (Real)j, (Real)k };
	const int numPos = clampMode == 1 ? 2 : 1;
	if (clampMode == 2) {
		int nb[3] = { i, j, k }; nb[c] -= 1;
		if (!(check_flag(flags, IDX(i, j, k)) && check_flag(flags, IDX(nb[0], nb[1], nb[2])))) return fwd;
	}
	for (int l = 0; l < numPos; l++) {
		int cp[3];
		for (int d = 0; d < 3; d++) cp[d] = (int)(l == 0 ? pos[d] - vel[d] : pos[d] + vel[d]);
		const int i0 = iclamp(cp[0], 0, sx - 2), j0 = iclamp(cp[1], 0, sy - 2), k0 = iclamp(cp[2], 0, IS3D ? sz - 2 : 0);
		const int i1 = i0 + 1, j1 = j0 + 1, k1 = IS3D ? k0 + 1 : k0;
		const int ii[2] = { i0, i1 }, jj[2] = { j0, j1 }, kk[2] = { k0, k1 };
		for (int cc = 0; cc < (IS3D ? 2 : 1); cc++) for (int b = 0; b < 2; b++) for (int a = 0; a < 2; a++) {
			const Real o = orig[3 * IDX(ii[a], jj[b], kk[cc]) + c];
			if (o < minv) minv = o;
			if (o > maxv) maxv = o;
		}
	}
	if (clampMode == 1) { if (dst < minv) return minv; if (dst > maxv) return maxv; return dst; }
	if (dst < minv || dst > maxv) dst = fwd;
	return dst;
}

/* applyOutflowBC: advection.cpp:323-392 (getBulkVel, extrapolateVelConvectiveBC, copyChangedVels) */
static void apply_outflow_bc(int sx, int sy, int sz, const int* flags, Real* vel, const Real* velPrev, double dt)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	const double ts_d = 1.0 > dt * 4 ? 1.0 : dt * 4;
	const Real timeStep = (Real)ts_d;
	Real* velDst = (Real*)calloc((size_t)n * 3, sizeof(Real));
	const int dim = IS3D ? 3 : 2, nmax = IS3D ? 1 : 0;
	const int size[3] = { sx, sy, sz };
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		const IndexInt idx = IDX(i, j, k);
		if (!(flags[idx] & TypeOutflow)) continue;
		Real avg[3] = { 0, 0, 0 }; int count = 0;
		for (int nn = -nmax; nn <= nmax; nn++) for (int m = -1; m <= 1; m++) for (int l = -1; l <= 1; l++) {
			const int a = i + l, b = j + m, c = k + nn;
			if (a < 0 || b < 0 || c < 0 || a >= sx || b >= sy || c >= sz) continue;
			const IndexInt q = IDX(a, b, c);
			if (flags[q] & (TypeFluid | TypeOutflow)) { for (int d = 0; d < 3; d++) avg[d] += vel[3 * q + d]; count++; }
		}
		if (count > 0) for (int d = 0; d < 3; d++) avg[d] = avg[d] / (Real)count;
		const int cur[3] = { i, j, k };
		int cnt = 0;
		Real* vd = velDst + 3 * idx;
		for (int c = 0; c < dim; c++) {
			int low[3] = { i, j, k }, up[3] = { i, j, k }, flLow[3] = { i, j, k }, flUp[3] = { i, j, k };
			const Real factor = timeStep * ((Real)1.0 > avg[c] ? (Real)1.0 : avg[c]);
			low[c] = flLow[c] = cur[c] - 1;
			up[c] = flUp[c] = cur[c] + 1;
			for (int d = 0; d < 2; d++) {
				const int inLo = flLow[c] >= 0 && flLow[c] < size[c], inUp = flUp[c] >= 0 && flUp[c] < size[c];
				const int fromLower = inLo && (flags[IDX(flLow[0], flLow[1], flLow[2])] & TypeFluid);
				const int fromUpper = inUp && (flags[IDX(flUp[0], flUp[1], flUp[2])] & TypeFluid);
				if (fromLower || fromUpper) {
					if (fromLower) { const IndexInt q = IDX(low[0], low[1], low[2]);
						for (int e = 0; e < 3; e++) vd[e] += ((vel[3 * idx + e] - velPrev[3 * idx + e]) / factor) + vel[3 * q + e];
						cnt++; }
					if (fromUpper) { const IndexInt q = IDX(up[0], up[1], up[2]);
						for (int e = 0; e < 3; e++) vd[e] += ((vel[3 * idx + e] - velPrev[3 * idx + e]) / factor) + vel[3 * q + e];
						cnt++; }
					break;
				}
				flLow[c]--; flUp[c]++;
			}
		}
		if (cnt > 0) for (int e = 0; e < 3; e++) vd[e] /= (Real)cnt;
	}
	for (IndexInt idx = 0; idx < n; idx++) if (flags[idx] & TypeOutflow) for (int e = 0; e < 3; e++) vel[3 * idx + e] = velDst[3 * idx + e];
	free(velDst);
}

/* advectSemiLagrange: advection.cpp:442-461, fnAdvectSemiLagrange :289-316 (Real) and :404-434 (MAC).
 * kind 0: Grid<Real> (density, level set), 1: MACGrid, 2: cell-centred Grid<Vec3>.  orderSpace 1 / 2, orderTrace 1 / 2. */
int mfo_advect_semi_lagrange(int sx, int sy, int sz, const int* flags, const Real* vel, Real* grid, int kind,
                             int order, double strength_, int orderSpace, int clampMode, int orderTrace, double dt_)
{
	STRIDES
	if (order != 1 && order != 2) { snprintf(g_err, sizeof g_err, "AdvectSemiLagrange: Only order 1 (regular SL) and 2 (MacCormack) supported"); return 1; }
	if (orderSpace != 1 && orderSpace != 2) { snprintf(g_err, sizeof g_err, "Unknown interpolation order %d", orderSpace); return 1; }
	if (orderTrace != 1 && orderTrace != 2) { snprintf(g_err, sizeof g_err, "Unknown backtracing order %d", orderTrace); return 1; }
	if (kind != 0 && kind != 1 && kind != 2) { snprintf(g_err, sizeof g_err, "oracle: grid kind %d not restated", kind); return 1; }
	const IndexInt n = (IndexInt)sx * sy * sz;
	const int nc = kind == 0 ? 1 : 3;
	const Real dt = (Real)dt_, strength = (Real)strength_;
	Real* fwd = (Real*)calloc((size_t)n * nc, sizeof(Real));
	if (kind == 0) semi_lagrange_real(sx, sy, sz, vel, fwd, grid, dt, orderSpace, orderTrace);
	else if (kind == 1) semi_lagrange_mac(sx, sy, sz, vel, fwd, grid, dt, orderSpace, orderTrace);
	else semi_lagrange_vec3(sx, sy, sz, vel, fwd, grid, dt, orderSpace, orderTrace);
	if (order == 1) {
		if (kind == 1) apply_outflow_bc(sx, sy, sz, flags, fwd, grid, (double)dt);      /* MAC grids only */
		memcpy(grid, fwd, sizeof(Real) * (size_t)n * nc);
		free(fwd);
		return 0;
	}
	Real* bwd = (Real*)calloc((size_t)n * nc, sizeof(Real));
	Real* neu = (Real*)calloc((size_t)n * nc, sizeof(Real));
	if (kind == 0) semi_lagrange_real(sx, sy, sz, vel, bwd, fwd, -dt, orderSpace, orderTrace);
	else if (kind == 1) semi_lagrange_mac(sx, sy, sz, vel, bwd, fwd, -dt, orderSpace, orderTrace);
	else semi_lagrange_vec3(sx, sy, sz, vel, bwd, fwd, -dt, orderSpace, orderTrace);
	const double half = (double)strength * 0.5;
	if (kind == 0) {
		/* MacCormackCorrect :81-91 (all cells) */
		for (IndexInt idx = 0; idx < n; idx++) {
			neu[idx] = fwd[idx];
			if (flags[idx] & TypeFluid) neu[idx] = (Real)((double)neu[idx] + half * (double)(grid[idx] - bwd[idx]));
		}
		/* MacCormackClamp :241-267 */
		FOR_BND1 {
			const IndexInt idx = IDX(i, j, k);
			Real c[3]; mac_centered(vel, sx, sy, sz, idx, c);
			const Real v[3] = { c[0] * dt, c[1] * dt, c[2] * dt };
			const Real pos[3] = { (Real)i, (Real)j, (Real)k };
			Real dval = clamp_component(sx, sy, sz, flags, neu[idx], grid, fwd[idx], pos, v, clampMode);
			if (clampMode == 1) {
				int pf[3], pb[3];
				for (int d = 0; d < 3; d++) { pf[d] = (int)((pos[d] + (Real)0.5) - v[d]); pb[d] = (int)((pos[d] + (Real)0.5) + v[d]); }
				const int ux = sx - 1, uy = sy - 1, uz = sz - 1;
				int bad = pf[0] < 0 || pf[1] < 0 || pf[2] < 0 || pb[0] < 0 || pb[1] < 0 || pb[2] < 0 ||
				          pf[0] > ux || pf[1] > uy || ((pf[2] > uz) && IS3D) || pb[0] > ux || pb[1] > uy || ((pb[2] > uz) && IS3D);
				if (!bad) bad = (flags[IDX(pf[0], pf[1], pf[2])] & TypeObstacle) || (flags[IDX(pb[0], pb[1], pb[2])] & TypeObstacle);
				if (bad) dval = fwd[idx];
			}
			neu[idx] = dval;
		}
	} else if (kind == 2) {
		/* MacCormackCorrect<Vec3> :81-91: the correction is narrowed before it is added (double * Vec3, then Vec3 += Vec3) */
		for (IndexInt idx = 0; idx < n; idx++) for (int c = 0; c < 3; c++) {
			const IndexInt q = 3 * idx + c;
			neu[q] = fwd[q];
			if (flags[idx] & TypeFluid) neu[q] = neu[q] + (Real)(half * (double)(grid[q] - bwd[q]));
		}
		/* MacCormackClamp<Vec3> :241-267 with doClampComponent<Vec3> :141-186 (getMinMax / cmpMinMax<Vec3> :119-136, clamp<Vec3> vectorbase.h:605-609) */
		FOR_BND1 {
			const IndexInt idx = IDX(i, j, k);
			Real cv[3]; mac_centered(vel, sx, sy, sz, idx, cv);
			const Real v[3] = { cv[0] * dt, cv[1] * dt, cv[2] * dt };
			const Real pos[3] = { (Real)i, (Real)j, (Real)k };
			Real minv[3] = { REAL_MAX_, REAL_MAX_, REAL_MAX_ }, maxv[3] = { -REAL_MAX_, -REAL_MAX_, -REAL_MAX_ };
			int haveFl = 0;
			for (int l = 0; l < (clampMode == 1 ? 2 : 1); l++) {
				int cp[3];
				for (int a = 0; a < 3; a++) cp[a] = (int)(l == 0 ? pos[a] - v[a] : pos[a] + v[a]);
				const int i0 = iclamp(cp[0], 0, sx - 2), j0 = iclamp(cp[1], 0, sy - 2), k0 = iclamp(cp[2], 0, IS3D ? sz - 2 : 1);
				const int k1 = IS3D ? k0 + 1 : k0;
				for (int cc = 0; cc < (IS3D ? 2 : 1); cc++) for (int b = 0; b < 2; b++) for (int a = 0; a < 2; a++) {
					const IndexInt q = IDX(i0 + a, j0 + b, cc ? k1 : k0);
					if (!check_flag(flags, q)) continue;
					for (int c = 0; c < 3; c++) { const Real o = grid[3 * q + c]; if (o < minv[c]) minv[c] = o; if (o > maxv[c]) maxv[c] = o; }
					haveFl = 1;
				}
			}
			Real* dv = neu + 3 * idx; const Real* f = fwd + 3 * idx;
			if (!haveFl) { dv[0] = f[0]; dv[1] = f[1]; dv[2] = f[2]; }
			else if (clampMode == 1) { for (int c = 0; c < 3; c++) dv[c] = dv[c] < minv[c] ? minv[c] : (dv[c] > maxv[c] ? maxv[c] : dv[c]); }
			else if (dv[0] < minv[0] || dv[0] > maxv[0] || dv[1] < minv[1] || dv[1] > maxv[1] || dv[2] < minv[2] || dv[2] > maxv[2]) { dv[0] = f[0]; dv[1] = f[1]; dv[2] = f[2]; }
			if (clampMode == 1) {
				int pf[3], pb[3];
				for (int d = 0; d < 3; d++) { pf[d] = (int)((pos[d] + (Real)0.5) - v[d]); pb[d] = (int)((pos[d] + (Real)0.5) + v[d]); }
				const int ux = sx - 1, uy = sy - 1, uz = sz - 1;
				int bad = pf[0] < 0 || pf[1] < 0 || pf[2] < 0 || pb[0] < 0 || pb[1] < 0 || pb[2] < 0 ||
				          pf[0] > ux || pf[1] > uy || ((pf[2] > uz) && IS3D) || pb[0] > ux || pb[1] > uy || ((pb[2] > uz) && IS3D);
				if (!bad) bad = (flags[IDX(pf[0], pf[1], pf[2])] & TypeObstacle) || (flags[IDX(pb[0], pb[1], pb[2])] & TypeObstacle);
				if (bad) { dv[0] = f[0]; dv[1] = f[1]; dv[2] = f[2]; }
			}
		}
	} else {
		/* MacCormackCorrectMAC :94-117 (all cells, isMAC) */
		for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
			const IndexInt idx = IDX(i, j, k);
			int skip[3] = { 0, 0, 0 };
			if (!(flags[idx] & TypeFluid)) skip[0] = skip[1] = skip[2] = 1;
			if (i > 0 && !(flags[idx - X] & TypeFluid)) skip[0] = 1;
			if (j > 0 && !(flags[idx - Y] & TypeFluid)) skip[1] = 1;
			if (k > 0 && !(flags[idx - Z] & TypeFluid)) skip[2] = 1;
			for (int c = 0; c < 3; c++) {
				const IndexInt q = 3 * idx + c;
				neu[q] = skip[c] ? fwd[q] : (Real)((double)fwd[q] + half * (double)(grid[q] - bwd[q]));
			}
		}
		/* MacCormackClampMAC :270-287 */
		FOR_BND1 {
			const IndexInt idx = IDX(i, j, k);
			for (int c = 0; c < (IS3D ? 3 : 2); c++) {
				Real m[3]; mac_at(vel, sx, sy, sz, idx, c, m);
				const Real v[3] = { m[0] * dt, m[1] * dt, m[2] * dt };
				neu[3 * idx + c] = clamp_component_mac(sx, sy, sz, flags, c, neu[3 * idx + c], grid, fwd[3 * idx + c], i, j, k, v, clampMode);
			}
		}
		apply_outflow_bc(sx, sy, sz, flags, neu, grid, (double)dt);
	}
	memcpy(grid, neu, sizeof(Real) * (size_t)n * nc);
	free(fwd); free(bwd); free(neu);
	return 0;
}

/* ---------------------------------------------------------------------------------------------
 * cgSolveWE plugin/waves.cpp:86-147 (implicit wave equation step, another GridCg caller);
 * MakeRhsWE :72-80.  ut / utm1 / out are updated as the plugin leaves them: utm1 <- old ut, ut <- out.        */
int mfo_cg_solve_we(int sx, int sy, int sz, const int* flags, Real* ut, Real* utm1, Real* out, int crankNic, double cSqr_, double cgMaxIterFac,
                    double cgAccuracy, double dt_)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	const Real dt = (Real)dt_, cSqr = (Real)cSqr_;
	const Real s = (Real)((double)(dt * dt * cSqr) * 0.5);                         /* :111 */
	Real *A0 = (Real*)calloc((size_t)n, sizeof(Real)), *Ai = (Real*)calloc((size_t)n, sizeof(Real)), *Aj = (Real*)calloc((size_t)n, sizeof(Real)), *Ak = (Real*)calloc((size_t)n, sizeof(Real));
	Real* rhs = (Real*)calloc((size_t)n, sizeof(Real));
	memset(out, 0, sizeof(Real) * (size_t)n);                                      /* out.clear() :106 */
	mfo_make_matrix(sx, sy, sz, flags, 0, 0, 0., A0, Ai, Aj, Ak);
	for (IndexInt q = 0; q < n; q++) { Ai[q] *= s; Aj[q] *= s; Ak[q] *= s; A0[q] *= s; A0[q] = (Real)((double)A0[q] + 1.); }      /* :112-118 */
	FOR_BND1 {
		const IndexInt idx = IDX(i, j, k);
		rhs[idx] = (Real)(2. * (double)ut[idx] - (double)utm1[idx]);
		if (crankNic) rhs[idx] = (Real)((double)rhs[idx] + (double)s * (-4. * (double)ut[idx] + 1. * (double)ut[idx - X] + 1. * (double)ut[idx + X] + 1. * (double)ut[idx - Y] + 1. * (double)ut[idx + Y]));
	}
	const int maxDim = sx > sy ? (sx > sz ? sx : sz) : (sy > sz ? sy : sz);
	const int maxIter = (int)((Real)cgMaxIterFac * maxDim) * (IS3D ? 1 : 4);
	const int rc = cg_run(sx, sy, sz, flags, rhs, out, A0, Ai, Aj, Ak, 0, (Real)cgAccuracy, 1, maxIter, 0, 0, 0);    /* GridCg defaults: no preconditioner, L2 stop test */
	for (IndexInt q = 0; q < n; q++) { utm1[q] = ut[q]; ut[q] = out[q]; }           /* utm1.swap(ut); ut.copyFrom(out) :145-146 */
	free(A0); free(Ai); free(Aj); free(Ak); free(rhs);
	return rc;
}

/* ---------------------------------------------------------------------------------------------
 * PD_fluid_guiding plugin/fluidguiding.cpp:294-353: primal-dual guiding loop around solvePressure.
 * Helpers: get1DGaussianBlurKernel :30-45 (through the sparse Matrix class util/rcmatrix.h: increments <= VECTOR_EPSILON are dropped,
 * :186-187), apply1DKernelDirX/Y/Z :49-82, applySeparableKernel2D/3D :85-130, getRNorm :140-145, getEpsDual :165-168,
 * applyApproxInvM :229-239, precomputeQ :243-250, precomputeInvA :254-264, prox_f :267-272.
 * MACGrid algebra (grid.h:478-486 via grid.cpp add/sub/mult/multConst/addScaled) is componentwise in Real.                       */
#ifndef M_PI
#define M_PI 3.14159265358979323846   /* math.h value (hidden by -std=c11) */
#endif
#if MF_REAL_IS_DOUBLE
#define R_EXP exp
#define VEC_EPS 1e-10
#else
#define R_EXP expf
#define VEC_EPS 1e-6f
#endif
static void pd_blur_kernel(int n, Real* G)
{
	const int sigma = n;
	Real sumG = 0;
	for (int j = 0; j < n; j++) {
		Real xv = (Real)(-(n - 1) * 0.5), yv = (Real)(j - (n - 1) * 0.5);
		if (!(R_FABS(xv) > VEC_EPS)) xv = 0;
		if (!(R_FABS(yv) > VEC_EPS)) yv = 0;
		Real g = (Real)(1 / (2 * M_PI * sigma * sigma) * R_EXP(-(xv * xv + yv * yv) / (2 * sigma * sigma)));
		if (!(R_FABS(g) > VEC_EPS)) g = 0;
		G[j] = g;
		sumG += G[j];
	}
	const double k = 1.0 / sumG;
	for (int j = 0; j < n; j++) { Real v = (Real)(G[j] * k); if (!(R_FABS(v) > VEC_EPS)) v = 0; G[j] = v; }
}
static void pd_conv1d(int sx, int sy, int sz, int dir, const Real* in, Real* out, const Real* kernel, int kn)
{
	STRIDES
	const int size[3] = { sx, sy, sz }, kCentre = kn / 2;
	const IndexInt stride[3] = { X, Y, Z };
	#pragma omp parallel for schedule(static)
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		const IndexInt idx = IDX(i, j, k);
		const int pos[3] = { i, j, k };
		Real acc[3] = { 0, 0, 0 };
		for (int m = 0, ind = kn - 1, q = pos[dir] - kCentre; m < kn; m++, ind--, q++) {
			if (q < 0) continue;
			else if (q >= size[dir]) break;
			const Real* v = in + 3 * (idx + (IndexInt)(q - pos[dir]) * stride[dir]);
			for (int c = 0; c < 3; c++) acc[c] += v[c] * kernel[ind];
		}
		for (int c = 0; c < 3; c++) out[3 * idx + c] = acc[c];
	}
}
static void pd_blur(int sx, int sy, int sz, const int* flags, Real* grid, const Real* kernel, int kn, Real* t1, Real* t2)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	Real* orig = (Real*)malloc(sizeof(Real) * 3 * (size_t)n);
	memcpy(orig, grid, sizeof(Real) * 3 * (size_t)n);
	pd_conv1d(sx, sy, sz, 0, grid, t1, kernel, kn);
	pd_conv1d(sx, sy, sz, 1, t1, t2, kernel, kn);
	if (IS3D) { pd_conv1d(sx, sy, sz, 2, t2, t1, kernel, kn); memcpy(grid, t1, sizeof(Real) * 3 * (size_t)n); }
	else memcpy(grid, t2, sizeof(Real) * 3 * (size_t)n);
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		const IndexInt idx = IDX(i, j, k);
		if ((i > 0 && (flags[idx - X] & TypeObstacle)) || (j > 0 && (flags[idx - Y] & TypeObstacle)) || (IS3D && k > 0 && (flags[idx - Z] & TypeObstacle)) || (flags[idx] & TypeObstacle))
			for (int c = 0; c < 3; c++) grid[3 * idx + c] = orig[3 * idx + c];
	}
	free(orig);
}
static Real pd_max_abs(const Real* v, IndexInt n)      /* Grid<Vec3>::getMaxAbs grid.cpp:330-332, CompMaxVec :198-203 */
{
	Real m = -REAL_MAX_;
	for (IndexInt q = 0; q < n; q++) { const Real s = v[3 * q] * v[3 * q] + v[3 * q + 1] * v[3 * q + 1] + v[3 * q + 2] * v[3 * q + 2]; if (s > m) m = s; }
	return R_SQRT(m);
}
int mfo_pd_fluid_guiding(int sx, int sy, int sz, const int* flags, Real* vel, const Real* velT, Real* pressure, const Real* weight,
	int blurRadius, double theta_, double tau_, double sigma_, double epsRel_, double epsAbs_, int maxIters,
	double cgMaxIterFac, double cgAccuracy, int preconditioner, int zeroPressureFixing, int* iterations)
{
	const IndexInt n = (IndexInt)sx * sy * sz; const size_t n3 = 3 * (size_t)n;
	const Real theta = (Real)theta_, tau = (Real)tau_, sigma = (Real)sigma_, epsRel = (Real)epsRel_, epsAbs = (Real)epsAbs_;
	const int kn = 2 * blurRadius + 1;
	Real* G = (Real*)calloc((size_t)kn, sizeof(Real)); pd_blur_kernel(kn, G);
	Real *velC = (Real*)malloc(n3 * sizeof(Real)), *x = (Real*)calloc(n3, sizeof(Real)), *y = (Real*)calloc(n3, sizeof(Real)), *z = (Real*)calloc(n3, sizeof(Real));
	Real *x0 = (Real*)calloc(n3, sizeof(Real)), *z0 = (Real*)calloc(n3, sizeof(Real)), *Q = (Real*)malloc(n3 * sizeof(Real)), *invA = (Real*)malloc(n3 * sizeof(Real));
	Real *vn = (Real*)malloc(n3 * sizeof(Real)), *t1 = (Real*)malloc(n3 * sizeof(Real)), *t2 = (Real*)malloc(n3 * sizeof(Real));
	memcpy(velC, vel, n3 * sizeof(Real));
	/* precomputeQ */
	for (size_t q = 0; q < n3; q++) Q[q] = velT[q] - velC[q];
	pd_blur(sx, sy, sz, flags, Q, G, kn, t1, t2); pd_blur(sx, sy, sz, flags, Q, G, kn, t1, t2);
	for (size_t q = 0; q < n3; q++) { Q[q] = Q[q] * (Real)2.0; Q[q] = Q[q] + (-sigma) * velC[q]; }
	/* precomputeInvA */
	for (IndexInt q = 0; q < n; q++) {
		Real val = 2 * weight[q] * weight[q] + sigma;
		if (val < 0.01) val = (Real)0.01;
		const Real inv = (Real)(1.0 / val);
		invA[3 * q] = invA[3 * q + 1] = invA[3 * q + 2] = inv;
	}
	const Real invSigma = (Real)(1.0 / sigma);
	int iter = 0, rc = 0;
	for (iter = 0; iter < maxIters; iter++) {
		/* x-update */
		for (size_t q = 0; q < n3; q++) { x0[q] = x[q]; x[q] = x[q] * invSigma; x[q] = x[q] + y[q]; }
		/* prox_f(x) */
		for (size_t q = 0; q < n3; q++) { x[q] = x[q] * sigma; x[q] = x[q] + Q[q]; }
		/*   applyApproxInvM(x) */
		for (size_t q = 0; q < n3; q++) vn[q] = x[q] * invA[q];
		pd_blur(sx, sy, sz, flags, vn, G, kn, t1, t2); pd_blur(sx, sy, sz, flags, vn, G, kn, t1, t2);
		for (size_t q = 0; q < n3; q++) { vn[q] = vn[q] * (Real)2.0; vn[q] = vn[q] * invA[q]; x[q] = x[q] * invA[q]; x[q] = x[q] - vn[q]; }
		for (size_t q = 0; q < n3; q++) x[q] = x[q] + velC[q];
		for (size_t q = 0; q < n3; q++) { x[q] = x[q] * (-sigma); x[q] = x[q] + sigma * y[q]; x[q] = x[q] + x0[q]; }
		/* z-update */
		for (size_t q = 0; q < n3; q++) { z0[q] = z[q]; z[q] = z[q] + (-tau) * x[q]; }
		rc = mfo_solve_pressure(0, sx, sy, sz, flags, z, pressure, 0, 0, 0, 0, 0, 0, cgAccuracy, 1e-04, cgMaxIterFac, 1, preconditioner, 0, 0, zeroPressureFixing, 0., 0, 0);
		if (rc) break;
		/* y-update */
		for (size_t q = 0; q < n3; q++) { y[q] = z[q]; y[q] = y[q] - z0[q]; y[q] = y[q] * theta; y[q] = y[q] + z[q]; }
		/* stopping criterion: getRNorm(z, z0) < getEpsDual(epsAbs, epsRel, z) */
		int stop = 0;
		if (iter > 0) {
			for (size_t q = 0; q < n3; q++) t1[q] = z[q] - z0[q];
			const Real rnorm = pd_max_abs(t1, n);
			const Real epsDual = (Real)(sqrt(sz > 1 ? 3.0 : 2.0) * (double)epsAbs + (double)(epsRel * pd_max_abs(z, n)));
			stop = rnorm < epsDual;
		}
		if (stop || (iter == maxIters - 1)) break;
	}
	memcpy(vel, z, n3 * sizeof(Real));
	if (iterations) *iterations = iter;
	free(G); free(velC); free(x); free(y); free(z); free(x0); free(z0); free(Q); free(invA); free(vn); free(t1); free(t2);
	return rc;
}

/* =============================================================================================
 * Liquid neighbours (first slice): the functions that make the level-set free-surface loop of
 * scenes/freesurface.py:54-84 (useMarching = False) device-resident around solvePressure(phi=...).            */

static const int NB6[6][3] = { { 1, 0, 0 }, { -1, 0, 0 }, { 0, 1, 0 }, { 0, -1, 0 }, { 0, 0, 1 }, { 0, 0, -1 } };   /* fastmarch.cpp:233-236, :433-436 */

/* FlagGrid::updateFromLevelset grid.cpp:844-854; invalidTimeValue = -1000 (levelset.cpp:103 -> fastmarch.h:134) */
int mfo_update_from_levelset(int sx, int sy, int sz, int* flags, const Real* phi)
{
	const IndexInt n = (IndexInt)sx * sy * sz;
	for (IndexInt idx = 0; idx < n; idx++) {
		if ((flags[idx] & TypeObstacle) || (flags[idx] & TypeOutflow)) continue;
		const Real p = phi[idx];
		if (p <= (Real)-1000) continue;
		flags[idx] &= ~(TypeEmpty | TypeFluid);
		flags[idx] |= (p <= 0) ? TypeFluid : TypeEmpty;
	}
	return 0;
}

/* Grid<T>::setBound grid.cpp:591-593, knSetBoundary :585-589 (ncomp 1: Grid<Real>, 3: Grid<Vec3>; the same value in every component) */
int mfo_set_bound(int sx, int sy, int sz, Real* grid, int ncomp, double value, int w)
{
	STRIDES
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		const int bnd = (i <= w || i >= sx - 1 - w || j <= w || j >= sy - 1 - w || (IS3D && (k <= w || k >= sz - 1 - w)));
		if (bnd) for (int c = 0; c < ncomp; c++) grid[ncomp * IDX(i, j, k) + c] = (Real)value;
	}
	return 0;
}

/* normalize util/vectorbase.h:415-429 (the comparisons against 1. and the reciprocal are double expressions) */
static Real normalize3(Real v[3])
{
	Real norm;
	const Real l = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
#if MF_REAL_IS_DOUBLE
	const double eps2 = 1e-10 * 1e-10;            /* VECTOR_EPSILON vectorbase.h:55 */
#else
	const float eps2 = 1e-6f * 1e-6f;             /* vectorbase.h:52 */
#endif
	if (fabs((double)l - 1.) < eps2) norm = 1.;
	else if (l > eps2) {
		norm = R_SQRT(l);
		const Real s = (Real)(1. / (double)norm);
		v[0] *= s; v[1] *= s; v[2] *= s;
	} else { v[0] = v[1] = v[2] = 0; norm = 0.; }
	return norm;
}

/* extrapolateMACSimple fastmarch.cpp:337-375: knExtrapolateMACSimple :231-258, knUnprojectNormalComp :319-331 (getNormal :302-318),
 * knExtrapolateIntoBnd :260-299 */
int mfo_extrapolate_mac_simple(int sx, int sy, int sz, const int* flags, Real* vel, int distance, const Real* phiObs, int intoObs)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	const int dim = IS3D ? 3 : 2;
	if (sx < 3 || sy < 3 || (IS3D && sz < 3)) { snprintf(g_err, sizeof g_err, "extrapolateMACSimple: grid without interior cells"); return 1; }
	int* tmp = (int*)malloc(sizeof(int) * (size_t)n);
	for (int c = 0; c < dim; c++) {
		const IndexInt dir = c == 0 ? X : (c == 1 ? Y : Z);
		memset(tmp, 0, sizeof(int) * (size_t)n);
		{ FOR_BND1 {
			const IndexInt p = IDX(i, j, k);
			int mark = 0;
			if (!intoObs) { if ((flags[p] & TypeFluid) || (flags[p - dir] & TypeFluid)) mark = 1; }
			else { if (((flags[p] & TypeFluid) || (flags[p - dir] & TypeFluid)) && !(flags[p] & TypeObstacle) && !(flags[p - dir] & TypeObstacle)) mark = 1; }
			if (mark) tmp[p] = 1;
		} }
		for (int d = 1; d < 1 + distance; d++) {
			FOR_BND1 {
				const IndexInt p = IDX(i, j, k);
				if (tmp[p] != 0) continue;
				int nbs = 0; Real avgVel = 0.;
				for (int q = 0; q < 2 * dim; q++) {
					const IndexInt pn = IDX(i + NB6[q][0], j + NB6[q][1], k + NB6[q][2]);
					if (tmp[pn] == d) { avgVel += vel[3 * pn + c]; nbs++; }
				}
				if (nbs > 0) { tmp[p] = d + 1; vel[3 * p + c] = avgVel / nbs; }
			}
		}
	}
	free(tmp);
	if (phiObs) {
		const Real maxDist = (Real)distance;
		FOR_BND1 {
			const IndexInt p = IDX(i, j, k);
			if (phiObs[p] > 0. || phiObs[p] < -maxDist) continue;
			Real nrm[3] = { phiObs[p + X] - phiObs[p - X], phiObs[p + Y] - phiObs[p - Y], phiObs[p + Z] - phiObs[p - Z] };   /* 2-D: Z == 0 -> 0 */
			Real* v = vel + 3 * p;
			if (nrm[0] * v[0] + nrm[1] * v[1] + nrm[2] * v[2] < 0.) {
				normalize3(nrm);
				const Real l = nrm[0] * v[0] + nrm[1] * v[1] + nrm[2] * v[2];
				v[0] -= nrm[0] * l; v[1] -= nrm[1] * l; v[2] -= nrm[2] * l;
			}
		}
	}
	Real* velTmp = (Real*)malloc(sizeof(Real) * 3 * (size_t)n);
	memcpy(velTmp, vel, sizeof(Real) * 3 * (size_t)n);
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		const IndexInt p = IDX(i, j, k);
		int c = 0;
		Real v[3] = { 0, 0, 0 };
		const int isObs = flags[p] & TypeObstacle;
		if (i == 0)           { memcpy(v, velTmp + 3 * (p + X), sizeof v); if (isObs && v[0] < 0.) v[0] = 0.; c++; }
		else if (i == sx - 1) { memcpy(v, velTmp + 3 * (p - X), sizeof v); if (isObs && v[0] > 0.) v[0] = 0.; c++; }
		if (j == 0)           { memcpy(v, velTmp + 3 * (p + Y), sizeof v); if (isObs && v[1] < 0.) v[1] = 0.; c++; }
		else if (j == sy - 1) { memcpy(v, velTmp + 3 * (p - Y), sizeof v); if (isObs && v[1] > 0.) v[1] = 0.; c++; }
		if (IS3D) {
			if (k == 0)           { memcpy(v, velTmp + 3 * (p + Z), sizeof v); if (isObs && v[2] < 0.) v[2] = 0.; c++; }
			else if (k == sz - 1) { memcpy(v, velTmp + 3 * (p - Z), sizeof v); if (isObs && v[2] > 0.) v[2] = 0.; c++; }
		}
		if (c > 0) { vel[3 * p] = v[0] / (Real)c; vel[3 * p + 1] = v[1] / (Real)c; vel[3 * p + 2] = v[2] / (Real)c; }
	}
	free(velTmp);
	return 0;
}

/* marks of extrapolateLsSimple / extrapolateVec3Simple fastmarch.cpp:475-498, :516-536: 1 on the chosen side of phi, 2 on the first layer next to it */
static int* ls_marks(int sx, int sy, int sz, const Real* phi, int inside)
{
	STRIDES
	const int dim = IS3D ? 3 : 2;
	int* tmp = (int*)calloc((size_t)sx * sy * sz, sizeof(int));
	{ FOR_BND1 { const IndexInt p = IDX(i, j, k); if (!inside ? (phi[p] < 0.) : (phi[p] > 0.)) tmp[p] = 1; } }
	{ FOR_BND1 {
		const IndexInt p = IDX(i, j, k);
		if (tmp[p]) continue;
		for (int q = 0; q < 2 * dim; q++) if (tmp[IDX(i + NB6[q][0], j + NB6[q][1], k + NB6[q][2])] == 1) { tmp[p] = 2; break; }
	} }
	return tmp;
}

/* knExtrapolateLsSimple<S> fastmarch.cpp:439-460 for d = 2 .. distance, then knSetRemaining :463-467 */
static void ls_extrapolate(int sx, int sy, int sz, Real* val, int ncomp, int* tmp, int distance, Real direction, Real remaining)
{
	STRIDES
	const int dim = IS3D ? 3 : 2;
	for (int d = 2; d < 1 + distance; d++) {
		FOR_BND1 {
			const IndexInt p = IDX(i, j, k);
			if (tmp[p] != 0) continue;
			int nbs = 0; Real avg[3] = { 0, 0, 0 };
			for (int q = 0; q < 2 * dim; q++) {
				const IndexInt pn = IDX(i + NB6[q][0], j + NB6[q][1], k + NB6[q][2]);
				if (tmp[pn] == d) { for (int c = 0; c < ncomp; c++) avg[c] += val[ncomp * pn + c]; nbs++; }
			}
			if (nbs > 0) { tmp[p] = d + 1; for (int c = 0; c < ncomp; c++) val[ncomp * p + c] = avg[c] / nbs + direction; }
		}
	}
	{ FOR_BND1 { const IndexInt p = IDX(i, j, k); if (tmp[p] == 0) for (int c = 0; c < ncomp; c++) val[ncomp * p + c] = remaining; } }
}

/* extrapolateLsSimple fastmarch.cpp:470-507 */
int mfo_extrapolate_ls_simple(int sx, int sy, int sz, Real* phi, int distance, int inside)
{
	if (sx < 3 || sy < 3 || (IS3D && sz < 3)) { snprintf(g_err, sizeof g_err, "extrapolateLsSimple: grid without interior cells"); return 1; }
	const Real direction = inside ? -1. : 1.;
	int* tmp = ls_marks(sx, sy, sz, phi, inside);
	ls_extrapolate(sx, sy, sz, phi, 1, tmp, distance, direction, (Real)(direction * (distance + 2)));
	free(tmp);
	return 0;
}

/* extrapolateVec3Simple fastmarch.cpp:510-542 */
int mfo_extrapolate_vec3_simple(int sx, int sy, int sz, Real* vel, const Real* phi, int distance, int inside)
{
	if (sx < 3 || sy < 3 || (IS3D && sz < 3)) { snprintf(g_err, sizeof g_err, "extrapolateVec3Simple: grid without interior cells"); return 1; }
	int* tmp = ls_marks(sx, sy, sz, phi, inside);
	ls_extrapolate(sx, sy, sz, vel, 3, tmp, distance, (Real)0, (Real)0);
	free(tmp);
	return 0;
}

/* getLaplacian plugin/flip.cpp:710-712 -> LaplaceOp commonkernels.h:75-80; getCurvature flip.cpp:714-716 -> CurvatureOp commonkernels.h:83-101.
 * The double literals promote every product to double; each named Real narrows. */
int mfo_get_laplacian(int sx, int sy, int sz, Real* laplace, const Real* grid)
{
	STRIDES
	FOR_BND1 {
		const IndexInt p = IDX(i, j, k);
		laplace[p] = (Real)(((double)grid[p + X] - 2.0 * (double)grid[p]) + (double)grid[p - X]);
		laplace[p] = (Real)((double)laplace[p] + (((double)grid[p + Y] - 2.0 * (double)grid[p]) + (double)grid[p - Y]));
		if (IS3D) laplace[p] = (Real)((double)laplace[p] + (((double)grid[p + Z] - 2.0 * (double)grid[p]) + (double)grid[p - Z]));
	}
	return 0;
}
int mfo_get_curvature(int sx, int sy, int sz, Real* curv, const Real* grid, double h_)
{
	STRIDES
	const Real h = (Real)h_;
	FOR_BND1 {
		const IndexInt p = IDX(i, j, k);
		const Real over_h = (Real)(1.0 / (double)h);
		const double oh = (double)over_h, g2 = 2.0 * (double)grid[p];
		const Real x = (Real)((0.5 * (double)(grid[p + X] - grid[p - X])) * oh);
		const Real y = (Real)((0.5 * (double)(grid[p + Y] - grid[p - Y])) * oh);
		const Real xx = (Real)(((((double)grid[p + X] - g2) + (double)grid[p - X]) * oh) * oh);
		const Real yy = (Real)(((((double)grid[p + Y] - g2) + (double)grid[p - Y]) * oh) * oh);
		const Real xy = (Real)(((0.25 * (double)(((grid[p + X + Y] + grid[p - X - Y]) - grid[p - X + Y]) - grid[p + X - Y])) * oh) * oh);
		Real c = (Real)(((double)(x * x * yy + y * y * xx)) - (2.0 * (double)x) * (double)y * (double)xy);
		Real denom = x * x + y * y;
		if (IS3D) {
			const Real z = (Real)((0.5 * (double)(grid[p + Z] - grid[p - Z])) * oh);
			const Real zz = (Real)(((((double)grid[p + Z] - g2) + (double)grid[p - Z]) * oh) * oh);
			const Real xz = (Real)(((0.25 * (double)(((grid[p + X + Z] + grid[p - X - Z]) - grid[p - X + Z]) - grid[p + X - Z])) * oh) * oh);
			const Real yz = (Real)(((0.25 * (double)(((grid[p + Y + Z] + grid[p - Y - Z]) - grid[p + Y - Z]) - grid[p - Y + Z])) * oh) * oh);
			c = (Real)((double)c + ((double)(x * x * zz + z * z * xx + y * y * zz + z * z * yy) - 2.0 * (double)(x * z * xz + y * z * yz)));
			denom += z * z;
		}
#if MF_REAL_IS_DOUBLE
		const Real dmax = denom > 1e-10 ? denom : 1e-10;
#else
		const Real dmax = denom > 1e-6f ? denom : 1e-6f;
#endif
		curv[p] = (Real)((double)c / pow((double)dmax, 1.5));
	}
	return 0;
}

/* setWallBcs with fractions + phiObs: plugin/extforces.cpp:307-316 -> KnSetWallBcsFrac :220-303 (second-order obstacle boundaries; the
 * kernel writes a fresh grid that is then swapped in).  mac_at = MACGrid::getAtMACX/Y/Z grid.h:437-470, normalize3 as above. */
static inline Real half_sum(Real a, Real b) { return (Real)((double)(a + b) * .5); }
int mfo_set_wall_bcs_frac(int sx, int sy, int sz, const int* flags, Real* vel, const Real* phiObs)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	Real* tgt = (Real*)malloc(sizeof(Real) * 3 * (size_t)n);
	memcpy(tgt, vel, sizeof(Real) * 3 * (size_t)n);                  /* velTarget(i,j,k) = vel(i,j,k) for every cell */
	const IndexInt S[3] = { X, Y, Z };
	FOR_BND1 {
		const IndexInt p = IDX(i, j, k);
		const int curFluid = flags[p] & TypeFluid, curObs = flags[p] & TypeObstacle;
		if (!curFluid && !curObs) continue;
		const int dim = IS3D ? 3 : 2;
		for (int c = 0; c < dim; c++) {
			if (!(curObs || (flags[p - S[c]] & TypeObstacle))) continue;
			Real dphi[3] = { 0, 0, 0 };
			const Real tmp1 = half_sum(phiObs[p], phiObs[p - S[c]]);
			for (int a = 0; a < dim; a++) {
				if (a == c) { dphi[a] = phiObs[p] - phiObs[p - S[c]]; continue; }
				Real tmp2 = half_sum(phiObs[p + S[a]], phiObs[p + S[a] - S[c]]);
				const Real phi1 = half_sum(tmp1, tmp2);
				tmp2 = half_sum(phiObs[p - S[a]], phiObs[p - S[a] - S[c]]);
				const Real phi2 = half_sum(tmp1, tmp2);
				dphi[a] = phi1 - phi2;
			}
			normalize3(dphi);
			Real vm[3];
			mac_at(vel, sx, sy, sz, p, c, vm);
			tgt[3 * p + c] = vm[c] - (dphi[0] * vm[0] + dphi[1] * vm[1] + dphi[2] * vm[2]) * dphi[c];
		}
	}
	memcpy(vel, tgt, sizeof(Real) * 3 * (size_t)n);
	free(tgt);
	return 0;
}

/* extrapolateMACFromWeight fastmarch.cpp:410-432 (knExtrapolateMACFromWeight :378-403): like extrapolateMACSimple, but the marks live in a
 * Vec3 weight grid (what mapPartsToMAC leaves behind): > 0 becomes 1 = initialised, pass d turns reached cells into d+1.  The weight grid is destroyed. */
int mfo_extrapolate_mac_from_weight(int sx, int sy, int sz, Real* vel, Real* weight, int distance)
{
	STRIDES
	const int dim = IS3D ? 3 : 2;
	if (sx < 3 || sy < 3 || (IS3D && sz < 3)) { snprintf(g_err, sizeof g_err, "extrapolateMACFromWeight: grid without interior cells"); return 1; }
	for (int c = 0; c < dim; c++) {
		{ FOR_BND1 { const IndexInt p = IDX(i, j, k); if (weight[3 * p + c] > 0.) weight[3 * p + c] = 1.0; } }
		for (int d = 1; d < 1 + distance; d++) {
			FOR_BND1 {
				const IndexInt p = IDX(i, j, k);
				if (weight[3 * p + c] != 0) continue;
				int nbs = 0; Real avgVel = 0.;
				for (int q = 0; q < 2 * dim; q++) {
					const IndexInt pn = IDX(i + NB6[q][0], j + NB6[q][1], k + NB6[q][2]);
					if (weight[3 * pn + c] == d) { avgVel += vel[3 * pn + c]; nbs++; }
				}
				if (nbs > 0) { weight[3 * p + c] = d + 1; vel[3 * p + c] = avgVel / nbs; }
			}
		}
	}
	return 0;
}

/* updateFractions plugin/initplugins.cpp:437-440 (KnUpdateFractions :371-434, calcFraction :356-369) and setObstacleFlags :473-475
 * (KnUpdateFlagsObs :443-470): the producers of the `fractions` argument of solvePressure / setWallBcs (second-order obstacle boundaries).
 * KnUpdateFractions writes neighbour cells on the max sides; this is the SERIAL order of the loop (with OpenMP the reference races there
 * for boundaryWidth > 0).  The max-z test reads j, as the reference does (:423). */
static inline Real calc_fraction(Real phi1, Real phi2, Real fracThreshold)
{
	if (phi1 > 0. && phi2 > 0.) return 1.;
	if (phi1 < 0. && phi2 < 0.) return 0.;
	if (phi2 < phi1) { Real t = phi1; phi1 = phi2; phi2 = t; }
	const Real denom = phi1 - phi2;
	if (denom > -1e-04) return 0.5;
	Real frac = (Real)(1. - (double)(phi1 / denom));
	if (frac < fracThreshold) frac = 0.;
	return rmin((Real)1, frac);
}
static inline int is_open_kind(int f) { return (f & TypeInflow) || (f & TypeOutflow) || (f & TypeOpen); }
int mfo_update_fractions(int sx, int sy, int sz, const int* flags, const Real* phiObs, Real* fractions, int boundaryWidth, double fracThreshold_)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	const Real thr = (Real)fracThreshold_;
	const int w = boundaryWidth;
	for (IndexInt q = 0; q < 3 * n; q++) fractions[q] = 0;
	FOR_BND1 {
		const IndexInt p = IDX(i, j, k);
		Real* f = fractions + 3 * p;
		f[0] = calc_fraction(phiObs[p], phiObs[p - X], thr);
		f[1] = calc_fraction(phiObs[p], phiObs[p - Y], thr);
		if (IS3D) f[2] = calc_fraction(phiObs[p], phiObs[p - Z], thr);
		if (phiObs[p] < 0.) continue;
#define SET1(cell) { Real* g_ = fractions + 3 * (cell); g_[0] = g_[1] = 1.; if (IS3D) g_[2] = 1.; }
		if (i <= w + 1 && is_open_kind(flags[p - X])) SET1(p)
		if (i >= sx - w - 2 && is_open_kind(flags[p + X])) SET1(p + X)
		if (j <= w + 1 && is_open_kind(flags[p - Y])) SET1(p)
		if (j >= sy - w - 2 && is_open_kind(flags[p + Y])) SET1(p + Y)
		if (IS3D) {
			if (k <= w + 1 && is_open_kind(flags[p - Z])) SET1(p)
			if (j >= sz - w - 2 && is_open_kind(flags[p + Z])) SET1(p + Z)
		}
#undef SET1
	}
	return 0;
}
int mfo_set_obstacle_flags(int sx, int sy, int sz, int* flags, const Real* phiObs, const Real* fractions, const Real* phiOut, const Real* phiIn, int bw)
{
	STRIDES
	const int k0 = IS3D ? bw : 0, k1 = IS3D ? sz - bw : 1;
	for (int k = k0; k < k1; k++) for (int j = bw; j < sy - bw; j++) for (int i = bw; i < sx - bw; i++) {
		const IndexInt p = IDX(i, j, k);
		int isObs = 0;
		if (fractions) {
			Real f = 0.;
			f += fractions[3 * p]; f += fractions[3 * (p + X)];
			f += fractions[3 * p + 1]; f += fractions[3 * (p + Y) + 1];
			if (IS3D) { f += fractions[3 * p + 2]; f += fractions[3 * (p + Z) + 2]; }
			if (f == 0.) isObs = 1;
		} else if (phiObs[p] < 0.) isObs = 1;
		const int isOutflow = phiOut && phiOut[p] < 0., isInflow = phiIn && phiIn[p] < 0.;
		if (isObs) flags[p] = TypeObstacle;
		else if (isInflow) flags[p] = TypeFluid | TypeInflow;
		else if (isOutflow) flags[p] = TypeEmpty | TypeOutflow;
		else flags[p] = TypeEmpty;
	}
	return 0;
}

/* =============================================================================================
 * FLIP particle <-> grid plugins (mapPartsToMAC, gridParticleIndex + unionParticleLevelset are 8.7 % + 13.5 % of a
 * benchmark_dam step; markFluidCells, mapMACToParts, flipVelocityUpdate sit beside them in the loop, scenes/benchmark_dam.py:100-125).
 * Restated here ahead of the device versions ("the oracle first").  Particles are plain arrays: pos [N][3], pflag [N]
 * (BasicParticleData particle.h:182-191, active = !(flag & PDELETE)), optional ptype [N], velocities [N][3].                        */
enum { PDELETE = 1 << 10 };                                  /* particle.h:41 */
#define P_SKIP(idx) ((pflag[idx] & PDELETE) || (ptype && (ptype[idx] & exclude)))
static inline int in_bounds0(int sx, int sy, int sz, int x, int y, int z)     /* GridBase::isInBounds(p, 0) grid.h */
{ return x >= 0 && y >= 0 && x < sx && y < sy && (sz > 1 ? (z >= 0 && z < sz) : z == 0); }

/* markFluidCells plugin/flip.cpp:158-177 (knClearFluidFlags :137-141, knSetNbObstacle :142-157) */
int mfo_mark_fluid_cells(int sx, int sy, int sz, int* flags, long long np, const Real* pos, const int* pflag, const Real* phiObs, const int* ptype, int exclude)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	for (IndexInt q = 0; q < n; q++) if (flags[q] & TypeFluid) flags[q] = (flags[q] | TypeEmpty) & ~TypeFluid;
	for (long long idx = 0; idx < np; idx++) {
		if (P_SKIP(idx)) continue;
		const int x = (int)pos[3 * idx], y = (int)pos[3 * idx + 1], z = (int)pos[3 * idx + 2];
		if (!in_bounds0(sx, sy, sz, x, y, z)) continue;
		const IndexInt p = IDX(x, y, z);
		if (flags[p] & TypeEmpty) flags[p] = (flags[p] | TypeFluid) & ~TypeEmpty;
	}
	if (phiObs) {
		int* tmp = (int*)malloc(sizeof(int) * (size_t)n);
		memcpy(tmp, flags, sizeof(int) * (size_t)n);
		FOR_BND1 {
			const IndexInt p = IDX(i, j, k);
			if (phiObs[p] > 0.) continue;
			if (!(flags[p] & TypeEmpty)) continue;
			int set = 0;
			if ((flags[p - X] & TypeFluid) && (phiObs[p + X] <= 0.)) set = 1;
			if ((flags[p + X] & TypeFluid) && (phiObs[p - X] <= 0.)) set = 1;
			if ((flags[p - Y] & TypeFluid) && (phiObs[p + Y] <= 0.)) set = 1;
			if ((flags[p + Y] & TypeFluid) && (phiObs[p - Y] <= 0.)) set = 1;
			if (IS3D) {
				if ((flags[p - Z] & TypeFluid) && (phiObs[p + Z] <= 0.)) set = 1;
				if ((flags[p + Z] & TypeFluid) && (phiObs[p - Z] <= 0.)) set = 1;
			}
			if (set) tmp[p] = (flags[p] | TypeFluid) & ~TypeEmpty;
		}
		memcpy(flags, tmp, sizeof(int) * (size_t)n);
		free(tmp);
	}
	return 0;
}

/* gridParticleIndex plugin/flip.cpp:260-306: index[cell] = first slot of the cell in indexSys, indexSys[slot] = particle (cells in grid
 * order, particles of a cell in ascending order); returns the number of indexed particles in *count */
int mfo_grid_particle_index(int sx, int sy, int sz, long long np, const Real* pos, const int* pflag, int* index, int* indexSys, long long* count)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	int* counter = (int*)calloc((size_t)n, sizeof(int));
	memset(index, 0, sizeof(int) * (size_t)n);
	for (long long idx = 0; idx < np; idx++) {
		if (pflag[idx] & PDELETE) continue;
		const int x = (int)pos[3 * idx], y = (int)pos[3 * idx + 1], z = (int)pos[3 * idx + 2];
		if (!in_bounds0(sx, sy, sz, x, y, z)) continue;
		index[IDX(x, y, z)]++;
	}
	IndexInt run = 0;
	for (IndexInt q = 0; q < n; q++) { const int num = index[q]; index[q] = (int)run; run += num; }
	for (long long idx = 0; idx < np; idx++) {
		if (pflag[idx] & PDELETE) continue;
		const int x = (int)pos[3 * idx], y = (int)pos[3 * idx + 1], z = (int)pos[3 * idx + 2];
		if (!in_bounds0(sx, sy, sz, x, y, z)) continue;
		const IndexInt p = IDX(x, y, z);
		indexSys[index[p] + counter[p]] = (int)idx;
		counter[p]++;
	}
	free(counter);
	*count = run;
	return 0;
}

/* unionParticleLevelset plugin/flip.cpp:340-350 (ComputeUnionLevelsetPindex :308-338, calculateRadiusFactor :186-188), then phi.setBound(0.5, 0) */
int mfo_union_particle_levelset(int sx, int sy, int sz, long long np, const Real* pos, const int* index, const int* indexSys, long long count,
                                Real* phi, double radiusFactor_, const int* ptype, int exclude)
{
	STRIDES
	(void)np;
	const IndexInt n = (IndexInt)sx * sy * sz;
	const Real radiusFactor = (Real)radiusFactor_;
	const Real radius = (Real)(0.5 * (double)(Real)((IS3D ? sqrt(3.) : sqrt(2.)) * ((double)radiusFactor + .01)));
	const int r = (int)radius + 1, rZ = IS3D ? r : 0;
	for (int k = 0; k < sz; k++) for (int j = 0; j < sy; j++) for (int i = 0; i < sx; i++) {
		const Real gx = (Real)i + (Real)0.5, gy = (Real)j + (Real)0.5, gz = (Real)k + (Real)0.5;
		Real phiv = (Real)((double)radius * 1.0);
		for (int zj = k - rZ; zj <= k + rZ; zj++) for (int yj = j - r; yj <= j + r; yj++) for (int xj = i - r; xj <= i + r; xj++) {
			if (!in_bounds0(sx, sy, sz, xj, yj, zj)) continue;
			const IndexInt c = IDX(xj, yj, zj);
			const IndexInt pStart = index[c], pEnd = (c + 1 < n) ? index[c + 1] : count;
			for (IndexInt p = pStart; p < pEnd; p++) {
				const int psrc = indexSys[p];
				if (ptype && (ptype[psrc] & exclude)) continue;
				const Real dx = gx - pos[3 * psrc], dy = gy - pos[3 * psrc + 1], dz = gz - pos[3 * psrc + 2];
				const Real l = dx * dx + dy * dy + dz * dz;
				const Real nrm = l <= (Real)(MF_REAL_IS_DOUBLE ? 1e-10 * 1e-10 : 1e-6f * 1e-6f) ? (Real)0 : ((double)fabs((double)l - 1.) < (double)(Real)(MF_REAL_IS_DOUBLE ? 1e-10 * 1e-10 : 1e-6f * 1e-6f) ? (Real)1 : R_SQRT(l));
				const Real v = (Real)fabs((double)nrm) - radius;
				phiv = rmin(phiv, v);
			}
		}
		phi[IDX(i, j, k)] = phiv;
	}
	return mfo_set_bound(sx, sy, sz, phi, 1, 0.5, 0);
}

/* BUILD_INDEX / BUILD_INDEX_SHIFT util/interpol.h:50-66,:112-125 */
typedef struct { int xi, yi, zi, sxi, syi, szi; Real s0, s1, t0, t1, f0, f1, ss0, ss1, st0, st1, sf0, sf1; } MacIdx;
static inline MacIdx build_index_shift(int sx, int sy, int sz, const Real* pos)
{
	MacIdx m;
	const Real px = pos[0] - 0.5f, py = pos[1] - 0.5f, pz = pos[2] - 0.5f;
	m.xi = (int)px; m.yi = (int)py; m.zi = (int)pz;
	m.s1 = px - (Real)m.xi; m.s0 = (Real)(1. - m.s1); m.t1 = py - (Real)m.yi; m.t0 = (Real)(1. - m.t1); m.f1 = pz - (Real)m.zi; m.f0 = (Real)(1. - m.f1);
	if (px < 0.) { m.xi = 0; m.s0 = 1.0; m.s1 = 0.0; }
	if (py < 0.) { m.yi = 0; m.t0 = 1.0; m.t1 = 0.0; }
	if (pz < 0.) { m.zi = 0; m.f0 = 1.0; m.f1 = 0.0; }
	if (m.xi >= sx - 1) { m.xi = sx - 2; m.s0 = 0.0; m.s1 = 1.0; }
	if (m.yi >= sy - 1) { m.yi = sy - 2; m.t0 = 0.0; m.t1 = 1.0; }
	if (sz > 1) { if (m.zi >= sz - 1) { m.zi = sz - 2; m.f0 = 0.0; m.f1 = 1.0; } }
	m.sxi = (int)pos[0]; m.syi = (int)pos[1]; m.szi = (int)pos[2];
	m.ss1 = pos[0] - (Real)m.sxi; m.ss0 = (Real)(1. - m.ss1); m.st1 = pos[1] - (Real)m.syi; m.st0 = (Real)(1. - m.st1); m.sf1 = pos[2] - (Real)m.szi; m.sf0 = (Real)(1. - m.sf1);
	if (pos[0] < 0) { m.sxi = 0; m.ss0 = 1.0; m.ss1 = 0.0; }
	if (pos[1] < 0) { m.syi = 0; m.st0 = 1.0; m.st1 = 0.0; }
	if (pos[2] < 0) { m.szi = 0; m.sf0 = 1.0; m.sf1 = 0.0; }
	if (m.sxi >= sx - 1) { m.sxi = sx - 2; m.ss0 = 0.0; m.ss1 = 1.0; }
	if (m.syi >= sy - 1) { m.syi = sy - 2; m.st0 = 0.0; m.st1 = 1.0; }
	if (sz > 1) { if (m.szi >= sz - 1) { m.szi = sz - 2; m.sf0 = 0.0; m.sf1 = 1.0; } }
	return m;
}
/* interpolMAC util/interpol.h:127-157 */
static void interpol_mac(int sx, int sy, int sz, const Real* data, const Real* pos, Real out[3])
{
	const MacIdx m = build_index_shift(sx, sy, sz, pos);
	const IndexInt X = 1, Y = sx, Z = (sz > 1) ? (IndexInt)sx * sy : 0;
#define RF(o, c) ref[3 * (o) + (c)]
	{ const Real* ref = data + 3 * (((IndexInt)m.zi * sy + m.yi) * sx + m.sxi);
	  out[0] = m.f0 * ((RF(0, 0) * m.t0 + RF(Y, 0) * m.t1) * m.ss0 + (RF(X, 0) * m.t0 + RF(X + Y, 0) * m.t1) * m.ss1) +
	           m.f1 * ((RF(Z, 0) * m.t0 + RF(Z + Y, 0) * m.t1) * m.ss0 + (RF(X + Z, 0) * m.t0 + RF(X + Y + Z, 0) * m.t1) * m.ss1); }
	{ const Real* ref = data + 3 * (((IndexInt)m.zi * sy + m.syi) * sx + m.xi);
	  out[1] = m.f0 * ((RF(0, 1) * m.st0 + RF(Y, 1) * m.st1) * m.s0 + (RF(X, 1) * m.st0 + RF(X + Y, 1) * m.st1) * m.s1) +
	           m.f1 * ((RF(Z, 1) * m.st0 + RF(Z + Y, 1) * m.st1) * m.s0 + (RF(X + Z, 1) * m.st0 + RF(X + Y + Z, 1) * m.st1) * m.s1); }
	{ const Real* ref = data + 3 * (((IndexInt)m.szi * sy + m.yi) * sx + m.xi);
	  out[2] = m.sf0 * ((RF(0, 2) * m.t0 + RF(Y, 2) * m.t1) * m.s0 + (RF(X, 2) * m.t0 + RF(X + Y, 2) * m.t1) * m.s1) +
	           m.sf1 * ((RF(Z, 2) * m.t0 + RF(Z + Y, 2) * m.t1) * m.s0 + (RF(X + Z, 2) * m.t0 + RF(X + Y + Z, 2) * m.t1) * m.s1); }
#undef RF
}
/* one component of setInterpolMAC util/interpol.h:159-203: weights and the order of the sixteen updates as in the reference */
static void scatter8(Real* ref, Real* sum, int c, IndexInt X, IndexInt Y, IndexInt Z, Real a0, Real a1, Real b0, Real b1, Real c0, Real c1, Real v, int zFirst)
{	/* a: weights along x, b: along y, c: along z */
	const Real s0f0 = a0 * c0, s1f0 = a1 * c0, s0f1 = a0 * c1, s1f1 = a1 * c1;
	const Real w0 = b0 * s0f0, wx = b0 * s1f0, wy = b1 * s0f0, wxy = b1 * s1f0, wz = b0 * s0f1, wxz = b0 * s1f1, wyz = b1 * s0f1, wxyz = b1 * s1f1;
#define S_(o) sum[3 * (o) + c]
#define R_(o) ref[3 * (o) + c]
	if (zFirst) {
		S_(Z) += wz; S_(X + Z) += wxz; S_(Y + Z) += wyz; S_(X + Y + Z) += wxyz;
		R_(Z) += wz * v; R_(X + Z) += wxz * v; R_(Y + Z) += wyz * v; R_(X + Y + Z) += wxyz * v;
		S_(0) += w0; S_(X) += wx; S_(Y) += wy; S_(X + Y) += wxy;
		R_(0) += w0 * v; R_(X) += wx * v; R_(Y) += wy * v; R_(X + Y) += wxy * v;
	} else {
		S_(0) += w0; S_(X) += wx; S_(Y) += wy; S_(X + Y) += wxy;
		S_(Z) += wz; S_(X + Z) += wxz; S_(Y + Z) += wyz; S_(X + Y + Z) += wxyz;
		R_(0) += w0 * v; R_(X) += wx * v; R_(Y) += wy * v; R_(X + Y) += wxy * v;
		R_(Z) += wz * v; R_(X + Z) += wxz * v; R_(Y + Z) += wyz * v; R_(X + Y + Z) += wxyz * v;
	}
#undef S_
#undef R_
}
/* mapPartsToMAC plugin/flip.cpp:573-595 (knMapLinearVec3ToMACGrid :562-569 is a serial kernel: contributions are added in particle order),
 * weight->stomp(VECTOR_EPSILON) grid.cpp:224-226, vel.safeDivide(weight) general.h:148-151; weight (optional, may be NULL) receives the weights */
int mfo_map_parts_to_mac(int sx, int sy, int sz, Real* vel, Real* velOld, long long np, const Real* pos, const int* pflag, const Real* pvel,
                         Real* weight, const int* ptype, int exclude)
{
	const IndexInt n = (IndexInt)sx * sy * sz, X = 1, Y = sx, Z = (sz > 1) ? (IndexInt)sx * sy : 0;
	Real* w = weight ? weight : (Real*)malloc(sizeof(Real) * 3 * (size_t)n);
	memset(w, 0, sizeof(Real) * 3 * (size_t)n);
	memset(vel, 0, sizeof(Real) * 3 * (size_t)n);
	for (long long idx = 0; idx < np; idx++) {
		if (P_SKIP(idx)) continue;
		const MacIdx m = build_index_shift(sx, sy, sz, pos + 3 * idx);
		const Real* v = pvel + 3 * idx;
		{ const IndexInt q = ((IndexInt)m.zi * sy + m.yi) * sx + m.sxi;  scatter8(vel + 3 * q, w + 3 * q, 0, X, Y, Z, m.ss0, m.ss1, m.t0, m.t1, m.f0, m.f1, v[0], 1); }
		{ const IndexInt q = ((IndexInt)m.zi * sy + m.syi) * sx + m.xi;  scatter8(vel + 3 * q, w + 3 * q, 1, X, Y, Z, m.s0, m.s1, m.st0, m.st1, m.f0, m.f1, v[1], 1); }
		{ const IndexInt q = ((IndexInt)m.szi * sy + m.yi) * sx + m.xi;  scatter8(vel + 3 * q, w + 3 * q, 2, X, Y, Z, m.s0, m.s1, m.t0, m.t1, m.sf0, m.sf1, v[2], 0); }
	}
#if MF_REAL_IS_DOUBLE
	const Real eps = 1e-10;
#else
	const Real eps = 1e-6f;
#endif
	for (IndexInt q = 0; q < 3 * n; q++) { if (w[q] < eps) w[q] = 0; vel[q] = w[q] ? (vel[q] / w[q]) : vel[q]; }
	memcpy(velOld, vel, sizeof(Real) * 3 * (size_t)n);
	if (!weight) free(w);
	return 0;
}

/* mapMACToParts plugin/flip.cpp:651-656 (flipRatio < 0) and flipVelocityUpdate :669-677 */
int mfo_flip_velocity_update(int sx, int sy, int sz, const Real* vel, const Real* velOld, long long np, const Real* pos, const int* pflag, Real* pvel,
                             double flipRatio_, const int* ptype, int exclude)
{
	const Real flipRatio = (Real)flipRatio_;
	for (long long idx = 0; idx < np; idx++) {
		if (P_SKIP(idx)) continue;
		Real v[3];
		interpol_mac(sx, sy, sz, vel, pos + 3 * idx, v);
		if (flipRatio_ < 0) { for (int c = 0; c < 3; c++) pvel[3 * idx + c] = v[c]; continue; }
		Real o[3];
		interpol_mac(sx, sy, sz, velOld, pos + 3 * idx, o);
		for (int c = 0; c < 3; c++) {
			const Real delta = v[c] - o[c];
			pvel[3 * idx + c] = (Real)((double)(flipRatio * (pvel[3 * idx + c] + delta)) + (double)(Real)((1.0 - (double)flipRatio) * (double)v[c]));
		}
	}
	return 0;
}

/* ParticleSystem<S>::advectInGrid particle.h:512-536 (GridAdvectKernel :446-467, integratePointSet util/integrator.h:26-68, KnClampPositions :494-509
 * with bisectBacktracePos :480-490, KnDeleteInObstacle :471-477).  mode: 0 IntEuler, 1 IntRK2, 2 IntRK4.  pos and pflag are updated in place.
 * The velocity kernel leaves u untouched for a particle in an obstacle when stopInObstacle is off (the reference's result vector persists between runs). */
enum { PNEW_ = 1 << 0 };                                 /* particle.h:36 */
static inline int is_obstacle_at(int sx, int sy, int sz, const int* flags, const Real* p)   /* FlagGrid::isObstacle(const Vec3&) grid.h:313 */
{ return flags[(IndexInt)(int)p[0] + (IndexInt)sx * (int)p[1] + (sz > 1 ? (IndexInt)sx * sy * (int)p[2] : 0)] & TypeObstacle; }
static inline int in_bounds_b(int sx, int sy, int sz, const Real* p, int b)                  /* GridBase::isInBounds(Vec3, bnd) grid.h:60,:407-415 */
{ const int x = (int)p[0], y = (int)p[1], z = (int)p[2];
  return x >= b && y >= b && x < sx - b && y < sy - b && (sz > 1 ? (z >= b && z < sz - b) : z == 0); }
static void advect_vel_kernel(int sx, int sy, int sz, const int* flags, const Real* vel, const Real* pos, int* pflag, int ptype, int exclude, Real dt,
                              int deleteInObstacle, int stopInObstacle, int skipNew, Real u[3])
{
	if ((*pflag & PDELETE) || (ptype & exclude) || (skipNew && (*pflag & PNEW_))) { u[0] = u[1] = u[2] = 0; return; }
	if (deleteInObstacle || stopInObstacle) {
		if (!in_bounds_b(sx, sy, sz, pos, 1) || is_obstacle_at(sx, sy, sz, flags, pos)) {
			if (stopInObstacle) u[0] = u[1] = u[2] = 0;
			if (deleteInObstacle) *pflag |= PDELETE;
			return;
		}
	}
	Real v[3];
	interpol_mac(sx, sy, sz, vel, pos, v);
	for (int c = 0; c < 3; c++) u[c] = v[c] * dt;
}
int mfo_advect_in_grid(int sx, int sy, int sz, const int* flags, const Real* vel, long long np, Real* pos, int* pflag, double dt_, int mode,
                       int deleteInObstacle, int stopInObstacle, int skipNew, const int* ptype, int exclude)
{
	const Real dt = (Real)dt_;
	if (mode < 0 || mode > 2) { snprintf(g_err, sizeof g_err, "unknown integration type"); return 1; }
	for (long long idx = 0; idx < np; idx++) {
		Real* x = pos + 3 * idx;
		int* fl = pflag + idx;
		const int pt = ptype ? ptype[idx] : 0;
		const Real x0[3] = { x[0], x[1], x[2] };
		Real u[3] = { 0, 0, 0 }, uTotal[3];
#define RUN_ advect_vel_kernel(sx, sy, sz, flags, vel, x, fl, pt, exclude, dt, deleteInObstacle, stopInObstacle, skipNew, u)
		RUN_;
		if (mode == 0) { for (int c = 0; c < 3; c++) x[c] += u[c]; }
		else if (mode == 1) {
			for (int c = 0; c < 3; c++) x[c] = x0[c] + (Real)(0.5 * (double)u[c]);
			RUN_;
			for (int c = 0; c < 3; c++) x[c] = x0[c] + u[c];
		} else {
			for (int c = 0; c < 3; c++) { uTotal[c] = u[c]; x[c] = x0[c] + (Real)(0.5 * (double)u[c]); }
			RUN_;
			for (int c = 0; c < 3; c++) { x[c] = x0[c] + (Real)(0.5 * (double)u[c]); uTotal[c] += (Real)(2 * u[c]); }
			RUN_;
			for (int c = 0; c < 3; c++) { x[c] = x0[c] + u[c]; uTotal[c] += (Real)(2 * u[c]); }
			RUN_;
			for (int c = 0; c < 3; c++) x[c] = x0[c] + (Real)(1. / 6.) * (uTotal[c] + u[c]);
		}
#undef RUN_
		if (!deleteInObstacle) {                                      /* KnClampPositions */
			if (*fl & PDELETE) continue;
			if (pt & exclude) { for (int c = 0; c < 3; c++) x[c] = x0[c]; continue; }
			if (!in_bounds_b(sx, sy, sz, x, 0)) {
				const Real hi[3] = { (Real)sx - (Real)1, (Real)sy - (Real)1, (Real)sz - (Real)1 };
				for (int c = 0; c < 3; c++) { if (x[c] < (Real)0) x[c] = 0; else if (x[c] > hi[c]) x[c] = hi[c]; }
			}
			if (stopInObstacle && is_obstacle_at(sx, sy, sz, flags, x)) {      /* bisectBacktracePos */
				Real sacc = 0.;
				for (int i = 1; i < 5; ++i) {
					const Real ds = (Real)(1. / (double)(Real)(1 << i));
					const Real sd = sacc + ds;
					Real q[3];
					for (int c = 0; c < 3; c++) q[c] = (Real)((double)x0[c] * (1. - (double)sd)) + x[c] * sd;
					if (!is_obstacle_at(sx, sy, sz, flags, q)) sacc += ds;
				}
				for (int c = 0; c < 3; c++) x[c] = (Real)((double)x0[c] * (1. - (double)sacc)) + x[c] * sacc;
			}
		} else {                                                       /* KnDeleteInObstacle */
			if (*fl & PDELETE) continue;
			if (!in_bounds_b(sx, sy, sz, x, 1) || is_obstacle_at(sx, sy, sz, flags, x)) *fl |= PDELETE;
		}
	}
	return 0;
}

/* pushOutofObs plugin/flip.cpp:528-545 (knPushOutofObs; getGradient grid.h:520-537, interpol util/interpol.h:68-78, normalize vectorbase.h:415-429) */
int mfo_push_out_of_obs(int sx, int sy, int sz, long long np, Real* pos, const int* pflag, const Real* phiObs, double shift_, double thresh_, const int* ptype, int exclude)
{
	const Real shift = (Real)shift_, thresh = (Real)thresh_;
	const IndexInt Y = sx, Z = (sz > 1) ? (IndexInt)sx * sy : 0;
#if MF_REAL_IS_DOUBLE
	const Real eps = 1e-10;
#else
	const Real eps = 1e-6f;
#endif
	for (long long idx = 0; idx < np; idx++) {
		if (P_SKIP(idx)) continue;
		Real* x = pos + 3 * idx;
		int i = (int)x[0], j = (int)x[1], k = (int)x[2];
		if (!(i >= 0 && j >= 0 && k >= 0 && i < sx && j < sy && k < sz)) continue;          /* GridBase::isInBounds(Vec3i) grid.h:403-405 */
		const Real v = interpol_s(phiObs, 1, sx, sy, sz, x);
		if (v < thresh) {
			if (i > sx - 2) i = sx - 2;
			if (j > sy - 2) j = sy - 2;
			if (i < 1) i = 1;
			if (j < 1) j = 1;
			Real g[3];
			g[0] = phiObs[(i + 1) + Y * j + Z * k] - phiObs[(i - 1) + Y * j + Z * k];
			g[1] = phiObs[i + Y * (j + 1) + Z * k] - phiObs[i + Y * (j - 1) + Z * k];
			g[2] = 0.;
			if (sz > 1) {
				if (k > sz - 2) k = sz - 2;
				if (k < 1) k = 1;
				g[2] = phiObs[i + Y * j + Z * (k + 1)] - phiObs[i + Y * j + Z * (k - 1)];
			}
			if (normalize3(g) < eps) continue;
			const Real a = thresh - v + shift;
			for (int c = 0; c < 3; c++) x[c] = x[c] + g[c] * a;
		}
	}
	return 0;
}

/* ParticleSystem<S>::projectOutOfBnd particle.h:565-590; axis: bit q set <-> the q-th letter of "xXyYzZ" is in `plane` */
int mfo_project_out_of_bnd(int sx, int sy, int sz, long long np, Real* pos, const int* pflag, double bnd_, int axis, const int* ptype, int exclude)
{
	const Real bnd = (Real)bnd_;
	for (long long idx = 0; idx < np; idx++) {
		if (P_SKIP(idx)) continue;
		Real* x = pos + 3 * idx;
		if (axis & 1) x[0] = x[0] < bnd ? bnd : x[0];                                       /* std::max(pos.x, bnd) */
		if (axis & 2) { const Real hi = (Real)sx - bnd; x[0] = hi < x[0] ? hi : x[0]; }     /* std::min(pos.x, size - bnd) */
		if (axis & 4) x[1] = x[1] < bnd ? bnd : x[1];
		if (axis & 8) { const Real hi = (Real)sy - bnd; x[1] = hi < x[1] ? hi : x[1]; }
		if (sz > 1) {
			if (axis & 16) x[2] = x[2] < bnd ? bnd : x[2];
			if (axis & 32) { const Real hi = (Real)sz - bnd; x[2] = hi < x[2] ? hi : x[2]; }
		}
	}
	return 0;
}

/* The Lagrangian-particle helpers of scenes/benchmark_dam.py:118-134 (plugin/ptsplugins.cpp:17-70, ParticleSystem::getPosPdata particle.h:422-427,
 * markIsolatedFluidCell grid.cpp:866-890).  None of them looks at the PDELETE flag; an excluded type is skipped. */
#define T_SKIP(idx) (ptype && (ptype[idx] & exclude))
int mfo_add_force_pvel(long long np, Real* pvel, double ax, double ay, double az, double dt_, const int* ptype, int exclude)
{
	const Real dt = (Real)dt_, da[3] = { (Real)ax * dt, (Real)ay * dt, (Real)az * dt };
	for (long long idx = 0; idx < np; idx++) { if (T_SKIP(idx)) continue; for (int c = 0; c < 3; c++) pvel[3 * idx + c] += da[c]; }
	return 0;
}
int mfo_update_velocity_from_delta_pos(long long np, const Real* pos, Real* pvel, const Real* xPrev, double dt_, const int* ptype, int exclude)
{
	const Real overDt = (Real)(1.0 / (double)(Real)dt_);
	for (long long idx = 0; idx < np; idx++) { if (T_SKIP(idx)) continue; for (int c = 0; c < 3; c++) pvel[3 * idx + c] = (pos[3 * idx + c] - xPrev[3 * idx + c]) * overDt; }
	return 0;
}
int mfo_euler_step(long long np, Real* pos, const Real* pvel, double dt_, const int* ptype, int exclude)
{
	const Real dt = (Real)dt_;
	for (long long idx = 0; idx < np; idx++) { if (T_SKIP(idx)) continue; for (int c = 0; c < 3; c++) pos[3 * idx + c] += pvel[3 * idx + c] * dt; }
	return 0;
}
int mfo_set_part_type(int sx, int sy, int sz, long long np, const Real* pos, int* ptype, int mark, int stype, const int* flags, int cflag)
{
	for (long long idx = 0; idx < np; idx++) {
		const Real* x = pos + 3 * idx;
		if (in_bounds_b(sx, sy, sz, x, 0) && (flags[(IndexInt)(int)x[0] + (IndexInt)sx * (int)x[1] + (sz > 1 ? (IndexInt)sx * sy * (int)x[2] : 0)] & cflag) && (ptype[idx] & stype)) ptype[idx] = mark;
	}
	return 0;
}
int mfo_mark_isolated_fluid_cell(int sx, int sy, int sz, int* flags, int mark)
{
	STRIDES
	const IndexInt n = (IndexInt)sx * sy * sz;
	for (IndexInt q = 0; q < n; q++) {
		if (!(flags[q] & TypeFluid)) continue;
		if ((flags[q - X] & TypeFluid) || (flags[q + X] & TypeFluid) || (flags[q - Y] & TypeFluid) || (flags[q + Y] & TypeFluid)) continue;
		if (IS3D && ((flags[q - Z] & TypeFluid) || (flags[q + Z] & TypeFluid))) continue;
		flags[q] = mark;
	}
	return 0;
}

/* Grid<Vec3>::getMaxAbs grid.cpp:330-332 = sqrt(CompMaxVec :198-203) */
int mfo_vec_max_abs(int sx, int sy, int sz, const Real* v, double* out)
{
	const IndexInt n = (IndexInt)sx * sy * sz;
	Real m = -REAL_MAX_;
	for (IndexInt q = 0; q < n; q++) { const Real s = v[3 * q] * v[3 * q] + v[3 * q + 1] * v[3 * q + 1] + v[3 * q + 2] * v[3 * q + 2]; if (s > m) m = s; }
	*out = (double)R_SQRT(m);
	return 0;
}

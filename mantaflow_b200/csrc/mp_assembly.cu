// Assembly kernels of the pressure projection (bit-exact w.r.t. the reference; compiled -fmad=false):
//   k_make_rhs            MakeRhs                       plugin/pressure.cpp:32-84
//   k_make_matrix         MakeLaplaceMatrix             conjugategrad.h:154-187
//                         (+ ApplyGhostFluidDiagonal    plugin/pressure.cpp:136-151, fused)
//   k_scan_flags/k_choose CountEmptyCells + cell choice plugin/pressure.cpp:217-220,:352-382
//   k_fix_pressure        fixPressure                   plugin/pressure.cpp:226-245
//   k_correct_velocity    knCorrectVelocity + knCorrectVelocityGhostFluid (fused)  :87-109,:154-187
//   k_replace_clamped     knReplaceClampedGhostFluidVels :198-214
// All are single-pass streaming kernels bounded by HBM bandwidth; they run once per solve.
#include "mp_common.cuh"

// ---------------------------------------------------------------- ghost-fluid helpers (pressure.cpp:115-133,:191-196)
template <typename Real> __device__ __forceinline__ Real thetaHelper(Real inside, Real outside) {
	const Real denom = inside - outside;
	if ((double)denom > -1e-04) return (Real)0.5;
	const Real q = inside / denom;
	const Real m = (q < (Real)1) ? q : (Real)1;        // std::min(Real(1), q)
	return ((Real)0 < m) ? m : (Real)0;                // std::max(Real(0), m)
}
template <typename Real> __device__ __forceinline__ Real ghostFluidHelper(IndexInt idx, IndexInt offset, const Real* __restrict__ phi, Real gfClamp) {
	const Real alpha = thetaHelper<Real>(phi[idx], phi[idx + offset]);
	if (alpha < gfClamp) return gfClamp;
	return (Real)(1. - (1. / (double)alpha));          // double arithmetic, then narrowed
}
template <typename Real> __device__ __forceinline__ Real surfTensHelper(IndexInt idx, IndexInt offset, const Real* __restrict__ phi, const Real* __restrict__ curv, Real surfTens, Real gfClamp) {
	return surfTens * (curv[idx + offset] - ghostFluidHelper<Real>(idx, offset, phi, gfClamp) * curv[idx]);
}
template <typename Real> __device__ __forceinline__ bool ghostFluidWasClamped(IndexInt idx, IndexInt offset, const Real* __restrict__ phi, Real gfClamp) {
	return thetaHelper<Real>(phi[idx], phi[idx + offset]) < gfClamp;
}

// interior test of KERNEL(bnd=1): kernel.cpp:21-30
__device__ __forceinline__ bool interior(const Dims& d, IndexInt idx, int& i, int& j, int& k) {
	cellOf(d, idx, i, j, k);
	if (i < 1 || i >= d.sx - 1 || j < 1 || j >= d.sy - 1) return false;
	if (d.is3D) { const int kg = k + d.kOff; return k >= d.kb && k < d.ke && kg >= 1 && kg < d.gsz - 1; }   // owned plane, global interior
	return true;
}

// ---------------------------------------------------------------- MakeRhs
// FULL = false: no face fractions, no surface tension -- the smoke / plain liquid case, compiled without those paths (the full kernel needs 113
// registers and runs at a quarter of the occupancy; same cells per thread and same reduction order in both)
template <typename Real, bool FULL>
__global__ void __launch_bounds__(256, FULL ? 1 : 3) k_make_rhs(Dims d, const int* __restrict__ flags, Real* __restrict__ rhs, const Real* __restrict__ vel,
	const Real* __restrict__ perCellCorr, const Real* __restrict__ fractions, const Real* __restrict__ obvel,
	const Real* __restrict__ phi, const Real* __restrict__ curv, Real surfTens, Real gfClamp,
	double* partials, unsigned int* ticket, double* out)
{
	double v[2] = { 0.0, 0.0 };
	if (!FULL) {
		// The plain case four cells at a time: a thread's cells are the same and are summed in the same order as in the loop below, but the
		// flag and velocity loads of four cells are in flight together instead of two dependent round trips per cell (1.35 -> 0.5 ms at 512^3).
		constexpr int U = 4;
		const IndexInt stride = (IndexInt)gridDim.x * blockDim.x;
		const IndexInt X = d.X, Y = d.Y, Z = d.Z;
		for (IndexInt base = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; base < d.n; base += U * stride) {
			bool in[U]; int fl[U]; Real c0[U], c1[U], c2[U], x0[U], y1[U], z2[U], pc[U];
			#pragma unroll
			for (int u = 0; u < U; u++) {
				const IndexInt idx = base + u * stride;
				int i, j, k;
				in[u] = idx < d.n && interior(d, idx, i, j, k);
				fl[u] = 0; c0[u] = c1[u] = c2[u] = x0[u] = y1[u] = z2[u] = pc[u] = (Real)0;
				if (in[u]) {          // interior cells have all their plus neighbours: the velocity is requested before the flag is known
					fl[u] = flags[idx];
					const Real* c = vel + 3 * idx;
					c0[u] = c[0]; c1[u] = c[1]; x0[u] = c[3 * X]; y1[u] = c[3 * Y + 1];
					if (d.is3D) { c2[u] = c[2]; z2[u] = c[3 * Z + 2]; }
					if (perCellCorr) pc[u] = perCellCorr[idx];
				}
			}
			#pragma unroll
			for (int u = 0; u < U; u++) {
				if (!in[u]) continue;
				const IndexInt idx = base + u * stride;
				if (!(fl[u] & TypeFluid)) { rhs[idx] = 0; continue; }
				Real set = c0[u] - x0[u] + c1[u] - y1[u];
				if (d.is3D) set += c2[u] - z2[u];
				if (perCellCorr) set += pc[u];
				v[0] += (double)set; v[1] += 1.0;
				rhs[idx] = set;
			}
		}
	} else
	for (IndexInt idx = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x; idx < d.n; idx += (IndexInt)gridDim.x * blockDim.x) {
		int i, j, k;
		if (!interior(d, idx, i, j, k)) continue;
		if (!(flags[idx] & TypeFluid)) rhs[idx] = 0;
		else {
			const IndexInt X = d.X, Y = d.Y, Z = d.Z;
			const Real* c = vel + 3 * idx; const Real* cx = vel + 3 * (idx + X); const Real* cy = vel + 3 * (idx + Y); const Real* cz = vel + 3 * (idx + Z);
			Real set;
			if (!FULL || !fractions) {
				set = c[0] - cx[0] + c[1] - cy[1];
				if (d.is3D) set += c[2] - cz[2];
			} else {
				const Real* f = fractions + 3 * idx; const Real* fx = fractions + 3 * (idx + X); const Real* fy = fractions + 3 * (idx + Y); const Real* fz = fractions + 3 * (idx + Z);
				set = f[0] * c[0] - fx[0] * cx[0] + f[1] * c[1] - fy[1] * cy[1];
				if (d.is3D) set += f[2] * c[2] - fz[2] * cz[2];
				if (obvel) {
					const Real* o = obvel + 3 * idx; const Real* ox = obvel + 3 * (idx + X); const Real* oy = obvel + 3 * (idx + Y); const Real* oz = obvel + 3 * (idx + Z);
					set += (1 - f[0]) * o[0] - (1 - fx[0]) * ox[0] + (1 - f[1]) * o[1] - (1 - fy[1]) * oy[1];
					if (d.is3D) set += (1 - f[2]) * o[2] - (1 - fz[2]) * oz[2];
				}
			}
			if (FULL && phi && curv) {
				if (flags[idx - X] & TypeEmpty) set += surfTensHelper<Real>(idx, -X, phi, curv, surfTens, gfClamp);
				if (flags[idx + X] & TypeEmpty) set += surfTensHelper<Real>(idx, +X, phi, curv, surfTens, gfClamp);
				if (flags[idx - Y] & TypeEmpty) set += surfTensHelper<Real>(idx, -Y, phi, curv, surfTens, gfClamp);
				if (flags[idx + Y] & TypeEmpty) set += surfTensHelper<Real>(idx, +Y, phi, curv, surfTens, gfClamp);
				if (d.is3D) {
					if (flags[idx - Z] & TypeEmpty) set += surfTensHelper<Real>(idx, -Z, phi, curv, surfTens, gfClamp);
					if (flags[idx + Z] & TypeEmpty) set += surfTensHelper<Real>(idx, +Z, phi, curv, surfTens, gfClamp);
				}
			}
			if (perCellCorr) set += perCellCorr[idx];
			v[0] += (double)set; v[1] += 1.0;
			rhs[idx] = set;
		}
	}
	const bool isMax[2] = { false, false };
	double fin[2];
	if (blockReduceFinal<2>(v, isMax, partials, ticket, fin) && threadIdx.x == 0) { out[0] = fin[0]; out[1] = fin[1]; }
}

// ---------------------------------------------------------------- MakeLaplaceMatrix (+ ghost fluid diagonal)
template <typename Real>
__global__ void __launch_bounds__(256) k_make_matrix(Dims d, const int* __restrict__ flags, const Real* __restrict__ fractions,
	const Real* __restrict__ phi, Real gfClamp, Real* __restrict__ A0, Real* __restrict__ Ai, Real* __restrict__ Aj, Real* __restrict__ Ak)
{
	const IndexInt idx = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= d.n) return;
	Real a0 = 0, ai = 0, aj = 0, ak = 0;
	int i, j, k;
	if (interior(d, idx, i, j, k) && (flags[idx] & TypeFluid)) {
		const IndexInt X = d.X, Y = d.Y, Z = d.Z;
		const int fxm = flags[idx - X], fxp = flags[idx + X], fym = flags[idx - Y], fyp = flags[idx + Y];
		const int fzm = d.is3D ? flags[idx - Z] : 0, fzp = d.is3D ? flags[idx + Z] : 0;
		if (!fractions) {
			if (!(fxm & TypeObstacle)) a0 += (Real)1;
			if (!(fxp & TypeObstacle)) a0 += (Real)1;
			if (!(fym & TypeObstacle)) a0 += (Real)1;
			if (!(fyp & TypeObstacle)) a0 += (Real)1;
			if (d.is3D && !(fzm & TypeObstacle)) a0 += (Real)1;
			if (d.is3D && !(fzp & TypeObstacle)) a0 += (Real)1;
			if (fxp & TypeFluid) ai = (Real)-1;
			if (fyp & TypeFluid) aj = (Real)-1;
			if (d.is3D && (fzp & TypeFluid)) ak = (Real)-1;
		} else {
			a0 += fractions[3 * idx + 0];
			a0 += fractions[3 * (idx + X) + 0];
			a0 += fractions[3 * idx + 1];
			a0 += fractions[3 * (idx + Y) + 1];
			if (d.is3D) a0 += fractions[3 * idx + 2];
			if (d.is3D) a0 += fractions[3 * (idx + Z) + 2];
			if (fxp & TypeFluid) ai = -fractions[3 * (idx + X) + 0];
			if (fyp & TypeFluid) aj = -fractions[3 * (idx + Y) + 1];
			if (d.is3D && (fzp & TypeFluid)) ak = -fractions[3 * (idx + Z) + 2];
		}
		if (phi) {
			if (fxm & TypeEmpty) a0 -= ghostFluidHelper<Real>(idx, -X, phi, gfClamp);
			if (fxp & TypeEmpty) a0 -= ghostFluidHelper<Real>(idx, +X, phi, gfClamp);
			if (fym & TypeEmpty) a0 -= ghostFluidHelper<Real>(idx, -Y, phi, gfClamp);
			if (fyp & TypeEmpty) a0 -= ghostFluidHelper<Real>(idx, +Y, phi, gfClamp);
			if (d.is3D) {
				if (fzm & TypeEmpty) a0 -= ghostFluidHelper<Real>(idx, -Z, phi, gfClamp);
				if (fzp & TypeEmpty) a0 -= ghostFluidHelper<Real>(idx, +Z, phi, gfClamp);
			}
		}
	}
	A0[idx] = a0; Ai[idx] = ai; Aj[idx] = aj; Ak[idx] = ak;
}

// ghost-fluid diagonal as a separate in-place pass (API parity with ApplyGhostFluidDiagonal)
template <typename Real>
__global__ void __launch_bounds__(256) k_ghost_diag(Dims d, const int* __restrict__ flags, const Real* __restrict__ phi, Real gfClamp, Real* __restrict__ A0)
{
	const IndexInt idx = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	int i, j, k;
	if (idx >= d.n || !interior(d, idx, i, j, k) || !(flags[idx] & TypeFluid)) return;
	const IndexInt X = d.X, Y = d.Y, Z = d.Z;
	Real a0 = A0[idx];
	if (flags[idx - X] & TypeEmpty) a0 -= ghostFluidHelper<Real>(idx, -X, phi, gfClamp);
	if (flags[idx + X] & TypeEmpty) a0 -= ghostFluidHelper<Real>(idx, +X, phi, gfClamp);
	if (flags[idx - Y] & TypeEmpty) a0 -= ghostFluidHelper<Real>(idx, -Y, phi, gfClamp);
	if (flags[idx + Y] & TypeEmpty) a0 -= ghostFluidHelper<Real>(idx, +Y, phi, gfClamp);
	if (d.is3D) {
		if (flags[idx - Z] & TypeEmpty) a0 -= ghostFluidHelper<Real>(idx, -Z, phi, gfClamp);
		if (flags[idx + Z] & TypeEmpty) a0 -= ghostFluidHelper<Real>(idx, +Z, phi, gfClamp);
	}
	A0[idx] = a0;
}

// ---------------------------------------------------------------- flag scans
// slots: [0] #empty cells, [1] min linear index of an interior fluid cell, [2] #fluid cells on the outer layer
__global__ void __launch_bounds__(256) k_scan_flags(Dims d, const int* __restrict__ flags, unsigned long long* slots)
{
	// a warp walks whole rows of the owned planes [kb, ke): (j, k) come from one division per row, not per cell
	unsigned long long nEmpty = 0, nBad = 0, minFluid = ~0ull;
	const int lane = threadIdx.x & 31, warpsPerBlock = blockDim.x >> 5;
	const long long nrows = (long long)d.sy * (d.ke - d.kb), plane = (long long)d.sx * d.sy;
	for (long long r = (long long)blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); r < nrows; r += (long long)gridDim.x * warpsPerBlock) {
		const int k = d.kb + (int)(r / d.sy), j = (int)(r - (long long)(k - d.kb) * d.sy);
		const int kg = k + d.kOff;
		const bool rowInterior = j >= 1 && j < d.sy - 1 && (!d.is3D || (kg >= 1 && kg < d.gsz - 1));
		const IndexInt row = (IndexInt)k * plane + (IndexInt)j * d.sx;
		for (int i = lane; i < d.sx; i += 32) {
			const int f = flags[row + i];
			if (f & TypeEmpty) nEmpty++;
			if (f & TypeFluid) {
				if (rowInterior && i >= 1 && i < d.sx - 1) {
					const unsigned long long gidx = (unsigned long long)(row + i + (IndexInt)d.kOff * plane);      // index in the global grid
					if (gidx < minFluid) minFluid = gidx;
				} else nBad++;
			}
		}
	}
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		nEmpty += __shfl_xor_sync(0xffffffffu, nEmpty, o);
		nBad += __shfl_xor_sync(0xffffffffu, nBad, o);
		const unsigned long long m = __shfl_xor_sync(0xffffffffu, minFluid, o);
		minFluid = m < minFluid ? m : minFluid;
	}
	if ((threadIdx.x & 31) == 0) {
		if (nEmpty) atomicAdd(&slots[0], nEmpty);
		if (minFluid != ~0ull) atomicMin(&slots[1], minFluid);
		if (nBad) atomicAdd(&slots[2], nBad);
	}
}
__global__ void k_scan_init(unsigned long long* slots) { slots[0] = 0; slots[1] = ~0ull; slots[2] = 0; slots[3] = 0; }

// the three preferred cells of pressure.cpp:357-367 (top centre, one below, two below): bit q of slots[3] = cell q is fluid
__global__ void k_scan_preferred(Dims d, const int* __restrict__ flags, unsigned long long* slots)
{
	const int cx = d.sx / 2, czg = d.is3D ? d.gsz / 2 : 0, cz = czg - d.kOff;
	if (d.is3D && (cz < d.kb || cz >= d.ke)) return;          // another rank owns that plane
	unsigned long long bits = 0;
	for (int q = 0; q < 3; q++) {
		const int cy = d.sy - 1 - q;
		if (cy < 0) continue;
		if (flags[(IndexInt)cx + (IndexInt)d.sx * cy + d.Z * cz] & TypeFluid) bits |= 1ull << q;
	}
	slots[3] = bits;
}

// combine the ranks' scan results: gathered[r*8 + q] holds rank r's slots as doubles (exact below 2^53)
__global__ void k_scan_to_double(const unsigned long long* slots, double* out) {
	out[0] = (double)slots[0]; out[1] = slots[1] == ~0ull ? -1.0 : (double)slots[1]; out[2] = (double)slots[2]; out[3] = (double)slots[3];
}
__global__ void k_scan_combine(const double* gathered, int world, unsigned long long* slots) {
	unsigned long long nEmpty = 0, nBad = 0, bits = 0, minFluid = ~0ull;
	for (int r = 0; r < world; r++) {
		const double* g = gathered + 8 * r;
		nEmpty += (unsigned long long)g[0]; nBad += (unsigned long long)g[2]; bits |= (unsigned long long)g[3];
		if (g[1] >= 0 && (unsigned long long)g[1] < minFluid) minFluid = (unsigned long long)g[1];
	}
	slots[0] = nEmpty; slots[1] = minFluid; slots[2] = nBad; slots[3] = bits;
}

// pressure.cpp:352-382: -1 if any empty cell; else top centre, one below, two below; else first interior fluid cell.
// The result is an index into the GLOBAL grid.
__global__ void k_choose_fix(Dims d, const unsigned long long* slots, long long* outIdx)
{
	long long fix = -1;
	if (slots[0] == 0) {
		const int cx = d.sx / 2, cz = d.is3D ? d.gsz / 2 : 0;
		for (int q = 0; q < 3 && fix < 0; q++) {
			const int cy = d.sy - 1 - q;
			if (cy < 0) continue;
			if (slots[3] & (1ull << q)) fix = (long long)cx + (long long)d.sx * cy + d.Z * cz;
		}
		if (fix < 0 && slots[1] != ~0ull) fix = (long long)slots[1];
	}
	*outIdx = fix;
}

// fixPressure pressure.cpp:226-245; idx taken from *pIdx (device) so that no host round trip is needed
template <typename Real>
__global__ void k_fix_pressure(Dims d, const long long* pIdx, Real value, Real* rhs, Real* A0, Real* Ai, Real* Aj, Real* Ak)
{
	const long long pg = *pIdx;                               // index in the global grid
	if (pg < 0) return;
	const IndexInt X = d.X, Y = d.Y, Z = d.Z;
	// local index; in slab mode every rank applies the edits that land on its planes (ghost copies included), so
	// owned data and ghost copies stay consistent without another exchange
	const long long p = pg - (long long)d.kOff * Z;
	const int kp = d.is3D ? (int)(p >= 0 ? p / Z : -1 - (-p - 1) / Z) : 0;    // local plane of the pinned cell (floor division)
	const bool here = kp >= 0 && kp < d.sz, below = kp - 1 >= 0 && kp - 1 < d.sz, above = kp + 1 >= 0 && kp + 1 < d.sz;
	if (here) {
		rhs[p + X] -= Ai[p] * value;
		rhs[p + Y] -= Aj[p] * value;
		rhs[p - X] -= Ai[p - X] * value;
		rhs[p - Y] -= Aj[p - Y] * value;
	}
	if (d.is3D) {
		if (here && above) rhs[p + Z] -= Ak[p] * value;
		if (below) rhs[p - Z] -= Ak[p - Z] * value;
	}
	if (here) {
		rhs[p] = value;
		A0[p] = (Real)1;
		Ai[p] = Aj[p] = Ak[p] = (Real)0;
		Ai[p - X] = (Real)0;
		Aj[p - Y] = (Real)0;
	}
	if (d.is3D && below) Ak[p - Z] = (Real)0;
}

// ---------------------------------------------------------------- correctVelocity
// GHOST = false: no level set -- compiled without the ghost-fluid / surface-tension paths (fewer registers, full occupancy)
template <typename Real, bool GHOST>
__global__ void __launch_bounds__(256, GHOST ? 4 : 8) k_correct_velocity(Dims d, const int* __restrict__ flags, Real* __restrict__ vel, const Real* __restrict__ pressure,
	const Real* __restrict__ phi, const Real* __restrict__ curv, Real gfClamp, Real surfTens)
{
	const IndexInt idx = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	int i, j, k;
	if (idx >= d.n || !interior(d, idx, i, j, k)) return;
	const IndexInt X = d.X, Y = d.Y, Z = d.Z;
	// an interior cell has its minus neighbours: everything is requested before the cell's own flag is known (one round trip instead of two)
	const int f = flags[idx];
	const int fx = flags[idx - X], fy = flags[idx - Y], fz = d.is3D ? flags[idx - Z] : 0;
	Real vx = vel[3 * idx + 0], vy = vel[3 * idx + 1], vz = vel[3 * idx + 2];
	const Real p = pressure[idx];
	const Real pxm = pressure[idx - X], pym = pressure[idx - Y], pzm = d.is3D ? pressure[idx - Z] : (Real)0;
	const bool fl = f & TypeFluid, em = (f & TypeEmpty) && !(f & TypeOutflow);
	if (!fl && !em) return;
	// knCorrectVelocity :87-109
	if (fl) {
		if (fx & TypeFluid) vx -= (p - pxm);
		if (fy & TypeFluid) vy -= (p - pym);
		if (d.is3D && (fz & TypeFluid)) vz -= (p - pzm);
		if (fx & TypeEmpty) vx -= p;
		if (fy & TypeEmpty) vy -= p;
		if (d.is3D && (fz & TypeEmpty)) vz -= p;
	} else {
		if (fx & TypeFluid) vx += pxm; else vx = 0.f;
		if (fy & TypeFluid) vy += pym; else vy = 0.f;
		if (d.is3D) { if (fz & TypeFluid) vz += pzm; else vz = 0.f; }
	}
	// knCorrectVelocityGhostFluid :154-187 (touches only this cell's velocity -> fused)
	if (GHOST && phi) {
		if (fl) {
			if (fx & TypeEmpty) vx += p * ghostFluidHelper<Real>(idx, -X, phi, gfClamp);
			if (fy & TypeEmpty) vy += p * ghostFluidHelper<Real>(idx, -Y, phi, gfClamp);
			if (d.is3D && (fz & TypeEmpty)) vz += p * ghostFluidHelper<Real>(idx, -Z, phi, gfClamp);
		} else {
			if (fx & TypeFluid) vx -= pressure[idx - X] * ghostFluidHelper<Real>(idx - X, +X, phi, gfClamp); else vx = 0.f;
			if (fy & TypeFluid) vy -= pressure[idx - Y] * ghostFluidHelper<Real>(idx - Y, +Y, phi, gfClamp); else vy = 0.f;
			if (d.is3D) { if (fz & TypeFluid) vz -= pressure[idx - Z] * ghostFluidHelper<Real>(idx - Z, +Z, phi, gfClamp); else vz = 0.f; }
		}
		if (curv) {
			if (fl) {
				if (fx & TypeEmpty) vx += surfTensHelper<Real>(idx, -X, phi, curv, surfTens, gfClamp);
				if (fy & TypeEmpty) vy += surfTensHelper<Real>(idx, -Y, phi, curv, surfTens, gfClamp);
				if (d.is3D && (fz & TypeEmpty)) vz += surfTensHelper<Real>(idx, -Z, phi, curv, surfTens, gfClamp);
			} else {
				vx -= (fx & TypeFluid) ? surfTensHelper<Real>(idx - X, +X, phi, curv, surfTens, gfClamp) : (Real)0.f;
				vy -= (fy & TypeFluid) ? surfTensHelper<Real>(idx - Y, +Y, phi, curv, surfTens, gfClamp) : (Real)0.f;
				if (d.is3D) vz -= (fz & TypeFluid) ? surfTensHelper<Real>(idx - Z, +Z, phi, curv, surfTens, gfClamp) : (Real)0.f;
			}
		}
	}
	vel[3 * idx + 0] = vx; vel[3 * idx + 1] = vy; vel[3 * idx + 2] = vz;
}

// knReplaceClampedGhostFluidVels :198-214.  Empty cells copy components from FLUID neighbours, which this
// kernel never writes, so a separate pass after k_correct_velocity reproduces the serial result exactly.
template <typename Real>
__global__ void __launch_bounds__(256) k_replace_clamped(Dims d, const int* __restrict__ flags, Real* __restrict__ vel, const Real* __restrict__ phi, Real gfClamp)
{
	const IndexInt idx = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	int i, j, k;
	if (idx >= d.n || !interior(d, idx, i, j, k)) return;
	if (!(flags[idx] & TypeEmpty)) return;
	const IndexInt X = d.X, Y = d.Y, Z = d.Z;
	if ((flags[idx - X] & TypeFluid) && ghostFluidWasClamped<Real>(idx - X, +X, phi, gfClamp)) vel[3 * idx + 0] = vel[3 * (idx - X) + 0];
	if ((flags[idx - Y] & TypeFluid) && ghostFluidWasClamped<Real>(idx - Y, +Y, phi, gfClamp)) vel[3 * idx + 1] = vel[3 * (idx - Y) + 1];
	if (d.is3D && (flags[idx - Z] & TypeFluid) && ghostFluidWasClamped<Real>(idx - Z, +Z, phi, gfClamp)) vel[3 * idx + 2] = vel[3 * (idx - Z) + 2];
	if ((flags[idx + X] & TypeFluid) && ghostFluidWasClamped<Real>(idx + X, -X, phi, gfClamp)) vel[3 * idx + 0] = vel[3 * (idx + X) + 0];
	if ((flags[idx + Y] & TypeFluid) && ghostFluidWasClamped<Real>(idx + Y, -Y, phi, gfClamp)) vel[3 * idx + 1] = vel[3 * (idx + Y) + 1];
	if (d.is3D && (flags[idx + Z] & TypeFluid) && ghostFluidWasClamped<Real>(idx + Z, -Z, phi, gfClamp)) vel[3 * idx + 2] = vel[3 * (idx + Z) + 2];
}

// ================================================================ host side
static int scanFlags(mp_context* ctx, const mp_grid* flags) {
	unsigned long long* slots = (unsigned long long*)(ctx->dScal + 8);
	const Dims d = dimsOf(flags);
	MP_TRY(mp_dist_check_grid(flags));
	k_scan_init<<<1, 1, 0, ctx->stream>>>(slots); MP_CHECK_LAUNCH(ctx);
	unsigned int blocks = gridFor(d.i1 - d.i0, 256 * 8); if (blocks > (unsigned)ctx->smCount * 16) blocks = ctx->smCount * 16;
	k_scan_flags<<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, slots); MP_CHECK_LAUNCH(ctx);
	k_scan_preferred<<<1, 1, 0, ctx->stream>>>(d, (const int*)flags->d, slots); MP_CHECK_LAUNCH(ctx);
	if (d.world > 1) {
		k_scan_to_double<<<1, 1, 0, ctx->stream>>>(slots, ctx->dist->dLocal); MP_CHECK_LAUNCH(ctx);
		MP_TRY(mp_dist_allgather(ctx, 4));
		k_scan_combine<<<1, 1, 0, ctx->stream>>>(ctx->dist->dGather, d.world, slots); MP_CHECK_LAUNCH(ctx);
	}
	return MP_OK;
}

int mp_check_flags_interior(mp_context* ctx, const mp_grid* flags) {
	MP_TRY(scanFlags(ctx, flags));
	MP_CUDA(cudaMemcpyAsync(ctx->hScal + 8, ctx->dScal + 8, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	const unsigned long long nBad = ((unsigned long long*)(ctx->hScal + 8))[2];
	if (nBad) MP_FAIL(MP_ERR_INVALID, "FlagGrid has %llu fluid cell(s) on the outer layer of the domain; the reference reads out of bounds there (conjugategrad.h:126-132)", nBad);
	return MP_OK;
}

// device-side cell choice + pin, no host round trip; result index left in ctx->dScal[12] (as long long)
int mp_fix_pressure_auto(mp_context* ctx, const mp_grid* flags, mp_grid* rhs, mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak) {
	const Dims d = dimsOf(flags);
	MP_TRY(scanFlags(ctx, flags));
	long long* pIdx = (long long*)(ctx->dScal + 12);
	k_choose_fix<<<1, 1, 0, ctx->stream>>>(d, (const unsigned long long*)(ctx->dScal + 8), pIdx); MP_CHECK_LAUNCH(ctx);
	if (rhs->prec == 4) k_fix_pressure<float><<<1, 1, 0, ctx->stream>>>(d, pIdx, 0.f, (float*)rhs->d, (float*)A0->d, (float*)Ai->d, (float*)Aj->d, (float*)Ak->d);
	else                k_fix_pressure<double><<<1, 1, 0, ctx->stream>>>(d, pIdx, 0., (double*)rhs->d, (double*)A0->d, (double*)Ai->d, (double*)Aj->d, (double*)Ak->d);
	MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

extern "C" {

int mp_make_rhs(mp_context* ctx, const mp_grid* flags, mp_grid* rhs, const mp_grid* vel,
                const mp_grid* perCellCorr, const mp_grid* fractions, const mp_grid* obvel,
                const mp_grid* phi, const mp_grid* curv, double surfTens, double gfClamp, double* sum, int* cnt)
{
	if (!ctx || !flags || !rhs || !vel) MP_FAIL(MP_ERR_INVALID, "mp_make_rhs: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_make_rhs: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, rhs, MP_GRID_REAL, "rhs", false));
	MP_TRY(mp_check_same(rhs, vel, MP_GRID_MAC, "vel", false));
	MP_TRY(mp_check_same(rhs, perCellCorr, MP_GRID_REAL, "perCellCorr", true));
	MP_TRY(mp_check_same(rhs, fractions, MP_GRID_MAC, "fractions", true));
	MP_TRY(mp_check_same(rhs, obvel, MP_GRID_MAC, "obvel", true));
	MP_TRY(mp_check_same(rhs, phi, MP_GRID_REAL, "phi", true));
	MP_TRY(mp_check_same(rhs, curv, MP_GRID_REAL, "curv", true));
	MP_CUDA(cudaSetDevice(ctx->device));
	const Dims d = dimsOf(flags);
	unsigned int blocks = gridFor(d.n, 256);
	if (blocks > (unsigned)kMaxPartials) blocks = kMaxPartials;   // grid-stride beyond that
	const bool full = fractions || (phi && curv);
	if (rhs->prec == 4) {
		if (full) k_make_rhs<float, true><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (float*)rhs->d, (const float*)vel->d, dptr<float>(perCellCorr), dptr<float>(fractions),
			dptr<float>(obvel), dptr<float>(phi), dptr<float>(curv), (float)surfTens, (float)gfClamp, ctx->partials, ctx->tickets + 1, ctx->dScal + 2);
		else k_make_rhs<float, false><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (float*)rhs->d, (const float*)vel->d, dptr<float>(perCellCorr), nullptr,
			nullptr, nullptr, nullptr, 0.f, (float)gfClamp, ctx->partials, ctx->tickets + 1, ctx->dScal + 2);
	} else {
		if (full) k_make_rhs<double, true><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (double*)rhs->d, (const double*)vel->d, dptr<double>(perCellCorr), dptr<double>(fractions),
			dptr<double>(obvel), dptr<double>(phi), dptr<double>(curv), surfTens, gfClamp, ctx->partials, ctx->tickets + 1, ctx->dScal + 2);
		else k_make_rhs<double, false><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (double*)rhs->d, (const double*)vel->d, dptr<double>(perCellCorr), nullptr,
			nullptr, nullptr, nullptr, 0., gfClamp, ctx->partials, ctx->tickets + 1, ctx->dScal + 2);
	}
	MP_CHECK_LAUNCH(ctx);
	MP_TRY(mp_dist_sum(ctx, ctx->dScal + 2, 2));          // slab mode: global sum / cnt
	if (sum || cnt) {
		MP_CUDA(cudaMemcpyAsync(ctx->hScal + 2, ctx->dScal + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
		MP_CUDA(cudaStreamSynchronize(ctx->stream));
		if (sum) *sum = ctx->hScal[2];
		if (cnt) *cnt = (int)ctx->hScal[3];
	}
	return MP_OK;
}

static int makeMatrix(mp_context* ctx, const mp_grid* flags, mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak, const mp_grid* fractions, const mp_grid* phi, double gfClamp)
{
	if (!ctx || !flags || !A0 || !Ai || !Aj || !Ak) MP_FAIL(MP_ERR_INVALID, "mp_make_laplace_matrix: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_make_laplace_matrix: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, A0, MP_GRID_REAL, "A0", false));
	MP_TRY(mp_check_same(A0, Ai, MP_GRID_REAL, "Ai", false)); MP_TRY(mp_check_same(A0, Aj, MP_GRID_REAL, "Aj", false)); MP_TRY(mp_check_same(A0, Ak, MP_GRID_REAL, "Ak", false));
	MP_TRY(mp_check_same(A0, fractions, MP_GRID_MAC, "fractions", true));
	MP_TRY(mp_check_same(A0, phi, MP_GRID_REAL, "phi", true));
	MP_CUDA(cudaSetDevice(ctx->device));
	const Dims d = dimsOf(flags);
	const unsigned int blocks = gridFor(d.n, 256);
	if (A0->prec == 4) k_make_matrix<float><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, dptr<float>(fractions), dptr<float>(phi), (float)gfClamp, (float*)A0->d, (float*)Ai->d, (float*)Aj->d, (float*)Ak->d);
	else               k_make_matrix<double><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, dptr<double>(fractions), dptr<double>(phi), gfClamp, (double*)A0->d, (double*)Ai->d, (double*)Aj->d, (double*)Ak->d);
	MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}
int mp_make_laplace_matrix(mp_context* ctx, const mp_grid* flags, mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak, const mp_grid* fractions)
{ return makeMatrix(ctx, flags, A0, Ai, Aj, Ak, fractions, nullptr, 0.); }
}
// fused variant used by solvePressureSystem (MakeLaplaceMatrix + ApplyGhostFluidDiagonal in one pass)
int mp_make_matrix_fused(mp_context* ctx, const mp_grid* flags, mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak, const mp_grid* fractions, const mp_grid* phi, double gfClamp)
{ return makeMatrix(ctx, flags, A0, Ai, Aj, Ak, fractions, phi, gfClamp); }

extern "C" {
int mp_apply_ghost_fluid_diagonal(mp_context* ctx, mp_grid* A0, const mp_grid* flags, const mp_grid* phi, double gfClamp)
{
	if (!ctx || !flags || !A0 || !phi) MP_FAIL(MP_ERR_INVALID, "mp_apply_ghost_fluid_diagonal: NULL argument");
	MP_TRY(mp_check_same(flags, A0, MP_GRID_REAL, "A0", false)); MP_TRY(mp_check_same(A0, phi, MP_GRID_REAL, "phi", false));
	const Dims d = dimsOf(flags);
	const unsigned int blocks = gridFor(d.n, 256);
	if (A0->prec == 4) k_ghost_diag<float><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (const float*)phi->d, (float)gfClamp, (float*)A0->d);
	else               k_ghost_diag<double><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (const double*)phi->d, gfClamp, (double*)A0->d);
	MP_CHECK_LAUNCH(ctx); return MP_OK;
}

int mp_count_empty_cells(mp_context* ctx, const mp_grid* flags, long long* numEmpty)
{
	if (!flags || flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_count_empty_cells: flags is not a FlagGrid");
	MP_TRY(scanFlags(ctx, flags));
	MP_CUDA(cudaMemcpyAsync(ctx->hScal + 8, ctx->dScal + 8, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	*numEmpty = (long long)((unsigned long long*)(ctx->hScal + 8))[0];
	return MP_OK;
}

int mp_choose_fix_cell(mp_context* ctx, const mp_grid* flags, long long* fixPidx)
{
	if (!flags || flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_choose_fix_cell: flags is not a FlagGrid");
	const Dims d = dimsOf(flags);
	MP_TRY(scanFlags(ctx, flags));
	long long* pIdx = (long long*)(ctx->dScal + 12);
	k_choose_fix<<<1, 1, 0, ctx->stream>>>(d, (const unsigned long long*)(ctx->dScal + 8), pIdx); MP_CHECK_LAUNCH(ctx);
	MP_CUDA(cudaMemcpyAsync(ctx->hScal + 12, pIdx, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	*fixPidx = *(long long*)(ctx->hScal + 12);
	return MP_OK;
}

int mp_fix_pressure(mp_context* ctx, long long fixPidx, double value, mp_grid* rhs, mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak)
{
	if (!rhs || !A0 || !Ai || !Aj || !Ak) MP_FAIL(MP_ERR_INVALID, "mp_fix_pressure: NULL argument");
	MP_TRY(mp_check_same(rhs, A0, MP_GRID_REAL, "A0", false)); MP_TRY(mp_check_same(rhs, Ai, MP_GRID_REAL, "Ai", false));
	MP_TRY(mp_check_same(rhs, Aj, MP_GRID_REAL, "Aj", false)); MP_TRY(mp_check_same(rhs, Ak, MP_GRID_REAL, "Ak", false));
	const Dims d = dimsOf(rhs);
	const IndexInt nGlobal = d.is3D ? d.Z * d.gsz : d.n;     // fixPidx indexes the GLOBAL grid
	if (fixPidx < d.Y + (d.is3D ? d.Z : 0) + 1 || fixPidx >= nGlobal - d.Y - (d.is3D ? d.Z : 0) - 1) MP_FAIL(MP_ERR_INVALID, "mp_fix_pressure: cell %lld has neighbours outside the grid", fixPidx);
	long long* pIdx = (long long*)(ctx->dScal + 13);
	MP_CUDA(cudaMemcpyAsync(pIdx, &fixPidx, sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
	MP_CUDA(cudaStreamSynchronize(ctx->stream));   // fixPidx lives on the caller's stack
	if (rhs->prec == 4) k_fix_pressure<float><<<1, 1, 0, ctx->stream>>>(d, pIdx, (float)value, (float*)rhs->d, (float*)A0->d, (float*)Ai->d, (float*)Aj->d, (float*)Ak->d);
	else                k_fix_pressure<double><<<1, 1, 0, ctx->stream>>>(d, pIdx, value, (double*)rhs->d, (double*)A0->d, (double*)Ai->d, (double*)Aj->d, (double*)Ak->d);
	MP_CHECK_LAUNCH(ctx); return MP_OK;
}

int mp_correct_velocity(mp_context* ctx, mp_grid* vel, const mp_grid* pressure, const mp_grid* flags,
                        const mp_grid* phi, const mp_grid* curv, const mp_pressure_params* params)
{
	if (!ctx || !vel || !pressure || !flags) MP_FAIL(MP_ERR_INVALID, "mp_correct_velocity: NULL argument");
	if (flags->kind != MP_GRID_FLAGS) MP_FAIL(MP_ERR_INVALID, "mp_correct_velocity: flags is not a FlagGrid");
	MP_TRY(mp_check_same(flags, pressure, MP_GRID_REAL, "pressure", false));
	MP_TRY(mp_check_same(pressure, vel, MP_GRID_MAC, "vel", false));
	MP_TRY(mp_check_same(pressure, phi, MP_GRID_REAL, "phi", true));
	MP_TRY(mp_check_same(pressure, curv, MP_GRID_REAL, "curv", true));
	mp_pressure_params def; if (!params) { mp_pressure_params_default(&def); params = &def; }
	MP_CUDA(cudaSetDevice(ctx->device));
	const Dims d = dimsOf(flags);
	const unsigned int blocks = gridFor(d.n, 256);
	const size_t planeReal = (size_t)d.Z * pressure->prec;
	if (d.world > 1) MP_TRY(mp_dist_halo(ctx, pressure->d, planeReal, pressure->sz));    // p(k-1) of the first owned plane
	if (vel->prec == 4) {
		if (phi) k_correct_velocity<float, true><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (float*)vel->d, (const float*)pressure->d, dptr<float>(phi), dptr<float>(curv), (float)params->gfClamp, (float)params->surfTens);
		else     k_correct_velocity<float, false><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (float*)vel->d, (const float*)pressure->d, nullptr, nullptr, (float)params->gfClamp, 0.f);
		MP_CHECK_LAUNCH(ctx);
		if (phi && d.world > 1) MP_TRY(mp_dist_halo(ctx, vel->d, planeReal * 3, vel->sz));   // k_replace_clamped reads neighbours' updated velocity
		if (phi) { k_replace_clamped<float><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (float*)vel->d, (const float*)phi->d, (float)params->gfClamp); MP_CHECK_LAUNCH(ctx); }
	} else {
		if (phi) k_correct_velocity<double, true><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (double*)vel->d, (const double*)pressure->d, dptr<double>(phi), dptr<double>(curv), params->gfClamp, params->surfTens);
		else     k_correct_velocity<double, false><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (double*)vel->d, (const double*)pressure->d, nullptr, nullptr, params->gfClamp, 0.);
		MP_CHECK_LAUNCH(ctx);
		if (phi && d.world > 1) MP_TRY(mp_dist_halo(ctx, vel->d, planeReal * 3, vel->sz));
		if (phi) { k_replace_clamped<double><<<blocks, 256, 0, ctx->stream>>>(d, (const int*)flags->d, (double*)vel->d, (const double*)phi->d, params->gfClamp); MP_CHECK_LAUNCH(ctx); }
	}
	return MP_OK;
}

} // extern "C"

// MIC(0) in BLOCK RED-BLACK ordering: the reformulated preconditioner north_star (4) names ("level-scheduled or colored triangular
// solves", iteration counts reported).  Selected with mp_set_mic_ordering(ctx, 1, tileY, tileZ) / MP_MIC_RB="TY,TZ"; the default PcMIC
// keeps the reference's lexicographic ordering (mp_mic.cu: bit-identical to conjugategrad.cpp:66-97,:135-159, but a chain of
// sx+sy+sz dependent hyperplanes -- DESIGN section 5).
//
// Ordering.  The y-z plane of rows is cut into tiles of TY x TZ rows (each tile holds its rows' whole x extent); tiles with even
// (J + K) come first ("red"), then the odd ones ("black"), cells in lexicographic order inside a tile.  A 7-point stencil couples tiles
// through faces only, so all tiles of a colour are independent: an application is four launches (forward red, forward black,
// backward black, backward red) without any hand-off between CTAs, and every cell's operands are read once.  The factor and the sweeps
// are those of an exact MIC(0) of the PERMUTED matrix (nothing is dropped); specification: the CPU restatement micrb_init / micrb_apply of the test oracle
// (with one tile it is the reference's MIC(0) bit for bit, tests/test_micrb.py), and the device results equal that specification bit
// for bit (tests/test_gpu_micrb.py).  Iteration counts rise by 1.25-1.6x over the lexicographic ordering, depending on the tile.
//
// Schedule.  One warp per tile; a tile is walked in sub-blocks of 8 x 4 rows (lane (lj, lk) owns row (8 sa + lj, 4 sb + lk)), rows in
// chunks of 32 bytes.  A lane trails its -y / -z neighbour lane by ONE CHUNK: in round T it works on chunk T - (lj + lk), takes the
// neighbours' products of that chunk by warp shuffle (they were formed a round earlier) and runs the x-recurrence of its chunk on
// registers -- no shared memory, no barrier, no mailbox.  Operands of round T+1 are requested in round T (two register sets).  Rows next
// to another tile (or to an earlier sub-block of the same tile) read the neighbour row's chunk from global memory.
// The matrix is the byte mask k_micrb_mask builds (every off-diagonal 0 or -1, verified on the arrays; anything else -- face
// fractions -- keeps the lexicographic kernels): per cell and application 4 Real read + 2 written + 2 mask bytes + the edge rows.
// Measured (512^3 float, one B200): 1.6-2.0 ms per application against 4.5 ms of the lexicographic sweeps, ~2 TB/s.
// ncu: 350-400 instructions per warp and 8-cell round at 3.5 cycles per instruction (dependent ALU chain + selects on the mask bits) and a
// third of the stall samples on the first use of the operands requested a round earlier.  Tried and dropped: L2 prefetch 2-4 rounds ahead
// (slower), a register cap for 16 warps per SM (spills, slower), a cp.async ring in shared memory 3 rounds deep (long-scoreboard stalls gone,
// but ~50 more instructions per round: 1.7-2.4 ms), the ring plus a mask-free path for rounds whose chunks are all fluid and fully coupled
// (1.7-2.3 ms: no pipe is busy -- ALU 37 %, FMA 10 %, LSU 13 % -- a warp simply issues every 3.5 cycles and there are 2-3 warps per scheduler).
#include "mp_common.cuh"
#include <cstdlib>

namespace {

struct RbGeom { int sx, sy, sz; IndexInt Y, Z; int nch, pitch, TY, TZ, nJ, nK, na, nb; };

template <typename Real> struct RbChunk { static constexpr int CH = 32 / (int)sizeof(Real); };
// mask bits of a cell: 1 fluid row, 2 / 4 coupled to -x / +x, 8 / 16 -y / +y, 32 / 64 -z / +z (coupled = the off-diagonal is -1)
enum : unsigned { mFluid = 1u, mXm = 2u, mXp = 4u, mYm = 8u, mYp = 16u, mZm = 32u, mZp = 64u };

template <typename Real>
__global__ void __launch_bounds__(256) k_micrb_mask(Dims d, int pitch, const int* __restrict__ flags, const Real* __restrict__ Ai, const Real* __restrict__ Aj,
	const Real* __restrict__ Ak, unsigned char* __restrict__ mask8, int* bad)
{
	const IndexInt t = (IndexInt)blockIdx.x * blockDim.x + threadIdx.x;
	const IndexInt rows = (IndexInt)d.sy * d.sz;
	if (t >= rows * pitch) return;
	const int i = (int)(t % pitch); const IndexInt row = t / pitch;
	unsigned m = 0;
	if (i < d.sx) {
		const int j = (int)(row % d.sy), k = (int)(row / d.sy);
		const IndexInt idx = (IndexInt)i + d.Y * j + d.Z * k;
		if (flags[idx] & TypeFluid) {
			m = mFluid;
			const Real c[6] = { Ai[idx - 1], Ai[idx], Aj[idx - d.Y], Aj[idx], Ak[idx - d.Z], Ak[idx] };
			#pragma unroll
			for (int q = 0; q < 6; q++) {
				if (c[q] == (Real)-1) m |= 2u << q;
				else if (c[q] != (Real)0) *bad = 1;
			}
		}
	}
	mask8[t] = (unsigned char)m;
}

// ---- chunk moves: CH cells = 32 bytes (VEC: rows are whole aligned chunks; else ragged rows, element by element)
template <typename Real, int CH, bool VEC>
__device__ __forceinline__ void ldChunk(const Real* row, int i0, int sx, Real (&v)[CH]) {
	if (VEC) {
		union { uint4 u[2]; Real r[CH]; } t;
		const uint4* p = reinterpret_cast<const uint4*>(row + i0);
		t.u[0] = __ldcg(p); t.u[1] = __ldcg(p + 1);
		#pragma unroll
		for (int s = 0; s < CH; s++) v[s] = t.r[s];
	} else {
		#pragma unroll
		for (int s = 0; s < CH; s++) v[s] = (i0 + s < sx) ? __ldcg(row + i0 + s) : (Real)0;
	}
}
template <typename Real, int CH, bool VEC>
__device__ __forceinline__ void stChunk(Real* row, int i0, unsigned fluidBits, const Real (&v)[CH]) {
	if (VEC && fluidBits == (1u << CH) - 1u) {
		union { uint4 u[2]; Real r[CH]; } t;
		#pragma unroll
		for (int s = 0; s < CH; s++) t.r[s] = v[s];
		uint4* p = reinterpret_cast<uint4*>(row + i0);
		p[0] = t.u[0]; p[1] = t.u[1];
	} else {
		#pragma unroll
		for (int s = 0; s < CH; s++) if ((fluidBits >> s) & 1u) row[i0 + s] = v[s];
	}
}
// the CH mask bytes of a chunk as one word (byte s = cell s); rows of the mask are padded to whole chunks
template <int CH>
__device__ __forceinline__ unsigned long long ldMask(const unsigned char* mrow, int i0) {
	if (CH == 8) return __ldg(reinterpret_cast<const unsigned long long*>(mrow + i0));
	return (unsigned long long)__ldg(reinterpret_cast<const unsigned int*>(mrow + i0));
}
__device__ __forceinline__ unsigned maskOf(unsigned long long M, int s) { return (unsigned)(M >> (8 * s)) & 0xffu; }
template <typename Real> __device__ __forceinline__ Real coup(unsigned m, unsigned bit) { return (m & bit) ? (Real)-1 : (Real)0; }

// sum of the couplings of a cell to ALL its successors (every entry 0 or -1: exact in any order).  ymS...: is the -y / +y / -z / +z
// neighbour a successor (it lies in the same tile on the plus side, or in another tile and this cell's tile is red)
template <typename Real>
__device__ __forceinline__ Real succSum(unsigned m, bool ymS, bool ypS, bool zmS, bool zpS) {
	int n = (m & mXp) ? 1 : 0;
	if (ymS && (m & mYm)) n++;
	if (ypS && (m & mYp)) n++;
	if (zmS && (m & mZm)) n++;
	if (zpS && (m & mZp)) n++;
	return (Real)(-n);
}
template <typename Real>
__device__ __forceinline__ Real micrbFactorEnd(Real e, Real inner, Real a0) {     // conjugategrad.cpp:89-95
	const Real tau = (Real)0.97, sigma = (Real)0.25;
	e = (Real)((double)e - (double)tau * ((double)inner + 0.));
	if (e < sigma * a0) e = a0;
	return (Real)(1. / (double)sqrt(e));
}

// MODE 0: factor (P written), 1: forward substitution (dst = L^-1 src), 2: backward substitution (dst = L^-T dst)
template <typename Real, int MODE, bool VEC>
__global__ void __launch_bounds__(32) k_micrb(RbGeom g, int colour, const unsigned char* __restrict__ mask8, Real* dst, const Real* __restrict__ src,
	Real* P, const Real* __restrict__ A0, const int* __restrict__ doneFlag)
{
	constexpr int CH = RbChunk<Real>::CH;
	constexpr unsigned FULL = 0xffffffffu;
	constexpr bool bwd = (MODE == 2);
	if (doneFlag && *doneFlag) return;
	const int J = blockIdx.x % g.nJ, K = blockIdx.x / g.nJ;
	if (((J + K) & 1) != colour) return;
	const int lane = threadIdx.x, lj = lane & 7, lk = lane >> 3;
	const int skew = bwd ? (7 - lj) + (3 - lk) : lj + lk;
	const int nRounds = g.nch + 10;
	const bool red = colour == 0;

	#pragma unroll 1
	for (int sbi = 0; sbi < g.nb; sbi++) {
	#pragma unroll 1
	for (int sai = 0; sai < g.na; sai++) {
		const int sa = bwd ? g.na - 1 - sai : sai, sb = bwd ? g.nb - 1 - sbi : sbi;
		const int jt = 8 * sa + lj, kt = 4 * sb + lk;                  // row inside the tile
		const int j = J * g.TY + jt, k = K * g.TZ + kt;
		const bool rowIn = j < g.sy && k < g.sz;
		const IndexInt rowOff = rowIn ? g.Y * j + g.Z * k : 0;
		const unsigned char* mrow = mask8 + (rowIn ? ((IndexInt)k * g.sy + j) * g.pitch : 0);
		// Y side / Z side neighbour row that is not a lane of this sub-block: -1 / +1 = the row at j-1 / j+1 takes part, 0 = none.
		//   forward / factor: predecessors = minus side inside the tile (an earlier sub-block), both sides across a tile face when black
		//   backward: successors = plus side inside the tile (a later sub-block, done before), both sides across a tile face when red
		int ySide = 0, zSide = 0; bool yOut = false, zOut = false;
		if (rowIn) {
			const bool yLo = jt == 0, yHi = jt == g.TY - 1, zLo = kt == 0, zHi = kt == g.TZ - 1;
			if (!bwd) {
				if (lj == 0) { if (!yLo) ySide = -1; else if (!red && j > 0) { ySide = -1; yOut = true; } }
				if (lj == 7 || yHi) { if (yHi && !red && j + 1 < g.sy && ySide == 0) { ySide = 1; yOut = true; } }
				if (lk == 0) { if (!zLo) zSide = -1; else if (!red && k > 0) { zSide = -1; zOut = true; } }
				if (lk == 3 || zHi) { if (zHi && !red && k + 1 < g.sz && zSide == 0) { zSide = 1; zOut = true; } }
			} else {
				if (lj == 7) { if (!yHi && j + 1 < g.sy) ySide = 1; else if (yHi && red && j + 1 < g.sy) { ySide = 1; yOut = true; } }
				if (lj == 0 && yLo && red && j > 0 && ySide == 0) { ySide = -1; yOut = true; }
				if (lk == 3) { if (!zHi && k + 1 < g.sz) zSide = 1; else if (zHi && red && k + 1 < g.sz) { zSide = 1; zOut = true; } }
				if (lk == 0 && zLo && red && k > 0 && zSide == 0) { zSide = -1; zOut = true; }
			}
		}
		// out-of-tile terms are taken in the order -y, -z, +y, +z (as in the specification): the y term first unless it is +y and the z term is -z
		const bool yFirst = !(ySide == 1 && zSide == -1);
		const IndexInt eyOff = rowOff + (IndexInt)ySide * g.Y, ezOff = rowOff + (IndexInt)zSide * g.Z;
		const unsigned char* eyM = mrow + (IndexInt)ySide * g.pitch; const unsigned char* ezM = mrow + (IndexInt)zSide * g.sy * g.pitch;
		const unsigned yBit = ySide < 0 ? mYm : mYp, zBit = zSide < 0 ? mZm : mZp;    // the cell's own coupling towards that row
		// in-warp neighbours exist for lj > 0 / lk > 0 (forward), lj < 7 / lk < 3 (backward)
		const bool inY = bwd ? lj < 7 : lj > 0, inZ = bwd ? lk < 3 : lk > 0;
		// successor relations of this row's cells (factor): see succSum
		const bool ymS = (jt == 0) ? red : false, ypS = (jt == g.TY - 1) ? red : true, zmS = (kt == 0) ? red : false, zpS = (kt == g.TZ - 1) ? red : true;
		// ... and of the cells of the edge rows (factor).  A row across a tile face lies in a tile of the other colour, at the opposite face.
		bool eYymS = false, eYypS = true, eYzmS = zmS, eYzpS = zpS, eZymS = ymS, eZypS = ypS, eZzmS = false, eZzpS = true;
		if (MODE == 0) {
			if (yOut) { eYymS = ySide > 0 ? !red : false; eYypS = ySide < 0 ? !red : true; eYzmS = (kt == 0) ? !red : false; eYzpS = (kt == g.TZ - 1) ? !red : true; }
			else if (ySide) { const int jn = jt + ySide; eYymS = (jn == 0) ? red : false; eYypS = (jn == g.TY - 1) ? red : true; }
			if (zOut) { eZzmS = zSide > 0 ? !red : false; eZzpS = zSide < 0 ? !red : true; eZymS = (jt == 0) ? !red : false; eZypS = (jt == g.TY - 1) ? !red : true; }
			else if (zSide) { const int kn = kt + zSide; eZzmS = (kn == 0) ? red : false; eZzpS = (kn == g.TZ - 1) ? red : true; }
		}

		Real tx = (Real)0, txS = (Real)0;          // carried along the row: forward the product of the previous cell, backward its value; factor: its P and successor sum
		Real oy[CH], oz[CH], oyS[CH], ozS[CH];     // handed to the next lanes: forward products, backward values, factor P (+ successor sums)
		#pragma unroll
		for (int s = 0; s < CH; s++) { oy[s] = oz[s] = (Real)0; oyS[s] = ozS[s] = (Real)0; }
		// warp-uniform: does any lane of this sub-block read an edge row at all (red tiles of one sub-block never do in the forward sweep)
		const bool anyY = __any_sync(FULL, ySide != 0), anyZ = __any_sync(FULL, zSide != 0);

		// Operands of a round live in one of two register sets: round T computes on one while the loads of round T+1 land in the other.
		struct Ops { Real P[CH], R[CH], EY[CH], EYP[CH], EZ[CH], EZP[CH]; unsigned long long M, EYM, EZM; };
		auto request = [&](int c, Ops& o) {
			const bool act = rowIn && (unsigned)c < (unsigned)g.nch;
			o.M = o.EYM = o.EZM = 0;                                // a round without fluid cells computes on stale operands and selects zeros
			if (!act) return;
			const int i0 = CH * (bwd ? g.nch - 1 - c : c);
			o.M = ldMask<CH>(mrow, i0);                             // (no test of the mask here: a branch on it would put its latency in front of the loads below)
			if (MODE == 0) ldChunk<Real, CH, VEC>(A0 + rowOff, i0, g.sx, o.R);
			else { ldChunk<Real, CH, VEC>(P + rowOff, i0, g.sx, o.P); ldChunk<Real, CH, VEC>((MODE == 1 ? src : (const Real*)dst) + rowOff, i0, g.sx, o.R); }
			if (ySide) {
				if (MODE == 0) { ldChunk<Real, CH, VEC>(P + eyOff, i0, g.sx, o.EYP); o.EYM = ldMask<CH>(eyM, i0); }
				else { ldChunk<Real, CH, VEC>((const Real*)dst + eyOff, i0, g.sx, o.EY); if (MODE == 1) ldChunk<Real, CH, VEC>(P + eyOff, i0, g.sx, o.EYP); }
			}
			if (zSide) {
				if (MODE == 0) { ldChunk<Real, CH, VEC>(P + ezOff, i0, g.sx, o.EZP); o.EZM = ldMask<CH>(ezM, i0); }
				else { ldChunk<Real, CH, VEC>((const Real*)dst + ezOff, i0, g.sx, o.EZ); if (MODE == 1) ldChunk<Real, CH, VEC>(P + ezOff, i0, g.sx, o.EZP); }
			}
		};
		// Arithmetic.  Every off-diagonal a is 0 or -1, so z a P is -(z P) or a zero and "acc - z a P" is "acc + z P" or acc: the products
		// below are formed without the coefficient and added where the coupling bit is set -- the value the specification gets (signs of
		// exact zeros aside, which no later value depends on).
		auto compute = [&](int c, const Ops& o) {
			const bool act = rowIn && (unsigned)c < (unsigned)g.nch && o.M != 0;
			Real q[CH]; unsigned fluidBits = 0;
			// edge rows: products of this round's cells, formed off the dependent chain; 0 where the row does not take part
			Real ey[CH], ez[CH], sY[CH], sZ[CH];
			#pragma unroll
			for (int s = 0; s < CH; s++) { ey[s] = ez[s] = (Real)0; sY[s] = sZ[s] = (Real)0; }
			if (anyY) {
				#pragma unroll
				for (int s = 0; s < CH; s++) {
					const bool cp = ySide && (maskOf(o.M, s) & yBit);
					if (MODE == 1) ey[s] = cp ? o.EY[s] * o.EYP[s] : (Real)0;
					if (MODE == 2) ey[s] = cp ? o.EY[s] * o.P[s] : (Real)0;
					if (MODE == 0) { ey[s] = cp ? o.EYP[s] : (Real)0; sY[s] = succSum<Real>(maskOf(o.EYM, s), eYymS, eYypS, eYzmS, eYzpS); }
				}
			}
			if (anyZ) {
				#pragma unroll
				for (int s = 0; s < CH; s++) {
					const bool cp = zSide && (maskOf(o.M, s) & zBit);
					if (MODE == 1) ez[s] = cp ? o.EZ[s] * o.EZP[s] : (Real)0;
					if (MODE == 2) ez[s] = cp ? o.EZ[s] * o.P[s] : (Real)0;
					if (MODE == 0) { ez[s] = cp ? o.EZP[s] : (Real)0; sZ[s] = succSum<Real>(maskOf(o.EZM, s), eZymS, eZypS, eZzmS, eZzpS); }
				}
			}
			#pragma unroll
			for (int ss = 0; ss < CH; ss++) {
				const int s = bwd ? CH - 1 - ss : ss;                    // processing order along the row
				// the neighbour lanes' output for this cell (they worked on this chunk one round ago); every lane takes part
				Real iy = bwd ? __shfl_down_sync(FULL, oy[s], 1) : __shfl_up_sync(FULL, oy[s], 1);
				Real iz = bwd ? __shfl_down_sync(FULL, oz[s], 8) : __shfl_up_sync(FULL, oz[s], 8);
				Real iyS = (Real)0, izS = (Real)0;
				if (MODE == 0) { iyS = __shfl_up_sync(FULL, oyS[s], 1); izS = __shfl_up_sync(FULL, ozS[s], 8); }
				const unsigned m = maskOf(o.M, s);
				const bool fl = act && (m & mFluid);
				if (fl) fluidBits |= 1u << s;
				Real out = (Real)0;
				if (MODE == 1) {
					// iy / iz / tx arrive as z P of the neighbour, already zero where THAT cell has no coupling towards this one
					if (!inY) iy = yOut ? (Real)0 : ey[s];
					if (!inZ) iz = zOut ? (Real)0 : ez[s];
					Real acc = o.R[s];
					if (yOut && yFirst) acc = acc + ey[s];
					if (zOut) acc = acc + ez[s];
					if (yOut && !yFirst) acc = acc + ey[s];
					acc = ((acc + tx) + iy) + iz;
					const Real p = o.P[s];
					out = fl ? p * acc : (Real)0;
					const Real t = out * p;
					tx = (m & mXp) ? t : (Real)0; oy[s] = (m & mYp) ? t : (Real)0; oz[s] = (m & mZp) ? t : (Real)0;
				} else if (MODE == 2) {
					const Real p = o.P[s];
					if (!inY) iy = (Real)0;
					if (!inZ) iz = (Real)0;
					const Real px = (m & mXp) ? tx * p : (Real)0;
					const Real py = inY ? ((m & mYp) ? iy * p : (Real)0) : (yOut ? (Real)0 : ey[s]);
					const Real pz = inZ ? ((m & mZp) ? iz * p : (Real)0) : (zOut ? (Real)0 : ez[s]);
					Real acc = o.R[s];
					if (yOut && yFirst) acc = acc + ey[s];
					if (zOut) acc = acc + ez[s];
					if (yOut && !yFirst) acc = acc + ey[s];
					acc = ((acc + px) + py) + pz;
					out = fl ? p * acc : (Real)0;
					tx = out; oy[s] = out; oz[s] = out;
				} else {
					// factor: a predecessor n contributes (a P_n)^2 to e and a (S_n - a) P_n^2 to the bracket, S_n = sum over all successors of n
					Real e = o.R[s], inner = (Real)0;
					auto pred = [&](bool coupled, Real pn, Real sn) { if (coupled) { e = e - pn * pn; inner = inner + (Real)-1 * (sn - (Real)-1) * (pn * pn); } };
					if (yOut && yFirst) pred(m & yBit, o.EYP[s], sY[s]);
					if (zOut) pred(m & zBit, o.EZP[s], sZ[s]);
					if (yOut && !yFirst) pred(m & yBit, o.EYP[s], sY[s]);
					pred(m & mXm, tx, txS);
					if (inY) pred(m & mYm, iy, iyS); else if (ySide && !yOut) pred(m & yBit, o.EYP[s], sY[s]);
					if (inZ) pred(m & mZm, iz, izS); else if (zSide && !zOut) pred(m & zBit, o.EZP[s], sZ[s]);
					out = fl ? micrbFactorEnd<Real>(e, inner, o.R[s]) : (Real)0;
					const Real sn = fl ? succSum<Real>(m, ymS, ypS, zmS, zpS) : (Real)0;
					tx = out; txS = sn; oy[s] = out; oyS[s] = sn; oz[s] = out; ozS[s] = sn;
				}
				q[s] = out;
			}
			if (fluidBits) stChunk<Real, CH, VEC>((MODE == 0 ? P : dst) + rowOff, CH * (bwd ? g.nch - 1 - c : c), fluidBits, q);
		};
		__syncwarp();                        // rows of the sub-blocks before this one are written
		Ops A, B;
		#pragma unroll
		for (int s = 0; s < CH; s++) { A.P[s] = A.R[s] = A.EY[s] = A.EYP[s] = A.EZ[s] = A.EZP[s] = (Real)0; B.P[s] = B.R[s] = B.EY[s] = B.EYP[s] = B.EZ[s] = B.EZP[s] = (Real)0; }
		request(-skew, A);
		#pragma unroll 1
		for (int T = 0; T < nRounds; T += 2) {
			request(T + 1 - skew, B);
			compute(T - skew, A);
			request(T + 2 - skew, A);
			compute(T + 1 - skew, B);
		}
	}
	}
}

struct RbState {     // per context (keyed by the context pointer: one solver thread per context)
	unsigned char* mask8 = nullptr; size_t maskBytes = 0;
	const void* forFlags = nullptr; const void* forP = nullptr; int prec = 0; bool valid = false;
	RbGeom g;
};

}  // namespace

struct mp_micrb_state { RbState s; };

static int rbTiles(const mp_context* ctx, int& ty, int& tz) {
	ty = ctx->micRbTY; tz = ctx->micRbTZ;
	const char* e = getenv("MP_MIC_RB");          // "TY,TZ" switches the ordering on for every context (read per solve)
	int on = ctx->micRb;
	if (e && *e) { int a = 0, b = 0; if (sscanf(e, "%d,%d", &a, &b) == 2 && a > 0 && b > 0) { on = 1; ty = a; tz = b; } else on = atoi(e) != 0; }
	ty = ty > 0 ? (ty + 7) / 8 * 8 : 0; tz = tz > 0 ? (tz + 3) / 4 * 4 : 0;     // 0: chosen from the grid (rbAutoTiles)
	return on;
}

// Tile of a grid when none is given: larger tiles need fewer iterations (512^3: 612 / 571 / 483 with 8x4 / 8x8 / 16x8) but there is one warp per tile and
// only half of the tiles run at a time, so the largest tile that still leaves ~900 warps per launch (measured: 16x8 at 512^3, 8x4 at 256^3).
static void rbAutoTiles(const Dims& d, int& ty, int& tz) {
	static const int cand[3][2] = { { 16, 8 }, { 8, 8 }, { 8, 4 } };
	for (int q = 0; q < 3; q++) {
		ty = cand[q][0]; tz = cand[q][1];
		if ((long long)((d.sy + ty - 1) / ty) * ((d.sz + tz - 1) / tz) / 2 >= 900) return;
	}
}

// Is the block red-black ordering selected, and can it run on this matrix?  Builds the byte mask on the way (once per factorisation).
int mp_micrb_prepare(mp_context* ctx, const mp_grid* flags, const mp_grid* P, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak, bool* use)
{
	*use = false;
	if (ctx->micRbState) ctx->micRbState->s.valid = false;       // a new factorisation: whatever the state said about an earlier factor is void
	int ty, tz;
	if (!rbTiles(ctx, ty, tz)) return MP_OK;
	const Dims d = dimsOf(flags);
	if (!d.is3D || d.world > 1 || d.sx < 3 || d.sy < 3 || d.sz < 3) return MP_OK;
	if (ty == 0 || tz == 0) rbAutoTiles(d, ty, tz);
	if (!ctx->micRbState) ctx->micRbState = new mp_micrb_state();
	RbState& st = ctx->micRbState->s;
	const int CH = 32 / P->prec;
	RbGeom g; g.sx = d.sx; g.sy = d.sy; g.sz = d.sz; g.Y = d.Y; g.Z = d.Z; g.nch = (d.sx + CH - 1) / CH; g.pitch = (g.nch * CH + 7) / 8 * 8;
	g.TY = ty; g.TZ = tz; g.nJ = (d.sy + ty - 1) / ty; g.nK = (d.sz + tz - 1) / tz; g.na = ty / 8; g.nb = tz / 4;
	const size_t need = (size_t)g.pitch * d.sy * d.sz;
	if (st.maskBytes < need) {
		if (st.mask8) { MP_CUDA(cudaStreamSynchronize(ctx->stream)); MP_CUDA(cudaFree(st.mask8)); st.mask8 = nullptr; st.maskBytes = 0; }
		MP_CUDA(cudaMalloc((void**)&st.mask8, need)); st.maskBytes = need;
	}
	int* bad = (int*)(ctx->dScal + 21);
	MP_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
	const unsigned blocks = (unsigned)((need + 255) / 256);
	if (P->prec == 4) k_micrb_mask<float><<<blocks, 256, 0, ctx->stream>>>(d, g.pitch, (const int*)flags->d, (const float*)Ai->d, (const float*)Aj->d, (const float*)Ak->d, st.mask8, bad);
	else              k_micrb_mask<double><<<blocks, 256, 0, ctx->stream>>>(d, g.pitch, (const int*)flags->d, (const double*)Ai->d, (const double*)Aj->d, (const double*)Ak->d, st.mask8, bad);
	MP_CHECK_LAUNCH(ctx);
	MP_CUDA(cudaMemcpyAsync(ctx->hScal + 21, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	MP_CUDA(cudaStreamSynchronize(ctx->stream));
	st.valid = *(int*)(ctx->hScal + 21) == 0;      // an off-diagonal that is neither 0 nor -1 (face fractions): the lexicographic kernels take over
	st.forFlags = flags; st.forP = P; st.prec = P->prec; st.g = g;
	*use = st.valid;
	return MP_OK;
}

bool mp_micrb_active(const mp_context* ctx, const mp_grid* flags, const mp_grid* P) {
	if (!ctx->micRbState) return false;
	const RbState& st = ctx->micRbState->s;
	(void)flags;                 // the factor grid identifies the factorisation; the mask built with it already holds what the sweeps need of the flags
	return st.valid && st.forP == P && st.prec == P->prec && st.g.sx == P->sx && st.g.sy == P->sy && st.g.sz == P->sz;
}
void mp_micrb_release(mp_context* ctx) {
	if (!ctx->micRbState) return;
	if (ctx->micRbState->s.mask8) cudaFree(ctx->micRbState->s.mask8);
	delete ctx->micRbState; ctx->micRbState = nullptr;
}
void mp_micrb_tiles(const mp_context* ctx, int* ty, int* tz) {
	*ty = *tz = 0;
	if (ctx->micRbState && ctx->micRbState->s.valid) { *ty = ctx->micRbState->s.g.TY; *tz = ctx->micRbState->s.g.TZ; }
}

template <typename Real, int MODE>
static int rbLaunch(mp_context* ctx, const RbState& st, int colour, Real* dst, const Real* src, Real* P, const Real* A0, const int* doneFlag) {
	constexpr int CH = RbChunk<Real>::CH;
	const RbGeom& g = st.g;
	const unsigned grid = (unsigned)(g.nJ * g.nK);
	if (g.sx % CH == 0) k_micrb<Real, MODE, true><<<grid, 32, 0, ctx->stream>>>(g, colour, st.mask8, dst, src, P, A0, doneFlag);
	else                k_micrb<Real, MODE, false><<<grid, 32, 0, ctx->stream>>>(g, colour, st.mask8, dst, src, P, A0, doneFlag);
	MP_CHECK_LAUNCH(ctx);
	return MP_OK;
}

int mp_micrb_init_launch(mp_context* ctx, mp_grid* P, const mp_grid* A0)
{
	const RbState& st = ctx->micRbState->s;
	MP_CUDA(cudaMemsetAsync(P->d, 0, P->bytes, ctx->stream));
	for (int colour = 0; colour < 2; colour++) {
		if (P->prec == 4) MP_TRY((rbLaunch<float, 0>(ctx, st, colour, nullptr, nullptr, (float*)P->d, (const float*)A0->d, nullptr)));
		else              MP_TRY((rbLaunch<double, 0>(ctx, st, colour, nullptr, nullptr, (double*)P->d, (const double*)A0->d, nullptr)));
	}
	return MP_OK;
}

int mp_micrb_apply_launch(mp_context* ctx, mp_grid* dst, const mp_grid* var1, const mp_grid* P, const int* doneFlag)
{
	const RbState& st = ctx->micRbState->s;
	for (int pass = 0; pass < 4; pass++) {
		const int colour = (pass == 0 || pass == 3) ? 0 : 1;
		if (dst->prec == 4) {
			if (pass < 2) MP_TRY((rbLaunch<float, 1>(ctx, st, colour, (float*)dst->d, (const float*)var1->d, (float*)P->d, nullptr, doneFlag)));
			else          MP_TRY((rbLaunch<float, 2>(ctx, st, colour, (float*)dst->d, nullptr, (float*)P->d, nullptr, doneFlag)));
		} else {
			if (pass < 2) MP_TRY((rbLaunch<double, 1>(ctx, st, colour, (double*)dst->d, (const double*)var1->d, (double*)P->d, nullptr, doneFlag)));
			else          MP_TRY((rbLaunch<double, 2>(ctx, st, colour, (double*)dst->d, nullptr, (double*)P->d, nullptr, doneFlag)));
		}
	}
	return MP_OK;
}

extern "C" {

// mode 0: the reference's lexicographic ordering (default; bit-identical to conjugategrad.cpp:66-159), 1: block red-black ordering with tiles of
// tileY x tileZ rows (rounded up to multiples of 8 and 4; 0 = chosen from the grid size).  Takes effect at the next factorisation.
int mp_set_mic_ordering(mp_context* ctx, int mode, int tileY, int tileZ)
{
	if (!ctx) MP_FAIL(MP_ERR_INVALID, "mp_set_mic_ordering: NULL context");
	if (mode != 0 && mode != 1) MP_FAIL(MP_ERR_INVALID, "mp_set_mic_ordering: mode %d (0 lexicographic, 1 block red-black)", mode);
	if (tileY < 0 || tileZ < 0) MP_FAIL(MP_ERR_INVALID, "mp_set_mic_ordering: negative tile size");
	ctx->micRb = mode; ctx->micRbTY = tileY; ctx->micRbTZ = tileZ;
	return MP_OK;
}

// what the last MIC(0) factorisation of this context used: *mode 1 and the tile when it ran in block red-black ordering, 0 / 0 / 0 otherwise
int mp_get_mic_ordering(const mp_context* ctx, int* mode, int* tileY, int* tileZ)
{
	if (!ctx || !mode || !tileY || !tileZ) MP_FAIL(MP_ERR_INVALID, "mp_get_mic_ordering: NULL argument");
	mp_micrb_tiles(ctx, tileY, tileZ);
	*mode = *tileY > 0 ? 1 : 0;
	return MP_OK;
}

}

// TEST INFRASTRUCTURE ONLY -- never linked into the product path.
//
// C entry points over the UNMODIFIED reference implementation (thunil/mantaflow, built
// -DNOPYTHON from /root/reference by oracle/Makefile into oracle/_ref/).  The functions
// here wrap caller-owned arrays in the reference's own Grid<T> objects (external-data
// constructors, grid.cpp:62-72) and call the reference's own kernels / plugins:
//   computePressureRhs      plugin/pressure.cpp:277-299
//   MakeLaplaceMatrix       conjugategrad.h:154-187
//   ApplyGhostFluidDiagonal plugin/pressure.cpp:136-151
//   fixPressure             plugin/pressure.cpp:226-245
//   GridCg<ApplyMatrix[2D]> conjugategrad.cpp:201-339
//   GridMg                  multigrid.cpp
//   correctVelocity         plugin/pressure.cpp:455-476
//   solvePressure           plugin/pressure.cpp:480-521
//   setWallBcs              plugin/extforces.cpp:307-316
//   addGravity, addBuoyancy plugin/extforces.cpp:61-90
//   advectSemiLagrange      plugin/advection.cpp:442-461
//   cgSolveWE               plugin/waves.cpp:86-147
//   PD_fluid_guiding        plugin/fluidguiding.cpp:294-353
//   extrapolateMACSimple, extrapolateLsSimple, extrapolateVec3Simple   fastmarch.cpp:337-375, :470-542
//   FlagGrid::updateFromLevelset, Grid<T>::setBound                    grid.cpp:844-854, :591-593
//   updateFractions, setObstacleFlags                                  plugin/initplugins.cpp:437-440,:473-475
//   markFluidCells, gridParticleIndex, unionParticleLevelset, mapPartsToMAC, mapMACToParts, flipVelocityUpdate   plugin/flip.cpp:158-177,:260-350,:562-595,:643-677
//   LaplaceOp, CurvatureOp (getLaplacian / getCurvature)               commonkernels.h:75-101, plugin/flip.cpp:710-716
//   Grid<T>::save / load (.uni, .raw, .npz)                            grid.cpp:113-156, fileio/iogrids.cpp
// Nothing of the reference is copied: its sources are compiled where they lie.
//
// The signatures are shared with oracle/mf_oracle.c (the restatement) so the same
// python driver (tests/refapi.py) can run either.
#include <vector>
#include <string>
#include <map>
#include <sstream>
#include <iostream>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <algorithm>
#include <omp.h>

#define private public      // dump GridMg internals (levels, types, operators) for parity tests
#include "multigrid.h"
#undef private

// pull in the preprocessed plugin so file-local kernels (ApplyGhostFluidDiagonal,
// CountEmptyCells, MakeRhs, fixPressure, gMapMG ...) are visible to the harness
#include "plugin/pressure.cpp"
#include "commonkernels.h"
#include "levelset.h"
#include "particle.h"
#include "vortexsheet.h"

namespace Manta {
// --- link stubs for the OpenVDB entry points referenced by grid.cpp (the .uni / .raw / .npz readers and writers are the reference's own, fileio/iogrids.cpp) ---
int writeObjectsVDB(const std::string&, std::vector<PbClass*>*, float, bool, int, bool) { return 0; }
int readObjectsVDB (const std::string&, std::vector<PbClass*>*, float) { return 0; }
void setWallBcs(const FlagGrid& flags, MACGrid& vel, const MACGrid* obvel, const MACGrid* fractions, const Grid<Real>* phiObs, int boundaryWidth);
void cgSolveDiffusion(const FlagGrid& flags, GridBase& grid, Real alpha, Real cgMaxIterFac, Real cgAccuracy);
void cgSolveWE(const FlagGrid& flags, Grid<Real>& ut, Grid<Real>& utm1, Grid<Real>& out, bool crankNic, Real cSqr, Real cgMaxIterFac, Real cgAccuracy);
void PD_fluid_guiding(MACGrid& vel, MACGrid& velT, Grid<Real>& pressure, FlagGrid& flags, Grid<Real>& weight, int blurRadius, Real theta, Real tau, Real sigma,
	Real epsRel, Real epsAbs, int maxIters, Grid<Real>* phi, Grid<Real>* perCellCorr, MACGrid* fractions, MACGrid* obvel, Real gfClamp, Real cgMaxIterFac,
	Real cgAccuracy, int preconditioner, bool zeroPressureFixing, const Grid<Real>* curv, const Real surfTens);
void releaseBlurPrecomp();
void extrapolateMACSimple(FlagGrid& flags, MACGrid& vel, int distance, LevelsetGrid* phiObs, bool intoObs);
void extrapolateLsSimple(Grid<Real>& phi, int distance, bool inside);
void extrapolateMACFromWeight(MACGrid& vel, Grid<Vec3>& weight, int distance);
void addForcePvel(ParticleDataImpl<Vec3>& vel, const Vec3& a, const Real dt, const ParticleDataImpl<int>* ptype, const int exclude);
void updateVelocityFromDeltaPos(const BasicParticleSystem& parts, ParticleDataImpl<Vec3>& vel, const ParticleDataImpl<Vec3>& x_prev, const Real dt, const ParticleDataImpl<int>* ptype, const int exclude);
void eulerStep(BasicParticleSystem& parts, const ParticleDataImpl<Vec3>& vel, const ParticleDataImpl<int>* ptype, const int exclude);
void setPartType(const BasicParticleSystem& parts, ParticleDataImpl<int>& ptype, const int mark, const int stype, const FlagGrid& flags, const int cflag);
void markIsolatedFluidCell(FlagGrid& flags, const int mark);
void pushOutofObs(BasicParticleSystem& parts, const FlagGrid& flags, const Grid<Real>& phiObs, const Real shift, const Real thresh, const ParticleDataImpl<int>* ptype, const int exclude);
void markFluidCells(const BasicParticleSystem& parts, FlagGrid& flags, const Grid<Real>* phiObs, const ParticleDataImpl<int>* ptype, const int exclude);
void gridParticleIndex(const BasicParticleSystem& parts, ParticleIndexSystem& indexSys, const FlagGrid& flags, Grid<int>& index, Grid<int>* counter);
void unionParticleLevelset(const BasicParticleSystem& parts, const ParticleIndexSystem& indexSys, const FlagGrid& flags, const Grid<int>& index, LevelsetGrid& phi,
	const Real radiusFactor, const ParticleDataImpl<int>* ptype, const int exclude);
void mapPartsToMAC(const FlagGrid& flags, MACGrid& vel, MACGrid& velOld, const BasicParticleSystem& parts, const ParticleDataImpl<Vec3>& partVel, Grid<Vec3>* weight,
	const ParticleDataImpl<int>* ptype, const int exclude);
void mapMACToParts(const FlagGrid& flags, const MACGrid& vel, const BasicParticleSystem& parts, ParticleDataImpl<Vec3>& partVel, const ParticleDataImpl<int>* ptype, const int exclude);
void flipVelocityUpdate(const FlagGrid& flags, const MACGrid& vel, const MACGrid& velOld, const BasicParticleSystem& parts, ParticleDataImpl<Vec3>& partVel, const Real flipRatio,
	const ParticleDataImpl<int>* ptype, const int exclude);
void updateFractions(const FlagGrid& flags, const Grid<Real>& phiObs, MACGrid& fractions, const int& boundaryWidth, const Real fracThreshold);
void setObstacleFlags(FlagGrid& flags, const Grid<Real>& phiObs, const MACGrid* fractions, const Grid<Real>* phiOut, const Grid<Real>* phiIn, int boundaryWidth);
void extrapolateVec3Simple(Grid<Vec3>& vel, Grid<Real>& phi, int distance, bool inside);
void addGravity(const FlagGrid& flags, MACGrid& vel, Vec3 gravity, const Grid<Real>* exclude, bool scale);
void addBuoyancy(const FlagGrid& flags, const Grid<Real>& density, MACGrid& vel, Vec3 gravity, Real coefficient, bool scale);
void advectSemiLagrange(const FlagGrid* flags, const MACGrid* vel, GridBase* grid, int order, Real strength, int orderSpace, bool openBounds, int boundaryWidth, int clampMode, int orderTrace);
void VICintegration(VortexSheetMesh& mesh, Real sigma, Grid<Vec3>& vel, const FlagGrid& flags, Grid<Vec3>* vorticity, Real cgMaxIterFac, Real cgAccuracy, Real scale, int precondition);
void InitPreconditionIncompCholesky(const FlagGrid& flags, Grid<Real>& A0, Grid<Real>& Ai, Grid<Real>& Aj, Grid<Real>& Ak, Grid<Real>& orgA0, Grid<Real>& orgAi, Grid<Real>& orgAj, Grid<Real>& orgAk);
void ApplyPreconditionIncompCholesky(Grid<Real>& dst, Grid<Real>& Var1, const FlagGrid& flags, Grid<Real>& A0, Grid<Real>& Ai, Grid<Real>& Aj, Grid<Real>& Ak, Grid<Real>& orgA0, Grid<Real>& orgAi, Grid<Real>& orgAj, Grid<Real>& orgAk);
void InitPreconditionModifiedIncompCholesky2(const FlagGrid& flags, Grid<Real>& Aprecond, Grid<Real>& A0, Grid<Real>& Ai, Grid<Real>& Aj, Grid<Real>& Ak);
void ApplyPreconditionModifiedIncompCholesky2(Grid<Real>& dst, Grid<Real>& Var1, const FlagGrid& flags, Grid<Real>& Aprecond, Grid<Real>& A0, Grid<Real>& Ai, Grid<Real>& Aj, Grid<Real>& Ak);
}

using namespace Manta;

static std::string gLastError;
static thread_local std::stringstream gLog;

struct CoutCapture {   // the reference reports iteration counts only through debMsg (pressure.cpp:440)
	std::streambuf* old; std::stringstream ss;
	CoutCapture() { old = std::cout.rdbuf(ss.rdbuf()); }
	~CoutCapture() { std::cout.rdbuf(old); }
};

static FluidSolver* mkSolver(int sx, int sy, int sz) {
	return new FluidSolver(Vec3i(sx, sy, sz), sz > 1 ? 3 : 2);
}

#define TRY try {
#define CATCH } catch (std::exception& e) { gLastError = e.what(); return 1; } return 0;

template <class G> static int saveLoad(G& g, const char* name, int load) { return load ? g.load(name) : g.save(name); }

extern "C" {

const char* ref_last_error() { return gLastError.c_str(); }
int ref_real_size() { return (int)sizeof(Real); }
int ref_is_reference() { return 1; }
int ref_set_debug_level(int l) { gDebugLevel = l; return 0; }

int ref_set_wall_bcs(int sx, int sy, int sz, const int* flags, Real* vel)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags); MACGrid V(s, (Vec3*)vel);
	  setWallBcs(F, V, 0, 0, 0, 0); }
	delete s;
  CATCH }

int ref_set_wall_bcs_obvel(int sx, int sy, int sz, const int* flags, Real* vel, const Real* obvel)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags); MACGrid V(s, (Vec3*)vel);
	  MACGrid* Ov = obvel ? new MACGrid(s, (Vec3*)obvel) : 0;
	  setWallBcs(F, V, Ov, 0, 0, 0);
	  delete Ov; }
	delete s;
  CATCH }

// the second-order variant (KnSetWallBcsFrac extforces.cpp:220-303) is taken when both fractions and phiObs are given; it never reads the fractions
int ref_set_wall_bcs_frac(int sx, int sy, int sz, const int* flags, Real* vel, const Real* phiObs)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	const size_t n = (size_t)sx * sy * sz;
	{ FlagGrid F(s, (int*)flags); Grid<Real> P(s, (Real*)phiObs); MACGrid Fr(s);
	  MACGrid V(s); memcpy(&V[0], vel, n * sizeof(Vec3));            // the plugin swaps its result in: work on a solver-owned copy
	  setWallBcs(F, V, 0, &Fr, &P, 0);
	  memcpy(vel, &V[0], n * sizeof(Vec3)); }
	delete s;
  CATCH }

int ref_add_gravity(int sx, int sy, int sz, const int* flags, Real* vel, double gx, double gy, double gz, const Real* exclude, int scale, double dt)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz); s->mDt = (Real)dt;
	{ FlagGrid F(s, (int*)flags); MACGrid V(s, (Vec3*)vel);
	  Grid<Real>* E = exclude ? new Grid<Real>(s, (Real*)exclude) : 0;
	  addGravity(F, V, Vec3((Real)gx, (Real)gy, (Real)gz), E, scale != 0);
	  delete E; }
	delete s;
  CATCH }

int ref_add_buoyancy(int sx, int sy, int sz, const int* flags, const Real* density, Real* vel, double gx, double gy, double gz, double coefficient, int scale, double dt)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz); s->mDt = (Real)dt;
	{ FlagGrid F(s, (int*)flags); MACGrid V(s, (Vec3*)vel); Grid<Real> D(s, (Real*)density);
	  addBuoyancy(F, D, V, Vec3((Real)gx, (Real)gy, (Real)gz), (Real)coefficient, scale != 0); }
	delete s;
  CATCH }

int ref_extrapolate_mac_simple(int sx, int sy, int sz, const int* flags, Real* vel, int distance, const Real* phiObs, int intoObs)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags); MACGrid V(s, (Vec3*)vel);
	  LevelsetGrid* P = phiObs ? new LevelsetGrid(s, (Real*)phiObs) : 0;
	  extrapolateMACSimple(F, V, distance, P, intoObs != 0);
	  delete P; }
	delete s;
  CATCH }

// updateFractions / setObstacleFlags plugin/initplugins.cpp:437-440,:473-475.  The kernel of updateFractions writes neighbour cells on the
// max sides; single-threaded here so that the result is the serial loop's also for boundaryWidth > 0.
int ref_update_fractions(int sx, int sy, int sz, const int* flags, const Real* phiObs, Real* fractions, int boundaryWidth, double fracThreshold)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	const int nt = omp_get_max_threads(); omp_set_num_threads(1);
	try { FlagGrid F(s, (int*)flags); Grid<Real> P(s, (Real*)phiObs); MACGrid Fr(s, (Vec3*)fractions); updateFractions(F, P, Fr, boundaryWidth, (Real)fracThreshold); }
	catch (...) { omp_set_num_threads(nt); delete s; throw; }
	omp_set_num_threads(nt);
	delete s;
  CATCH }
int ref_set_obstacle_flags(int sx, int sy, int sz, int* flags, const Real* phiObs, const Real* fractions, const Real* phiOut, const Real* phiIn, int boundaryWidth)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, flags); Grid<Real> P(s, (Real*)phiObs);
	  MACGrid* Fr = fractions ? new MACGrid(s, (Vec3*)fractions) : 0;
	  Grid<Real>* Po = phiOut ? new Grid<Real>(s, (Real*)phiOut) : 0; Grid<Real>* Pi = phiIn ? new Grid<Real>(s, (Real*)phiIn) : 0;
	  setObstacleFlags(F, P, Fr, Po, Pi, boundaryWidth);
	  delete Fr; delete Po; delete Pi; }
	delete s;
  CATCH }

// ---- FLIP particle <-> grid plugins (plugin/flip.cpp): the caller's arrays are copied into the reference's own particle system
struct RefParts {
	BasicParticleSystem pp; ParticleDataImpl<int>* pt; ParticleDataImpl<Vec3>* pv;
	RefParts(FluidSolver* s, long long np, const Real* pos, const int* pflag, const int* ptype, const Real* pvel) : pp(s), pt(0), pv(0) {
		pp.resizeAll(np);
		for (long long q = 0; q < np; q++) { pp[q].pos = Vec3(pos[3 * q], pos[3 * q + 1], pos[3 * q + 2]); pp[q].flag = pflag[q]; }
		if (ptype) { pt = new ParticleDataImpl<int>(s); pp.registerPdata(pt); pt->resize(np); for (long long q = 0; q < np; q++) (*pt)[q] = ptype[q]; }
		if (pvel) { pv = new ParticleDataImpl<Vec3>(s); pp.registerPdata(pv); pv->resize(np); for (long long q = 0; q < np; q++) (*pv)[q] = Vec3(pvel[3 * q], pvel[3 * q + 1], pvel[3 * q + 2]); }
	}
	~RefParts() { delete pt; delete pv; }
};
int ref_mark_fluid_cells(int sx, int sy, int sz, int* flags, long long np, const Real* pos, const int* pflag, const Real* phiObs, const int* ptype, int exclude)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	const size_t n = (size_t)sx * sy * sz;
	{ RefParts P(s, np, pos, pflag, ptype, 0);
	  FlagGrid F(s); memcpy(&F[0], flags, n * sizeof(int));            // the plugin swaps a copy in when phiObs is given: solver-owned storage
	  Grid<Real>* Po = phiObs ? new Grid<Real>(s, (Real*)phiObs) : 0;
	  markFluidCells(P.pp, F, Po, P.pt, exclude);
	  memcpy(flags, &F[0], n * sizeof(int)); delete Po; }
	delete s;
  CATCH }
int ref_grid_particle_index(int sx, int sy, int sz, long long np, const Real* pos, const int* pflag, int* index, int* indexSys, long long* count)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ RefParts P(s, np, pos, pflag, 0, 0);
	  FlagGrid F(s); Grid<int> I(s, index); ParticleIndexSystem IS(s);
	  gridParticleIndex(P.pp, IS, F, I, 0);
	  *count = IS.size();
	  for (IndexInt q = 0; q < IS.size(); q++) indexSys[q] = (int)IS[q].sourceIndex; }
	delete s;
  CATCH }
int ref_union_particle_levelset(int sx, int sy, int sz, long long np, const Real* pos, const int* index, const int* indexSys, long long count,
                                Real* phi, double radiusFactor, const int* ptype, int exclude)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ std::vector<int> active(np, 0);
	  RefParts P(s, np, pos, active.data(), ptype, 0);
	  FlagGrid F(s); Grid<int> I(s, (int*)index); ParticleIndexSystem IS(s); IS.resizeAll(count);
	  for (long long q = 0; q < count; q++) IS[q].sourceIndex = indexSys[q];
	  LevelsetGrid Ph(s, phi);
	  unionParticleLevelset(P.pp, IS, F, I, Ph, (Real)radiusFactor, P.pt, exclude); }
	delete s;
  CATCH }
int ref_map_parts_to_mac(int sx, int sy, int sz, Real* vel, Real* velOld, long long np, const Real* pos, const int* pflag, const Real* pvel,
                         Real* weight, const int* ptype, int exclude)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ RefParts P(s, np, pos, pflag, ptype, pvel);
	  FlagGrid F(s); MACGrid V(s, (Vec3*)vel), Vo(s, (Vec3*)velOld);
	  Grid<Vec3>* W = weight ? new Grid<Vec3>(s, (Vec3*)weight) : 0;
	  mapPartsToMAC(F, V, Vo, P.pp, *P.pv, W, P.pt, exclude);
	  delete W; }
	delete s;
  CATCH }
int ref_flip_velocity_update(int sx, int sy, int sz, const Real* vel, const Real* velOld, long long np, const Real* pos, const int* pflag, Real* pvel,
                             double flipRatio, const int* ptype, int exclude)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ RefParts P(s, np, pos, pflag, ptype, pvel);
	  FlagGrid F(s); MACGrid V(s, (Vec3*)vel), Vo(s, (Vec3*)velOld);
	  if (flipRatio < 0) mapMACToParts(F, V, P.pp, *P.pv, P.pt, exclude);
	  else flipVelocityUpdate(F, V, Vo, P.pp, *P.pv, (Real)flipRatio, P.pt, exclude);
	  for (long long q = 0; q < np; q++) { pvel[3 * q] = (*P.pv)[q].x; pvel[3 * q + 1] = (*P.pv)[q].y; pvel[3 * q + 2] = (*P.pv)[q].z; } }
	delete s;
  CATCH }

// ParticleSystem::advectInGrid particle.h:512-536 through the unmodified class (mode: IntEuler 0, IntRK2 1, IntRK4 2); pos and pflag are updated
int ref_advect_in_grid(int sx, int sy, int sz, const int* flags, const Real* vel, long long np, Real* pos, int* pflag, double dt, int mode,
                       int deleteInObstacle, int stopInObstacle, int skipNew, const int* ptype, int exclude)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz); s->mDt = (Real)dt;
	{ RefParts P(s, np, pos, pflag, ptype, 0);
	  FlagGrid F(s, (int*)flags); MACGrid V(s, (Vec3*)vel);
	  P.pp.advectInGrid(F, V, mode, deleteInObstacle != 0, stopInObstacle != 0, skipNew != 0, P.pt, exclude);
	  for (long long q = 0; q < np; q++) { pos[3 * q] = P.pp[q].pos.x; pos[3 * q + 1] = P.pp[q].pos.y; pos[3 * q + 2] = P.pp[q].pos.z; pflag[q] = P.pp[q].flag; } }
	delete s;
  CATCH }

int ref_push_out_of_obs(int sx, int sy, int sz, long long np, Real* pos, const int* pflag, const Real* phiObs, double shift, double thresh, const int* ptype, int exclude)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ RefParts P(s, np, pos, pflag, ptype, 0);
	  FlagGrid F(s); Grid<Real> Po(s, (Real*)phiObs);
	  pushOutofObs(P.pp, F, Po, (Real)shift, (Real)thresh, P.pt, exclude);
	  for (long long q = 0; q < np; q++) { pos[3 * q] = P.pp[q].pos.x; pos[3 * q + 1] = P.pp[q].pos.y; pos[3 * q + 2] = P.pp[q].pos.z; } }
	delete s;
  CATCH }
int ref_project_out_of_bnd(int sx, int sy, int sz, long long np, Real* pos, const int* pflag, double bnd, int axis, const int* ptype, int exclude)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ RefParts P(s, np, pos, pflag, ptype, 0);
	  FlagGrid F(s);
	  std::string plane; for (int q = 0; q < 6; q++) if (axis & (1 << q)) plane += "xXyYzZ"[q];
	  P.pp.projectOutOfBnd(F, (Real)bnd, plane, P.pt, exclude);
	  for (long long q = 0; q < np; q++) { pos[3 * q] = P.pp[q].pos.x; pos[3 * q + 1] = P.pp[q].pos.y; pos[3 * q + 2] = P.pp[q].pos.z; } }
	delete s;
  CATCH }

// plugin/ptsplugins.cpp:17-70 and grid.cpp:866-890 through the unmodified plugins
#define PV_OUT for (long long q = 0; q < np; q++) { pvel[3 * q] = (*P.pv)[q].x; pvel[3 * q + 1] = (*P.pv)[q].y; pvel[3 * q + 2] = (*P.pv)[q].z; }
#define POS_OUT for (long long q = 0; q < np; q++) { pos[3 * q] = P.pp[q].pos.x; pos[3 * q + 1] = P.pp[q].pos.y; pos[3 * q + 2] = P.pp[q].pos.z; }
int ref_add_force_pvel(long long np, Real* pvel, double ax, double ay, double az, double dt, const int* ptype, int exclude)
{ TRY
	FluidSolver* s = mkSolver(4, 4, 4);
	{ std::vector<Real> pos(3 * np, 1); std::vector<int> pf(np, 0);
	  RefParts P(s, np, pos.data(), pf.data(), ptype, pvel);
	  addForcePvel(*P.pv, Vec3((Real)ax, (Real)ay, (Real)az), (Real)dt, P.pt, exclude); PV_OUT }
	delete s;
  CATCH }
int ref_update_velocity_from_delta_pos(long long np, const Real* pos, Real* pvel, const Real* xPrev, double dt, const int* ptype, int exclude)
{ TRY
	FluidSolver* s = mkSolver(4, 4, 4);
	{ std::vector<int> pf(np, 0);
	  RefParts P(s, np, pos, pf.data(), ptype, pvel);
	  ParticleDataImpl<Vec3> xp(s); P.pp.registerPdata(&xp); xp.resize(np);
	  for (long long q = 0; q < np; q++) xp[q] = Vec3(xPrev[3 * q], xPrev[3 * q + 1], xPrev[3 * q + 2]);
	  updateVelocityFromDeltaPos(P.pp, *P.pv, xp, (Real)dt, P.pt, exclude); PV_OUT }      // xp leaves the system in its destructor, before P goes
	delete s;
  CATCH }
int ref_euler_step(long long np, Real* pos, const Real* pvel, double dt, const int* ptype, int exclude)
{ TRY
	FluidSolver* s = mkSolver(4, 4, 4); s->mDt = (Real)dt;
	{ std::vector<int> pf(np, 0);
	  RefParts P(s, np, pos, pf.data(), ptype, pvel);
	  eulerStep(P.pp, *P.pv, P.pt, exclude); POS_OUT }
	delete s;
  CATCH }
int ref_set_part_type(int sx, int sy, int sz, long long np, const Real* pos, int* ptype, int mark, int stype, const int* flags, int cflag)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ std::vector<int> pf(np, 0);
	  RefParts P(s, np, pos, pf.data(), ptype, 0);
	  FlagGrid F(s, (int*)flags);
	  setPartType(P.pp, *P.pt, mark, stype, F, cflag);
	  for (long long q = 0; q < np; q++) ptype[q] = (*P.pt)[q]; }
	delete s;
  CATCH }
int ref_mark_isolated_fluid_cell(int sx, int sy, int sz, int* flags, int mark)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, flags); markIsolatedFluidCell(F, mark); }
	delete s;
  CATCH }

int ref_vec_max_abs(int sx, int sy, int sz, const Real* v, double* out)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ MACGrid V(s, (Vec3*)v); *out = (double)V.getMaxAbs(); }
	delete s;
  CATCH }

int ref_extrapolate_mac_from_weight(int sx, int sy, int sz, Real* vel, Real* weight, int distance)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ MACGrid V(s, (Vec3*)vel); Grid<Vec3> W(s, (Vec3*)weight); extrapolateMACFromWeight(V, W, distance); }
	delete s;
  CATCH }

int ref_extrapolate_ls_simple(int sx, int sy, int sz, Real* phi, int distance, int inside)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ Grid<Real> P(s, phi); extrapolateLsSimple(P, distance, inside != 0); }
	delete s;
  CATCH }

int ref_extrapolate_vec3_simple(int sx, int sy, int sz, Real* vel, const Real* phi, int distance, int inside)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ Grid<Vec3> V(s, (Vec3*)vel); Grid<Real> P(s, (Real*)phi); extrapolateVec3Simple(V, P, distance, inside != 0); }
	delete s;
  CATCH }

int ref_update_from_levelset(int sx, int sy, int sz, int* flags, const Real* phi)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, flags); LevelsetGrid P(s, (Real*)phi); F.updateFromLevelset(P); }
	delete s;
  CATCH }

int ref_set_bound(int sx, int sy, int sz, Real* grid, int ncomp, double value, int boundaryWidth)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	if (ncomp == 1) { Grid<Real> G(s, grid); G.setBound((Real)value, boundaryWidth); }
	else { Grid<Vec3> G(s, (Vec3*)grid); G.setBound(Vec3((Real)value), boundaryWidth); }
	delete s;
  CATCH }

// getLaplacian / getCurvature plugin/flip.cpp:710-716 are one-line wrappers of these two kernels (commonkernels.h:75-101)
int ref_get_laplacian(int sx, int sy, int sz, Real* laplace, const Real* grid)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ Grid<Real> L(s, laplace); Grid<Real> G(s, (Real*)grid); LaplaceOp(L, G); }
	delete s;
  CATCH }
int ref_get_curvature(int sx, int sy, int sz, Real* curv, const Real* grid, double h)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ Grid<Real> C(s, curv); Grid<Real> G(s, (Real*)grid); CurvatureOp(C, G, (Real)h); }
	delete s;
  CATCH }

// kind 0: Grid<Real>, 1: MACGrid, 2: FlagGrid, 3: LevelsetGrid-typed Grid<Real>, 4: Grid<Vec3>.  Grid<T>::save / load grid.cpp:113-156.
int ref_grid_file(int sx, int sy, int sz, int kind, void* data, const char* name, int load)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	int ok = 0;
	if (kind == 0) { Grid<Real> G(s, (Real*)data); ok = saveLoad(G, name, load); }
	else if (kind == 1) { MACGrid G(s, (Vec3*)data); ok = saveLoad(G, name, load); }
	else if (kind == 2) { FlagGrid G(s, (int*)data); ok = saveLoad(G, name, load); }
	else if (kind == 3) { LevelsetGrid G(s, (Real*)data); ok = saveLoad(G, name, load); }
	else { Grid<Vec3> G(s, (Vec3*)data); ok = saveLoad(G, name, load); }
	delete s;
	if (!ok) { gLastError = std::string("Grid::") + (load ? "load" : "save") + " returned 0 for " + name; return 1; }
  CATCH }

// kind 0: Grid<Real>, 1: MACGrid, 2: Grid<Vec3>.  The plugin swaps its result in, which the reference forbids for external data: work on solver-owned copies.
int ref_advect_semi_lagrange(int sx, int sy, int sz, const int* flags, const Real* vel, Real* grid, int kind,
	int order, double strength, int orderSpace, int clampMode, int orderTrace, double dt)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz); s->mDt = (Real)dt;
	const size_t n = (size_t)sx * sy * sz;
	{ FlagGrid F(s, (int*)flags); MACGrid V(s, (Vec3*)vel);
	  if (kind == 0) { Grid<Real> G(s); memcpy(&G[0], grid, n * sizeof(Real));
	    advectSemiLagrange(&F, &V, &G, order, (Real)strength, orderSpace, false, -1, clampMode, orderTrace);
	    memcpy(grid, &G[0], n * sizeof(Real)); }
	  else if (kind == 2) { Grid<Vec3> G(s); memcpy(&G[0], grid, n * sizeof(Vec3));
	    advectSemiLagrange(&F, &V, &G, order, (Real)strength, orderSpace, false, -1, clampMode, orderTrace);
	    memcpy(grid, &G[0], n * sizeof(Vec3)); }
	  else { MACGrid G(s); memcpy(&G[0], grid, n * sizeof(Vec3));
	    MACGrid Vc(s); memcpy(&Vc[0], vel, n * sizeof(Vec3));      // self-advection passes the same grid as vel and grid: keep that aliasing out of the harness
	    advectSemiLagrange(&F, &Vc, &G, order, (Real)strength, orderSpace, false, -1, clampMode, orderTrace);
	    memcpy(grid, &G[0], n * sizeof(Vec3)); } }
	delete s;
  CATCH }

int ref_pd_fluid_guiding(int sx, int sy, int sz, const int* flags, Real* vel, const Real* velT, Real* pressure, const Real* weight,
	int blurRadius, double theta, double tau, double sigma, double epsRel, double epsAbs, int maxIters,
	double cgMaxIterFac, double cgAccuracy, int preconditioner, int zeroPressureFixing, int* iterations)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	std::string log;
	{ FlagGrid F(s, (int*)flags); MACGrid V(s, (Vec3*)vel); MACGrid VT(s, (Vec3*)velT); Grid<Real> P(s, pressure); Grid<Real> W(s, (Real*)weight);
	  int oldLevel = gDebugLevel; gDebugLevel = 1;
	  try {
		CoutCapture cap;
		try {
			PD_fluid_guiding(V, VT, P, F, W, blurRadius, (Real)theta, (Real)tau, (Real)sigma, (Real)epsRel, (Real)epsAbs, maxIters, 0, 0, 0, 0, (Real)1e-04,
				(Real)cgMaxIterFac, (Real)cgAccuracy, preconditioner, zeroPressureFixing != 0, 0, (Real)0.);
		} catch (...) { log = cap.ss.str(); throw; }
		log = cap.ss.str();
	  } catch (...) { gDebugLevel = oldLevel; releaseBlurPrecomp(); releaseMG(s); throw; }
	  gDebugLevel = oldLevel; }
	releaseBlurPrecomp();              // the plugin keeps its blur kernel in globals (fluidguiding.cpp:22-24)
	releaseMG(s);
	size_t p = log.rfind("PD_fluid_guiding iterations:");      // :350
	if (iterations) *iterations = (p != std::string::npos) ? atoi(log.c_str() + p + 28) : -1;
	delete s;
  CATCH }

// ut / utm1 are swapped by the plugin (not allowed for external data): solver-owned copies, results copied back
// VICintegration plugin/vortexplugins.cpp:195-300 on a vortex sheet given as triangles (tri: [ntri][3 corners][3], vort: [ntri][3]); returns the
// vorticity grid the Peskin kernel leaves (:203-250) and the velocity of the three Poisson solves (:253-299).  velIsMac: vel is a MACGrid
// (GetShiftedComponent) or a centred Grid<Vec3> (GetComponent).  iters: the three GridCg iteration counts, parsed from the debMsg line :295.
int ref_vic_integration(int sx, int sy, int sz, const int* flags, int ntri, const Real* tri, const Real* vort, double sigma, Real* vel, int velIsMac,
	Real* vorticity, double cgMaxIterFac, double cgAccuracy, double scale, int precondition, int* iters)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	const size_t n = (size_t)sx * sy * sz;
	std::string log;
	{ FlagGrid F(s, (int*)flags); Grid<Vec3> W(s);
	  Grid<Vec3>* V = velIsMac ? new MACGrid(s) : new Grid<Vec3>(s);
	  memcpy(&(*V)[0], vel, 3 * n * sizeof(Real));
	  struct Sheet : VortexSheetMesh { Sheet(FluidSolver* p) : VortexSheetMesh(p) {} void sync() { rebuildChannels(); } } mesh(s);   // Mesh::load does the same after reading triangles
	  for (int t = 0; t < ntri; t++) {
		int c[3];
		for (int q = 0; q < 3; q++) c[q] = mesh.addNode(Node(Vec3(tri[9 * t + 3 * q], tri[9 * t + 3 * q + 1], tri[9 * t + 3 * q + 2])));
		mesh.addTri(Triangle(c[0], c[1], c[2]));
	  }
	  mesh.sync();
	  for (int t = 0; t < ntri; t++) mesh.sheet(t).vorticity = Vec3(vort[3 * t], vort[3 * t + 1], vort[3 * t + 2]);
	  const int dl = gDebugLevel; gDebugLevel = 1;
	  { CoutCapture cap;
	    VICintegration(mesh, (Real)sigma, *V, F, &W, (Real)cgMaxIterFac, (Real)cgAccuracy, (Real)scale, precondition);
	    log = cap.ss.str(); }
	  gDebugLevel = dl;
	  memcpy(vel, &(*V)[0], 3 * n * sizeof(Real)); memcpy(vorticity, &W[0], 3 * n * sizeof(Real));
	  delete V; }
	delete s;
	if (iters) {
		size_t pos = 0;
		for (int c = 0; c < 3; c++) {
			iters[c] = -1;
			pos = log.find("VICintegration CG iterations:", pos);
			if (pos == std::string::npos) break;
			pos += strlen("VICintegration CG iterations:");
			iters[c] = atoi(log.c_str() + pos);
		}
	}
  CATCH }

int ref_cg_solve_we(int sx, int sy, int sz, const int* flags, Real* ut, Real* utm1, Real* out, int crankNic, double cSqr, double cgMaxIterFac, double cgAccuracy, double dt)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz); s->mDt = (Real)dt;
	const size_t n = (size_t)sx * sy * sz;
	{ FlagGrid F(s, (int*)flags); Grid<Real> U(s), Um(s), O(s);
	  memcpy(&U[0], ut, n * sizeof(Real)); memcpy(&Um[0], utm1, n * sizeof(Real)); memcpy(&O[0], out, n * sizeof(Real));
	  cgSolveWE(F, U, Um, O, crankNic != 0, (Real)cSqr, (Real)cgMaxIterFac, (Real)cgAccuracy);
	  memcpy(ut, &U[0], n * sizeof(Real)); memcpy(utm1, &Um[0], n * sizeof(Real)); memcpy(out, &O[0], n * sizeof(Real)); }
	delete s;
  CATCH }

int ref_compute_rhs(int sx, int sy, int sz, const int* flags, const Real* vel, Real* rhs,
	const Real* phi, const Real* perCellCorr, const Real* fractions, const Real* obvel, const Real* curv,
	double gfClamp, double surfTens, int enforceCompatibility, double* sum, int* cnt)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags); MACGrid V(s, (Vec3*)vel); Grid<Real> R(s, rhs);
	  Grid<Real>* P = phi ? new Grid<Real>(s, (Real*)phi) : 0;
	  Grid<Real>* C = perCellCorr ? new Grid<Real>(s, (Real*)perCellCorr) : 0;
	  MACGrid* Fr = fractions ? new MACGrid(s, (Vec3*)fractions) : 0;
	  MACGrid* Ov = obvel ? new MACGrid(s, (Vec3*)obvel) : 0;
	  Grid<Real>* Cu = curv ? new Grid<Real>(s, (Real*)curv) : 0;
	  MakeRhs k(F, R, V, C, Fr, Ov, P, Cu, (Real)surfTens, (Real)gfClamp);
	  if (enforceCompatibility) R += (Real)(-k.sum / (Real)k.cnt);
	  if (sum) *sum = k.sum; if (cnt) *cnt = k.cnt;
	  delete P; delete C; delete Fr; delete Ov; delete Cu; }
	delete s;
  CATCH }

// MakeLaplaceMatrix (+ ApplyGhostFluidDiagonal when phi != NULL); A* must be zero-filled by the caller
int ref_make_matrix(int sx, int sy, int sz, const int* flags, const Real* fractions, const Real* phi, double gfClamp,
	Real* A0, Real* Ai, Real* Aj, Real* Ak)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags); Grid<Real> a0(s, A0), ai(s, Ai), aj(s, Aj), ak(s, Ak);
	  MACGrid* Fr = fractions ? new MACGrid(s, (Vec3*)fractions) : 0;
	  MakeLaplaceMatrix(F, a0, ai, aj, ak, Fr);
	  if (phi) { Grid<Real> P(s, (Real*)phi); ApplyGhostFluidDiagonal(a0, F, P, (Real)gfClamp); }
	  delete Fr; }
	delete s;
  CATCH }

int ref_fix_pressure(int sx, int sy, int sz, long long idx, double value, Real* rhs, Real* A0, Real* Ai, Real* Aj, Real* Ak)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ Grid<Real> r(s, rhs), a0(s, A0), ai(s, Ai), aj(s, Aj), ak(s, Ak);
	  fixPressure((int)idx, (Real)value, r, a0, ai, aj, ak); }
	delete s;
  CATCH }

int ref_apply_matrix(int sx, int sy, int sz, const int* flags, Real* dst, const Real* src,
	const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags); Grid<Real> d(s, dst), sr(s, (Real*)src), a0(s, (Real*)A0), ai(s, (Real*)Ai), aj(s, (Real*)Aj), ak(s, (Real*)Ak);
	  if (sz > 1) ApplyMatrix(F, d, sr, a0, ai, aj, ak); else ApplyMatrix2D(F, d, sr, a0, ai, aj, ak); }
	delete s;
  CATCH }

int ref_mic_init(int sx, int sy, int sz, const int* flags, Real* precond, const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags); Grid<Real> p(s, precond), a0(s, (Real*)A0), ai(s, (Real*)Ai), aj(s, (Real*)Aj), ak(s, (Real*)Ak);
	  InitPreconditionModifiedIncompCholesky2(F, p, a0, ai, aj, ak); }
	delete s;
  CATCH }

int ref_mic_apply(int sx, int sy, int sz, const int* flags, Real* dst, const Real* src, const Real* precond,
	const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags); Grid<Real> d(s, dst), sr(s, (Real*)src), p(s, (Real*)precond), a0(s, (Real*)A0), ai(s, (Real*)Ai), aj(s, (Real*)Aj), ak(s, (Real*)Ak);
	  ApplyPreconditionModifiedIncompCholesky2(d, sr, F, p, a0, ai, aj, ak); }
	delete s;
  CATCH }

// IC(0) "a la Wavelet Turbulence" conjugategrad.cpp:26-63,:109-132 (PC_ICP, the preconditioner of the VIC Poisson solve)
int ref_ic_init(int sx, int sy, int sz, const int* flags, Real* P0, Real* Pi, Real* Pj, Real* Pk, const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags); Grid<Real> p0(s, P0), pi(s, Pi), pj(s, Pj), pk(s, Pk), a0(s, (Real*)A0), ai(s, (Real*)Ai), aj(s, (Real*)Aj), ak(s, (Real*)Ak);
	  InitPreconditionIncompCholesky(F, p0, pi, pj, pk, a0, ai, aj, ak); }
	delete s;
  CATCH }

int ref_ic_apply(int sx, int sy, int sz, const int* flags, Real* dst, const Real* src, const Real* P0, const Real* Pi, const Real* Pj, const Real* Pk)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags); Grid<Real> d(s, dst), sr(s, (Real*)src), p0(s, (Real*)P0), pi(s, (Real*)Pi), pj(s, (Real*)Pj), pk(s, (Real*)Pk);
	  ApplyPreconditionIncompCholesky(d, sr, F, p0, pi, pj, pk, p0, pi, pj, pk); }      // the org* arguments are unused (:109-132)
	delete s;
  CATCH }

// GridCg driven directly (the only unpatched route to PcNone in 3-D).
// pc: 0 none, 1 mICP, 2 MG (fresh GridMg), 3 ICP.  x is overwritten; returns iterations / resNorm.
int ref_cg_solve(int sx, int sy, int sz, const int* flags, const Real* rhs, Real* x,
	const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak,
	int pc, double accuracy, int useL2, int maxIter, int* iterations, double* resNorm)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags);
	  Grid<Real> X(s, x), B(s, (Real*)rhs), a0(s, (Real*)A0), ai(s, (Real*)Ai), aj(s, (Real*)Aj), ak(s, (Real*)Ak);
	  Grid<Real> residual(s), search(s), tmp(s);
	  GridCgInterface* gcg;
	  if (sz > 1) gcg = new GridCg<ApplyMatrix>(X, B, residual, search, F, tmp, &a0, &ai, &aj, &ak);
	  else        gcg = new GridCg<ApplyMatrix2D>(X, B, residual, search, F, tmp, &a0, &ai, &aj, &ak);
	  gcg->setAccuracy((Real)accuracy);
	  gcg->setUseL2Norm(useL2 != 0);
	  Grid<Real>* pca0 = 0, *pca1 = 0, *pca2 = 0, *pca3 = 0; GridMg* mg = 0;
	  if (pc == 1) {
		pca0 = new Grid<Real>(s); pca1 = new Grid<Real>(s); pca2 = new Grid<Real>(s); pca3 = new Grid<Real>(s);
		gcg->setICPreconditioner(GridCgInterface::PC_mICP, pca0, pca1, pca2, pca3);
	  } else if (pc == 3) {
		pca0 = new Grid<Real>(s); pca1 = new Grid<Real>(s); pca2 = new Grid<Real>(s); pca3 = new Grid<Real>(s);
		gcg->setICPreconditioner(GridCgInterface::PC_ICP, pca0, pca1, pca2, pca3);
	  } else if (pc == 2) {
		mg = new GridMg(Vec3i(sx, sy, sz));
		gcg->setMGPreconditioner(GridCgInterface::PC_MGP, mg);
	  }
	  try {
		for (int iter = 0; iter < maxIter; iter++) if (!gcg->iterate()) iter = maxIter;
	  } catch (...) { delete gcg; delete pca0; delete pca1; delete pca2; delete pca3; delete mg; throw; }
	  if (iterations) *iterations = (int)gcg->getIterations();
	  if (resNorm) *resNorm = gcg->getResNorm();
	  delete gcg; delete pca0; delete pca1; delete pca2; delete pca3; delete mg; }
	delete s;
  CATCH }

int ref_correct_velocity(int sx, int sy, int sz, const int* flags, Real* vel, const Real* pressure,
	const Real* phi, const Real* curv, double gfClamp, double surfTens)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags); MACGrid V(s, (Vec3*)vel); Grid<Real> P(s, (Real*)pressure);
	  Grid<Real>* Ph = phi ? new Grid<Real>(s, (Real*)phi) : 0;
	  Grid<Real>* Cu = curv ? new Grid<Real>(s, (Real*)curv) : 0;
	  correctVelocity(V, P, F, 1e-3, Ph, 0, 0, (Real)gfClamp, 1.5, true, PcMIC, false, false, false, Cu, (Real)surfTens);
	  delete Ph; delete Cu; }
	delete s;
  CATCH }

// the reference plugin itself; `solver_key` lets PcMGStatic reuse its hierarchy across calls
// (gMapMG is keyed by FluidSolver*, pressure.cpp:250): pass the same key for the same scene.
static std::map<long long, FluidSolver*> gSolvers;

int ref_release_solver(long long key)
{ TRY
	auto it = gSolvers.find(key);
	if (it != gSolvers.end()) { releaseMG(it->second); delete it->second; gSolvers.erase(it); }
  CATCH }

int ref_solve_pressure(long long solver_key, int sx, int sy, int sz, const int* flags, Real* vel, Real* pressure,
	const Real* phi, const Real* perCellCorr, const Real* fractions, const Real* obvel, const Real* curv, Real* retRhs,
	double cgAccuracy, double gfClamp, double cgMaxIterFac, int precondition, int preconditioner,
	int enforceCompatibility, int useL2Norm, int zeroPressureFixing, double surfTens,
	int* iterations, double* resNorm)
{ TRY
	FluidSolver* s;
	if (solver_key) { auto it = gSolvers.find(solver_key); if (it == gSolvers.end()) { s = mkSolver(sx, sy, sz); gSolvers[solver_key] = s; } else s = it->second; }
	else s = mkSolver(sx, sy, sz);
	std::string log;
	{ FlagGrid F(s, (int*)flags); MACGrid V(s, (Vec3*)vel); Grid<Real> P(s, pressure);
	  Grid<Real>* Ph = phi ? new Grid<Real>(s, (Real*)phi) : 0;
	  Grid<Real>* C = perCellCorr ? new Grid<Real>(s, (Real*)perCellCorr) : 0;
	  MACGrid* Fr = fractions ? new MACGrid(s, (Vec3*)fractions) : 0;
	  MACGrid* Ov = obvel ? new MACGrid(s, (Vec3*)obvel) : 0;
	  Grid<Real>* Cu = curv ? new Grid<Real>(s, (Real*)curv) : 0;
	  Grid<Real>* RR = retRhs ? new Grid<Real>(s, retRhs) : 0;
	  int oldLevel = gDebugLevel; gDebugLevel = 2;
	  try {
		CoutCapture cap;
		try {
			solvePressure(V, P, F, (Real)cgAccuracy, Ph, C, Fr, Ov, (Real)gfClamp, (Real)cgMaxIterFac, precondition != 0, preconditioner,
				enforceCompatibility != 0, useL2Norm != 0, zeroPressureFixing != 0, Cu, (Real)surfTens, RR);
		} catch (...) { log = cap.ss.str(); throw; }
		log = cap.ss.str();
	  } catch (...) { gDebugLevel = oldLevel; delete Ph; delete C; delete Fr; delete Ov; delete Cu; delete RR; if (!solver_key) { releaseMG(s); } throw; }
	  gDebugLevel = oldLevel;
	  delete Ph; delete C; delete Fr; delete Ov; delete Cu; delete RR; }
	// parse "FluidSolver::solvePressure done. Iterations:<n>, residual:<r>"  (pressure.cpp:440)
	size_t p = log.rfind("Iterations:");
	if (p != std::string::npos) {
		int it = 0; double rn = 0;
		sscanf(log.c_str() + p, "Iterations:%d, residual:%lf", &it, &rn);
		if (iterations) *iterations = it; if (resNorm) *resNorm = rn;
	}
	if (!solver_key) { releaseMG(s); delete s; }
  CATCH }

// cgSolveDiffusion conjugategrad.cpp:350-423 (another GridCg caller).  ncomp 1: Grid<Real>, 3: Grid<Vec3>
int ref_cg_solve_diffusion(int sx, int sy, int sz, const int* flags, Real* data, int ncomp, double alpha, double cgMaxIterFac, double cgAccuracy)
{ TRY
	FluidSolver* s = mkSolver(sx, sy, sz);
	{ FlagGrid F(s, (int*)flags);
	  if (ncomp == 1) { Grid<Real> G(s, data); cgSolveDiffusion(F, G, (Real)alpha, (Real)cgMaxIterFac, (Real)cgAccuracy); }
	  else            { Grid<Vec3> G(s, (Vec3*)data); cgSolveDiffusion(F, G, (Real)alpha, (Real)cgMaxIterFac, (Real)cgAccuracy); } }
	delete s;
  CATCH }

// ---------------- GridMg probes (hierarchy + V-cycle) ----------------
static GridMg* gMg = 0; static FluidSolver* gMgSolver = 0;

int ref_mg_create(int sx, int sy, int sz)
{ TRY
	delete gMg; delete gMgSolver;
	gMgSolver = mkSolver(sx, sy, sz);
	gMg = new GridMg(Vec3i(sx, sy, sz));
  CATCH }
int ref_mg_destroy() { delete gMg; gMg = 0; delete gMgSolver; gMgSolver = 0; return 0; }
int ref_mg_set_a(const Real* A0, const Real* Ai, const Real* Aj, const Real* Ak)
{ TRY
	FluidSolver* s = gMgSolver;
	Grid<Real> a0(s, (Real*)A0), ai(s, (Real*)Ai), aj(s, (Real*)Aj), ak(s, (Real*)Ak);
	gMg->setA(&a0, &ai, &aj, &ak);
  CATCH }
int ref_mg_num_levels() { return gMg ? (int)gMg->mA.size() : 0; }
int ref_mg_level_size(int l, int* out3) { out3[0] = gMg->mSize[l].x; out3[1] = gMg->mSize[l].y; out3[2] = gMg->mSize[l].z; return 0; }
int ref_mg_stencil_size(int l) { return l == 0 ? gMg->mStencilSize0 : gMg->mStencilSize; }
int ref_mg_get_type(int l, signed char* out) { for (size_t i = 0; i < gMg->mType[l].size(); i++) out[i] = (signed char)gMg->mType[l][i]; return 0; }
int ref_mg_get_a(int l, Real* out) { memcpy(out, gMg->mA[l].data(), gMg->mA[l].size() * sizeof(Real)); return 0; }
int ref_mg_get_x(int l, Real* out) { memcpy(out, gMg->mx[l].data(), gMg->mx[l].size() * sizeof(Real)); return 0; }
int ref_mg_get_b(int l, Real* out) { memcpy(out, gMg->mb[l].data(), gMg->mb[l].size() * sizeof(Real)); return 0; }
int ref_mg_get_r(int l, Real* out) { memcpy(out, gMg->mr[l].data(), gMg->mr[l].size() * sizeof(Real)); return 0; }
// one V-cycle as the preconditioner applies it (conjugategrad.cpp:100-106, :162-167)
int ref_mg_vcycle(const Real* rhs, Real* dst, double coarsestAccuracy, int pre, int post)
{ TRY
	FluidSolver* s = gMgSolver;
	Grid<Real> b(s, (Real*)rhs), d(s, dst);
	gMg->setCoarsestLevelAccuracy((Real)coarsestAccuracy);
	gMg->setSmoothing(pre, post);
	gMg->setRhs(b);
	gMg->doVCycle(d);
  CATCH }

} // extern "C"

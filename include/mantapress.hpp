// mantapress.hpp -- C++ host mirror of the reference's types and plugins for the pressure-projection path, header-only, over the C-ABI
// of mantapress.h.  The reference is compiled C++ driven from Python; this is the C++ side of that pair: the same class and function
// names, argument order, defaults and error behaviour as
//     FluidSolver                         fluidsolver.h:27-91
//     Grid<T>, MACGrid, FlagGrid          grid.h:92-365           (LevelsetGrid levelset.h:25)
//     enum Preconditioner, releaseMG, computePressureRhs, solvePressureSystem, correctVelocity, solvePressure
//                                         plugin/pressure.cpp:27,:252,:277-292,:312-326,:455-468,:480-495
//     setWallBcs, addGravity, addGravityNoScale, addBuoyancy          plugin/extforces.cpp:61-90,:307-316
//     advectSemiLagrange                                              plugin/advection.cpp:442-461
//     extrapolateMACSimple, extrapolateLsSimple, extrapolateVec3Simple fastmarch.cpp:337-375,:470-542
//     BasicParticleSystem, ParticleDataImpl<T>, ParticleIndexSystem (particle.h) and the FLIP plugins of plugin/flip.cpp (markFluidCells ... pushOutofObs)
//     getLaplacian, getCurvature                                      plugin/flip.cpp:710-716
//     updateFractions, setObstacleFlags                               plugin/initplugins.cpp:437-440,:473-475
//     cgSolveDiffusion, cgSolveWE, vicPoisson                         conjugategrad.cpp:350, plugin/waves.cpp:86, plugin/vortexplugins.cpp:253
// so that C++ callers of these plugins (e.g. plugin/fluidguiding.cpp:276-335) and tests written against the reference's headers compile
// against this header unchanged.  Every grid owns a host array in the reference layout (grid.h:70) AND an mp_grid in HBM; two dirty
// bits keep them coherent lazily, so a sequence of plugins never leaves the device.
// Errors: a failing C-ABI call throws Manta::Error with the library's message, as errMsg / assertMsg do in the reference
// (general.h:42-57).  There is no CPU fallback: constructing a FluidSolver without a CUDA device throws.
//
// Precision is a compile-time choice like the reference's: default float ("fp1"), -DDOUBLEPRECISION=1 for the double build.
#pragma once
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>
#include "mantapress.h"

namespace Manta {

#if DOUBLEPRECISION
typedef double Real;
#else
typedef float Real;
#endif

class Error : public std::runtime_error {           // general.h:42-57
public:
	int status;
	Error(int status_, const std::string& s) : std::runtime_error(s), status(status_) {}
};
inline void mpCheck(int status) { if (status != MP_OK) throw Error(status, mp_last_error()); }

struct Vec3i { int x, y, z; Vec3i(int x_ = 0, int y_ = 0, int z_ = 0) : x(x_), y(y_), z(z_) {} int max() const { return std::max(x, std::max(y, z)); } };
struct Vec3 {                                        // vectorbase.h:199-213, AoS {x,y,z}
	Real x, y, z;
	Vec3(Real v = 0) : x(v), y(v), z(v) {}
	Vec3(Real x_, Real y_, Real z_) : x(x_), y(y_), z(z_) {}
	Real& operator[](int c) { return (&x)[c]; }
	const Real& operator[](int c) const { return (&x)[c]; }
};

enum Preconditioner { PcNone = 0, PcMIC = 1, PcMGDynamic = 2, PcMGStatic = 3 };      // plugin/pressure.cpp:27

class FluidSolver {                                  // fluidsolver.h:27-91
public:
	FluidSolver(Vec3i gridSize, int dim = 3, int device = 0) : mDt(1.0), mGridSize(gridSize), mDim(dim), mCtx(nullptr) {
		if (dim != 2 && dim != 3) throw Error(MP_ERR_INVALID, "Only 2D and 3D solvers allowed.");
		if (dim == 2 && gridSize.z != 1) throw Error(MP_ERR_INVALID, "Trying to create 2D solver with size.z != 1");
		mpCheck(mp_context_create(device, &mCtx));
	}
	~FluidSolver() { if (mCtx) mp_context_destroy(mCtx); }
	FluidSolver(const FluidSolver&) = delete;
	FluidSolver& operator=(const FluidSolver&) = delete;
	Vec3i getGridSize() const { return mGridSize; }
	bool is2D() const { return mDim == 2; }
	bool is3D() const { return mDim == 3; }
	Real getDt() const { return mDt; }
	Real getDx() const { return (Real)(1.0 / mGridSize.max()); }
	mp_context* ctx() const { return mCtx; }
	void synchronize() const { mpCheck(mp_context_synchronize(mCtx)); }
	long long kernelLaunches() const { long long n = 0; mpCheck(mp_context_kernel_launches(mCtx, &n)); return n; }
	// MIC(0) of PcMIC / PC_mICP in block red-black ordering (the reformulated preconditioner, no reference counterpart): mode 1 on, 0 the reference's ordering
	void setMicOrdering(int mode = 1, int tileY = 0, int tileZ = 0) { mpCheck(mp_set_mic_ordering(mCtx, mode, tileY, tileZ)); }
	Real mDt;                                        // timestep, public like the reference's Python-exposed member
private:
	Vec3i mGridSize; int mDim; mp_context* mCtx;
};

template <class T> struct GridKind;
template <> struct GridKind<Real> { static const int kind = MP_GRID_REAL; };
template <> struct GridKind<int>  { static const int kind = MP_GRID_FLAGS; };
template <> struct GridKind<Vec3> { static const int kind = MP_GRID_MAC; };

class GridBase {                                     // grid.h:27-89
public:
	enum GridType { TypeNone = 0, TypeReal = 1, TypeInt = 2, TypeVec3 = 4, TypeMAC = 8, TypeLevelset = 16, TypeFlags = 32 };
	GridBase(FluidSolver* parent) : mParent(parent), mType(TypeNone), mSize(parent->getGridSize()), m3D(parent->is3D()), mDev(nullptr), mHostDirty(false), mDevDirty(false) {
		mStrideZ = m3D ? (long long)mSize.x * mSize.y : 0;
	}
	virtual ~GridBase() { if (mDev) mp_grid_destroy(mDev); }
	GridBase(const GridBase&) = delete;
	GridBase& operator=(const GridBase&) = delete;
	FluidSolver* getParent() const { return mParent; }
	int getSizeX() const { return mSize.x; }
	int getSizeY() const { return mSize.y; }
	int getSizeZ() const { return mSize.z; }
	Vec3i getSize() const { return mSize; }
	bool is3D() const { return m3D; }
	GridType getType() const { return mType; }
	long long getStrideX() const { return 1; }
	long long getStrideY() const { return mSize.x; }
	long long getStrideZ() const { return mStrideZ; }
	long long index(int i, int j, int k) const { return (long long)i + (long long)mSize.x * j + mStrideZ * k; }
	bool isInBounds(const Vec3i& p, int bnd = 0) const {
		bool ret = p.x >= bnd && p.y >= bnd && p.x < mSize.x - bnd && p.y < mSize.y - bnd;
		if (m3D) ret &= p.z >= bnd && p.z < mSize.z - bnd; else ret &= p.z == 0;
		return ret;
	}
	// ---- device mirror ----
	void markDeviceWritten() { mDevDirty = true; mHostDirty = false; }
	void markHostWritten() { mHostDirty = true; mDevDirty = false; }
protected:
	FluidSolver* mParent; GridType mType; Vec3i mSize; bool m3D; long long mStrideZ;
	mutable mp_grid* mDev; mutable bool mHostDirty, mDevDirty;
};

template <class T>
class Grid : public GridBase {                       // grid.h:92-235; zeroed on construction (grid.cpp:47-59)
public:
	Grid(FluidSolver* parent) : GridBase(parent), mData((size_t)parent->getGridSize().x * parent->getGridSize().y * parent->getGridSize().z) {
		mType = std::is_same<T, Real>::value ? TypeReal : (std::is_same<T, int>::value ? TypeInt : TypeVec3);      // the vector value-initialises: a zeroed grid
		mpCheck(mp_grid_create(parent->ctx(), GridKind<T>::kind, (int)sizeof(Real), mSize.x, mSize.y, mSize.z, &mDev));
	}
	// host access: brings the host copy up to date first; writing through the non-const accessors marks it the newer one
	const T& operator()(int i, int j, int k) const { syncToHost(); return mData[index(i, j, k)]; }
	T& operator()(int i, int j, int k) { syncToHost(); mHostDirty = true; return mData[index(i, j, k)]; }
	const T& operator[](long long idx) const { syncToHost(); return mData[idx]; }
	T& operator[](long long idx) { syncToHost(); mHostDirty = true; return mData[idx]; }
	const T* hostData() const { syncToHost(); return mData.data(); }
	T* hostDataForWriting() { syncToHost(); mHostDirty = true; return mData.data(); }
	// device access: uploads only if the host copy is the newer one
	mp_grid* dev() const {
		if (mHostDirty) { mpCheck(mp_grid_upload(mDev, mData.data())); mHostDirty = false; }
		return mDev;
	}
	void syncToHost() const { if (mDevDirty) { mpCheck(mp_grid_download(mDev, (void*)mData.data())); mDevDirty = false; } }
	void clear() { mpCheck(mp_grid_clear(mDev)); std::fill(mData.begin(), mData.end(), T()); mHostDirty = mDevDirty = false; }   // grid.cpp:93-96
	void copyFrom(const Grid<T>& a) { mpCheck(mp_grid_copy_from(dev(), a.dev())); markDeviceWritten(); }                                      // grid.cpp:205-210
	// element-wise arithmetic on the device, grid.cpp:258-284
	void setConst(T s) { arith(MP_OP_SET_CONST, nullptr, s); }
	void addConst(T s) { arith(MP_OP_ADD_CONST, nullptr, s); }
	void multConst(T s) { arith(MP_OP_MULT_CONST, nullptr, s); }
	void add(const Grid<T>& a) { arith(MP_OP_ADD, &a, T()); }
	void sub(const Grid<T>& a) { arith(MP_OP_SUB, &a, T()); }
	void mult(const Grid<T>& a) { arith(MP_OP_MULT, &a, T()); }
	void addScaled(const Grid<T>& a, const T& factor) { arith(MP_OP_ADD_SCALED, &a, factor); }
	void clamp(Real min, Real max) { mpCheck(mp_grid_arith(mParent->ctx(), dev(), MP_OP_CLAMP, nullptr, min, max, 0)); markDeviceWritten(); }
	void stomp(const T& threshold) { arith(MP_OP_STOMP, nullptr, threshold); }
	Grid<T>& safeDivide(const Grid<T>& a) { arith(MP_OP_SAFE_DIVIDE, &a, T()); return *this; }
	void setBound(T value, int boundaryWidth = 1);                                                                                             // grid.cpp:591-593
	Real getMaxAbs() const;                                                                                                                    // grid.cpp:319-323 (Grid<Real>)
protected:
	static void xyz(const Vec3& v, double (&c)[3]) { c[0] = v.x; c[1] = v.y; c[2] = v.z; }
	template <class S> static void xyz(const S& v, double (&c)[3]) { c[0] = c[1] = c[2] = (double)v; }
	void arith(int op, const Grid<T>* other, const T& value) {
		double c[3]; xyz(value, c);
		mpCheck(mp_grid_arith(mParent->ctx(), dev(), op, other ? other->dev() : nullptr, c[0], c[1], c[2])); markDeviceWritten();
	}
	mutable std::vector<T> mData;
};
template <> inline void Grid<Real>::setBound(Real value, int w) { mpCheck(mp_grid_set_bound(mParent->ctx(), dev(), value, value, value, w)); markDeviceWritten(); }
template <> inline void Grid<int>::setBound(int value, int w) { mpCheck(mp_grid_set_bound(mParent->ctx(), dev(), value, value, value, w)); markDeviceWritten(); }
template <> inline void Grid<Vec3>::setBound(Vec3 value, int w) { mpCheck(mp_grid_set_bound(mParent->ctx(), dev(), value.x, value.y, value.z, w)); markDeviceWritten(); }
template <> inline Real Grid<Real>::getMaxAbs() const { double v = 0; mpCheck(mp_grid_max_abs(mParent->ctx(), dev(), &v)); return (Real)v; }
template <> inline Real Grid<Vec3>::getMaxAbs() const { double v = 0; mpCheck(mp_grid_max_abs(mParent->ctx(), dev(), &v)); return (Real)v; }      // grid.cpp:330-332

class LevelsetGrid : public Grid<Real> {             // levelset.h:25-60
public:
	LevelsetGrid(FluidSolver* parent) : Grid<Real>(parent) { mType = (GridType)(TypeLevelset | TypeReal); }
	static Real invalidTimeValue() { return -1000; }  // levelset.cpp:103 -> fastmarch.h:134
};

class MACGrid : public Grid<Vec3> {                  // grid.h:243-281
public:
	MACGrid(FluidSolver* parent) : Grid<Vec3>(parent) { mType = (GridType)(TypeMAC | TypeVec3); }
};

class FlagGrid : public Grid<int> {                  // grid.h:284-365
public:
	enum CellType { TypeNone = 0, TypeFluid = 1, TypeObstacle = 2, TypeEmpty = 4, TypeInflow = 8, TypeOutflow = 16, TypeOpen = 32, TypeStick = 64 };
	FlagGrid(FluidSolver* parent) : Grid<int>(parent) { mType = (GridType)(TypeFlags | TypeInt); }
	int get(int i, int j, int k) const { return (*this)(i, j, k); }
	bool isFluid(int i, int j, int k) const { return get(i, j, k) & TypeFluid; }
	bool isObstacle(int i, int j, int k) const { return get(i, j, k) & TypeObstacle; }
	bool isEmpty(int i, int j, int k) const { return get(i, j, k) & TypeEmpty; }
	bool isOutflow(int i, int j, int k) const { return get(i, j, k) & TypeOutflow; }
	// grid.cpp:732-842 with the default strings: everything Empty, the six boundary slabs of width boundaryWidth+1 Obstacle
	void initDomain(int boundaryWidth = 0) {
		int* f = hostDataForWriting();
		const int w = boundaryWidth;
		for (int k = 0; k < mSize.z; k++) for (int j = 0; j < mSize.y; j++) for (int i = 0; i < mSize.x; i++) {
			const bool bnd = i <= w || i >= mSize.x - 1 - w || j <= w || j >= mSize.y - 1 - w || (m3D && (k <= w || k >= mSize.z - 1 - w));
			f[index(i, j, k)] = bnd ? TypeObstacle : TypeEmpty;
		}
	}
	void fillGrid(int type = TypeFluid) {              // grid.cpp:856-861
		int* f = hostDataForWriting();
		for (size_t q = 0; q < mData.size(); q++)
			if ((f[q] & TypeObstacle) == 0 && (f[q] & TypeInflow) == 0 && (f[q] & TypeOutflow) == 0 && (f[q] & TypeOpen) == 0) f[q] = (f[q] & ~(TypeEmpty | TypeFluid)) | type;
	}
	void updateFromLevelset(LevelsetGrid& levelset) {  // grid.cpp:844-854, on the device
		mpCheck(mp_flags_update_from_levelset(mParent->ctx(), dev(), levelset.dev())); markDeviceWritten();
	}
};

// ---------------------------------------------------------------- particles (particle.h): device-resident arrays, coherent lazily like the grids
template <class T>
class DevArray {                                     // one per-particle array: std::vector on the host, an mp_grid of size (capacity, 1, 1) on the device
public:
	DevArray(FluidSolver* parent) : mParent(parent), mDev(nullptr), mCap(0), mHostDirty(false), mDevDirty(false) {}
	~DevArray() { if (mDev) mp_grid_destroy(mDev); }
	DevArray(const DevArray&) = delete;
	DevArray& operator=(const DevArray&) = delete;
	FluidSolver* getParent() const { return mParent; }
	long long size() const { return (long long)mData.size(); }
	void resize(long long n) { syncToHost(); mData.resize((size_t)n); if (n) mHostDirty = true; }
	const T& operator[](long long idx) const { syncToHost(); return mData[(size_t)idx]; }
	T& operator[](long long idx) { syncToHost(); mHostDirty = true; return mData[(size_t)idx]; }
	mp_grid* dev() const {                             // nullptr for an empty array (the C-ABI accepts it with np == 0)
		const long long n = size();
		if (n == 0) return nullptr;
		if (n > mCap) { reserveDev(n); mHostDirty = true; }
		if (mHostDirty) {
			std::vector<T> buf(mData); buf.resize((size_t)mCap);      // the upload moves `capacity` entries
			mpCheck(mp_grid_upload(mDev, buf.data())); mHostDirty = false;
		}
		return mDev;
	}
	void reserveDev(long long n) const {
		if (n <= mCap) return;
		if (mDev) mpCheck(mp_grid_destroy(mDev));
		mDev = nullptr; mCap = n;
		mpCheck(mp_grid_create(mParent->ctx(), GridKind<T>::kind, (int)sizeof(Real), (int)n, 1, 1, &mDev));
	}
	void syncToHost() const {
		if (!mDevDirty || !mDev) return;
		std::vector<T> buf((size_t)mCap);
		mpCheck(mp_grid_download(mDev, buf.data()));
		std::copy(buf.begin(), buf.begin() + (long long)mData.size(), mData.begin()); mDevDirty = false;
	}
	void markDeviceWritten() { mDevDirty = true; mHostDirty = false; }
protected:
	FluidSolver* mParent; mutable std::vector<T> mData; mutable mp_grid* mDev; mutable long long mCap; mutable bool mHostDirty, mDevDirty;
};
template <class T> class ParticleDataImpl : public DevArray<T> {      // particle.h:392-470
public:
	ParticleDataImpl(FluidSolver* parent) : DevArray<T>(parent) {}
};
enum IntegrationMode { IntEuler = 0, IntRK2, IntRK4 };                 // util/integrator.h:23
class BasicParticleSystem {                          // particle.h:193-274 (BasicParticleData {Vec3 pos; int flag;} as two arrays)
public:
	enum ParticleStatus { PNONE = 0, PNEW = 1 << 0, PSPRAY = 1 << 1, PBUBBLE = 1 << 2, PFOAM = 1 << 3, PTRACER = 1 << 4, PDELETE = 1 << 10, PINVALID = 1 << 30 };   // particle.h:34-43
	BasicParticleSystem(FluidSolver* parent) : mParent(parent), mPos(parent), mFlag(parent) {}
	FluidSolver* getParent() const { return mParent; }
	long long size() const { return mPos.size(); }
	void resizeAll(long long n) { mPos.resize(n); mFlag.resize(n); }
	Vec3 getPos(long long idx) const { return mPos[idx]; }
	void setPos(long long idx, const Vec3& p) { mPos[idx] = p; }
	int getStatus(long long idx) const { return mFlag[idx]; }
	void setStatus(long long idx, int f) { mFlag[idx] = f; }
	bool isActive(long long idx) const { return (mFlag[idx] & PDELETE) == 0; }
	mp_grid* devPos() const { return mPos.dev(); }
	mp_grid* devFlag() const { return mFlag.dev(); }
	void advectInGrid(const FlagGrid& flags, const MACGrid& vel, const int integrationMode, const bool deleteInObstacle = true, const bool stopInObstacle = true,
	                  const bool skipNew = false, const ParticleDataImpl<int>* ptype = NULL, const int exclude = 0) {                       // particle.h:154
		if (!size()) return;
		mpCheck(mp_parts_advect_in_grid(mParent->ctx(), flags.dev(), vel.dev(), size(), devPos(), devFlag(), mParent->getDt(), integrationMode, deleteInObstacle, stopInObstacle,
		                                skipNew, ptype ? ptype->dev() : nullptr, exclude));
		mPos.markDeviceWritten(); mFlag.markDeviceWritten();
	}
	void projectOutOfBnd(const FlagGrid& flags, const Real bnd, const std::string& plane = "xXyYzZ", const ParticleDataImpl<int>* ptype = NULL, const int exclude = 0) {   // particle.h:158
		if (!size()) return;
		mpCheck(mp_parts_project_out_of_bnd(mParent->ctx(), flags.dev(), size(), devPos(), devFlag(), bnd, plane.c_str(), ptype ? ptype->dev() : nullptr, exclude));
		mPos.markDeviceWritten();
	}
	void markPosDeviceWritten() { mPos.markDeviceWritten(); }
private:
	FluidSolver* mParent; DevArray<Vec3> mPos; DevArray<int> mFlag;
};
class ParticleIndexSystem : public DevArray<int> {   // particle.h:276-300: sourceIndex per slot
public:
	ParticleIndexSystem(FluidSolver* parent) : DevArray<int>(parent) {}
	void setCountOnDevice(long long n) { mData.resize((size_t)n); mHostDirty = false; mDevDirty = n > 0; }
};

// ---------------------------------------------------------------- plugins
namespace detail {
template <class G> inline const mp_grid* dv(const G* g) { return g ? g->dev() : nullptr; }
inline mp_pressure_params params(Real cgAccuracy, Real gfClamp, Real cgMaxIterFac, bool precondition, int preconditioner, bool enforceCompatibility,
                                 bool useL2Norm, bool zeroPressureFixing, Real surfTens) {
	mp_pressure_params p;
	p.cgAccuracy = cgAccuracy; p.gfClamp = gfClamp; p.cgMaxIterFac = cgMaxIterFac; p.precondition = precondition; p.preconditioner = preconditioner;
	p.enforceCompatibility = enforceCompatibility; p.useL2Norm = useL2Norm; p.zeroPressureFixing = zeroPressureFixing; p.surfTens = surfTens;
	return p;
}
inline mp_solve_info& lastInfo() { static thread_local mp_solve_info info = mp_solve_info(); return info; }
}  // namespace detail

//! what the reference only prints at debug level 2 (pressure.cpp:440): iterations and residual of the last solve on this thread
inline const mp_solve_info& lastSolveInfo() { return detail::lastInfo(); }

inline void releaseMG(FluidSolver* solver = nullptr) {                       // pressure.cpp:252-266 (all solvers: one context per solver here)
	if (solver) mpCheck(mp_release_mg(solver->ctx()));
}

inline void computePressureRhs(Grid<Real>& rhs, const MACGrid& vel, const Grid<Real>& pressure, const FlagGrid& flags, Real cgAccuracy = 1e-3,
	const Grid<Real>* phi = 0, const Grid<Real>* perCellCorr = 0, const MACGrid* fractions = 0, const MACGrid* obvel = 0, Real gfClamp = 1e-04,
	Real cgMaxIterFac = 1.5, bool precondition = true, int preconditioner = PcMIC, bool enforceCompatibility = false, bool useL2Norm = false,
	bool zeroPressureFixing = false, const Grid<Real>* curv = NULL, const Real surfTens = 0.)
{
	const mp_pressure_params p = detail::params(cgAccuracy, gfClamp, cgMaxIterFac, precondition, preconditioner, enforceCompatibility, useL2Norm, zeroPressureFixing, surfTens);
	mpCheck(mp_compute_pressure_rhs(flags.getParent()->ctx(), rhs.dev(), vel.dev(), pressure.dev(), flags.dev(), detail::dv(phi), detail::dv(perCellCorr),
	                                detail::dv(fractions), detail::dv(obvel), detail::dv(curv), &p));
	rhs.markDeviceWritten();
}

inline void solvePressureSystem(Grid<Real>& rhs, MACGrid& vel, Grid<Real>& pressure, const FlagGrid& flags, Real cgAccuracy = 1e-3,
	const Grid<Real>* phi = 0, const Grid<Real>* perCellCorr = 0, const MACGrid* fractions = 0, Real gfClamp = 1e-04, Real cgMaxIterFac = 1.5,
	bool precondition = true, int preconditioner = PcMIC, const bool enforceCompatibility = false, const bool useL2Norm = false,
	const bool zeroPressureFixing = false, const Grid<Real>* curv = NULL, const Real surfTens = 0.)
{
	const mp_pressure_params p = detail::params(cgAccuracy, gfClamp, cgMaxIterFac, precondition, preconditioner, enforceCompatibility, useL2Norm, zeroPressureFixing, surfTens);
	mpCheck(mp_solve_pressure_system(flags.getParent()->ctx(), rhs.dev(), vel.dev(), pressure.dev(), flags.dev(), detail::dv(phi), detail::dv(perCellCorr),
	                                 detail::dv(fractions), detail::dv(curv), &p, &detail::lastInfo()));
	rhs.markDeviceWritten(); pressure.markDeviceWritten();
}

inline void correctVelocity(MACGrid& vel, Grid<Real>& pressure, const FlagGrid& flags, Real cgAccuracy = 1e-3, const Grid<Real>* phi = 0,
	const Grid<Real>* perCellCorr = 0, const MACGrid* fractions = 0, Real gfClamp = 1e-04, Real cgMaxIterFac = 1.5, bool precondition = true,
	int preconditioner = PcMIC, bool enforceCompatibility = false, bool useL2Norm = false, bool zeroPressureFixing = false,
	const Grid<Real>* curv = NULL, const Real surfTens = 0.)
{
	(void)perCellCorr; (void)fractions;
	const mp_pressure_params p = detail::params(cgAccuracy, gfClamp, cgMaxIterFac, precondition, preconditioner, enforceCompatibility, useL2Norm, zeroPressureFixing, surfTens);
	mpCheck(mp_correct_velocity(flags.getParent()->ctx(), vel.dev(), pressure.dev(), flags.dev(), detail::dv(phi), detail::dv(curv), &p));
	vel.markDeviceWritten();
}

inline void solvePressure(MACGrid& vel, Grid<Real>& pressure, const FlagGrid& flags, Real cgAccuracy = 1e-3, const Grid<Real>* phi = 0,
	const Grid<Real>* perCellCorr = 0, const MACGrid* fractions = 0, const MACGrid* obvel = 0, Real gfClamp = 1e-04, Real cgMaxIterFac = 1.5,
	bool precondition = true, int preconditioner = PcMIC, bool enforceCompatibility = false, bool useL2Norm = false, bool zeroPressureFixing = false,
	const Grid<Real>* curv = NULL, const Real surfTens = 0., Grid<Real>* retRhs = NULL)
{
	const mp_pressure_params p = detail::params(cgAccuracy, gfClamp, cgMaxIterFac, precondition, preconditioner, enforceCompatibility, useL2Norm, zeroPressureFixing, surfTens);
	mpCheck(mp_solve_pressure(flags.getParent()->ctx(), vel.dev(), pressure.dev(), flags.dev(), detail::dv(phi), detail::dv(perCellCorr), detail::dv(fractions),
	                          detail::dv(obvel), detail::dv(curv), retRhs ? retRhs->dev() : nullptr, &p, &detail::lastInfo()));
	vel.markDeviceWritten(); pressure.markDeviceWritten();
	if (retRhs) retRhs->markDeviceWritten();
}

// ---- the steps either side of the projection ----
inline void setWallBcs(const FlagGrid& flags, MACGrid& vel, const MACGrid* obvel = 0, const MACGrid* fractions = 0, const Grid<Real>* phiObs = 0, int boundaryWidth = 0) {
	mpCheck(mp_set_wall_bcs(flags.getParent()->ctx(), flags.dev(), vel.dev(), detail::dv(obvel), detail::dv(fractions), detail::dv(phiObs), boundaryWidth));
	vel.markDeviceWritten();
}
inline void addGravity(const FlagGrid& flags, MACGrid& vel, Vec3 gravity, const Grid<Real>* exclude = NULL, bool scale = true) {
	mpCheck(mp_add_gravity(flags.getParent()->ctx(), flags.dev(), vel.dev(), gravity.x, gravity.y, gravity.z, detail::dv(exclude), scale, flags.getParent()->getDt()));
	vel.markDeviceWritten();
}
inline void addGravityNoScale(const FlagGrid& flags, MACGrid& vel, const Vec3& gravity, const Grid<Real>* exclude = NULL) { addGravity(flags, vel, gravity, exclude, false); }
inline void addBuoyancy(const FlagGrid& flags, const Grid<Real>& density, MACGrid& vel, Vec3 gravity, Real coefficient = 1., bool scale = true) {
	mpCheck(mp_add_buoyancy(flags.getParent()->ctx(), flags.dev(), density.dev(), vel.dev(), gravity.x, gravity.y, gravity.z, coefficient, scale, flags.getParent()->getDt()));
	vel.markDeviceWritten();
}
template <class G>          // G: Grid<Real> (incl. LevelsetGrid), MACGrid or Grid<Vec3> -- the reference takes a GridBase* and dispatches on its type (advection.cpp:447-459)
inline void advectSemiLagrange(const FlagGrid* flags, const MACGrid* vel, G* grid, int order = 1, Real strength = 1.0, int orderSpace = 1, bool openBounds = false,
	int boundaryWidth = -1, int clampMode = 2, int orderTrace = 1)
{
	(void)openBounds; (void)boundaryWidth;             // deprecated in the reference, no effect (advection.cpp:446)
	if (order != 1 && order != 2) throw Error(MP_ERR_INVALID, "AdvectSemiLagrange: Only order 1 (regular SL) and 2 (MacCormack) supported");
	mpCheck((std::is_same<G, Grid<Vec3> >::value ? mp_advect_semi_lagrange_vec3 : mp_advect_semi_lagrange)(flags->getParent()->ctx(), flags->dev(), vel->dev(), grid->dev(),
	        order, strength, orderSpace, clampMode, orderTrace, flags->getParent()->getDt()));
	grid->markDeviceWritten();
}

// ---- liquid neighbours ----
inline void extrapolateMACSimple(FlagGrid& flags, MACGrid& vel, int distance = 4, LevelsetGrid* phiObs = NULL, bool intoObs = false) {
	mpCheck(mp_extrapolate_mac_simple(flags.getParent()->ctx(), flags.dev(), vel.dev(), distance, detail::dv(phiObs), intoObs));
	vel.markDeviceWritten();
}
inline void extrapolateMACFromWeight(MACGrid& vel, Grid<Vec3>& weight, int distance = 2) {      // fastmarch.cpp:410
	mpCheck(mp_extrapolate_mac_from_weight(vel.getParent()->ctx(), vel.dev(), weight.dev(), distance));
	vel.markDeviceWritten(); weight.markDeviceWritten();
}
inline void extrapolateLsSimple(Grid<Real>& phi, int distance = 4, bool inside = false) {
	mpCheck(mp_extrapolate_ls_simple(phi.getParent()->ctx(), phi.dev(), distance, inside));
	phi.markDeviceWritten();
}
inline void extrapolateVec3Simple(Grid<Vec3>& vel, Grid<Real>& phi, int distance = 4, bool inside = false) {
	mpCheck(mp_extrapolate_vec3_simple(vel.getParent()->ctx(), vel.dev(), phi.dev(), distance, inside));
	vel.markDeviceWritten();
}

inline void updateFractions(const FlagGrid& flags, const Grid<Real>& phiObs, MACGrid& fractions, const int& boundaryWidth = 0, const Real fracThreshold = 0.01) {   // initplugins.cpp:437
	mpCheck(mp_update_fractions(flags.getParent()->ctx(), flags.dev(), phiObs.dev(), fractions.dev(), boundaryWidth, fracThreshold));
	fractions.markDeviceWritten();
}
inline void setObstacleFlags(FlagGrid& flags, const Grid<Real>& phiObs, const MACGrid* fractions = NULL, const Grid<Real>* phiOut = NULL, const Grid<Real>* phiIn = NULL,
	int boundaryWidth = 1) {                                                                                                                                          // initplugins.cpp:473
	mpCheck(mp_set_obstacle_flags(flags.getParent()->ctx(), flags.dev(), phiObs.dev(), detail::dv(fractions), detail::dv(phiOut), detail::dv(phiIn), boundaryWidth));
	flags.markDeviceWritten();
}
inline void getLaplacian(Grid<Real>& laplacian, const Grid<Real>& grid) {            // plugin/flip.cpp:710-712
	mpCheck(mp_get_laplacian(grid.getParent()->ctx(), laplacian.dev(), grid.dev()));
	laplacian.markDeviceWritten();
}
inline void getCurvature(Grid<Real>& curv, const Grid<Real>& grid, const Real h = 1.0) {   // plugin/flip.cpp:714-716
	mpCheck(mp_get_curvature(grid.getParent()->ctx(), curv.dev(), grid.dev(), h));
	curv.markDeviceWritten();
}

// ---- the other GridCg callers ----
template <class G>          // G: Grid<Real> or a Vec3 / MAC grid (the reference takes a GridBase&, conjugategrad.cpp:350-351)
inline void cgSolveDiffusion(const FlagGrid& flags, G& grid, Real alpha = 0.25, Real cgMaxIterFac = 1.0, Real cgAccuracy = 1e-4) {
	mpCheck(mp_cg_solve_diffusion(flags.getParent()->ctx(), flags.dev(), grid.dev(), alpha, cgMaxIterFac, cgAccuracy, &detail::lastInfo()));
	grid.markDeviceWritten();
}
inline void cgSolveWE(const FlagGrid& flags, Grid<Real>& ut, Grid<Real>& utm1, Grid<Real>& out, bool crankNic = false, Real cSqr = 0.25, Real cgMaxIterFac = 1.5,
	Real cgAccuracy = 1e-5)                              // plugin/waves.cpp:87-91
{
	mpCheck(mp_cg_solve_we(flags.getParent()->ctx(), flags.dev(), ut.dev(), utm1.dev(), out.dev(), crankNic, cSqr, cgMaxIterFac, cgAccuracy, flags.getParent()->getDt(),
	                       &detail::lastInfo()));
	ut.markDeviceWritten(); utm1.markDeviceWritten(); out.markDeviceWritten();
}

// The grid half of VICintegration plugin/vortexplugins.cpp:253-299 (parameter names and defaults of the plugin :195-196; the mesh and its Peskin
// mapping :203-250 stay with the caller): vorticity grid -> vel (MACGrid: shifted components, Grid<Vec3>: centred).  precondition: 1 PC_ICP,
// 2 PC_mICP; the plugin's default 0 throws setICPreconditioner's error as in the reference (conjugategrad.cpp:312).
inline void vicPoisson(Grid<Vec3>& vel, const FlagGrid& flags, const Grid<Vec3>& vorticity, Real cgMaxIterFac = 1.5, Real cgAccuracy = 1e-3, Real scale = 0.01,
	int precondition = 0, int* iterations = NULL)
{
	mpCheck(mp_vic_poisson(flags.getParent()->ctx(), flags.dev(), vorticity.dev(), vel.dev(), (vel.getType() & GridBase::TypeMAC) ? 1 : 0, cgMaxIterFac, cgAccuracy, scale,
	                       precondition, iterations));
	vel.markDeviceWritten();
}

// ---- FLIP particle <-> grid plugins plugin/flip.cpp (second slice)
namespace detail { template <class T> inline const mp_grid* dp(const ParticleDataImpl<T>* p) { return p ? p->dev() : nullptr; } }
inline void markFluidCells(const BasicParticleSystem& parts, FlagGrid& flags, const Grid<Real>* phiObs = NULL, const ParticleDataImpl<int>* ptype = NULL, const int exclude = 0) {   // :158
	mpCheck(mp_mark_fluid_cells(flags.getParent()->ctx(), parts.size(), parts.devPos(), parts.devFlag(), flags.dev(), detail::dv(phiObs), detail::dp(ptype), exclude));
	flags.markDeviceWritten();
}
inline void gridParticleIndex(const BasicParticleSystem& parts, ParticleIndexSystem& indexSys, const FlagGrid& flags, Grid<int>& index, Grid<int>* counter = NULL) {   // :260
	(void)counter;                                   // the slots of a cell are filled by a stable sort, no counter grid is needed
	long long count = 0;
	if (parts.size()) indexSys.reserveDev(parts.size());
	indexSys.resize(parts.size());
	mpCheck(mp_grid_particle_index(index.getParent()->ctx(), parts.size(), parts.devPos(), parts.devFlag(), parts.size() ? indexSys.dev() : nullptr, flags.dev(), index.dev(), &count));
	indexSys.setCountOnDevice(count); index.markDeviceWritten();
}
inline void unionParticleLevelset(const BasicParticleSystem& parts, const ParticleIndexSystem& indexSys, const FlagGrid& flags, const Grid<int>& index, LevelsetGrid& phi,
                                  const Real radiusFactor = 1., const ParticleDataImpl<int>* ptype = NULL, const int exclude = 0) {       // :340
	mpCheck(mp_union_particle_levelset(phi.getParent()->ctx(), parts.size(), parts.devPos(), indexSys.dev(), indexSys.size(), flags.dev(), index.dev(), phi.dev(), radiusFactor,
	                                   detail::dp(ptype), exclude));
	phi.markDeviceWritten();
}
inline void mapPartsToMAC(const FlagGrid& flags, MACGrid& vel, MACGrid& velOld, const BasicParticleSystem& parts, const ParticleDataImpl<Vec3>& partVel, Grid<Vec3>* weight = NULL,
                          const ParticleDataImpl<int>* ptype = NULL, const int exclude = 0) {                                            // :573
	mpCheck(mp_map_parts_to_mac(vel.getParent()->ctx(), flags.dev(), vel.dev(), velOld.dev(), parts.size(), parts.devPos(), parts.devFlag(), partVel.dev(),
	                            weight ? weight->dev() : nullptr, detail::dp(ptype), exclude));
	vel.markDeviceWritten(); velOld.markDeviceWritten(); if (weight) weight->markDeviceWritten();
}
inline void mapMACToParts(const FlagGrid& flags, const MACGrid& vel, const BasicParticleSystem& parts, ParticleDataImpl<Vec3>& partVel, const ParticleDataImpl<int>* ptype = NULL,
                          const int exclude = 0) {                                                                                      // :651
	mpCheck(mp_map_mac_to_parts(vel.getParent()->ctx(), flags.dev(), vel.dev(), parts.size(), parts.devPos(), parts.devFlag(), partVel.dev(), detail::dp(ptype), exclude));
	if (parts.size()) partVel.markDeviceWritten();
}
inline void flipVelocityUpdate(const FlagGrid& flags, const MACGrid& vel, const MACGrid& velOld, const BasicParticleSystem& parts, ParticleDataImpl<Vec3>& partVel, const Real flipRatio,
                               const ParticleDataImpl<int>* ptype = NULL, const int exclude = 0) {                                       // :669
	mpCheck(mp_flip_velocity_update(vel.getParent()->ctx(), flags.dev(), vel.dev(), velOld.dev(), parts.size(), parts.devPos(), parts.devFlag(), partVel.dev(), flipRatio,
	                                detail::dp(ptype), exclude));
	if (parts.size()) partVel.markDeviceWritten();
}
// ---- the Lagrangian-particle helpers of scenes/benchmark_dam.py:118-134 (plugin/ptsplugins.cpp:26-65, grid.cpp:885-890)
inline void addForcePvel(ParticleDataImpl<Vec3>& vel, const Vec3& a, const Real dt, const ParticleDataImpl<int>* ptype, const int exclude) {
	if (!vel.size()) return;
	mpCheck(mp_add_force_pvel(vel.getParent()->ctx(), vel.size(), vel.dev(), a.x, a.y, a.z, dt, detail::dp(ptype), exclude)); vel.markDeviceWritten();
}
inline void updateVelocityFromDeltaPos(const BasicParticleSystem& parts, ParticleDataImpl<Vec3>& vel, const ParticleDataImpl<Vec3>& x_prev, const Real dt, const ParticleDataImpl<int>* ptype,
                                       const int exclude) {
	if (!parts.size()) return;
	mpCheck(mp_update_velocity_from_delta_pos(parts.getParent()->ctx(), parts.size(), parts.devPos(), vel.dev(), x_prev.dev(), dt, detail::dp(ptype), exclude)); vel.markDeviceWritten();
}
inline void eulerStep(BasicParticleSystem& parts, const ParticleDataImpl<Vec3>& vel, const ParticleDataImpl<int>* ptype, const int exclude) {
	if (!parts.size()) return;
	mpCheck(mp_euler_step(parts.getParent()->ctx(), parts.size(), parts.devPos(), vel.dev(), parts.getParent()->getDt(), detail::dp(ptype), exclude)); parts.markPosDeviceWritten();
}
inline void setPartType(const BasicParticleSystem& parts, ParticleDataImpl<int>& ptype, const int mark, const int stype, const FlagGrid& flags, const int cflag) {
	if (!parts.size()) return;
	mpCheck(mp_set_part_type(flags.getParent()->ctx(), parts.size(), parts.devPos(), ptype.dev(), mark, stype, flags.dev(), cflag)); ptype.markDeviceWritten();
}
inline void markIsolatedFluidCell(FlagGrid& flags, const int mark) {
	mpCheck(mp_mark_isolated_fluid_cell(flags.getParent()->ctx(), flags.dev(), mark)); flags.markDeviceWritten();
}
inline void pushOutofObs(BasicParticleSystem& parts, const FlagGrid& flags, const Grid<Real>& phiObs, const Real shift = 0, const Real thresh = 0,
                         const ParticleDataImpl<int>* ptype = NULL, const int exclude = 0) {                                             // :542
	if (!parts.size()) return;
	mpCheck(mp_push_out_of_obs(phiObs.getParent()->ctx(), parts.size(), parts.devPos(), parts.devFlag(), flags.dev(), phiObs.dev(), shift, thresh, detail::dp(ptype), exclude));
	parts.markPosDeviceWritten();
}

}  // namespace Manta

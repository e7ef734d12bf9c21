/* mantapress.h -- C-ABI of the B200-native (sm_100a) pressure-projection path of mantaflow.
 *
 * This is the drop-in boundary: everything the reference's plugin/pressure.cpp, conjugategrad.{h,cpp},
 * multigrid.{h,cpp} and the Grid storage of grid.cpp need from the device is reachable through the
 * entry points below (plain pointers and sizes, int status codes, no C++ / torch types).  Each entry
 * point cites the reference interface (file:line, relative to the mantaflow tree) it replaces.
 * The reference-side binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - Layout is the reference's (grid.h:70): idx = i + sx*(j + sy*k), x fastest; flags are int32 bit
 *    masks (grid.h:292-304); MAC grids are AoS {x,y,z} (vectorbase.h:199-213). 2-D grids have sz == 1.
 *  - "Real" is chosen per grid at run time: prec = 4 (float, reference build fp1) or 8 (double, fp2,
 *    -DDOUBLEPRECISION).  All grids of one call must share size and precision.  Scalars cross the ABI
 *    as double and are narrowed to Real inside, exactly like the reference's Python->Real conversion.
 *  - Every function returns MP_OK (0) or an error code; mp_last_error() gives the message of the last
 *    failure on the calling thread.  Nothing throws across the ABI.  The host mirror turns codes into
 *    Manta::Error / RuntimeError (general.h:42-57, pclass.cpp:57-61).
 *  - All work is enqueued on the context's CUDA stream; functions that return host-visible results
 *    synchronise that stream.  A context is bound to one device; one host thread drives it
 *    (the reference runs plugins on the interpreter thread).
 *  - There is NO CPU fallback: without a CUDA device mp_context_create fails with MP_ERR_CUDA.
 */
#ifndef MANTAPRESS_H
#define MANTAPRESS_H

#ifdef __cplusplus
extern "C" {
#endif

#define MP_VERSION 100

enum mp_status {
	MP_OK = 0,
	MP_ERR_INVALID = 1,      /* bad argument / size or precision mismatch / fluid cell on the outer layer */
	MP_ERR_CUDA = 2,         /* CUDA runtime error (message has file:line) */
	MP_ERR_DIVERGED = 3,     /* "GridCg::iterate: The CG solver diverged" conjugategrad.cpp:288-295 */
	MP_ERR_NOT_SET = 4,      /* "GridMg::setRhs Error: A has not been set." multigrid.cpp:428,:453 */
	MP_ERR_UNSUPPORTED = 5,
	MP_ERR_COMM = 6          /* NCCL / peer-memory error in the multi-GPU path */
};

/* Grid element kinds (GridBase::GridType grid.h:29) */
enum mp_grid_kind { MP_GRID_REAL = 1, MP_GRID_FLAGS = 2, MP_GRID_MAC = 8 };

/* enum Preconditioner plugin/pressure.cpp:27, python/defines.py:46-50 */
enum mp_preconditioner { MP_PC_NONE = 0, MP_PC_MIC = 1, MP_PC_MG_DYNAMIC = 2, MP_PC_MG_STATIC = 3 };

/* GridCgInterface::PreconditionType conjugategrad.h:29 */
enum mp_cg_pc_type { MP_CG_PC_NONE = 0, MP_CG_PC_ICP = 1, MP_CG_PC_MICP = 2, MP_CG_PC_MGP = 3 };

typedef struct mp_context mp_context;   /* device, stream, scratch; also plays FluidSolver for gMapMG (pressure.cpp:250) */
typedef struct mp_grid mp_grid;         /* device-resident mirror of Grid<Real> / MACGrid / FlagGrid storage */
typedef struct mp_cg mp_cg;             /* GridCg<ApplyMatrix|ApplyMatrix2D> conjugategrad.h:65-114 */
typedef struct mp_mg mp_mg;             /* GridMg multigrid.h:31-137 */

/* The keyword tail shared by the four pressure plugins (pressure.cpp:277-292,:312-326,:455-468,:480-495).
 * Defaults are the reference's: see mp_pressure_params_default(). */
typedef struct mp_pressure_params {
	double cgAccuracy;            /* 1e-3 */
	double gfClamp;               /* 1e-4 */
	double cgMaxIterFac;          /* 1.5  */
	int    precondition;          /* true; deprecated switch: false forces PcNone (pressure.cpp:328) */
	int    preconditioner;        /* PcMIC */
	int    enforceCompatibility;  /* false */
	int    useL2Norm;             /* false */
	int    zeroPressureFixing;    /* false */
	double surfTens;              /* 0. */
} mp_pressure_params;

/* What the reference only prints (pressure.cpp:440 debMsg level 2) plus device timings. */
typedef struct mp_solve_info {
	int       iterations;         /* gcg->getIterations() */
	double    resNorm;            /* gcg->getResNorm() */
	int       maxIter;            /* the cap that was applied (pressure.cpp:408,:419) */
	long long fixedCell;          /* pinned cell index or -1 (pressure.cpp:383-387) */
	int       mgLevels;           /* GridMg levels (0 unless PcMG*) */
	float     msRhs, msMatrix, msSolve, msCorrect, msTotal;  /* CUDA-event times of the stages */
	float     msH2D, msD2H;       /* only filled by the *_host entry points */
	/* per-kernel averages over sampled launches of the CG loop (mp_context_set_profiling), 0 when off */
	float     msMatvecAvg, msAxpyAvg, msUpdateAvg, msPrecondAvg;
	int       profSamples;
	int       matvecKernel;       /* 0 k_matvec_dot (L2 reuse), 1 k_matvec_zmarch (4+6w B/cell), 2 k_matvec_zmarch_masked (4+3w), 3 k_matvec_fused (PcNone loop, 4+7w incl. the s and x updates), 4 k_matvec_fused_tma (the same loop staged by TMA, matrix as 2 B/cell: 2+6w) */
} mp_solve_info;

/* ---- library / errors ---- */
int         mp_version(void);
const char* mp_last_error(void);
const char* mp_status_string(int status);
int         mp_device_count(int* count);                       /* 0 devices is MP_OK with *count = 0 */
void        mp_pressure_params_default(mp_pressure_params* p);

/* ---- context ---- */
int   mp_context_create(int device, mp_context** out);
int   mp_context_destroy(mp_context* ctx);
int   mp_context_synchronize(mp_context* ctx);
/* Freed grids are kept in a per-context pool (FluidSolver::GridStorage, fluidsolver.cpp:33-50) bounded by bytes (16 blocks of the
 * largest size seen); this returns all of them to the driver. */
int   mp_context_trim(mp_context* ctx);
void* mp_context_stream(mp_context* ctx);                      /* cudaStream_t all work is enqueued on */
int   mp_context_device(const mp_context* ctx);
int   mp_context_sm_count(const mp_context* ctx);
int   mp_context_kernel_launches(const mp_context* ctx, long long* count);  /* kernels launched so far */
/* every `period`-th CG iteration brackets its kernels with CUDA events (0 = off); results land in mp_solve_info */
int   mp_context_set_profiling(mp_context* ctx, int period);

/* ---- Grid storage mirror: Grid<T> ctor/dtor grid.cpp:47-91, FluidSolver::GridStorage fluidsolver.cpp:33-50 ---- */
int   mp_grid_create(mp_context* ctx, int kind, int prec, int sx, int sy, int sz, mp_grid** out);  /* zero-filled like Grid<T>(parent) */
int   mp_grid_destroy(mp_grid* g);
int   mp_grid_upload(mp_grid* g, const void* host);            /* host: sx*sy*sz (*3 for MAC) elements, reference layout */
int   mp_grid_download(const mp_grid* g, void* host);
int   mp_grid_upload_async(mp_grid* g, const void* pinned_host);
int   mp_grid_download_async(const mp_grid* g, void* pinned_host);
int   mp_grid_clear(mp_grid* g);                               /* Grid<T>::clear grid.cpp:93-96 */
int   mp_grid_copy_from(mp_grid* dst, const mp_grid* src);     /* Grid<T>::copyFrom grid.cpp:205-210 */
void* mp_grid_device_ptr(mp_grid* g);
int   mp_grid_info(const mp_grid* g, int* kind, int* prec, int* sx, int* sy, int* sz);
/* pinned host staging for plugin-boundary transfers */
int   mp_host_alloc(void** out, unsigned long long bytes);
int   mp_host_free(void* p);

/* ---- Grid<Real> reductions / BLAS-1 used by GridCg ---- */
int mp_grid_dot(mp_context* ctx, const mp_grid* a, const mp_grid* b, double* out);        /* GridDotProduct conjugategrad.cpp:175-178 */
int mp_grid_max_abs(mp_context* ctx, const mp_grid* a, double* out);                      /* Grid<Real>::getMaxAbs grid.cpp:319-323; Vec3 / MAC grids: Grid<Vec3>::getMaxAbs :330-332 */
int mp_grid_sum_sqr(mp_context* ctx, const mp_grid* a, double* out);                      /* GridSumSqr commonkernels.h:32-35 */
int mp_grid_scaled_add(mp_context* ctx, mp_grid* me, const mp_grid* other, double factor);/* gridScaledAdd grid.h:478 */
int mp_grid_add_const(mp_context* ctx, mp_grid* me, double value);                        /* Grid<T>::operator+=(S) grid.h:490 */
/* Element-wise Grid<T> arithmetic on the device, grid.cpp:258-284 (Real, int and Vec3 / MAC grids): setConst, addConst, multConst, add, sub, mult,
 * addScaled, clamp, stomp, safeDivide.  (x, y, z): the constant / factor / threshold -- x alone for Real and int grids, per component for Vec3 grids;
 * clamp: x = min, y = max for every component (Grid<T>::clamp(Real min, Real max) builds T(min), T(max)).  `other` is NULL for the constant operations. */
enum mp_grid_arith_op { MP_OP_SET_CONST = 0, MP_OP_ADD_CONST = 1, MP_OP_MULT_CONST = 2, MP_OP_ADD = 3, MP_OP_SUB = 4, MP_OP_MULT = 5, MP_OP_ADD_SCALED = 6,
                        MP_OP_CLAMP = 7, MP_OP_STOMP = 8, MP_OP_SAFE_DIVIDE = 9 };
int mp_grid_arith(mp_context* ctx, mp_grid* me, int op, const mp_grid* other, double x, double y, double z);

/* ---- assembly kernels ---- */
/* MakeRhs pressure.cpp:32-84 (+ the optional mean subtraction of computePressureRhs :297-298 is NOT
 * applied here, see mp_compute_pressure_rhs).  Optional grids may be NULL. */
int mp_make_rhs(mp_context* ctx, const mp_grid* flags, mp_grid* rhs, const mp_grid* vel,
                const mp_grid* perCellCorr, const mp_grid* fractions, const mp_grid* obvel,
                const mp_grid* phi, const mp_grid* curv, double surfTens, double gfClamp,
                double* sum, int* cnt);
/* MakeLaplaceMatrix conjugategrad.h:154-187; writes every cell (the reference relies on cleared grids) */
int mp_make_laplace_matrix(mp_context* ctx, const mp_grid* flags, mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak,
                           const mp_grid* fractions);
/* ApplyGhostFluidDiagonal pressure.cpp:136-151 */
int mp_apply_ghost_fluid_diagonal(mp_context* ctx, mp_grid* A0, const mp_grid* flags, const mp_grid* phi, double gfClamp);
/* CountEmptyCells pressure.cpp:217-220 */
int mp_count_empty_cells(mp_context* ctx, const mp_grid* flags, long long* numEmpty);
/* the cell choice of pressure.cpp:352-382: -1 when an empty cell exists or no fluid cell is found */
int mp_choose_fix_cell(mp_context* ctx, const mp_grid* flags, long long* fixPidx);
/* fixPressure pressure.cpp:226-245 */
int mp_fix_pressure(mp_context* ctx, long long fixPidx, double value, mp_grid* rhs, mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak);
/* ApplyMatrix / ApplyMatrix2D conjugategrad.h:118-151 (chosen by sz) */
int mp_apply_matrix(mp_context* ctx, const mp_grid* flags, mp_grid* dst, const mp_grid* src,
                    const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak);
/* InitPreconditionModifiedIncompCholesky2 / ApplyPreconditionModifiedIncompCholesky2 conjugategrad.cpp:66-97,:135-159 */
/* The ordering MIC(0) is formed in.  mode 0 (default): the reference's lexicographic ordering -- Aprecond and the sweeps are bit-identical to
 * conjugategrad.cpp:66-97,:135-159 and PcMIC keeps the reference's iteration counts, but the sweeps are a chain of sx+sy+sz dependent
 * hyperplanes.  mode 1: block red-black ordering, the reformulated ("coloured") triangular solves: tiles of tileY x tileZ rows (rounded up to
 * multiples of 8 and 4; 0 = chosen from the grid size), an exact MIC(0) of the permuted matrix, bandwidth-bound; more iterations (reported in mp_solve_info as usual).
 * Applies to mp_mic_init / mp_mic_apply, GridCg with PC_mICP and solvePressure with PcMIC on unsharded 3-D grids whose off-diagonals are 0 / -1
 * (no face fractions); other cases keep mode 0.  No reference counterpart: specified by the test oracle's restatement micrb_init / micrb_apply. */
int mp_set_mic_ordering(mp_context* ctx, int mode, int tileY, int tileZ);
/* what the last MIC(0) factorisation of the context used: mode 1 + its tile when it ran in block red-black ordering, else 0 / 0 / 0 */
int mp_get_mic_ordering(const mp_context* ctx, int* mode, int* tileY, int* tileZ);
int mp_mic_init(mp_context* ctx, const mp_grid* flags, mp_grid* Aprecond, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak);
int mp_mic_apply(mp_context* ctx, mp_grid* dst, const mp_grid* var1, const mp_grid* flags, const mp_grid* Aprecond,
                 const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak);
/* IC(0) "a la Wavelet Turbulence" (PC_ICP): InitPreconditionIncompCholesky conjugategrad.cpp:26-63, ApplyPreconditionIncompCholesky :109-132.
 * P0..Pk receive the factor (a scaled copy of the matrix, diagonal inverted; bit-identical to the reference's four grids). 3-D only, like the reference. */
int mp_ic_init(mp_context* ctx, const mp_grid* flags, mp_grid* P0, mp_grid* Pi, mp_grid* Pj, mp_grid* Pk, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak);
int mp_ic_apply(mp_context* ctx, mp_grid* dst, const mp_grid* var1, const mp_grid* flags, const mp_grid* P0, const mp_grid* Pi, const mp_grid* Pj, const mp_grid* Pk);

/* ---- GridCg conjugategrad.h:65-114 ---- */
int mp_cg_create(mp_context* ctx, mp_grid* dst, mp_grid* rhs, mp_grid* residual, mp_grid* search, const mp_grid* flags, mp_grid* tmp,
                 mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak, mp_cg** out);                 /* ctor conjugategrad.cpp:201-207 */
int mp_cg_destroy(mp_cg* cg);
int mp_cg_set_accuracy(mp_cg* cg, double accuracy);                                                /* setAccuracy :87 */
int mp_cg_set_use_l2_norm(mp_cg* cg, int useL2);                                                   /* setUseL2Norm :52 */
int mp_cg_set_ic_preconditioner(mp_cg* cg, int method, mp_grid* A0, mp_grid* Ai, mp_grid* Aj, mp_grid* Ak); /* :310-326; MP_CG_PC_NONE accepted */
int mp_cg_set_mg_preconditioner(mp_cg* cg, int method, mp_mg* mg);                                 /* :328-335 */
int mp_cg_force_reinit(mp_cg* cg);                                                                 /* forceReinit :79 */
int mp_cg_iterate(mp_cg* cg, int* keepGoing);                                                      /* iterate :237-299 (one iteration, synchronous) */
int mp_cg_solve(mp_cg* cg, int maxIter);                                                           /* solve :301-307 (whole loop on the device) */
int mp_cg_get(mp_cg* cg, int* iterations, double* resNorm, double* sigma);                         /* getIterations/getResNorm/getSigma :82-85 */

/* ---- GridMg multigrid.h:31-137 ---- */
int mp_mg_create(mp_context* ctx, int prec, int sx, int sy, int sz, mp_mg** out);                  /* GridMg::GridMg multigrid.cpp:220-319 */
int mp_mg_destroy(mp_mg* mg);
int mp_mg_set_a(mp_mg* mg, const mp_grid* A0, const mp_grid* Ai, const mp_grid* Aj, const mp_grid* Ak);  /* setA :386-415 */
int mp_mg_set_rhs(mp_mg* mg, const mp_grid* rhs);                                                  /* setRhs :426-433 */
int mp_mg_is_a_set(const mp_mg* mg, int* isSet);
int mp_mg_do_vcycle(mp_mg* mg, mp_grid* dst, const mp_grid* src /* may be NULL */, double* resNorm /* may be NULL */); /* doVCycle :448-504 */
int mp_mg_set_coarsest_level_accuracy(mp_mg* mg, double accuracy);
int mp_mg_set_smoothing(mp_mg* mg, int numPreSmooth, int numPostSmooth);
int mp_mg_num_levels(const mp_mg* mg, int* levels);
int mp_mg_level_info(const mp_mg* mg, int level, int* sx, int* sy, int* sz, int* stencil);
/* parity probes: copy one level's vertex types (int8) / operator (interleaved, multigrid.cpp:208-218) / x,b,r to the host */
int mp_mg_download(const mp_mg* mg, int level, const char* what /* "type","a","x","b","r" */, void* host);
/* 1 when level 0 of the V-cycle (doVCycle :458-466,:484-494 on level 0) runs as the fused single-pass kernels: the operator set by the last
   setA is codable as 2 bytes per vertex (no face fractions), rows are 16-byte aligned, single GPU.  0: the per-colour kernels run. */
int mp_mg_level0_fused(const mp_mg* mg, int* fused);

/* ---- the plugins, device-resident grids (fields stay in HBM for the whole projection) ---- */
/* releaseMG pressure.cpp:252-266 (the context plays the FluidSolver key of gMapMG) */
int mp_release_mg(mp_context* ctx);
/* computePressureRhs pressure.cpp:277-299 */
int mp_compute_pressure_rhs(mp_context* ctx, mp_grid* rhs, const mp_grid* vel, const mp_grid* pressure, const mp_grid* flags,
                            const mp_grid* phi, const mp_grid* perCellCorr, const mp_grid* fractions, const mp_grid* obvel,
                            const mp_grid* curv, const mp_pressure_params* params);
/* solvePressureSystem pressure.cpp:312-452 */
int mp_solve_pressure_system(mp_context* ctx, mp_grid* rhs, mp_grid* vel, mp_grid* pressure, const mp_grid* flags,
                             const mp_grid* phi, const mp_grid* perCellCorr, const mp_grid* fractions,
                             const mp_grid* curv, const mp_pressure_params* params, mp_solve_info* info);
/* correctVelocity pressure.cpp:455-476 */
int mp_correct_velocity(mp_context* ctx, mp_grid* vel, const mp_grid* pressure, const mp_grid* flags,
                        const mp_grid* phi, const mp_grid* curv, const mp_pressure_params* params);
/* solvePressure pressure.cpp:480-521 */
int mp_solve_pressure(mp_context* ctx, mp_grid* vel, mp_grid* pressure, const mp_grid* flags,
                      const mp_grid* phi, const mp_grid* perCellCorr, const mp_grid* fractions, const mp_grid* obvel,
                      const mp_grid* curv, mp_grid* retRhs, const mp_pressure_params* params, mp_solve_info* info);

/* ---- another GridCg caller on the same device solver ----
 * cgSolveDiffusion conjugategrad.cpp:350-423: implicit diffusion (I + alpha*L) u = u_old of a Real grid or, component by
 * component, of a Vec3 / MAC grid; plain CG, GridCg's default L2 stop test, maxIter = (int)(cgMaxIterFac*maxDim) (x4 in 2-D).
 * info (optional) reports the iterations / residual of the last component solved. */
int mp_cg_solve_diffusion(mp_context* ctx, const mp_grid* flags, mp_grid* grid, double alpha, double cgMaxIterFac, double cgAccuracy,
                          mp_solve_info* info);

/* cgSolveWE plugin/waves.cpp:86-147: one implicit (optionally Crank-Nicolson) step of the wave equation; afterwards utm1 holds the old ut and
 * ut the new solution (= out), as the plugin leaves them.  dt = FluidSolver::getDt(). */
int mp_cg_solve_we(mp_context* ctx, const mp_grid* flags, mp_grid* ut, mp_grid* utm1, mp_grid* out, int crankNic, double cSqr, double cgMaxIterFac,
                   double cgAccuracy, double dt, mp_solve_info* info);

/* The grid half of VICintegration plugin/vortexplugins.cpp:253-299 (the VIC Poisson solve): from the vorticity grid the Peskin
 * kernel of :203-250 leaves (a centred Grid<Vec3>) to the velocity -- MakeLaplaceMatrix, CurlOp, then for each of the three components
 * GetShiftedComponent (velIsMac != 0: vel is a MACGrid) or GetComponent, a GridCg<ApplyMatrix> solve with setUseL2Norm(true) and
 * setICPreconditioner(PreconditionType(precondition)), solution *= scale, SetComponent.  precondition: 1 = PC_ICP, 2 = PC_mICP; as in the
 * reference every other value (the plugin's default 0 included) fails with setICPreconditioner's message (conjugategrad.cpp:312).
 * iterations: int[3], the GridCg iteration count per component (the plugin's debMsg line :295), may be NULL. */
int mp_vic_poisson(mp_context* ctx, const mp_grid* flags, const mp_grid* vorticity, mp_grid* vel, int velIsMac, double cgMaxIterFac, double cgAccuracy,
                   double scale, int precondition, int* iterations);

/* ---- the steps either side of the projection, so that a whole smoke step keeps its fields in HBM ----
 * setWallBcs          plugin/extforces.cpp:186-218, :307-316  (KnSetWallBcs; with phiObs AND fractions the second-order variant KnSetWallBcsFrac :220-303)
 * addGravity          plugin/extforces.cpp:45-65              (scale != 0: divided by the grid's dx = 1/max(size))
 * addBuoyancy         plugin/extforces.cpp:75-90
 * advectSemiLagrange  plugin/advection.cpp:442-461            (grid: Real / Levelset or MAC, _vec3: a cell-centred Grid<Vec3> in MP_GRID_MAC storage;
 *                     order 1 | 2 (MacCormack), clampMode 1 | 2, orderSpace 1 | 2 (cubic lookups, util/interpolHigh.h), orderTrace 1 | 2 (explicit
 *                     midpoint, advection.cpp:32-37, :58-73), convective outflow boundary for MAC grids).  dt = FluidSolver::getDt().
 * Results are bit-identical to the reference's in both precisions. */
int mp_set_wall_bcs(mp_context* ctx, const mp_grid* flags, mp_grid* vel, const mp_grid* obvel, const mp_grid* fractions, const mp_grid* phiObs, int boundaryWidth);
int mp_add_gravity(mp_context* ctx, const mp_grid* flags, mp_grid* vel, double gx, double gy, double gz, const mp_grid* exclude, int scale, double dt);
int mp_add_buoyancy(mp_context* ctx, const mp_grid* flags, const mp_grid* density, mp_grid* vel, double gx, double gy, double gz, double coefficient, int scale, double dt);
int mp_advect_semi_lagrange(mp_context* ctx, const mp_grid* flags, const mp_grid* vel, mp_grid* grid, int order, double strength, int orderSpace,
                            int clampMode, int orderTrace, double dt);
int mp_advect_semi_lagrange_vec3(mp_context* ctx, const mp_grid* flags, const mp_grid* vel, mp_grid* grid, int order, double strength, int orderSpace,
                                 int clampMode, int orderTrace, double dt);

/* ---- liquid neighbours (first slice): with them the level-set free-surface loop of scenes/freesurface.py:54-84
 * (extrapolateLsSimple x2, extrapolateMACSimple, advect phi, phi.setBound, flags.updateFromLevelset, advect vel, addGravity, setWallBcs,
 * solvePressure with phi) keeps every field in HBM.
 * extrapolateMACSimple          fastmarch.cpp:337-375  (distance <= 250; phiObs may be NULL)
 * extrapolateMACFromWeight      fastmarch.cpp:410-432  (weight: a Vec3 grid whose positive entries mark initialised faces; it is destroyed)
 * extrapolateLsSimple           fastmarch.cpp:470-507
 * extrapolateVec3Simple         fastmarch.cpp:510-542  (vel: a Vec3 grid, same storage as MP_GRID_MAC)
 * FlagGrid::updateFromLevelset  grid.cpp:844-854
 * Grid<T>::setBound             grid.cpp:585-593       (value = vx for Real / flag grids, (vx,vy,vz) for Vec3 grids)
 * Results are bit-identical to the reference's in both precisions. */
int mp_extrapolate_mac_simple(mp_context* ctx, const mp_grid* flags, mp_grid* vel, int distance, const mp_grid* phiObs, int intoObs);
int mp_extrapolate_mac_from_weight(mp_context* ctx, mp_grid* vel, mp_grid* weight, int distance);
int mp_extrapolate_ls_simple(mp_context* ctx, mp_grid* phi, int distance, int inside);
int mp_extrapolate_vec3_simple(mp_context* ctx, mp_grid* vel, const mp_grid* phi, int distance, int inside);
int mp_flags_update_from_levelset(mp_context* ctx, mp_grid* flags, const mp_grid* levelset);
int mp_grid_set_bound(mp_context* ctx, mp_grid* g, double vx, double vy, double vz, int boundaryWidth);
/* updateFractions / setObstacleFlags plugin/initplugins.cpp:437-440,:473-475: the producers of the `fractions` argument of solvePressure and
 * setWallBcs (second-order obstacle boundaries from an obstacle level set).  updateFractions gives the result of the reference's serial loop
 * (its OpenMP build races on the max-side walls for boundaryWidth > 0); setObstacleFlags needs boundaryWidth >= 1. Optional grids may be NULL. */
int mp_update_fractions(mp_context* ctx, const mp_grid* flags, const mp_grid* phiObs, mp_grid* fractions, int boundaryWidth, double fracThreshold);
int mp_set_obstacle_flags(mp_context* ctx, mp_grid* flags, const mp_grid* phiObs, const mp_grid* fractions, const mp_grid* phiOut, const mp_grid* phiIn, int boundaryWidth);
/* getLaplacian / getCurvature plugin/flip.cpp:710-716 (LaplaceOp, CurvatureOp commonkernels.h:75-101): the `curv` input of the surface-tension
 * variant of solvePressure.  Cells of the outer layer keep their content (KERNEL(bnd=1)); result and input must be different grids. */
int mp_get_laplacian(mp_context* ctx, mp_grid* laplacian, const mp_grid* grid);
int mp_get_curvature(mp_context* ctx, mp_grid* curv, const mp_grid* grid, double h);

/* ---- FLIP particle <-> grid plugins (second slice; the loop of scenes/benchmark_dam.py:100-125 / flip02_surface.py).
 * A particle system is a set of device arrays, each created with mp_grid_create(ctx, kind, prec, capacity, 1, 1):
 *   pos      MP_GRID_MAC    BasicParticleData::pos   particle.h:182-191        pflag    MP_GRID_FLAGS  BasicParticleData::flag (PDELETE = 1 << 10)
 *   partVel  MP_GRID_MAC    ParticleDataImpl<Vec3>   particle.h:392            ptype    MP_GRID_FLAGS  ParticleDataImpl<int> (optional, may be NULL)
 *   indexSys MP_GRID_FLAGS  ParticleIndexSystem      particle.h:276 (sourceIndex of every slot)
 * np = particles in use (<= capacity).  `flags` arguments that the reference only takes for the grid size may be NULL.
 * markFluidCells         plugin/flip.cpp:158-177
 * gridParticleIndex      plugin/flip.cpp:260-306  (*count = indexed particles = indexSys.size(); one stream synchronisation)
 * unionParticleLevelset  plugin/flip.cpp:340-350  (count = what gridParticleIndex returned)
 * mapPartsToMAC          plugin/flip.cpp:573-595  (weight may be NULL; faces sum in particle order like the reference's serial kernel)
 * mapMACToParts          plugin/flip.cpp:651-656
 * flipVelocityUpdate     plugin/flip.cpp:669-677
 * Results are bit-identical to the reference's in both precisions and independent of the launch geometry (no floating-point atomics). */
/* ParticleSystem<S>::advectInGrid particle.h:512-536: positions (and, with deleteInObstacle, the PDELETE flags) are updated in place on the device.
 * integrationMode: 0 IntEuler, 1 IntRK2, 2 IntRK4 (util/integrator.h:23); dt = FluidSolver::getDt(). One pass over the particles for all stages. */
int mp_parts_advect_in_grid(mp_context* ctx, const mp_grid* flags, const mp_grid* vel, long long np, mp_grid* pos, mp_grid* pflag, double dt, int integrationMode,
                            int deleteInObstacle, int stopInObstacle, int skipNew, const mp_grid* ptype, int exclude);
/* pushOutofObs plugin/flip.cpp:542-545 and ParticleSystem<S>::projectOutOfBnd particle.h:578-590 (`plane`: any of the letters "xXyYzZ"): positions updated in place. */
int mp_push_out_of_obs(mp_context* ctx, long long np, mp_grid* pos, const mp_grid* pflag, const mp_grid* flags, const mp_grid* phiObs, double shift, double thresh,
                       const mp_grid* ptype, int exclude);
int mp_parts_project_out_of_bnd(mp_context* ctx, const mp_grid* flags, long long np, mp_grid* pos, const mp_grid* pflag, double bnd, const char* plane,
                                const mp_grid* ptype, int exclude);
/* The Lagrangian-particle helpers of scenes/benchmark_dam.py:118-134: addForcePvel, updateVelocityFromDeltaPos, eulerStep, setPartType
 * (plugin/ptsplugins.cpp:26-29,:38-41,:50-53,:62-65; dt of eulerStep = FluidSolver::getDt()), ParticleSystem::getPosPdata (particle.h:422-427),
 * markIsolatedFluidCell (grid.cpp:885-890).  With them every plugin of that scene's main loop has a device version. */
int mp_add_force_pvel(mp_context* ctx, long long np, mp_grid* vel, double ax, double ay, double az, double dt, const mp_grid* ptype, int exclude);
int mp_update_velocity_from_delta_pos(mp_context* ctx, long long np, const mp_grid* pos, mp_grid* vel, const mp_grid* xPrev, double dt, const mp_grid* ptype, int exclude);
int mp_euler_step(mp_context* ctx, long long np, mp_grid* pos, const mp_grid* vel, double dt, const mp_grid* ptype, int exclude);
int mp_set_part_type(mp_context* ctx, long long np, const mp_grid* pos, mp_grid* ptype, int mark, int stype, const mp_grid* flags, int cflag);
int mp_parts_get_pos_pdata(mp_context* ctx, long long np, const mp_grid* pos, mp_grid* target);
int mp_mark_isolated_fluid_cell(mp_context* ctx, mp_grid* flags, int mark);
int mp_mark_fluid_cells(mp_context* ctx, long long np, const mp_grid* pos, const mp_grid* pflag, mp_grid* flags, const mp_grid* phiObs, const mp_grid* ptype, int exclude);
int mp_grid_particle_index(mp_context* ctx, long long np, const mp_grid* pos, const mp_grid* pflag, mp_grid* indexSys, const mp_grid* flags, mp_grid* index, long long* count);
int mp_union_particle_levelset(mp_context* ctx, long long np, const mp_grid* pos, const mp_grid* indexSys, long long count, const mp_grid* flags, const mp_grid* index,
                               mp_grid* phi, double radiusFactor, const mp_grid* ptype, int exclude);
int mp_map_parts_to_mac(mp_context* ctx, const mp_grid* flags, mp_grid* vel, mp_grid* velOld, long long np, const mp_grid* pos, const mp_grid* pflag, const mp_grid* partVel,
                        mp_grid* weight, const mp_grid* ptype, int exclude);
int mp_map_mac_to_parts(mp_context* ctx, const mp_grid* flags, const mp_grid* vel, long long np, const mp_grid* pos, const mp_grid* pflag, mp_grid* partVel,
                        const mp_grid* ptype, int exclude);
int mp_flip_velocity_update(mp_context* ctx, const mp_grid* flags, const mp_grid* vel, const mp_grid* velOld, long long np, const mp_grid* pos, const mp_grid* pflag,
                            mp_grid* partVel, double flipRatio, const mp_grid* ptype, int exclude);

/* ---- PD_fluid_guiding plugin/fluidguiding.cpp:294-353: primal-dual guiding of vel towards velT with per-cell weight;
 * up to maxIters solvePressure calls on device-resident copies, separable Gaussian blurs of radius blurRadius, stop test as in the reference.
 * vel receives the guided, divergence-free field; *iterations = the loop index at exit (what the reference prints).  The optional grids and
 * the solver settings are passed to every inner solvePressure as the plugin does (precondition = true, enforceCompatibility = useL2Norm = false). */
int mp_pd_fluid_guiding(mp_context* ctx, mp_grid* vel, const mp_grid* velT, mp_grid* pressure, const mp_grid* flags, const mp_grid* weight,
                        int blurRadius, double theta, double tau, double sigma, double epsRel, double epsAbs, int maxIters,
                        const mp_grid* phi, const mp_grid* perCellCorr, const mp_grid* fractions, const mp_grid* obvel, double gfClamp, double cgMaxIterFac,
                        double cgAccuracy, int preconditioner, int zeroPressureFixing, const mp_grid* curv, double surfTens, int* iterations);

/* ---- the plugin with HOST buffers (what pressure.cpp calls when grids have no device mirror yet):
 * uploads flags/vel(/phi...), runs mp_solve_pressure, downloads vel/pressure(/retRhs).  Optional
 * pointers may be NULL.  Buffers may be pageable or pinned (mp_host_alloc). ---- */
int mp_solve_pressure_host(mp_context* ctx, int prec, int sx, int sy, int sz,
                           void* vel, void* pressure, const int* flags,
                           const void* phi, const void* perCellCorr, const void* fractions, const void* obvel,
                           const void* curv, void* retRhs, const mp_pressure_params* params, mp_solve_info* info);

/* ---- multi-GPU: z-slab sharding of one solve over the GPUs of one box (one process per GPU) ----
 * The host (torch.distributed or anything else) only moves the opaque unique id; halo exchange and
 * the per-iteration scalar all-reduce run over NCCL/NVLink inside the library. */
int mp_dist_unique_id(void* out128 /* 128 bytes */);
int mp_dist_init(mp_context* ctx, int rank, int world, const void* id128, const char* nccl_library_path /* NULL: search */);
int mp_dist_shutdown(mp_context* ctx);
/* slab of rank r: global planes [k0, k1), computed exactly like the library does */
int mp_dist_slab(int sz_global, int rank, int world, int* k0, int* k1);
/* switch the context to slab mode for a global grid of sz_global planes: from now on every 3-D grid of the context
 * holds the rank's owned planes [k0,k1) plus one ghost plane on each side (local sz = k1-k0+2, local plane kl is global
 * plane k0-1+kl; ghost planes outside the domain are ignored).  Inputs (flags, vel, phi ...) are uploaded WITH their ghost
 * planes; outputs are valid on owned planes (and on ghosts after mp_dist_exchange_halo).  PcNone runs the reference's
 * algorithm unchanged (bit-identical results in the float build).  PcMIC / PcMG* are applied block-Jacobi over the slabs
 * (each rank factorises / coarsens its own slab; same solution within the solver tolerance, more iterations). */
int mp_dist_set_domain(mp_context* ctx, int sz_global);
/* refresh the two ghost planes of a slab grid from the neighbouring ranks (NCCL send/recv on the context stream) */
int mp_dist_exchange_halo(mp_context* ctx, mp_grid* g);
/* how the per-iteration exchanges of the CG loop travel: 0 single GPU (none), 1 NCCL send/recv + all-gather,
 * 2 peer memory (NVLink stores into the neighbours' CUDA-IPC mapped arenas + release/acquire flags) */
int mp_dist_exchange_mode(const mp_context* ctx, int* mode);

#ifdef __cplusplus
}
#endif
#endif

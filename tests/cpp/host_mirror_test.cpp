// C++ parity test of the host mirror (include/mantapress.hpp): written like the reference's tools/tests/test_0100_psolve.py and
// test_0110_mgsolve.py, but in C++ against the reference's class and plugin names -- the way a C++ caller such as
// plugin/fluidguiding.cpp:276-335 uses solvePressure.  The checker is the CPU oracle (oracle/libmf_oracle_f32.so, dlopen'ed: TEST
// INFRASTRUCTURE, never linked into the product), run on the same inputs.
//
//   host_mirror_test --no-device    no GPU needed: the library loads, every plugin symbol links, and constructing a FluidSolver without
//                                   a CUDA device throws Manta::Error (there is no CPU fallback)
//   host_mirror_test <liboracle.so> on a B200: solvePressure with PcMIC / PcMGStatic / PcNone on a closed box with a velocity source,
//                                   the smoke-step plugins and the liquid neighbours, each compared with the oracle
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include "../../include/mantapress.hpp"

using namespace Manta;

#if !MIRROR_COMPILE_ONLY
static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { fails++; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)

// oracle entry points (oracle/mf_oracle.c), same argument meaning as the plugins
typedef int (*fn_solve)(long long, int, int, int, const int*, float*, float*, const float*, const float*, const float*, const float*, const float*, float*, double, double,
                        double, int, int, int, int, int, double, int*, double*);
typedef int (*fn_wall)(int, int, int, const int*, float*, const float*);
typedef int (*fn_grav)(int, int, int, const int*, float*, double, double, double, const float*, int, double);
typedef int (*fn_adv)(int, int, int, const int*, const float*, float*, int, int, double, int, int, int, double);
typedef int (*fn_mac)(int, int, int, const int*, float*, int, const float*, int);
typedef int (*fn_ls)(int, int, int, float*, int, int);

static void boxSource(MACGrid& vel, int res, Vec3 value) {      // test_0100_psolve.py:29,:35: a box of constant velocity
	for (int k = 0; k < res; k++) for (int j = 0; j < res; j++) for (int i = 0; i < res; i++)
		if (i > res * .3 && i < res * .7 && j > res * .1 && j < res * .3 && k > res * .3 && k < res * .7) vel(i, j, k) = value;
}
static double maxDiff(const float* a, const float* b, size_t n) { double m = 0; for (size_t q = 0; q < n; q++) m = std::max(m, (double)std::fabs(a[q] - b[q])); return m; }

#endif

int main(int argc, char** argv) {
	if (argc < 2) { std::printf("usage: host_mirror_test --no-device | <liboracle.so>\n"); return 2; }
	if (std::string(argv[1]) == "--no-device") {
		int n = -1;
		if (mp_device_count(&n) != MP_OK) { std::printf("mp_device_count failed\n"); return 1; }
		if (n > 0) { std::printf("a device is present: run the parity mode instead\n"); return 0; }
		try { FluidSolver s(Vec3i(8, 8, 8), 3); std::printf("FAIL: a FluidSolver was created without a CUDA device\n"); return 1; }
		catch (const Error& e) { std::printf("OK: %s\n", e.what()); }
		try { FluidSolver s(Vec3i(8, 8, 2), 2); return 1; } catch (const Error& e) { std::printf("OK: %s\n", e.what()); }
		// take the address of every plugin so that a missing C-ABI symbol is a link error of this test
		void* p[] = { (void*)&computePressureRhs, (void*)&solvePressureSystem, (void*)&correctVelocity, (void*)&solvePressure, (void*)&releaseMG, (void*)&setWallBcs,
		              (void*)&addGravity, (void*)&addGravityNoScale, (void*)&addBuoyancy, (void*)&advectSemiLagrange<Grid<Real> >, (void*)&advectSemiLagrange<MACGrid>,
		              (void*)&extrapolateMACSimple, (void*)&extrapolateLsSimple, (void*)&extrapolateVec3Simple, (void*)&extrapolateMACFromWeight, (void*)&updateFractions, (void*)&setObstacleFlags, (void*)&getLaplacian, (void*)&getCurvature, (void*)&cgSolveDiffusion<Grid<Real> >, (void*)&cgSolveWE,
		              (void*)&markFluidCells, (void*)&gridParticleIndex, (void*)&unionParticleLevelset, (void*)&mapPartsToMAC, (void*)&mapMACToParts, (void*)&flipVelocityUpdate, (void*)&pushOutofObs,
		              (void*)&addForcePvel, (void*)&updateVelocityFromDeltaPos, (void*)&eulerStep, (void*)&setPartType, (void*)&markIsolatedFluidCell };
		{ typedef void (BasicParticleSystem::*Adv)(const FlagGrid&, const MACGrid&, const int, const bool, const bool, const bool, const ParticleDataImpl<int>*, const int);
		  Adv a = &BasicParticleSystem::advectInGrid; if (!a) return 1; }
		std::printf("OK: %d plugins link\n", (int)(sizeof p / sizeof p[0]));
		return 0;
	}
#if MIRROR_COMPILE_ONLY
	std::printf("built with MIRROR_COMPILE_ONLY: only --no-device is available\n"); return 2;
#else
	void* O = dlopen(argv[1], RTLD_NOW);
	if (!O) { std::printf("cannot load the oracle %s: %s\n", argv[1], dlerror()); return 2; }
	fn_solve o_solve = (fn_solve)dlsym(O, "mfo_solve_pressure"); fn_wall o_wall = (fn_wall)dlsym(O, "mfo_set_wall_bcs_obvel"); fn_grav o_grav = (fn_grav)dlsym(O, "mfo_add_gravity");
	fn_adv o_adv = (fn_adv)dlsym(O, "mfo_advect_semi_lagrange"); fn_mac o_mac = (fn_mac)dlsym(O, "mfo_extrapolate_mac_simple"); fn_ls o_ls = (fn_ls)dlsym(O, "mfo_extrapolate_ls_simple");
	if (!o_solve || !o_wall || !o_grav || !o_adv || !o_mac || !o_ls) { std::printf("oracle symbols missing\n"); return 2; }

	const int res = 40;
	const size_t n = (size_t)res * res * res;
	try {
		FluidSolver s(Vec3i(res, res, res), 3);
		FlagGrid flags(&s); MACGrid vel(&s); Grid<Real> pressure(&s);
		flags.initDomain(); flags.fillGrid();
		std::vector<int> f0(flags.hostData(), flags.hostData() + n);

		// ---- test_0100_psolve.py: PcMIC (the default), then the multigrid preconditioners of test_0110_mgsolve.py, then PcNone
		const int pcs[4] = { PcMIC, PcMGDynamic, PcMGStatic, PcNone };
		for (int q = 0; q < 4; q++) {
			const int pc = pcs[q];
			const bool fix = pc == PcMGDynamic || pc == PcMGStatic;
			vel.clear(); boxSource(vel, res, Vec3(0.15, 0.3, 0.21)); setWallBcs(flags, vel);
			std::vector<float> v_o((const float*)vel.hostData(), (const float*)vel.hostData() + 3 * n), p_o(n, 0.f);
			int it_o = -1; double rn_o = -1;
			const int rc = o_solve(0, res, res, res, f0.data(), v_o.data(), p_o.data(), 0, 0, 0, 0, 0, 0, 1e-4, 1e-4, 99, 1, pc, 0, 0, fix, 0., &it_o, &rn_o);
			CHECK(rc == 0, "oracle solve failed");
			solvePressure(vel, pressure, flags, 1e-4, 0, 0, 0, 0, 1e-04, 99, true, pc, false, false, fix);
			const mp_solve_info& info = lastSolveInfo();
			const double dp = maxDiff(pressure.hostData(), p_o.data(), n), dv = maxDiff((const float*)vel.hostData(), v_o.data(), 3 * n);
			std::printf("preconditioner %d: iterations %d (oracle %d), max |dp| %.3g, max |dv| %.3g\n", pc, info.iterations, it_o, dp, dv);
			CHECK(std::abs(info.iterations - it_o) <= 1, "iteration count");
			CHECK(dp <= 1e-4 && dv <= 1e-4, "pressure / velocity differ from the oracle (threshold of test_0100_psolve.py:40-41)");
		}
		releaseMG(&s);

		// ---- a smoke step either side of the projection, and the liquid neighbours: bit-identical to the oracle
		vel.clear(); boxSource(vel, res, Vec3(0.4, -0.3, 0.2));
		std::vector<float> v_o((const float*)vel.hostData(), (const float*)vel.hostData() + 3 * n);
		s.mDt = 0.7;
		addGravity(flags, vel, Vec3(0.1, -0.3, 0.2));                    o_grav(res, res, res, f0.data(), v_o.data(), 0.1, -0.3, 0.2, 0, 1, 0.7);
		setWallBcs(flags, vel);                                          o_wall(res, res, res, f0.data(), v_o.data(), 0);
		{ std::vector<float> vcopy(v_o);
		  advectSemiLagrange(&flags, &vel, &vel, 2);                     o_adv(res, res, res, f0.data(), vcopy.data(), v_o.data(), 1, 2, 1.0, 1, 2, 1, 0.7); }
		extrapolateMACSimple(flags, vel, 3);                             o_mac(res, res, res, f0.data(), v_o.data(), 3, 0, 0);
		CHECK(maxDiff((const float*)vel.hostData(), v_o.data(), 3 * n) == 0.0, "addGravity / setWallBcs / advectSemiLagrange / extrapolateMACSimple not bit-identical");
		LevelsetGrid phi(&s);
		for (int k = 0; k < res; k++) for (int j = 0; j < res; j++) for (int i = 0; i < res; i++) phi(i, j, k) = (Real)(j + 0.5 - 0.3 * res + 0.05 * ((i * 7 + k * 3) % 11));
		std::vector<float> ph_o(phi.hostData(), phi.hostData() + n);
		extrapolateLsSimple(phi, 4, false);                              o_ls(res, res, res, ph_o.data(), 4, 0);
		extrapolateLsSimple(phi, 4, true);                               o_ls(res, res, res, ph_o.data(), 4, 1);
		CHECK(maxDiff(phi.hostData(), ph_o.data(), n) == 0.0, "extrapolateLsSimple not bit-identical");
		CHECK(s.kernelLaunches() > 0, "no kernel was launched");

		// ---- error behaviour: what the reference reports through errMsg arrives as Manta::Error
		bool threw = false;
		try { advectSemiLagrange(&flags, &vel, &phi, 3); } catch (const Error&) { threw = true; }
		CHECK(threw, "order 3 must throw");
		FluidSolver t(Vec3i(res, res, res + 1), 3);
		MACGrid other(&t);
		threw = false;
		try { solvePressure(other, pressure, flags); } catch (const Error& e) { threw = true; std::printf("expected error: %s\n", e.what()); }
		CHECK(threw, "grids of different solvers must be rejected");
	} catch (const Error& e) { std::printf("FAIL: Manta::Error: %s\n", e.what()); return 1; }
	std::printf(fails ? "%d check(s) FAILED\n" : "all checks passed\n", fails);
	return fails ? 1 : 0;
#endif
}

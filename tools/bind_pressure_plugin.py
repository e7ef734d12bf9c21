"""Applies INTEGRATION.md section 2 to a PREPROCESSED copy of the reference's source/plugin/pressure.cpp (never to the reference tree):
the bodies of solvePressure() (pressure.cpp:480-521) and releaseMG() (:252-266) are replaced by calls into libmantapress through the C-ABI
(include/mantapress.h).  Signatures, parameter names, defaults and the generated Python wrappers stay exactly as the reference has them, so
scenes call solvePressure(flags=..., vel=..., pressure=..., ...) unchanged.  Grids stay host objects in this minimal binding: one call
uploads flags / vel (/ phi, fractions, ...), projects on the GPU and downloads vel / pressure (mp_solve_pressure_host).
MANTA_CPU_PRESSURE=1 in the environment keeps the reference's own CPU body (for A/B runs of the same binary).

usage: python tools/bind_pressure_plugin.py <path to pp/source/plugin/pressure.cpp>"""
import re
import sys

PRELUDE = r'''
// ---- libmantapress binding (tools/bind_pressure_plugin.py, INTEGRATION.md section 2) ----
#include "mantapress.h"
#include <cstdlib>
namespace {
mp_context* gMpCtx = nullptr;                      // process singleton, created at the first solve
mp_context* mpCtx() {
	if (!gMpCtx && mp_context_create(0, &gMpCtx) != MP_OK) errMsg(mp_last_error());
	return gMpCtx;
}
// Grid<T>::mData is protected and the const operator[] returns by value (grid.h:135-137): the non-const accessor of the same object gives the address
template <class G> void* rawOf(const G& g) { return (void*)&const_cast<G&>(g)[0]; }
bool mpUseCpu() { static const bool cpu = getenv("MANTA_CPU_PRESSURE") && atoi(getenv("MANTA_CPU_PRESSURE")) != 0; return cpu; }
}
'''

SOLVE_BODY = r'''
	if (!mpUseCpu()) {
		const Vec3i s_ = flags.getSize();
		mp_pressure_params p_; mp_pressure_params_default(&p_);
		p_.cgAccuracy = cgAccuracy; p_.gfClamp = gfClamp; p_.cgMaxIterFac = cgMaxIterFac; p_.precondition = precondition; p_.preconditioner = preconditioner;
		p_.enforceCompatibility = enforceCompatibility; p_.useL2Norm = useL2Norm; p_.zeroPressureFixing = zeroPressureFixing; p_.surfTens = surfTens;
		mp_solve_info info_;
		// Grid<T>::mData is the raw x-fastest array the ABI expects (grid.h:70); Vec3 is 3 packed Reals (vectorbase.h:199-213)
		const int rc_ = mp_solve_pressure_host(mpCtx(), (int)sizeof(Real), s_.x, s_.y, flags.is3D() ? s_.z : 1,
			rawOf(vel), rawOf(pressure), (const int*)rawOf(flags), phi ? rawOf(*phi) : 0, perCellCorr ? rawOf(*perCellCorr) : 0,
			fractions ? rawOf(*fractions) : 0, obvel ? rawOf(*obvel) : 0, curv ? rawOf(*curv) : 0, retRhs ? rawOf(*retRhs) : 0, &p_, &info_);
		if (rc_ != MP_OK) errMsg(mp_last_error());          // Manta::Error -> RuntimeError (general.h:42-57, pclass.cpp:57-61)
		debMsg("FluidSolver::solvePressure (libmantapress) iterations:" << info_.iterations << ", residual norm: " << info_.resNorm
			<< ", device ms " << info_.msTotal << " (H2D " << info_.msH2D << ", D2H " << info_.msD2H << ")", 2);
		return;
	}
'''

RELEASE_BODY = r'''
	if (!mpUseCpu()) { if (gMpCtx) mp_release_mg(gMpCtx); return; }
'''


def body_span(src, start):
    """(index of the opening brace, index after the matching closing brace) of the function whose definition starts at `start`"""
    i = src.index("{", src.index(")", start))
    # the parameter list may contain parentheses: find the brace that follows the matching ')'
    depth, j = 0, src.index("(", start)
    while True:
        if src[j] == "(":
            depth += 1
        elif src[j] == ")":
            depth -= 1
            if depth == 0:
                break
        j += 1
    i = src.index("{", j)
    depth, k = 0, i
    while True:
        if src[k] == "{":
            depth += 1
        elif src[k] == "}":
            depth -= 1
            if depth == 0:
                return i, k + 1
        k += 1


def main(path):
    src = open(path).read()
    if "libmantapress binding" in src:
        print("already bound:", path)
        return
    m_rel = re.search(r"^void releaseMG\s*\(", src, re.M)
    m_sol = re.search(r"^void solvePressure\s*\(", src, re.M)
    assert m_rel and m_sol, "solvePressure / releaseMG not found"
    # insert from the back so that earlier offsets stay valid
    i, _ = body_span(src, m_sol.start())
    src = src[:i + 1] + SOLVE_BODY + src[i + 1:]
    i, _ = body_span(src, m_rel.start())
    src = src[:i + 1] + RELEASE_BODY + src[i + 1:]
    # the prelude goes in front of releaseMG's file-level neighbours (after the includes / namespace opening that precede it)
    k = src.rfind("\n", 0, src.index("static std::map<FluidSolver*, GridMg*> gMapMG;"))
    src = src[:k] + "\n" + PRELUDE + src[k:]
    open(path, "w").write(src)
    print("bound:", path)


if __name__ == "__main__":
    main(sys.argv[1])

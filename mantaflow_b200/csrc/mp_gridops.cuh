// Element-wise Grid<T> arithmetic of grid.cpp:212-284 / grid.h:472-480 (setConst, addConst, multConst, add, sub, mult, addScaled, clamp, stomp,
// safeDivide), shared by the CUDA kernel of mp_api.cu and the host emulation (tests/emul/particles_emul.cpp, test infrastructure).
// A Vec3 grid is n * 3 scalars; the constant of element e is c[e % comps].  Every operation is one rounding per element, as in the reference.
#pragma once
#include "mp_common.cuh"

#ifndef MP_HD
#ifdef __CUDACC__
#define MP_HD __host__ __device__ __forceinline__
#else
#define MP_HD inline
#endif
#endif

namespace gridops {
template <typename T>
MP_HD T apply(int op, T me, T other, T c0, T c1) {      // c0: the constant / factor / clamp minimum / threshold, c1: clamp maximum
	switch (op) {
		case MP_OP_SET_CONST:  return c0;                                   // knGridSetConstReal grid.cpp:215
		case MP_OP_ADD_CONST:  return me + c0;                              // knGridAddConstReal :216
		case MP_OP_MULT_CONST: return me * c0;                              // knGridMultConst :217
		case MP_OP_ADD:        return me + other;                           // gridAdd grid.h:472
		case MP_OP_SUB:        return me - other;                           // gridSub :473
		case MP_OP_MULT:       return me * other;                           // gridMult :474
		case MP_OP_ADD_SCALED: return me + c0 * other;                      // gridScaledAdd :478
		case MP_OP_CLAMP:      return me < c0 ? c0 : (me > c1 ? c1 : me);   // knGridClamp grid.cpp:222, clamp general.h:137-141
		case MP_OP_STOMP:      return me < c0 ? (T)0 : me;                  // knGridStomp :224-226
		case MP_OP_SAFE_DIVIDE: return other ? (me / other) : me;           // knGridSafeDiv :219, safeDivide general.h:148-151
	}
	return me;
}
template <typename T> struct Op {
	T* me; const T* other; int op, comps; T c0[3], c1[3];
	MP_HD void operator()(IndexInt e) const { const int c = comps == 3 ? (int)(e % 3) : 0; me[e] = apply<T>(op, me[e], other ? other[e] : (T)0, c0[c], c1[c]); }
};
}  // namespace gridops

// Host build of the serial coarse-vertex selection that libmantapress uses for the order-dependent levels of GridMg::setA
// (mantaflow_b200/csrc/mp_mg_coarsen.h), for tests/test_mg_coarsen_host.py.
#include "../../mantaflow_b200/csrc/mp_mg_coarsen.h"

extern "C" int mg_coarsen_level(int fsx, int fsy, int fsz, int csx, int csy, int csz, int is3D, const signed char* tf, signed char* tc)
{
	std::vector<signed char> f(tf, tf + (size_t)fsx * fsy * fsz), c((size_t)csx * csy * csz);
	mgcoarsen::selectCoarseVertices(mgcoarsen::Dim3i{ fsx, fsy, fsz }, mgcoarsen::Dim3i{ csx, csy, csz }, is3D != 0, f, c);
	for (size_t i = 0; i < c.size(); i++) tc[i] = c[i];
	return 0;
}

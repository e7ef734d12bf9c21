// Device-resident scalar state of one GridCg (conjugategrad.h:108-113): Real scalars, int counters.
#pragma once
template <typename Real> struct CgScal {
	Real sigma, alpha, beta, dp, resNorm, accuracy;
	int iterations, done, diverged, useL2;
};
struct CgScalHost { double sigma, resNorm; int iterations, done, diverged, pad; };

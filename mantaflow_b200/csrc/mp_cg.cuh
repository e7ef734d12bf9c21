// Device-resident scalar state of one GridCg (conjugategrad.h:108-113): Real scalars, int counters.
#pragma once
template <typename Real> struct CgScal {
	Real sigma, alpha, beta, dp, resNorm, accuracy;
	int iterations, done, diverged, useL2;
	int xPending;      // fused PcNone loop: x += alpha s of the last iteration has not been applied yet (the next matvec or the flush does it)
};
struct CgScalHost { double sigma, resNorm; int iterations, done, diverged, pad; };

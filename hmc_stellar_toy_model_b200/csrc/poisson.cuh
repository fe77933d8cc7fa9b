// Counter-based Poisson sampling shared by the mock-data kernels (mock_kernels.cu, big_tile.cuh).
//
// Pixel `idx` of a batch draws from Philox4x32-10 with key = seed and counter (idx_lo, attempt, idx_hi, 3): the data
// depend neither on the launch geometry nor on how fields / strips are sharded.  Algorithm: the one NumPy's legacy
// generator uses for np.random.poisson (utils.poisson_realization, utils.py:488-496, calls it once per pixel) --
// multiplication method for lambda < 10, Hoermann's PTRS transformed rejection for lambda >= 10 -- restated from the
// published description (NumPy is an un-vendored, unpinned dependency of the reference).
#pragma once
#include "common.cuh"

namespace srhmc {

// two uniforms in (0,1) for (pixel, attempt)
__device__ __forceinline__ void poisson_uniforms(uint64_t seed, uint64_t idx, uint32_t attempt, double& u, double& v) {
    uint32_t r[4];
    Philox::block(seed, (uint32_t)idx, attempt, (uint32_t)(idx >> 32), 3u, r);
    u = u01(r[0], r[1]);
    v = u01(r[2], r[3]);
}

// One Poisson variate.  lambda <= 0 (or NaN) gives 0, like a zero-rate pixel.
static __device__ __noinline__ double poisson_draw(double lam, uint64_t seed, uint64_t idx) {
    if (!(lam > 0.0)) return 0.0;
    if (lam < 10.0) {
        // multiplication method: count uniforms until their product drops below exp(-lambda)
        const double enlam = exp(-lam);
        double prod = 1.0;
        int k = 0;
        for (uint32_t attempt = 0;; ++attempt) {
            double u, v;
            poisson_uniforms(seed, idx, attempt, u, v);
            prod *= u;
            if (!(prod > enlam)) return (double)k;
            ++k;
            prod *= v;
            if (!(prod > enlam)) return (double)k;
            ++k;
        }
    }
    // PTRS
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (uint32_t attempt = 0; attempt < 4096u; ++attempt) {
        double U, V;
        poisson_uniforms(seed, idx, attempt, U, V);
        U -= 0.5;
        const double us = 0.5 - fabs(U);
        const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
        if ((us >= 0.07) && (V <= vr)) return k;
        if ((k < 0.0) || ((us < 0.013) && (V > us))) continue;
        if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + k * loglam - lgamma(k + 1.0))) return k;
    }
    return floor(lam);  // unreachable in practice (acceptance > 0.9 per attempt)
}

}  // namespace srhmc

// FP64 exp / log tuned for the pixel loops: no special-case branches, coefficients read straight from the
// constant bank (no per-use register moves), log through a 64-entry reciprocal table.
//
// exp_neg(x), x <= 0:   n = rint(x log2 e), r = x - n ln2 (two-term Cody-Waite), degree-12 Taylor polynomial on
//                       |r| <= ln2/2 (truncation 1.7e-16), result scaled by 2^n through the exponent field.
// log_pos(x), x > 0 normal: x = 2^e m, m in [1,2); k = top 6 mantissa bits; table holds rc_k ~ 1/c_k (c_k = bin centre)
//                       and lc_k = -ln(rc_k) to long-double accuracy, so with r = fma(m, rc_k, -1) (exact to 2^-60,
//                       |r| <= 2^-7) the identity ln x = e ln2 + lc_k + log1p(r) holds exactly; log1p by the
//                       degree-7 alternating series (truncation r^8/8 <= 2^-59).  Max error ~2e-16 absolute + 1 ulp.
#pragma once
#include <cuda_runtime.h>

namespace srhmc {

constexpr int kLogTableSize = 64;

struct FastMathConst {
    double expc[13];   // 1/k!
    double log1p_c[10]; // (-1)^(k+1)/k, k = 1..9 at [1..9]
};

static __constant__ FastMathConst kFM = {
    {1.0, 1.0, 0.5, 0.16666666666666666, 0.041666666666666664, 0.008333333333333333, 0.001388888888888889,
     0.0001984126984126984, 2.48015873015873e-05, 2.7557319223985893e-06, 2.755731922398589e-07,
     2.505210838544172e-08, 2.08767569878681e-09},
    {0.0, 1.0, -0.5, 0.3333333333333333, -0.25, 0.2, -0.16666666666666666, 0.14285714285714285, -0.125,
     0.1111111111111111}};

__device__ __forceinline__ double exp_neg(double x) {
    x = fmax(x, -700.0);
    const double magic = 6755399441055744.0;  // 1.5 * 2^52: round-to-nearest integer in the low word
    const double t = fma(x, 1.4426950408889634, magic);
    const int n = __double2loint(t);
    const double nf = t - magic;
    double r = fma(nf, -6.93147180369123816490e-01, x);
    r = fma(nf, -1.90821492927058770002e-10, r);
    double p = kFM.expc[12];
#pragma unroll
    for (int k = 11; k >= 0; --k) p = fma(p, r, kFM.expc[k]);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// tab: kLogTableSize entries of (rc_k, lc_k), in shared memory
__device__ __forceinline__ double log_pos(double x, const double2* __restrict__ tab) {
    const int hi = __double2hiint(x);
    const int e = (hi >> 20) - 1023;
    const int k = (hi >> 14) & (kLogTableSize - 1);
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(x));
    const double2 tk = tab[k];
    const double r = fma(m, tk.x, -1.0);
    double p = kFM.log1p_c[7];
#pragma unroll
    for (int j = 6; j >= 1; --j) p = fma(p, r, kFM.log1p_c[j]);
    // e as double without a conversion instruction: (2^52 + 2^31 + e) - (2^52 + 2^31)
    const double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;
    return fma(ed, 0.6931471805599453, fma(p, r, tk.y));
}

}  // namespace srhmc

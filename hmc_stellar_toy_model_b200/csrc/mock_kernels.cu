// Device-side mock data (SURVEY.md 8f "next" row 3): the input side of every configuration.
//
//   model image   Lambda = B + sum_s f_s PSF_s        base_class.gen_model / gen_mock_data, sampler_RHMC.py:77-117
//                                                      (lightsource_gym.gen_mock_data, samplers.py:44-67)
//   Poisson data  D_ij ~ Poisson(Lambda_ij)            utils.poisson_realization, utils.py:488-496 (one np.random.poisson
//                                                      call per pixel in a Python double loop)
//
// The reference draws from NumPy's global legacy stream, which no parallel generator can reproduce; the device
// generator is counter based instead: pixel `idx` of the batch draws from Philox4x32-10 with key = seed and counter
// (idx_lo, attempt, idx_hi, 3), so the data depend neither on the launch geometry nor on how fields / strips are
// sharded over GPUs.  The sampling ALGORITHM is the one NumPy's legacy generator uses (numpy/random/src/legacy,
// un-vendored third-party code restated from its published description): multiplication method for lambda < 10,
// Hoermann's PTRS transformed rejection (1993) for lambda >= 10; oracle/stellar_oracle.py restates both with the same
// Philox counters, so the test-suite checks the device data value for value.
#include <math_constants.h>

#include <algorithm>

#include "kernels_api.h"
#include "poisson.cuh"

namespace srhmc {

namespace {

// One CTA per (field, 16-row band): stars in chunks of kStarChunk with their separable Gaussian factors in shared memory,
// summed per pixel in star order (the reference adds one star image at a time, sampler_RHMC.py:88-90).
constexpr int kStarChunk = 32;
constexpr int kBandRows = 16;
constexpr int kMaxCols = 128;  // columns handled per CTA pass

__global__ void __launch_bounds__(256) model_kernel(const FieldParams P, int n_fields, const double* __restrict__ q,
                                                    const int* __restrict__ nstars, const double* __restrict__ background,
                                                    double* __restrict__ out) {
    __shared__ double ex[kStarChunk][kBandRows];
    __shared__ double fey[kStarChunk][kMaxCols];
    const int bands = (P.R + kBandRows - 1) / kBandRows;
    const int field = blockIdx.x / bands, band = blockIdx.x % bands;
    if (field >= n_fields) return;
    const int N = nstars ? nstars[field] : P.Nmax;
    const int i0 = band * kBandRows, nr = min(kBandRows, P.R - i0);
    const double* qf = q + (size_t)field * 3 * P.Nmax;
    double* img = out + (size_t)field * P.R * P.C;
    const double* bg = background ? background + (size_t)field * P.R * P.C : nullptr;
    for (int c0 = 0; c0 < P.C; c0 += kMaxCols) {
        const int nc = min(kMaxCols, P.C - c0);
        // pixels of this pass owned by the thread: p = tid + 256 m  (row = p / nc, col = p % nc)
        constexpr int kMaxOwn = (kBandRows * kMaxCols) / 256;
        double acc[kMaxOwn];
#pragma unroll
        for (int m = 0; m < kMaxOwn; ++m) {
            const int pix = threadIdx.x + 256 * m;
            acc[m] = (pix < nr * nc) ? (bg ? bg[(size_t)(i0 + pix / nc) * P.C + c0 + pix % nc] : P.B) : 0.0;
        }
        for (int s0 = 0; s0 < N; s0 += kStarChunk) {
            const int ns = min(kStarChunk, N - s0);
            __syncthreads();
            for (int t = threadIdx.x; t < ns * (nr + nc); t += 256) {
                const int s = t / (nr + nc), e = t % (nr + nc);
                const double f = qf[3 * (s0 + s)], x = qf[3 * (s0 + s) + 1], y = qf[3 * (s0 + s) + 2];
                if (e < nr) {
                    const double d = ((double)(i0 + e) + 0.5) - x;
                    const bool in = P.rad == 0 || (abs(i0 + e - (int)fmin(fmax(floor(x), 0.0), (double)(P.R - 1))) <= P.rad);
                    ex[s][e] = in ? exp(-(d * d) * P.inv2s2) : 0.0;
                } else {
                    const int j = c0 + e - nr;
                    const double d = ((double)j + 0.5) - y;
                    const bool in = P.rad == 0 || (abs(j - (int)fmin(fmax(floor(y), 0.0), (double)(P.C - 1))) <= P.rad);
                    fey[s][e - nr] = in ? f * (exp(-(d * d) * P.inv2s2) * P.norm) : 0.0;
                }
            }
            __syncthreads();
#pragma unroll
            for (int m = 0; m < kMaxOwn; ++m) {
                const int pix = threadIdx.x + 256 * m;
                if (pix < nr * nc) {
                    const int r = pix / nc, c = pix % nc;
                    double a = acc[m];
                    for (int s = 0; s < ns; ++s) a = fma(ex[s][r], fey[s][c], a);
                    acc[m] = a;
                }
            }
        }
#pragma unroll
        for (int m = 0; m < kMaxOwn; ++m) {
            const int pix = threadIdx.x + 256 * m;
            if (pix < nr * nc) img[(size_t)(i0 + pix / nc) * P.C + c0 + pix % nc] = acc[m];
        }
    }
}

__global__ void poisson_kernel(const double* __restrict__ lam, double* __restrict__ D, size_t n, unsigned long long seed,
                               unsigned long long index_base) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        D[i] = poisson_draw(lam[i], seed, index_base + i);
}

}  // namespace

int model_launch(cudaStream_t stream, const FieldParams& P, int n_fields, const double* q, const int* nstars,
                 const double* background, double* out) {
    const int bands = (P.R + kBandRows - 1) / kBandRows;
    model_kernel<<<n_fields * bands, 256, 0, stream>>>(P, n_fields, q, nstars, background, out);
    return (int)cudaGetLastError();
}

int poisson_launch(cudaStream_t stream, const double* lam, double* D, size_t n, unsigned long long seed,
                   unsigned long long index_base) {
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, 148 * 16));
    poisson_kernel<<<blocks, 256, 0, stream>>>(lam, D, n, seed, index_base);
    return (int)cudaGetLastError();
}

}  // namespace srhmc

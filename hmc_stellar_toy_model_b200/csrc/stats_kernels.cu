// On-device chain statistics (SURVEY.md 8f "next" row 4): split-chain Gelman-Rubin R and effective sample size over
// thousands of chains, so that a batch (BASELINE configs[1], [3]) can return summaries instead of its full chains.
//
// Reference: utils.convergence_stats / utils.variogram, utils.py:86-188 (broken under Python 3 upstream: `n = L_chain/2`
// at utils.py:111 is an integer division in the reference's Python 2).  Its quirks are behaviour and are kept:
//   * W is the mean of the within-chain standard DEVIATIONS (np.std, utils.py:120), not variances;
//   * the autocorrelation sum stops at the first odd t with rho_{t+1} + rho_{t+2} < 0 and sums lags 1..t.
//
// One CTA per (group of chains, variable).  Chain layout X[chain][row][d] (the layout of q_chain).  All reductions are
// fixed-order (block_sum), so the numbers are reproducible run to run.
#include <algorithm>

#include "kernels_api.h"

namespace srhmc {

namespace {

struct StatsGeom {
    long long rows;   // iterations stored per chain
    int d;            // variables per row
    int cpg;          // chains per group
    int thin, warm;
    int n;            // samples per split chain
};

// sample k of split chain j (j = 2 * chain + half) of group g, variable v
__device__ __forceinline__ const double* sample_ptr(const double* X, const StatsGeom& G, int g, int v, int j) {
    const long long chain = (long long)g * G.cpg + (j >> 1);
    const long long first = (long long)G.warm + (long long)G.thin * ((j & 1) ? G.n : 0);
    return X + (chain * G.rows + first) * G.d + v;
}

__global__ void __launch_bounds__(256) conv_stats_kernel(const double* __restrict__ X, const StatsGeom G, double* __restrict__ means,
                                                         double* __restrict__ R_out, double* __restrict__ neff_out) {
    __shared__ double red[2 * 32];
    const int g = blockIdx.x / G.d, v = blockIdx.x % G.d;
    const int m = 2 * G.cpg, n = G.n;
    const long long stride = (long long)G.thin * G.d;
    double* mu = means + (size_t)blockIdx.x * m;
    // within-chain mean and sample standard deviation of every split chain (np.mean / np.std(ddof=1), utils.py:118-126)
    double part[2] = {0.0, 0.0};
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const double* x = sample_ptr(X, G, g, v, j);
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += x[k * stride];
        const double mean = s / (double)n;
        double ss = 0.0;
        for (int k = 0; k < n; ++k) {
            const double dlt = x[k * stride] - mean;
            ss += dlt * dlt;
        }
        mu[j] = mean;
        part[0] += sqrt(ss / (double)(n - 1));
        part[1] += mean;
    }
    block_sum<2>(part, red);
    const double W = part[0] / (double)m, mean_all = part[1] / (double)m;
    __syncthreads();
    double b[1] = {0.0};
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const double dlt = mu[j] - mean_all;
        b[0] += dlt * dlt;
    }
    block_sum<1>(b, red);
    const double B = b[0] * (double)n / (double)(m - 1);
    const double var = W * (double)(n - 1) / (double)n + B / (double)n;
    const double R = sqrt(var / W);

    // variogram at lag t (utils.py:170-188), valid in every thread
    auto variogram = [&](int t) {
        double a[1] = {0.0};
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            const double* x = sample_ptr(X, G, g, v, j);
            double s = 0.0;
            for (int k = 0; k + t < n; ++k) {
                const double dlt = x[(k + t) * stride] - x[k * stride];
                s += dlt * dlt;
            }
            a[0] += s;
        }
        block_sum<1>(a, red);
        return a[0] / ((double)m * (double)(n - t));
    };
    // effective sample size (utils.py:137-165); every thread follows the same control flow
    const double rho1 = 1.0 - variogram(1) / (2.0 * var);
    const double rho2 = 1.0 - variogram(2) / (2.0 * var);
    double sum_rho = 0.0;
    if (!(rho1 < 5e-2)) {
        // rho_t list of the reference: index i holds lag i + 1.  Loop invariant: prefix = sum(rho_t[:t]), cur = rho_t[t].
        double prefix = rho1, cur = rho2;
        int t = 1;
        while (t < n - 2) {
            const double next = 1.0 - variogram(t + 2) / (2.0 * var);  // appended as rho_t[t + 1]
            if ((t & 1) && (cur + next) < 0.0) break;
            prefix += cur;
            cur = next;
            ++t;
        }
        sum_rho = prefix < 0.0 ? 0.0 : prefix;
    }
    if (threadIdx.x == 0) {
        R_out[blockIdx.x] = R;
        neff_out[blockIdx.x] = (double)m * (double)n / (1.0 + 2.0 * sum_rho);
    }
}

}  // namespace

// X [n_groups * cpg chains][rows][d] on the device; means: scratch of n_groups * d * 2 * cpg doubles; R, n_eff [n_groups, d].
int conv_stats_launch(cudaStream_t stream, const double* X, long long rows, int d, int n_groups, int cpg, int thin, int warm,
                      double* means, double* R, double* neff) {
    StatsGeom G;
    G.rows = rows; G.d = d; G.cpg = cpg; G.thin = thin; G.warm = warm;
    const long long Lw = rows - warm;
    const long long Lc = Lw > 0 ? (Lw + thin - 1) / thin : 0;  // len(chain[warm:][::thin])
    G.n = (int)(Lc / 2);                                       // utils.py:111 (Python-2 integer division)
    conv_stats_kernel<<<n_groups * d, 256, 0, stream>>>(X, G, means, R, neff);
    return (int)cudaGetLastError();
}

}  // namespace srhmc

// Warp-resident one-star chain kernel: instantiation, occupancy plan and launcher.
#include <algorithm>

#include "chain_kernel.cuh"
#include "kernels_api.h"

namespace srhmc {

static size_t chain_smem_bytes(const FieldParams& P, int lpc) {
    const size_t gpw = 32 / lpc;
    return gpw * ((size_t)P.R * kChainCS * sizeof(double) + (size_t)P.R * sizeof(double2));
}

int chain_kernel_configure(const FieldParams& P, ChainLaunchPlan& plan) {
    plan.lpc = 16;
    plan.smem = chain_smem_bytes(P, plan.lpc);
    cudaError_t e = cudaFuncSetAttribute(chain_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(chain_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, chain_kernel<16>, 32, plan.smem);
    if (e != cudaSuccess) return (int)e;
    if (nb < 1) return (int)cudaErrorInvalidConfiguration;
    plan.max_warps_per_sm = nb;
    return 0;
}

// Grid = (warps per SM) x SMs with the warps-per-SM count chosen so that every SM runs the same number of equally
// long rounds (chains of one launch all take the same number of iterations).
static int chain_grid(const ChainLaunchPlan& plan, int n_fields, int sms) {
    const int gpw = 32 / plan.lpc;
    const long long warps = ((long long)n_fields + gpw - 1) / gpw;
    const long long cap = (long long)plan.max_warps_per_sm * sms;
    if (warps <= cap) return (int)warps;
    int best_k = plan.max_warps_per_sm;
    long long best_cost = -1;
    for (int k = plan.max_warps_per_sm; k >= std::max(1, plan.max_warps_per_sm / 2); --k) {
        const long long rounds = (warps + (long long)k * sms - 1) / ((long long)k * sms);
        const long long cost = rounds * k;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best_k = k;
        }
    }
    return best_k * sms;
}

int chain_kernel_launch(const FieldParams& P, const LaunchArgs& A, const ChainLaunchPlan& plan, int sms,
                               cudaStream_t stream) {
    const int grid = chain_grid(plan, A.n_fields, sms);
    chain_kernel<16><<<grid, 32, plan.smem, stream>>>(P, A);
    return (int)cudaGetLastError();
}

}  // namespace srhmc

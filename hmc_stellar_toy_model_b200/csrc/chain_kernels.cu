// Warp-resident one-star chain kernel: instantiations, occupancy plan, launcher, and the device-math test hook.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "chain_kernel.cuh"
#include "kernels_api.h"

namespace srhmc {

namespace {

constexpr int kLPC = 8;  // lanes per chain: 4 chains per warp (measured faster than 16 lanes / 2 chains)
constexpr int kWarpsPerBlock = 1;

size_t chain_smem_bytes(const FieldParams& P, int lpc, int nw, size_t elem) {
    const size_t gpw = 32 / lpc;
    const size_t per_warp = gpw * (((size_t)P.R * kChainCS + lpc) * elem + (size_t)P.R * sizeof(double2));
    return kLogTableSize * sizeof(double2) + (size_t)nw * per_warp;
}

template <int LPC, typename DT, int MODE>
int configure_mode(size_t smem, int nw, int& blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(chain_kernel<LPC, DT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(chain_kernel<LPC, DT, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, chain_kernel<LPC, DT, MODE>, 32 * nw, smem);
    if (e != cudaSuccess) return (int)e;
    if (nb < 1) return (int)cudaErrorInvalidConfiguration;
    blocks_per_sm = blocks_per_sm == 0 ? nb : std::min(blocks_per_sm, nb);
    return 0;
}

template <int LPC, typename DT>
int configure_one(const FieldParams& P, int nw, size_t& smem, int& blocks_per_sm) {
    smem = chain_smem_bytes(P, LPC, nw, sizeof(DT));
    blocks_per_sm = 0;
    if (int rc = configure_mode<LPC, DT, MODE_EVAL>(smem, nw, blocks_per_sm)) return rc;
    if (int rc = configure_mode<LPC, DT, MODE_STEP>(smem, nw, blocks_per_sm)) return rc;
    if (int rc = configure_mode<LPC, DT, MODE_RUN>(smem, nw, blocks_per_sm)) return rc;
    return configure_mode<LPC, DT, MODE_SINGLE>(smem, nw, blocks_per_sm);
}

template <int LPC, typename DT>
void launch_mode(int grid, int threads, size_t smem, cudaStream_t stream, const FieldParams& P, const LaunchArgs& A) {
    switch (A.mode) {
        case MODE_EVAL: chain_kernel<LPC, DT, MODE_EVAL><<<grid, threads, smem, stream>>>(P, A); break;
        case MODE_STEP: chain_kernel<LPC, DT, MODE_STEP><<<grid, threads, smem, stream>>>(P, A); break;
        case MODE_SINGLE: chain_kernel<LPC, DT, MODE_SINGLE><<<grid, threads, smem, stream>>>(P, A); break;
        default: chain_kernel<LPC, DT, MODE_RUN><<<grid, threads, smem, stream>>>(P, A); break;
    }
}

// Grid = (blocks per SM) x SMs with the per-SM count chosen so that every SM runs the same number of equally long
// rounds (all chains of one launch run the same number of iterations).
int balanced_grid(int max_blocks_per_sm, long long blocks_needed, int sms) {
    if (const char* env = std::getenv("SRHMC_CHAIN_BLOCKS_PER_SM")) {  // tuning / experiments
        const int k = std::atoi(env);
        if (k >= 1) return (int)std::min<long long>(blocks_needed, (long long)std::min(k, max_blocks_per_sm) * sms);
    }
    const long long cap = (long long)max_blocks_per_sm * sms;
    if (blocks_needed <= cap) return (int)std::max<long long>(1, blocks_needed);
    // k resident warps per SM run ceil(needed / (k sms)) rounds; a round costs ~k (throughput-bound) but never less
    // than ~kSat/2 (latency-bound), so prefer the largest k among near-ties.
    const int k_lo = std::max(1, (3 * max_blocks_per_sm) / 4);
    int best_k = max_blocks_per_sm;
    double best_cost = -1.0;
    for (int k = max_blocks_per_sm; k >= k_lo; --k) {
        const long long rounds = (blocks_needed + (long long)k * sms - 1) / ((long long)k * sms);
        const double cost = (double)rounds * k;
        if (best_cost < 0 || cost < best_cost * 0.97) {
            best_cost = cost;
            best_k = k;
        }
    }
    return best_k * sms;
}

__global__ void math_test_kernel(int which, const double* x, double* y, int n, const double2* log_table) {
    __shared__ double2 tab[kLogTableSize];
    for (int i = threadIdx.x; i < kLogTableSize; i += blockDim.x) tab[i] = log_table[i];
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double v = x[i];
        y[i] = which == 0 ? exp_neg(v) : (which == 1 ? log_pos(v, tab) : rcp_fast(v));
    }
}

// exact-count image check + conversion (1 flag word: set when a pixel is not a uint32-representable integer)
__global__ void to_u32_kernel(const double* src, unsigned int* dst, size_t n, int* not_exact) {
    int bad = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = src[i];
        const bool ok = (v >= 0.0) && (v <= 4294967295.0) && (v == floor(v));
        bad |= !ok;
        dst[i] = ok ? (unsigned int)v : 0u;
    }
    if (bad) atomicOr(not_exact, 1);
}

}  // namespace

void fill_log_table(double* host_table /* [kLogTableSize*2] */) {
    for (int k = 0; k < kLogTableSize; ++k) {
        const long double c = 1.0L + ((long double)k + 0.5L) / (long double)kLogTableSize;
        const double rc = (double)(1.0L / c);
        host_table[2 * k] = rc;
        host_table[2 * k + 1] = (double)(-logl((long double)rc));
    }
}

int chain_kernel_configure(const FieldParams& P, ChainLaunchPlan& plan) {
    plan.lpc = kLPC;
    if (const char* env = std::getenv("SRHMC_CHAIN_LPC")) {  // experiments: 16 lanes per chain (2 chains per warp)
        if (std::atoi(env) == 16) plan.lpc = 16;
    }
    plan.nw = kWarpsPerBlock;
    if (plan.lpc == 8) {
        int rc = configure_one<8, double>(P, plan.nw, plan.smem_f64, plan.blocks_per_sm_f64);
        if (rc != 0) return rc;
        return configure_one<8, unsigned int>(P, plan.nw, plan.smem_u32, plan.blocks_per_sm_u32);
    }
    int rc = configure_one<16, double>(P, plan.nw, plan.smem_f64, plan.blocks_per_sm_f64);
    if (rc != 0) return rc;
    rc = configure_one<16, unsigned int>(P, plan.nw, plan.smem_u32, plan.blocks_per_sm_u32);
    return rc;
}

int chain_kernel_launch(const FieldParams& P, const LaunchArgs& A, const ChainLaunchPlan& plan, int sms, cudaStream_t stream) {
    const int chains_per_block = plan.nw * (32 / plan.lpc);
    const long long blocks = ((long long)A.n_fields + chains_per_block - 1) / chains_per_block;
    if (A.D_u32 != nullptr) {
        const int grid = balanced_grid(plan.blocks_per_sm_u32, blocks, sms);
        if (plan.lpc == 8)
            launch_mode<8, unsigned int>(grid, 32 * plan.nw, plan.smem_u32, stream, P, A);
        else
            launch_mode<16, unsigned int>(grid, 32 * plan.nw, plan.smem_u32, stream, P, A);
    } else {
        const int grid = balanced_grid(plan.blocks_per_sm_f64, blocks, sms);
        if (plan.lpc == 8)
            launch_mode<8, double>(grid, 32 * plan.nw, plan.smem_f64, stream, P, A);
        else
            launch_mode<16, double>(grid, 32 * plan.nw, plan.smem_f64, stream, P, A);
    }
    return (int)cudaGetLastError();
}

int math_test_launch(cudaStream_t stream, int which, const double* x, double* y, int n, const double* log_table) {
    math_test_kernel<<<std::max(1, std::min(1024, (n + 255) / 256)), 256, 0, stream>>>(
        which, x, y, n, reinterpret_cast<const double2*>(log_table));
    return (int)cudaGetLastError();
}

int to_u32_launch(cudaStream_t stream, const double* src, unsigned int* dst, size_t n, int* not_exact_flag) {
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, 4096));
    to_u32_kernel<<<blocks, 256, 0, stream>>>(src, dst, n, not_exact_flag);
    return (int)cudaGetLastError();
}

}  // namespace srhmc

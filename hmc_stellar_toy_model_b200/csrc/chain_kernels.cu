// Warp-resident one-star chain kernel: instantiations, occupancy plan, launcher, and the device-math test hook.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "chain_kernel.cuh"
#include "kernels_api.h"

namespace srhmc {

namespace {

constexpr int kLPC = kChainLPC;  // lanes per chain: 4 -> 8 chains per warp (measured: 16 lanes 734, 8 lanes 1104, 4 lanes 1210 M star-steps/s)
#ifndef SRHMC_CHAIN_WARPS_PER_BLOCK
#define SRHMC_CHAIN_WARPS_PER_BLOCK 1
#endif
constexpr int kWarpsPerBlock = SRHMC_CHAIN_WARPS_PER_BLOCK;
// Column slots per lane: the full 32-column image, or a 24-column window around the star when the PSF weight of
// every dropped column is below 2^-46 of its peak (sigma <= 1.50 px: |dy| >= 12 px).  SRHMC_CHAIN_WINDOW=0 at run time
// forces the full width (A/B measurements).
constexpr int kSlotsFull = 32 / kLPC;
#ifndef SRHMC_CHAIN_WIN_COLS
#define SRHMC_CHAIN_WIN_COLS 24
#endif
constexpr int kWinCols = SRHMC_CHAIN_WIN_COLS;
constexpr int kSlotsWin = (kWinCols % kLPC == 0 && kWinCols / kLPC < kSlotsFull) ? kWinCols / kLPC : kSlotsFull;

bool use_column_window(const FieldParams& P) {
    if (kSlotsWin == kSlotsFull) return false;
    if (const char* env = std::getenv("SRHMC_CHAIN_WINDOW"))
        if (std::atoi(env) == 0) return false;
    const double half = 0.5 * kWinCols;
    return half * half * P.inv2s2 >= 46.0 * M_LN2;
}

size_t chain_smem_bytes(const FieldParams& P, int lpc, int nw, size_t elem, bool f32 = false) {
    const size_t gpw = 32 / lpc;
    const size_t rowtab = sizeof(double2) + (f32 ? sizeof(float2) : 0);
    const size_t per_warp = gpw * (((size_t)P.R * kChainCS + lpc) * elem + (size_t)P.R * rowtab + kChainTabPad);
    return kLogTableSize * sizeof(double2) + (size_t)nw * per_warp;
}

template <typename DT, int MODE, int NCS, typename PT = double>
int configure_mode(size_t smem, int nw, int& blocks_per_sm) {
    auto fn = chain_kernel<kLPC, DT, MODE, NCS, kChainMaxReg, PT>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, 32 * nw, smem);
    if (e != cudaSuccess) return (int)e;
    if (nb < 1) return (int)cudaErrorInvalidConfiguration;
    blocks_per_sm = blocks_per_sm == 0 ? nb : std::min(blocks_per_sm, nb);
    return 0;
}

template <typename DT, int NCS, typename PT = double>
int configure_slots(size_t smem, int nw, int& blocks_per_sm) {
    if (int rc = configure_mode<DT, MODE_EVAL, NCS, PT>(smem, nw, blocks_per_sm)) return rc;
    if (int rc = configure_mode<DT, MODE_STEP, NCS, PT>(smem, nw, blocks_per_sm)) return rc;
    if (int rc = configure_mode<DT, MODE_RUN, NCS, PT>(smem, nw, blocks_per_sm)) return rc;
    return configure_mode<DT, MODE_SINGLE, NCS, PT>(smem, nw, blocks_per_sm);
}

template <typename DT, typename PT = double>
int configure_one(const FieldParams& P, int nw, size_t& smem, int& blocks_per_sm) {
    smem = chain_smem_bytes(P, kLPC, nw, sizeof(DT), sizeof(PT) == 4);
    blocks_per_sm = 0;
    const int rc = use_column_window(P) ? configure_slots<DT, kSlotsWin, PT>(smem, nw, blocks_per_sm)
                                        : configure_slots<DT, kSlotsFull, PT>(smem, nw, blocks_per_sm);
    if (const char* env = std::getenv("SRHMC_CHAIN_MAX_RESIDENT")) {   // occupancy experiments: cap the resident blocks per SM
        const int k = std::atoi(env);
        if (rc == 0 && k >= 1) blocks_per_sm = std::min(blocks_per_sm, k);
    }
    return rc;
}

template <typename DT, int NCS, typename PT = double>
void launch_slots(int grid, int threads, size_t smem, cudaStream_t stream, const FieldParams& P, const LaunchArgs& A) {
    switch (A.mode) {
        case MODE_EVAL: chain_kernel<kLPC, DT, MODE_EVAL, NCS, kChainMaxReg, PT><<<grid, threads, smem, stream>>>(P, A); break;
        case MODE_STEP: chain_kernel<kLPC, DT, MODE_STEP, NCS, kChainMaxReg, PT><<<grid, threads, smem, stream>>>(P, A); break;
        case MODE_SINGLE: chain_kernel<kLPC, DT, MODE_SINGLE, NCS, kChainMaxReg, PT><<<grid, threads, smem, stream>>>(P, A); break;
        default: chain_kernel<kLPC, DT, MODE_RUN, NCS, kChainMaxReg, PT><<<grid, threads, smem, stream>>>(P, A); break;
    }
}

template <typename DT, typename PT = double>
void launch_mode(int grid, int threads, size_t smem, cudaStream_t stream, const FieldParams& P, const LaunchArgs& A) {
    if (use_column_window(P))
        launch_slots<DT, kSlotsWin, PT>(grid, threads, smem, stream, P, A);
    else
        launch_slots<DT, kSlotsFull, PT>(grid, threads, smem, stream, P, A);
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

// exact-count image check + conversion.  flags bit 0: a pixel is not a uint32-representable integer; bit 1: a pixel
// exceeds 65535 (the uint16 copy is then unusable).
__global__ void to_counts_kernel(const double* src, unsigned int* dst32, unsigned short* dst16, size_t n, int* flags) {
    int bad = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = src[i];
        const bool ok = (v >= 0.0) && (v <= 4294967295.0) && (v == floor(v));
        bad |= ok ? 0 : 1;
        bad |= (ok && v <= 65535.0) ? 0 : 2;
        dst32[i] = ok ? (unsigned int)v : 0u;
        dst16[i] = (ok && v <= 65535.0) ? (unsigned short)v : (unsigned short)0;
    }
    if (bad) atomicOr(flags, bad);
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
    plan.nw = kWarpsPerBlock;
    if (int rc = configure_one<double>(P, plan.nw, plan.smem_f64, plan.blocks_per_sm_f64)) return rc;
    if (int rc = configure_one<unsigned int>(P, plan.nw, plan.smem_u32, plan.blocks_per_sm_u32)) return rc;
    if (int rc = configure_one<unsigned short>(P, plan.nw, plan.smem_u16, plan.blocks_per_sm_u16)) return rc;
    // FP32 pixel arithmetic (precision-32 contexts) on the uint16 count images
    return configure_one<unsigned short, float>(P, plan.nw, plan.smem_u16_f32, plan.blocks_per_sm_u16_f32);
}

// Iteration chunks per chain for a MODE_RUN launch of `groups` warp-sized work items on `warps` resident warps.
// One chunk when everything is resident at once (or the chains are short); otherwise enough chunks that the
// under-filled last round costs a few percent instead of up to half the launch.
// The kernel cuts L iterations into n chunks of ceil(L/n) iterations; every chunk must be non-empty (an empty last chunk
// would wait for a predecessor that, being the real last one, never publishes): largest n' <= n with (n'-1) ceil(L/n') < L.
static int valid_chunks(int n, int L) {
    n = std::max(1, std::min(n, L));
    while (n > 1 && (long long)(n - 1) * (((long long)L + n - 1) / n) >= (long long)L) --n;
    return n;
}

int pick_chunks(long long groups, long long warps, int L) {
    if (const char* env = std::getenv("SRHMC_CHAIN_CHUNKS")) {
        const int c = std::atoi(env);
        if (c >= 1) return valid_chunks(c, L);
    }
    if (groups <= warps || L < 64) return 1;
    return valid_chunks(std::min(16, L / 32), L);
}

long long chain_kernel_resident_warps(const LaunchArgs& A, const ChainLaunchPlan& plan, int sms, int n_fields) {
    const int chains_per_block = plan.nw * (32 / plan.lpc);
    const long long blocks = ((long long)n_fields + chains_per_block - 1) / chains_per_block;
    const bool u16 = A.D_int != nullptr && A.D_int_bytes == 2;
    const int max_k = (u16 && A.pix_f32) ? plan.blocks_per_sm_u16_f32
                                         : (u16 ? plan.blocks_per_sm_u16 : (A.D_int != nullptr ? plan.blocks_per_sm_u32 : plan.blocks_per_sm_f64));
    return std::min<long long>(blocks, (long long)max_k * sms) * plan.nw;
}

int chain_kernel_launch(const FieldParams& P, const LaunchArgs& A_in, const ChainLaunchPlan& plan, int sms, cudaStream_t stream) {
    LaunchArgs A = A_in;
    const int chains_per_block = plan.nw * (32 / plan.lpc);
    const int n_range = (A.field_end > 0 ? A.field_end : A.n_fields) - A.field_begin;
    const long long blocks = ((long long)n_range + chains_per_block - 1) / chains_per_block;
    const bool u16 = A.D_int != nullptr && A.D_int_bytes == 2;
    // 4 lanes per chain: 240 of 255 registers -> 8 resident single-warp blocks per SM (2 per scheduler: the register file is
    // partitioned per scheduler, a third warp would need <= 168 registers and measured slower); residency comes from the
    // occupancy query in chain_kernel_configure.  Register-capped 8-lane builds were measured earlier and are slower
    // (128 registers / 16 warps 1057 M star-steps/s, 112 / 18 warps 892 against 1104 at 168 / 12).
    const bool f32 = u16 && A.pix_f32;
    const int max_k = f32 ? plan.blocks_per_sm_u16_f32
                          : (u16 ? plan.blocks_per_sm_u16 : (A.D_int != nullptr ? plan.blocks_per_sm_u32 : plan.blocks_per_sm_f64));
    int grid = balanced_grid(max_k, blocks, sms);
    if (!(A.mode == MODE_RUN && A.sched_done != nullptr && A.chunk_count > 0)) A.n_chunks = 1;
    if (A.mode == MODE_RUN && A.sched_done != nullptr && A.chunk_count > 0) {
        // one launch of a run that the host has cut along the iteration axis: the chunking is fixed by the caller
        grid = (int)std::min<long long>(blocks, (long long)max_k * sms);
    } else if (A.mode == MODE_RUN && A.sched_done != nullptr) {
        // with the chunked scheduler a partly filled last round costs little, so run at full residency
        const int full = (int)std::min<long long>(blocks, (long long)max_k * sms);
        const int chunks = pick_chunks(blocks, (long long)full * plan.nw, A.niter + 1);
        if (chunks > 1 && !std::getenv("SRHMC_CHAIN_BLOCKS_PER_SM")) grid = full;
        A.n_chunks = pick_chunks(blocks, (long long)grid * plan.nw, A.niter + 1);
    }
    if (f32) {
        launch_mode<unsigned short, float>(grid, 32 * plan.nw, plan.smem_u16_f32, stream, P, A);
    } else if (u16) {
        launch_mode<unsigned short>(grid, 32 * plan.nw, plan.smem_u16, stream, P, A);
    } else if (A.D_int != nullptr) {
        launch_mode<unsigned int>(grid, 32 * plan.nw, plan.smem_u32, stream, P, A);
    } else {
        launch_mode<double>(grid, 32 * plan.nw, plan.smem_f64, stream, P, A);
    }
    return (int)cudaGetLastError();
}

int math_test_launch(cudaStream_t stream, int which, const double* x, double* y, int n, const double* log_table) {
    math_test_kernel<<<std::max(1, std::min(1024, (n + 255) / 256)), 256, 0, stream>>>(
        which, x, y, n, reinterpret_cast<const double2*>(log_table));
    return (int)cudaGetLastError();
}

int to_counts_launch(cudaStream_t stream, const double* src, unsigned int* dst32, unsigned short* dst16, size_t n,
                     int* flags) {
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, 4096));
    to_counts_kernel<<<blocks, 256, 0, stream>>>(src, dst32, dst16, n, flags);
    return (int)cudaGetLastError();
}

}  // namespace srhmc

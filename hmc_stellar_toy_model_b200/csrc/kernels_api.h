// Host-callable launchers of the device kernels; each family is compiled in its own translation unit so the
// in-tree build can run them in parallel.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "common.cuh"

namespace srhmc {

struct ChainLaunchPlan {
    int lpc = 16;
    int max_warps_per_sm = 0;
    size_t smem = 0;
};

// CTA-per-field kernel (field_kernel.cuh); precision 64|32, (mr, mc) in {(2,4), (2,2), (1,2)}
size_t field_layout_total(int precision, const FieldParams& P, bool d_in_smem);
int field_kernel_configure(int precision, int mr, int mc, size_t smem);
int field_kernel_launch(int precision, int mr, int mc, int grid, int threads, size_t smem, cudaStream_t stream,
                        const FieldParams& P, const LaunchArgs& A, double* scratch, int d_in_smem);
int philox_dump_launch(cudaStream_t stream, unsigned long long seed, int n_fields, int L, int Nmax, double* normals,
                       double* lnu);
int convert_image_launch(cudaStream_t stream, const double* src, float* dst, size_t n);

// warp-resident one-star kernel (chain_kernel.cuh), FP64
int chain_kernel_configure(const FieldParams& P, ChainLaunchPlan& plan);
int chain_kernel_launch(const FieldParams& P, const LaunchArgs& A, const ChainLaunchPlan& plan, int sms, cudaStream_t stream);

// FMA-chain roofline microbenchmark
int fma_peak_run(int precision, int sms, double* tflops, float* ms);

}  // namespace srhmc

// lightsource_gym-family kernel instantiations (double pixels only: these samplers are parity paths).
#include "ls_kernel.cuh"
#include "kernels_api.h"

namespace srhmc {

typedef void (*LsKernelFn)(const FieldParams, const LsArgs, double*, int);

static LsKernelFn pick_ls(int mr, int mc) {
    if (mr == 2 && mc == 4) return ls_kernel<double, 2, 4>;
    if (mr == 2 && mc == 2) return ls_kernel<double, 2, 2>;
    return ls_kernel<double, 1, 2>;
}

int ls_kernel_configure(int mr, int mc, size_t smem) {
    return (int)cudaFuncSetAttribute(pick_ls(mr, mc), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

int ls_kernel_launch(int mr, int mc, int grid, int threads, size_t smem, cudaStream_t stream, const FieldParams& P,
                     const LsArgs& A, double* scratch, int dsm) {
    pick_ls(mr, mc)<<<grid, threads, smem, stream>>>(P, A, scratch, dsm);
    return (int)cudaGetLastError();
}

}  // namespace srhmc

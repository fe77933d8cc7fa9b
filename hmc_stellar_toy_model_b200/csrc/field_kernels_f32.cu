// CTA-per-field kernel instantiations, float pixels.
#include "field_kernel.cuh"
#include "kernels_api.h"

namespace srhmc {

typedef void (*FieldKernelFn32)(const FieldParams, const LaunchArgs, double*, int);

static FieldKernelFn32 pick32(int mr, int mc) {
    if (mr == 2 && mc == 4) return field_kernel<float, 2, 4>;
    if (mr == 2 && mc == 2) return field_kernel<float, 2, 2>;
    return field_kernel<float, 1, 2>;
}

size_t field_layout_total_f32(const FieldParams& P, bool dsm) { return make_layout<float>(P, dsm).total; }

int field_kernel_configure_f32(int mr, int mc, size_t smem) {
    return (int)cudaFuncSetAttribute(pick32(mr, mc), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

int field_kernel_launch_f32(int mr, int mc, int grid, int threads, size_t smem, cudaStream_t stream, const FieldParams& P,
                            const LaunchArgs& A, double* scratch, int dsm) {
    pick32(mr, mc)<<<grid, threads, smem, stream>>>(P, A, scratch, dsm);
    return (int)cudaGetLastError();
}

}  // namespace srhmc

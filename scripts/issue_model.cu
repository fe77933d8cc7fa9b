// Issue-slot model of the FP64 pipe on sm_100a (evidence for DESIGN section 5; not part of the library).
//
// Question: does a DFMA only occupy the FP64 pipe for two cycles per scheduler (16 lanes per cycle), leaving the issue
// port free for other instructions in the second cycle -- or does it also hold the scheduler's issue port, so that every
// non-FP64 instruction in the stream takes time away from the FP64 pipe?  The one-star chain kernel issues 45 % non-FP64
// instructions; its attainable FP64-pipe fraction is 1.0 under the first model and 2 f / (2 f + (1 - f)) = 0.71 under the
// second (f = FP64 share of the instruction stream).
//
// Method: every thread runs 8 independent FP64 dependency chains (as srhmc_measure_fma_peak) and, per FP64 instruction,
// K independent instructions of another kind on other registers (IMAD / FFMA / LOP3 chains).  We report FP64 warp-instructions
// per cycle per scheduler for K = 0, 1, 2 at 2 and 16 resident warps per scheduler.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_model scripts/issue_model.cu && ./issue_model
#include <cstdio>
#include <cuda_runtime.h>

enum { OP_DFMA = 0, OP_DMUL = 1, OP_DADD = 2 };
enum { MIX_NONE = 0, MIX_IMAD = 1, MIX_FFMA = 2, MIX_LOP = 3 };

template <int OP, int MIX, int K>
__global__ void __launch_bounds__(256) mix_kernel(double* out, int iters, double a, double b, int ia, float fa) {
    double x[8];
    int y[8];
    float z[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        x[u] = (double)(threadIdx.x + u);
        y[u] = threadIdx.x + u;
        z[u] = (float)(threadIdx.x + u);
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (OP == OP_DFMA) x[u] = fma(x[u], a, b);
                if (OP == OP_DMUL) x[u] = x[u] * a;
                if (OP == OP_DADD) x[u] = x[u] + b;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (MIX == MIX_IMAD) y[u] = y[u] * ia + 12345;
                    if (MIX == MIX_FFMA) z[u] = fmaf(z[u], fa, 0.5f);
                    if (MIX == MIX_LOP) y[u] = (y[u] ^ ia) + (y[u] >> 3);
                }
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += x[u] + (double)y[u] + (double)z[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP, int MIX, int K>
static void run(const char* name, int sms, int warps_per_sm, double mhz, double* out) {
    const int threads = 32, blocks = sms * warps_per_sm, iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        mix_kernel<OP, MIX, K><<<blocks, threads>>>(out, iters, 0.999, 0.001, 3, 0.999f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float t;
        cudaEventElapsedTime(&t, e0, e1);
        if (rep && t < best) best = t;
    }
    const double fp64_per_warp = 64.0 * iters;
    const double cycles = best * 1e-3 * mhz * 1e6;
    const double per_sched = fp64_per_warp * (warps_per_sm / 4.0) / cycles;   // FP64 warp-instructions per cycle per scheduler
    printf("%-28s warps/scheduler %2d  K=%d  %.3f ms  FP64 instr/cycle/scheduler %.3f  (all instr %.3f)\n", name, warps_per_sm / 4, K, best,
           per_sched, per_sched * (1 + K));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs, %.0f MHz (max SM clock: the cycle counts assume the GPU runs at it)\n", p.name, sms, mhz);
    double* out;
    cudaMalloc(&out, (size_t)sms * 64 * 32 * sizeof(double));
    for (int w : {8, 64}) {
        if (w == 8) {
            run<OP_DFMA, MIX_NONE, 0>("DFMA alone", sms, 8, mhz, out);
            run<OP_DMUL, MIX_NONE, 0>("DMUL alone", sms, 8, mhz, out);
            run<OP_DADD, MIX_NONE, 0>("DADD alone", sms, 8, mhz, out);
            run<OP_DFMA, MIX_IMAD, 1>("DFMA + 1 IMAD each", sms, 8, mhz, out);
            run<OP_DFMA, MIX_IMAD, 2>("DFMA + 2 IMAD each", sms, 8, mhz, out);
            run<OP_DFMA, MIX_FFMA, 1>("DFMA + 1 FFMA each", sms, 8, mhz, out);
            run<OP_DFMA, MIX_FFMA, 2>("DFMA + 2 FFMA each", sms, 8, mhz, out);
            run<OP_DFMA, MIX_LOP, 1>("DFMA + (LOP3, SHF, IADD) each", sms, 8, mhz, out);
        } else {
            run<OP_DFMA, MIX_NONE, 0>("DFMA alone", sms, 64, mhz, out);
            run<OP_DMUL, MIX_NONE, 0>("DMUL alone", sms, 64, mhz, out);
            run<OP_DADD, MIX_NONE, 0>("DADD alone", sms, 64, mhz, out);
            run<OP_DFMA, MIX_IMAD, 1>("DFMA + 1 IMAD each", sms, 64, mhz, out);
            run<OP_DFMA, MIX_IMAD, 2>("DFMA + 2 IMAD each", sms, 64, mhz, out);
            run<OP_DFMA, MIX_FFMA, 1>("DFMA + 1 FFMA each", sms, 64, mhz, out);
            run<OP_DFMA, MIX_FFMA, 2>("DFMA + 2 FFMA each", sms, 64, mhz, out);
            run<OP_DFMA, MIX_LOP, 1>("DFMA + (LOP3, SHF, IADD) each", sms, 64, mhz, out);
        }
    }
    cudaFree(out);
    return 0;
}

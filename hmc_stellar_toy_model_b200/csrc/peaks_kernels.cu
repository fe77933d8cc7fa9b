// Gradient-descent leg of lightsource_gym.find_peaks (samplers.py:196-226; SURVEY.md 8f "next" row 2): every seed
// rolls downhill on its own, on a pure-background model, until its potential changes by less than 1e-9 relative or its
// flux drops below f_lim.  Seeds are independent, so one warp owns one seed for its whole descent; the data image sits
// in shared memory (one copy per CTA) and nothing returns to the host in between.
//
// Per position the warp makes ONE pass over the full image (the reference's full-image PSF) that yields both
// V_single (samplers.py:121-127) and the three components of dVdq_single (samplers.py:77-112) through the separable
// Gaussian:  Lambda_ij = B + f n ex_i ey_j,  rho = D/Lambda - 1,
//   g_f = -n sum_j ey_j c0_j,  g_x = -(f/s^2) n sum_j ey_j c1_j,  g_y = -(f/s^2) n sum_j ey_j dy_j c0_j,
//   c0_j = sum_i rho_ij ex_i,  c1_j = sum_i rho_ij ex_i dx_i,     V = -sum(D ln Lambda - Lambda).
// The reference evaluates V at the new position and the gradient at the same position in the next iteration; fusing the
// two passes changes no value.
#include <algorithm>

#include "kernels_api.h"

namespace srhmc {

namespace {

constexpr int kPeakWarps = 4;

__global__ void __launch_bounds__(32 * kPeakWarps) peaks_kernel(const FieldParams P, const double* __restrict__ D, int d_in_smem,
                                                                int n, int nstep, double dt_f_coeff, double dt_xy_coeff,
                                                                double f_lim, double* __restrict__ q, unsigned char* __restrict__ alive,
                                                                int* __restrict__ steps) {
    extern __shared__ __align__(16) unsigned char peaks_smem[];
    const int R = P.R, C = P.C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* tab = reinterpret_cast<double*>(peaks_smem) + (size_t)warp * (2 * R + C);  // ex[R], exdx[R], ey[C]
    double* sD = reinterpret_cast<double*>(peaks_smem) + (size_t)kPeakWarps * (2 * R + C);
    if (d_in_smem) {
        for (int i = threadIdx.x; i < R * C; i += blockDim.x) sD[i] = D[i];
        __syncthreads();
    }
    const double* img = d_in_smem ? sD : D;
    const int seed = blockIdx.x * kPeakWarps + warp;
    if (seed >= n) return;
    double f = q[3 * seed], x = q[3 * seed + 1], y = q[3 * seed + 2];
    double V = 0.0, gf = 0.0, gx = 0.0, gy = 0.0;

    auto pass = [&]() {
        __syncwarp();
        for (int i = lane; i < R; i += 32) {
            const double dx = ((double)i + 0.5) - x;
            const double e = exp(-(dx * dx) * P.inv2s2);
            tab[i] = e;
            tab[R + i] = e * dx;
        }
        for (int j = lane; j < C; j += 32) {
            const double dy = ((double)j + 0.5) - y;
            tab[2 * R + j] = exp(-(dy * dy) * P.inv2s2) * P.norm;
        }
        __syncwarp();
        double sv = 0.0, sf = 0.0, sx = 0.0, sy = 0.0;
        for (int j = lane; j < C; j += 32) {
            const double ey = tab[2 * R + j], fey = f * ey, dy = ((double)j + 0.5) - y;
            double c0 = 0.0, c1 = 0.0;
            for (int i = 0; i < R; ++i) {
                const double lam = fma(tab[i], fey, P.B);
                const double d = img[i * C + j];
                const double rho = d / lam - 1.0;
                c0 = fma(rho, tab[i], c0);
                c1 = fma(rho, tab[R + i], c1);
                sv += d * log(lam) - lam;
            }
            sf = fma(ey, c0, sf);
            sx = fma(ey, c1, sx);
            sy = fma(ey * dy, c0, sy);
        }
        V = -warp_sum(sv);
        gf = -warp_sum(sf);
        gx = -warp_sum(sx) * f * P.inv_s2;
        gy = -warp_sum(sy) * f * P.inv_s2;
    };

    pass();
    double V_prev = V;
    bool live = true;
    int it = 0;
    for (int i = 0; i < nstep; ++i) {
        const double dt_f = f * dt_f_coeff, dt_xy = dt_xy_coeff / f;  // samplers.py:207-208
        f -= gf * dt_f;
        x -= gx * dt_xy;
        y -= gy * dt_xy;
        it = i + 1;
        if (f < f_lim) {  // faint: the seed disappears (samplers.py:215-217)
            live = false;
            break;
        }
        pass();
        if (fabs((V - V_prev) / V_prev) < 1e-9) break;
        V_prev = V;
    }
    if (lane == 0) {
        q[3 * seed] = f; q[3 * seed + 1] = x; q[3 * seed + 2] = y;
        alive[seed] = live ? 1 : 0;
        if (steps) steps[seed] = it;
    }
}

}  // namespace

int peaks_launch(cudaStream_t stream, const FieldParams& P, const double* D, int n, int nstep, double dt_f_coeff,
                 double dt_xy_coeff, double f_lim, double* q, unsigned char* alive, int* steps) {
    const size_t tabs = (size_t)kPeakWarps * (2 * P.R + P.C) * 8, img = (size_t)P.R * P.C * 8;
    const int d_in_smem = tabs + img <= 200 * 1024 ? 1 : 0;
    const size_t smem = tabs + (d_in_smem ? img : 0);
    cudaError_t e = cudaFuncSetAttribute(peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    peaks_kernel<<<(n + kPeakWarps - 1) / kPeakWarps, 32 * kPeakWarps, smem, stream>>>(P, D, d_in_smem, n, nstep, dt_f_coeff,
                                                                                        dt_xy_coeff, f_lim, q, alive, steps);
    return (int)cudaGetLastError();
}

}  // namespace srhmc

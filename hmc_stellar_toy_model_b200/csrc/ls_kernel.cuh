// samplers.lightsource_gym family on the device (reference samplers.py): the older random-trajectory-length HMC /
// RHMC samplers of the toy model, one CTA per field, images and star state resident in shared memory for the
// whole chain.  Built from the crowded-field kernel's pixel machinery (field_kernel.cuh).
//
//   LS_EVAL_HESS  RHMC_efficient_computation (samplers.py:828-927): dV/dq, diagonal d2V/dq2, d3V/dq3 as
//                 residual-weighted PSF reductions (17 separable sums per star), dq/dt, dp/dt, E
//   LS_HMC        HMC_random (samplers.py:460-572): identity mass, per-coordinate dt vector
//   LS_DIAG       RHMC_random_diag (samplers.py:668-825): M(f) = [1/f, f factor1, f factor1], global dt
//   LS_HESS       RHMC_random (samplers.py:930-1105): Hessian-metric explicit scheme
//   LS_TRIAL      the single-star HMC trial inside HMC_find_best_dt (samplers.py:327-370, 395-432) on a model_data
//                 background image
// The reference's quirks are part of the behaviour and are reproduced (hard part 2): the flip
// set is never cleared inside an iteration; the flip branch of the final half step leaves p_tmp untouched; the
// diagonal sampler evaluates dpMpdq with the iteration's INITIAL momentum; RHMC_random's position test
// `(x < 0) or (x < num_rows)` flags every in-image star so its position momenta flip sign every step, and its
// in-place `q_tmp += ...` aliases q_initial so a rejection does not restore the state; the trial HMC kicks all
// three momentum components with the scalar FLUX gradient (dVdq_single's default f_only=True).
#pragma once
#include "field_kernel.cuh"

namespace srhmc {

// ------------------------------------------------------------------------------------------ K7: Hessian reductions
// One warp per star.  Needs c.sL = rho0 = D/Lambda, c.sL2 = 1/Lambda and tables built with scale_f = false.
// With a_i = ex_i, b_j = ey_j (normalised), dx_i = i+.5-x, dy_j = j+.5-y, rho1 = 1-rho0, rho2 = rho0/Lambda,
// rho3 = rho2/Lambda the 17 sums are S{1,2,3}_{mn} = sum rho_k (a^k dx^m)(b^k dy^n).
template <typename T>
__device__ void hess_chunk(const Ctx<T>& c, int k0, int nk, double* d1, double* d2, double* d3) {
    const FieldParams& P = *c.P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int kk = warp; kk < nk; kk += nwarps) {
        const int k = k0 + kk;
        const double f = c.q[3 * k], x = c.q[3 * k + 1], y = c.q[3 * k + 2];
        const short4 sp = *reinterpret_cast<const short4*>(c.span + 4 * kk);
        const int i0 = sp.x, i1 = sp.y, j0 = sp.z, j1 = sp.w;
        const T* tx = c.tabx + (size_t)kk * P.sx;
        const T* ty = c.taby + (size_t)kk * P.sy;
        double S1[7] = {0, 0, 0, 0, 0, 0, 0};  // 00 10 20 30 01 02 03
        double S2[7] = {0, 0, 0, 0, 0, 0, 0};
        double S3[3] = {0, 0, 0};              // 00 30 03
        for (int jc = j0; jc <= j1; jc += 32) {
            const int j = jc + lane;
            const bool ok = j <= j1;
            const int jj = ok ? j : j1;
            const double b = ok ? (double)ty[jj] : 0.0;
            const double dy = ((double)jj + 0.5) - y;
            double c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c30 = 0.0, c33 = 0.0;
            for (int i = i0; i <= i1; ++i) {
                const double r0 = (double)c.sL[i * P.C + jj], il = (double)c.sL2[i * P.C + jj];
                const double r1 = 1.0 - r0, r2 = r0 * il, r3 = r2 * il;
                const double a = (double)tx[i], dx = ((double)i + 0.5) - x;
                const double a1 = a * dx, a2 = a1 * dx, a3 = a2 * dx;
                const double aa = a * a, aaa = aa * a;
                c1[0] = fma(r1, a, c1[0]);
                c1[1] = fma(r1, a1, c1[1]);
                c1[2] = fma(r1, a2, c1[2]);
                c1[3] = fma(r1, a3, c1[3]);
                const double w2 = r2 * aa;
                c2[0] += w2;
                c2[1] = fma(w2, dx, c2[1]);
                c2[2] = fma(w2, dx * dx, c2[2]);
                c2[3] = fma(w2, dx * dx * dx, c2[3]);
                const double w3 = r3 * aaa;
                c30 += w3;
                c33 = fma(w3, dx * dx * dx, c33);
            }
            const double b1 = b * dy, b2 = b1 * dy, b3 = b2 * dy, bb = b * b, bbb = bb * b;
            S1[0] = fma(b, c1[0], S1[0]); S1[1] = fma(b, c1[1], S1[1]); S1[2] = fma(b, c1[2], S1[2]);
            S1[3] = fma(b, c1[3], S1[3]); S1[4] = fma(b1, c1[0], S1[4]); S1[5] = fma(b2, c1[0], S1[5]);
            S1[6] = fma(b3, c1[0], S1[6]);
            S2[0] = fma(bb, c2[0], S2[0]); S2[1] = fma(bb, c2[1], S2[1]); S2[2] = fma(bb, c2[2], S2[2]);
            S2[3] = fma(bb, c2[3], S2[3]); S2[4] = fma(bb * dy, c2[0], S2[4]); S2[5] = fma(bb * dy * dy, c2[0], S2[5]);
            S2[6] = fma(bb * dy * dy * dy, c2[0], S2[6]);
            S3[0] = fma(bbb, c30, S3[0]); S3[1] = fma(bbb, c33, S3[1]); S3[2] = fma(bbb * dy * dy * dy, c30, S3[2]);
        }
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            S1[m] = warp_sum(S1[m]);
            S2[m] = warp_sum(S2[m]);
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) S3[m] = warp_sum(S3[m]);
        if (lane == 0) {
            const double iv = P.inv_s2, iv2 = iv * iv, iv3 = iv2 * iv, f2 = f * f, f3 = f2 * f;
            d1[3 * k] = S1[0];
            d1[3 * k + 1] = f * (S1[1] * iv);
            d1[3 * k + 2] = f * (S1[4] * iv);
            d2[3 * k] = S2[0];
            d2[3 * k + 1] = f2 * (S2[2] * iv2) + f * ((S1[2] * iv - S1[0]) * iv);
            d2[3 * k + 2] = f2 * (S2[5] * iv2) + f * ((S1[5] * iv - S1[0]) * iv);
            d3[3 * k] = -2.0 * S3[0];
            d3[3 * k + 1] = -f3 * (S3[1] * iv3) + 3.0 * f2 * (S2[3] * iv3 - S2[1] * iv2) +
                            f * ((S1[3] * iv2 - 3.0 * S1[1] * iv) * iv);
            d3[3 * k + 2] = -f3 * (S3[2] * iv3) + 3.0 * f2 * (S2[6] * iv3 - S2[4] * iv2) +
                            f * ((S1[6] * iv2 - 3.0 * S1[4] * iv) * iv);
        }
    }
}

// Pixel potential -sum(D ln Lambda - Lambda) plus d1 (-> c.g), d2 (-> c.a1), d3 (-> c.a2) at the current c.q.
template <typename T, int MR, int MC>
__device__ double eval_hess_pixels(Ctx<T>& c) {
    const FieldParams& P = *c.P;
    const int nchunks = c.N > 0 ? (c.N + P.Kc - 1) / P.Kc : 1;
    double vacc = 0.0;
    c.hess_out = true;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int k0 = ch * P.Kc, nk = min(P.Kc, c.N - k0);
        build_tables<T>(c, k0, nk, true);
        __syncthreads();
        render_chunk<T, MR, MC>(c, nk, ch == 0, ch == nchunks - 1, true, vacc);
        __syncthreads();
    }
    c.hess_out = false;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int k0 = ch * P.Kc, nk = min(P.Kc, c.N - k0);
        if (nk <= 0) break;
        build_tables<T>(c, k0, nk, false);
        __syncthreads();
        hess_chunk<T>(c, k0, nk, c.g, c.a1, c.a2);
        __syncthreads();
    }
    double v[1] = {vacc};
    block_sum<1>(v, c.red);
    return v[0];
}

// dq/dt, dp/dt, E of RHMC_efficient_computation from the cached derivatives (c.g, c.a1, c.a2 = d1, d2, d3) and the
// momentum `pm` (samplers.py:905-912).  Returns E in every thread; all-inf when a flux is below f_lim (:832-837).
template <typename T>
__device__ double hess_flow(const Ctx<T>& c, double Vpix, double f_lim, const double* pm, double* dqdt, double* dpdt) {
    // sum p^2/d2; ln|prod d2| as sum ln|d2| with the sign of the product tracked (np.log of a negative product is
    // NaN, of a positive product of two negative curvatures it is not); count of fluxes below the floor
    double v[4] = {0, 0, 0, 0};
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < 3 * c.N; i += nt) {
        const double h = c.a1[i], dq = pm[i] / h;
        if (dqdt) dqdt[i] = dq;
        if (dpdt) dpdt[i] = -c.a2[i] * ((1.0 / h) - dq * dq) / 2.0 - c.g[i];
        v[0] += (pm[i] * pm[i]) / h;
        v[1] += log(fabs(h));
        if (h < 0.0) v[3] += 1.0;
        if ((i % 3) == 0 && c.q[i] < f_lim) v[2] += 1.0;
    }
    block_sum<4>(v, c.red);
    if (fmod(v[3], 2.0) != 0.0) v[1] = CUDART_NAN;
    if (v[2] > 0.0) {
        __syncthreads();
        for (int i = tid; i < 3 * c.N; i += nt) {
            if (dqdt) dqdt[i] = CUDART_INF;
            if (dpdt) dpdt[i] = CUDART_INF;
        }
        __syncthreads();
        return CUDART_INF;
    }
    __syncthreads();
    return v[0] / 2.0 + v[1] / 2.0 + Vpix;
}

// E(q, p[, M]) of lightsource_gym (samplers.py:1152-1173): inf below the flux floor, V + K otherwise.
template <typename T>
__device__ double ls_energy(const Ctx<T>& c, double Vpix, double f_lim, const double* pm, bool with_mass, double factor1) {
    double v[3] = {0, 0, 0};
    for (int k = threadIdx.x; k < c.N; k += blockDim.x) {
        const double f = c.q[3 * k];
        const double pf = pm[3 * k], px = pm[3 * k + 1], py = pm[3 * k + 2];
        if (with_mass) {
            const double G = 1.0 / f, Fm = f * factor1;
            v[0] += pf * pf / G + px * px / Fm + py * py / Fm;
            v[1] += log(fabs(G)) + 2.0 * log(fabs(Fm));
        } else {
            v[0] += pf * pf + px * px + py * py;
        }
        if (f < f_lim) v[2] += 1.0;
    }
    block_sum<3>(v, c.red);
    if (v[2] > 0.0) return CUDART_INF;
    return Vpix + (v[0] + v[1]) / 2.0;
}

template <typename T, int MR, int MC>
__global__ void __launch_bounds__(512, 1)
ls_kernel(const __grid_constant__ FieldParams P, const __grid_constant__ LsArgs A, double* scratch, int d_in_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SmemLayout lay = make_layout<T>(P, d_in_smem != 0);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int S = 3 * P.Nmax;

    for (int field = blockIdx.x; field < A.n_fields; field += gridDim.x) {
        Ctx<T> c;
        c.P = &P;
        c.q = reinterpret_cast<double*>(smem_raw + lay.q);
        c.p = reinterpret_cast<double*>(smem_raw + lay.p);
        c.g = reinterpret_cast<double*>(smem_raw + lay.g);
        c.a1 = reinterpret_cast<double*>(smem_raw + lay.a1);
        c.a2 = reinterpret_cast<double*>(smem_raw + lay.a2);
        c.red = reinterpret_cast<double*>(smem_raw + lay.red);
        c.sD = d_in_smem ? reinterpret_cast<T*>(smem_raw + lay.D) : nullptr;
        c.sL = reinterpret_cast<T*>(smem_raw + lay.L);
        c.sL2 = P.hess ? reinterpret_cast<T*>(smem_raw + lay.L2) : nullptr;
        c.tabx = reinterpret_cast<T*>(smem_raw + lay.tabx);
        c.taby = reinterpret_cast<T*>(smem_raw + lay.taby);
        c.span = reinterpret_cast<short*>(smem_raw + lay.span);
        c.gD = reinterpret_cast<const T*>(A.D) + (P.D_shared ? 0 : (size_t)field * P.R * P.C);
        c.bg = A.background ? A.background + (size_t)field * P.R * P.C : nullptr;
        c.hess_out = false;
        c.N = A.nstars ? A.nstars[field] : P.Nmax;
        c.g_ff2 = 1.0;
        c.beta = 0.0;
        c.h = 0.0;
        const int n3 = 3 * c.N;
        // global scratch per field: q_init, g_init, p_tmp, flags / dqdt, dpdt
        double* qs = scratch + (size_t)field * 5 * S;
        double* gs = qs + S;
        double* pt = gs + S;
        double* w1 = pt + S;
        double* w2 = w1 + S;

        __syncthreads();
        if (c.sD)
            for (int i = tid; i < P.R * P.C; i += nt) c.sD[i] = c.gD[i];
        for (int i = tid; i < n3; i += nt) {
            c.q[i] = A.q0[(size_t)field * S + i];
            c.p[i] = A.p0 ? A.p0[(size_t)field * S + i] : 0.0;
        }
        __syncthreads();
        const int niter = A.niter, L = niter + 1;
        const double* Z = A.normals ? A.normals + (size_t)field * L * S : nullptr;
        const int* ST = A.steps ? A.steps + (size_t)field * niter : nullptr;
        const double* LU = A.lnu ? A.lnu + (size_t)field * niter : nullptr;
        double* QC = A.q_chain ? A.q_chain + (size_t)field * L * S : nullptr;
        double* EC = A.E_chain ? A.E_chain + (size_t)field * L : nullptr;
        double* DC = A.dE_chain ? A.dE_chain + (size_t)field * L : nullptr;
        unsigned char* AC = A.A_chain ? A.A_chain + (size_t)field * niter : nullptr;
        int n_acc = 0;

        if (A.variant == LS_EVAL_HESS) {
            const double Vpix = eval_hess_pixels<T, MR, MC>(c);
            const size_t o = (size_t)field * S;
            if (A.d2_only) {
                for (int i = tid; i < n3; i += nt) A.d2[o + i] = c.a1[i];
            } else {
                const double E = hess_flow(c, Vpix, A.f_lim, c.p, w1, w2);
                for (int i = tid; i < n3; i += nt) {
                    if (A.d1) A.d1[o + i] = c.g[i];
                    if (A.d2) A.d2[o + i] = c.a1[i];
                    if (A.d3) A.d3[o + i] = c.a2[i];
                    if (A.dqdt) A.dqdt[o + i] = w1[i];
                    if (A.dpdt) A.dpdt[o + i] = w2[i];
                }
                if (tid == 0 && A.E_out) A.E_out[field] = E;
            }
        } else if (A.variant == LS_EVAL_BG) {
            // V_single / dVdq_single: potential and gradient on a model_data background (samplers.py:77-127)
            const double Vpix = eval_pixels<T, MR, MC>(c, true);
            for (int i = tid; i < n3; i += nt) A.d1[(size_t)field * S + i] = c.g[i];
            if (tid == 0) A.E_out[field] = Vpix;
        } else if (A.variant == LS_HMC || A.variant == LS_DIAG) {
            const bool diag = A.variant == LS_DIAG;
            const double dtg = A.dt[0];
            auto dt_of = [&](int i) { return diag ? dtg : A.dt[i]; };
            // force on component i at the current q (c.g = pixel gradient) with the iteration's initial momentum pt
            auto force = [&](int i) -> double {
                double fo = c.g[i];
                if (diag && (i % 3) == 0) {
                    const double f = c.q[i], G = 1.0 / f, dG = -1.0 / (f * f), Fm = f * A.factor1, dF = A.factor1;
                    const double pf = pt[i], px = pt[i + 1], py = pt[i + 2];
                    fo += 0.5 * (dG / G + 2.0 * dF / Fm);
                    fo += -0.5 * (dG * (pf * pf) / (G * G) + (px * px + py * py) * dF / (Fm * Fm));
                }
                return fo;
            };
            auto draw_p = [&](int l) {  // p_sample() [* sqrt(M(q))] -> pt
                for (int i = tid; i < n3; i += nt) {
                    double z = Z[(size_t)l * S + i];
                    if (diag) {
                        const double f = c.q[i - (i % 3)];
                        z *= sqrt((i % 3) == 0 ? 1.0 / f : f * A.factor1);
                    }
                    pt[i] = z;
                }
                __syncthreads();
            };
            double Vpix = eval_pixels<T, MR, MC>(c, true);
            draw_p(0);
            double e_prev = ls_energy(c, Vpix, A.f_lim, pt, diag, A.factor1);
            if (QC) for (int i = tid; i < n3; i += nt) QC[i] = c.q[i];
            if (tid == 0) { if (EC) EC[0] = e_prev; if (DC) DC[0] = 0.0; }
            for (int it = 1; it <= niter; ++it) {
                for (int i = tid; i < n3; i += nt) { qs[i] = c.q[i]; gs[i] = c.g[i]; }
                const double Vpix0 = Vpix;
                draw_p(it);
                const double e0 = ls_energy(c, Vpix, A.f_lim, pt, diag, A.factor1);
                if (tid == 0) { if (EC) EC[it] = e0; if (DC) DC[it] = e0 - e_prev; }
                const int nst = ST[it - 1];
                for (int i = tid; i < n3; i += nt) {
                    c.p[i] = pt[i] - dt_of(i) * force(i) / 2.0;  // first half step
                    c.a2[i] = 0.0;                                // iflip
                }
                __syncthreads();
                int flip = 0;
                for (int s = 0; s < nst; ++s) {
                    int fl = 0;
                    for (int k = tid; k < c.N; k += nt) {
                        const double f = c.q[3 * k];
                        double m0 = 1.0, m1 = 1.0;
                        if (diag) { m0 = 1.0 / f; m1 = f * A.factor1; }
                        const double nf = f + dt_of(3 * k) * c.p[3 * k] / m0;
                        c.q[3 * k] = nf;
                        c.q[3 * k + 1] += dt_of(3 * k + 1) * c.p[3 * k + 1] / m1;
                        c.q[3 * k + 2] += dt_of(3 * k + 2) * c.p[3 * k + 2] / m1;
                        if (nf < A.f_lim) { c.a2[3 * k] = 1.0; fl = 1; }
                    }
                    flip = __syncthreads_or(fl);
                    const double v = eval_pixels<T, MR, MC>(c, s == nst - 1);
                    if (s == nst - 1) Vpix = v;
                    for (int i = tid; i < n3; i += nt) {
                        const double old = c.p[i];
                        c.p[i] = (flip && c.a2[i] != 0.0) ? -old : old - dt_of(i) * force(i);
                    }
                    __syncthreads();
                }
                if (!flip) {  // final half-step correction; in the flip branch the reference never updates p_tmp
                    for (int i = tid; i < n3; i += nt) w1[i] = force(i);  // reads the OLD p_tmp (dpMpdq quirk)
                    __syncthreads();
                    for (int i = tid; i < n3; i += nt) pt[i] = c.p[i] + dt_of(i) * w1[i] / 2.0;
                    __syncthreads();
                }
                const double e1 = ls_energy(c, Vpix, A.f_lim, pt, diag, A.factor1);
                const double dE = e1 - e0;
                e_prev = e0;
                const bool accept = (dE < 0.0) || (LU[it - 1] < -dE);
                if (tid == 0 && AC) AC[it - 1] = accept ? 1 : 0;
                if (accept) {
                    ++n_acc;
                } else {
                    __syncthreads();
                    for (int i = tid; i < n3; i += nt) { c.q[i] = qs[i]; c.g[i] = gs[i]; }
                    Vpix = Vpix0;
                }
                __syncthreads();
                if (QC) for (int i = tid; i < n3; i += nt) QC[(size_t)it * S + i] = c.q[i];
            }
        } else if (A.variant == LS_TRIAL) {
            // single star on the model_data background; scalar flux gradient kicks every component
            double Vpix = eval_pixels<T, MR, MC>(c, true);
            for (int it = 1; it <= niter; ++it) {
                for (int i = tid; i < n3; i += nt) { qs[i] = c.q[i]; gs[i] = c.g[i]; }
                const double Vpix0 = Vpix;
                for (int i = tid; i < n3; i += nt) pt[i] = (A.zero_xy && (i % 3) != 0) ? 0.0 : Z[(size_t)it * S + i];
                __syncthreads();
                const double e0 = ls_energy(c, Vpix, -CUDART_INF, pt, false, 0.0);
                const int nst = ST[it - 1];
                for (int i = tid; i < n3; i += nt) c.p[i] = pt[i] - A.dt[i] * c.g[i - (i % 3)] / 2.0;
                __syncthreads();
                for (int s = 0; s < nst; ++s) {
                    for (int i = tid; i < n3; i += nt) c.q[i] += A.dt[i] * c.p[i];
                    __syncthreads();
                    const double v = eval_pixels<T, MR, MC>(c, s == nst - 1);
                    if (s == nst - 1) Vpix = v;
                    for (int i = tid; i < n3; i += nt) c.p[i] -= A.dt[i] * c.g[i - (i % 3)];
                    __syncthreads();
                }
                for (int i = tid; i < n3; i += nt) pt[i] = c.p[i] + A.dt[i] * c.g[i - (i % 3)] / 2.0;
                __syncthreads();
                const double e1 = ls_energy(c, Vpix, -CUDART_INF, pt, false, 0.0);
                const double dE = e1 - e0;
                const bool accept = (dE < 0.0) || (LU[it - 1] < -dE);
                if (tid == 0 && AC) AC[it - 1] = accept ? 1 : 0;
                if (accept) {
                    ++n_acc;
                } else {
                    __syncthreads();
                    for (int i = tid; i < n3; i += nt) { c.q[i] = qs[i]; c.g[i] = gs[i]; }
                    Vpix = Vpix0;
                }
                __syncthreads();
            }
        } else {  // LS_HESS: RHMC_random
            double e_prev = 0.0;
            for (int i = tid; i < n3; i += nt) c.p[i] = 0.0;
            double Vpix = eval_hess_pixels<T, MR, MC>(c);  // derivatives at q0
            for (int it = 0; it <= niter; ++it) {
                // p_tmp = p_sample() * sqrt(dVdqq)
                for (int i = tid; i < n3; i += nt) pt[i] = Z[(size_t)it * S + i] * sqrt(c.a1[i]);
                __syncthreads();
                double E = hess_flow(c, Vpix, A.f_lim, pt, w1, w2);  // dqdt -> w1, dpdt -> w2
                if (it == 0) {
                    if (QC) for (int i = tid; i < n3; i += nt) QC[i] = c.q[i];
                    if (tid == 0) { if (EC) EC[0] = E; if (DC) DC[0] = 0.0; }
                    e_prev = E;
                    continue;
                }
                const double e0 = E;
                if (tid == 0) { if (EC) EC[it] = e0; if (DC) DC[it] = e0 - e_prev; }
                const int nst = ST[it - 1];
                for (int i = tid; i < n3; i += nt) {
                    c.p[i] = pt[i] + A.dt[i] * w2[i] / 2.0;  // p_half
                    qs[i] = 0.0;                              // flip flags (flux / x / y share the slot index)
                }
                __syncthreads();
                for (int s = 0; s < nst; ++s) {
                    E = hess_flow(c, Vpix, A.f_lim, c.p, w1, w2);
                    if (E == CUDART_INF) break;  // "Divergence encountered"
                    for (int k = tid; k < c.N; k += nt) {
                        const double f = c.q[3 * k] + A.dt[3 * k] * w1[3 * k];
                        const double x = c.q[3 * k + 1] + A.dt[3 * k + 1] * w1[3 * k + 1];
                        const double y = c.q[3 * k + 2] + A.dt[3 * k + 2] * w1[3 * k + 2];
                        c.q[3 * k] = f; c.q[3 * k + 1] = x; c.q[3 * k + 2] = y;
                        if (f < A.f_lim) qs[3 * k] = 1.0;
                        if ((x < 0.0) || (x < (double)P.R)) qs[3 * k + 1] = 1.0;   // sic (samplers.py:1038)
                        if ((y < 0.0) || (y < (double)P.R)) qs[3 * k + 2] = 1.0;   // sic: num_rows for y too (:1041)
                    }
                    __syncthreads();
                    Vpix = eval_hess_pixels<T, MR, MC>(c);
                    hess_flow(c, Vpix, A.f_lim, c.p, (double*)nullptr, w2);
                    for (int i = tid; i < n3; i += nt) {
                        const double dtt = (s == nst - 1) ? A.dt[i] / 2.0 : A.dt[i];
                        const double old = c.p[i];
                        c.p[i] = (qs[i] != 0.0) ? -old : old + dtt * w2[i];
                    }
                    __syncthreads();
                }
                const double e1 = hess_flow(c, Vpix, A.f_lim, c.p, (double*)nullptr, (double*)nullptr);
                const double dE = e1 - e0;
                e_prev = e0;
                const bool accept = (dE < 0.0) || (LU[it - 1] < -dE);
                if (tid == 0 && AC) AC[it - 1] = accept ? 1 : 0;
                if (accept) ++n_acc;
                // no restore on rejection: q_tmp += ... aliased q_initial in the reference (samplers.py:1031)
                if (QC) for (int i = tid; i < n3; i += nt) QC[(size_t)it * S + i] = c.q[i];
                if ((it % 100) == 0 && (100.0 * n_acc) / (double)it < 50.0) break;  // samplers.py:1098-1101
            }
        }
        if (A.variant != LS_EVAL_HESS && A.variant != LS_EVAL_BG) {
            if (A.q_final) for (int i = tid; i < n3; i += nt) A.q_final[(size_t)field * S + i] = c.q[i];
            if (tid == 0 && A.accept_count) A.accept_count[field] = (double)n_acc;
        }
        __syncthreads();
    }
}

}  // namespace srhmc

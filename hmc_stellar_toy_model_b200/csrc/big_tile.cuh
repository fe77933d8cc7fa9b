// Fused tile evaluation of the large-field engine (BASELINE configs[4]; north_star: "one CTA per image tile, the
// residual image staged in shared memory").  One gradient evaluation of the whole field (reference:
// base_class.dVdq / base_class.V, sampler_RHMC.py:365-425 / 294-351, PSF truncated to the (2r+1)^2 patch) reads
// the data image from HBM exactly ONCE and never materialises Lambda or rho in global memory:
//
//   bin        one thread per star (own stars: inside the scalar kernel that finishes the position update; ghosts: a
//              small kernel after the exchange): append a pair record (star id, box patch x tile, footprint slot) to
//              the fixed-capacity list of each of the <= 2x2 tiles its patch touches (one atomic per pair)
//   tile       one CTA per 64x64 tile: sort the tile's list by star id (fixed summation order), render
//              Lambda = B + sum f PSF for the tile in registers (4x4 pixels per thread; per-star row/column
//              Gaussian tables in shared memory), read D, rho = D/Lambda - 1 into the shared tile, V partial of
//              the owned rows, then one warp per owned star of the list: the three residual-weighted PSF
//              reductions over patch x tile  -> gpart[star][slot of this tile in the star's 2x2 footprint]
//              (EVAL_V: the last tile to finish sums the per-tile potential partials in tile order); the CTA
//              re-zeroes its list counter for the next evaluation
//   gsum       per owned star: sum of its footprint slots in fixed order, scale -> g[3n]; fused into the scalar kernel
//              that consumes the gradient (big_tail_kernel) or run on its own (big_gsum_kernel)
//
// No floating-point atomics anywhere: results are bit-reproducible run to run.
#pragma once

namespace {

constexpr int kTile = 64;        // tile edge in pixels (>= 2 kMaxRad + 1, so a patch touches at most 2x2 tiles)
constexpr int kTileThreads = 256;
constexpr int kTileChunk = 16;   // pairs per table build (each warp builds the tables of two pairs)
constexpr int kTileMaxList = 1024;  // stars per tile list (0.13 stars/px over tile + halo)

struct TileSrc {
    const double* q;
    const double* ga;
    const double* gb;
    int n_own, cap;
};

// star record of source id `sid`: own stars [0, n_own), then the two ghost lists (capacity `cap` each)
__device__ __forceinline__ const double* tile_source(const TileSrc& S, int sid) {
    if (sid < S.n_own) return S.q + 3 * (size_t)sid;
    sid -= S.n_own;
    if (sid < S.cap) return (S.ga && sid < (int)S.ga[0]) ? S.ga + 1 + 3 * (size_t)sid : nullptr;
    sid -= S.cap;
    return (S.gb && sid < S.cap && sid < (int)S.gb[0]) ? S.gb + 1 + 3 * (size_t)sid : nullptr;
}

struct TileSpan {
    int ti0, ti1, tj0, tj1;
};

__device__ __forceinline__ TileSpan tile_span(const BigParams& P, int i0, int i1, int j0, int j1) {
    TileSpan t;
    t.ti0 = (i0 - P.row0) / kTile;
    t.ti1 = (i1 - P.row0) / kTile;
    t.tj0 = j0 / kTile;
    t.tj1 = j1 / kTile;
    return t;
}

// (star, tile) pair record: star id and the box patch x tile in tile coordinates (6 bits each) + the index of the
// tile in the star's 2x2 footprint
__device__ __forceinline__ int pack_box(int ia, int ib, int ja, int jb, int slot) {
    return ia | (ib << 6) | (ja << 12) | (jb << 18) | (slot << 24);
}

// Append the pair records of one source star to the lists of the tiles its patch touches.  `owned`: the star belongs
// to this rank, so leaving the local data window is an error (flag 1).  List overflow raises flag 3.
__device__ __forceinline__ void bin_star(const BigParams& P, int ntx, int sid, double x, double y, bool owned, int* cnt,
                                         int2* list, int* err) {
    int i0, i1, j0, j1, mi, mj;
    bool clipped;
    if (!patch_of(P, x, y, i0, i1, j0, j1, mi, mj, clipped)) {
        if (owned) atomicExch(err, 1);  // an owned star left the local data window
        return;
    }
    if (owned && clipped) atomicExch(err, 1);
    const TileSpan t = tile_span(P, i0, i1, j0, j1);
    for (int ti = t.ti0; ti <= t.ti1; ++ti)
        for (int tj = t.tj0; tj <= t.tj1; ++tj) {
            const int tile = ti * ntx + tj;
            const int pos = atomicAdd(&cnt[tile], 1);
            if (pos >= kTileMaxList) {
                atomicExch(err, 3);  // list capacity exceeded (density above 0.13 stars/px)
                continue;
            }
            const int r0 = P.row0 + ti * kTile, c0 = tj * kTile;
            list[(size_t)tile * kTileMaxList + pos] =
                make_int2(sid, pack_box(max(i0, r0) - r0, min(i1, r0 + kTile - 1) - r0, max(j0, c0) - c0,
                                        min(j1, c0 + kTile - 1) - c0, (ti - t.ti0) * 2 + (tj - t.tj0)));
        }
}

// sources [sid0, sid1): own stars are [0, n_own), the ghost lists follow (tile_source)
__global__ void big_bin_kernel(const BigParams P, const TileSrc S, int ntx, int sid0, int sid1, int* cnt, int2* list, int* err) {
    for (int sid = sid0 + blockIdx.x * blockDim.x + threadIdx.x; sid < sid1; sid += gridDim.x * blockDim.x) {
        const double* src = tile_source(S, sid);
        if (!src) continue;
        bin_star(P, ntx, sid, src[1], src[2], sid < S.n_own, cnt, list, err);
    }
}

// Data loads before (1) or after (0) the render phase: early loads overlap HBM latency with the render inside one CTA
// but hold 32 more registers per thread.
#ifndef SRHMC_TILE_EARLY_LOAD
#define SRHMC_TILE_EARLY_LOAD 0
#endif
#ifndef SRHMC_TILE_MIN_CTAS
#define SRHMC_TILE_MIN_CTAS 4
#endif

constexpr int kTabPad = 4;                 // zero guard entries on each side of a 32-entry factor table
constexpr int kTabLen = 32 + 2 * kTabPad;  // a thread reads 4 consecutive entries starting anywhere in [-3, 31]

// Factor tables of one (star, tile) pair, built by one warp with two warp-wide exponentials (exp_neg: the kernels'
// own branch-free FP64 exponential, fastmath.cuh) and used by BOTH the render and the gather phase:
//   rowf[kTabPad + k] = ex_k      for row  ia + k of the box,  ex = exp(-(i+.5-x)^2/2s^2)
//   colf[kTabPad + k] = f ey_k    for column ja + k,           ey = norm exp(-(j+.5-y)^2/2s^2)
// zero past the box and in the guards; (dx0, dy0) = offsets of the box's first row / column from the star, from which
// the gather phase forms ex dx and ey dy.
struct PairTab {
    double rowf[kTabLen];
    double colf[kTabLen];
    double dx0, dy0;
};

struct TileSmem {
    double rho[kTile][kTile];          // 32 KB
    PairTab tab[kTileChunk];           // 656 B each
    int2 list[kTileMaxList];           // sorted pair records
    int box[kTileChunk][4];
    double red[32];
    double2 ltab[kLogTableSize];       // table of log_pos (fastmath.cuh), built per CTA when the potential is wanted
    bool is_last;
};

__device__ __forceinline__ void build_pair_tab(const BigParams& P, const TileSrc& S, int2 rec, int r0, int c0, int lane,
                                               PairTab& T, int* box) {
    const double* src = tile_source(S, rec.x);
    const double f = src[0], x = src[1], y = src[2];
    const int ia = rec.y & 63, ib = (rec.y >> 6) & 63, ja = (rec.y >> 12) & 63, jb = (rec.y >> 18) & 63;
    const double dx = ((double)(r0 + ia + lane) + 0.5) - x, dy = ((double)(c0 + ja + lane) + 0.5) - y;
    T.rowf[kTabPad + lane] = (ia + lane <= ib) ? exp_neg(-(dx * dx) * P.inv2s2) : 0.0;
    T.colf[kTabPad + lane] = (ja + lane <= jb) ? exp_neg(-(dy * dy) * P.inv2s2) * (P.norm * f) : 0.0;
    if (lane == 0) {
        T.dx0 = dx;
        T.dy0 = dy;
        if (box) {
            box[0] = ia; box[1] = ib; box[2] = ja; box[3] = jb;
        }
    }
}

// 16-byte asynchronous copy global -> shared (LDGSTS): no registers held while the data are in flight
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// MODE 0: gradient; 1: gradient + pixel potential; 2: mock data -- the rendered model of the tile is Poisson-sampled
// (poisson.cuh, counter = global pixel index) into Dout and nothing else happens
template <int MODE>
__global__ void __launch_bounds__(kTileThreads, SRHMC_TILE_MIN_CTAS) big_tile_kernel(const BigParams P, const TileSrc S, int ntx,
                                                                const double* __restrict__ D,
                                                                int* __restrict__ cnt,
                                                                const int2* __restrict__ list, double* __restrict__ gpart,
                                                                double* vpart, unsigned int* ticket, double* scalars,
                                                                int* fp_counters, double* __restrict__ Dout,
                                                                unsigned long long mock_seed) {
    constexpr bool WANT_V = MODE == 1;
    extern __shared__ __align__(16) unsigned char tile_smem_raw[];
    TileSmem& sm = *reinterpret_cast<TileSmem*>(tile_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kTileThreads / 32;
    const int ti = blockIdx.x / ntx, tj = blockIdx.x % ntx;
    const int r0 = P.row0 + ti * kTile, c0 = tj * kTile;  // global pixel of the tile origin
    const int vr = min(kTile, P.row0 + P.nrows - r0), vc = min(kTile, P.C - c0);  // valid rows / columns

    const int ty = tid >> 4, tx = tid & 15;
    const int pr = 4 * ty, pc = 4 * tx;  // thread owns the 4x4 pixel block at (pr, pc)
    // The tile's data pixels first: every thread copies its own 4x4 block asynchronously straight into the rho buffer, so
    // the HBM latency is covered by the list sort, the table build and the render; the residual later overwrites the
    // block in place (a thread only ever touches its own block before the barrier that precedes the gather).
    if (MODE != 2) {
        const bool vec = ((P.C & 1) == 0) && (pc + 3 < vc);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int li = pr + a;
            if (li < vr) {
                const double* row = D + (size_t)(r0 + li - P.row0) * P.C + c0 + pc;
                if (vec) {
                    cp_async16(&sm.rho[li][pc], row);
                    cp_async16(&sm.rho[li][pc + 2], row + 2);
                } else {
#pragma unroll
                    for (int b = 0; b < 4; ++b) sm.rho[li][pc + b] = (pc + b < vc) ? __ldg(row + b) : 0.0;
                }
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) sm.rho[li][pc + b] = 0.0;
            }
        }
        cp_async_commit();
    }

    // ---- the tile's pair list, sorted by star id so that every sum below has a fixed order
    list += (size_t)blockIdx.x * kTileMaxList;
    constexpr int b0 = 0;
    const int nl = min(cnt[blockIdx.x], kTileMaxList);  // an overflow was flagged when the list was filled
    // the fixed-point iteration counters of the leapfrog step are free between its last reader and the next step
    if (fp_counters && blockIdx.x == 0 && tid < 2) fp_counters[tid] = 0;
    // guard entries of the factor tables stay zero for the whole launch; the 32 body entries are rewritten per pair
    for (int k = tid; k < kTileChunk * 4 * kTabPad; k += kTileThreads) {
        const int pair = k / (4 * kTabPad), e = k % (4 * kTabPad), side = e / kTabPad, g = e % kTabPad;
        double* t = (side & 1) ? sm.tab[pair].colf : sm.tab[pair].rowf;
        t[(side & 2) ? kTabPad + 32 + g : g] = 0.0;
    }
    if (WANT_V && tid < kLogTableSize) {
        // (rc_k, -ln rc_k) with rc_k ~ 1/(bin centre): ln x = e ln2 - ln rc_k + log1p(m rc_k - 1) is an identity for the
        // ROUNDED rc_k, so the table only needs -ln rc_k to double accuracy
        const double rc = 1.0 / (1.0 + ((double)tid + 0.5) / (double)kLogTableSize);
        sm.ltab[tid] = make_double2(rc, -log(rc));
    }
    if (nl == 1) {
        if (tid == 0) sm.list[0] = __ldcg(&list[b0]);
    } else if (nl > 1) {
        for (int k = tid; k < nl; k += kTileThreads) {
            const int2 rec = __ldcg(&list[b0 + k]);
            int r = 0;
            for (int m = 0; m < nl; ++m) r += __ldcg(&list[b0 + m].x) < rec.x;
            sm.list[r] = rec;
        }
    }
    __syncthreads();
    if (tid == 0) cnt[blockIdx.x] = 0;  // every thread has read the count: ready for the next evaluation's binning

    // ---- render Lambda = B + sum f PSF over the list, in list order
    double lam[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) lam[a][b] = P.F.B;
    for (int base = 0; base < nl; base += kTileChunk) {
        const int nc = min(kTileChunk, nl - base);
        for (int s = warp; s < nc; s += kWarps) build_pair_tab(P, S, sm.list[base + s], r0, c0, lane, sm.tab[s], sm.box[s]);
        __syncthreads();
        for (int s = 0; s < nc; ++s) {
            const int ia = sm.box[s][0], ja = sm.box[s][2];
            if (pr + 3 < ia || pr > sm.box[s][1] || pc + 3 < ja || pc > sm.box[s][3]) continue;
            const double* te = &sm.tab[s].rowf[kTabPad + pr - ia];  // pr - ia in [-3, 31]: inside the padded table
            const double* tf = &sm.tab[s].colf[kTabPad + pc - ja];
            const double ex[4] = {te[0], te[1], te[2], te[3]}, fy[4] = {tf[0], tf[1], tf[2], tf[3]};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) lam[a][b] = fma(ex[a], fy[b], lam[a][b]);
        }
        if (base + kTileChunk < nl) __syncthreads();  // the tables are rewritten by the next chunk
    }

    if (MODE == 2) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int li = pr + a, lj = pc + b;
                if (li < vr && lj < vc)
                    Dout[(size_t)(r0 + li - P.row0) * P.C + c0 + lj] =
                        poisson_draw(lam[a][b], mock_seed, (unsigned long long)(r0 + li) * (unsigned long long)P.C + (unsigned long long)(c0 + lj));
            }
        return;
    }

    // ---- residual into the shared tile; pixel potential of the owned rows
    cp_async_wait_all();  // this thread's own copies (it reads nothing else here)
    double v = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int li = pr + a, gi = r0 + li;
        const bool own = WANT_V && li < vr && gi >= P.own_lo && gi < P.own_hi;
        const double2 d0 = *reinterpret_cast<const double2*>(&sm.rho[li][pc]);
        const double2 d1 = *reinterpret_cast<const double2*>(&sm.rho[li][pc + 2]);
        const double d[4] = {d0.x, d0.y, d1.x, d1.y};
        double rho[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const bool in = li < vr && pc + b < vc;
            rho[b] = in ? fma(d[b], rcp_fast(lam[a][b]), -1.0) : 0.0;
            if (WANT_V && own && pc + b < vc) v += lam[a][b] - d[b] * (lam[a][b] >= 2.3e-308 ? log_pos(lam[a][b], sm.ltab) : CUDART_NAN);  // ln of a non-positive model is NaN, as in NumPy
        }
        *reinterpret_cast<double2*>(&sm.rho[li][pc]) = make_double2(rho[0], rho[1]);
        *reinterpret_cast<double2*>(&sm.rho[li][pc + 2]) = make_double2(rho[2], rho[3]);
    }
    if (WANT_V) {
        double a1[1] = {v};
        block_sum<1>(a1, sm.red);
        if (tid == 0) vpart[blockIdx.x] = a1[0];
    }
    __syncthreads();

    // ---- gather: one warp per owned star of the list (sampler_RHMC.py:404-406 restricted to patch x tile);
    //      lanes own the columns of the box.  A single-chunk list still has its tables in shared memory; otherwise
    //      the warp rebuilds the pair's tables in its own slot.
    const bool keep = nl <= kTileChunk;
    for (int s = warp; s < nl; s += kWarps) {
        const int2 rec = sm.list[s];
        if (rec.x >= S.n_own) continue;  // ghosts are rendered only
        PairTab& T = sm.tab[keep ? s : warp];
        if (!keep) {
            __syncwarp();
            build_pair_tab(P, S, rec, r0, c0, lane, T, nullptr);
            __syncwarp();
        }
        const int ia = rec.y & 63, ib = (rec.y >> 6) & 63, ja = (rec.y >> 12) & 63, jb = (rec.y >> 18) & 63;
        const double fy = T.colf[kTabPad + lane];                // f ey of this lane's column, 0 past the box
        const double dyl = T.dy0 + (double)lane;
        const double* col = &sm.rho[ia][min(ja + lane, kTile - 1)];
        const double* rf = &T.rowf[kTabPad];
        double a0 = 0.0, a1 = 0.0, dxk = T.dx0;
        const int nr = ib - ia + 1;
#pragma unroll 4
        for (int k = 0; k < nr; ++k) {
            const double e = rf[k];
            const double rho = col[k * kTile];
            a0 = fma(rho, e, a0);
            a1 = fma(rho * e, dxk, a1);
            dxk += 1.0;
        }
        // three warp sums with six shuffles: fold the values onto lane groups first
        if (ja + lane > jb) a0 = a1 = 0.0;  // lanes past the box read a clamped column
        double sf = fy * a0, sx = fy * a1, sy = (fy * dyl) * a0, sz = 0.0;
        {
            const bool hi = lane & 16;
            const double k0 = hi ? sy : sf, k1 = hi ? sz : sx, t0 = hi ? sf : sy, t1 = hi ? sx : sz;
            const double x0 = k0 + __shfl_xor_sync(0xffffffffu, t0, 16), x1 = k1 + __shfl_xor_sync(0xffffffffu, t1, 16);
            const bool h8 = lane & 8;
            double yv = (h8 ? x1 : x0) + __shfl_xor_sync(0xffffffffu, h8 ? x0 : x1, 8);
            yv += __shfl_xor_sync(0xffffffffu, yv, 4);
            yv += __shfl_xor_sync(0xffffffffu, yv, 2);
            yv += __shfl_xor_sync(0xffffffffu, yv, 1);
            // lane 0: sum sf, lane 8: sum sx, lane 16: sum sy
            if ((lane & 7) == 0 && lane < 24) gpart[((size_t)rec.x * 4 + (rec.y >> 24)) * 3 + (lane >> 3)] = yv;
        }
    }
    // ---- pixel potential: the last tile to finish sums the per-tile partials in tile order -> scalars[0]
    if (WANT_V) {
        if (last_block_ticket(ticket, &sm.is_last)) {
            double w[1] = {0.0};
            for (int b = tid; b < (int)gridDim.x; b += kTileThreads) w[0] += __ldcg(vpart + b);
            block_sum<1>(w, sm.red);
            if (tid == 0) {
                scalars[0] = w[0];
                *ticket = 0u;
            }
        }
    }
}

// g = scaled sum of the footprint slots of star k, in slot order (sampler_RHMC.py:404-406)
__device__ __forceinline__ void gsum_star(const BigParams& P, int k, double f, double x, double y, const double* gpart,
                                          double& gf, double& gx, double& gy) {
    int i0, i1, j0, j1, mi, mj;
    bool clipped;
    double sf = 0.0, sx = 0.0, sy = 0.0;
    if (patch_of(P, x, y, i0, i1, j0, j1, mi, mj, clipped)) {
        const TileSpan t = tile_span(P, i0, i1, j0, j1);
        for (int ti = t.ti0; ti <= t.ti1; ++ti)
            for (int tj = t.tj0; tj <= t.tj1; ++tj) {
                const double* o = gpart + ((size_t)k * 4 + (ti - t.ti0) * 2 + (tj - t.tj0)) * 3;
                sf += __ldcg(o); sx += __ldcg(o + 1); sy += __ldcg(o + 2);
            }
    }
    // the tile kernel's sums carry the factor f (its column table is f ey)
    gf = -sf / f;
    gx = -sx * P.inv_s2;
    gy = -sy * P.inv_s2;
}

__global__ void big_gsum_kernel(const BigParams P, const double* q, int n_own, const double* gpart, double* g) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_own; k += gridDim.x * blockDim.x)
        gsum_star(P, k, q[3 * k], q[3 * k + 1], q[3 * k + 2], gpart, g[3 * k], g[3 * k + 1], g[3 * k + 2]);
}

}  // namespace

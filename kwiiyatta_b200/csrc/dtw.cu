// Batched DTW / FastDTW for sm_100a.
//
// Replaces fastdtw.fastdtw(x_feature, y_feature, dist=2, radius=radius) at
// kwiiyatta/vocoder/align.py:71 (fastdtw==0.3.2 semantics, see oracle/fastdtw_ref.py):
//   * multi-resolution pyramid ((a+b)/2, odd tail dropped) down to the first level with
//     tx < radius+2 or ty < radius+2, exhaustive DP there;
//   * per finer level the window of row a is the closed form
//       S = { j : (i, j) on the coarser path, |i - a/2| <= radius }
//       cols [ 2*(min S - radius), 2*(max S + radius) + 1 ] clipped to [0, ty-1]
//     so only first_j / last_j per coarse row are carried between levels;
//   * cell rule D = min(up + d, left + d, diag + d), compared after the addition, first
//     minimum wins in the order up, left, diag; FP64 throughout; the local distance is
//     sqrt(sum_k fma(d_k, d_k, s)) summed left to right (p = 2) or sum |d_k| (p = 1).
//
// Kernel structure (one CTA per pair per level):
//   strip-mined systolic wavefront.  A strip is 2*NT consecutive rows; thread t keeps rows
//   i0+2t and i0+2t+1 of x in registers and walks the columns with a skew of one step per
//   thread, so the only inter-thread traffic per step is one double (its lower row's D value)
//   through a double-buffered shared array.  y is staged k-major in a shared-memory ring of
//   2*NT columns, refilled NT columns at a time with coalesced loads.  The D values of a
//   strip's last row go through a small global ping-pong buffer to the next strip.
//   Back-pointers are 2 bits per cell, packed 16 per word by the owning thread.
//
// Two variants of the sweep:
//   * exhaustive levels (the coarsest level of a pair, or radius < 0): dtw_dp_kernel computes the
//     local distances inside the sweep; the distance matrix never exists in memory;
//   * banded levels: the sweep of a narrow band is latency-bound (one barrier per anti-diagonal
//     step, few cells in flight), and with ~72 FP64 instructions per cell inside that chain the
//     FP64 pipe idles most of the time.  dtw_dist_kernel therefore computes the band's local
//     distances first, dependency-free at FP64 issue rate, into a per-row-pair buffer of
//     DIST_WCAP columns, and dtw_dpw_kernel sweeps over them, one warp per pair (3 additions and
//     2 integer comparisons per cell).  Columns of a window beyond the buffer capacity are computed
//     inside the sweep, so the capacity only affects speed.
#include <algorithm>
#include <type_traits>
#include <vector>

#include <climits>
#include <limits>
#include <cstdlib>

#include "common.cuh"

namespace kw {

constexpr int MAXLEV = 20;
constexpr int DPW_NW = 4;             // warps per pair in the banded sweep (strips in flight)
constexpr int DPW_NB = DPW_NW + 1;    // boundary-row buffers per pair

struct PairDesc {
    int nlev;
    int pad_;
    int tx[MAXLEV];
    int ty[MAXLEV];
    long long xoff[MAXLEV];    // doubles, k-major (F x tx) block per level
    long long yoff[MAXLEV];
    long long rj_off[MAXLEV];  // int32: first_j[tx] then last_j[tx] for levels >= 1
    long long xrow0, yrow0;    // first row in the caller's concatenated inputs
    long long bp_off;          // uint32 words, tx0 * ceil(ty0/16)
    long long brow_off;        // doubles, DPW_NB * ty0 (strip boundary rows in flight)
    long long path_off;        // points, capacity tx0 + ty0
    long long dist_off;        // double2 (rows 2q, 2q+1), ceil(tx0/2) * wcap when nlev > 1
    long long win_off;         // int2 (lo, hi) per row pair, ceil(tx0/2) when nlev > 1
};

struct DtwPlan {
    std::vector<PairDesc> descs;
    std::vector<int> order;
    int maxlev = 0;
    size_t n_pyr_x = 0, n_pyr_y = 0, n_rowj = 0, n_bp = 0, n_brow = 0, n_dist = 0, n_win = 0;
    std::vector<int> level_max_tx, level_max_ty;
    std::vector<char> level_has_full, level_has_band;
    int wcap = 0;
};

// Columns of local distances kept per row pair of a banded level.  A window is
// 2 * (span of the coarser path over 2 radius + 1 rows) + 4 radius + 2 columns wide: 8 radius + 2
// for a diagonal path; wider ones spill into the sweep's inline path.
static int dist_wcap(int radius) { return radius < 0 ? 0 : std::max(64, 12 * radius + 16); }
static int make_plan(int n_pairs, const int32_t* tx, const int32_t* ty, int radius, int F,
                     DtwPlan& plan) {
    plan.descs.resize(n_pairs);
    plan.wcap = dist_wcap(radius);
    long long xrow = 0, yrow = 0, path = 0;
    for (int p = 0; p < n_pairs; ++p) {
        PairDesc& d = plan.descs[p];
        if (tx[p] <= 0 || ty[p] <= 0) {
            set_error("pair %d has an empty sequence (tx=%d, ty=%d)", p, tx[p], ty[p]);
            return KW_ERR_INVALID;
        }
        int a = tx[p], b = ty[p], l = 0;
        for (;;) {
            if (l >= MAXLEV) {
                set_error("pair %d needs more than %d resolution levels", p, MAXLEV);
                return KW_ERR_INVALID;
            }
            d.tx[l] = a;
            d.ty[l] = b;
            d.xoff[l] = (long long)plan.n_pyr_x;
            d.yoff[l] = (long long)plan.n_pyr_y;
            plan.n_pyr_x += (size_t)a * F;
            plan.n_pyr_y += (size_t)b * F;
            if (l >= 1) {
                d.rj_off[l] = (long long)plan.n_rowj;
                plan.n_rowj += 2 * (size_t)a;
            } else {
                d.rj_off[l] = 0;
            }
            if ((int)plan.level_max_tx.size() <= l) {
                plan.level_max_tx.push_back(0);
                plan.level_max_ty.push_back(0);
            }
            plan.level_max_tx[l] = std::max(plan.level_max_tx[l], a);
            plan.level_max_ty[l] = std::max(plan.level_max_ty[l], b);
            ++l;
            if (radius < 0 || a < radius + 2 || b < radius + 2) break;
            a /= 2;
            b /= 2;
        }
        d.nlev = l;
        d.pad_ = 0;
        plan.maxlev = std::max(plan.maxlev, l);
        if ((int)plan.level_has_full.size() < l) {
            plan.level_has_full.resize(l, 0);
            plan.level_has_band.resize(l, 0);
        }
        plan.level_has_full[l - 1] = 1;
        for (int q = 0; q + 1 < l; ++q) plan.level_has_band[q] = 1;
        d.dist_off = (long long)plan.n_dist;
        d.win_off = (long long)plan.n_win;
        if (l > 1) {
            plan.n_dist += (size_t)((tx[p] + 1) / 2) * (size_t)plan.wcap;
            plan.n_win += (size_t)((tx[p] + 1) / 2);
        }
        d.xrow0 = xrow;
        d.yrow0 = yrow;
        xrow += tx[p];
        yrow += ty[p];
        d.bp_off = (long long)plan.n_bp;
        plan.n_bp += (size_t)((tx[p] + 7) / 8) * (size_t)((ty[p] + 15) / 16) * 8;
        d.brow_off = (long long)plan.n_brow;
        plan.n_brow += (size_t)DPW_NB * (size_t)ty[p];
        d.path_off = path;
        path += (long long)tx[p] + ty[p];
    }
    plan.order.resize(n_pairs);
    for (int p = 0; p < n_pairs; ++p) plan.order[p] = p;
    std::stable_sort(plan.order.begin(), plan.order.end(), [&](int a, int b) {
        return (long long)tx[a] * ty[a] > (long long)tx[b] * ty[b];
    });
    return KW_OK;
}

struct DtwWorkspace {
    PairDesc* descs;
    int* order;
    double* xpyr;
    double* ypyr;
    int* rowj;
    uint32_t* bp;
    double* brow;
    double* browm;     // running margins of the boundary rows (margin variants only)
    double2* dist;
    int2* win;
    size_t bytes;
};

static DtwWorkspace carve(const DtwPlan& plan, void* base) {
    Carver c(base);
    DtwWorkspace w;
    w.descs = c.take<PairDesc>(plan.descs.size());
    w.order = c.take<int>(plan.order.size());
    w.xpyr = c.take<double>(plan.n_pyr_x);
    w.ypyr = c.take<double>(plan.n_pyr_y);
    w.rowj = c.take<int>(plan.n_rowj + 1);
    w.bp = c.take<uint32_t>(plan.n_bp);
    w.brow = c.take<double>(plan.n_brow);
    w.browm = c.take<double>(plan.n_brow);
    w.dist = c.take<double2>(plan.n_dist);
    w.win = c.take<int2>(plan.n_win + 1);
    w.bytes = align_up(c.used, 256);
    return w;
}

// ---------------------------------------------------------------------------------------
// Pyramid: level 0 = transpose of the caller's row-major rows into k-major; level l>0 =
// pairwise mean of level l-1.  grid (n_pairs, 2): y = 0 -> x side, 1 -> y side.
// ---------------------------------------------------------------------------------------
__global__ void dtw_pyramid_kernel(const PairDesc* __restrict__ descs, int level, int F,
                                   const double* __restrict__ x_in,
                                   const double* __restrict__ y_in, double* __restrict__ xpyr,
                                   double* __restrict__ ypyr) {
    const PairDesc& d = descs[blockIdx.x];
    if (level >= d.nlev) return;
    const bool is_y = blockIdx.y == 1;
    const int T = is_y ? d.ty[level] : d.tx[level];
    double* out = is_y ? ypyr + d.yoff[level] : xpyr + d.xoff[level];
    if (level == 0) {
        // row-major (T x F) -> k-major (F x T) through a shared tile of 64 rows: both the
        // global reads and the global writes are contiguous runs
        __shared__ double tile[64][33];
        const double* in = is_y ? y_in + d.yrow0 * F : x_in + d.xrow0 * F;
        for (int r0 = 64 * blockIdx.z; r0 < T; r0 += 64 * gridDim.z) {
            const int nr = min(64, T - r0);
            for (int e = threadIdx.x; e < nr * F; e += blockDim.x) {
                const int i = e / F, k = e - i * F;
                tile[i][k] = in[(size_t)r0 * F + e];
            }
            __syncthreads();
            for (int e = threadIdx.x; e < F * 64; e += blockDim.x) {
                const int k = e >> 6, i = e & 63;
                if (i < nr) out[(size_t)k * T + r0 + i] = tile[i][k];
            }
            __syncthreads();
        }
    } else {
        const int Tp = is_y ? d.ty[level - 1] : d.tx[level - 1];
        const double* in = is_y ? ypyr + d.yoff[level - 1] : xpyr + d.xoff[level - 1];
        const long long n = (long long)T * F;
        for (long long e = threadIdx.x + (long long)blockDim.x * blockIdx.z; e < n;
             e += (long long)blockDim.x * gridDim.z) {
            const int k = (int)(e / T), i = (int)(e % T);
            const double a = in[(size_t)k * Tp + 2 * i];
            const double b = in[(size_t)k * Tp + 2 * i + 1];
            out[e] = __dmul_rn(__dadd_rn(a, b), 0.5);
        }
    }
}

// Back-pointers: 2 bits per cell, 16 cells per 32-bit word; the 8 words of rows 8a..8a+7 for one
// 16-column group form one 32-byte sector, so the backtrace (which moves to a neighbouring row
// or column every step) stays inside one L1 sector for ~8-12 steps.  ncg = ceil(ty / 16).
__host__ __device__ __forceinline__ size_t bp_word(int i, int j, int ncg) {
    return ((size_t)(i >> 3) * ncg + (j >> 4)) * 8 + (size_t)(i & 7);
}

template <int P>
__device__ __forceinline__ double dist_acc(double s, double d) {
    if (P == 2) return __fma_rn(d, d, s);
    return __dadd_rn(s, fabs(d));
}
template <int P>
__device__ __forceinline__ double dist_fin(double s) {
    if (P == 2) return __dsqrt_rn(s);
    return s;
}
// fp32 local-distance mode (precision = 1): differences, squares and the square root in fp32,
// the DP sums stay fp64 (so the path cost carries only the rounding of the local distances).
template <int P>
__device__ __forceinline__ float dist_acc(float s, float d) {
    if (P == 2) return __fmaf_rn(d, d, s);
    return __fadd_rn(s, fabsf(d));
}
template <int P>
__device__ __forceinline__ double dist_fin(float s) {
    if (P == 2) return (double)__fsqrt_rn(s);
    return (double)s;
}
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }

// One DP cell.  TIE = 0: fastdtw's pure-Python back-end -- the three candidates are compared
// AFTER the local distance is added and the first minimum wins in the order up (i-1, j), left
// (i, j-1), diagonal.  TIE = 1: the order of its Cython back-end as recalled in SURVEY.md
// section 8a -- the predecessors are compared BEFORE the addition and the diagonal wins ties,
// then left, then up.  Back-pointer codes: 0 = up, 1 = left, 2 = diagonal.
// MARGIN: also the decision margin (runner-up minus winner among the compared values, +inf with
// a single finite candidate) and its running minimum along the winning path, so that
// m(tx-1, ty-1) is the smallest margin of any decision ON THE RETURNED PATH: when it is far above
// the rounding of the sums, neither the tie order nor the last-bit rounding of the local
// distances can change the path.
template <int TIE, bool MARGIN>
__device__ __forceinline__ void dtw_cell(double up, double left, double diag, double dt,
                                         double mup, double mleft, double mdiag, double& v,
                                         uint32_t& code, double& m) {
    double cu, cl, cd;
    if (TIE == 0) {
        cu = __dadd_rn(up, dt);
        cl = __dadd_rn(left, dt);
        cd = __dadd_rn(diag, dt);
        v = cu;
        code = 0u;
        if (cl < v) { v = cl; code = 1u; }
        if (cd < v) { v = cd; code = 2u; }
    } else {
        cu = up;
        cl = left;
        cd = diag;
        double best = cd;
        code = 2u;
        if (cl < best) { best = cl; code = 1u; }
        if (cu < best) { best = cu; code = 0u; }
        v = __dadd_rn(best, dt);
    }
    if (MARGIN) {
        const double win = code == 0u ? cu : (code == 1u ? cl : cd);
        const double other = code == 0u ? fmin(cl, cd) : (code == 1u ? fmin(cu, cd) : fmin(cu, cl));
        const double here = (win < CUDART_INF) ? other - win : CUDART_INF;
        const double along = code == 0u ? mup : (code == 1u ? mleft : mdiag);
        m = fmin(here, along);
    }
}

// ---------------------------------------------------------------------------------------
// Windowed DP for one resolution level.
// ---------------------------------------------------------------------------------------
template <int FP, int NT, int P, typename T, int TIE, bool MARGIN>
__global__ void __launch_bounds__(NT)
dtw_dp_kernel(const PairDesc* __restrict__ descs, const int* __restrict__ order, int level,
              int radius, int F, const double* __restrict__ xpyr,
              const double* __restrict__ ypyr, const int* __restrict__ rowj,
              uint32_t* __restrict__ bp, double* __restrict__ brow, double* __restrict__ cost,
              unsigned long long* __restrict__ cells, int only_full,
              double* __restrict__ browm, double* __restrict__ margin) {
    constexpr int CH = NT;
    constexpr int RING = 2 * NT;
    constexpr int RS = RING + 1;
    extern __shared__ double smem[];
    double* xch = smem;            // 2 * NT: lower-row D values, double-buffered by step parity
    double* brs = xch + 2 * NT;    // RING: previous strip's last row, staged with the y ring
    double* xchm = brs + RING;     // MARGIN: the same two exchanges for the running margins
    double* brsm = xchm + (MARGIN ? 2 * NT : 0);
    T* ys = reinterpret_cast<T*>(brsm + (MARGIN ? RING : 0));   // FP * RS, k-major ring of y columns
    __shared__ unsigned long long cell_count;

    const int pair = order[blockIdx.x];
    const PairDesc& d = descs[pair];
    if (level >= d.nlev) return;
    const int t = threadIdx.x;
    const int tx = d.tx[level], ty = d.ty[level];
    const double* __restrict__ xT = xpyr + d.xoff[level];
    const double* __restrict__ yT = ypyr + d.yoff[level];
    const bool full = (level == d.nlev - 1);
    if (only_full && !full) return;      // banded levels go through dtw_dist / dtw_dpb
    const int ctx = full ? 0 : d.tx[level + 1];
    const int* __restrict__ cfirst = full ? nullptr : rowj + d.rj_off[level + 1];
    const int* __restrict__ clast = full ? nullptr : cfirst + ctx;
    const int tiles_x = (ty + 15) >> 4;
    uint32_t* bp_pair = bp + d.bp_off;
    double* brow_pair = brow + d.brow_off;
    double* browm_pair = MARGIN ? browm + d.brow_off : nullptr;
    const double INF = CUDART_INF;

    auto window = [&](int a, int& lo, int& hi) {
        if (full) {
            lo = 0;
            hi = ty - 1;
            return;
        }
        const int ca = a >> 1;
        int r0 = max(0, ca - radius);
        const int r1 = min(ctx - 1, ca + radius);
        r0 = min(r0, r1);
        lo = max(0, 2 * (cfirst[r0] - radius));
        hi = min(ty - 1, 2 * (clast[r1] + radius) + 1);
    };

    if (t == 0) cell_count = 0ull;
    // zero the padded feature rows of the ring once
    for (int e = t; e < (FP - F) * RS; e += NT) ys[F * RS + e] = (T)0;

    unsigned int my_cells = 0;
    int strip = 0;
    for (int i0 = 0; i0 < tx; i0 += 2 * NT, ++strip) {
        const int ia = i0 + 2 * t, ib = ia + 1;
        int loa = INT_MAX, hia = INT_MIN, lob = INT_MAX, hib = INT_MIN;
        if (ia < tx) window(ia, loa, hia);
        if (ib < tx) window(ib, lob, hib);
        T xa[FP], xb[FP];
#pragma unroll
        for (int k = 0; k < FP; ++k) {
            xa[k] = (T)((k < F && ia < tx) ? xT[(size_t)k * tx + ia] : 0.0);
            xb[k] = (T)((k < F && ib < tx) ? xT[(size_t)k * tx + ib] : 0.0);
        }
        // strip-wide quantities (uniform across the CTA)
        int jstart, dummy;
        window(i0, jstart, dummy);
        const int il = min(tx, i0 + 2 * NT) - 1;
        int hil;
        window(il, dummy, hil);
        const int n_steps = hil - jstart + ((il - i0) >> 1) + 1;
        int plo = INT_MAX, phi = INT_MIN;
        if (i0 > 0) window(i0 - 1, plo, phi);
        const double* brow_in = brow_pair + ((strip & 1) ? 0 : ty);
        double* brow_out = brow_pair + ((strip & 1) ? ty : 0);
        const double* browm_in = MARGIN ? browm_pair + ((strip & 1) ? 0 : ty) : nullptr;
        double* browm_out = MARGIN ? browm_pair + ((strip & 1) ? ty : 0) : nullptr;
        const bool writes_boundary = (t == NT - 1) && (i0 + 2 * NT < tx);

        double va_prev = INF, vb_prev = INF, diag_in = INF;
        double ma_prev = INF, mb_prev = INF, mdiag_in = INF;
        if (t == 0) {
            const int jm = jstart - 1;
            if (i0 == 0) {
                diag_in = (jm == -1) ? 0.0 : INF;  // virtual origin D[0][0] = 0
            } else {
                const bool in = jm >= plo && jm <= phi;
                diag_in = in ? __ldcg(brow_in + jm) : INF;
                if (MARGIN) mdiag_in = in ? __ldcg(browm_in + jm) : INF;
            }
        }
        xch[NT + t] = INF;  // parity 1 is read at step 0
        if (MARGIN) xchm[NT + t] = INF;
        uint32_t wa = 0u, wb = 0u;
        __syncthreads();

        for (int s = 0; s < n_steps; ++s) {
            if ((s % CH) == 0) {
                const int jbase = jstart + s;
                for (int e = t; e < F * CH; e += NT) {
                    const int k = e / CH, jj = e - k * CH;
                    const int j = jbase + jj;
                    if (j < ty) ys[k * RS + (j & (RING - 1))] = (T)yT[(size_t)k * ty + j];
                }
                {
                    const int j = jbase + t;
                    const bool in = i0 > 0 && j >= plo && j <= phi;
                    brs[j & (RING - 1)] = in ? __ldcg(brow_in + j) : INF;
                    if (MARGIN) brsm[j & (RING - 1)] = in ? __ldcg(browm_in + j) : INF;
                }
                __syncthreads();
            }
            const int j = jstart + s - t;
            const double up_in =
                (t == 0) ? brs[j & (RING - 1)] : xch[((s + 1) & 1) * NT + t - 1];
            double mup_in = INF;
            if (MARGIN)
                mup_in = (t == 0) ? brsm[j & (RING - 1)] : xchm[((s + 1) & 1) * NT + t - 1];
            const bool act_a = (j >= loa) && (j <= hia);
            const bool act_b = (j >= lob) && (j <= hib);
            double va = INF, vb = INF, ma = INF, mb = INF;
            if (act_a || act_b) {
                T sa = (T)0, sb = (T)0;
                const T* yp = ys + (j & (RING - 1));
#pragma unroll
                for (int k = 0; k < FP; ++k) {
                    const T yv = yp[k * RS];
                    sa = dist_acc<P>(sa, sub_rn(xa[k], yv));
                    sb = dist_acc<P>(sb, sub_rn(xb[k], yv));
                }
                if (act_a) {
                    const double dt = dist_fin<P>(sa);
                    uint32_t code;
                    dtw_cell<TIE, MARGIN>(up_in, va_prev, diag_in, dt, mup_in, ma_prev, mdiag_in,
                                          va, code, ma);
                    wa |= code << (2 * (j & 15));
                    if ((j & 15) == 15 || j == hia) { bp_pair[bp_word(ia, j, tiles_x)] = wa; wa = 0u; }
                    if (ia == tx - 1 && j == ty - 1) {
                        cost[pair] = va;
                        if (MARGIN) {
                            if (level == 0) margin[2 * pair] = ma;
                            margin[2 * pair + 1] = fmin(margin[2 * pair + 1], ma);
                        }
                    }
                    ++my_cells;
                }
                if (act_b) {
                    const double dt = dist_fin<P>(sb);
                    uint32_t code;
                    dtw_cell<TIE, MARGIN>(va, vb_prev, va_prev, dt, ma, mb_prev, ma_prev, vb, code,
                                          mb);
                    wb |= code << (2 * (j & 15));
                    if ((j & 15) == 15 || j == hib) { bp_pair[bp_word(ib, j, tiles_x)] = wb; wb = 0u; }
                    if (ib == tx - 1 && j == ty - 1) {
                        cost[pair] = vb;
                        if (MARGIN) {
                            if (level == 0) margin[2 * pair] = mb;
                            margin[2 * pair + 1] = fmin(margin[2 * pair + 1], mb);
                        }
                    }
                    if (writes_boundary) {
                        __stcg(brow_out + j, vb);
                        if (MARGIN) __stcg(browm_out + j, mb);
                    }
                    ++my_cells;
                }
            }
            xch[(s & 1) * NT + t] = vb;
            if (MARGIN) xchm[(s & 1) * NT + t] = mb;
            va_prev = va;
            vb_prev = vb;
            diag_in = up_in;
            if (MARGIN) {
                ma_prev = ma;
                mb_prev = mb;
                mdiag_in = mup_in;
            }
            __syncthreads();
        }
    }
    if (cells != nullptr) {
        atomicAdd(&cell_count, (unsigned long long)my_cells);
        __syncthreads();
        if (t == 0) atomicAdd(cells + pair, cell_count);
    }
}

// ---------------------------------------------------------------------------------------
// Banded levels, part 1: local distances of every window cell.  A CTA takes 8 row pairs (rows
// 2q and 2q+1 share their window) and the union of their windows; a thread owns one column of
// that union and keeps the 16 running sums in registers, so a y value is loaded once per 16
// cells and the x rows are shared-memory broadcasts.  Also records each row pair's window and
// counts the level's cells.  grid (n_pairs, ceil(row pairs / 8)), 256 threads.
// ---------------------------------------------------------------------------------------
template <int FP, int P, typename T, bool FULLF>
__global__ void __launch_bounds__(256)
dtw_dist_kernel(const PairDesc* __restrict__ descs, int level, int radius, int F, int wcap,
                const double* __restrict__ xpyr, const double* __restrict__ ypyr,
                const int* __restrict__ rowj, int2* __restrict__ win, double2* __restrict__ dist,
                unsigned long long* __restrict__ cells) {
    __shared__ __align__(16) T xs[FP][8][2];
    __shared__ int los[8], his[8];
    const PairDesc& d = descs[blockIdx.x];
    if (level >= d.nlev - 1) return;
    const int tx = d.tx[level], ty = d.ty[level];
    const int n_rp = (tx + 1) >> 1;
    const int rp0 = blockIdx.y * 8;
    if (rp0 >= n_rp) return;
    const int nrp = min(8, n_rp - rp0);
    const double* __restrict__ xT = xpyr + d.xoff[level];
    const double* __restrict__ yT = ypyr + d.yoff[level];
    const int tid = threadIdx.x;
    if (tid < 8) {
        const int ctx = d.tx[level + 1];
        const int* __restrict__ cfirst = rowj + d.rj_off[level + 1];
        const int* __restrict__ clast = cfirst + ctx;
        const int rp = min(rp0 + tid, n_rp - 1);
        int r0 = max(0, rp - radius);
        const int r1 = min(ctx - 1, rp + radius);
        r0 = min(r0, r1);
        const int lo = max(0, 2 * (cfirst[r0] - radius));
        const int hi = min(ty - 1, 2 * (clast[r1] + radius) + 1);
        los[tid] = lo;
        his[tid] = hi;
        if (tid < nrp) {
            win[d.win_off + rp] = make_int2(lo, hi);
            if (cells != nullptr) {
                const int rows = (2 * rp + 1 < tx) ? 2 : 1;
                atomicAdd(cells + blockIdx.x, (unsigned long long)rows * (unsigned)(hi - lo + 1));
            }
        }
    }
    for (int e = tid; e < FP * 16; e += 256) {
        const int k = e >> 4, r = (e >> 1) & 7, ab = e & 1;
        const int i = 2 * (rp0 + r) + ab;
        xs[k][r][ab] = (T)((k < F && i < tx) ? xT[(size_t)k * tx + i] : 0.0);
    }
    __syncthreads();
    const int LO = los[0], HI = his[nrp - 1];
    double2* out = dist + d.dist_off + (size_t)rp0 * wcap;
    for (int j = LO + tid; j <= HI; j += 256) {
        T acc[8][2];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r][0] = acc[r][1] = (T)0;
#pragma unroll
        for (int k = 0; k < FP; ++k) {
            if (FULLF || k < F) {       // FULLF: F == FP, no per-coefficient test in the loop
                const T yv = (T)__ldg(yT + (size_t)k * ty + j);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    acc[r][0] = dist_acc<P>(acc[r][0], sub_rn(xs[k][r][0], yv));
                    acc[r][1] = dist_acc<P>(acc[r][1], sub_rn(xs[k][r][1], yv));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r < nrp) {
                const int lo = los[r], hi = his[r];
                if (j >= lo && j <= hi && j - lo < wcap)
                    out[(size_t)r * wcap + (j - lo)] =
                        make_double2(dist_fin<P>(acc[r][0]), dist_fin<P>(acc[r][1]));
            }
        }
    }
}

// D values are sums of non-negative terms (or +inf): their IEEE bit patterns order like the
// values, so the "strictly less" of the cell rule runs on the integer pipe.
__device__ __forceinline__ bool lt_nonneg(double a, double b) {
    return __double_as_longlong(a) < __double_as_longlong(b);
}

// Local distance of one cell straight from the pyramids: the rare columns of a window that is
// wider than the stored slice.  Kept out of line so the sweep's loop stays small.
template <int P, typename T>
__device__ __noinline__ double slow_dist(const double* __restrict__ xT,
                                         const double* __restrict__ yT, int tx, int ty, int F,
                                         int i, int j) {
    T sacc = (T)0;
    for (int k = 0; k < F; ++k)
        sacc = dist_acc<P>(sacc, sub_rn((T)xT[(size_t)k * tx + i], (T)yT[(size_t)k * ty + j]));
    return dist_fin<P>(sacc);
}

// ---------------------------------------------------------------------------------------
// Banded levels, part 2: the sweep over the stored local distances, one WARP per pair.
// What is left per cell is 3 additions and 2 comparisons on a chain that runs through every
// anti-diagonal, so the sweep is bound by per-step latency, not throughput: the systolic
// wavefront of dtw_dp_kernel (lane t owns rows i0+2t, i0+2t+1 of a 64-row strip and lags t
// columns) is kept, but the neighbour's value comes through a shuffle instead of shared memory
// plus a CTA barrier, the previous strip's boundary row is read 32 columns at a time into a
// register per lane, and the two distances of a step arrive through a per-lane shared-memory
// ring filled by cp.async 24 steps ahead (the slices were just written by dtw_dist_kernel and
// mostly live in DRAM).
// ---------------------------------------------------------------------------------------
template <int P, typename T, int TIE, bool MARGIN>
__global__ void __launch_bounds__(32 * DPW_NW)
dtw_dpw_kernel(const PairDesc* __restrict__ descs, const int* __restrict__ order, int level,
               int F, int wcap, const double* __restrict__ xpyr, const double* __restrict__ ypyr,
               const int2* __restrict__ win, const double2* __restrict__ dist,
               uint32_t* __restrict__ bp, double* __restrict__ brow, double* __restrict__ cost,
               double* __restrict__ browm, double* __restrict__ margin) {
    constexpr int RD = 16, PD = 12;       // ring slots per lane, prefetch distance (steps)
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ double2 ring_all[DPW_NW][RD][32];   // local distances in flight: [step % RD][lane]
    __shared__ unsigned long long prog_s[DPW_NB];  // (strip << 32 | last finished boundary column + 1)
    __shared__ double bch_all[DPW_NW][32];         // boundary-row chunk of each warp, for its lane 0
    __shared__ double bchm_all[MARGIN ? DPW_NW : 1][32];   // ... and its running margins
    const int pair = order[blockIdx.x];
    const PairDesc& d = descs[pair];
    if (level >= d.nlev - 1) return;
    const int t = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2 (*ring)[32] = ring_all[warp];
    double* bch = bch_all[warp];
    double* bchm = bchm_all[MARGIN ? warp : 0];
    volatile unsigned long long* prog = prog_s;
    if (threadIdx.x < DPW_NB) prog_s[threadIdx.x] = 0ull;
    __syncthreads();
    const int tx = d.tx[level], ty = d.ty[level];
    const double* __restrict__ xT = xpyr + d.xoff[level];
    const double* __restrict__ yT = ypyr + d.yoff[level];
    const int2* __restrict__ winp = win + d.win_off;
    const double2* __restrict__ dist_pair = dist + d.dist_off;
    const int tiles_x = (ty + 15) >> 4;
    uint32_t* bp_pair = bp + d.bp_off;
    double* brow_pair = brow + d.brow_off;
    double* browm_pair = MARGIN ? browm + d.brow_off : nullptr;
    const double INF = CUDART_INF;

    // Strips are pipelined over the CTA's warps: warp w takes strips w, w + NW, ...; strip k
    // starts as soon as strip k-1 has finished the first 32 columns of its last row, and keeps
    // checking the producer's progress once per 32 columns.
    const unsigned ring_base = (unsigned)__cvta_generic_to_shared(&ring[0][t]);
    const int n_strips = (tx + 63) >> 6;
    for (int strip = warp; strip < n_strips; strip += DPW_NW) {
        const int i0 = strip << 6;
        const int ia = i0 + 2 * t, ib = ia + 1;
        const bool has_b = ib < tx;
        int lo = 0, hi = -1;                 // lanes past the last row: empty window
        if (ia < tx) {
            const int2 w = winp[ia >> 1];
            lo = w.x;
            hi = w.y;
        }
        const int width = hi - lo + 1, lim = min(width, wcap);
        const double2* drow = dist_pair + (size_t)(ia >> 1) * wcap;
        const int il = min(tx, i0 + 64) - 1;
        const int tl = (il - i0) >> 1;                        // lane of the strip's last row
        const int jstart = __shfl_sync(FULL, lo, 0);
        const int hil = __shfl_sync(FULL, hi, tl);
        const int n_steps = (hil - jstart + tl + 1 + 3) & ~3;   // padded steps touch no window
        int plo = INT_MAX, phi = INT_MIN;
        if (i0 > 0) {
            const int2 w = winp[(i0 - 1) >> 1];
            plo = w.x;
            phi = w.y;
        }
        const double* brow_in = brow_pair + (size_t)((strip + DPW_NB - 1) % DPW_NB) * ty;
        double* brow_out = brow_pair + (size_t)(strip % DPW_NB) * ty;
        const double* browm_in =
            MARGIN ? browm_pair + (size_t)((strip + DPW_NB - 1) % DPW_NB) * ty : nullptr;
        double* browm_out = MARGIN ? browm_pair + (size_t)(strip % DPW_NB) * ty : nullptr;
        const bool writes_boundary = (t == 31) && (i0 + 64 < tx);
        volatile unsigned long long* prog_in = prog + (strip + DPW_NB - 1) % DPW_NB;
        volatile unsigned long long* prog_out = prog + strip % DPW_NB;
        uint32_t* bpa = bp_pair + bp_word(ia, 0, tiles_x);    // + 8 * (j >> 4); row ib: + 1
        const int col0 = jstart - t - lo;                     // window-relative column at step 0
        // wait until the previous strip's last row is final up to column `col`
        auto wait_boundary = [&](int col) {
            if (i0 > 0) {
                col = min(col, phi);
                if (col >= plo) {
                    const unsigned long long need =
                        ((unsigned long long)(unsigned)(strip - 1) << 32) | (unsigned)(col + 1);
                    const long long t0 = clock64();
                    while (*prog_in < need) {
                        __nanosleep(100);       // leave the issue slots to the producers
                        if (clock64() - t0 > 4000000000LL) __trap();   // protocol bug: never hang
                    }
                }
            }
        };
        // This lane's two distances for step s go to ring slot s % RD: an asynchronous copy of
        // the stored value, +inf outside the window (so the cell code needs no activity test),
        // computed in place for the columns of a window wider than the stored slice.
        const bool wide = width > wcap;
        auto prefetch = [&](int s) {
            const int c = col0 + s;
            const unsigned dst = ring_base + ((unsigned)(s & (RD - 1)) << 9);
            const bool v = (unsigned)c < (unsigned)lim;
            asm volatile(
                "{\n .reg .pred p;\n setp.ne.b32 p, %0, 0;\n"
                " @p cp.async.cg.shared.global [%1], [%2], 16;\n"
                " @!p st.shared.v2.f64 [%1], {%3, %3};\n}"
                ::"r"((int)v), "r"(dst), "l"(drow + c), "d"(INF) : "memory");
            if (wide && !v && (unsigned)c < (unsigned)width) {
                const double da = slow_dist<P, T>(xT, yT, tx, ty, F, ia, lo + c);
                const double db = has_b ? slow_dist<P, T>(xT, yT, tx, ty, F, ib, lo + c) : INF;
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(dst), "d"(da), "d"(db)
                             : "memory");
            }
        };
        // back-pointer word of the 16-column group ending at column e, from a 64-bit shift
        // register that holds the codes of columns jl-31 .. jl (column jl in the top two bits)
        auto store_words = [&](int e, int jl, unsigned long long qa, unsigned long long qb) {
            const int sh = 32 - 2 * (jl - e);          // e in [jl - 15, jl + 15]
            uint32_t* wp = bpa + ((e >> 4) << 3);
            wp[0] = (uint32_t)(qa >> sh);
            if (has_b) wp[1] = (uint32_t)(qb >> sh);
        };

        double va_prev = INF, vb_prev = INF, diag_in = INF;
        double ma_prev = INF, mb_prev = INF, mdiag_in = INF;
        wait_boundary(jstart + 31);
        if (t == 0) {
            const int jm = jstart - 1;
            if (i0 == 0) {
                diag_in = (jm == -1) ? 0.0 : INF;  // virtual origin D[0][0] = 0
            } else {
                const bool in = jm >= plo && jm <= phi;
                diag_in = in ? __ldcg(brow_in + jm) : INF;
                if (MARGIN) mdiag_in = in ? __ldcg(browm_in + jm) : INF;
            }
        }
        unsigned long long qa = 0ull, qb = 0ull;
        double ha[4], hb[4];                  // D of the last four steps (for the final cell)
        double hma[4], hmb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) ha[q] = hb[q] = hma[q] = hmb[q] = INF;
        int e_done = -1;                      // last column group already stored
        __syncwarp();                         // the previous strip's reads of the ring are done
#pragma unroll
        for (int g = 0; g < PD / 4; ++g) {
#pragma unroll
            for (int q = 0; q < 4; ++q) prefetch(4 * g + q);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }

        for (int s0 = 0; s0 < n_steps; s0 += 4) {
            if ((s0 & 31) == 0) {
                // boundary row of the previous strip, columns jstart+s0 .. +31 (lane = column),
                // staged for lane 0
                if (s0 > 0) wait_boundary(jstart + s0 + 31);
                const int j = jstart + s0 + t;
                const bool in = i0 > 0 && j >= plo && j <= phi;
                __syncwarp();
                bch[t] = in ? __ldcg(brow_in + j) : INF;
                if (MARGIN) bchm[t] = in ? __ldcg(browm_in + j) : INF;
                __syncwarp();
            }
            asm volatile("cp.async.wait_group %0;" ::"n"(PD / 4 - 1) : "memory");
#pragma unroll
            for (int q = 0; q < 4; ++q) prefetch(s0 + PD + q);
            asm volatile("cp.async.commit_group;" ::: "memory");
            // this iteration's distances and (for lane 0) boundary values, ahead of the chain
            const unsigned slot0 = ring_base + ((unsigned)(s0 & (RD - 1)) << 9);
            double da[4], db[4], bq[4], bqm[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(da[q]), "=d"(db[q])
                             : "r"(slot0 + (unsigned)q * 512u) : "memory");
                bq[q] = bch[(s0 + q) & 31];
                bqm[q] = MARGIN ? bchm[(s0 + q) & 31] : INF;
            }
            const int cbase = col0 + s0;              // window-relative column of step s0
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double up_in = __shfl_up_sync(FULL, vb_prev, 1);
                up_in = (t == 0) ? bq[q] : up_in;
                double mup_in = INF;
                if (MARGIN) {
                    mup_in = __shfl_up_sync(FULL, mb_prev, 1);
                    mup_in = (t == 0) ? bqm[q] : mup_in;
                }
                double va, vb, ma = INF, mb = INF;
                uint32_t ca, cb;
                dtw_cell<TIE, MARGIN>(up_in, va_prev, diag_in, da[q], mup_in, ma_prev, mdiag_in, va,
                                      ca, ma);
                qa = (qa >> 2) | ((unsigned long long)ca << 62);
                dtw_cell<TIE, MARGIN>(va, vb_prev, va_prev, db[q], ma, mb_prev, ma_prev, vb, cb, mb);
                qb = (qb >> 2) | ((unsigned long long)cb << 62);
                if (writes_boundary && (unsigned)(cbase + q) < (unsigned)width) {
                    __stcg(brow_out + lo + cbase + q, vb);
                    if (MARGIN) __stcg(browm_out + lo + cbase + q, mb);
                }
                ha[q] = va;
                hb[q] = vb;
                va_prev = va;
                vb_prev = vb;
                diag_in = up_in;
                if (MARGIN) {
                    hma[q] = ma;
                    hmb[q] = mb;
                    ma_prev = ma;
                    mb_prev = mb;
                    mdiag_in = mup_in;
                }
            }
            if ((s0 & 15) == 12) {
                // every 16 steps (all lanes together): store the column group that was
                // completed since the last time, and publish the boundary row's progress
                const int jl = lo + cbase + 3;
                const int e = jl - ((jl + 1) & 15);
                if (e >= lo && e - 15 <= hi) store_words(e, jl, qa, qb);
                e_done = e;
                if (writes_boundary) {
                    const int jp = min(jl, hi);
                    if (jp >= lo) {
                        __threadfence_block();
                        *prog_out = ((unsigned long long)(unsigned)strip << 32) | (unsigned)(jp + 1);
                    }
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        {
            // column groups not stored yet: the last complete one and the partial one after it
            const int jl = lo + col0 + n_steps - 1;
            const int e = jl - ((jl + 1) & 15);
            if (width > 0) {
                if (e > e_done && e >= lo && e - 15 <= hi) store_words(e, jl, qa, qb);
                if (e != jl && hi > e) store_words(e + 16, jl, qa, qb);
            }
            if (writes_boundary) {
                __threadfence_block();
                *prog_out = ((unsigned long long)(unsigned)strip << 32) | (unsigned)(hi + 1);
            }
        }
        // D[tx-1][ty-1]: the last row's last column, reached n_steps - 1 - (padding) steps in
        if (strip == n_strips - 1 && t == tl) {
            const int pad = (n_steps - 1) - (hil - jstart + tl);      // 0..3
            double fa = ha[3], fb = hb[3], fma_ = hma[3], fmb_ = hmb[3];
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (pad == 3 - q) { fa = ha[q]; fb = hb[q]; fma_ = hma[q]; fmb_ = hmb[q]; }
            cost[pair] = (il == ia) ? fa : fb;
            if (MARGIN) {
                const double mfin = (il == ia) ? fma_ : fmb_;
                if (level == 0) margin[2 * pair] = mfin;
                margin[2 * pair + 1] = fmin(margin[2 * pair + 1], mfin);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Backtrace: one warp per pair walks the 2-bit codes from (tx-1, ty-1) to the origin.  The warp
// keeps a 2 x 2 window of back-pointer sectors (16 rows x 32 columns, one word per lane) in
// registers and looks codes up with shuffles; global memory is touched once per window, not
// once per step.  Level 0 writes the path (backwards, into the tail of the pair's region);
// levels >= 1 only record first_j / last_j per row for the next finer level's window.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
dtw_backtrace_kernel(const PairDesc* __restrict__ descs, int n_pairs, int level,
                     const uint32_t* __restrict__ bp, int* __restrict__ rowj,
                     int32_t* __restrict__ path, int32_t* __restrict__ path_begin,
                     int32_t* __restrict__ path_len) {
    const int lane = threadIdx.x & 31;
    const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= n_pairs) return;
    const PairDesc& d = descs[pair];
    if (level >= d.nlev) return;
    const int tx = d.tx[level], ty = d.ty[level];
    const int ncg = (ty + 15) >> 4;
    const uint32_t* bpp = bp + d.bp_off;
    int* first = (level > 0) ? rowj + d.rj_off[level] : nullptr;
    int* last = (level > 0) ? first + tx : nullptr;
    const int cap = d.tx[0] + d.ty[0];
    int32_t* out = path + 2 * d.path_off;
    const int q = lane >> 3, r = lane & 7;      // window sector and row-in-sector of this lane
    int i = tx - 1, j = ty - 1, n = 0, prev_i = -1, prev_j = -1;
    int si = -1, sj = -1;                        // anchor (bottom-right sector) of the window
    int bi = 0, bj = 0;                          // path point buffered by this lane (level 0)
    uint32_t w = 0u;
    while (i >= 0 && j >= 0 && n < tx + ty) {
        const int ci = i >> 3, cj = j >> 4;
        if (si < 0 || si - ci > 1 || sj - cj > 1 || ci > si || cj > sj) {
            si = ci;
            sj = cj;
            const int ii = si - (q & 1), jj = sj - (q >> 1);
            w = (ii >= 0 && jj >= 0) ? __ldcg(bpp + ((size_t)ii * ncg + jj) * 8 + r) : 0u;
        }
        if (level == 0) {
            // lane n % 32 keeps the point; every 32 steps the warp stores its 32 points at once
            if (lane == (n & 31)) { bi = i; bj = j; }
            if ((n & 31) == 31)
                reinterpret_cast<int2*>(out)[cap - 1 - (n - 31 + lane)] = make_int2(bi, bj);
        } else if (lane == 0 && i != prev_i) {
            last[i] = j;
            if (prev_i >= 0) first[prev_i] = prev_j;
        }
        prev_i = i;
        prev_j = j;
        const int src = (((si - ci) + 2 * (sj - cj)) << 3) + (i & 7);
        const uint32_t word = __shfl_sync(0xffffffffu, w, src);
        const uint32_t code = (word >> (2 * (j & 15))) & 3u;
        if (code == 0u) {
            --i;
        } else if (code == 1u) {
            --j;
        } else {
            --i;
            --j;
        }
        ++n;
    }
    if (level == 0 && lane < (n & 31))       // the points since the last full group of 32
        reinterpret_cast<int2*>(out)[cap - 1 - ((n & ~31) + lane)] = make_int2(bi, bj);
    if (lane == 0) {
        if (level > 0 && prev_i >= 0) first[prev_i] = prev_j;
        if (level == 0) {
            path_begin[pair] = cap - n;
            path_len[pair] = n;
        }
    }
}

// Variant of the cell rule and of what is carried along (see dtw_cell).
struct SweepOpts {
    int tie;            // 0 = pure-Python order, 1 = Cython order
    double* margin;     // nullptr, or 2 doubles per pair (level 0, minimum over all levels)
};

__global__ void dtw_fill_kernel(double* __restrict__ p, long long n, double v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

template <int FP, int NT, int P, typename T, int TIE, bool MARGIN>
static int launch_dp_v(int n_pairs, const DtwWorkspace& w, int level, int radius, int F,
                       double* cost, unsigned long long* cells, int only_full, double* margin,
                       cudaStream_t st) {
    constexpr int RS = 2 * NT + 1;
    const size_t smem = sizeof(double) * (size_t)(2 * NT + 2 * NT) * (MARGIN ? 2 : 1) +
                        sizeof(T) * (size_t)FP * RS;
    auto kern = dtw_dp_kernel<FP, NT, P, T, TIE, MARGIN>;
    KW_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    kern<<<n_pairs, NT, smem, st>>>(w.descs, w.order, level, radius, F, w.xpyr, w.ypyr, w.rowj,
                                    w.bp, w.brow, cost, cells, only_full, w.browm, margin);
    KW_CUDA_CHECK(cudaGetLastError());
    return KW_OK;
}

// The margin variants exist for exact (double) local distances only.
template <int FP, int NT, int P, typename T>
static int launch_dp(int n_pairs, const DtwWorkspace& w, int level, int radius, int F,
                     double* cost, unsigned long long* cells, int only_full, SweepOpts o,
                     cudaStream_t st) {
    if constexpr (std::is_same<T, double>::value) {
        if (o.margin != nullptr) {
            if (o.tie == 0)
                return launch_dp_v<FP, NT, P, T, 0, true>(n_pairs, w, level, radius, F, cost, cells,
                                                          only_full, o.margin, st);
            return launch_dp_v<FP, NT, P, T, 1, true>(n_pairs, w, level, radius, F, cost, cells,
                                                      only_full, o.margin, st);
        }
    }
    if (o.tie == 0)
        return launch_dp_v<FP, NT, P, T, 0, false>(n_pairs, w, level, radius, F, cost, cells,
                                                   only_full, nullptr, st);
    return launch_dp_v<FP, NT, P, T, 1, false>(n_pairs, w, level, radius, F, cost, cells,
                                               only_full, nullptr, st);
}

template <int FP, int P, typename T>
static int launch_dp_nt(int nt, int n_pairs, const DtwWorkspace& w, int level, int radius, int F,
                        double* cost, unsigned long long* cells, int only_full, SweepOpts o,
                        cudaStream_t st) {
    if (nt == 32)
        return launch_dp<FP, 32, P, T>(n_pairs, w, level, radius, F, cost, cells, only_full, o, st);
    if (nt == 64)
        return launch_dp<FP, 64, P, T>(n_pairs, w, level, radius, F, cost, cells, only_full, o, st);
    return launch_dp<FP, 128, P, T>(n_pairs, w, level, radius, F, cost, cells, only_full, o, st);
}

template <int P, typename T>
static int launch_dp_fp(int F, int nt, int n_pairs, const DtwWorkspace& w, int level, int radius,
                        double* cost, unsigned long long* cells, int only_full, SweepOpts o,
                        cudaStream_t st) {
    if (F <= 8)
        return launch_dp_nt<8, P, T>(nt, n_pairs, w, level, radius, F, cost, cells, only_full, o, st);
    if (F <= 16)
        return launch_dp_nt<16, P, T>(nt, n_pairs, w, level, radius, F, cost, cells, only_full, o, st);
    if (F <= 26)
        return launch_dp_nt<26, P, T>(nt, n_pairs, w, level, radius, F, cost, cells, only_full, o, st);
    return launch_dp_nt<32, P, T>(nt, n_pairs, w, level, radius, F, cost, cells, only_full, o, st);
}

// Banded level: local distances, then the sweep over them.
template <int FP, int P, typename T>
static int launch_dist(int n_pairs, const DtwWorkspace& w, int level, int radius, int F, int wcap,
                       int max_tx, unsigned long long* cells, cudaStream_t st) {
    const int rp_blocks = ((max_tx + 1) / 2 + 7) / 8;
    if (F == FP)
        dtw_dist_kernel<FP, P, T, true><<<dim3(n_pairs, rp_blocks), 256, 0, st>>>(
            w.descs, level, radius, F, wcap, w.xpyr, w.ypyr, w.rowj, w.win, w.dist, cells);
    else
        dtw_dist_kernel<FP, P, T, false><<<dim3(n_pairs, rp_blocks), 256, 0, st>>>(
            w.descs, level, radius, F, wcap, w.xpyr, w.ypyr, w.rowj, w.win, w.dist, cells);
    KW_CUDA_CHECK(cudaGetLastError());
    return KW_OK;
}

template <int P, typename T, int TIE, bool MARGIN>
static void launch_dpw_v(int n_pairs, const DtwWorkspace& w, int level, int F, int wcap,
                         double* cost, double* margin, cudaStream_t st) {
    dtw_dpw_kernel<P, T, TIE, MARGIN><<<n_pairs, 32 * DPW_NW, 0, st>>>(
        w.descs, w.order, level, F, wcap, w.xpyr, w.ypyr, w.win, w.dist, w.bp, w.brow, cost,
        w.browm, margin);
}

template <int P, typename T>
static int launch_banded(int F, int n_pairs, const DtwWorkspace& w, int level, int radius,
                         int wcap, int max_tx, double* cost, unsigned long long* cells,
                         SweepOpts o, cudaStream_t st) {
    int rc;
    if (F <= 8) rc = launch_dist<8, P, T>(n_pairs, w, level, radius, F, wcap, max_tx, cells, st);
    else if (F <= 16) rc = launch_dist<16, P, T>(n_pairs, w, level, radius, F, wcap, max_tx, cells, st);
    else if (F <= 26) rc = launch_dist<26, P, T>(n_pairs, w, level, radius, F, wcap, max_tx, cells, st);
    else rc = launch_dist<32, P, T>(n_pairs, w, level, radius, F, wcap, max_tx, cells, st);
    if (rc != KW_OK) return rc;
    bool done = false;
    if constexpr (std::is_same<T, double>::value) {
        if (o.margin != nullptr) {
            if (o.tie == 0) launch_dpw_v<P, T, 0, true>(n_pairs, w, level, F, wcap, cost, o.margin, st);
            else launch_dpw_v<P, T, 1, true>(n_pairs, w, level, F, wcap, cost, o.margin, st);
            done = true;
        }
    }
    if (!done) {
        if (o.tie == 0) launch_dpw_v<P, T, 0, false>(n_pairs, w, level, F, wcap, cost, nullptr, st);
        else launch_dpw_v<P, T, 1, false>(n_pairs, w, level, F, wcap, cost, nullptr, st);
    }
    KW_CUDA_CHECK(cudaGetLastError());
    return KW_OK;
}

}  // namespace kw

using namespace kw;

extern "C" size_t kw_dtw_workspace_bytes(int n_pairs, const int32_t* tx_host,
                                         const int32_t* ty_host, int feat_dim, int radius) {
    if (n_pairs <= 0 || feat_dim <= 0) return 0;
    DtwPlan plan;
    if (make_plan(n_pairs, tx_host, ty_host, radius, feat_dim, plan) != KW_OK) return 0;
    return carve(plan, nullptr).bytes;
}

extern "C" int kw_dtw_batch(int n_pairs, const double* x_dev, const double* y_dev,
                            const int32_t* tx_host, const int32_t* ty_host, int feat_dim,
                            int radius, int p_norm, int precision, int tie_mode,
                            double* cost_dev, int32_t* path_dev, int32_t* path_begin_dev,
                            int32_t* path_len_dev, int64_t* cells_dev, double* margin_dev,
                            void* workspace_dev, size_t workspace_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_pairs == 0) return KW_OK;
    KW_REQUIRE(n_pairs > 0, "n_pairs must be >= 0");
    KW_REQUIRE(feat_dim > 0, "feat_dim must be positive");
    KW_REQUIRE(p_norm == 1 || p_norm == 2, "p_norm must be 1 or 2 (got %d)", p_norm);
    if (feat_dim > 32) {
        set_error("feat_dim %d > 32 is not supported by the sm_100a DTW kernels", feat_dim);
        return KW_ERR_UNSUPPORTED;
    }
    KW_REQUIRE(precision == 0 || precision == 1,
               "DTW precision must be 0 (fp64 exact) or 1 (fp32 local distances)");
    KW_REQUIRE(tie_mode == 0 || tie_mode == 1,
               "DTW tie_mode must be 0 (pure-Python order) or 1 (Cython order)");
    if (margin_dev != nullptr && precision != 0) {
        set_error("decision margins are reported for precision 0 (exact local distances) only");
        return KW_ERR_UNSUPPORTED;
    }
    DtwPlan plan;
    int rc = make_plan(n_pairs, tx_host, ty_host, radius, feat_dim, plan);
    if (rc != KW_OK) return rc;
    DtwWorkspace w = carve(plan, workspace_dev);
    if (w.bytes > workspace_bytes) {
        set_error("DTW workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
        return KW_ERR_WORKSPACE;
    }
    KW_CUDA_CHECK(cudaMemcpyAsync(w.descs, plan.descs.data(), sizeof(PairDesc) * n_pairs,
                                  cudaMemcpyHostToDevice, st));
    KW_CUDA_CHECK(cudaMemcpyAsync(w.order, plan.order.data(), sizeof(int) * n_pairs,
                                  cudaMemcpyHostToDevice, st));
    if (cells_dev != nullptr)
        KW_CUDA_CHECK(cudaMemsetAsync(cells_dev, 0, sizeof(int64_t) * n_pairs, st));
    if (margin_dev != nullptr) {
        dtw_fill_kernel<<<(2 * n_pairs + 255) / 256, 256, 0, st>>>(margin_dev, 2LL * n_pairs,
                                                                  std::numeric_limits<double>::infinity());
        KW_CUDA_CHECK(cudaGetLastError());
    }
    const SweepOpts opts{tie_mode, margin_dev};
    for (int l = 0; l < plan.maxlev; ++l) {
        // (pair, x | y, slice of the rows): one block per sequence walked its 64-row tiles one
        // after the other, two barriers each -- 0.1 ms of latency at level 0
        const int longest = std::max(plan.level_max_tx[l], plan.level_max_ty[l]);
        const int slices = std::max(1, std::min(16, (longest + 63) / 64));
        dtw_pyramid_kernel<<<dim3(n_pairs, 2, slices), 256, 0, st>>>(w.descs, l, feat_dim, x_dev,
                                                                     y_dev, w.xpyr, w.ypyr);
        KW_CUDA_CHECK(cudaGetLastError());
    }
    for (int l = plan.maxlev - 1; l >= 0; --l) {
        const int mtx = plan.level_max_tx[l];
        const int nt = (mtx <= 64) ? 32 : ((mtx <= 128 || radius >= 0) ? 64 : 128);
        unsigned long long* cells = reinterpret_cast<unsigned long long*>(cells_dev);
        const bool split = plan.wcap > 0;
        if (split && plan.level_has_band[l]) {
            if (precision == 0) {
                if (p_norm == 2)
                    rc = launch_banded<2, double>(feat_dim, n_pairs, w, l, radius, plan.wcap, mtx,
                                                  cost_dev, cells, opts, st);
                else
                    rc = launch_banded<1, double>(feat_dim, n_pairs, w, l, radius, plan.wcap, mtx,
                                                  cost_dev, cells, opts, st);
            } else {
                if (p_norm == 2)
                    rc = launch_banded<2, float>(feat_dim, n_pairs, w, l, radius, plan.wcap, mtx,
                                                 cost_dev, cells, opts, st);
                else
                    rc = launch_banded<1, float>(feat_dim, n_pairs, w, l, radius, plan.wcap, mtx,
                                                 cost_dev, cells, opts, st);
            }
            if (rc != KW_OK) return rc;
        }
        if (!split || plan.level_has_full[l]) {
            const int only_full = split ? 1 : 0;
            if (precision == 0) {
                if (p_norm == 2)
                    rc = launch_dp_fp<2, double>(feat_dim, nt, n_pairs, w, l, radius, cost_dev,
                                                 cells, only_full, opts, st);
                else
                    rc = launch_dp_fp<1, double>(feat_dim, nt, n_pairs, w, l, radius, cost_dev,
                                                 cells, only_full, opts, st);
            } else {
                if (p_norm == 2)
                    rc = launch_dp_fp<2, float>(feat_dim, nt, n_pairs, w, l, radius, cost_dev,
                                                cells, only_full, opts, st);
                else
                    rc = launch_dp_fp<1, float>(feat_dim, nt, n_pairs, w, l, radius, cost_dev,
                                                cells, only_full, opts, st);
            }
        }
        if (rc != KW_OK) return rc;
        dtw_backtrace_kernel<<<(n_pairs + 3) / 4, 128, 0, st>>>(
            w.descs, n_pairs, l, w.bp, w.rowj, path_dev, path_begin_dev, path_len_dev);
        KW_CUDA_CHECK(cudaGetLastError());
    }
    // the host-side plan vectors were consumed by the (staged) async copies above
    return KW_OK;
}

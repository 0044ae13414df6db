// Conversion: hard posterior, conditional Gaussian, MLPG -- sm_100a, FP64.
//
// Replaces nnmnkwii.baseline.gmm.MLPG(gmm, windows, diff).transform(src) as called at
// kwiiyatta/converter/gmm.py:28-34 (restated in oracle/mlpg_ref.py):
//   m_t = argmax_k log N(x_t; mu_x, S_xx) + log w          (px.predict)
//   E_t = mu_y[m] + S_yx[m] S_xx[m]^-1 (x_t - mu_x[m])       = offset[m] + A[m] x_t
//   D_t = diag(S_yy[m]) - diag(S_yx[m]) / diag(S_xx[m]) * diag(S_xy[m])   (per mixture table)
//   y[:, d] = solve( W^T D^-1 W , W^T D^-1 E )  per static dim d, pentadiagonal SPD system,
//   windows = kwiiyatta DELTA_WINDOWS (kwiiyatta/converter/delta.py:8-12).
#include "common.cuh"

namespace kw {

struct PreparedView {
    double* px_prec_chol;  // (K, Dh, Dh)
    double* px_aux;        // (K, Dh + 2)
    double* at;            // (K, Dh, Dh)  A^T: at[c][r] = A[r][c]
    double* offset;        // (K, Dh)
    double* var;           // (K, Dh)
    double* scratch;       // (K, Dh, Dh)  S_xx, then S_yx P
    double* src_means;     // (K, Dh)
    double* tgt_means;     // (K, Dh)
    size_t len;
};

static PreparedView view_prepared(double* base, int K, int Dh) {
    PreparedView v;
    size_t o = 0;
    const size_t kdd = (size_t)K * Dh * Dh, kd = (size_t)K * Dh;
    v.px_prec_chol = base + o; o += kdd;
    v.px_aux = base + o;       o += (size_t)K * (Dh + 2);
    v.at = base + o;           o += kdd;
    v.offset = base + o;       o += kd;
    v.var = base + o;          o += kd;
    v.scratch = base + o;      o += kdd;
    v.src_means = base + o;    o += kd;
    v.tgt_means = base + o;    o += kd;
    v.len = o;
    return v;
}

// Slice the joint model (MLPGBase.__init__), with the diff rewrite when asked.
__global__ void convert_slice_kernel(int K, int Dh, int diff, const double* __restrict__ means,
                                     const double* __restrict__ cov, PreparedView v) {
    const int k = blockIdx.x;
    const int D = 2 * Dh;
    const double* C = cov + (size_t)k * D * D;
    const double* mu = means + (size_t)k * D;
    for (int e = threadIdx.x; e < Dh * Dh; e += blockDim.x) {
        const int r = e / Dh, c = e - r * Dh;
        v.scratch[(size_t)k * Dh * Dh + e] = C[(size_t)r * D + c];  // S_xx
    }
    for (int d = threadIdx.x; d < Dh; d += blockDim.x) {
        const double cxx = C[(size_t)d * D + d];
        double cxy = C[(size_t)d * D + Dh + d];
        double cyx = C[(size_t)(Dh + d) * D + d];
        double cyy = C[(size_t)(Dh + d) * D + Dh + d];
        double ty = mu[Dh + d];
        if (diff) {
            ty = ty - mu[d];
            cyy = cxx + cyy - cxy - cyx;
            cxy = cxy - cxx;
            cyx = cxy;
        }
        v.src_means[(size_t)k * Dh + d] = mu[d];
        v.tgt_means[(size_t)k * Dh + d] = ty;
        v.var[(size_t)k * Dh + d] = cyy - cyx / cxx * cxy;
    }
}

// A = S_yx P P^T (P = px_prec_chol, upper), offset = mu_y - A mu_x.  One CTA per mixture.
__global__ void convert_regress_kernel(int K, int Dh, int diff, const double* __restrict__ cov,
                                       PreparedView v) {
    extern __shared__ double t1[];  // Dh * Dh : S_yx P
    const int k = blockIdx.x;
    const int D = 2 * Dh;
    const double* C = cov + (size_t)k * D * D;
    const double* P = v.px_prec_chol + (size_t)k * Dh * Dh;
    auto syx = [&](int r, int c) -> double {
        if (!diff) return C[(size_t)(Dh + r) * D + c];
        return C[(size_t)c * D + Dh + r] - C[(size_t)c * D + r];  // (S_xy - S_xx)^T
    };
    for (int e = threadIdx.x; e < Dh * Dh; e += blockDim.x) {
        const int r = e / Dh, c = e - r * Dh;
        double s = 0.0;
        for (int q = 0; q <= c; ++q) s = fma(syx(r, q), P[(size_t)q * Dh + c], s);
        t1[e] = s;
    }
    __syncthreads();
    double* at = v.at + (size_t)k * Dh * Dh;
    for (int e = threadIdx.x; e < Dh * Dh; e += blockDim.x) {
        const int r = e / Dh, c = e - r * Dh;
        double s = 0.0;
        for (int q = c; q < Dh; ++q) s = fma(t1[r * Dh + q], P[(size_t)c * Dh + q], s);
        at[(size_t)c * Dh + r] = s;
    }
    __syncthreads();
    const double* mx = v.src_means + (size_t)k * Dh;
    for (int r = threadIdx.x; r < Dh; r += blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < Dh; ++c) s = fma(at[(size_t)c * Dh + r], mx[c], s);
        v.offset[(size_t)k * Dh + r] = v.tgt_means[(size_t)k * Dh + r] - s;
    }
}

// Conditional means E[t] = offset[m_t] + A[m_t] x[t], grouped by mixture: a counting sort of the
// frames by hard label (histogram, scan, scatter), then one CTA per (mixture, 64 frames of it)
// with A^T of that mixture and the frames staged in shared memory.
constexpr int CM_FT = 64;

__global__ void convert_hist_kernel(long long N, int K, const int32_t* __restrict__ mix,
                                    int* __restrict__ counts) {
    extern __shared__ int h[];
    for (int i = threadIdx.x; i < K; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N;
         n += (long long)gridDim.x * blockDim.x)
        atomicAdd(&h[mix[n]], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x)
        if (h[i] != 0) atomicAdd(&counts[i], h[i]);
}

// offsets[m] = first slot of mixture m in the sorted order; blockstart[m] = first CTA of it.
__global__ void convert_scan_kernel(int K, const int* __restrict__ counts, int* __restrict__ offsets,
                                    int* __restrict__ blockstart, int* __restrict__ cursor) {
    if (threadIdx.x == 0) {
        int o = 0, b = 0;
        for (int m = 0; m < K; ++m) {
            offsets[m] = o;
            blockstart[m] = b;
            cursor[m] = 0;
            o += counts[m];
            b += (counts[m] + CM_FT - 1) / CM_FT;
        }
        offsets[K] = o;
        blockstart[K] = b;
    }
}

__global__ void convert_scatter_kernel(long long N, const int32_t* __restrict__ mix,
                                       const int* __restrict__ offsets, int* __restrict__ cursor,
                                       int* __restrict__ order) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int m = mix[n];
    order[offsets[m] + atomicAdd(&cursor[m], 1)] = (int)n;
}

__global__ void __launch_bounds__(256)
convert_condmean_kernel(int K, int Dh, const double* __restrict__ src,
                        const int* __restrict__ offsets, const int* __restrict__ blockstart,
                        const int* __restrict__ order, PreparedView v, double* __restrict__ E) {
    extern __shared__ double sm[];
    const int XS = Dh + 1;
    double* as = sm;                    // Dh * Dh : A^T[c][r]
    double* xs = as + Dh * Dh;          // CM_FT * XS
    __shared__ int idx[CM_FT];
    if ((int)blockIdx.x >= blockstart[K]) return;
    int lo = 0, hi = K;                 // mixture with blockstart[m] <= blockIdx.x < blockstart[m+1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (blockstart[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    const int m = lo;
    const int first = offsets[m] + ((int)blockIdx.x - blockstart[m]) * CM_FT;
    const int count = min(CM_FT, offsets[m + 1] - first);
    const int tid = threadIdx.x;
    if (tid < CM_FT) idx[tid] = tid < count ? order[first + tid] : -1;
    const double* at = v.at + (size_t)m * Dh * Dh;
    for (int e = tid; e < Dh * Dh; e += 256) as[e] = at[e];
    __syncthreads();
    for (int e = tid; e < CM_FT * Dh; e += 256) {
        const int f = e / Dh, c = e - f * Dh;
        xs[f * XS + c] = idx[f] >= 0 ? src[(size_t)idx[f] * Dh + c] : 0.0;
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    double acc[4][5];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int q = 0; q < 5; ++q) acc[a][q] = 0.0;
    for (int c = 0; c < Dh; ++c) {
        double a[5], x[4];
#pragma unroll
        for (int q = 0; q < 5; ++q) a[q] = (tx + 16 * q < Dh) ? as[c * Dh + tx + 16 * q] : 0.0;
#pragma unroll
        for (int f = 0; f < 4; ++f) x[f] = xs[(ty * 4 + f) * XS + c];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int q = 0; q < 5; ++q) acc[f][q] = fma(a[q], x[f], acc[f][q]);
    }
    const double* off = v.offset + (size_t)m * Dh;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        const int n = idx[ty * 4 + f];
        if (n < 0) continue;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int r = tx + 16 * q;
            if (r < Dh) E[(size_t)n * Dh + r] = off[r] + acc[f][q];
        }
    }
}

// Soft-posterior mapping (nnmnkwii MLPGBase.transform, the mlpg=False branch of
// kwiiyatta/converter/gmm.py:30-31):  y_t = sum_m p(m | x_t) (offset[m] + A[m] x_t).
// CTA = 32 frames; thread r owns output dimension r for all 32 frames; mixtures whose posterior is
// below 1e-14 for every frame of the tile are skipped.
constexpr int SOFT_FT = 32;
__global__ void __launch_bounds__(96)
convert_soft_kernel(long long N, long long Npad, int K, int Dh, const double* __restrict__ src,
                    const double* __restrict__ respT, PreparedView v, double* __restrict__ out) {
    extern __shared__ double xs[];          // SOFT_FT * Dh frames, then SOFT_FT responsibilities
    double* rs = xs + SOFT_FT * Dh;
    __shared__ int any_active;
    const long long n0 = (long long)blockIdx.x * SOFT_FT;
    for (int e = threadIdx.x; e < SOFT_FT * Dh; e += blockDim.x) {
        const long long n = n0 + e / Dh;
        xs[e] = (n < N) ? src[n * Dh + (e % Dh)] : 0.0;
    }
    const int r = threadIdx.x;
    double acc[SOFT_FT];
#pragma unroll
    for (int f = 0; f < SOFT_FT; ++f) acc[f] = 0.0;
    for (int m = 0; m < K; ++m) {
        __syncthreads();
        if (threadIdx.x == 0) any_active = 0;
        __syncthreads();
        if (threadIdx.x < SOFT_FT) {
            const long long n = n0 + threadIdx.x;
            const double p = (n < N) ? respT[(size_t)m * Npad + n] : 0.0;
            rs[threadIdx.x] = p;
            if (p > 1e-14) any_active = 1;
        }
        __syncthreads();
        if (!any_active || r >= Dh) continue;
        const double* at = v.at + (size_t)m * Dh * Dh;
        const double off = v.offset[(size_t)m * Dh + r];
        double e[SOFT_FT];
#pragma unroll
        for (int f = 0; f < SOFT_FT; ++f) e[f] = off;
        for (int c = 0; c < Dh; ++c) {
            const double a = at[(size_t)c * Dh + r];
#pragma unroll
            for (int f = 0; f < SOFT_FT; ++f) e[f] = fma(a, xs[f * Dh + c], e[f]);
        }
#pragma unroll
        for (int f = 0; f < SOFT_FT; ++f) acc[f] = fma(rs[f], e[f], acc[f]);
    }
    if (r < Dh) {
#pragma unroll
        for (int f = 0; f < SOFT_FT; ++f)
            if (n0 + f < N) out[(n0 + f) * Dh + r] = acc[f];
    }
}

// MLPG in two kernels.
//  1. convert_band_kernel, one thread per (frame, static dim): the pentadiagonal normal
//     equations W^T P W y = W^T P mu of that dimension, row by row: a_i = P[i][i],
//     b_i = P[i][i+1], c_i = P[i][i+2] and the right-hand side.  Fully parallel; this is where
//     the per-frame gathers (mixture -> variance) and the divisions happen.
//  2. convert_mlpg_kernel, one thread per (utterance, static dim): banded Cholesky
//     (bandwidth 2) and the two substitutions, in place over the four arrays, reading four
//     steps ahead of the recurrence.  The chain per frame is two fused multiply-adds and one
//     reciprocal square root: with r_i = 1 / L[i][i], e_i = L[i][i-1], f_i = L[i][i-2]
//        f_i = c_{i-2} r_{i-2};  e_i = (b_{i-1} - f_i e_{i-1}) r_{i-1};
//        r_i = rsqrt(a_i - f_i^2 - e_i^2);  z_i = (rhs_i - e_i z_{i-1} - f_i z_{i-2}) r_i
//     and back:  y_i = (z_i - e_{i+1} y_{i+1} - f_{i+2} y_{i+2}) r_i  with f_{i+2} = c_i r_i.
struct Windows {
    double c[3][3];  // c[w][o+1]: coefficient of window w at offset o in {-1, 0, 1}
};

// flags[n]: bit 0 = first frame of its utterance, bit 1 = last frame
__global__ void convert_flags_kernel(int n_utts, const int64_t* __restrict__ off,
                                     unsigned char* __restrict__ flags) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_utts) return;
    const long long t0 = off[u], t1 = off[u + 1];
    if (t1 <= t0) return;
    if (t1 - t0 == 1) {
        flags[t0] = 3;
    } else {
        flags[t0] = 1;
        flags[t1 - 1] = 2;
    }
}

__global__ void __launch_bounds__(256)
convert_band_kernel(long long total, int sd, int Dh, const double* __restrict__ E,
                    const int32_t* __restrict__ mix, const double* __restrict__ var,
                    const unsigned char* __restrict__ flags, Windows win,
                    double* __restrict__ band) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total * sd) return;
    const long long n = gid / sd;
    const int d = (int)(gid - n * sd);
    const unsigned fl = flags[n];
    // p[w], bs[w] at frames n-1 (m), n (c), n+1 (x)
    double pm[3] = {0, 0, 0}, bm[3] = {0, 0, 0}, pc[3], bc[3], px[3] = {0, 0, 0}, bx[3] = {0, 0, 0};
    auto load = [&](long long t, double* p, double* b) {
        const int m = mix[t];
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            const double prec = 1.0 / var[(size_t)m * Dh + w * sd + d];
            p[w] = prec;
            b[w] = prec * E[t * Dh + w * sd + d];
        }
    };
    load(n, pc, bc);
    if (!(fl & 1u)) load(n - 1, pm, bm);
    if (!(fl & 2u)) load(n + 1, px, bx);
    double a_i = 0.0, b_i = 0.0, c_i = 0.0, rhs = 0.0;
#pragma unroll
    for (int w = 0; w < 3; ++w) {
        const double cm = win.c[w][0], c0 = win.c[w][1], cp = win.c[w][2];
        a_i += cp * cp * pm[w] + c0 * c0 * pc[w] + cm * cm * px[w];
        rhs += cp * bm[w] + c0 * bc[w] + cm * bx[w];
        b_i += c0 * cp * pc[w] + cm * c0 * px[w];
        c_i += cm * cp * px[w];
    }
    const size_t plane = (size_t)total * sd;
    band[gid] = a_i;
    band[plane + gid] = b_i;
    band[2 * plane + gid] = c_i;
    band[3 * plane + gid] = rhs;
}

__global__ void __launch_bounds__(96)
convert_mlpg_kernel(int n_utts, const int64_t* __restrict__ off, int sd, double* __restrict__ ws,
                    long long total, double* __restrict__ out) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int u = gid / sd, d = gid - u * sd;
    if (u >= n_utts) return;
    const long long t0 = off[u];
    const int T = (int)(off[u + 1] - t0);
    if (T <= 0) return;
    const size_t plane = (size_t)total * sd;
    double* pa = ws + (size_t)t0 * sd + d;     // a -> r
    double* pb = pa + plane;                   // b -> e
    double* pcc = pb + plane;                  // c (kept)
    double* pr = pcc + plane;                  // rhs -> z
    constexpr int PF = 4;
    double c_m2 = 0, c_m1 = 0, b_m1 = 0, r_m1 = 1, r_m2 = 1, e_m1 = 0, z_m1 = 0, z_m2 = 0;
    double na[PF], nb[PF], nc[PF], nr[PF];
    auto fetch = [&](int i0) {
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            const int i = min(i0 + q, T - 1);
            const size_t o = (size_t)i * sd;
            na[q] = pa[o]; nb[q] = pb[o]; nc[q] = pcc[o]; nr[q] = pr[o];
        }
    };
    fetch(0);
    for (int i0 = 0; i0 < T; i0 += PF) {
        double ca[PF], cb[PF], cc[PF], cr[PF];
#pragma unroll
        for (int q = 0; q < PF; ++q) { ca[q] = na[q]; cb[q] = nb[q]; cc[q] = nc[q]; cr[q] = nr[q]; }
        if (i0 + PF < T) fetch(i0 + PF);
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            if (i0 + q < T) {
                const double f = c_m2 * r_m2;
                const double e = (b_m1 - f * e_m1) * r_m1;
                const double r = rsqrt(ca[q] - f * f - e * e);
                const double z = (cr[q] - e * z_m1 - f * z_m2) * r;
                const size_t o = (size_t)(i0 + q) * sd;
                pa[o] = r; pb[o] = e; pr[o] = z;      // c stays: f_{i+2} = c_i r_i on the way back
                c_m2 = c_m1; c_m1 = cc[q]; b_m1 = cb[q];
                r_m2 = r_m1; r_m1 = r; e_m1 = e;
                z_m2 = z_m1; z_m1 = z;
            }
        }
    }
    // back substitution L^T y = z
    double* po = out + (size_t)t0 * sd + d;
    double y1 = 0, y2 = 0, e_n1 = 0;
    auto fetch_back = [&](int i0) {      // steps i0, i0-1, ...
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            const int i = max(i0 - q, 0);
            const size_t o = (size_t)i * sd;
            na[q] = pa[o]; nb[q] = pb[o]; nc[q] = pcc[o]; nr[q] = pr[o];
        }
    };
    fetch_back(T - 1);
    for (int i0 = T - 1; i0 >= 0; i0 -= PF) {
        double ca[PF], cb[PF], cc[PF], cr[PF];
#pragma unroll
        for (int q = 0; q < PF; ++q) { ca[q] = na[q]; cb[q] = nb[q]; cc[q] = nc[q]; cr[q] = nr[q]; }
        if (i0 - PF >= 0) fetch_back(i0 - PF);
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            if (i0 - q >= 0) {
                const double y = (cr[q] - e_n1 * y1 - (cc[q] * ca[q]) * y2) * ca[q];
                po[(size_t)(i0 - q) * sd] = y;
                y2 = y1; y1 = y;
                e_n1 = cb[q];
            }
        }
    }
}

struct ConvertWorkspace {
    int32_t* mix;
    int* counts;      // K
    int* offsets;     // K + 1
    int* blockstart;  // K + 1
    int* cursor;      // K
    int* order;       // total
    double* E;
    double* band;
    unsigned char* flags;
    size_t bytes;
};

static ConvertWorkspace carve_convert(long long total, int K, int Dh, int sd, void* base) {
    Carver c(base);
    ConvertWorkspace w;
    w.mix = c.take<int32_t>((size_t)total);
    w.counts = c.take<int>((size_t)K);
    w.offsets = c.take<int>((size_t)K + 1);
    w.blockstart = c.take<int>((size_t)K + 1);
    w.cursor = c.take<int>((size_t)K);
    w.order = c.take<int>((size_t)total);
    w.E = c.take<double>((size_t)total * Dh);
    w.band = c.take<double>(4 * (size_t)total * sd);
    w.flags = c.take<unsigned char>((size_t)total);
    w.bytes = align_up(c.used, 256);
    return w;
}

}  // namespace kw

using namespace kw;

extern "C" size_t kw_convert_prepared_len(int K, int Dh) {
    return view_prepared(nullptr, K, Dh).len;
}

extern "C" int kw_convert_prepare(int K, int Dh, int diff, const double* weights_dev,
                                  const double* means_dev, const double* covariances_dev,
                                  double* prepared_dev, int32_t* info_dev, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    KW_REQUIRE(K > 0 && Dh > 0, "kw_convert_prepare: K, dim_half must be positive");
    PreparedView v = view_prepared(prepared_dev, K, Dh);
    convert_slice_kernel<<<K, 128, 0, st>>>(K, Dh, diff, means_dev, covariances_dev, v);
    KW_CUDA_CHECK(cudaGetLastError());
    int rc = finalize_launch(K, Dh, 0.0, 0, 0, nullptr, nullptr,
                             const_cast<double*>(weights_dev), v.src_means, v.scratch,
                             v.px_prec_chol, v.px_aux, info_dev, st);
    if (rc != KW_OK) return rc;
    const size_t smem = sizeof(double) * (size_t)Dh * Dh;
    KW_CUDA_CHECK(cudaFuncSetAttribute(convert_regress_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    convert_regress_kernel<<<K, 256, smem, st>>>(K, Dh, diff, covariances_dev, v);
    KW_CUDA_CHECK(cudaGetLastError());
    return KW_OK;
}

extern "C" size_t kw_convert_soft_workspace_bytes(int64_t total_frames, int K, int Dh) {
    Carver c(nullptr);
    c.take<double>((size_t)K * (size_t)resp_pad(total_frames));
    c.take<double>((size_t)((total_frames + 63) / 64) + 1);
    return align_up(c.used, 256);
}

extern "C" int kw_convert_soft_batch(int64_t total, const double* src_dev, int K, int Dh,
                                     const double* prepared_dev, double* out_dev,
                                     void* workspace_dev, size_t workspace_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (total == 0) return KW_OK;
    KW_REQUIRE(total > 0 && K > 0 && Dh > 0, "kw_convert_soft_batch: bad sizes");
    KW_REQUIRE(Dh <= 96, "dim_half %d > 96 unsupported", Dh);
    if (kw_convert_soft_workspace_bytes(total, K, Dh) > workspace_bytes) {
        set_error("soft conversion workspace too small");
        return KW_ERR_WORKSPACE;
    }
    Carver c(workspace_dev);
    double* resp = c.take<double>((size_t)K * (size_t)resp_pad(total));
    double* lse = c.take<double>((size_t)((total + 63) / 64) + 1);
    PreparedView v = view_prepared(const_cast<double*>(prepared_dev), K, Dh);
    int rc = estep_fp64(total, src_dev, K, Dh, v.px_prec_chol, v.px_aux, resp, lse, 0, nullptr, st);
    if (rc != KW_OK) return rc;
    const size_t smem = sizeof(double) * ((size_t)SOFT_FT * Dh + SOFT_FT);
    convert_soft_kernel<<<(unsigned)((total + SOFT_FT - 1) / SOFT_FT), 96, smem, st>>>(
        total, resp_pad(total), K, Dh, src_dev, resp, v, out_dev);
    KW_CUDA_CHECK(cudaGetLastError());
    return KW_OK;
}

extern "C" size_t kw_convert_workspace_bytes(int64_t total_frames, int K, int Dh, int precision) {
    size_t b = carve_convert(total_frames, K, Dh, Dh / 3, nullptr).bytes;
    if (precision == 1) b += tc_workspace_bytes(total_frames, K, Dh);
    return b;
}

extern "C" int kw_convert_batch(int n_utts, const int64_t* off_dev, int64_t total, int max_frames,
                                const double* src_dev, int K, int Dh, const double* prepared_dev,
                                double* out_dev, int32_t* mix_dev, int precision,
                                void* workspace_dev, size_t workspace_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    (void)max_frames;
    if (n_utts == 0 || total == 0) return KW_OK;
    KW_REQUIRE(n_utts > 0 && total > 0 && K > 0, "kw_convert_batch: bad sizes");
    KW_REQUIRE(Dh % 3 == 0, "dim_half %d is not static+delta+delta2 (multiple of 3)", Dh);
    KW_REQUIRE(precision == 0 || precision == 1, "convert precision must be 0 or 1");
    const int sd = Dh / 3;
    ConvertWorkspace w = carve_convert(total, K, Dh, sd, workspace_dev);
    if (w.bytes > workspace_bytes) {
        set_error("convert workspace too small: need %zu bytes, got %zu", w.bytes,
                  workspace_bytes);
        return KW_ERR_WORKSPACE;
    }
    PreparedView v = view_prepared(const_cast<double*>(prepared_dev), K, Dh);
    int rc;
    if (precision == 1) {
        rc = pack_frames_tc(total, src_dev, K, Dh, static_cast<char*>(workspace_dev) + w.bytes,
                            workspace_bytes - w.bytes, st, /*mstep_parts=*/false);
        if (rc != KW_OK) return rc;
    }
    if (precision == 1)
        rc = estep_tc(total, src_dev, K, Dh, v.src_means, v.px_prec_chol, v.px_aux, nullptr,
                      nullptr, 1, w.mix, static_cast<char*>(workspace_dev) + w.bytes,
                      workspace_bytes - w.bytes, st);
    else
        rc = estep_fp64(total, src_dev, K, Dh, v.px_prec_chol, v.px_aux, nullptr, nullptr, 1,
                        w.mix, st);
    if (rc != KW_OK) return rc;
    {
        KW_REQUIRE(Dh <= 80, "dim_half %d > 80 unsupported", Dh);
        KW_REQUIRE(total < 2147483647LL, "too many frames for one conversion batch");
        KW_CUDA_CHECK(cudaMemsetAsync(w.counts, 0, sizeof(int) * K, st));
        convert_hist_kernel<<<296, 256, sizeof(int) * K, st>>>(total, K, w.mix, w.counts);
        KW_CUDA_CHECK(cudaGetLastError());
        convert_scan_kernel<<<1, 32, 0, st>>>(K, w.counts, w.offsets, w.blockstart, w.cursor);
        KW_CUDA_CHECK(cudaGetLastError());
        convert_scatter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
            total, w.mix, w.offsets, w.cursor, w.order);
        KW_CUDA_CHECK(cudaGetLastError());
        const size_t smem = sizeof(double) * ((size_t)Dh * Dh + (size_t)CM_FT * (Dh + 1));
        KW_CUDA_CHECK(cudaFuncSetAttribute(convert_condmean_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long max_blocks = (total + CM_FT - 1) / CM_FT + K;
        convert_condmean_kernel<<<(unsigned)max_blocks, 256, smem, st>>>(
            K, Dh, src_dev, w.offsets, w.blockstart, w.order, v, w.E);
        KW_CUDA_CHECK(cudaGetLastError());
    }
    {
        Windows win = {{{0.0, 1.0, 0.0}, {-0.5, 0.0, 0.5}, {1.0, -2.0, 1.0}}};
        KW_CUDA_CHECK(cudaMemsetAsync(w.flags, 0, (size_t)total, st));
        convert_flags_kernel<<<(n_utts + 127) / 128, 128, 0, st>>>(n_utts, off_dev, w.flags);
        KW_CUDA_CHECK(cudaGetLastError());
        const long long cells = total * sd;
        convert_band_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(
            total, sd, Dh, w.E, w.mix, v.var, w.flags, win, w.band);
        KW_CUDA_CHECK(cudaGetLastError());
        const long long threads = (long long)n_utts * sd;
        convert_mlpg_kernel<<<(unsigned)((threads + 95) / 96), 96, 0, st>>>(
            n_utts, off_dev, sd, w.band, total, out_dev);
        KW_CUDA_CHECK(cudaGetLastError());
    }
    if (mix_dev != nullptr)
        KW_CUDA_CHECK(cudaMemcpyAsync(mix_dev, w.mix, sizeof(int32_t) * total,
                                      cudaMemcpyDeviceToDevice, st));
    return KW_OK;
}

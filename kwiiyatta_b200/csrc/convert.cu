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

// E[t][r] = offset[m][r] + sum_c A^T[m][c][r] x[t][c].  CTA = 8 frames; runs of equal mixture
// share the loads of A^T.
constexpr int C_FT = 8;
__global__ void __launch_bounds__(128)
convert_condmean_kernel(long long N, int Dh, const double* __restrict__ src,
                        const int32_t* __restrict__ mix, PreparedView v,
                        double* __restrict__ E) {
    extern __shared__ double xs[];  // C_FT * Dh
    __shared__ int ms[C_FT];
    const long long n0 = (long long)blockIdx.x * C_FT;
    for (int e = threadIdx.x; e < C_FT * Dh; e += blockDim.x) {
        const long long n = n0 + e / Dh;
        xs[e] = (n < N) ? src[n * Dh + (e % Dh)] : 0.0;
    }
    if (threadIdx.x < C_FT) ms[threadIdx.x] = (n0 + threadIdx.x < N) ? mix[n0 + threadIdx.x] : -1;
    __syncthreads();
    const int r = threadIdx.x;
    if (r >= Dh) return;
    int f0 = 0;
    while (f0 < C_FT && ms[f0] >= 0) {
        const int m = ms[f0];
        int f1 = f0 + 1;
        while (f1 < C_FT && ms[f1] == m) ++f1;
        const double* at = v.at + (size_t)m * Dh * Dh;
        double acc[C_FT];
        const double off = v.offset[(size_t)m * Dh + r];
#pragma unroll
        for (int q = 0; q < C_FT; ++q) acc[q] = 0.0;
        for (int c = 0; c < Dh; ++c) {
            const double a = at[(size_t)c * Dh + r];
#pragma unroll
            for (int q = 0; q < C_FT; ++q)
                if (f0 + q < f1) acc[q] = fma(a, xs[(f0 + q) * Dh + c], acc[q]);
        }
#pragma unroll
        for (int q = 0; q < C_FT; ++q)
            if (f0 + q < f1) E[(n0 + f0 + q) * Dh + r] = off + acc[q];
        f0 = f1;
    }
}

// MLPG: one thread per (utterance, static dim).  Banded Cholesky (bandwidth 2) with the
// factor and the forward solution kept in a global workspace laid out [frame][static dim].
struct Windows {
    double c[3][3];  // c[w][o+1]: coefficient of window w at offset o in {-1, 0, 1}
};

__global__ void __launch_bounds__(96)
convert_mlpg_kernel(int n_utts, const int64_t* __restrict__ off, int sd, int Dh,
                    const double* __restrict__ E, const int32_t* __restrict__ mix,
                    const double* __restrict__ var, Windows win, double* __restrict__ ws,
                    long long total, double* __restrict__ out) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int u = gid / sd, d = gid - u * sd;
    if (u >= n_utts) return;
    const long long t0 = off[u];
    const int T = (int)(off[u + 1] - t0);
    if (T <= 0) return;
    double* l0 = ws;
    double* l1 = ws + (size_t)total * sd;
    double* l2 = ws + 2 * (size_t)total * sd;
    double* zz = ws + 3 * (size_t)total * sd;
    // p[w], bs[w] at frames i-1 (m), i (c), i+1 (n)
    double pm[3] = {0, 0, 0}, bm[3] = {0, 0, 0}, pc[3], bc[3], pn[3] = {0, 0, 0}, bn[3] = {0, 0, 0};
    auto load = [&](int t, double* p, double* b) {
        const long long n = t0 + t;
        const int m = mix[n];
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            const double prec = 1.0 / var[(size_t)m * Dh + w * sd + d];
            p[w] = prec;
            b[w] = prec * E[n * Dh + w * sd + d];
        }
    };
    load(0, pc, bc);
    // banded Cholesky P = L L^T with d_i = L[i][i], e_i = L[i][i-1], f_i = L[i][i-2]:
    //   f_i = c_{i-2} / d_{i-2};  e_i = (b_{i-1} - f_i e_{i-1}) / d_{i-1};
    //   d_i = sqrt(a_i - f_i^2 - e_i^2);  z_i = (rhs_i - e_i z_{i-1} - f_i z_{i-2}) / d_i
    // where a_i = P[i][i], b_i = P[i][i+1], c_i = P[i][i+2].
    double c_m2 = 0, c_m1 = 0, b_m1 = 0, d_m1 = 1, d_m2 = 1, e_m1 = 0, z_m1 = 0, z_m2 = 0;
    for (int i = 0; i < T; ++i) {
        if (i + 1 < T) {
            load(i + 1, pn, bn);
        } else {
#pragma unroll
            for (int w = 0; w < 3; ++w) { pn[w] = 0.0; bn[w] = 0.0; }
        }
        double a_i = 0.0, b_i = 0.0, c_i = 0.0, rhs = 0.0;
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            const double cm = win.c[w][0], c0 = win.c[w][1], cp = win.c[w][2];
            a_i += cp * cp * pm[w] + c0 * c0 * pc[w] + cm * cm * pn[w];
            rhs += cp * bm[w] + c0 * bc[w] + cm * bn[w];
            b_i += c0 * cp * pc[w] + cm * c0 * pn[w];
            c_i += cm * cp * pn[w];
        }
        const double f = (i >= 2) ? c_m2 / d_m2 : 0.0;
        const double e = (i >= 1) ? (b_m1 - f * e_m1) / d_m1 : 0.0;
        const double dg = sqrt(a_i - f * f - e * e);
        const double z = (rhs - e * z_m1 - f * z_m2) / dg;
        const size_t idx = (size_t)(t0 + i) * sd + d;
        l0[idx] = dg; l1[idx] = e; l2[idx] = f; zz[idx] = z;
        c_m2 = c_m1; c_m1 = c_i; b_m1 = b_i;
        d_m2 = d_m1; d_m1 = dg; e_m1 = e;
        z_m2 = z_m1; z_m1 = z;
#pragma unroll
        for (int w = 0; w < 3; ++w) { pm[w] = pc[w]; bm[w] = bc[w]; pc[w] = pn[w]; bc[w] = bn[w]; }
    }
    // back substitution L^T y = z:  y_i = (z_i - e_{i+1} y_{i+1} - f_{i+2} y_{i+2}) / d_i
    double y1 = 0, y2 = 0, e_n1 = 0, f_n1 = 0, f_n2 = 0;
    for (int i = T - 1; i >= 0; --i) {
        const size_t idx = (size_t)(t0 + i) * sd + d;
        const double y = (zz[idx] - e_n1 * y1 - f_n2 * y2) / l0[idx];
        out[idx] = y;
        y2 = y1; y1 = y;
        f_n2 = f_n1; f_n1 = l2[idx]; e_n1 = l1[idx];
    }
}

struct ConvertWorkspace {
    int32_t* mix;
    double* E;
    double* band;
    size_t bytes;
};

static ConvertWorkspace carve_convert(long long total, int Dh, int sd, void* base) {
    Carver c(base);
    ConvertWorkspace w;
    w.mix = c.take<int32_t>((size_t)total);
    w.E = c.take<double>((size_t)total * Dh);
    w.band = c.take<double>(4 * (size_t)total * sd);
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

extern "C" size_t kw_convert_workspace_bytes(int64_t total_frames, int K, int Dh, int precision) {
    size_t b = carve_convert(total_frames, Dh, Dh / 3, nullptr).bytes;
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
    ConvertWorkspace w = carve_convert(total, Dh, sd, workspace_dev);
    if (w.bytes > workspace_bytes) {
        set_error("convert workspace too small: need %zu bytes, got %zu", w.bytes,
                  workspace_bytes);
        return KW_ERR_WORKSPACE;
    }
    PreparedView v = view_prepared(const_cast<double*>(prepared_dev), K, Dh);
    int rc;
    if (precision == 1) {
        rc = pack_frames_tc(total, src_dev, K, Dh, static_cast<char*>(workspace_dev) + w.bytes,
                            workspace_bytes - w.bytes, st);
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
        const size_t smem = sizeof(double) * C_FT * Dh;
        const long long grid = (total + C_FT - 1) / C_FT;
        const int bd = (Dh + 31) / 32 * 32;
        KW_REQUIRE(bd <= 128, "dim_half %d > 128 unsupported", Dh);
        convert_condmean_kernel<<<(unsigned)grid, bd, smem, st>>>(total, Dh, src_dev, w.mix, v,
                                                                 w.E);
        KW_CUDA_CHECK(cudaGetLastError());
    }
    {
        Windows win = {{{0.0, 1.0, 0.0}, {-0.5, 0.0, 0.5}, {1.0, -2.0, 1.0}}};
        const long long threads = (long long)n_utts * sd;
        convert_mlpg_kernel<<<(unsigned)((threads + 95) / 96), 96, 0, st>>>(
            n_utts, off_dev, sd, Dh, w.E, w.mix, v.var, win, w.band, total, out_dev);
        KW_CUDA_CHECK(cudaGetLastError());
    }
    if (mix_dev != nullptr)
        KW_CUDA_CHECK(cudaMemcpyAsync(mix_dev, w.mix, sizeof(int32_t) * total,
                                      cudaMemcpyDeviceToDevice, st));
    return KW_OK;
}

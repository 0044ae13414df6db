// Device-side training-array assembly: everything between the mel-cepstra and the (N, 2*3*order)
// matrix GaussianMixture.fit receives, on the GPU that ran the DTW, so that the joint frames never
// visit the host.  Reference code paths (all per utterance pair, in Python, on the host):
//   kwiiyatta/vocoder/align.py:20-58    make_feature      -> kw_dtw_features
//   kwiiyatta/vocoder/align.py:73-94    strict filter     \  kw_path_select
//   kwiiyatta/vocoder/align.py:139-145  pad trim          /
//   kwiiyatta/vocoder/abc/feature.py:170-194 gather, converter/mcep.py:33 drop c0,
//   converter/delta.py:30 delta features, converter/dataset.py:68-70 hstack + zero-frame test
//                                                         -> kw_joint_frames
// Byte / index work: every kernel is bound by HBM or by its own latency, none by arithmetic.
#include "common.cuh"

namespace kw {

// (T, width) mel-cepstra -> (T, width + 1) DTW features: column 0 the power flag, column 1 the
// voicing flag, then mcep[1:].  power_mode 0: flag = power_weight where c0 >= thr[utt];
// 1: raw c0; 2: zero.  One thread per output element.
__global__ void dtw_features_kernel(int n_utts, const int64_t* __restrict__ off, long long total,
                                    int width, const double* __restrict__ mcep,
                                    const unsigned char* __restrict__ voiced,
                                    const double* __restrict__ thr, int power_mode,
                                    double power_weight, double vuv_weight,
                                    double* __restrict__ out) {
    const int ow = width + 1;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total * ow) return;
    const long long n = e / ow;
    const int c = (int)(e - n * ow);
    double v;
    if (c >= 2) {
        v = mcep[n * width + (c - 1)];
    } else if (c == 1) {
        v = (voiced != nullptr && voiced[n]) ? vuv_weight : 0.0;
    } else if (power_mode == 1) {
        v = mcep[n * width];
    } else if (power_mode == 2) {
        v = 0.0;
    } else {
        int lo = 0, hi = n_utts;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (off[mid] <= n) lo = mid; else hi = mid;
        }
        v = (mcep[n * width] >= thr[lo]) ? power_weight : 0.0;
    }
    out[e] = v;
}

// Strict filter + pad trim of one pair's path, one warp per pair.
//   strict: interior points are kept when the binary flags agree -- x power vs y power when
//           check_power, x VOICING vs y POWER when check_vuv (the reference's :78 quirk);
//   trim:   [first index with i >= pad and j >= pad, first index with i >= tx - pad and
//           j >= ty - pad), both 0 when no such index exists (np.argmax of an all-False mask).
// The selected points are written to the FRONT of the pair's path region (in place is safe: the
// warp reads a chunk of 32 points before it writes any of them, and writes never pass reads).
__global__ void __launch_bounds__(128)
path_select_kernel(int n_pairs, const int64_t* __restrict__ region_off,
                   const int32_t* __restrict__ path_begin, const int32_t* __restrict__ path_len,
                   const int32_t* __restrict__ tx, const int32_t* __restrict__ ty,
                   const int64_t* __restrict__ xoff, const int64_t* __restrict__ yoff,
                   const double* __restrict__ xf, const double* __restrict__ yf, int fdim,
                   int strict, int check_power, int check_vuv, int trim, int pad_len,
                   int32_t* __restrict__ path, int32_t* __restrict__ sel_len) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n_pairs) return;
    const int2* src = reinterpret_cast<const int2*>(path) + region_off[p] + path_begin[p];
    int2* dst = reinterpret_cast<int2*>(path) + region_off[p];
    const int len = path_len[p];
    const double* xfp = xf + xoff[p] * fdim;
    const double* yfp = yf + yoff[p] * fdim;
    const int a_len = tx[p], b_len = ty[p];
    // pass 1: position of every point in the strict-filtered path, begin / end of the trim
    // (indices into the FILTERED path, as the reference trims after filtering)
    int kept = 0, begin = -1, end = -1;
    for (int c0 = 0; c0 < len; c0 += 32) {
        const int q = c0 + lane;
        int2 pt = make_int2(0, 0);
        bool keep = false;
        if (q < len) {
            pt = src[q];
            keep = true;
            if (strict && q > 0 && q < len - 1) {
                const bool ypow = yfp[(size_t)pt.y * fdim] > 0.0;
                if (check_power && ((xfp[(size_t)pt.x * fdim] > 0.0) != ypow)) keep = false;
                if (check_vuv && ((xfp[(size_t)pt.x * fdim + 1] > 0.0) != ypow)) keep = false;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int pos = kept + __popc(m & ((1u << lane) - 1u));
        if (trim) {
            const bool in = keep && pt.x >= pad_len && pt.y >= pad_len;
            const bool past = keep && pt.x >= a_len - pad_len && pt.y >= b_len - pad_len;
            const unsigned mi = __ballot_sync(0xffffffffu, in);
            const unsigned mp = __ballot_sync(0xffffffffu, past);
            if (begin < 0 && mi) begin = __shfl_sync(0xffffffffu, pos, __ffs(mi) - 1);
            if (end < 0 && mp) end = __shfl_sync(0xffffffffu, pos, __ffs(mp) - 1);
        }
        kept += __popc(m);
    }
    if (strict && len == 1) kept = 2;     // the reference chains path[0], (), path[-1]
    if (!trim) { begin = 0; end = kept; }
    if (begin < 0) begin = 0;
    if (end < 0) end = 0;
    // pass 2: write the points with begin <= position < end
    int seen = 0;
    for (int c0 = 0; c0 < len; c0 += 32) {
        const int q = c0 + lane;
        int2 pt = make_int2(0, 0);
        bool keep = false;
        if (q < len) {
            pt = src[q];
            keep = true;
            if (strict && q > 0 && q < len - 1) {
                const bool ypow = yfp[(size_t)pt.y * fdim] > 0.0;
                if (check_power && ((xfp[(size_t)pt.x * fdim] > 0.0) != ypow)) keep = false;
                if (check_vuv && ((xfp[(size_t)pt.x * fdim + 1] > 0.0) != ypow)) keep = false;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int pos = seen + __popc(m & ((1u << lane) - 1u));
        __syncwarp();
        if (keep && pos >= begin && pos < end) dst[pos - begin] = pt;
        seen += __popc(m);
    }
    if (strict && len == 1 && lane == 0 && begin == 0 && end == 2) dst[1] = src[0];
    if (lane == 0) sel_len[p] = max(0, end - begin);
}

// Joint frames: row r of the output belongs to pair p (binary search in out_off) and is its
// l-th selected path point (i, j):
//   [ x[i][1:], d x, dd x, y[j][1:], d y, dd y ]   with the deltas taken over the SELECTED sequence
// (the reference gathers first and differentiates the aligned features, converter/delta.py:30),
// zero padded at the pair's ends, in np.correlate's order of operations.  zero_flag[r] = 1 when
// the row's absolute sum is not above 1e-7 (remove_zeros_frames, converter/dataset.py:70).
// One warp per row; lane = static coefficient (order <= 32).
__global__ void __launch_bounds__(256)
joint_frames_kernel(int n_pairs, const int64_t* __restrict__ out_off,
                    const int64_t* __restrict__ region_off, const int32_t* __restrict__ path,
                    const int64_t* __restrict__ xoff, const int64_t* __restrict__ yoff,
                    const double* __restrict__ xm, const double* __restrict__ ym, int width,
                    int use_delta, double* __restrict__ out,
                    unsigned char* __restrict__ zero_flag) {
    const int lane = threadIdx.x & 31;
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total = out_off[n_pairs];
    if (r >= total) return;
    int lo = 0, hi = n_pairs;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (out_off[mid] <= r) lo = mid; else hi = mid;
    }
    const int p = lo;
    const long long l = r - out_off[p], L = out_off[p + 1] - out_off[p];
    const int2* pts = reinterpret_cast<const int2*>(path) + region_off[p];
    const int order = width - 1;
    const int blocks = use_delta ? 3 : 1;
    double* o = out + r * (size_t)(2 * blocks * order);
    double asum = 0.0;
    for (int side = 0; side < 2; ++side) {
        const double* base = side == 0 ? xm + xoff[p] * width : ym + yoff[p] * width;
        const int2 pc = pts[l];
        const int ic = side == 0 ? pc.x : pc.y;
        double xc = 0.0, xmn = 0.0, xpl = 0.0;
        if (lane < order) {
            xc = base[(size_t)ic * width + 1 + lane];
            if (use_delta) {
                if (l > 0) {
                    const int2 q = pts[l - 1];
                    xmn = base[(size_t)(side == 0 ? q.x : q.y) * width + 1 + lane];
                }
                if (l + 1 < L) {
                    const int2 q = pts[l + 1];
                    xpl = base[(size_t)(side == 0 ? q.x : q.y) * width + 1 + lane];
                }
            }
            double* os = o + (size_t)side * blocks * order;
            os[lane] = xc;
            asum += fabs(xc);
            if (use_delta) {
                const double d1 = __dadd_rn(__dadd_rn(__dmul_rn(-0.5, xmn), __dmul_rn(0.0, xc)),
                                            __dmul_rn(0.5, xpl));
                const double d2 = __dadd_rn(__dadd_rn(xmn, __dmul_rn(-2.0, xc)), xpl);
                os[order + lane] = d1;
                os[2 * order + lane] = d2;
                asum += fabs(d1) + fabs(d2);
            }
        }
    }
    asum = warp_sum(asum);
    if (lane == 0) zero_flag[r] = !(asum > 1e-7);
}

}  // namespace kw

using namespace kw;

extern "C" int kw_dtw_features(int n_utts, const int64_t* off_dev, int64_t total_frames, int width,
                               const double* mcep_dev, const uint8_t* voiced_dev,
                               const double* threshold_dev, int power_mode, double power_weight,
                               double vuv_weight, double* out_dev, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (total_frames == 0) return KW_OK;
    KW_REQUIRE(n_utts > 0 && total_frames > 0 && width >= 2, "kw_dtw_features: bad sizes");
    KW_REQUIRE(power_mode >= 0 && power_mode <= 2, "kw_dtw_features: power_mode must be 0..2");
    KW_REQUIRE(power_mode != 0 || threshold_dev != nullptr,
               "kw_dtw_features: power_mode 0 needs per-utterance thresholds");
    const long long n = (long long)total_frames * (width + 1);
    dtw_features_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        n_utts, off_dev, total_frames, width, mcep_dev, voiced_dev, threshold_dev, power_mode,
        power_weight, vuv_weight, out_dev);
    KW_CUDA_CHECK(cudaGetLastError());
    return KW_OK;
}

extern "C" int kw_path_select(int n_pairs, const int64_t* region_off_dev,
                              const int32_t* path_begin_dev, const int32_t* path_len_dev,
                              const int32_t* tx_dev, const int32_t* ty_dev,
                              const int64_t* xoff_dev, const int64_t* yoff_dev,
                              const double* xfeat_dev, const double* yfeat_dev, int feat_dim,
                              int strict, int check_power, int check_vuv, int trim, int pad_len,
                              int32_t* path_dev, int32_t* selected_len_dev, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_pairs == 0) return KW_OK;
    KW_REQUIRE(n_pairs > 0 && feat_dim >= 2 && pad_len >= 0, "kw_path_select: bad arguments");
    path_select_kernel<<<(n_pairs + 3) / 4, 128, 0, st>>>(
        n_pairs, region_off_dev, path_begin_dev, path_len_dev, tx_dev, ty_dev, xoff_dev, yoff_dev,
        xfeat_dev, yfeat_dev, feat_dim, strict, check_power, check_vuv, trim, pad_len, path_dev,
        selected_len_dev);
    KW_CUDA_CHECK(cudaGetLastError());
    return KW_OK;
}

extern "C" int kw_joint_frames(int n_pairs, const int64_t* out_off_dev, int64_t total_rows,
                               const int64_t* region_off_dev, const int32_t* path_dev,
                               const int64_t* xoff_dev, const int64_t* yoff_dev,
                               const double* xmcep_dev, const double* ymcep_dev, int width,
                               int use_delta, double* out_dev, uint8_t* zero_flag_dev,
                               void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_pairs == 0 || total_rows == 0) return KW_OK;
    KW_REQUIRE(n_pairs > 0 && total_rows > 0, "kw_joint_frames: bad sizes");
    if (width - 1 > 32 || width < 2) {
        set_error("kw_joint_frames: mel-cepstrum order %d not in 1..32", width - 1);
        return KW_ERR_UNSUPPORTED;
    }
    const long long threads = (long long)total_rows * 32;
    joint_frames_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        n_pairs, out_off_dev, region_off_dev, path_dev, xoff_dev, yoff_dev, xmcep_dev, ymcep_dev,
        width, use_delta, out_dev, zero_flag_dev);
    KW_CUDA_CHECK(cudaGetLastError());
    return KW_OK;
}

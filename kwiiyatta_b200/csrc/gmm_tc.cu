// Tensor-core E-step (precision = 1): the Mahalanobis contraction of the joint-GMM E-step on
// tcgen05 with split-fp16 operands, for sm_100a.
//
//   y_nk = (x_n - mu_k) L_k,   q_nk = || y_nk ||^2      (kwiiyatta/converter/gmm.py:25-26 ->
//   sklearn _estimate_log_gaussian_prob; restated in oracle/gmm_ref.py weighted_log_prob)
//
// Numerics.  X is centred on its column means c and scaled per column by a power of two sigma_d
// so |x'| < 1;  L'_k = diag(sigma) L_k is scaled per output column by a power of two 2^t so its
// largest entry is in [0.5, 1).  Each operand is split  v = hi + lo * 2^-11  with hi = fp16(v),
// lo = fp16((v - hi) * 2^11)  (22 significant bits).  Per (128-frame tile, component):
//   acc  = x_hi . l_lo + x_lo . l_hi          (18 MMAs,  fp32 accumulate in TMEM)
//   acc  = acc * 2^-11 + x_hi . l_hi          (scale-input-d on the first of 9 MMAs)
//   y_j  = acc_j * 2^t_j - (mu_k - c) L_k     (epilogue, fp32; q accumulated in fp64 per 16 columns)
// so the result carries ~2^-22 relative operand error plus the fp32 accumulation of the MMA.
//
// Data movement.  Operands are pre-packed in global memory in the UMMA canonical K-major,
// no-swizzle layout (8 x 8 fp16 core matrices, 128 B each), so a tile is one contiguous block and
// is fetched with a single cp.async.bulk (1-D TMA) onto an mbarrier -- no tensor maps.  The same
// packed X tile is the MN-major operand a tensor-core M-step would need.
//
// Kernel.  Persistent, one CTA per SM, 6 warps: warp 0 = bulk-copy producer, warp 1 = MMA issuer
// (one elected thread), warps 2-5 = epilogue (one TMEM lane quarter each).  Two accumulator stages
// in TMEM (2 x DP columns) overlap the epilogue of component k with the MMAs of k+1.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace kw {
namespace tc {

constexpr int TILE_M = 128;
constexpr int LO_SHIFT = 11;          // lo parts are stored scaled by 2^11
constexpr double LO_SCALE = 2048.0;

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("kwiiyatta_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x,
                   threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; accumulate = 0 overwrites D.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D = A * B + D * 2^-11  (scale-input-d)
__device__ __forceinline__ void umma_f16_scaled(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 11;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, canonical K-major layout without swizzle:
//   element (row r, k-col c) at  (r/8)*SBO + (c/8)*LBO + (r%8)*16 B + (c%8)*2 B.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;                // base offset 0, layout type 0 = no swizzle
}
// Instruction descriptor: kind::f16, A = B = fp16, D = fp32, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Packed layouts (in halves).  X tile: [part][row group 16][k group DP/8][8][8];
// B (per component): [part][col group DP/8][k group DP/8][8][8].
__host__ __device__ inline size_t tile_elems(int DP) { return (size_t)TILE_M * DP; }
__host__ __device__ inline size_t bmat_elems(int DP) { return (size_t)DP * DP; }

// ------------------------------------------------------------------------------------------
// Column statistics of X: partial sums / min / max per chunk, then centre and power-of-two scale.
// ------------------------------------------------------------------------------------------
__global__ void colstats_partial_kernel(long long N, int D, const double* __restrict__ X,
                                        double* __restrict__ partial, long long frames_per_chunk) {
    const long long n0 = (long long)blockIdx.x * frames_per_chunk;
    const long long n1 = min(N, n0 + frames_per_chunk);
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double s = 0.0, mn = CUDART_INF, mx = -CUDART_INF;
        for (long long n = n0; n < n1; ++n) {
            const double v = X[n * D + d];
            s += v;
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        }
        double* p = partial + (size_t)blockIdx.x * 3 * D;
        p[d] = s;
        p[D + d] = mn;
        p[2 * D + d] = mx;
    }
}
// xinfo: [centre (DP) | sigma = 2^e (DP)]
__global__ void colstats_final_kernel(long long N, int D, int DP, int chunks,
                                      const double* __restrict__ partial,
                                      double* __restrict__ xinfo) {
    for (int d = threadIdx.x; d < DP; d += blockDim.x) {
        double c = 0.0, sigma = 1.0;
        if (d < D) {
            double s = 0.0, mn = CUDART_INF, mx = -CUDART_INF;
            for (int q = 0; q < chunks; ++q) {
                const double* p = partial + (size_t)q * 3 * D;
                s += p[d];
                mn = fmin(mn, p[D + d]);
                mx = fmax(mx, p[2 * D + d]);
            }
            c = s / (double)N;
            const double amax = fmax(mx - c, c - mn);
            if (amax > 0.0 && isfinite(amax)) {
                int e;
                frexp(amax, &e);        // amax = m * 2^e, m in [0.5, 1)  ->  amax < 2^e
                sigma = ldexp(1.0, e);
            }
        }
        xinfo[d] = c;
        xinfo[DP + d] = sigma;
    }
}

__device__ __forceinline__ void split_store(double v, __half* hi_dst, __half* lo_dst) {
    const __half h = __double2half(v);
    const double r = (v - (double)__half2float(h)) * LO_SCALE;
    *hi_dst = h;
    *lo_dst = __double2half(r);
}

// One CTA per 128-frame tile.
__global__ void pack_x_kernel(long long N, int D, int DP, const double* __restrict__ X,
                              const double* __restrict__ xinfo, __half* __restrict__ xt) {
    const long long tile = blockIdx.x;
    __half* hi = xt + (size_t)tile * 2 * tile_elems(DP);
    __half* lo = hi + tile_elems(DP);
    const int kg_n = DP / 8;
    for (int e = threadIdx.x; e < TILE_M * DP; e += blockDim.x) {
        const int r = e / DP, c = e - r * DP;
        const long long n = tile * TILE_M + r;
        double v = 0.0;
        if (n < N && c < D) v = (X[n * D + c] - xinfo[c]) / xinfo[DP + c];
        const size_t o = ((size_t)(r >> 3) * kg_n + (c >> 3)) * 64 + (r & 7) * 8 + (c & 7);
        split_store(v, hi + o, lo + o);
    }
}

// One CTA per component: B[j][d] = sigma_d L[d][j] 2^-t_j (split), column scale 2^t_j and
// b'_j = sum_d (mu_d - c_d) L[d][j], plus [log|L|, log w] from aux.
__global__ void pack_l_kernel(int K, int D, int DP, const double* __restrict__ means,
                              const double* __restrict__ prec_chol, const double* __restrict__ aux,
                              const double* __restrict__ xinfo, __half* __restrict__ bt,
                              float* __restrict__ sc, double* __restrict__ cst) {
    extern __shared__ double colinv[];  // DP: 2^-t_j
    const int k = blockIdx.x;
    const double* L = prec_chol + (size_t)k * D * D;
    const double* mu = means + (size_t)k * D;
    float* sck = sc + (size_t)k * 2 * DP;
    for (int j = threadIdx.x; j < DP; j += blockDim.x) {
        double amax = 0.0, bp = 0.0;
        if (j < D) {
            for (int d = 0; d <= j; ++d) {
                const double l = L[(size_t)d * D + j];
                amax = fmax(amax, fabs(l * xinfo[DP + d]));
                bp = fma(mu[d] - xinfo[d], l, bp);
            }
        }
        double scale = 1.0;
        if (amax > 0.0 && isfinite(amax)) {
            int e;
            frexp(amax, &e);
            scale = ldexp(1.0, e);
        }
        colinv[j] = 1.0 / scale;
        sck[j] = (float)scale;
        sck[DP + j] = (float)bp;
    }
    if (threadIdx.x == 0) {
        cst[2 * k] = aux[(size_t)k * (D + 2) + D];
        cst[2 * k + 1] = aux[(size_t)k * (D + 2) + D + 1];
    }
    __syncthreads();
    __half* hi = bt + (size_t)k * 2 * bmat_elems(DP);
    __half* lo = hi + bmat_elems(DP);
    const int kg_n = DP / 8;
    for (int e = threadIdx.x; e < DP * DP; e += blockDim.x) {
        const int j = e / DP, d = e - j * DP;
        double v = 0.0;
        if (j < D && d <= j) v = L[(size_t)d * D + j] * xinfo[DP + d] * colinv[j];
        const size_t o = ((size_t)(j >> 3) * kg_n + (d >> 3)) * 64 + (j & 7) * 8 + (d & 7);
        split_store(v, hi + o, lo + o);
    }
}

// ------------------------------------------------------------------------------------------
// The tcgen05 E-step kernel.
// ------------------------------------------------------------------------------------------
struct EstepSmem {
    // byte offsets into dynamic shared memory
    uint32_t a, b_hi, b_lo, scl, cst, bars, tmem_ptr, total;
};
__host__ __device__ inline EstepSmem estep_smem(int DP) {
    EstepSmem s;
    uint32_t o = 0;
    s.a = o;     o += 2u * TILE_M * DP * 2;            // hi then lo
    s.b_hi = o;  o += 2u * DP * DP * 2;                // two stages
    s.b_lo = o;  o += (uint32_t)DP * DP * 2;
    s.scl = o;   o += 2u * 2 * DP * 4;                 // two stages of [scale | bprime]
    s.cst = o;   o += 2u * 2 * 8;
    s.bars = o;  o += 16 * 8;
    s.tmem_ptr = o; o += 16;
    s.total = o;
    return s;
}

enum { BAR_A_FULL = 0, BAR_A_EMPTY, BAR_BLO_FULL, BAR_BLO_EMPTY, BAR_BHI_FULL0, BAR_BHI_FULL1,
       BAR_BHI_EMPTY0, BAR_BHI_EMPTY1, BAR_TM_FULL0, BAR_TM_FULL1, BAR_TM_EMPTY0, BAR_TM_EMPTY1 };

__global__ void __launch_bounds__(192, 1)
estep_tc_kernel(long long N, long long Npad, int n_tiles, int K, int D, int DP,
                uint32_t tmem_cols, const __half* __restrict__ xt, const __half* __restrict__ bt,
                const float* __restrict__ sc, const double* __restrict__ cst,
                double* __restrict__ wlpT, int mode, int32_t* __restrict__ mix,
                int32_t* __restrict__ cand, double near_tie) {
    extern __shared__ __align__(128) unsigned char smem[];
    const EstepSmem L = estep_smem(DP);
    __half* a_hi = reinterpret_cast<__half*>(smem + L.a);
    __half* a_lo = a_hi + tile_elems(DP);
    __half* b_hi0 = reinterpret_cast<__half*>(smem + L.b_hi);
    __half* b_lo = reinterpret_cast<__half*>(smem + L.b_lo);
    float* scl = reinterpret_cast<float*>(smem + L.scl);
    double* cst_s = reinterpret_cast<double*>(smem + L.cst);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L.tmem_ptr);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t a_bytes = 2u * TILE_M * DP * 2;
    const uint32_t b_bytes = (uint32_t)DP * DP * 2;
    const uint32_t lbo = 128, sbo = (uint32_t)(DP / 8) * 128;
    const int ksteps = DP / 16;
    const uint32_t idesc = make_idesc(TILE_M, DP);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 12; ++i) mbar_init(bars + i, (i >= BAR_TM_EMPTY0) ? 4u : 1u);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ---------------- producer ----------------
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                for (int k = 0; k < K; ++k) {
                    const uint32_t g = (uint32_t)it * K + k, s = g & 1u, u = g >> 1;
                    const __half* bk = bt + (size_t)k * 2 * bmat_elems(DP);
                    mbar_wait(bars + BAR_BLO_EMPTY, (g & 1u) ^ 1u);
                    mbar_expect_tx(bars + BAR_BLO_FULL, b_bytes);
                    bulk_g2s(b_lo, bk + bmat_elems(DP), b_bytes, bars + BAR_BLO_FULL);
                    mbar_wait(bars + BAR_BHI_EMPTY0 + s, (u & 1u) ^ 1u);
                    mbar_expect_tx(bars + BAR_BHI_FULL0 + s, b_bytes);
                    bulk_g2s(b_hi0 + (size_t)s * bmat_elems(DP), bk, b_bytes,
                             bars + BAR_BHI_FULL0 + s);
                    if (k == 0) {
                        mbar_wait(bars + BAR_A_EMPTY, ((uint32_t)it & 1u) ^ 1u);
                        mbar_expect_tx(bars + BAR_A_FULL, a_bytes);
                        bulk_g2s(a_hi, xt + (size_t)tile * 2 * tile_elems(DP), a_bytes,
                                 bars + BAR_A_FULL);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t a_hi_addr = smem_u32(a_hi), a_lo_addr = smem_u32(a_lo);
            const uint32_t b_lo_addr = smem_u32(b_lo);
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                for (int k = 0; k < K; ++k) {
                    const uint32_t g = (uint32_t)it * K + k, s = g & 1u, u = g >> 1;
                    const uint32_t b_hi_addr = smem_u32(b_hi0 + (size_t)s * bmat_elems(DP));
                    const uint32_t acc = tmem_base + s * (uint32_t)DP;
                    mbar_wait(bars + BAR_TM_EMPTY0 + s, (u & 1u) ^ 1u);
                    mbar_wait(bars + BAR_BLO_FULL, g & 1u);
                    if (k == 0) mbar_wait(bars + BAR_A_FULL, (uint32_t)it & 1u);
                    tc_fence_after();
                    for (int ks = 0; ks < ksteps; ++ks)       // x_hi . l_lo
                        umma_f16(acc, make_desc(a_hi_addr + ks * 256, lbo, sbo),
                                 make_desc(b_lo_addr + ks * 256, lbo, sbo), idesc, ks > 0);
                    umma_commit(bars + BAR_BLO_EMPTY);
                    mbar_wait(bars + BAR_BHI_FULL0 + s, u & 1u);
                    tc_fence_after();
                    for (int ks = 0; ks < ksteps; ++ks)       // x_lo . l_hi
                        umma_f16(acc, make_desc(a_lo_addr + ks * 256, lbo, sbo),
                                 make_desc(b_hi_addr + ks * 256, lbo, sbo), idesc, 1u);
                    // acc = acc * 2^-11 + x_hi . l_hi
                    umma_f16_scaled(acc, make_desc(a_hi_addr, lbo, sbo),
                                    make_desc(b_hi_addr, lbo, sbo), idesc);
                    for (int ks = 1; ks < ksteps; ++ks)
                        umma_f16(acc, make_desc(a_hi_addr + ks * 256, lbo, sbo),
                                 make_desc(b_hi_addr + ks * 256, lbo, sbo), idesc, 1u);
                    umma_commit(bars + BAR_BHI_EMPTY0 + s);
                    umma_commit(bars + BAR_TM_FULL0 + s);
                }
                umma_commit(bars + BAR_A_EMPTY);
            }
        }
    } else {
        // ---------------- epilogue (warps 2..5) ----------------
        const int et = threadIdx.x - 64;            // 0..127
        const uint32_t quarter = (uint32_t)(warp & 3);
        const int row = (int)quarter * 32 + lane;   // TMEM lane = row of the tile
        const double LOG2PI = 1.8378770664093453;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const long long n = (long long)tile * TILE_M + row;
            double v1 = -CUDART_INF, v2 = -CUDART_INF;
            int k1 = 0, k2 = -1;
            for (int k = 0; k < K; ++k) {
                const uint32_t g = (uint32_t)it * K + k, s = g & 1u, u = g >> 1;
                float* sk = scl + (size_t)s * 2 * DP;
                for (int i = et; i < 2 * DP; i += 128) sk[i] = sc[(size_t)k * 2 * DP + i];
                if (et < 2) cst_s[s * 2 + et] = cst[2 * k + et];
                asm volatile("bar.sync 1, 128;" ::: "memory");
                mbar_wait(bars + BAR_TM_FULL0 + s, u & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + s * (uint32_t)DP;
                double q = 0.0;
                for (int c = 0; c < ksteps; ++c) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c * 16, v);
                    tmem_ld_wait();
                    float part = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float y = fmaf(__uint_as_float(v[j]), sk[c * 16 + j],
                                             -sk[DP + c * 16 + j]);
                        part = fmaf(y, y, part);
                    }
                    q += (double)part;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + BAR_TM_EMPTY0 + s);
                const double wlp = (-0.5 * ((double)D * LOG2PI + q) + cst_s[s * 2]) + cst_s[s * 2 + 1];
                if (mode == 0) {
                    if (n < N) wlpT[(size_t)k * Npad + n] = wlp;
                } else if (wlp > v1) {
                    v2 = v1; k2 = k1; v1 = wlp; k1 = k;
                } else if (wlp > v2) {
                    v2 = wlp; k2 = k;
                }
            }
            if (mode == 1 && n < N) {
                mix[n] = k1;
                cand[n] = (K > 1 && v1 - v2 < near_tie) ? k2 : -1;   // runner-up to re-check in fp64
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// Responsibilities from the weighted log-probabilities (component-major), one thread per frame;
// mode 1: hard argmax (first maximum) instead.
__global__ void lse_kernel(long long N, long long Npad, int K, double* __restrict__ respT,
                           double* __restrict__ lse_partial, int mode, int32_t* __restrict__ mix) {
    __shared__ double sh[128];
    const long long n = (long long)blockIdx.x * 128 + threadIdx.x;
    double lse = 0.0;
    if (n < N) {
        double* col = respT + n;
        double mx = -CUDART_INF;
        int arg = 0;
        for (int k = 0; k < K; ++k) {
            const double v = col[(size_t)k * Npad];
            if (v > mx) { mx = v; arg = k; }
        }
        if (mode == 1) {
            mix[n] = arg;
        } else {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += exp(col[(size_t)k * Npad] - mx);
            lse = log(s) + mx;
            for (int k = 0; k < K; ++k) col[(size_t)k * Npad] = exp(col[(size_t)k * Npad] - lse);
        }
    }
    if (mode == 1) return;
    sh[threadIdx.x] = lse;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 128; ++i) t += sh[i];
        lse_partial[blockIdx.x] = t;
    }
}

// Hard-assignment near-ties: frames whose two best components are closer than the tensor-core
// error budget are re-evaluated for those two components in FP64 (one warp per frame), so the
// mixture sequence equals the FP64 argmax.
__global__ void refine_argmax_kernel(long long N, int D, const double* __restrict__ X,
                                     const double* __restrict__ prec_chol,
                                     const double* __restrict__ aux, int32_t* __restrict__ mix,
                                     const int32_t* __restrict__ cand) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const double LOG2PI = 1.8378770664093453;
    for (long long n = warp; n < N; n += n_warps) {
        const int kb = cand[n];
        if (kb < 0) continue;
        const int ka = mix[n];
        double w[2];
        for (int c = 0; c < 2; ++c) {
            const int k = c == 0 ? ka : kb;
            const double* Lk = prec_chol + (size_t)k * D * D;
            const double* ak = aux + (size_t)k * (D + 2);
            double q = 0.0;
            for (int j = lane; j < D; j += 32) {
                double y = 0.0;
                for (int d = 0; d <= j; ++d) y = fma(X[n * D + d], Lk[(size_t)d * D + j], y);
                y -= ak[j];
                q = fma(y, y, q);
            }
            q = warp_sum(q);
            w[c] = (-0.5 * ((double)D * LOG2PI + q) + ak[D]) + ak[D + 1];
        }
        if (lane == 0) {
            const bool take_b = (w[1] > w[0]) || (w[1] == w[0] && kb < ka);
            if (take_b) mix[n] = kb;
        }
    }
}

}  // namespace tc

constexpr int TC_STAT_CHUNKS = 296;

struct TcWorkspace {
    double* colpartial;
    double* xinfo;
    __half* xt;
    __half* bt;
    float* sc;
    double* cst;
    double* lse_partial;
    int32_t* cand;
    size_t bytes;
};

static inline int tc_dp(int D) { return (D + 15) / 16 * 16; }

static TcWorkspace carve_tc(long long N, int K, int D, void* base) {
    const int DP = tc_dp(D);
    const long long n_tiles = (N + tc::TILE_M - 1) / tc::TILE_M;
    Carver c(base);
    TcWorkspace w;
    w.colpartial = c.take<double>((size_t)TC_STAT_CHUNKS * 3 * D);
    w.xinfo = c.take<double>(2 * (size_t)DP);
    w.xt = c.take<__half>((size_t)n_tiles * 2 * tc::tile_elems(DP));
    w.bt = c.take<__half>((size_t)K * 2 * tc::bmat_elems(DP));
    w.sc = c.take<float>((size_t)K * 2 * DP);
    w.cst = c.take<double>(2 * (size_t)K);
    w.lse_partial = c.take<double>((size_t)n_tiles + 1);
    w.cand = c.take<int32_t>((size_t)N);
    w.bytes = align_up(c.used, 256);
    return w;
}

size_t tc_workspace_bytes(long long N, int K, int D) { return carve_tc(N, K, D, nullptr).bytes; }

// wlp (or resp, or argmax) through the tensor-core path.  mode 0: resp + sum of logsumexp into
// lse_out[0] and N into lse_out[1];  mode 1: hard labels into mix (resp is scratch, K x Npad).
int estep_tc(long long N, const double* X, int K, int D, const double* means, const double* pc,
             const double* aux, double* resp, double* lse_out, int mode, int32_t* mix,
             void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int DP = tc_dp(D);
    if (DP > 144) {
        set_error("dim %d > 144 is not supported by the tensor-core E-step", D);
        return KW_ERR_UNSUPPORTED;
    }
    TcWorkspace w = carve_tc(N, K, D, workspace);
    if (w.bytes > workspace_bytes) {
        set_error("tensor-core E-step workspace too small: need %zu bytes, got %zu", w.bytes,
                  workspace_bytes);
        return KW_ERR_WORKSPACE;
    }
    const long long n_tiles = (N + tc::TILE_M - 1) / tc::TILE_M;
    const long long Npad = resp_pad(N);
    const long long fpc = (N + TC_STAT_CHUNKS - 1) / TC_STAT_CHUNKS;
    tc::colstats_partial_kernel<<<TC_STAT_CHUNKS, 160, 0, st>>>(N, D, X, w.colpartial, fpc);
    KW_CUDA_CHECK(cudaGetLastError());
    tc::colstats_final_kernel<<<1, 160, 0, st>>>(N, D, DP, TC_STAT_CHUNKS, w.colpartial, w.xinfo);
    KW_CUDA_CHECK(cudaGetLastError());
    tc::pack_x_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(N, D, DP, X, w.xinfo, w.xt);
    KW_CUDA_CHECK(cudaGetLastError());
    tc::pack_l_kernel<<<K, 256, sizeof(double) * DP, st>>>(K, D, DP, means, pc, aux, w.xinfo,
                                                          w.bt, w.sc, w.cst);
    KW_CUDA_CHECK(cudaGetLastError());

    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const tc::EstepSmem L = tc::estep_smem(DP);
    KW_CUDA_CHECK(cudaFuncSetAttribute(tc::estep_tc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    uint32_t cols = 32;
    while (cols < 2u * DP) cols <<= 1;
    const int grid = (int)std::min<long long>(n_tiles, sms);
    tc::estep_tc_kernel<<<grid, 192, L.total, st>>>(N, Npad, (int)n_tiles, K, D, DP, cols, w.xt,
                                                    w.bt, w.sc, w.cst, resp, mode, mix, w.cand,
                                                    0.05);
    KW_CUDA_CHECK(cudaGetLastError());
    if (mode == 1) {
        tc::refine_argmax_kernel<<<sms * 4, 256, 0, st>>>(N, D, X, pc, aux, mix, w.cand);
        KW_CUDA_CHECK(cudaGetLastError());
        return KW_OK;
    }
    const unsigned lgrid = (unsigned)((N + 127) / 128);
    tc::lse_kernel<<<lgrid, 128, 0, st>>>(N, Npad, K, resp, w.lse_partial, mode, mix);
    KW_CUDA_CHECK(cudaGetLastError());
    if (mode == 0) {
        launch_reduce_fixed(w.lse_partial, (long long)lgrid, (double)N, lse_out, st);
        KW_CUDA_CHECK(cudaGetLastError());
    }
    return KW_OK;
}

}  // namespace kw

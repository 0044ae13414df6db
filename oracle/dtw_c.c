/* CPU restatement of FastDTW in plain C -- TEST INFRASTRUCTURE (see oracle/__init__.py).
 *
 * Restates fastdtw==0.3.2 (Pipfile.lock:68; third-party, absent from /root/reference)
 * as the reference calls it at kwiiyatta/vocoder/align.py:71
 *     fastdtw.fastdtw(x_feature, y_feature, dist=2, radius=radius)
 * PARITY UNPINNED against that package (not installable here); cross-checked against
 * oracle/fastdtw_ref.py (literal set-based window) in tests/test_oracle_dtw.py.
 *
 * Cell rule: D[0][0]=0 at the virtual origin, unwritten cells are +inf,
 *   D[i][j] = min over (up, left, diag) of (D[pred] + d(x[i-1], y[j-1])), the three sums
 *   compared AFTER the addition, first minimum wins in the order up, left, diag.
 * Window: one inclusive column interval [lo, hi] per row, derived from the coarser path
 *   (closed form of the radius expansion + 2x projection; see window_intervals in
 *   oracle/fastdtw_ref.py).
 * The rule above is fastdtw's pure-Python back-end (tie rule 0).  kwo_fastdtw_ex takes a general
 *   tie rule: the preference order of the three predecessors (a permutation of 0 = up, 1 = left,
 *   2 = diag) and whether they are compared before or after the local distance is added, so that
 *   tests can show the returned path to be the same under EVERY such rule (the Cython back-end
 *   of fastdtw 0.3.2, whose exact rule cannot be read here, is one of them), and reports the
 *   smallest decision margin on the returned path (see include/kwiiyatta_b200.h, margin_dev).
 * Local distance p=2: sqrt(sum_k (x_k - y_k)^2), summed left to right; use_fma selects
 *   s = fma(d, d, s) (what the CUDA path computes) or s = s + d*d (what a Python float
 *   loop computes).  p=1: sum_k |x_k - y_k|.
 *
 * Build:  gcc -O2 -ffp-contract=off -mfma -shared -fPIC -o oracle/_build/liboracle_dtw.so oracle/dtw_c.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static double local_dist(const double* a, const double* b, int F, int p, int use_fma) {
    double s = 0.0;
    if (p == 1) {
        for (int k = 0; k < F; ++k) s = s + fabs(a[k] - b[k]);
        return s;
    }
    if (use_fma) {
        for (int k = 0; k < F; ++k) { double d = a[k] - b[k]; s = fma(d, d, s); }
    } else {
        for (int k = 0; k < F; ++k) { double d = a[k] - b[k]; double q = d * d; s = s + q; }
    }
    return sqrt(s);
}

/* DP over rows with inclusive windows lo[i]..hi[i] (both non-decreasing).
 * path_out receives (i,j) pairs in forward order; returns path length. */
typedef struct {
    int order[3];   /* preference order of the candidates: 0 = up, 1 = left, 2 = diag */
    int before;     /* 1: compare the predecessors before adding the local distance */
} TieRule;

static int dp_window(const double* x, int Tx, const double* y, int Ty, int F, int p, int use_fma,
                     const int32_t* lo, const int32_t* hi, double* cost, int32_t* path_out,
                     int64_t* cells, const TieRule* tie, double* margin) {
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(Tx + 1));
    int64_t total = 0;
    for (int i = 0; i < Tx; ++i) { off[i] = total; total += (int64_t)(hi[i] - lo[i] + 1); }
    off[Tx] = total;
    if (cells) *cells = total;
    uint8_t* bp = (uint8_t*)malloc((size_t)total);
    double* prev = (double*)malloc(sizeof(double) * (size_t)(Ty + 1));
    double* cur = (double*)malloc(sizeof(double) * (size_t)(Ty + 1));
    double* mprev = (double*)malloc(sizeof(double) * (size_t)(Ty + 1));   /* running margins */
    double* mcur = (double*)malloc(sizeof(double) * (size_t)(Ty + 1));
    int plo = 0, phi = -1; /* previous row's window; row -1 is the virtual row */
    for (int i = 0; i < Tx; ++i) {
        const int l = lo[i], h = hi[i];
        for (int j = l; j <= h; ++j) {
            double pred[3], pm[3];       /* 0 = up, 1 = left, 2 = diag */
            if (i == 0) {
                pred[0] = INFINITY; pm[0] = INFINITY;
                pred[2] = (j == 0) ? 0.0 : INFINITY; pm[2] = INFINITY;
            } else {
                const int in_up = (j >= plo && j <= phi), in_dg = (j - 1 >= plo && j - 1 <= phi);
                pred[0] = in_up ? prev[j] : INFINITY;      pm[0] = in_up ? mprev[j] : INFINITY;
                pred[2] = in_dg ? prev[j - 1] : INFINITY;  pm[2] = in_dg ? mprev[j - 1] : INFINITY;
            }
            pred[1] = (j - 1 >= l) ? cur[j - 1] : INFINITY;
            pm[1] = (j - 1 >= l) ? mcur[j - 1] : INFINITY;
            const double dt = local_dist(x + (size_t)i * F, y + (size_t)j * F, F, p, use_fma);
            double cand[3];
            for (int q = 0; q < 3; ++q) cand[q] = tie->before ? pred[q] : pred[q] + dt;
            int code = tie->order[0];
            for (int q = 1; q < 3; ++q)
                if (cand[tie->order[q]] < cand[code]) code = tie->order[q];
            cur[j] = tie->before ? cand[code] + dt : cand[code];
            double other = INFINITY;
            for (int q = 0; q < 3; ++q)
                if (q != code && cand[q] < other) other = cand[q];
            const double here = (cand[code] < INFINITY) ? other - cand[code] : INFINITY;
            mcur[j] = here < pm[code] ? here : pm[code];
            bp[off[i] + (j - l)] = (uint8_t)code;
        }
        double* t = prev; prev = cur; cur = t;
        t = mprev; mprev = mcur; mcur = t;
        plo = l; phi = h;
    }
    *cost = (Ty - 1 >= plo && Ty - 1 <= phi) ? prev[Ty - 1] : INFINITY;
    if (margin) *margin = (Ty - 1 >= plo && Ty - 1 <= phi) ? mprev[Ty - 1] : INFINITY;
    /* backtrace */
    int n = 0, i = Tx - 1, j = Ty - 1;
    while (i >= 0 && j >= 0) {
        path_out[2 * n] = i; path_out[2 * n + 1] = j; ++n;
        if (j < lo[i] || j > hi[i]) { n = -1; break; } /* fell out of the window: malformed */
        const uint8_t code = bp[off[i] + (j - lo[i])];
        if (code == 0) --i; else if (code == 1) --j; else { --i; --j; }
        if (n > Tx + Ty) { n = -1; break; }
    }
    if (n > 0) {
        for (int a = 0, b = n - 1; a < b; ++a, --b) {
            int32_t ti = path_out[2 * a], tj = path_out[2 * a + 1];
            path_out[2 * a] = path_out[2 * b]; path_out[2 * a + 1] = path_out[2 * b + 1];
            path_out[2 * b] = ti; path_out[2 * b + 1] = tj;
        }
    }
    free(off); free(bp); free(prev); free(cur); free(mprev); free(mcur);
    return n;
}

static int fastdtw_rec(const double* x, int Tx, const double* y, int Ty, int F, int radius, int p,
                       int use_fma, double* cost, int32_t* path_out, int64_t* cells,
                       const TieRule* tie, double* margin_all) {
    int32_t* lo = (int32_t*)malloc(sizeof(int32_t) * (size_t)Tx);
    int32_t* hi = (int32_t*)malloc(sizeof(int32_t) * (size_t)Tx);
    int n;
    if (radius < 0 || Tx < radius + 2 || Ty < radius + 2) {
        for (int i = 0; i < Tx; ++i) { lo[i] = 0; hi[i] = Ty - 1; }
    } else {
        const int cx = Tx / 2, cy = Ty / 2;
        double* xs = (double*)malloc(sizeof(double) * (size_t)cx * F);
        double* ys = (double*)malloc(sizeof(double) * (size_t)cy * F);
        for (int i = 0; i < cx; ++i)
            for (int k = 0; k < F; ++k)
                xs[(size_t)i * F + k] = (x[(size_t)(2 * i) * F + k] + x[(size_t)(2 * i + 1) * F + k]) / 2;
        for (int i = 0; i < cy; ++i)
            for (int k = 0; k < F; ++k)
                ys[(size_t)i * F + k] = (y[(size_t)(2 * i) * F + k] + y[(size_t)(2 * i + 1) * F + k]) / 2;
        int32_t* cpath = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)(cx + cy + 2));
        double ccost;
        int cn = fastdtw_rec(xs, cx, ys, cy, F, radius, p, use_fma, &ccost, cpath, cells, tie,
                             margin_all);
        free(xs); free(ys);
        if (cn <= 0) { free(cpath); free(lo); free(hi); return -1; }
        int32_t* first_j = (int32_t*)malloc(sizeof(int32_t) * (size_t)cx);
        int32_t* last_j = (int32_t*)malloc(sizeof(int32_t) * (size_t)cx);
        for (int i = 0; i < cx; ++i) first_j[i] = -1;
        for (int t = 0; t < cn; ++t) {
            int ci = cpath[2 * t], cj = cpath[2 * t + 1];
            if (first_j[ci] < 0) first_j[ci] = cj;
            last_j[ci] = cj;
        }
        for (int a = 0; a < Tx; ++a) {
            int ca = a / 2;
            int r0 = ca - radius; if (r0 < 0) r0 = 0;
            int r1 = ca + radius; if (r1 > cx - 1) r1 = cx - 1;
            if (r0 > r1) { free(cpath); free(first_j); free(last_j); free(lo); free(hi); return -1; }
            int l = 2 * (first_j[r0] - radius); if (l < 0) l = 0;
            int h = 2 * (last_j[r1] + radius) + 1; if (h > Ty - 1) h = Ty - 1;
            lo[a] = l; hi[a] = h;
        }
        free(cpath); free(first_j); free(last_j);
    }
    int64_t c = 0;
    double m = INFINITY;
    n = dp_window(x, Tx, y, Ty, F, p, use_fma, lo, hi, cost, path_out, &c, tie, &m);
    if (cells) *cells += c;
    /* margin_all[0]: this (finally: the finest) level; [1]: minimum over all levels so far */
    margin_all[0] = m;
    if (m < margin_all[1]) margin_all[1] = m;
    free(lo); free(hi);
    return n;
}

/* radius < 0: exhaustive DTW.  path_out: capacity 2*(Tx+Ty) int32.  cells_out: sum of
 * window sizes over all resolution levels.  Returns path length, or -1 on a malformed window. */
int kwo_fastdtw_ex(const double* x, int Tx, const double* y, int Ty, int F, int radius, int p,
                   int use_fma, const int32_t* tie_order, int tie_before, double* cost,
                   int32_t* path_out, int64_t* cells_out, double* margin_out) {
    int64_t cells = 0;
    double margins[2] = {INFINITY, INFINITY};
    TieRule tie = {{tie_order[0], tie_order[1], tie_order[2]}, tie_before};
    if (Tx <= 0 || Ty <= 0) { *cost = 0.0; if (cells_out) *cells_out = 0; return 0; }
    int n = fastdtw_rec(x, Tx, y, Ty, F, radius, p, use_fma, cost, path_out, &cells, &tie, margins);
    if (cells_out) *cells_out = cells;
    if (margin_out) { margin_out[0] = margins[0]; margin_out[1] = margins[1]; }
    return n;
}

int kwo_fastdtw(const double* x, int Tx, const double* y, int Ty, int F, int radius, int p,
                int use_fma, double* cost, int32_t* path_out, int64_t* cells_out) {
    const int32_t python_order[3] = {0, 1, 2};
    return kwo_fastdtw_ex(x, Tx, y, Ty, F, radius, p, use_fma, python_order, 0, cost, path_out,
                          cells_out, NULL);
}

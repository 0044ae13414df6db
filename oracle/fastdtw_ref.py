"""Pure-Python restatement of FastDTW (TEST INFRASTRUCTURE - see oracle/__init__.py).

Follows the published algorithm of ``fastdtw==0.3.2`` (Salvador & Chan,
"FastDTW: Toward Accurate Dynamic Time Warping in Linear Time and Space"),
pure-Python back-end, as called by the reference at
kwiiyatta/vocoder/align.py:71 ``fastdtw.fastdtw(x_feature, y_feature, dist=2,
radius=radius)`` and at tests/kwiiyatta/test_vocoder.py:281-286.

PARITY UNPINNED: the package is not installed here and the reference tree has
no golden vectors; this file is written from the published algorithm.

Two formulations of the search window are kept on purpose:

* ``expand_window_sets``      - the literal set construction (every coarse
  path cell grown by ``radius`` in both axes, each coarse cell projected to a
  2x2 block, then one contiguous run per row with non-decreasing starts);
* ``expand_window_intervals`` - the closed form the device code uses
  (per fine row ``a``: ``S = {j : (i, j) in path, |i - a//2| <= radius}``,
  columns ``[2*(min S - radius), 2*(max S + radius) + 1]`` clipped).

tests/test_oracle_dtw.py checks that they agree.
"""
import math

import numpy as np

INF = float('inf')


def local_dist_numpy(a, b):
    """``np.linalg.norm(a - b, 2)``: what fastdtw's ``dist=2`` evaluates."""
    return float(np.linalg.norm(a - b, 2))


def local_dist_seq(a, b):
    """sqrt of the left-to-right sum of squares, no fused multiply-add."""
    s = 0.0
    for u, v in zip(a.tolist(), b.tolist()):
        d = u - v
        s = s + d * d
    return math.sqrt(s)


def local_dist_l1(a, b):
    s = 0.0
    for u, v in zip(a.tolist(), b.tolist()):
        s = s + abs(u - v)
    return s


_DISTS = {'numpy': local_dist_numpy, 'seq': local_dist_seq, 'l1': local_dist_l1}


def prep_inputs(x, y):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.ndim == y.ndim > 1 and x.shape[1] != y.shape[1]:
        raise ValueError('second dimension of x and y must be the same')
    if x.ndim == 1:
        x = x[:, None]
    if y.ndim == 1:
        y = y[:, None]
    return x, y


def reduce_by_half(x):
    """(x[2i] + x[2i+1]) / 2, odd tail dropped."""
    n = len(x) - len(x) % 2
    return (x[0:n:2] + x[1:n:2]) / 2


def dtw_window(x, y, window, dist):
    """DP restricted to ``window`` = iterable of 0-based (i, j) in row-major order.

    D[0,0] = 0 at the virtual origin, everything not yet written is +inf.
    Candidates are compared AFTER adding the local distance and the first
    minimum wins in the order (i-1, j), (i, j-1), (i-1, j-1).
    """
    len_x, len_y = len(x), len(y)
    D = {(0, 0): (0.0, 0, 0)}

    def get(i, j):
        e = D.get((i, j))
        return INF if e is None else e[0]

    for i0, j0 in window:
        i, j = i0 + 1, j0 + 1
        dt = dist(x[i0], y[j0])
        best = (get(i - 1, j) + dt, i - 1, j)
        cand = (get(i, j - 1) + dt, i, j - 1)
        if cand[0] < best[0]:
            best = cand
        cand = (get(i - 1, j - 1) + dt, i - 1, j - 1)
        if cand[0] < best[0]:
            best = cand
        D[i, j] = best
    path = []
    i, j = len_x, len_y
    while not (i == 0 and j == 0):
        path.append((i - 1, j - 1))
        _, i, j = D[i, j]
    path.reverse()
    return D[len_x, len_y][0], path


def full_window(len_x, len_y):
    return [(i, j) for i in range(len_x) for j in range(len_y)]


def expand_window_sets(path, len_x, len_y, radius):
    grown = set(path)
    for i, j in path:
        for a in range(-radius, radius + 1):
            for b in range(-radius, radius + 1):
                grown.add((i + a, j + b))
    fine = set()
    for i, j in grown:
        fine.add((2 * i, 2 * j))
        fine.add((2 * i, 2 * j + 1))
        fine.add((2 * i + 1, 2 * j))
        fine.add((2 * i + 1, 2 * j + 1))
    window = []
    start_j = 0
    for i in range(len_x):
        new_start = None
        for j in range(start_j, len_y):
            if (i, j) in fine:
                window.append((i, j))
                if new_start is None:
                    new_start = j
            elif new_start is not None:
                break
        start_j = new_start
    return window


def window_intervals(path, len_x, len_y, radius):
    """Closed-form per-row column interval [lo, hi] (inclusive) of the window."""
    path = np.asarray(path, dtype=np.int64)
    n_coarse = int(path[-1, 0]) + 1
    first_j = np.full(n_coarse, -1, dtype=np.int64)
    last_j = np.full(n_coarse, -1, dtype=np.int64)
    for i, j in path:
        if first_j[i] < 0:
            first_j[i] = j
        last_j[i] = j
    lo = np.empty(len_x, dtype=np.int64)
    hi = np.empty(len_x, dtype=np.int64)
    for a in range(len_x):
        ca = a // 2
        r0 = max(0, ca - radius)
        r1 = min(n_coarse - 1, ca + radius)
        if r0 > r1:
            raise ValueError('window row is empty (radius too small)')
        lo[a] = max(0, 2 * (first_j[r0] - radius))
        hi[a] = min(len_y - 1, 2 * (last_j[r1] + radius) + 1)
    return lo, hi


def expand_window_intervals(path, len_x, len_y, radius):
    lo, hi = window_intervals(path, len_x, len_y, radius)
    return [(i, j) for i in range(len_x) for j in range(lo[i], hi[i] + 1)]


def _fastdtw(x, y, radius, dist, expand, levels):
    min_time_size = radius + 2
    if len(x) < min_time_size or len(y) < min_time_size:
        out = dtw_window(x, y, full_window(len(x), len(y)), dist)
        if levels is not None:
            levels.append((len(x), len(y), len(x) * len(y), out[1]))
        return out
    xs = reduce_by_half(x)
    ys = reduce_by_half(y)
    _, path = _fastdtw(xs, ys, radius, dist, expand, levels)
    window = expand(path, len(x), len(y), radius)
    out = dtw_window(x, y, window, dist)
    if levels is not None:
        levels.append((len(x), len(y), len(window), out[1]))
    return out


def fastdtw(x, y, radius=1, dist=2, dist_mode='numpy', window='sets', levels=None):
    """(distance, path) with the semantics of ``fastdtw.fastdtw(x, y, radius, dist)``.

    ``dist`` : 2 (Euclidean, the reference's setting) or 1 / None (L1, the
    package default for 2-D input).  ``dist_mode`` picks how the Euclidean
    local distance is rounded ('numpy' = BLAS dot as the package does, 'seq' =
    sequential sum).  ``levels`` (a list) receives (Tx, Ty, window cells, path)
    per resolution level, coarsest first.
    """
    x, y = prep_inputs(x, y)
    if dist is not None and dist <= 0:
        raise ValueError('dist cannot be a negative integer')
    if dist is None or dist == 1:
        fn = _DISTS['l1']
    elif dist == 2:
        fn = _DISTS[dist_mode]
    else:
        raise NotImplementedError('oracle restates p=1 and p=2 only')
    expand = expand_window_sets if window == 'sets' else expand_window_intervals
    return _fastdtw(x, y, radius, fn, expand, levels)


def dtw(x, y, dist=2, dist_mode='numpy'):
    """Exhaustive DTW with the same cell rule (``fastdtw.dtw``)."""
    x, y = prep_inputs(x, y)
    if dist is None or dist == 1:
        fn = _DISTS['l1']
    else:
        fn = _DISTS[dist_mode]
    return dtw_window(x, y, full_window(len(x), len(y)), fn)

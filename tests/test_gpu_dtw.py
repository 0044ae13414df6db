"""CUDA DTW / FastDTW vs the oracle (through the C ABI via kwiiyatta_b200.fastdtw)."""
import numpy as np
import pytest

from kwiiyatta_b200 import fastdtw as kfd
from kwiiyatta_b200 import synth
from kwiiyatta_b200.alignment import make_feature
from oracle import dtw_c, fastdtw_ref

pytestmark = pytest.mark.gpu


def _rand_pair(rng, tx, ty, f):
    return rng.standard_normal((tx, f)), rng.standard_normal((ty, f))


def _check(pairs, radius, dist=2, exact_cost=True):
    got = kfd.fastdtw_batch(pairs, radius=radius, dist=dist)
    for (x, y), (cost, path) in zip(pairs, got):
        ecost, epath = dtw_c.fastdtw(x, y, radius=radius, dist=dist, use_fma=True)
        assert path.shape == epath.shape, (x.shape, y.shape, radius)
        assert np.array_equal(path, epath), (x.shape, y.shape, radius)
        if exact_cost:
            assert cost == ecost       # same summation order -> bit-exact
        else:
            assert abs(cost - ecost) <= 1e-12 * abs(ecost)


@pytest.mark.parametrize('radius', [1, 2, 5, 32])
def test_fastdtw_random_shapes(cuda, radius):
    rng = np.random.default_rng(radius)
    shapes = [(1, 1), (1, 7), (9, 1), (2, 2), (3, 5), (17, 33), (40, 55), (64, 64), (100, 37),
              (129, 257), (300, 280), (513, 300), (70, 400)]
    pairs = [_rand_pair(rng, tx, ty, 6) for tx, ty in shapes]
    _check(pairs, radius)


@pytest.mark.parametrize('f', [1, 2, 8, 9, 16, 25, 26, 32])
def test_fastdtw_feature_dims(cuda, f):
    rng = np.random.default_rng(100 + f)
    pairs = [_rand_pair(rng, 90, 120, f), _rand_pair(rng, 300, 260, f)]
    _check(pairs, 3)
    _check(pairs, -1)


def test_exhaustive_dtw_matches_python_reference(cuda):
    rng = np.random.default_rng(7)
    x, y = _rand_pair(rng, 45, 61, 4)
    cost, path = kfd.dtw(x, y, dist=2)
    ecost, epath = fastdtw_ref.dtw(x, y, dist=2, dist_mode='seq')
    assert path == epath
    assert abs(cost - ecost) <= 1e-13 * ecost
    # numpy-norm rounding of the local distance (what the package itself evaluates)
    ncost, npath = fastdtw_ref.dtw(x, y, dist=2, dist_mode='numpy')
    assert path == npath
    assert abs(cost - ncost) <= 1e-12 * ncost


def test_l1_and_default_dist(cuda):
    rng = np.random.default_rng(8)
    pairs = [_rand_pair(rng, 80, 95, 5), _rand_pair(rng, 33, 31, 5)]
    _check(pairs, 2, dist=1)
    x, y = pairs[0]
    cost, path = kfd.fastdtw(x, y, radius=2)            # dist=None -> 1-norm
    ecost, epath = fastdtw_ref.fastdtw(x, y, radius=2, dist=None)
    assert path == epath and abs(cost - ecost) <= 1e-12 * ecost
    # 1-D series
    a, b = rng.standard_normal(50), rng.standard_normal(64)
    cost, path = kfd.fastdtw(a, b, radius=1)
    ecost, epath = fastdtw_ref.fastdtw(a, b, radius=1, dist=None)
    assert path == epath and abs(cost - ecost) <= 1e-12 * ecost


def test_large_radius_equals_exhaustive(cuda):
    rng = np.random.default_rng(9)
    x, y = _rand_pair(rng, 150, 170, 8)
    c1, p1 = kfd.fastdtw(x, y, radius=200, dist=2)
    c2, p2 = kfd.dtw(x, y, dist=2)
    assert p1 == p2 and c1 == c2


def test_path_properties_and_cost_recomputed(cuda):
    rng = np.random.default_rng(10)
    x, y = _rand_pair(rng, 400, 333, 26)
    cost, path = kfd.fastdtw_batch([(x, y)], radius=32, dist=2)[0]
    assert tuple(path[0]) == (0, 0) and tuple(path[-1]) == (399, 332)
    step = np.diff(path, axis=0)
    assert ((step >= 0) & (step <= 1)).all() and (step.sum(axis=1) >= 1).all()
    local = np.sqrt(((x[path[:, 0]] - y[path[:, 1]]) ** 2).sum(axis=1))
    assert abs(local.sum() - cost) <= 1e-10 * cost


def test_errors(cuda):
    rng = np.random.default_rng(11)
    with pytest.raises(ValueError, match='second dimension of x and y must be the same'):
        kfd.fastdtw(rng.standard_normal((5, 3)), rng.standard_normal((5, 4)), dist=2)
    with pytest.raises(ValueError):
        kfd.fastdtw(rng.standard_normal((5, 3)), rng.standard_normal((5, 3)), dist=-1)
    with pytest.raises(ValueError):
        kfd.fastdtw(np.zeros((0, 3)), rng.standard_normal((5, 3)), dist=2)
    assert kfd.fastdtw_batch([]) == []


def test_synthetic_corpus_pairs_radius32(cuda):
    """Config-1 shaped: 10 padded pairs, 26-dim DTW features, radius 32."""
    pairs = []
    for i in range(10):
        a, b = synth.make_padded_pair(i)
        pairs.append((make_feature(a, a.fs), make_feature(b, b.fs)))
    _check(pairs, 32)
    # cells = sum of window sizes over all levels, as the oracle counts them
    import torch
    tx = np.array([len(x) for x, _ in pairs], dtype=np.int32)
    ty = np.array([len(y) for _, y in pairs], dtype=np.int32)
    res = kfd.fastdtw_batch_device(torch.from_numpy(np.concatenate([x for x, _ in pairs])).cuda(),
                                   torch.from_numpy(np.concatenate([y for _, y in pairs])).cuda(),
                                   tx, ty, radius=32, dist=2)
    cells = res.cells.cpu().numpy()
    for p, (x, y) in enumerate(pairs):
        assert cells[p] == dtw_c.fastdtw(x, y, 32, 2, return_cells=True)[2]


def test_ragged_batch_and_order_independence(cuda):
    rng = np.random.default_rng(12)
    shapes = [(600, 500), (35, 900), (900, 40), (2, 3), (257, 255)]
    pairs = [_rand_pair(rng, tx, ty, 26) for tx, ty in shapes]
    a = kfd.fastdtw_batch(pairs, radius=4, dist=2)
    b = kfd.fastdtw_batch(pairs[::-1], radius=4, dist=2)[::-1]
    for (c1, p1), (c2, p2) in zip(a, b):
        assert c1 == c2 and np.array_equal(p1, p2)
    _check(pairs, 4)


def test_long_unconstrained_pair(cuda):
    """Config-4 shaped (scaled to what the C oracle finishes in seconds): 1536 x 1536, F=26."""
    a, b = synth.make_pair(3, length=1536)
    x, y = make_feature(a, a.fs), make_feature(b, b.fs)
    _check([(x, y)], -1)


def test_fp32_distance_mode_cost_within_1e6(cuda):
    """north_star: the fp32 mode is reported separately, path cost within 1e-6 relative."""
    pairs = []
    for i in range(6):
        a, b = synth.make_padded_pair(20 + i)
        pairs.append((make_feature(a, a.fs), make_feature(b, b.fs)))
    exact = kfd.fastdtw_batch(pairs, radius=32, dist=2, precision=0)
    fast = kfd.fastdtw_batch(pairs, radius=32, dist=2, precision=1)
    same = 0
    for (x, y), (c0, p0), (c1, p1) in zip(pairs, exact, fast):
        assert abs(c1 - c0) <= 1e-6 * c0
        # the fp32 path is a valid warping path whose fp64 cost is (near-)optimal as well
        step = np.diff(p1, axis=0)
        assert ((step >= 0) & (step <= 1)).all() and (step.sum(axis=1) >= 1).all()
        local = np.sqrt(((x[p1[:, 0]] - y[p1[:, 1]]) ** 2).sum(axis=1))
        assert abs(local.sum() - c0) <= 1e-6 * c0
        same += int(p0.shape == p1.shape and np.array_equal(p0, p1))
    assert same >= 4       # identical paths except on near-ties


def _plateau_pair(rng, tx, ty, f, jump_x, jump_y):
    """Two step-like sequences whose steps sit at very different places: the optimal path has
    long horizontal / vertical runs, so the windows of the finer levels are far wider than
    8 radius + 2 (wider than the stored distance slice: exercises the in-place distance path)."""
    x = np.where(np.arange(tx)[:, None] < jump_x, 0.0, 4.0) + 0.01 * rng.standard_normal((tx, f))
    y = np.where(np.arange(ty)[:, None] < jump_y, 0.0, 4.0) + 0.01 * rng.standard_normal((ty, f))
    return x, y


@pytest.mark.parametrize('radius', [1, 3])
def test_wide_windows_overflow_the_distance_slice(cuda, radius):
    rng = np.random.default_rng(40 + radius)
    pairs = [_plateau_pair(rng, 600, 640, 3, 60, 560), _plateau_pair(rng, 513, 300, 3, 450, 30),
             _plateau_pair(rng, 257, 700, 3, 128, 650), _rand_pair(rng, 333, 444, 3)]
    got = kfd.fastdtw_batch(pairs, radius=radius, dist=2)
    widest = 0
    for (x, y), (cost, path) in zip(pairs, got):
        ecost, epath, ecells = dtw_c.fastdtw(x, y, radius=radius, dist=2, use_fma=True,
                                             return_cells=True)
        assert np.array_equal(path, epath)
        assert cost == ecost
        # the longest run of the path in one row bounds the window width from below
        runs = np.bincount(path[:, 0])
        widest = max(widest, int(runs.max()))
    assert widest > max(64, 12 * radius + 16)      # the slice capacity really was exceeded


@pytest.mark.parametrize('tx,ty', [(63, 64), (64, 65), (65, 129), (127, 128), (128, 64), (191, 257)])
def test_strip_boundaries_of_the_banded_sweep(cuda, tx, ty):
    """Lengths around the 64-row strips / 16-column back-pointer groups of the banded sweep."""
    rng = np.random.default_rng(tx * 1000 + ty)
    pairs = [_rand_pair(rng, tx, ty, 5), _rand_pair(rng, ty, tx, 5)]
    _check(pairs, 2)
    _check(pairs, 7)
    got = kfd.fastdtw_batch(pairs, radius=2, dist=2, precision=1)
    for (x, y), (cost, path) in zip(pairs, got):
        ecost, _ = dtw_c.fastdtw(x, y, radius=2, dist=2, use_fma=True)
        assert abs(cost - ecost) <= 1e-6 * abs(ecost)


def test_batch_rejects_mixed_feature_dims(cuda):
    rng = np.random.default_rng(3)
    with pytest.raises(ValueError, match='same number of features'):
        kfd.fastdtw_batch([_rand_pair(rng, 10, 12, 3), _rand_pair(rng, 10, 12, 2)], radius=1, dist=2)


def _corpus_pairs(n):
    out = []
    for i in range(n):
        a, b = synth.make_padded_pair(i)
        out.append((make_feature(a, a.fs), make_feature(b, b.fs)))
    return out


@pytest.mark.parametrize('tie_mode', ['python', 'cython'])
@pytest.mark.parametrize('radius', [1, 32, -1])
def test_tie_modes_on_exact_ties(cuda, tie_mode, radius):
    """Integer-valued sequences (exact ties everywhere): each tie mode of kw_dtw_batch reproduces
    the oracle under the same rule, bit for bit, and the margin of such a path is 0."""
    rng = np.random.default_rng(17)
    pairs = [(rng.integers(0, 3, (tx, 2)).astype(float), rng.integers(0, 3, (ty, 2)).astype(float))
             for tx, ty in [(40, 45), (130, 90), (257, 300), (64, 64), (5, 9)]]
    got, margins = kfd.fastdtw_batch(pairs, radius=radius, dist=2, tie_mode=tie_mode,
                                     with_margin=True)
    for (x, y), (cost, path), m in zip(pairs, got, margins):
        ecost, epath, em = dtw_c.fastdtw(x, y, radius=radius, dist=2, tie=tie_mode,
                                         return_margin=True)
        assert np.array_equal(path, epath) and cost == ecost
        assert np.array_equal(m, em)
    assert (margins[:4, 0] == 0.0).all()
    # the two modes really are different rules
    other = kfd.fastdtw_batch(pairs, radius=radius, dist=2,
                              tie_mode='cython' if tie_mode == 'python' else 'python')
    assert any(not np.array_equal(a[1], b[1]) for a, b in zip(got, other))


def test_decision_margin_on_the_corpus(cuda):
    """configs[0-1] inputs: both tie modes return the same paths, and the smallest decision
    margin on every path equals the oracle's and is far above the rounding of the sums -- the
    path is pinned without the fastdtw package (kwiiyatta/vocoder/align.py:71)."""
    pairs = _corpus_pairs(24)
    for radius in (32, 1):
        got, margins = kfd.fastdtw_batch(pairs, radius=radius, dist=2, with_margin=True)
        cy = kfd.fastdtw_batch(pairs, radius=radius, dist=2, tie_mode='cython')
        for (x, y), (cost, path), (ccost, cpath), m in zip(pairs, got, cy, margins):
            ecost, epath, em = dtw_c.fastdtw(x, y, radius=radius, dist=2, return_margin=True)
            assert np.array_equal(path, epath) and cost == ecost
            assert np.array_equal(cpath, path) and abs(ccost - cost) <= 64 * np.spacing(cost)
            assert np.allclose(m, em, rtol=0, atol=16 * np.spacing(cost))
            assert m[1] <= m[0] and m[1] > 1e4 * np.spacing(cost)
    # margins are reported for exact local distances only
    with pytest.raises(NotImplementedError):
        kfd.fastdtw_batch(pairs[:1], radius=32, dist=2, precision=1, with_margin=True)


def test_margin_variant_returns_the_same_paths(cuda):
    rng = np.random.default_rng(5)
    pairs = [_rand_pair(rng, tx, ty, 6) for tx, ty in [(300, 280), (129, 257), (70, 400), (3, 5)]]
    for radius in (2, -1):
        plain = kfd.fastdtw_batch(pairs, radius=radius, dist=2)
        with_m, margins = kfd.fastdtw_batch(pairs, radius=radius, dist=2, with_margin=True)
        for (c1, p1), (c2, p2) in zip(plain, with_m):
            assert c1 == c2 and np.array_equal(p1, p2)
        assert (margins > 0).all()


def test_packed_host_api_equals_the_list_api(cuda):
    import torch
    rng = np.random.default_rng(21)
    pairs = [_rand_pair(rng, int(tx), int(ty), 6)
             for tx, ty in rng.integers(20, 200, (301, 2))]
    tx = [len(x) for x, _ in pairs]
    ty = [len(y) for _, y in pairs]
    xh = torch.from_numpy(np.concatenate([x for x, _ in pairs])).pin_memory()
    yh = torch.from_numpy(np.concatenate([y for _, y in pairs])).pin_memory()
    expected = kfd.fastdtw_batch(pairs, radius=3, dist=2)
    for n_chunks in (None, 1, 3):
        got = kfd.fastdtw_batch_packed(xh, yh, tx, ty, radius=3, dist=2, n_chunks=n_chunks)
        assert len(got) == len(expected)
        for (c1, p1), (c2, p2) in zip(got, expected):
            assert c1 == c2 and np.array_equal(p1, p2)


def test_configs3_pair_4096_unconstrained(cuda):
    """configs[3]: one 4096 x 4096 pair of 26-dim features, unconstrained window (fastdtw.dtw):
    path and cost bit-exact against the C oracle, and the path's decision margin."""
    a, b = synth.make_pair(0, length=4096)
    x, y = make_feature(a, a.fs), make_feature(b, b.fs)
    assert x.shape == (4096, 26) and y.shape == (4096, 26)
    (res,), margins = kfd.fastdtw_batch([(x, y)], radius=-1, dist=2, with_margin=True)
    ecost, epath, em = dtw_c.fastdtw(x, y, radius=-1, dist=2, return_margin=True)
    assert np.array_equal(res[1], epath) and res[0] == ecost
    assert np.allclose(margins[0], em, rtol=0, atol=64 * np.spacing(ecost))
    assert margins[0][0] > 1e4 * np.spacing(ecost)
    cy = kfd.fastdtw_batch([(x, y)], radius=-1, dist=2, tie_mode='cython')[0]
    assert np.array_equal(cy[1], epath)

"""CPU: the DTW oracle against itself (independent formulations) and the committed fixtures."""
import os

import numpy as np
import pytest

from oracle import dtw_c, fastdtw_ref

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'dtw.npz')


def _pairs(seed, shapes, f=5):
    rng = np.random.default_rng(seed)
    return [(rng.standard_normal((tx, f)), rng.standard_normal((ty, f))) for tx, ty in shapes]


@pytest.mark.parametrize('radius', [1, 2, 4])
def test_set_window_equals_interval_window(radius):
    for x, y in _pairs(radius, [(30, 41), (64, 33), (17, 90), (55, 55), (7, 9)]):
        c1, p1 = fastdtw_ref.fastdtw(x, y, radius, 2, 'seq', window='sets')
        c2, p2 = fastdtw_ref.fastdtw(x, y, radius, 2, 'seq', window='intervals')
        assert p1 == p2 and c1 == c2


@pytest.mark.parametrize('radius', [1, 3, 32])
def test_c_restatement_equals_python(radius):
    for x, y in _pairs(10 + radius, [(40, 55), (100, 90), (1, 6), (6, 1), (2, 2), (130, 77)]):
        c1, p1 = fastdtw_ref.fastdtw(x, y, radius, 2, 'seq')
        c2, p2 = dtw_c.fastdtw(x, y, radius, 2, use_fma=False)
        assert p1 == [tuple(t) for t in p2.tolist()]
        assert c1 == c2
        c3, p3 = dtw_c.fastdtw(x, y, radius, 2, use_fma=True)
        assert np.array_equal(p2, p3) and abs(c3 - c1) <= 1e-13 * c1
        c4, p4 = fastdtw_ref.fastdtw(x, y, radius, 2, 'numpy')   # BLAS-dot rounding
        assert p4 == p1 and abs(c4 - c1) <= 1e-13 * c1


def test_l1_and_1d():
    rng = np.random.default_rng(3)
    x, y = rng.standard_normal(40), rng.standard_normal(52)
    c1, p1 = fastdtw_ref.fastdtw(x, y, 1, None)
    c2, p2 = dtw_c.fastdtw(x, y, 1, None)
    assert p1 == [tuple(t) for t in p2.tolist()] and abs(c1 - c2) <= 1e-13 * c1


def test_large_radius_is_exhaustive():
    for x, y in _pairs(4, [(50, 60), (33, 20)]):
        assert fastdtw_ref.fastdtw(x, y, 100, 2, 'seq') == fastdtw_ref.dtw(x, y, 2, 'seq')
        c1, p1 = dtw_c.fastdtw(x, y, 100, 2)
        c2, p2 = dtw_c.dtw(x, y, 2)
        assert c1 == c2 and np.array_equal(p1, p2)


def test_exhaustive_against_full_matrix_dp():
    x, y = _pairs(5, [(25, 31)])[0]
    d = np.sqrt(((x[:, None, :] - y[None, :, :]) ** 2).sum(-1))
    D = np.full((26, 32), np.inf)
    D[0, 0] = 0
    for i in range(1, 26):
        for j in range(1, 32):
            D[i, j] = d[i - 1, j - 1] + min(D[i - 1, j], D[i, j - 1], D[i - 1, j - 1])
    cost, path = dtw_c.dtw(x, y, 2)
    assert abs(cost - D[25, 31]) <= 1e-12 * cost
    assert abs(d[path[:, 0], path[:, 1]].sum() - cost) <= 1e-12 * cost


def test_path_invariants_and_errors():
    x, y = _pairs(6, [(80, 70)])[0]
    cost, path = dtw_c.fastdtw(x, y, 2, 2)
    assert tuple(path[0]) == (0, 0) and tuple(path[-1]) == (79, 69)
    step = np.diff(path, axis=0)
    assert ((step >= 0) & (step <= 1)).all() and (step.sum(1) >= 1).all()
    with pytest.raises(ValueError, match='second dimension'):
        fastdtw_ref.fastdtw(np.zeros((3, 2)), np.zeros((3, 3)))
    with pytest.raises(ValueError):
        fastdtw_ref.fastdtw(x, y, 1, -2)


def test_golden_fixtures():
    g = np.load(GOLDEN)
    for i in range(int(g['n'])):
        x, y, r = g[f'x{i}'], g[f'y{i}'], int(g[f'radius{i}'])
        cost, path = dtw_c.fastdtw(x, y, r, 2, use_fma=False)
        assert np.array_equal(path, g[f'path{i}']) and cost == float(g[f'cost{i}'])


def _corpus_features(n):
    from kwiiyatta_b200 import synth
    from oracle import align_ref
    out = []
    for i in range(n):
        a, b = synth.make_padded_pair(i)
        out.append((align_ref.make_feature(a.mel_cepstrum.data, a.f0, a.is_voiced),
                    align_ref.make_feature(b.mel_cepstrum.data, b.f0, b.is_voiced)))
    return out


def test_path_does_not_depend_on_the_tie_rule_or_the_rounding():
    """The pin SURVEY.md sections 7 / 8b ask for in place of the (absent) fastdtw package: on the
    silence-padded synthetic corpus the path is the same under all 12 tie rules (3! preference
    orders x compared before / after the local distance is added; the pure-Python and the Cython
    back-end of fastdtw 0.3.2 are two of them) and under all three roundings of the local
    distance (fma, separate multiply-add, BLAS dot), because the smallest decision margin on the
    path is many orders of magnitude above the rounding of the sums."""
    for x, y in _corpus_features(10):
        cost, path, margin = dtw_c.fastdtw(x, y, radius=32, dist=2, return_margin=True)
        ulp = np.spacing(cost)
        assert margin[0] > 1e4 * ulp and margin[1] > 1e4 * ulp
        assert margin[1] <= margin[0]
        for rule in dtw_c.all_tie_rules():
            c, p = dtw_c.fastdtw(x, y, radius=32, dist=2, tie=rule)
            assert np.array_equal(p, path) and abs(c - cost) <= 64 * ulp
        c, p = dtw_c.fastdtw(x, y, radius=32, dist=2, use_fma=False)
        assert np.array_equal(p, path)
    # radius 1 (the reference's own tests call fastdtw(..., radius=1, dist=2),
    # tests/kwiiyatta/test_vocoder.py:281-286)
    for x, y in _corpus_features(3):
        cost, path, margin = dtw_c.fastdtw(x, y, radius=1, dist=2, return_margin=True)
        assert margin[1] > 1e4 * np.spacing(cost)
        for rule in dtw_c.all_tie_rules():
            assert np.array_equal(dtw_c.fastdtw(x, y, radius=1, dist=2, tie=rule)[1], path)


def test_tie_rules_do_differ_on_exact_ties():
    """The margin is not vacuous: integer-valued sequences tie exactly, the margin is 0 and
    the rules return different (equally cheap) paths."""
    rng = np.random.default_rng(8)
    x = rng.integers(0, 3, (40, 1)).astype(float)
    y = rng.integers(0, 3, (45, 1)).astype(float)
    results = {}
    for rule in dtw_c.all_tie_rules():
        cost, path, margin = dtw_c.fastdtw(x, y, radius=-1, dist=2, tie=rule, return_margin=True)
        assert margin[0] == 0.0
        results[rule] = (cost, path.tobytes())
    assert len({c for c, _ in results.values()}) == 1          # same optimal cost
    assert len({p for _, p in results.values()}) > 1           # different optimal paths
    py = dtw_c.fastdtw(x, y, radius=-1, dist=2, tie='python')[1]
    assert [tuple(t) for t in py.tolist()] == fastdtw_ref.dtw(x, y, 2, 'seq')[1]

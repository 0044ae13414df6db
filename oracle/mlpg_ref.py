"""Posterior + MLPG oracle (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates ``nnmnkwii.baseline.gmm.MLPG(gmm, windows, diff).transform(src)`` and
``nnmnkwii.paramgen.mlpg`` (nnmnkwii==0.0.17 with bandmat==0.7, Pipfile.lock:86,19; both
third-party and absent) as called at kwiiyatta/converter/gmm.py:28-34.
PARITY UNPINNED against the packages; ``mlpg`` is cross-checked between a dense solve of
(W^T D^-1 W) c = W^T D^-1 E and scipy's banded Cholesky in tests/test_oracle_mlpg.py.
"""
import numpy as np
import scipy.linalg

from . import gmm_ref
from .delta_ref import DELTA_WINDOWS


def split_joint(weights, means, covariances, diff=False):
    """MLPGBase.__init__: slice the joint model at Dh = D/2, apply the diff rewrite."""
    dh = means.shape[1] // 2
    src_means = means[:, :dh]
    tgt_means = means[:, dh:]
    cxx = covariances[:, :dh, :dh]
    cxy = covariances[:, :dh, dh:]
    cyx = covariances[:, dh:, :dh]
    cyy = covariances[:, dh:, dh:]
    if diff:
        tgt_means = tgt_means - src_means
        cyy = cxx + cyy - cxy - cyx
        cxy = cxy - cxx
        cyx = cxy.transpose(0, 2, 1)
    return dict(weights=weights, src_means=src_means, tgt_means=tgt_means,
                cxx=cxx, cxy=cxy, cyx=cyx, cyy=cyy,
                px_prec_chol=gmm_ref.precision_cholesky(np.ascontiguousarray(cxx)))


def predict(src, model):
    """px.predict: argmax_k of log N(x; mu_x, Sigma_xx) + log w."""
    wlp = gmm_ref.weighted_log_prob(src, model['weights'], model['src_means'],
                                    model['px_prec_chol'])
    return wlp.argmax(axis=1), wlp


def window_matrix(window, t):
    """T x T matrix with W[t, t + k - l] = coeff[k], truncated at the edges."""
    l, u, coeff = window
    w = np.zeros((t, t))
    for k, c in enumerate(coeff):
        o = k - l
        for r in range(t):
            if 0 <= r + o < t:
                w[r, r + o] = c
    return w


def mlpg_dense(mean_frames, variance_frames, windows=DELTA_WINDOWS):
    t, d = mean_frames.shape
    sd = d // len(windows)
    mats = [window_matrix(w, t) for w in windows]
    y = np.zeros((t, sd))
    for k in range(sd):
        p = np.zeros((t, t))
        b = np.zeros(t)
        for wi, w in enumerate(mats):
            prec = 1 / variance_frames[:, wi * sd + k]
            b += w.T @ (prec * mean_frames[:, wi * sd + k])
            p += w.T @ (prec[:, None] * w)
        y[:, k] = np.linalg.solve(p, b)
    return y


def mlpg_banded(mean_frames, variance_frames, windows=DELTA_WINDOWS):
    """Same systems assembled directly in banded (pentadiagonal) storage and solved with
    scipy's banded Cholesky, as bandmat's solveh does."""
    t, d = mean_frames.shape
    sd = d // len(windows)
    y = np.zeros((t, sd))
    for k in range(sd):
        ab = np.zeros((3, t))  # lower form: ab[b, i] = P[i + b, i]
        b = np.zeros(t)
        for wi, (l, _, coeff) in enumerate(windows):
            prec = 1 / variance_frames[:, wi * sd + k]
            bs = prec * mean_frames[:, wi * sd + k]
            for r in range(t):
                for ka, ca in enumerate(coeff):
                    i = r + ka - l
                    if not 0 <= i < t:
                        continue
                    b[i] += ca * bs[r]
                    for kb, cb in enumerate(coeff):
                        j = r + kb - l
                        if 0 <= j < t and j >= i:
                            ab[j - i, i] += ca * prec[r] * cb
        y[:, k] = scipy.linalg.solveh_banded(ab, b, lower=True)
    return y


def transform(src, weights, means, covariances, diff=False, windows=DELTA_WINDOWS,
              banded=True, return_internals=False):
    """MLPG.transform (hard mixture sequence, diagonal conditional covariance, MLPG)."""
    src = np.asarray(src, dtype=np.float64)
    model = split_joint(weights, means, covariances, diff)
    t, fd = src.shape
    mix, _ = predict(src, model)
    e = np.empty((t, fd))
    dv = np.empty((t, fd))
    for i in range(t):
        m = mix[i]
        xx = np.linalg.solve(model['cxx'][m], src[i] - model['src_means'][m])
        e[i] = model['tgt_means'][m] + model['cyx'][m] @ xx
        dv[i] = (np.diag(model['cyy'][m]) - np.diag(model['cyx'][m])
                 / np.diag(model['cxx'][m]) * np.diag(model['cxy'][m]))
    y = (mlpg_banded if banded else mlpg_dense)(e, dv, windows)
    if return_internals:
        return y, mix, e, dv
    return y


def transform_frames_soft(src, weights, means, covariances, diff=False):
    """MLPGBase.transform: per-frame soft-posterior conditional mean (the mlpg=False branch
    of kwiiyatta/converter/gmm.py:30-31)."""
    src = np.asarray(src, dtype=np.float64)
    model = split_joint(weights, means, covariances, diff)
    _, wlp = predict(src, model)
    post = np.exp(wlp - wlp.max(axis=1, keepdims=True))
    post /= post.sum(axis=1, keepdims=True)
    k = len(weights)
    out = np.zeros_like(src)
    for i, x in enumerate(src):
        em = np.empty((k, src.shape[1]))
        for m in range(k):
            xx = np.linalg.solve(model['cxx'][m], x - model['src_means'][m])
            em[m] = model['tgt_means'][m] + model['cyx'][m] @ xx
        out[i] = post[i] @ em
    return out


def transform_vectorised(src, weights, means, covariances, diff=False, model=None):
    """The same conversion written the way a numpy user would for speed (SURVEY.md section 8d asks
    for this beside the faithful per-frame loop as the CPU baseline): frames grouped by mixture,
    one solve per mixture (A_m = S_yx S_xx^-1), banded systems of all static dimensions assembled
    with array operations and solved by scipy's banded Cholesky.  ``model`` = split_joint(...)
    to amortise the slicing over utterances."""
    src = np.asarray(src, dtype=np.float64)
    if model is None:
        model = split_joint(weights, means, covariances, diff)
    t, fd = src.shape
    sd = fd // len(DELTA_WINDOWS)
    mix, _ = predict(src, model)
    e = np.empty((t, fd))
    dv = np.empty((t, fd))
    for m in np.unique(mix):
        rows = mix == m
        a = np.linalg.solve(model['cxx'][m], model['cyx'][m].T).T          # S_yx S_xx^-1
        e[rows] = model['tgt_means'][m] + (src[rows] - model['src_means'][m]) @ a.T
        dv[rows] = (np.diag(model['cyy'][m]) - np.diag(model['cyx'][m])
                    / np.diag(model['cxx'][m]) * np.diag(model['cxy'][m]))
    prec = 1.0 / dv
    pm = prec * e
    # window coefficients at offsets -1, 0, +1 (DELTA_WINDOWS), zero padded at the edges
    coeff = np.zeros((len(DELTA_WINDOWS), 3))
    for w, (l, _, c) in enumerate(DELTA_WINDOWS):
        coeff[w, 1 - l:1 - l + len(c)] = c
    y = np.empty((t, sd))
    ab = np.zeros((3, t))
    for k in range(sd):
        ab[:] = 0.0
        b = np.zeros(t)
        for w in range(len(DELTA_WINDOWS)):
            p = prec[:, w * sd + k]
            q = pm[:, w * sd + k]
            for oa in (-1, 0, 1):                  # row r touches column r + oa
                ca = coeff[w, oa + 1]
                if ca == 0.0:
                    continue
                r = np.arange(max(0, -oa), min(t, t - oa))
                np.add.at(b, r + oa, ca * q[r])
                for ob in (-1, 0, 1):
                    cb = coeff[w, ob + 1]
                    if cb == 0.0 or ob < oa:
                        continue
                    rr = r[(r + ob >= 0) & (r + ob < t)]
                    np.add.at(ab[ob - oa], rr + oa, ca * cb * p[rr])
        y[:, k] = scipy.linalg.solveh_banded(ab, b, lower=True)
    return y

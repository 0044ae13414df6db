#!/usr/bin/env python
"""Randomised parity sweep on the GPU (not part of the test suite; run before a release):
DTW batches of random shapes / radii / feature dims / norms against the C oracle, conversion
batches with ragged utterance lengths against the numpy oracle, tensor-core E-step with random
(K, D) against the FP64 oracle."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kwiiyatta_b200 import fastdtw as kfd
from kwiiyatta_b200.gmm import GaussianMixture
from kwiiyatta_b200.mlpg import MLPG
from oracle import dtw_c, gmm_ref, mlpg_ref

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rng = np.random.default_rng(seed)
t0 = time.time()
n_dtw = 0
for trial in range(40):
    f = int(rng.choice([1, 2, 3, 8, 13, 16, 25, 26, 32]))
    radius = int(rng.choice([-1, 1, 2, 3, 7, 15, 32]))   # (radius 0 breaks the package itself on odd lengths)
    p = int(rng.choice([1, 2]))
    n = int(rng.integers(1, 7))
    pairs = []
    for _ in range(n):
        tx, ty = (int(v) for v in rng.integers(1, 420, 2))
        kind = rng.integers(0, 3)
        if kind == 0:
            x, y = rng.standard_normal((tx, f)), rng.standard_normal((ty, f))
        elif kind == 1:      # smooth, correlated (speech-like)
            base = np.cumsum(rng.standard_normal((max(tx, ty) + 8, f)), 0) * 0.3
            x = base[np.sort(rng.integers(0, len(base), tx))] + 0.05 * rng.standard_normal((tx, f))
            y = base[np.sort(rng.integers(0, len(base), ty))] + 0.05 * rng.standard_normal((ty, f))
        else:                # plateaus -> wide windows
            x = np.where(np.arange(tx)[:, None] < tx * rng.uniform(), 0.0, 3.0) + 0.01 * rng.standard_normal((tx, f))
            y = np.where(np.arange(ty)[:, None] < ty * rng.uniform(), 0.0, 3.0) + 0.01 * rng.standard_normal((ty, f))
        pairs.append((x, y))
    got = kfd.fastdtw_batch(pairs, radius=radius, dist=p)
    for (x, y), (cost, path) in zip(pairs, got):
        ecost, epath = dtw_c.fastdtw(x, y, radius=radius, dist=p, use_fma=True)
        assert path.shape == epath.shape and np.array_equal(path, epath), (trial, x.shape, y.shape, radius, p)
        assert cost == ecost, (trial, cost, ecost)
        n_dtw += 1
print(f'dtw: {n_dtw} pairs bit-exact ({time.time() - t0:.1f} s)')

t0 = time.time()
for trial in range(6):
    k = int(rng.integers(1, 9))
    d = int(rng.choice([4, 9, 16, 33, 48, 72, 100, 144]))
    n = int(rng.integers(300, 2500))
    centres = rng.standard_normal((k, d)) * 2
    lab = rng.integers(0, k, n)
    a = rng.standard_normal((k, d, d)) * 0.3 + np.eye(d)
    x = centres[lab] + np.einsum('nij,nj->ni', a[lab], rng.standard_normal((n, d)))
    resp0 = gmm_ref.kmeans_like_resp(x, k, trial)
    ref = gmm_ref.numpy_em(x, resp0, max_iter=2, tol=0.0)
    gm = GaussianMixture(n_components=k, precision='tc').set_parameters(
        ref['weights'], ref['means'], ref['covariances'])
    lb, log_resp = gmm_ref.e_step(x, ref['weights'], ref['means'], ref['precisions_cholesky'])
    assert abs(gm.score(x) - lb) <= 1e-5 * abs(lb), (trial, k, d, gm.score(x), lb)
    assert np.abs(gm.predict_proba(x) - np.exp(log_resp)).max() <= 2e-3, (trial, k, d)
print(f'tc e-step: 6 random (K, D) within 1e-5 ({time.time() - t0:.1f} s)')
print('fuzz ok, seed', seed)

"""Generates tests/golden/*.npz from the oracle (run from the repo root:
``python tests/golden/make_golden.py``).

The reference tree holds no golden vectors for this path and its third-party callees
(fastdtw, nnmnkwii, bandmat) are not installable here, so these fixtures pin the ORACLE's
behaviour (regression vectors), not the reference packages'.  The EM fixture additionally
stores what the installed scikit-learn produced from the same injected initialisation."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from kwiiyatta_b200 import synth  # noqa: E402
from oracle import align_ref, delta_ref, dtw_c, fastdtw_ref, gmm_ref, mlpg_ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dtw_cases():
    rng = np.random.default_rng(20260101)
    out = {}
    shapes = [(1, 1, 3, 1), (7, 5, 2, 1), (40, 55, 6, 1), (37, 80, 6, 3), (100, 90, 26, 5),
              (64, 64, 26, 32), (130, 77, 4, 2), (90, 95, 26, -1)]
    for idx, (tx, ty, f, r) in enumerate(shapes):
        x = rng.standard_normal((tx, f))
        y = rng.standard_normal((ty, f))
        if r >= 0:
            cost, path = fastdtw_ref.fastdtw(x, y, radius=r, dist=2, dist_mode='seq')
        else:
            cost, path = fastdtw_ref.dtw(x, y, dist=2, dist_mode='seq')
        out[f'x{idx}'] = x
        out[f'y{idx}'] = y
        out[f'radius{idx}'] = r
        out[f'cost{idx}'] = cost
        out[f'path{idx}'] = np.array(path, dtype=np.int32)
    # one padded synthetic pair with kwiiyatta's defaults (26-dim features, radius 32)
    a, b = synth.make_padded_pair(0)
    xf = align_ref.make_feature(a.mel_cepstrum.data, a.f0, a.is_voiced)
    yf = align_ref.make_feature(b.mel_cepstrum.data, b.f0, b.is_voiced)
    cost, path, cells = dtw_c.fastdtw(xf, yf, radius=32, dist=2, use_fma=True, return_cells=True)
    out['synth_cost'] = cost
    out['synth_path'] = path
    out['synth_cells'] = cells
    out['n'] = len(shapes)
    np.savez_compressed(os.path.join(HERE, 'dtw.npz'), **out)


def em_case():
    rng = np.random.default_rng(20260102)
    n, d, k = 1500, 12, 4
    x = rng.standard_normal((n, d)) + rng.integers(0, k, n)[:, None] * 1.5
    resp0 = gmm_ref.kmeans_like_resp(x, k, 0)
    ref = gmm_ref.sklearn_em(x, resp0, max_iter=30)
    np.savez_compressed(os.path.join(HERE, 'em.npz'), x=x, labels0=resp0.argmax(1),
                        weights=ref['weights'], means=ref['means'],
                        covariances=ref['covariances'],
                        precisions_cholesky=ref['precisions_cholesky'],
                        lower_bounds=np.array(ref['lower_bounds']), n_iter=ref['n_iter'],
                        converged=ref['converged'])


def mlpg_case():
    w, m, c = synth.make_joint_gmm(4, dim_half=12, seed=5, static_dim=4)
    rng = np.random.default_rng(20260103)
    src = delta_ref.delta_features(rng.standard_normal((60, 4)).cumsum(0) * 0.1)
    out = dict(weights=w, means=m, covariances=c, src=src)
    for diff in (False, True):
        y, mix, e, dv = mlpg_ref.transform(src, w, m, c, diff=diff, return_internals=True)
        out[f'y_diff{int(diff)}'] = y
        out[f'mix_diff{int(diff)}'] = mix
    np.savez_compressed(os.path.join(HERE, 'mlpg.npz'), **out)


def em_scale_case():
    """sklearn's own fit to convergence (tol = 1e-3, the reference's default) of the 64-mix model
    on the joint frames of 60 synthetic pairs (N ~ 21 k, D = 144): minutes of CPU, so it is
    stored instead of recomputed by the GPU test.  Covariances are kept as their diagonals plus
    a fixed random sample of entries (the full array is 10 MB)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from util import oracle_joint_array
    x, _ = oracle_joint_array(60)
    resp0 = gmm_ref.kmeans_like_resp(x, 64, 0)
    ref = gmm_ref.sklearn_em(x, resp0, max_iter=100, tol=1e-3)
    cov = ref['covariances']
    idx = np.random.default_rng(20260104).integers(0, cov.size, 20000)
    np.savez_compressed(os.path.join(HERE, 'em_scale.npz'), n_frames=len(x),
                        x_checksum=float(x.sum()), labels0=resp0.argmax(1).astype(np.int16),
                        weights=ref['weights'], means=ref['means'],
                        cov_diag=np.einsum('kii->ki', cov), cov_idx=idx,
                        cov_sample=cov.ravel()[idx],
                        lower_bounds=np.array(ref['lower_bounds']), n_iter=ref['n_iter'],
                        converged=ref['converged'])


if __name__ == '__main__':
    which = sys.argv[1:] or ['dtw', 'em', 'mlpg', 'em_scale']
    if 'dtw' in which:
        dtw_cases()
    if 'em' in which:
        em_case()
    if 'mlpg' in which:
        mlpg_case()
    if 'em_scale' in which:
        em_scale_case()
    print('golden fixtures written to', HERE)

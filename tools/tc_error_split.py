"""Which stage of the tensor-core EM limits its accuracy?  Fits the same problems with the E-step
and the M-step statistics independently on the tcgen05 path or the FP64 path and prints the
error of each combination against the FP64 oracle (run on a GPU box)."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from kwiiyatta_b200.gmm import GaussianMixture  # noqa: E402
from oracle import gmm_ref  # noqa: E402
from util import rel_err  # noqa: E402


def blobs(rng, n, d, k, sep):
    centres = rng.standard_normal((k, d)) * sep
    lab = rng.integers(0, k, n)
    a = rng.standard_normal((k, d, d)) * 0.3 + np.eye(d)
    return centres[lab] + np.einsum('nij,nj->ni', a[lab], rng.standard_normal((n, d)))


def two_pass(gm):
    """Experiment: second accumulation of the M-step around the NEW means (no cancellation in
    S / n - delta delta^T)."""
    import torch
    from kwiiyatta_b200 import _lib  # noqa: F401
    orig = gm.em_iteration

    def em_iteration(x):
        lb = orig(x)
        c2 = gm._means[gm._cur]
        gm._accumulate(torch, x, c2)
        gm._finalize(torch, c2, weight_norm=0)
        return lb
    gm.em_iteration = em_iteration
    orig_init = gm.initialize

    def initialize(X):
        x = orig_init(X)
        c2 = gm._means[gm._cur]
        gm._stats[-1] = float(x.shape[0])
        gm._accumulate(torch, x, c2)
        gm._finalize(torch, c2, weight_norm=1)
        return x
    gm.initialize = initialize


def run(name, x, k, iters, seed):
    resp0 = gmm_ref.kmeans_like_resp(x, k, seed)
    ref = gmm_ref.numpy_em(x, resp0, max_iter=iters, tol=0.0)
    for pe, pm, tp in ((1, 1, 0), (1, 1, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0)):
        gm = GaussianMixture(n_components=k, max_iter=iters, tol=0.0, resp_init=resp0,
                             precision='tc', reorder_every=0)
        gm._precision_e, gm._precision_m = pe, pm
        if tp:
            two_pass(gm)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            gm.fit(x)
        lb = np.abs(np.array(gm.lower_bounds_) - np.array(ref['lower_bounds'])).max() / \
            abs(ref['lower_bound'])
        print(f'{name:28s} E={"tc" if pe else "f64"} M={"tc" if pm else "f64"}{"x2" if tp else "  "}  lb {lb:.2e}  '
              f'w {rel_err(gm.weights_, ref["weights"]):.2e}  '
              f'mu {rel_err(gm.means_, ref["means"]):.2e}  '
              f'cov {rel_err(gm.covariances_, ref["covariances"]):.2e}', flush=True)


if __name__ == '__main__':
    rng = np.random.default_rng(4097 * 3 + 48)
    run('4097x48 K5 sep2 8it', blobs(rng, 4097, 48, 5, 2.0), 5, 8, 0)
    rng = np.random.default_rng(11)
    x = blobs(rng, 20000, 48, 16, 0.6)
    run('20000x48 K16 sep.6 8it', x, 16, 8, 1)
    run('20000x48 K16 sep.6 1it', x, 16, 1, 1)

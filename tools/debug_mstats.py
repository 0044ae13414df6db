"""Compare the tensor-core and FP64 M-step statistics on the same inputs (debug aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kwiiyatta_b200 as kw
from kwiiyatta_b200 import _lib
from oracle import gmm_ref

def run(n, d, k, seed=0):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((k, d)) * 2.0
    lab = rng.integers(0, k, n)
    a = rng.standard_normal((k, d, d)) * 0.3 + np.eye(d)
    x = centres[lab] + np.einsum('nij,nj->ni', a[lab], rng.standard_normal((n, d)))
    resp0 = gmm_ref.kmeans_like_resp(x, k, 0)
    out = {}
    for prec in ('fp64', 'tc'):
        gm = kw.GaussianMixture(n_components=k, max_iter=1, tol=0.0, resp_init=resp0, precision='fp64')
        xd = gm.initialize(x)
        gm.em_iteration(xd)          # fp64 params after 1 iteration
        gm._estep(torch, xd)         # fp64 responsibilities
        resp = gm._resp.clone()
        cen = gm._means[gm._cur].clone()
        g2 = kw.GaussianMixture(n_components=k, precision=prec)
        g2._alloc(torch, n, d, xd.device); g2._pack(torch, xd)
        g2._resp.copy_(resp); g2._stats.zero_()
        g2._accumulate(torch, xd, cen)
        torch.cuda.synchronize()
        out[prec] = g2._stats.cpu().numpy()
    sb = 1 + d + d * d
    a64 = out['fp64'][:k * sb].reshape(k, sb); atc = out['tc'][:k * sb].reshape(k, sb)
    nk = a64[:, 0]
    print(f'n={n} d={d} k={k}')
    print('  n_k rel err      ', np.abs(atc[:, 0] - nk).max() / nk.max())
    m64, mtc = a64[:, 1:1 + d], atc[:, 1:1 + d]
    sd = np.sqrt(np.abs(np.stack([a64[i, 1 + d:].reshape(d, d).diagonal() for i in range(k)])) / nk[:, None])
    print('  m err / (n_k sd) ', (np.abs(mtc - m64) / (nk[:, None] * sd)).max(), ' signed mean', ((mtc - m64) / (nk[:, None] * sd)).mean())
    S64 = a64[:, 1 + d:].reshape(k, d, d); Stc = atc[:, 1 + d:].reshape(k, d, d)
    scale = nk[:, None, None] * sd[:, :, None] * sd[:, None, :]
    e = (Stc - S64) / scale
    print('  S err / (n sd sd) max', np.abs(e).max(), ' diag signed mean', np.stack([e[i].diagonal() for i in range(k)]).mean(), ' offdiag rms', np.sqrt((e**2).mean()))

for args in [(4097, 48, 5), (3000, 12, 4), (20000, 144, 8)]:
    run(*args)

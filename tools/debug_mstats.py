"""Compare the tensor-core and FP64 M-step statistics on the same inputs (debug aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kwiiyatta_b200 as kw
from kwiiyatta_b200 import _lib
from oracle import gmm_ref

def run(n, d, k, seed=0, sep=2.0, rscale=1.0):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((k, d)) * sep
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
        floor = float(os.environ.get('DBG_FLOOR', '0'))
        if floor > 0:
            resp[resp < floor] = 0.0
        if os.environ.get('DBG_SORT'):
            perm = torch.argsort(resp[:, :n].argmax(dim=0), stable=True)
            xd = xd.index_select(0, perm).contiguous()
            resp[:, :n] = resp[:, :n].index_select(1, perm)
        if os.environ.get('DBG_HIST') and prec == 'tc':
            r = resp[:, :n]
            tot = r.sum().item()
            for lo, hi in ((1e-1, 2), (1e-2, 1e-1), (1e-3, 1e-2), (1e-5, 1e-3), (1e-8, 1e-5), (0, 1e-8)):
                m = (r >= lo) & (r < hi)
                print(f'    r in [{lo:g},{hi:g}): count {int(m.sum())}  mass share {r[m].sum().item() / tot:.3e}')
        cen = gm._means[gm._cur].clone()
        g2 = kw.GaussianMixture(n_components=k, precision=prec)
        g2._alloc(torch, n, d, xd.device); g2._pack(torch, xd)
        g2._resp.copy_(resp * rscale); g2._stats.zero_()
        g2._accumulate(torch, xd, cen)
        torch.cuda.synchronize()
        out[prec] = g2._stats.cpu().numpy()
    sb = 1 + d + d * d
    a64 = out['fp64'][:k * sb].reshape(k, sb); atc = out['tc'][:k * sb].reshape(k, sb)
    nk = a64[:, 0]
    print(f'n={n} d={d} k={k} sep={sep} rscale={rscale}')
    print('  n_k rel err      ', np.abs(atc[:, 0] - nk).max() / nk.max())
    m64, mtc = a64[:, 1:1 + d], atc[:, 1:1 + d]
    sd = np.sqrt(np.abs(np.stack([a64[i, 1 + d:].reshape(d, d).diagonal() for i in range(k)])) / nk[:, None])
    print('  m err / (n_k sd) ', (np.abs(mtc - m64) / (nk[:, None] * sd)).max(), ' signed mean', ((mtc - m64) / (nk[:, None] * sd)).mean())
    S64 = a64[:, 1 + d:].reshape(k, d, d); Stc = atc[:, 1 + d:].reshape(k, d, d)
    scale = nk[:, None, None] * sd[:, :, None] * sd[:, None, :]
    e = (Stc - S64) / scale
    print('  S err / (n sd sd) max', np.abs(e).max(), ' diag signed mean', np.stack([e[i].diagonal() for i in range(k)]).mean(), ' offdiag rms', np.sqrt((e**2).mean()))

    # hypothesis: calibrate the tensor-core sums with the exact diagonal (ratio correction)
    g = np.stack([S64[i].diagonal() / Stc[i].diagonal() for i in range(k)])
    Scor = Stc * np.sqrt(g[:, :, None] * g[:, None, :])
    e2 = (Scor - S64) / scale
    off = ~np.eye(d, dtype=bool)
    print('  before: offdiag max', np.abs(e[:, off]).max(), ' after ratio correction: max', np.abs(e2).max(), ' rms', np.sqrt((e2**2).mean()))
    # derived parameters in the tests' norm (max abs error / max abs value)
    def params(a):
        nk_ = a[:, 0]
        dl = a[:, 1:1 + d] / nk_[:, None]
        cv = a[:, 1 + d:].reshape(k, d, d) / nk_[:, None, None] - dl[:, :, None] * dl[:, None, :]
        return dl, cv
    d64, c64 = params(a64); dtc, ctc = params(atc)
    print('  cov rel_err (test norm)', np.abs(ctc - c64).max() / np.abs(c64).max(),
          ' mean shift err / max|mu|', np.abs(dtc - d64).max())
    i = np.unravel_index(np.abs(ctc - c64).argmax(), c64.shape)
    print('  worst cov element', i, c64[i], ctc[i] - c64[i], 'diag there', c64[i[0], i[1], i[1]], c64[i[0], i[2], i[2]], 'n_k', nk[i[0]])

if len(sys.argv) > 1 and sys.argv[1] == 'rscale':
    for rs in (1.0, 0.1, 0.01, 1e-3):
        run(4097, 48, 5, 0, 2.0, rs)
elif len(sys.argv) > 1 and sys.argv[1] == 'soft':
    run(20000, 48, 16, 1, 0.6)
else:
    for args in [(4097, 48, 5), (20000, 48, 16, 1, 0.6), (20000, 144, 8)]:
        run(*args)

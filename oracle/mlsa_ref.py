"""Restatement of ``pysptk.mc2b`` (TEST INFRASTRUCTURE - see oracle/__init__.py).

Call site: kwiiyatta/filter/mlsa.py:24-29 (``b = pysptk.mc2b(mc.astype(np.float64), alpha=alpha)``
on the converted mel-cepstra with the power coefficient zeroed, :23).  pysptk (SPTK's mc2b) is
third-party and absent; the published recursion is  b[M] = mc[M],  b[m] = mc[m] - alpha b[m+1].
PARITY UNPINNED against the package; the inverse recursion (b2mc) is checked in
tests/test_oracle_mlpg.py."""
import numpy as np


def mc2b(mc, alpha):
    mc = np.asarray(mc, dtype=np.float64)
    b = np.empty_like(mc)
    b[..., -1] = mc[..., -1]
    for m in range(mc.shape[-1] - 2, -1, -1):
        b[..., m] = mc[..., m] - alpha * b[..., m + 1]
    return b


def b2mc(b, alpha):
    b = np.asarray(b, dtype=np.float64)
    mc = b.copy()
    mc[..., :-1] += alpha * b[..., 1:]
    return mc

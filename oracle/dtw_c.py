"""ctypes binding of oracle/dtw_c.c (TEST INFRASTRUCTURE - see oracle/__init__.py)."""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.kwo_fastdtw.restype = ctypes.c_int
        _lib.kwo_fastdtw.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.kwo_fastdtw_ex.restype = ctypes.c_int
        _lib.kwo_fastdtw_ex.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


# Tie rules: (preference order over 0 = up (i-1, j), 1 = left (i, j-1), 2 = diagonal; compare
# before the local distance is added?).  'python' is fastdtw.py's; 'cython' is _fastdtw.pyx's as
# recalled in SURVEY.md section 8a (the one kw_dtw_batch's tie_mode = 1 implements).
TIE_RULES = {'python': ((0, 1, 2), False), 'cython': ((2, 1, 0), True)}


def all_tie_rules():
    """Every preference order, compared before or after the addition (12 rules)."""
    import itertools
    return [(order, before) for order in itertools.permutations((0, 1, 2))
            for before in (False, True)]


def fastdtw(x, y, radius=1, dist=2, use_fma=True, return_cells=False, tie='python',
            return_margin=False):
    """(distance, path[(L,2) int32]) - radius < 0 means exhaustive DTW.  ``tie``: a name in
    TIE_RULES or an (order, before) pair.  ``return_margin`` appends the smallest decision
    margin on the path, (finest level, minimum over all levels)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    if y.ndim == 1:
        y = y[:, None]
    if x.shape[1] != y.shape[1]:
        raise ValueError('second dimension of x and y must be the same')
    if dist is not None and dist <= 0:
        raise ValueError('dist cannot be a negative integer')
    p = 1 if dist is None else int(dist)
    if p not in (1, 2):
        raise NotImplementedError('oracle restates p=1 and p=2 only')
    tx, ty = len(x), len(y)
    path = np.empty((tx + ty + 2, 2), dtype=np.int32)
    cost = ctypes.c_double()
    cells = ctypes.c_int64()
    order, before = TIE_RULES[tie] if isinstance(tie, str) else tie
    order = np.array(order, dtype=np.int32)
    margin = np.empty(2, dtype=np.float64)
    n = lib().kwo_fastdtw_ex(x.ctypes.data, tx, y.ctypes.data, ty, x.shape[1],
                             int(radius), p, int(bool(use_fma)), order.ctypes.data,
                             int(bool(before)), ctypes.byref(cost), path.ctypes.data,
                             ctypes.byref(cells), margin.ctypes.data)
    if n < 0:
        raise RuntimeError('malformed window')
    out = (cost.value, path[:n].copy())
    if return_cells:
        out = out + (cells.value,)
    if return_margin:
        out = out + (margin,)
    return out


def dtw(x, y, dist=2, use_fma=True):
    return fastdtw(x, y, radius=-1, dist=dist, use_fma=use_fma)

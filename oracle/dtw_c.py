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
    return _lib


def fastdtw(x, y, radius=1, dist=2, use_fma=True, return_cells=False):
    """(distance, path[(L,2) int32]) - radius < 0 means exhaustive DTW."""
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
    n = lib().kwo_fastdtw(x.ctypes.data, tx, y.ctypes.data, ty, x.shape[1],
                          int(radius), p, int(bool(use_fma)),
                          ctypes.byref(cost), path.ctypes.data, ctypes.byref(cells))
    if n < 0:
        raise RuntimeError('malformed window')
    out = (cost.value, path[:n].copy())
    if return_cells:
        out = out + (cells.value,)
    return out


def dtw(x, y, dist=2, use_fma=True):
    return fastdtw(x, y, radius=-1, dist=dist, use_fma=use_fma)

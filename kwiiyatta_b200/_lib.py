"""ctypes binding of libkwiiyatta_b200.so (the C ABI in include/kwiiyatta_b200.h).

There is no CPU fallback: if the library is missing or a call fails, this raises."""
import ctypes
import os

from . import build as _build

_c = ctypes
_lib = None

KW_ERR_INVALID = -1
KW_ERR_WORKSPACE = -2
KW_ERR_UNSUPPORTED = -3
KW_ERR_CUDA = -4

_vp, _i, _i64, _sz, _dbl = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_size_t, _c.c_double

SIGNATURES = {
    'kw_abi_version': (_i, []),
    'kw_last_error': (_c.c_char_p, []),
    'kw_device_info': (_i, [_i, _vp, _vp, _vp, _vp]),
    'kw_dtw_workspace_bytes': (_sz, [_i, _vp, _vp, _i, _i]),
    'kw_dtw_batch': (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp,
                          _vp, _sz, _vp]),
    'kw_delta_features': (_i, [_i, _vp, _i64, _i, _vp, _vp, _vp]),
    'kw_gmm_stats_len': (_sz, [_i, _i]),
    'kw_gmm_resp_len': (_sz, [_i64, _i]),
    'kw_gmm_workspace_bytes': (_sz, [_i64, _i, _i, _i]),
    'kw_gmm_pack_frames': (_i, [_i64, _vp, _i, _i, _i, _vp, _sz, _vp]),
    'kw_gmm_estep': (_i, [_i64, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    'kw_gmm_hard_labels': (_i, [_i64, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    'kw_gmm_mstep_accumulate': (_i, [_i64, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    'kw_gmm_mstep_finalize': (_i, [_i, _i, _dbl, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _vp]),
    'kw_gmm_precision_cholesky': (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'kw_convert_prepared_len': (_sz, [_i, _i]),
    'kw_convert_prepare': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'kw_convert_soft_workspace_bytes': (_sz, [_i64, _i, _i]),
    'kw_convert_soft_batch': (_i, [_i64, _vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    'kw_convert_workspace_bytes': (_sz, [_i64, _i, _i, _i]),
    'kw_convert_batch': (_i, [_i, _vp, _i64, _i, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _sz,
                              _vp]),
}


class KwError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = _build.LIB_PATH
        if not os.path.exists(path):
            raise KwError(
                f'{path} is missing: build it with `python -m kwiiyatta_b200.build` '
                '(nvcc, sm_100a). kwiiyatta_b200 has no CPU fallback.')
        handle = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc == 0:
        return
    msg = lib().kw_last_error().decode('utf-8', 'replace')
    text = f'{what} failed ({rc}): {msg}'
    if rc == KW_ERR_INVALID:
        raise ValueError(text)
    if rc == KW_ERR_UNSUPPORTED:
        raise NotImplementedError(text)
    raise KwError(text)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise KwError('kwiiyatta_b200 needs a CUDA device (B200, sm_100a); there is no CPU '
                      'fallback')
    return torch


def ptr(t):
    """Device (or host numpy) pointer as an int for ctypes."""
    if t is None:
        return None
    if hasattr(t, 'data_ptr'):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr(torch):
    return torch.cuda.current_stream().cuda_stream

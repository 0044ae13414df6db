"""Restatement of nnmnkwii.preprocessing helpers (TEST INFRASTRUCTURE - see oracle/__init__.py).

nnmnkwii==0.0.17 (Pipfile.lock:86) is third-party and absent; call sites:
kwiiyatta/converter/delta.py:30,46 (delta_features with DELTA_WINDOWS, :8-12) and
kwiiyatta/converter/dataset.py:51,70 (trim_zeros_frames, remove_zeros_frames).
PARITY UNPINNED against the package; ``np.correlate(..., mode='same')`` is the published
definition and is used literally here.
"""
import numpy as np

# kwiiyatta/converter/delta.py:8-12
DELTA_WINDOWS = [
    (0, 0, np.array([1.0])),
    (1, 1, np.array([-0.5, 0.0, 0.5])),
    (1, 1, np.array([1.0, -2.0, 1.0])),
]


def delta_features(x, windows=DELTA_WINDOWS):
    x = np.asarray(x)
    t, d = x.shape
    out = np.empty((t, d * len(windows)), dtype=x.dtype)
    for idx, (_, _, window) in enumerate(windows):
        for k in range(d):
            if t >= len(window):
                out[:, d * idx + k] = np.correlate(x[:, k], window, mode='same')
            else:
                # shorter than the window: numpy's 'same' would return len(window) values and
                # the package raises; the zero-padded definition is kept for T < 3
                h = len(window) // 2
                out[:, d * idx + k] = np.correlate(np.pad(x[:, k], (h, h)), window,
                                                   mode='valid')
    return out


def trim_zeros_frames(x, eps=1e-7):
    """Drop trailing frames whose absolute sum is below eps (trim='b')."""
    s = np.sum(np.abs(x), axis=1)
    s[s < eps] = 0.0
    end = len(np.trim_zeros(s, trim='b')) - len(x)
    return x if end == 0 else x[:end]


def remove_zeros_frames(x, eps=1e-7):
    s = np.sum(np.abs(x), axis=1)
    s[s < eps] = 0.0
    return x[s > eps]

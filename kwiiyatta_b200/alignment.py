"""Alignment front-end: the reference's ``kwiiyatta/vocoder/align.py`` on the B200.

Same functions, argument meaning and error behaviour as the reference; the one change is
that ``fastdtw`` is our CUDA module (kwiiyatta_b200.fastdtw) and that batched variants
(``dtw_feature_many`` / ``align_even_many``) exist underneath so many pairs share one launch
sequence.  Features are duck-typed: anything with ``fs``, ``resample_mel_cepstrum(fs).data``,
``is_voiced``, ``f0``, ``frame_len`` and array ``__getitem__`` works (kwiiyatta's Feature,
or kwiiyatta_b200.synth.SynthFeature).
"""
import numpy as np

from . import fastdtw as _fastdtw

pad_silence_fn = None  # set to kwiiyatta.pad_silence (or synth.pad_silence) by the integrator


def set_pad_silence(fn):
    """Install the silence padder used by ``align`` / ``align_even`` when ``pad_silence=True``
    (``kwiiyatta.pad_silence``, kwiiyatta/vocoder/feature.py:19-41 -- feature-producer code)."""
    global pad_silence_fn
    pad_silence_fn = fn


def binalize(x, threshold, ceil, floor=0, out=None):
    """kwiiyatta/vocoder/align.py:10-17."""
    if out is None:
        out = np.full_like(x, floor)
    else:
        out[:] = floor
    out[x >= threshold] = ceil
    return out


def make_feature(f, fs, vuv='voiced', vuv_weight=9.0,
                 power='binalize', power_weight=9.4,
                 power_pivot='max', power_threshold=1.636):
    """kwiiyatta/vocoder/align.py:20-58 -> (T, 2 + order) float64."""
    data = f.resample_mel_cepstrum(fs).data
    data_power = data[:, 0]
    feature = np.hstack((np.zeros((len(data), 2)), data[:, 1:]))

    if power == 'binalize':
        if power_pivot == 'max':
            threshold = data_power.max() - power_threshold
        elif power_pivot == 'median':
            threshold = np.median(data_power) - power_threshold
        elif power_pivot == 'min':
            threshold = data_power.min() + power_threshold
        elif power_pivot == 'fix':
            threshold = power_threshold
        else:
            raise ValueError(f'Unknown power_pivot parameter: {power_pivot!r}')
        binalize(data_power, threshold, power_weight, out=feature[:, 0])
    elif power == 'raw':
        feature[:, 0] = data_power
    elif power is None:
        pass
    else:
        raise ValueError(f'Unknown power parameter: {power!r}')

    if vuv == 'voiced':
        feature[:, 1][f.is_voiced] = vuv_weight
    elif vuv == 'f0':
        feature[:, 1][f.f0 > 0] = vuv_weight
    elif vuv is None:
        pass
    else:
        raise ValueError(f'Unknown vuv parameter: {vuv!r}')
    return feature


def _strict_filter(path, x_feature, y_feature, vuv, power):
    """kwiiyatta/vocoder/align.py:73-94, vectorised.  Keeps the first and last point and the
    interior points whose binary flags agree; reproduces the :78 quirk (x's V/UV column is
    tested against y's power column)."""
    path = np.asarray(path, dtype=np.int64).reshape((-1, 2))
    if len(path) <= 1:
        return np.concatenate((path, path))
    inner = path[1:-1]
    keep = np.ones(len(inner), dtype=bool)
    if power == 'binalize':
        keep &= ~((x_feature[inner[:, 0], 0] > 0) ^ (y_feature[inner[:, 1], 0] > 0))
    if vuv is not None:
        keep &= ~((x_feature[inner[:, 0], 1] > 0) ^ (y_feature[inner[:, 1], 0] > 0))
    return np.concatenate((path[:1], inner[keep], path[-1:]))


def dtw_feature_many(pairs, vuv='voiced', power='binalize', strict=True, radius=32, **kwargs):
    """Batched ``dtw_feature``: ``pairs`` is a sequence of (x, y) features; returns a list of
    ``(dist, path ndarray (L, 2))``."""
    kwargs['vuv'] = vuv
    kwargs['power'] = power
    feats = []
    for x, y in pairs:
        fs = min(x.fs, y.fs)
        feats.append((make_feature(x, fs, **kwargs), make_feature(y, fs, **kwargs)))
    results = _fastdtw.fastdtw_batch(feats, radius=radius, dist=2)
    out = []
    for (xf, yf), (dist, path) in zip(feats, results):
        if strict:
            path = _strict_filter(path, xf, yf, vuv, power)
        else:
            path = np.asarray(path, dtype=np.int64).reshape((-1, 2))
        out.append((dist, path))
    return out


def dtw_feature(x, y, vuv='voiced', power='binalize', strict=True, radius=32, **kwargs):
    """kwiiyatta/vocoder/align.py:61-96."""
    return dtw_feature_many([(x, y)], vuv=vuv, power=power, strict=strict, radius=radius,
                            **kwargs)[0]


def project_path_iter(path, trim=True, trim_len=1):
    """kwiiyatta/vocoder/align.py:99-120."""
    prev_x = prev_y = -1
    if trim:
        prev_y += trim_len
    len_y = path[-1][1] + 1
    if trim:
        len_y -= trim_len
    for x, y in path:
        if y <= prev_y:
            continue
        elif y - prev_y > 1:
            y = min(y, len_y-1)
            diff_x = x - prev_x
            diff_y = y - prev_y
            for i in range(diff_y):
                yield prev_x + diff_x * i // (diff_y-1)
        elif y >= len_y:
            break
        else:
            yield x
        prev_x = x
        prev_y = y


def _pad(feature, pad_len):
    if pad_silence_fn is None:
        raise RuntimeError('kwiiyatta_b200.align.pad_silence_fn is not set: assign '
                           'kwiiyatta.pad_silence (or kwiiyatta_b200.synth.pad_silence)')
    return pad_silence_fn(feature, pad_len)


def align(feature, target, vuv='f0', strict=False, pad_silence=True, pad_len=100, **kwargs):
    """kwiiyatta/vocoder/align.py:123-131: the source warped onto the target's time axis."""
    if pad_silence:
        feature = _pad(feature, pad_len)
        target = _pad(target, pad_len)
    _, path = dtw_feature(feature, target, vuv=vuv, strict=strict, **kwargs)
    return feature[list(project_path_iter(path.tolist(), trim=pad_silence, trim_len=pad_len))]


def _trim_even(path, a_len, b_len, pad_len):
    """kwiiyatta/vocoder/align.py:139-145 (np.argmax of an all-False mask is 0)."""
    path = np.array(path).T
    begin = np.argmax(np.logical_and(path[0] >= pad_len, path[1] >= pad_len))
    end = np.argmax(np.logical_and(path[0] >= a_len - pad_len, path[1] >= b_len - pad_len))
    return path[:, begin:end]


def align_even_many(pairs, pad_silence=True, pad_len=100, **kwargs):
    """Batched ``align_even`` over a sequence of (a, b) features."""
    if pad_silence:
        pairs = [(_pad(a, pad_len), _pad(b, pad_len)) for a, b in pairs]
    results = dtw_feature_many(pairs, **kwargs)
    out = []
    for (a, b), (_, path) in zip(pairs, results):
        path = np.array(path).T
        if pad_silence:
            path = _trim_even(path.T, a.frame_len, b.frame_len, pad_len)
        out.append((a[path[0]], b[path[1]]))
    return out


def align_even(a, b, pad_silence=True, pad_len=100, **kwargs):
    """kwiiyatta/vocoder/align.py:134-146."""
    return align_even_many([(a, b)], pad_silence=pad_silence, pad_len=pad_len, **kwargs)[0]

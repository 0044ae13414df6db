"""Alignment front-end on the B200: what ``kwiiyatta/vocoder/align.py`` computes, batched.

Every function keeps the reference's name, arguments, defaults, return types and error texts
(kwiiyatta/vocoder/align.py:20-146) so callers switch module and nothing else, but the work is
organised for a device that wants many pairs at once: the ``*_many`` functions take a list of
pairs, build all DTW feature matrices, run ONE batched FastDTW (kwiiyatta_b200.fastdtw ->
kw_dtw_batch) and post-process every path with array operations; the single-pair functions of
the reference are the batch-of-one case.  tests/test_reference_seam.py runs the reference's own
functions beside these on the same inputs.

Features are duck-typed: ``fs``, ``frame_len``, ``f0``, ``is_voiced``,
``resample_mel_cepstrum(fs).data`` and index-array ``__getitem__`` (kwiiyatta's Feature, or
kwiiyatta_b200.synth.SynthFeature).
"""
import numpy as np

from . import fastdtw as _fastdtw
from . import hooks

# where the power-flag threshold sits relative to the utterance's c0 track
# (kwiiyatta/vocoder/align.py:28-38); 'fix' takes the threshold as given
_POWER_PIVOTS = {
    'max': lambda c0, t: c0.max() - t,
    'median': lambda c0, t: np.median(c0) - t,
    'min': lambda c0, t: c0.min() + t,
    'fix': lambda c0, t: t,
}


def set_pad_silence(fn):
    """Bind the silence padder used when ``pad_silence=True`` (see kwiiyatta_b200.hooks)."""
    hooks.bind(pad_silence=fn)


def binalize(x, threshold, ceil, floor=0, out=None):
    """Two-level quantisation: ``ceil`` where ``x >= threshold``, else ``floor``
    (kwiiyatta/vocoder/align.py:10-17)."""
    levels = np.where(np.asarray(x) >= threshold, ceil, floor).astype(np.asarray(x).dtype)
    if out is None:
        return levels
    out[...] = levels
    return out


def make_feature(f, fs, vuv='voiced', vuv_weight=9.0, power='binalize', power_weight=9.4,
                 power_pivot='max', power_threshold=1.636):
    """The (T, 2 + order) float64 matrix FastDTW compares (kwiiyatta/vocoder/align.py:20-58):
    column 0 the power flag (or raw c0, or zero), column 1 the voicing flag, then mcep[1:]."""
    mcep = f.resample_mel_cepstrum(fs).data
    c0 = mcep[:, 0]
    out = np.zeros((len(mcep), mcep.shape[1] + 1))
    out[:, 2:] = mcep[:, 1:]
    if power == 'binalize':
        if power_pivot not in _POWER_PIVOTS:
            raise ValueError(f'Unknown power_pivot parameter: {power_pivot!r}')
        out[:, 0] = np.where(c0 >= _POWER_PIVOTS[power_pivot](c0, power_threshold),
                             power_weight, 0.0)
    elif power == 'raw':
        out[:, 0] = c0
    elif power is not None:
        raise ValueError(f'Unknown power parameter: {power!r}')
    if vuv == 'voiced':
        out[np.asarray(f.is_voiced, dtype=bool), 1] = vuv_weight
    elif vuv == 'f0':
        out[np.asarray(f.f0) > 0, 1] = vuv_weight
    elif vuv is not None:
        raise ValueError(f'Unknown vuv parameter: {vuv!r}')
    return out


def _strict_filter(path, x_feature, y_feature, vuv, power):
    """kwiiyatta/vocoder/align.py:73-94 as array operations.  Keeps the first and last point and
    the interior points whose binary flags agree; reproduces the :78 quirk (x's V/UV column is
    tested against y's POWER column)."""
    path = np.asarray(path, dtype=np.int64).reshape((-1, 2))
    if len(path) <= 1:
        return np.concatenate((path, path))
    inner = path[1:-1]
    keep = np.ones(len(inner), dtype=bool)
    y_power_on = y_feature[inner[:, 1], 0] > 0
    if power == 'binalize':
        keep &= (x_feature[inner[:, 0], 0] > 0) == y_power_on
    if vuv is not None:
        keep &= (x_feature[inner[:, 0], 1] > 0) == y_power_on
    return np.concatenate((path[:1], inner[keep], path[-1:]))


def dtw_feature_many(pairs, vuv='voiced', power='binalize', strict=True, radius=32, **kwargs):
    """``dtw_feature`` for a sequence of (x, y) features with one batched FastDTW underneath.
    Returns a list of ``(dist, path ndarray (L, 2))``."""
    options = dict(kwargs, vuv=vuv, power=power)
    matrices = []
    for x, y in pairs:
        fs = min(x.fs, y.fs)
        matrices.append((make_feature(x, fs, **options), make_feature(y, fs, **options)))
    aligned = _fastdtw.fastdtw_batch(matrices, radius=radius, dist=2)
    results = []
    for (xm, ym), (dist, path) in zip(matrices, aligned):
        path = np.asarray(path, dtype=np.int64).reshape((-1, 2))
        results.append((dist, _strict_filter(path, xm, ym, vuv, power) if strict else path))
    return results


def dtw_feature(x, y, vuv='voiced', power='binalize', strict=True, radius=32, **kwargs):
    """kwiiyatta/vocoder/align.py:61-96."""
    return dtw_feature_many([(x, y)], vuv=vuv, power=power, strict=strict, radius=radius,
                            **kwargs)[0]


def project_path(path, trim=True, trim_len=1):
    """One x index per y frame of the (trimmed) target, as an int64 array: the values
    ``project_path_iter`` of the reference yields (kwiiyatta/vocoder/align.py:99-120), computed
    without a Python loop over the path.

    The reference walks the path keeping the last accepted point (px, py); a point is accepted
    when its y exceeds py.  A jump of more than one y (only possible after the strict filter or
    at the trimmed start) is bridged by ``px + (x - px) * i // (dy - 1)`` for i < dy; the walk
    ends at the first accepted point with y >= len_y (bridged up to len_y - 1 when it is a
    jump).  Start state (-1, trim_len - 1), so a leading jump can yield -1 (Python's "last
    frame"), and a bridge of dy == 1 after clamping divides by zero exactly as the reference's
    does."""
    path = np.asarray(path, dtype=np.int64).reshape((-1, 2))
    first_y = trim_len - 1 if trim else -1
    len_y = int(path[-1, 1]) + 1 - (trim_len if trim else 0)
    xs, ys = path[:, 0], path[:, 1]
    # accepted points: y above everything seen before (and above the start state)
    before = np.maximum.accumulate(np.concatenate(([first_y], ys[:-1])))
    taken = ys > before
    xs, ys = xs[taken], ys[taken]
    if len(ys) == 0:
        return np.zeros(0, dtype=np.int64)
    px = np.concatenate(([-1], xs[:-1]))
    py = np.concatenate(([first_y], ys[:-1]))
    # the walk stops at the first accepted point at or beyond len_y
    over = np.flatnonzero(ys >= len_y)
    stop = int(over[0]) if len(over) else len(ys)
    jump_stop = stop < len(ys) and ys[stop] - py[stop] > 1
    upto = stop + 1 if jump_stop else stop
    xs, ys, px, py = xs[:upto], ys[:upto].copy(), px[:upto], py[:upto]
    if jump_stop:
        ys[-1] = min(ys[-1], len_y - 1)
    dy = ys - py
    dx = xs - px
    if jump_stop and dy[-1] == 1:
        raise ZeroDivisionError('integer division or modulo by zero')
    count = np.where(dy > 1, dy, 1)
    if jump_stop and dy[-1] <= 0:
        count[-1] = 0
    owner = np.repeat(np.arange(len(ys)), count)
    step = np.arange(len(owner)) - np.repeat(np.cumsum(count) - count, count)
    bridged = px[owner] + dx[owner] * step // np.maximum(dy[owner] - 1, 1)
    return np.where(dy[owner] > 1, bridged, xs[owner])


def project_path_iter(path, trim=True, trim_len=1):
    """kwiiyatta/vocoder/align.py:99-120 (an iterator, as there)."""
    return iter(project_path(path, trim, trim_len).tolist())


def _pad(feature, pad_len):
    return hooks.get('pad_silence')(feature, pad_len)


def align_many(pairs, vuv='f0', strict=False, pad_silence=True, pad_len=100, **kwargs):
    """``align`` for a sequence of (feature, target) pairs: each source warped onto its target's
    time axis, one batched FastDTW."""
    if pad_silence:
        pairs = [(_pad(a, pad_len), _pad(b, pad_len)) for a, b in pairs]
    results = dtw_feature_many(pairs, vuv=vuv, strict=strict, **kwargs)
    return [a[project_path(path, trim=pad_silence, trim_len=pad_len)]
            for (a, _), (_, path) in zip(pairs, results)]


def align(feature, target, vuv='f0', strict=False, pad_silence=True, pad_len=100, **kwargs):
    """kwiiyatta/vocoder/align.py:123-131: the source warped onto the target's time axis."""
    return align_many([(feature, target)], vuv=vuv, strict=strict, pad_silence=pad_silence,
                      pad_len=pad_len, **kwargs)[0]


def _trim_even(path, a_len, b_len, pad_len):
    """(2, L') slice of the path between the first point inside both un-padded regions and the
    first point past both (kwiiyatta/vocoder/align.py:139-145; ``np.argmax`` of an all-False mask
    is 0, which the reference relies on)."""
    path = np.asarray(path).reshape((-1, 2)).T
    inside = (path[0] >= pad_len) & (path[1] >= pad_len)
    past = (path[0] >= a_len - pad_len) & (path[1] >= b_len - pad_len)
    return path[:, np.argmax(inside):np.argmax(past)]


def align_even_many(pairs, pad_silence=True, pad_len=100, **kwargs):
    """``align_even`` for a sequence of (a, b) features, one batched FastDTW."""
    if pad_silence:
        pairs = [(_pad(a, pad_len), _pad(b, pad_len)) for a, b in pairs]
    out = []
    for (a, b), (_, path) in zip(pairs, dtw_feature_many(pairs, **kwargs)):
        idx = _trim_even(path, a.frame_len, b.frame_len, pad_len) if pad_silence \
            else np.asarray(path).reshape((-1, 2)).T
        out.append((a[idx[0]], b[idx[1]]))
    return out


def align_even(a, b, pad_silence=True, pad_len=100, **kwargs):
    """kwiiyatta/vocoder/align.py:134-146."""
    return align_even_many([(a, b)], pad_silence=pad_silence, pad_len=pad_len, **kwargs)[0]

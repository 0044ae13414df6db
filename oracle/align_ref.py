"""Restatement of the reference's alignment glue (TEST INFRASTRUCTURE - see oracle/__init__.py).

Follows kwiiyatta/vocoder/align.py:10-146 on plain arrays: the Feature container
(pyworld / pysptk) is the feature producer and stays outside the hot path, so a
"feature" here is a dict-like with ``mcep`` (T, order+1) float64, ``f0`` (T,),
``is_voiced`` (T,) bool.
"""
import numpy as np

from . import dtw_c, fastdtw_ref


def binalize(x, threshold, ceil, floor=0.0):
    """kwiiyatta/vocoder/align.py:10-17."""
    out = np.full_like(x, floor)
    out[x >= threshold] = ceil
    return out


def make_feature(mcep, f0=None, is_voiced=None, vuv='voiced', vuv_weight=9.0,
                 power='binalize', power_weight=9.4, power_pivot='max',
                 power_threshold=1.636):
    """kwiiyatta/vocoder/align.py:20-58 (resampling at :23 is the producer's job)."""
    data = np.asarray(mcep, dtype=np.float64)
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
        feature[:, 0] = binalize(data_power, threshold, power_weight)
    elif power == 'raw':
        feature[:, 0] = data_power
    elif power is None:
        pass
    else:
        raise ValueError(f'Unknown power parameter: {power!r}')
    if vuv == 'voiced':
        feature[:, 1][np.asarray(is_voiced, dtype=bool)] = vuv_weight
    elif vuv == 'f0':
        feature[:, 1][np.asarray(f0) > 0] = vuv_weight
    elif vuv is None:
        pass
    else:
        raise ValueError(f'Unknown vuv parameter: {vuv!r}')
    return feature


def strict_filter(path, x_feature, y_feature, vuv='voiced', power='binalize'):
    """kwiiyatta/vocoder/align.py:73-94, including the :78 quirk (x's V/UV column is
    compared with y's POWER column)."""
    def check(x, y):
        if power == 'binalize':
            if (x_feature[x, 0] > 0) ^ (y_feature[y, 0] > 0):
                return False
        if vuv is not None:
            if (x_feature[x, 1] > 0) ^ (y_feature[y, 0] > 0):
                return False
        return True
    path = [tuple(p) for p in np.asarray(path).tolist()]
    if len(path) < 2:
        # the reference chains path[0], path[1:-1], path[-1]: a 1-point path is duplicated
        kept = [path[0], path[-1]]
    else:
        kept = [path[0]] + [p for p in path[1:-1] if check(*p)] + [path[-1]]
    return np.array(kept, dtype=np.int64).reshape((-1, 2))


def dtw_feature(x_feature, y_feature, vuv='voiced', power='binalize', strict=True,
                radius=32, backend='c'):
    """kwiiyatta/vocoder/align.py:61-96 on ready-made (T, 26) feature matrices."""
    if backend == 'c':
        dist, path = dtw_c.fastdtw(x_feature, y_feature, radius=radius, dist=2)
    else:
        dist, path = fastdtw_ref.fastdtw(x_feature, y_feature, radius=radius, dist=2)
    if strict:
        path = strict_filter(path, x_feature, y_feature, vuv=vuv, power=power)
    else:
        path = np.array(path, dtype=np.int64).reshape((-1, 2))
    return dist, path


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
            y = min(y, len_y - 1)
            diff_x = x - prev_x
            diff_y = y - prev_y
            for i in range(diff_y):
                yield prev_x + diff_x * i // (diff_y - 1)
        elif y >= len_y:
            break
        else:
            yield x
        prev_x = x
        prev_y = y


def trim_even_path(path, len_a, len_b, pad_len=100):
    """kwiiyatta/vocoder/align.py:139-145: cut the (L, 2) path to the un-padded region.
    ``np.argmax`` of an all-False mask is 0, as in the reference."""
    p = np.asarray(path).T
    begin = np.argmax(np.logical_and(p[0] >= pad_len, p[1] >= pad_len))
    end = np.argmax(np.logical_and(p[0] >= len_a - pad_len, p[1] >= len_b - pad_len))
    return p[:, begin:end]

"""Batched joint-frame assembly: ``make_dataset_to_array`` (kwiiyatta/converter/dataset.py:61-77)
with one batched DTW underneath instead of a per-key Python loop."""
import numpy as np

from . import _lib
from . import alignment as _align
from .delta import delta_features_device


def remove_zeros_frames(x, eps=1e-7):
    """nnmnkwii.preprocessing.remove_zeros_frames as called at kwiiyatta/converter/dataset.py:70."""
    s = np.sum(np.abs(x), axis=1)
    s[s < eps] = 0.0
    return x[s > eps]


def make_dataset_to_array(dataset, keys=None):
    """Reference semantics: per key ``np.hstack`` of the tuple, drop zero frames, append."""
    if keys is None:
        keys = sorted(dataset.keys())
    chunks = []
    for key in keys:
        d = dataset[key]
        if isinstance(d, tuple):
            d = np.hstack(d)
        chunks.append(remove_zeros_frames(d))
    if not chunks:
        return None
    return np.concatenate(chunks)


def joint_array_from_pairs(pairs, use_delta=True, pad_silence=True, pad_len=100, **align_kwargs):
    """The whole training-array pipeline for a list of (source, target) features:
    align_even (batched DTW) -> mcep without c0 (kwiiyatta/converter/mcep.py:33) -> delta
    features on the device (kwiiyatta/converter/delta.py:30) -> hstack -> remove zero frames.
    Returns the (N, 2*3*order) float64 array GaussianMixture.fit receives."""
    torch = _lib.require_cuda()
    aligned = _align.align_even_many(pairs, pad_silence=pad_silence, pad_len=pad_len,
                                     **align_kwargs)
    src = [a.mel_cepstrum.data[:, 1:] for a, _ in aligned]
    tgt = [b.mel_cepstrum.data[:, 1:] for _, b in aligned]
    lens = np.array([len(s) for s in src], dtype=np.int64)
    off = np.concatenate(([0], np.cumsum(lens)))
    if off[-1] == 0:
        return np.zeros((0, 0))
    parts = []
    off_dev = torch.from_numpy(off).cuda()
    for side in (src, tgt):
        x = torch.from_numpy(np.ascontiguousarray(np.concatenate(side))).cuda()
        parts.append(delta_features_device(x, off_dev, len(lens)) if use_delta else x)
    joint = torch.cat(parts, dim=1).cpu().numpy()
    return remove_zeros_frames(joint)

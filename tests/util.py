"""Shared helpers for the parity tests (oracle side)."""
import numpy as np

from kwiiyatta_b200 import synth
from oracle import align_ref, delta_ref, dtw_c


def oracle_joint_array(n_pairs, radius=32, first=0):
    """Config-1 pipeline on the CPU oracle: pad -> make_feature -> FastDTW -> strict filter ->
    trim -> gather mcep[:, 1:] -> delta -> hstack -> remove zero frames."""
    chunks, paths = [], []
    for i in range(first, first + n_pairs):
        a, b = synth.make_padded_pair(i)
        xf = align_ref.make_feature(a.mel_cepstrum.data, a.f0, a.is_voiced)
        yf = align_ref.make_feature(b.mel_cepstrum.data, b.f0, b.is_voiced)
        _, path = dtw_c.fastdtw(xf, yf, radius=radius, dist=2)
        path = align_ref.strict_filter(path, xf, yf)
        p = align_ref.trim_even_path(path, a.frame_len, b.frame_len, synth.PAD_LEN)
        paths.append(p)
        src = delta_ref.delta_features(a.mel_cepstrum.data[p[0]][:, 1:])
        tgt = delta_ref.delta_features(b.mel_cepstrum.data[p[1]][:, 1:])
        chunks.append(delta_ref.remove_zeros_frames(np.hstack((src, tgt))))
    return np.concatenate(chunks), paths


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

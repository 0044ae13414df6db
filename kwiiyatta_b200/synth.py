"""Synthetic ATR503-shaped parallel corpus.

There is no network and no pyworld/pysptk in the image, so the feature producer
(WORLD analysis + sp2mc, kwiiyatta/vocoder/world.py:33-59, vocoder/mcep.py:68-71) is
replaced by a seeded generator of mel-cepstrum-like tracks.  Every consumer (tests, the
oracle, bench.py) calls this module so CPU and GPU arms see identical float64 inputs.

One pair = ``np.random.default_rng(seed0 + pair_idx)``: a source track of
``Tx ~ U{450..750}`` frames (order-24 mcep incl. c0, 5 ms frames) and a target that is a
monotone time-warped copy pushed through a fixed near-identity affine "speaker" map plus
N(0, 0.05^2) noise on every coefficient (the noise makes DTW ties measure-zero).
"""
import numpy as np
import scipy.signal

SEED0 = 1234
ORDER = 24
PAD_LEN = 100
FS = 16000
FRAME_PERIOD = 5
SPECTRUM_LEN = 33


class SynthFeature:
    """The slice of kwiiyatta's Feature interface the alignment path touches
    (kwiiyatta/vocoder/abc/feature.py:11-194): ``fs``, ``frame_period``, ``f0``,
    ``is_voiced``, ``mel_cepstrum.data`` / ``resample_mel_cepstrum(fs).data``,
    ``frame_len`` and fancy ``__getitem__``."""

    class _Mcep:
        def __init__(self, data, fs, frame_period):
            self.data = data
            self.fs = fs
            self.frame_period = frame_period
            self.order = data.shape[1] - 1

    def __init__(self, mcep, f0, is_voiced, fs=FS, frame_period=FRAME_PERIOD):
        self.fs = fs
        self.frame_period = frame_period
        self.f0 = np.asarray(f0, dtype=np.float64)
        self.is_voiced = np.asarray(is_voiced, dtype=bool)
        self.mel_cepstrum = SynthFeature._Mcep(np.asarray(mcep, dtype=np.float64), fs,
                                               frame_period)

    @property
    def mel_cepstrum_order(self):
        return self.mel_cepstrum.order

    @property
    def spectrum_envelope(self):
        """Stand-in spectrum: the mel-cepstrum zero-padded to SPECTRUM_LEN columns (only its
        all-zero trailing frames matter to the path: TrimmedDataset,
        kwiiyatta/converter/dataset.py:49-52)."""
        out = np.zeros((len(self.f0), SPECTRUM_LEN))
        out[:, :self.mel_cepstrum.data.shape[1]] = self.mel_cepstrum.data
        return out

    @property
    def frame_len(self):
        return len(self.f0)

    def __len__(self):
        return len(self.f0)

    def resample_mel_cepstrum(self, fs):
        if fs != self.fs:
            raise ValueError('SynthFeature cannot resample (feature producer is out of scope)')
        return self.mel_cepstrum

    def __getitem__(self, key):
        if isinstance(key, int):
            raise TypeError('SynthFeature supports slice / index-array access only')
        return SynthFeature(self.mel_cepstrum.data[key], self.f0[key], self.is_voiced[key],
                            self.fs, self.frame_period)


def feature(f):
    """Stand-in for ``kwiiyatta.feature(Feature)``: an independent copy."""
    return SynthFeature(f.mel_cepstrum.data.copy(), f.f0.copy(), f.is_voiced.copy(), f.fs,
                        f.frame_period)


def resample(mel_cepstrum, fs):
    """Stand-in for ``kwiiyatta.resample``: synthetic features exist at one rate only."""
    if fs != mel_cepstrum.fs:
        raise ValueError('synthetic mel-cepstra cannot be resampled (feature producer is out '
                         'of scope)')
    import copy
    return copy.copy(mel_cepstrum)


def pad_silence(feature, frame_len=PAD_LEN, rng=None):
    """Stand-in for kwiiyatta.pad_silence (kwiiyatta/vocoder/feature.py:19-41): ``frame_len``
    unvoiced low-power frames either side, mcep ~ N(0, 1e-3^2), c0 = -10."""
    if rng is None:
        rng = np.random.default_rng(0)
    dim = feature.mel_cepstrum.data.shape[1]

    def sil():
        m = rng.normal(0.0, 1e-3, (frame_len, dim))
        m[:, 0] += -10.0
        return m
    mcep = np.concatenate((sil(), feature.mel_cepstrum.data, sil()))
    f0 = np.concatenate((np.zeros(frame_len), feature.f0, np.zeros(frame_len)))
    voiced = np.concatenate((np.zeros(frame_len, bool), feature.is_voiced,
                             np.zeros(frame_len, bool)))
    return SynthFeature(mcep, f0, voiced, feature.fs, feature.frame_period)


def _smooth_noise(rng, t, width, cols=None):
    shape = (t + 2 * width,) if cols is None else (t + 2 * width, cols)
    z = rng.standard_normal(shape)
    z = scipy.signal.lfilter([1.0], [1.0, -0.9], z, axis=0)
    k = np.ones(width) / width
    if cols is None:
        s = np.convolve(z, k, mode='same')[width:width + t]
    else:
        s = scipy.signal.lfilter(k, [1.0], z, axis=0)[2 * width - 1:2 * width - 1 + t]
    return (s - s.mean(axis=0)) / (s.std(axis=0) + 1e-12)


def _speaker_map(seed0, order):
    rng = np.random.default_rng(seed0 - 1)
    a = np.eye(order) + 0.05 * rng.standard_normal((order, order))
    bias = 0.1 * rng.standard_normal(order) / (1.0 + np.arange(order))
    return a, bias


def make_pair(pair_idx, seed0=SEED0, order=ORDER, t_range=(450, 750), noise=0.05,
              length=None):
    """Returns (source, target) SynthFeatures, un-padded.  ``length`` fixes Tx = Ty (no
    warp-length change) for the long singing-transfer configuration."""
    rng = np.random.default_rng(seed0 + pair_idx)
    tx = int(rng.integers(t_range[0], t_range[1] + 1)) if length is None else int(length)
    sigma = 1.0 / (1.0 + np.arange(order))
    static = _smooth_noise(rng, tx, 9, order) * sigma + rng.normal(0.0, 1.0, order) * sigma
    energy = _smooth_noise(rng, tx, 15)
    gate = _smooth_noise(rng, tx, 31)
    silent = gate < -1.0
    c0 = 2.0 + 1.5 * energy - 6.0 * silent
    voiced = (_smooth_noise(rng, tx, 21) > -0.3) & ~silent
    f0 = np.where(voiced, 120.0 + 20.0 * energy, 0.0)

    ty = tx if length is not None else int(round(tx * rng.uniform(0.8, 1.25)))
    n_seg = 8
    slopes = rng.uniform(0.5, 2.0, n_seg)
    knots_y = np.linspace(0.0, ty - 1.0, n_seg + 1)
    knots_x = np.concatenate(([0.0], np.cumsum(slopes)))
    knots_x *= (tx - 1.0) / knots_x[-1]
    tau = np.interp(np.arange(ty), knots_y, knots_x)
    i0 = np.clip(np.floor(tau).astype(np.int64), 0, tx - 2)
    frac = (tau - i0)[:, None]
    a, bias = _speaker_map(seed0, order)
    warped = static[i0] * (1 - frac) + static[i0 + 1] * frac
    tgt_static = warped @ a + bias + rng.normal(0.0, noise, (ty, order))
    tgt_c0 = (c0[i0] * (1 - frac[:, 0]) + c0[i0 + 1] * frac[:, 0]
              + rng.normal(0.0, noise, ty))
    near = np.clip(np.rint(tau).astype(np.int64), 0, tx - 1)
    tgt_voiced = voiced[near]
    tgt_f0 = np.where(tgt_voiced, f0[near] * 1.8, 0.0)

    src = SynthFeature(np.hstack((c0[:, None], static)), f0, voiced)
    tgt = SynthFeature(np.hstack((tgt_c0[:, None], tgt_static)), tgt_f0, tgt_voiced)
    return src, tgt


def make_padded_pair(pair_idx, seed0=SEED0, pad_len=PAD_LEN, **kw):
    """The pair as align_even sees it after pad_silence (kwiiyatta/vocoder/align.py:135-137)."""
    src, tgt = make_pair(pair_idx, seed0=seed0, **kw)
    rng = np.random.default_rng((seed0 + pair_idx) * 7919 + 1)
    return pad_silence(src, pad_len, rng), pad_silence(tgt, pad_len, rng)


def make_joint_gmm(n_components, dim_half=72, seed=0, static_dim=24):
    """A random but well-conditioned joint GMM (weights, means (K, 2*dim_half), covariances)
    for the bulk-conversion configuration: per-component SPD covariance with a decaying
    spectrum and source/target correlation, mcep-like scales per static/delta block."""
    rng = np.random.default_rng(seed)
    d = 2 * dim_half
    blocks = dim_half // static_dim
    scale_half = np.concatenate([(1.0 / (1.0 + np.arange(static_dim))) * (0.5 ** b)
                                 for b in range(blocks)])
    scale = np.concatenate((scale_half, scale_half))
    weights = rng.uniform(0.5, 1.5, n_components)
    weights /= weights.sum()
    means = rng.standard_normal((n_components, d)) * scale
    cov = np.empty((n_components, d, d))
    for k in range(n_components):
        g = rng.standard_normal((d, d)) / np.sqrt(d)
        c = 0.35 * np.eye(d) + 0.25 * (g @ g.T)
        c[:dim_half, dim_half:] += 0.3 * np.eye(dim_half)
        c[dim_half:, :dim_half] += 0.3 * np.eye(dim_half)
        c = 0.5 * (c + c.T)
        cov[k] = c * scale[:, None] * scale[None, :]
    return weights, means, cov


def make_source_utterances(n_utts, frames=600, seed0=SEED0, order=ORDER):
    """Source static mcep (T, order) per utterance for the conversion benchmark."""
    out = []
    for u in range(n_utts):
        src, _ = make_pair(u, seed0=seed0 + 100000, order=order, length=frames)
        out.append(src.mel_cepstrum.data[:, 1:])
    return out

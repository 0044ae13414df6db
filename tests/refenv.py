"""The REAL reference package, imported from /root/reference with its absent third-party
dependencies stubbed (TEST INFRASTRUCTURE; only used where /root/reference exists, i.e. in the
build container -- never on the GPU box).

``import kwiiyatta`` fails in this image for two reasons only: the packages
fastdtw / nnmnkwii / pysptk / pyworld / pyaudio are not installed, and ``np.int`` is gone from
numpy.  With those six names provided, every line of the reference's own alignment, dataset and
converter code runs unmodified:

* ``fastdtw``                      -> a module handed in by the caller (the oracle's FastDTW, or
                                      kwiiyatta_b200.fastdtw to show the drop-in);
* ``nnmnkwii.preprocessing``       -> oracle.delta_ref (delta_features, trim/remove_zeros_frames);
* ``nnmnkwii.baseline.gmm.MLPG``   -> a class handed in by the caller (oracle.mlpg_ref wrapped, or
                                      kwiiyatta_b200.MLPG);
* ``pysptk``                       -> an exactly invertible stand-in for sp2mc / mc2sp: the
                                      "spectrum" of a frame is its mel-cepstrum zero-padded to
                                      the spectrum length, so features keep the synthetic
                                      mel-cepstra bit for bit (the real transforms are feature
                                      production, out of scope);
* ``pyworld``, ``pyaudio``         -> empty modules (analysis / playback are never called).

Synthetic utterances enter as real ``kwiiyatta.feature`` objects whose ``Synthesizer`` is
``SynthSynthesizer`` below; its silence frames reproduce kwiiyatta_b200.synth.pad_silence draw
for draw, so the reference's own ``kwiiyatta.pad_silence`` yields exactly the padded features the
GPU tests build without the reference.
"""
import contextlib
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = '/root/reference'
SPECTRUM_LEN = 33          # >= mcep order + 1; small keeps Feature.__getitem__ cheap


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'kwiiyatta'))


def _pysptk_stub():
    m = types.ModuleType('pysptk')

    def mc2sp(mc, alpha=0.0, fftlen=0):
        width = fftlen // 2 + 1
        out = np.zeros((len(mc), width))
        out[:, :mc.shape[1]] = mc
        return out

    def sp2mc(spectrum, order=24, alpha=0.0):
        return np.array(spectrum[:, :order + 1])

    m.mc2sp, m.sp2mc = mc2sp, sp2mc
    m.__path__ = []                       # a package: kwiiyatta imports pysptk.synthesis
    m.util = types.ModuleType('pysptk.util')
    m.util.mcepalpha = lambda fs: 0.41
    m.synthesis = types.ModuleType('pysptk.synthesis')      # MLSA filter: never called here
    m.synthesis.MLSADF = m.synthesis.Synthesizer = None
    return m


def _nnmnkwii_stub(mlpg_class):
    from oracle import delta_ref
    top = types.ModuleType('nnmnkwii')
    pre = types.ModuleType('nnmnkwii.preprocessing')
    pre.delta_features = delta_ref.delta_features
    pre.trim_zeros_frames = delta_ref.trim_zeros_frames
    pre.remove_zeros_frames = delta_ref.remove_zeros_frames
    base = types.ModuleType('nnmnkwii.baseline')
    gmm = types.ModuleType('nnmnkwii.baseline.gmm')
    gmm.MLPG = mlpg_class
    top.preprocessing, top.baseline, base.gmm = pre, base, gmm
    return {'nnmnkwii': top, 'nnmnkwii.preprocessing': pre, 'nnmnkwii.baseline': base,
            'nnmnkwii.baseline.gmm': gmm}


class OracleMLPG:
    """nnmnkwii.baseline.gmm.MLPG's constructor / transform on the oracle restatement."""

    def __init__(self, gmm, windows=None, swap=False, diff=False):
        assert not swap
        self.gmm, self.windows, self.diff = gmm, windows, diff

    def transform(self, src):
        from oracle import mlpg_ref
        w, m, c = self.gmm.weights_, self.gmm.means_, self.gmm.covariances_
        if len(self.windows) == 1:
            return mlpg_ref.transform_frames_soft(src, w, m, c, diff=self.diff)
        return mlpg_ref.transform(src, w, m, c, diff=self.diff, windows=self.windows)


def oracle_fastdtw_module():
    from oracle import dtw_c
    m = types.ModuleType('fastdtw')

    def fastdtw(x, y, radius=1, dist=None):
        cost, path = dtw_c.fastdtw(x, y, radius=radius, dist=dist)
        return cost, [tuple(p) for p in path.tolist()]

    def dtw(x, y, dist=None):
        cost, path = dtw_c.dtw(x, y, dist=dist)
        return cost, [tuple(p) for p in path.tolist()]

    m.fastdtw, m.dtw = fastdtw, dtw
    return m


@contextlib.contextmanager
def reference(fastdtw_module=None, mlpg_class=OracleMLPG):
    """``with reference() as kwiiyatta:`` -- the reference package, importable for the duration
    of the block; sys.modules / sys.path / numpy are restored afterwards."""
    if not available():
        raise RuntimeError(f'{REFERENCE_ROOT} is not present')
    if fastdtw_module is None:
        fastdtw_module = oracle_fastdtw_module()
    stubs = {'fastdtw': fastdtw_module, 'pysptk': _pysptk_stub(),
             'pyworld': types.ModuleType('pyworld'), 'pyaudio': types.ModuleType('pyaudio')}
    stubs['pysptk.util'] = stubs['pysptk'].util
    stubs['pysptk.synthesis'] = stubs['pysptk'].synthesis
    stubs.update(_nnmnkwii_stub(mlpg_class))
    saved = {name: sys.modules.get(name) for name in stubs}
    had_np_int = hasattr(np, 'int')
    sys.modules.update(stubs)
    if not had_np_int:
        np.int = int                       # removed in numpy 1.24 (kwiiyatta/vocoder/align.py:91)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        yield importlib.import_module('kwiiyatta')
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for name in [n for n in sys.modules if n == 'kwiiyatta' or n.startswith('kwiiyatta.')]:
            del sys.modules[name]
        for name, mod in saved.items():
            if mod is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = mod
        if not had_np_int:
            del np.int


def make_synthesizer(kwiiyatta, rng_holder):
    """A kwiiyatta Synthesizer whose silence matches kwiiyatta_b200.synth.pad_silence:
    ``rng_holder['rng']`` is consumed exactly as synth.pad_silence consumes its rng (one
    (frame_len, order + 1) normal draw per silence spectrum, leading then trailing)."""
    class SynthSynthesizer(kwiiyatta.vocoder.abc.Synthesizer):
        @staticmethod
        def _synthesize(feature):
            raise NotImplementedError('waveform synthesis is outside the hot path')

        @staticmethod
        def fs_spectrum_len(fs):
            return SPECTRUM_LEN

        @staticmethod
        def extract_is_voiced(feature):
            return feature.f0 > 0

        @staticmethod
        def silence_f0(frame_len, fs):
            return np.zeros(frame_len)

        @staticmethod
        def _silence_spectrum_envelope(frame_len, fs, spectrum_len):
            from kwiiyatta_b200 import synth
            m = rng_holder['rng'].normal(0.0, 1e-3, (frame_len, synth.ORDER + 1))
            m[:, 0] += -10.0
            out = np.zeros((frame_len, spectrum_len))
            out[:, :m.shape[1]] = m
            return out

        @staticmethod
        def _silence_aperiodicity(frame_len, fs, spectrum_len):
            return np.full((frame_len, spectrum_len), 0.5)

        @staticmethod
        def _resample_up_spectrum_envelope(feature, fs, new_fs, new_spectrum_len):
            raise NotImplementedError

        @staticmethod
        def _resample_up_aperiodicity(feature, fs, new_fs, new_spectrum_len):
            raise NotImplementedError

    return SynthSynthesizer


def to_reference_feature(kwiiyatta, synth_feature, synthesizer):
    """kwiiyatta_b200.synth.SynthFeature -> a real kwiiyatta Feature with the same mel-cepstrum."""
    mcep = synth_feature.mel_cepstrum.data
    f = kwiiyatta.feature(synth_feature.fs, frame_period=synth_feature.frame_period,
                          mcep_order=mcep.shape[1] - 1, Synthesizer=synthesizer)
    spec = np.zeros((len(mcep), SPECTRUM_LEN))
    spec[:, :mcep.shape[1]] = mcep
    f.f0 = np.array(synth_feature.f0)
    f.spectrum_envelope = spec
    f.aperiodicity = np.full((len(mcep), SPECTRUM_LEN), 0.5)
    return f

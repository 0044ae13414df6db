"""Device-resident training-array assembly (kwiiyatta_b200.assemble: kw_dtw_features,
kw_dtw_batch, kw_path_select, kw_joint_frames) against the oracle chain and the host path.

Reference: kwiiyatta/vocoder/align.py:20-96,134-146, kwiiyatta/converter/mcep.py:33,
converter/delta.py:30, converter/dataset.py:61-77."""
import warnings

import numpy as np
import pytest

import kwiiyatta_b200 as kw
from kwiiyatta_b200 import assemble, synth
from oracle import align_ref, delta_ref, dtw_c
from util import oracle_joint_array

pytestmark = pytest.mark.gpu


@pytest.fixture()
def padded_hook():
    kw.hooks.bind(pad_silence=lambda f, n: f, feature=synth.feature, resample=synth.resample)
    yield
    kw.hooks.bind(pad_silence=None, feature=None, resample=None)


def _oracle_chain(pairs, use_delta=True, trim=True, pad_len=synth.PAD_LEN, radius=32,
                  strict=True, **feat_kw):
    chunks = []
    for a, b in pairs:
        xf = align_ref.make_feature(a.mel_cepstrum.data, a.f0, a.is_voiced, **feat_kw)
        yf = align_ref.make_feature(b.mel_cepstrum.data, b.f0, b.is_voiced, **feat_kw)
        _, path = dtw_c.fastdtw(xf, yf, radius=radius, dist=2)
        if strict:
            path = align_ref.strict_filter(path, xf, yf, vuv=feat_kw.get('vuv', 'voiced'),
                                           power=feat_kw.get('power', 'binalize'))
        p = align_ref.trim_even_path(path, a.frame_len, b.frame_len, pad_len) if trim \
            else np.asarray(path).T
        src, tgt = a.mel_cepstrum.data[p[0]][:, 1:], b.mel_cepstrum.data[p[1]][:, 1:]
        if use_delta:
            src, tgt = delta_ref.delta_features(src), delta_ref.delta_features(tgt)
        chunks.append(delta_ref.remove_zeros_frames(np.hstack((src, tgt))))
    return np.concatenate(chunks)


def test_config1_and_2_arrays_bit_identical(cuda, padded_hook):
    """configs[0] (10 pairs) and configs[1] (503 pairs): the device-assembled (N, 144) matrix
    equals the oracle chain's bit for bit and never leaves the device."""
    for n in (10, 503):
        pairs = [synth.make_padded_pair(i) for i in range(n)]
        expected, _ = oracle_joint_array(n)
        x = assemble.joint_frames_device(pairs, pad_silence=True, pad_len=synth.PAD_LEN)
        assert x.is_cuda and tuple(x.shape) == expected.shape
        assert np.array_equal(x.cpu().numpy(), expected)
    # the host path (batched DTW, numpy post-processing) gives the same array
    host = kw.joint_array_from_pairs(pairs[:40], pad_silence=True, pad_len=synth.PAD_LEN)
    dev = kw.joint_array_from_pairs(pairs[:40], pad_silence=True, pad_len=synth.PAD_LEN,
                                    device_resident=True)
    assert np.array_equal(host, dev.cpu().numpy())


@pytest.mark.parametrize('options', [
    {'strict': False}, {'vuv': 'f0'}, {'vuv': None, 'power': 'raw', 'strict': False},
    {'power': None}, {'radius': 1}, {'power_pivot': 'median'}, {'power_pivot': 'fix',
                                                                'power_threshold': 1.0},
    {'use_delta': False}, {'pad_silence': False, 'radius': 4}])
def test_options(cuda, padded_hook, options):
    pairs = [synth.make_padded_pair(i) for i in range(6)]
    opts = dict(options)
    use_delta = opts.pop('use_delta', True)
    trim = opts.pop('pad_silence', True)
    feat_kw = {k: v for k, v in opts.items() if k not in ('strict', 'radius')}
    expected = _oracle_chain(pairs, use_delta=use_delta, trim=trim,
                             radius=opts.get('radius', 32), strict=opts.get('strict', True),
                             **feat_kw)
    got = assemble.joint_frames_device(pairs, use_delta=use_delta, pad_silence=trim,
                                       pad_len=synth.PAD_LEN, **opts)
    assert np.array_equal(got.cpu().numpy(), expected)


def test_zero_frames_are_removed(cuda, padded_hook):
    a, b = synth.make_padded_pair(2)
    ma, mb = a.mel_cepstrum.data.copy(), b.mel_cepstrum.data.copy()
    ma[300:340, 1:] = 0.0
    mb[:, 1:] *= 0.0          # every joint row of this pair: tgt zero; rows 300..339: all zero
    a0 = synth.SynthFeature(ma, a.f0, a.is_voiced)
    b0 = synth.SynthFeature(mb, b.f0, b.is_voiced)
    pairs = [(a0, b0), synth.make_padded_pair(3)]
    expected = _oracle_chain(pairs)
    got = assemble.joint_frames_device(pairs, pad_silence=True, pad_len=synth.PAD_LEN)
    assert np.array_equal(got.cpu().numpy(), expected)
    dropped = _oracle_chain(pairs[1:]).shape[0] + 10      # some rows of pair 0 must be gone
    assert expected.shape[0] > dropped - 10


def test_converter_trains_on_the_device_array(cuda, padded_hook):
    """MelCepstrumConverter.train -> FeatureConverter.train: the B200 back-end receives the
    device-assembled tensor (no D2H / H2D of X between DTW and EM) and fits the same model as
    from the host array."""
    from oracle import gmm_ref
    n = 8
    src = {f'u{i}': synth.make_padded_pair(i)[0] for i in range(n)}
    tgt = {f'u{i}': synth.make_padded_pair(i)[1] for i in range(n)}
    keys = sorted(src)
    x_host = kw.make_dataset_to_array(
        kw.DeltaFeatureDataset(kw.MelCepstrumDataset(kw.align(src, tgt))), keys)
    resp0 = gmm_ref.kmeans_like_resp(x_host, 4, 0)
    seen = {}

    class Spy(kw.B200GMMFeatureConverter):
        def _train(self, dataarray, **kwargs):
            seen['array'] = dataarray
            return super()._train(dataarray, **kwargs)

    conv = kw.MelCepstrumConverter(Converter=Spy, components=4, verbose=0, resp_init=resp0,
                                   max_iter=5, tol=0.0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        conv.train(kw.align(src, tgt), keys)
    assert hasattr(seen['array'], 'is_cuda') and seen['array'].is_cuda
    assert np.array_equal(seen['array'].cpu().numpy(), x_host)
    assert (conv.order, conv.fs, conv.frame_period) == (synth.ORDER, synth.FS, synth.FRAME_PERIOD)
    ref = gmm_ref.numpy_em(x_host, resp0, max_iter=5, tol=0.0)
    assert np.abs(conv.gmm.means_ - ref['means']).max() <= 1e-9

"""The CUDA path against outputs of the REFERENCE'S OWN CODE (tests/golden/reference_chain.npz,
written by tests/golden/make_reference_golden.py from /root/reference with the absent third-party
packages stubbed).  Same synthetic utterances, same flows:

* kwiiyatta.vocoder.align.dtw_feature on silence-padded pairs       -> distance, strict-filtered path
* kwiiyatta.align(Feature, Feature)                                 -> warped source
* kwiiyatta.align(Dataset, Dataset) -> MelCepstrumConverter.train   -> the (N, 144) training array,
                                                                       the fitted model
* MelCepstrumConverter.convert(mcep, diff=False / True, mlpg=False) -> converted mel-cepstra
"""
import os
import warnings

import numpy as np
import pytest

import kwiiyatta_b200 as kw
from kwiiyatta_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'reference_chain.npz')


class _KeyedRng:
    def __init__(self, n, start=0):
        self.rngs = [np.random.default_rng((synth.SEED0 + i) * 7919 + 1) for i in range(n)]
        self.calls = 4 * start

    def normal(self, *args):
        rng = self.rngs[self.calls // 4]
        self.calls += 1
        return rng.normal(*args)


@pytest.fixture()
def hooks():
    holder = {'rng': None}
    kw.hooks.bind(pad_silence=lambda f, n: synth.pad_silence(f, n, holder['rng']),
                  feature=synth.feature, resample=synth.resample)
    yield holder
    kw.hooks.bind(pad_silence=None, feature=None, resample=None)


def test_alignment_equals_the_reference(cuda, hooks):
    g = np.load(GOLDEN)
    n = int(g['n_pairs'])
    pairs = [synth.make_pair(i) for i in range(n)]
    hooks['rng'] = _KeyedRng(n)
    padded = [(kw.hooks.get('pad_silence')(a, 100), kw.hooks.get('pad_silence')(b, 100))
              for a, b in pairs]
    for i, (dist, path) in enumerate(kw.dtw_feature_many(padded)):
        assert np.array_equal(path, g[f'path{i}'])
        assert abs(dist - float(g[f'dist{i}'])) <= 1e-12 * dist
    hooks['rng'] = _KeyedRng(n)
    for i, warped in enumerate(kw.align_many(pairs)):
        assert np.array_equal(warped.mel_cepstrum.data[:, 1], g[f'warped_c1_{i}'])
    # single-pair calls are the batch of one
    hooks['rng'] = _KeyedRng(n, start=2)
    one = kw.align(*pairs[2])
    assert np.array_equal(one.mel_cepstrum.data[:, 1], g['warped_c1_2'])


def test_training_chain_and_conversion_equal_the_reference(cuda, hooks):
    g = np.load(GOLDEN)
    n, k = int(g['n_pairs']), int(g['n_mix'])
    src = {f'utt{i:03d}.wav': synth.make_pair(i)[0] for i in range(n)}
    tgt = {f'utt{i:03d}.wav': synth.make_pair(i)[1] for i in range(n)}
    keys = sorted(src)
    hooks['rng'] = _KeyedRng(n)
    dataset = kw.align(src, tgt)
    chain = kw.DeltaFeatureDataset(kw.MelCepstrumDataset(dataset))
    x = kw.make_dataset_to_array(chain, keys)
    assert tuple(x.shape) == tuple(g['x_shape'])
    assert np.array_equal(x.sum(axis=1), g['x_rowsum'])
    assert np.array_equal(x[:3], g['x_head']) and np.array_equal(x[-3:], g['x_tail'])
    # the converter, through its own train(): same initial responsibilities as the reference run
    resp0 = np.zeros((len(x), k))
    resp0[np.arange(len(x)), g['labels0']] = 1.0
    conv = kw.MelCepstrumConverter(components=k, random_state=0, verbose=0, resp_init=resp0)
    hooks['rng'] = _KeyedRng(n)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        conv.train(dataset, keys)
    gmm = conv.gmm
    assert gmm.n_iter_ == int(g['n_iter']) and gmm.converged_ == bool(g['converged'])
    assert abs(gmm.lower_bound_ - float(g['lower_bound'])) <= 1e-9 * abs(float(g['lower_bound']))
    assert np.abs(gmm.weights_ - g['weights']).max() <= 1e-9
    assert np.abs(gmm.means_ - g['means']).max() <= 1e-8
    assert np.abs(np.einsum('kii->ki', gmm.covariances_) - g['cov_diag']).max() <= 1e-8
    assert np.abs(gmm.covariances_.ravel()[g['cov_idx']] - g['cov_sample']).max() <= 1e-8
    for i in (7, 8):
        mcep = synth.make_pair(i)[0].mel_cepstrum
        got = {'diff0': conv.convert(mcep, diff=False).data,
               'diff1': conv.convert(mcep, diff=True).data,
               'soft': conv.convert(mcep, mlpg=False).data}
        for name, data in got.items():
            assert data.shape == (len(mcep.data), 25)
            assert np.array_equal(data[:, 0], mcep.data[:, 0])
            expected = g[f'converted{i}_{name}']
            if i == 8:
                data = data.sum(axis=1)
            assert np.abs(data - expected).max() <= 1e-6      # north_star: 1e-4 absolute

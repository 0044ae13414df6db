"""Config 1 end to end through the reference-shaped API: align_even -> joint array ->
B200GMMFeatureConverter._train -> convert (diff and non-diff), against the oracle chain."""
import warnings

import numpy as np
import pytest

import kwiiyatta_b200 as kw
from kwiiyatta_b200 import synth
from oracle import delta_ref, gmm_ref, mlpg_ref
from util import oracle_joint_array, rel_err

pytestmark = pytest.mark.gpu


def test_config1_chain(cuda):
    pairs = [synth.make_padded_pair(i) for i in range(10)]
    x_exp, paths = oracle_joint_array(10)
    kw.set_pad_silence(lambda f, n: f)   # the synthetic features are already padded
    # alignment: identical paths => identical gathered frames
    aligned = kw.align_even_many(pairs, pad_silence=True, pad_len=synth.PAD_LEN)
    for (a, b), (pa, pb), p in zip(aligned, pairs, paths):
        assert np.array_equal(a.mel_cepstrum.data, pa.mel_cepstrum.data[p[0]])
        assert np.array_equal(b.mel_cepstrum.data, pb.mel_cepstrum.data[p[1]])
    x = kw.joint_array_from_pairs(pairs, pad_silence=True, pad_len=synth.PAD_LEN)
    assert x.shape == x_exp.shape and np.array_equal(x, x_exp)

    resp0 = gmm_ref.kmeans_like_resp(x_exp, 16, 0)
    ref = gmm_ref.sklearn_em(x_exp, resp0, max_iter=6, tol=0.0)
    conv = kw.B200GMMFeatureConverter(components=16, max_iter=6, tol=0.0, resp_init=resp0,
                                      verbose=0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        conv._train(x)
    assert rel_err(conv.gmm.means_, ref['means']) <= 1e-9
    assert rel_err(conv.gmm.covariances_, ref['covariances']) <= 1e-9
    assert abs(conv.gmm.lower_bound_ - ref['lower_bound']) <= 1e-9 * abs(ref['lower_bound'])

    for i in range(3):
        src, _ = synth.make_pair(i)
        feat = delta_ref.delta_features(src.mel_cepstrum.data[:, 1:])
        for diff in (False, True):
            exp = mlpg_ref.transform(feat, ref['weights'], ref['means'], ref['covariances'],
                                     diff=diff)
            got = conv.convert(feat, diff=diff)
            assert got.shape == (len(feat), 24)
            assert np.abs(got - exp).max() <= 1e-4

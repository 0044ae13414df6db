"""Generates tests/golden/reference_chain.npz by running the REFERENCE'S OWN CODE
(/root/reference, imported through tests/refenv.py with its absent third-party packages stubbed:
FastDTW / delta / MLPG = the oracle restatements, EM = the installed scikit-learn, started from
injected responsibilities because KMeans is version dependent).

Run from the repo root in the build container:  python tests/golden/make_reference_golden.py
The GPU test tests/test_gpu_reference_golden.py replays the same flows through the CUDA path on
the same synthetic utterances (kwiiyatta_b200.synth) and compares with this file."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import refenv  # noqa: E402
from kwiiyatta_b200 import synth  # noqa: E402
from oracle import gmm_ref  # noqa: E402

N_PAIRS = 6
N_MIX = 4


class KeyedRng:
    """Silence noise seeded per key (pair index), consumed in pad_silence's call order."""

    def __init__(self, n):
        self.rngs = [np.random.default_rng((synth.SEED0 + i) * 7919 + 1) for i in range(n)]
        self.calls = 0

    def normal(self, *args):
        rng = self.rngs[self.calls // 4]
        self.calls += 1
        return rng.normal(*args)


def main():
    out = {}
    with refenv.reference() as kwiiyatta:
        holder = {'rng': None}
        synthesizer = refenv.make_synthesizer(kwiiyatta, holder)
        align_mod = sys.modules['kwiiyatta.vocoder.align']

        def ref_feature(f):
            return refenv.to_reference_feature(kwiiyatta, f, synthesizer)

        class Corpus(kwiiyatta.converter.abc.Dataset):
            def __init__(self, items):
                super().__init__()
                self.items = items

            def keys(self):
                return self.items.keys()

            def get_data(self, key):
                return self.items[key]

        src, tgt = {}, {}
        for i in range(N_PAIRS):
            a, b = synth.make_pair(i)
            key = f'utt{i:03d}.wav'
            src[key], tgt[key] = ref_feature(a), ref_feature(b)
            # kwiiyatta.vocoder.align.dtw_feature on the padded pair (defaults: strict, r = 32)
            holder['rng'] = KeyedRng(N_PAIRS)
            holder['rng'].calls = 4 * i
            pa = kwiiyatta.pad_silence(src[key], 100)
            pb = kwiiyatta.pad_silence(tgt[key], 100)
            dist, path = align_mod.dtw_feature(pa, pb)
            out[f'dist{i}'] = dist
            out[f'path{i}'] = np.asarray(path, dtype=np.int16)
            # kwiiyatta.align(Feature, Feature): source index per target frame
            holder['rng'] = KeyedRng(N_PAIRS)
            holder['rng'].calls = 4 * i
            warped = kwiiyatta.align(src[key], tgt[key])
            out[f'warped_c1_{i}'] = warped.mel_cepstrum.data[:, 1].copy()
        keys = sorted(src)
        # the training chain of Config.train_converter (config.py:95-104)
        holder['rng'] = KeyedRng(N_PAIRS)
        dataset = kwiiyatta.align(Corpus(src), Corpus(tgt))

        class InjectedGMMFeatureConverter(kwiiyatta.converter.GMMFeatureConverter):
            """The reference's back-end; only sklearn's KMeans initialisation is replaced by
            given responsibilities (gmm_ref.kmeans_like_resp of the training array)."""

            def _train(self, dataarray, **kwargs):
                resp0 = gmm_ref.kmeans_like_resp(dataarray, self.gmm.n_components, 0)
                out['x_shape'] = np.array(dataarray.shape)
                out['x_rowsum'] = dataarray.sum(axis=1)
                out['x_head'] = dataarray[:3].copy()
                out['x_tail'] = dataarray[-3:].copy()
                out['labels0'] = resp0.argmax(1).astype(np.int16)

                class Injected(type(self.gmm)):
                    def _initialize_parameters(self, X, random_state, xp=None):
                        self._initialize(X, resp0)
                self.gmm.__class__ = Injected
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    super()._train(dataarray, **kwargs)

        conv = kwiiyatta.MelCepstrumConverter(Converter=InjectedGMMFeatureConverter,
                                              components=N_MIX, random_state=0, verbose=0)
        conv.train(dataset, keys)
        gmm = conv.gmm
        cov = gmm.covariances_
        idx = np.random.default_rng(20260105).integers(0, cov.size, 5000)
        out.update(weights=gmm.weights_, means=gmm.means_, cov_diag=np.einsum('kii->ki', cov),
                   cov_idx=idx, cov_sample=cov.ravel()[idx],
                   lower_bound=gmm.lower_bound_, n_iter=gmm.n_iter_, converged=gmm.converged_)
        # conversions: utterance 7 in full, utterance 8 as per-frame sums (file size)
        for i in (7, 8):
            s, _ = synth.make_pair(i)
            mcep = ref_feature(s).mel_cepstrum
            results = {'diff0': conv.convert(mcep, diff=False).data,
                       'diff1': conv.convert(mcep, diff=True).data,
                       'soft': conv.convert(mcep, mlpg=False).data}
            for name, data in results.items():
                out[f'converted{i}_{name}'] = data if i == 7 else data.sum(axis=1)
    np.savez_compressed(os.path.join(HERE, 'reference_chain.npz'), n_pairs=N_PAIRS, n_mix=N_MIX,
                        **out)
    print('written', os.path.join(HERE, 'reference_chain.npz'), 'X', out['x_shape'],
          'n_iter', out['n_iter'])


if __name__ == '__main__':
    main()

"""CPU: the repo's host-side mirrors against the REFERENCE'S OWN CODE, run here from
/root/reference with its missing third-party packages stubbed (tests/refenv.py).

Both sides use the same FastDTW / delta engine (the oracle's, substituted for the CUDA one because
this container has no GPU), so what is compared is everything the reference itself owns on the
path: kwiiyatta/vocoder/align.py:20-146, kwiiyatta/align.py:7-19, kwiiyatta/converter/dataset.py
(Parallel / Trimmed / Aligned datasets, make_dataset_to_array), converter/mcep.py,
converter/delta.py, converter/abc, converter/__init__.py:9-18.  The CUDA engine itself is compared
with the outputs of these same reference flows through tests/golden/reference_chain.npz
(tests/test_gpu_reference_golden.py).

Skipped where /root/reference does not exist (the GPU box)."""
import functools
import sys
import types

import numpy as np
import pytest

import kwiiyatta_b200 as kw
import refenv
from kwiiyatta_b200 import alignment, dataset as kds, synth
from oracle import delta_ref, dtw_c

pytestmark = pytest.mark.skipif(not refenv.available(), reason='/root/reference not present')

N_PAIRS = 4


def _pad_rng(i):
    return np.random.default_rng((synth.SEED0 + i) * 7919 + 1)


def _oracle_batch(pairs, radius=1, dist=2, precision=0, device=None):
    return [dtw_c.fastdtw(x, y, radius=radius, dist=dist) for x, y in pairs]


def _oracle_delta_many(features):
    return [delta_ref.delta_features(np.asarray(f, dtype=np.float64)) for f in features]


@pytest.fixture()
def engine(monkeypatch):
    """The repo's host code with the oracle standing in for the two CUDA entry points it calls."""
    monkeypatch.setattr(kw.fastdtw, 'fastdtw_batch', _oracle_batch)
    monkeypatch.setattr(kds, 'delta_many', _oracle_delta_many)
    holder = {'rng': None}

    def pad(feature, frame_len):
        return synth.pad_silence(feature, frame_len, holder['rng'])
    kw.hooks.bind(pad_silence=pad, feature=synth.feature, resample=synth.resample)
    yield holder
    kw.hooks.bind(pad_silence=None, feature=None, resample=None)


@pytest.fixture()
def ref():
    with refenv.reference() as kwiiyatta:
        holder = {'rng': None}
        synthesizer = refenv.make_synthesizer(kwiiyatta, holder)
        yield types.SimpleNamespace(
            kwiiyatta=kwiiyatta, holder=holder,
            feature=lambda f: refenv.to_reference_feature(kwiiyatta, f, synthesizer))


def _same_feature(ref_feature, synth_feature):
    return (np.array_equal(ref_feature.mel_cepstrum.data, synth_feature.mel_cepstrum.data)
            and np.array_equal(ref_feature.f0, synth_feature.f0))


@pytest.mark.parametrize('kwargs', [
    {}, {'vuv': 'f0', 'strict': False}, {'power': 'raw', 'vuv': None, 'strict': False},
    {'radius': 1}, {'power_pivot': 'median'}, {'power': None, 'vuv': 'f0'},
    {'pad_silence': False, 'radius': 4}])
def test_align_even_and_dtw_feature(ref, engine, kwargs):
    """kwiiyatta.align_even / vocoder.align.dtw_feature / make_feature vs the repo's."""
    ref_align = sys.modules['kwiiyatta.vocoder.align']     # (the attribute is the function)
    for i in range(N_PAIRS):
        a, b = synth.make_pair(i)
        ra, rb = ref.feature(a), ref.feature(b)
        dtw_kw = {k: v for k, v in kwargs.items() if k != 'pad_silence'}
        feat_kw = {k: v for k, v in dtw_kw.items() if k not in ('strict', 'radius')}
        assert np.array_equal(ref_align.make_feature(ra, ra.fs, **feat_kw),
                              alignment.make_feature(a, a.fs, **feat_kw))
        rd, rpath = ref_align.dtw_feature(ra, rb, **dtw_kw)
        gd, gpath = alignment.dtw_feature(a, b, **dtw_kw)
        assert rd == gd and np.array_equal(rpath, gpath)
        ref.holder['rng'], engine['rng'] = _pad_rng(i), _pad_rng(i)
        rx, ry = ref.kwiiyatta.align_even(ra, rb, **kwargs)
        gx, gy = kw.align_even(a, b, **kwargs)
        assert rx.frame_len == ry.frame_len == gx.frame_len == gy.frame_len
        assert _same_feature(rx, gx) and _same_feature(ry, gy)
    # the batched entry point is the same computation
    engine['rng'] = _pad_rng(0)
    one = kw.align_even(*synth.make_pair(0), **kwargs)
    engine['rng'] = _pad_rng(0)
    many = kw.align_even_many([synth.make_pair(0)], **kwargs)[0]
    assert np.array_equal(one[0].mel_cepstrum.data, many[0].mel_cepstrum.data)


@pytest.mark.parametrize('kwargs', [{}, {'strict': True}, {'vuv': 'voiced', 'radius': 2},
                                    {'pad_silence': False}, {'pad_len': 37}])
def test_single_pair_align(ref, engine, kwargs):
    """kwiiyatta.align(Feature, Feature) (vocoder/align.py:123-131 + project_path_iter) and the
    type dispatch of kwiiyatta/align.py:7-19."""
    for i in range(N_PAIRS):
        a, b = synth.make_pair(i)
        ra, rb = ref.feature(a), ref.feature(b)
        ref.holder['rng'], engine['rng'] = _pad_rng(i), _pad_rng(i)
        try:
            expected = ref.kwiiyatta.align(ra, rb, **kwargs)
        except ZeroDivisionError:
            with pytest.raises(ZeroDivisionError):
                kw.align(a, b, **kwargs)
            continue
        got = kw.align(a, b, **kwargs)
        assert _same_feature(expected, got)
        if not kwargs.get('strict'):
            assert got.frame_len == b.frame_len        # one source frame per target frame
    (rs, _), (gs, _) = _datasets(ref, 1)
    for bad_ref, bad in (((ref.feature(a), 3), (a, 3)), ((3, ref.feature(a)), (3, a)),
                         ((rs, ref.feature(a)), (gs, a)), ((ref.feature(a), rs), (a, gs))):
        with pytest.raises(TypeError) as e_ref:
            ref.kwiiyatta.align(*bad_ref)
        with pytest.raises(TypeError) as e_got:
            kw.align(*bad)
        assert str(e_ref.value).split(':')[0] == str(e_got.value).split(':')[0]


def test_project_path_iter_on_reference_paths(ref, engine):
    ref_align = sys.modules['kwiiyatta.vocoder.align']     # (the attribute is the function)
    for i in range(N_PAIRS):
        a, b = synth.make_padded_pair(i)
        for strict in (False, True):
            _, path = alignment.dtw_feature(a, b, strict=strict)
            for trim, trim_len in ((True, 100), (True, 1), (False, 1), (True, 250)):
                try:
                    expected = list(ref_align.project_path_iter(path, trim, trim_len))
                except ZeroDivisionError:
                    with pytest.raises(ZeroDivisionError):
                        alignment.project_path(path, trim, trim_len)
                    continue
                assert list(alignment.project_path_iter(path, trim, trim_len)) == expected


def _datasets(ref, n_pairs, trailing_zeros=()):
    """Matching (reference, repo) source / target datasets of synthetic utterances; utterances
    listed in ``trailing_zeros`` end in all-zero frames (what TrimmedDataset removes)."""
    class RefDataset(ref.kwiiyatta.converter.abc.Dataset):
        def __init__(self, items):
            super().__init__()
            self.items = items

        def keys(self):
            return self.items.keys()

        def get_data(self, key):
            return self.items[key]

    src_r, tgt_r, src_g, tgt_g = {}, {}, {}, {}
    for i in range(n_pairs):
        a, b = synth.make_pair(i)
        if i in trailing_zeros:
            def zero_tail(f, n):
                m = f.mel_cepstrum.data.copy()
                m[-n:] = 0.0
                return synth.SynthFeature(m, f.f0, f.is_voiced, f.fs, f.frame_period)
            a, b = zero_tail(a, 7), zero_tail(b, 12)
        key = f'utt{i:03d}.wav'
        src_r[key], tgt_r[key] = ref.feature(a), ref.feature(b)
        src_g[key], tgt_g[key] = a, b
    return (RefDataset(src_r), RefDataset(tgt_r)), (src_g, tgt_g)


class _KeyedRng:
    """Silence noise seeded per key, in the order pad_silence is called for that key (a, then b),
    whatever the order in which a chain visits the keys."""

    def __init__(self, keys):
        self.by_key = {key: _pad_rng(i) for i, key in enumerate(sorted(keys))}
        self.calls = 0
        self.order = sorted(keys)

    def normal(self, *args):
        rng = self.by_key[self.order[self.calls // 4]]
        self.calls += 1
        return rng.normal(*args)


@pytest.mark.parametrize('align_kwargs', [{}, {'radius': 1, 'vuv': None, 'power': 'raw',
                                               'pad_silence': False}])
def test_dataset_chain_to_training_array(ref, engine, align_kwargs):
    """Parallel -> Trimmed -> Aligned -> MelCepstrum -> Delta -> make_dataset_to_array, the
    reference's objects against the repo's batched ones (kwiiyatta/converter/dataset.py:35-77,
    converter/mcep.py:10-33, converter/delta.py:15-30)."""
    k = ref.kwiiyatta
    (rs, rt), (gs, gt) = _datasets(ref, N_PAIRS, trailing_zeros=(1,))
    keys = sorted(gs.keys())
    ref.holder['rng'], engine['rng'] = _KeyedRng(keys), _KeyedRng(keys)
    if align_kwargs:
        r_aligned = k.converter.AlignedDataset(
            k.converter.TrimmedDataset(k.ParallelDataset(rs, rt)), **align_kwargs)
        g_aligned = kw.AlignedDataset(kw.TrimmedDataset(kw.ParallelDataset(gs, gt)),
                                      **align_kwargs)
    else:
        r_aligned = k.align(rs, rt)            # the dispatching entry point, dataset branch
        g_aligned = kw.align(gs, gt)
    r_chain = k.converter.DeltaFeatureDataset(k.converter.MelCepstrumDataset(r_aligned))
    g_chain = kw.DeltaFeatureDataset(kw.MelCepstrumDataset(g_aligned))
    expected = k.converter.make_dataset_to_array(r_chain, keys)
    got = kw.make_dataset_to_array(g_chain, keys)
    assert expected.shape == got.shape and expected.shape[1] == 144
    assert np.array_equal(expected, got)
    assert r_chain.frame_period == g_chain.frame_period == synth.FRAME_PERIOD
    assert r_chain.base.order == g_chain.base.order == synth.ORDER
    # per-key access is the batch of one
    engine['rng'] = _KeyedRng(keys[:1])
    single = g_chain[keys[0]]
    ref.holder['rng'] = _KeyedRng(keys[:1])
    single_ref = r_chain[keys[0]]
    assert all(np.array_equal(x, y) for x, y in zip(single, single_ref))
    # the trimmed utterance really lost its zero tail on both sides
    tr, tg = k.converter.TrimmedDataset(k.ParallelDataset(rs, rt)), \
        kw.TrimmedDataset(kw.ParallelDataset(gs, gt))
    assert tr[keys[1]][0].frame_len == tg[keys[1]][0].frame_len == gs[keys[1]].frame_len - 7
    assert tr[keys[1]][1].frame_len == tg[keys[1]][1].frame_len == gt[keys[1]].frame_len - 12


def test_frame_period_mismatch_message(ref, engine):
    k = ref.kwiiyatta
    a, _ = synth.make_pair(0)
    odd = synth.SynthFeature(a.mel_cepstrum.data, a.f0, a.is_voiced, a.fs, frame_period=3)
    messages = []
    for make, base in ((k.converter.DeltaFeatureDataset,
                        k.converter.MelCepstrumDataset({'a': ref.feature(a),
                                                        'b': ref.feature(odd)})),
                       (kw.DeltaFeatureDataset, kw.MelCepstrumDataset({'a': a, 'b': odd}))):
        ds = make(base)
        ds['a']
        with pytest.raises(ValueError) as e:
            ds['b']
        messages.append(str(e.value))
    assert messages[0] == messages[1] == 'frame_period of "b" is 3 but others are 5'


class _Recorder:
    """A back-end that records what reaches it (the reference's NopConverter pattern,
    tests/kwiiyatta/test_converter.py:13-25)."""

    def __init__(self, components=None, random_state=None):
        self.args = (components, random_state)
        self.trained = None
        self.seen = []

    def _train(self, dataarray, **kwargs):
        self.trained = np.array(dataarray)

    def train(self, dataset, keys, **kwargs):
        from kwiiyatta_b200.dataset import make_dataset_to_array
        self._train(make_dataset_to_array(dataset, keys), **kwargs)

    def convert(self, feature, **kwargs):
        self.seen.append((np.array(feature), kwargs))
        return feature * 0.5 + 1.0


def test_converter_wrappers(ref, engine):
    """MelCepstrumConverter(Converter=...) = MelCepstrumFeatureConverter(DeltaFeatureConverter(
    back-end)): what the back-end is trained on and asked to convert, what comes back, and the
    two ValueErrors (kwiiyatta/converter/__init__.py:9-14, delta.py:33-50, mcep.py:36-61)."""
    k = ref.kwiiyatta
    (rs, rt), (gs, gt) = _datasets(ref, N_PAIRS)
    keys = sorted(gs.keys())

    class RefRecorder(_Recorder, k.converter.abc.FeatureConverter):
        def train(self, dataset, keys, **kwargs):
            return k.converter.abc.FeatureConverter.train(self, dataset, keys, **kwargs)

    class RepoRecorder(_Recorder, kw.FeatureConverter):
        def train(self, dataset, keys, **kwargs):
            return kw.FeatureConverter.train(self, dataset, keys, **kwargs)

    ref.holder['rng'], engine['rng'] = _KeyedRng(keys), _KeyedRng(keys)
    r_conv = k.MelCepstrumConverter(Converter=RefRecorder, components=3, random_state=1)
    g_conv = kw.MelCepstrumConverter(Converter=RepoRecorder, components=3, random_state=1)
    r_conv.train(k.align(rs, rt), keys)
    g_conv.train(kw.align(gs, gt), keys)
    assert r_conv.base.base.args == g_conv.base.base.args == (3, 1)
    assert np.array_equal(r_conv.base.base.trained, g_conv.base.base.trained)
    assert (r_conv.order, r_conv.fs, r_conv.frame_period) == \
        (g_conv.order, g_conv.fs, g_conv.frame_period) == (24, synth.FS, synth.FRAME_PERIOD)
    src, _ = synth.make_pair(9)
    for kwargs in ({}, {'diff': True}, {'mlpg': False}):
        expected = r_conv.convert(ref.feature(src).mel_cepstrum, **kwargs)
        got = g_conv.convert(src.mel_cepstrum, **kwargs)
        assert np.array_equal(expected.data, got.data) and got.data.shape == (src.frame_len, 25)
        assert np.array_equal(got.data[:, 0], src.mel_cepstrum.data[:, 0])     # the SOURCE c0
        (rf, rk), (gf, gk) = r_conv.base.base.seen[-1], g_conv.base.base.seen[-1]
        assert rk == gk == kwargs and np.array_equal(rf, gf) and gf.shape[1] == 72
    # the two checks
    wrong_order = synth.SynthFeature._Mcep(src.mel_cepstrum.data[:, :21], synth.FS,
                                           synth.FRAME_PERIOD)
    wrong_period = synth.SynthFeature._Mcep(src.mel_cepstrum.data, synth.FS, 3)
    for bad, text in ((wrong_order, 'order is expected to 24 but 20'),
                      (wrong_period, 'frame_period is expected to 5 but 3')):
        for conv in (r_conv, g_conv):
            with pytest.raises(ValueError) as e:
                conv.convert(bad)
            assert str(e.value) == text
    # use_delta=False: two layers, 24-dim training frames
    flat = kw.MelCepstrumConverter(use_delta=False, Converter=RepoRecorder)
    assert isinstance(flat.base, RepoRecorder)


def test_b200_backend_under_the_reference_factories(ref, monkeypatch):
    """The seams a kwiiyatta maintainer uses: the reference's own factory with
    ``Converter=B200GMMFeatureConverter`` (kwiiyatta/converter/__init__.py:9-14) and
    ``Config.create_converter`` with the factory bound to it (kwiiyatta/config.py:59-69, which
    forwards ``mcep_fs`` / ``components`` / ``random_state`` to the TOP-LEVEL factory).  No CUDA
    is needed to build the chain."""
    k = ref.kwiiyatta
    conv = k.MelCepstrumConverter(Converter=kw.B200GMMFeatureConverter, components=4,
                                  random_state=0)
    assert isinstance(conv, k.converter.MelCepstrumFeatureConverter)
    assert isinstance(conv.base, k.converter.DeltaFeatureConverter)
    backend = conv.base.base
    assert isinstance(backend, kw.B200GMMFeatureConverter)
    assert conv.gmm is backend.gmm                      # __getattr__ falls through the layers
    assert (backend.gmm.n_components, backend.gmm.max_iter, backend.gmm.random_state) == (4, 100, 0)
    assert backend.gmm.covariance_type == 'full' and backend.gmm.verbose == 1
    monkeypatch.setattr(sys, 'argv', ['kwiiyatta'])     # Config appends sys.argv[1:] (config.py:45-50)
    conf = k.Config()
    conf.add_converter_arguments()
    conf.parse_args(['--converter-components', '8', '--converter-seed', '3', '--mcep-fs',
                     '16000'])
    factory = functools.partial(k.MelCepstrumConverter, Converter=kw.B200GMMFeatureConverter)
    conv = conf.create_converter(Converter=factory)
    assert conv.mcep_fs == 16000
    assert (conv.gmm.n_components, conv.gmm.random_state) == (8, 3)
    # handing the back-end class itself to Config is NOT the seam: it would receive mcep_fs
    with pytest.raises(TypeError):
        conf.create_converter(Converter=kw.B200GMMFeatureConverter)

"""Feature-producer functions the hot path calls back into.

Silence padding, mel-cepstrum resampling and Feature copying are WORLD / pysptk territory
(kwiiyatta/vocoder/feature.py:19-41, kwiiyatta/vocoder/__init__.py:24-36, vocoder/feature.py:10-17)
and stay with the reference (north_star: the feature producer is outside the hot path).  The
alignment and converter front-ends reach them through this table: inside kwiiyatta they are
kwiiyatta's own functions (bound lazily on first use), anywhere else the integrator binds
equivalents (tests and bench.py bind kwiiyatta_b200.synth's)."""

_NAMES = ('pad_silence', 'resample', 'feature')
_table = dict.fromkeys(_NAMES)


def bind(**functions):
    """``bind(pad_silence=fn, resample=fn, feature=fn)``; ``None`` unbinds."""
    for name, fn in functions.items():
        if name not in _table:
            raise KeyError(f'unknown hook {name!r}; known: {_NAMES}')
        _table[name] = fn


def get(name):
    fn = _table[name]
    if fn is None:
        try:
            import kwiiyatta
        except ImportError:
            raise RuntimeError(
                f'kwiiyatta_b200.hooks: {name!r} is not bound and kwiiyatta is not importable; '
                f'call kwiiyatta_b200.hooks.bind({name}=...)') from None
        fn = getattr(kwiiyatta, name)
    return fn

"""CPU: the MLPG / delta oracle against independent formulations and the fixtures."""
import os

import numpy as np
import pytest

from kwiiyatta_b200 import synth
from oracle import delta_ref, mlpg_ref

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'mlpg.npz')


def test_banded_equals_dense():
    rng = np.random.default_rng(0)
    for t in (1, 2, 3, 10, 57):
        e = rng.standard_normal((t, 12))
        v = rng.uniform(0.1, 2.0, (t, 12))
        assert np.abs(mlpg_ref.mlpg_dense(e, v) - mlpg_ref.mlpg_banded(e, v)).max() <= 1e-10


def test_mlpg_reproduces_smooth_trajectory():
    """With tiny variances on consistent static/delta means the solution is the trajectory."""
    t = np.linspace(0, 3, 80)
    c = np.stack([np.sin(t), np.cos(2 * t)], axis=1)
    e = delta_ref.delta_features(c)
    v = np.full_like(e, 1e-3)
    y = mlpg_ref.mlpg_banded(e, v)
    assert np.abs(y - c).max() <= 1e-9


def test_delta_definition():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((20, 3))
    d = delta_ref.delta_features(x)
    assert np.array_equal(d[:, :3], x)
    assert np.allclose(d[1:-1, 3:6], 0.5 * (x[2:] - x[:-2]))
    assert np.allclose(d[0, 3:6], 0.5 * x[1]) and np.allclose(d[-1, 3:6], -0.5 * x[-2])
    assert np.allclose(d[1:-1, 6:], x[:-2] - 2 * x[1:-1] + x[2:])
    # window matrices of the solver and the correlate-based features are the same operator
    for w_idx, w in enumerate(delta_ref.DELTA_WINDOWS):
        wm = mlpg_ref.window_matrix(w, 20)
        assert np.allclose(wm @ x, d[:, 3 * w_idx:3 * w_idx + 3])


def test_zero_frame_helpers():
    x = np.ones((6, 2))
    x[4:] = 0
    x[1] = 0
    assert len(delta_ref.trim_zeros_frames(x)) == 4
    assert len(delta_ref.remove_zeros_frames(x)) == 3


def test_diff_rewrite_consistency():
    """diff=True must equal converting then subtracting the source statics' conditional mean:
    E_diff = E - x on every window block (the model of y - x given x)."""
    w, m, c = synth.make_joint_gmm(3, dim_half=6, seed=2, static_dim=2)
    rng = np.random.default_rng(2)
    src = delta_ref.delta_features(rng.standard_normal((30, 2)))
    _, mix0, e0, _ = mlpg_ref.transform(src, w, m, c, diff=False, return_internals=True)
    _, mix1, e1, _ = mlpg_ref.transform(src, w, m, c, diff=True, return_internals=True)
    assert np.array_equal(mix0, mix1)
    assert np.abs((e0 - src) - e1).max() <= 1e-9


def test_golden_fixture():
    g = np.load(GOLDEN)
    for diff in (0, 1):
        y, mix, _, _ = mlpg_ref.transform(g['src'], g['weights'], g['means'], g['covariances'],
                                          diff=bool(diff), return_internals=True)
        assert np.array_equal(mix, g[f'mix_diff{diff}'])
        assert np.abs(y - g[f'y_diff{diff}']).max() <= 1e-10


def test_vectorised_variant_equals_the_per_frame_restatement():
    """bench.py times both as CPU baselines; they are the same function."""
    w, m, c = synth.make_joint_gmm(8, seed=4)
    model = mlpg_ref.split_joint(w, m, c, diff=True)
    for frames in (1, 2, 3, 40, 173):
        src = delta_ref.delta_features(
            synth.make_source_utterances(1, frames=max(frames, 2))[0][:frames])
        for diff in (False, True):
            exp = mlpg_ref.transform(src, w, m, c, diff=diff)
            got = mlpg_ref.transform_vectorised(src, w, m, c, diff=diff,
                                                model=model if diff else None)
            assert np.abs(got - exp).max() <= 1e-10


def test_mc2b_recursion_inverts():
    from oracle import mlsa_ref
    rng = np.random.default_rng(3)
    mc = rng.standard_normal((50, 25))
    for alpha in (0.0, 0.41, 0.55):
        b = mlsa_ref.mc2b(mc, alpha)
        assert np.abs(mlsa_ref.b2mc(b, alpha) - mc).max() <= 1e-14
        assert np.array_equal(b[:, -1], mc[:, -1])
    assert np.array_equal(mlsa_ref.mc2b(mc, 0.0), mc)

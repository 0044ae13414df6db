#!/usr/bin/env python
"""A/B on one box: E-step accumulator stage released early (default) or after the partial sums
(KW_TC_LATE_RELEASE=1): conversion stage and EM E-step entry, CUDA events."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kwiiyatta_b200 as kw
from kwiiyatta_b200 import synth
from kwiiyatta_b200.mlpg import MLPG
from kwiiyatta_b200.delta import delta_features_device
dev = torch.device('cuda', 0)
w, m, c = synth.make_joint_gmm(128, seed=0)
model = type('M', (), dict(weights_=w, means_=m, covariances_=c, covariance_type='full'))
pg = MLPG(model, diff=False, device=dev, precision='tc')
n_utts, frames = 600, 1200
base = synth.make_source_utterances(8, frames=frames, seed0=synth.SEED0)
rng = np.random.default_rng(1000)
statics = np.concatenate([base[i % 8] + rng.normal(0, 0.02, base[0].shape) for i in range(n_utts)])
off = torch.arange(0, (n_utts + 1) * frames, frames, dtype=torch.int64, device=dev)
src = delta_features_device(torch.from_numpy(statics).to(dev), off, n_utts)
rngx = np.random.default_rng(0)
n, d, k = 176323, 144, 64
cen = rngx.standard_normal((k, d)) * 2.0
lab = rngx.integers(0, k, n)
x = cen[lab] + rngx.standard_normal((n, d))
resp0 = np.zeros((n, k)); resp0[np.arange(n), lab] = 1.0
gm = kw.GaussianMixture(n_components=k, max_iter=1, tol=0.0, resp_init=resp0, precision='tc', device=dev)
xd = gm.initialize(x)
gm.em_iteration(xd)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps


for rep in range(2):
    for g in ('2', '1'):
        for late in ('0', '1'):
            os.environ['KW_TC_LATE_RELEASE'] = late
            os.environ['KW_TC_G'] = g
            tc = timed(lambda: pg.transform_device(src, off, n_utts, frames))
            te = timed(lambda: gm._estep(torch, xd, for_mstep=True))
            print(f'G<={g} late_release={late}: conversion {tc:.3f} ms, EM E-step entry {te:.3f} ms')

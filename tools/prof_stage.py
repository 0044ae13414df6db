#!/usr/bin/env python
"""Small single-stage driver for ncu captures (never a bench value):
    python tools/prof_stage.py em|dtw|convert [--precision tc|fp64] [--reps 2]
Uses the same synthetic inputs as bench.py at a reduced size that still fills the GPU."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kwiiyatta_b200 as kw  # noqa: E402
from kwiiyatta_b200 import fastdtw as kfd  # noqa: E402
from kwiiyatta_b200 import synth  # noqa: E402
from kwiiyatta_b200.alignment import make_feature  # noqa: E402
from kwiiyatta_b200.delta import delta_features_device  # noqa: E402
from kwiiyatta_b200.mlpg import MLPG  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('stage', choices=['em', 'em_real', 'dtw', 'dtw_long', 'convert'])
    ap.add_argument('--iters', type=int, default=12)
    ap.add_argument('--precision', default='tc')
    ap.add_argument('--reps', type=int, default=2)
    ap.add_argument('--pairs', type=int, default=503)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    if a.stage in ('dtw', 'dtw_long'):
        feats = []
        if a.stage == 'dtw':
            for i in range(a.pairs):
                p, q = synth.make_padded_pair(i)
                feats.append((make_feature(p, p.fs), make_feature(q, q.fs)))
            radius = 32
        else:
            for i in range(min(a.pairs, 64)):
                p, q = synth.make_pair(i, length=4096)
                feats.append((make_feature(p, p.fs), make_feature(q, q.fs)))
            radius = -1
        tx = np.array([len(x) for x, _ in feats], dtype=np.int32)
        ty = np.array([len(y) for _, y in feats], dtype=np.int32)
        xd = torch.from_numpy(np.concatenate([x for x, _ in feats])).to(dev)
        yd = torch.from_numpy(np.concatenate([y for _, y in feats])).to(dev)
        for _ in range(a.reps):
            kfd.fastdtw_batch_device(xd, yd, tx, ty, radius=radius, dist=2)
        torch.cuda.synchronize()
    elif a.stage == 'em_real':
        # the bench workload itself: joint frames of 503 pairs, labels from 5 Lloyd passes
        from kwiiyatta_b200 import kmeans
        kw.set_pad_silence(lambda f, n: f)
        pairs = [synth.make_padded_pair(i) for i in range(a.pairs)]
        x = kw.joint_array_from_pairs(pairs, pad_silence=True, pad_len=synth.PAD_LEN)
        n, k = len(x), 64
        xd0 = torch.from_numpy(x).to(dev)
        lab = kmeans.kmeans_labels(xd0, k, seed=0, n_lloyd=5)
        resp0 = torch.zeros((n, k), dtype=torch.float64, device=dev)
        resp0[torch.arange(n, device=dev), lab] = 1.0
        gm = kw.GaussianMixture(n_components=k, max_iter=1, tol=0.0, resp_init=resp0,
                                precision=a.precision, device=dev)
        xd = gm.initialize(xd0)
        for _ in range(a.iters):
            gm.em_iteration(xd)
        torch.cuda.synchronize()
        print('PROFILE MARK')
        for _ in range(a.reps):
            gm.em_iteration(xd)
        torch.cuda.synchronize()
    elif a.stage == 'em':
        rng = np.random.default_rng(0)
        n, d, k = 176323, 144, 64
        centres = rng.standard_normal((k, d)) * 0.5
        x = centres[rng.integers(0, k, n)] + 0.6 * rng.standard_normal((n, d))
        lab = rng.integers(0, k, n)
        resp0 = torch.zeros((n, k), dtype=torch.float64, device=dev)
        resp0[torch.arange(n, device=dev), torch.from_numpy(lab).to(dev)] = 1.0
        gm = kw.GaussianMixture(n_components=k, max_iter=1, tol=0.0, resp_init=resp0,
                                precision=a.precision, device=dev)
        xd = gm.initialize(x)
        for _ in range(a.reps):
            gm.em_iteration(xd)
        torch.cuda.synchronize()
    else:
        w, m, c = synth.make_joint_gmm(128, seed=0)
        model = type('M', (), dict(weights_=w, means_=m, covariances_=c, covariance_type='full'))
        pg = MLPG(model, precision=a.precision, device=dev)
        n_utts, frames = 1200, 600
        base = synth.make_source_utterances(8, frames=frames)
        rng = np.random.default_rng(1)
        st = np.concatenate([base[i % 8] + rng.normal(0, 0.02, base[0].shape)
                             for i in range(n_utts)])
        off = torch.arange(0, (n_utts + 1) * frames, frames, dtype=torch.int64, device=dev)
        src = delta_features_device(torch.from_numpy(st).to(dev), off, n_utts)
        for _ in range(a.reps):
            pg.transform_device(src, off, n_utts, frames)
        torch.cuda.synchronize()
    print('done', a.stage)


if __name__ == '__main__':
    main()

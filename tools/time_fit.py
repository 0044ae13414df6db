#!/usr/bin/env python
"""Whole-fit time (100 EM iterations, bench workload) for several re-sort periods."""
import os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kwiiyatta_b200 as kw
from kwiiyatta_b200 import synth, kmeans
kw.set_pad_silence(lambda f, n: f)
pairs = [synth.make_padded_pair(i) for i in range(503)]
x = kw.joint_array_from_pairs(pairs, pad_silence=True, pad_len=synth.PAD_LEN)
n = len(x); K = 64
xd = torch.from_numpy(x).cuda()
lab = kmeans.kmeans_labels(xd, K, seed=0, n_lloyd=5)
for reorder in (0, 5, 10, 20):
    for rep in range(2):
        resp0 = torch.zeros((n, K), dtype=torch.float64, device='cuda'); resp0[torch.arange(n, device='cuda'), lab] = 1
        gm = kw.GaussianMixture(n_components=K, max_iter=100, tol=0.0, resp_init=resp0, precision='tc',
                                reorder_every=reorder)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            gm.fit(xd)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'reorder_every={reorder}: fit 100 iterations {dt * 1e3:.1f} ms = {dt * 10:.3f} ms/iter, lower bound {gm.lower_bound_:.9f}')

#!/usr/bin/env python
"""M-step / EM-iteration time with and without the frame reordering, at several EM ages."""
import os, sys
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
def tm(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps
for reorder in (0, 10):
    resp0 = torch.zeros((n, K), dtype=torch.float64, device='cuda'); resp0[torch.arange(n, device='cuda'), lab] = 1
    gm = kw.GaussianMixture(n_components=K, max_iter=1, tol=0.0, resp_init=resp0, precision='tc',
                            reorder_every=reorder)
    xdev = gm.initialize(x)
    done = 0
    for target in (1, 9, 21, 41):
        while done < target:
            gm.em_iteration(xdev); done += 1
        cen = gm._means[gm._cur]
        m = tm(lambda: gm._accumulate(torch, xdev, cen))
        print(f'reorder_every={reorder} after {done} iterations: M-step {m:.3f} ms')

#!/usr/bin/env python
"""Phase timing of the host path of fastdtw_batch (bench workload)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kwiiyatta_b200 as kw
from kwiiyatta_b200 import fastdtw as kfd, synth, _lib
from kwiiyatta_b200.alignment import make_feature
feats = []
for i in range(503):
    p, q = synth.make_padded_pair(i)
    feats.append((make_feature(p, p.fs), make_feature(q, q.fs)))
dev = torch.device('cuda', 0)
for rep in range(3):
    t = [time.perf_counter()]
    xs = [kfd._prep(x) for x, _ in feats]; ys = [kfd._prep(y) for _, y in feats]
    tx = np.array([len(x) for x in xs], dtype=np.int32); ty = np.array([len(y) for y in ys], dtype=np.int32)
    t.append(time.perf_counter())
    xd = _lib.gather_to_device(torch, xs, dev, 'dtw_x'); t.append(time.perf_counter())
    yd = _lib.gather_to_device(torch, ys, dev, 'dtw_y'); t.append(time.perf_counter())
    res = kfd.fastdtw_batch_device(xd, yd, tx, ty, 32, 2); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    out = res.to_host(); t.append(time.perf_counter())
    names = ['prep', 'gather x', 'gather y', 'plan+launch', 'sync', 'to_host']
    print(' | '.join(f'{n} {1e3 * (b - a):.2f} ms' for n, a, b in zip(names, t[:-1], t[1:])), '| total %.2f ms' % (1e3 * (t[-1] - t[0])))

#!/usr/bin/env python
"""Wall time of the packed host APIs on the bench workload (several repetitions)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kwiiyatta_b200 import fastdtw as kfd, synth
from kwiiyatta_b200.alignment import make_feature
feats = []
for i in range(503):
    p, q = synth.make_padded_pair(i)
    feats.append((make_feature(p, p.fs), make_feature(q, q.fs)))
tx = np.array([len(x) for x, _ in feats], dtype=np.int32); ty = np.array([len(y) for _, y in feats], dtype=np.int32)
xh = torch.from_numpy(np.concatenate([x for x, _ in feats])).pin_memory()
yh = torch.from_numpy(np.concatenate([y for _, y in feats])).pin_memory()
for n_chunks in (1, 2, 3):
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = kfd.fastdtw_batch_packed(xh, yh, tx, ty, radius=32, dist=2, n_chunks=n_chunks)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'n_chunks={n_chunks}: {dt * 1e3:.2f} ms')

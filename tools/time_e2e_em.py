#!/usr/bin/env python
"""Where the end-to-end EM time (bench.py `e2e`) goes: B200GMMFeatureConverter._train on a pinned
host array, 100 iterations, repeated; wall clock of the pieces."""
import os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import kwiiyatta_b200 as kw
from kwiiyatta_b200 import synth
dev = torch.device('cuda', 0)
kw.hooks.bind(pad_silence=lambda f, n: f, feature=synth.feature, resample=synth.resample)
padded = [synth.make_padded_pair(i) for i in range(503)]
x = kw.joint_array_from_pairs(padded, pad_silence=True, pad_len=synth.PAD_LEN)
n = len(x)
lab = bench.lloyd_labels(x, 64, seed=0)
xp = torch.from_numpy(x).pin_memory()
for rep in range(5):
    conv = kw.B200GMMFeatureConverter(components=64, max_iter=100, tol=0.0, verbose=0, device=dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r0 = torch.zeros((n, 64), dtype=torch.float64, device=dev)
    r0[torch.arange(n, device=dev), torch.from_numpy(lab).to(dev)] = 1.0
    conv.gmm.resp_init = r0
    torch.cuda.synchronize(); t1 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        if rep == 4:
            import cProfile, pstats
            pr = cProfile.Profile(); pr.enable()
        conv._train(xp)
        if rep == 4:
            pr.disable()
    t2 = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f'rep {rep}: resp_init {t1 - t0:.3f} s, _train returns after {t2 - t1:.3f} s, '
          f'device idle after {t3 - t1:.3f} s -> {n * 100 / (t3 - t0):.3e} frames/s/iter')
    del conv, r0
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)

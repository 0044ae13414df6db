#!/usr/bin/env python
"""Times kw_gmm_precision_cholesky (the finalize kernel without the statistics read) and the
whole finalize on the prof_stage EM inputs with CUDA events."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kwiiyatta_b200 as kw
from kwiiyatta_b200 import _lib

dev = torch.device('cuda', 0)
rng = np.random.default_rng(0)
n, d, k = 40000, 144, 64
centres = rng.standard_normal((k, d)) * 0.5
lab = rng.integers(0, k, n)
x = centres[lab] + 0.6 * rng.standard_normal((n, d))
resp0 = torch.zeros((n, k), dtype=torch.float64, device=dev)
resp0[torch.arange(n, device=dev), torch.from_numpy(lab).to(dev)] = 1.0
gm = kw.GaussianMixture(n_components=k, max_iter=1, tol=0.0, resp_init=resp0, precision='tc', device=dev)
xd = gm.initialize(x)
gm.em_iteration(xd)
torch.cuda.synchronize()
cen = gm._means[gm._cur].clone()
def fin():
    gm._finalize(torch, cen, 1)
for _ in range(3): fin()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): fin()
e1.record(); e1.synchronize()
print('finalize us per call', e0.elapsed_time(e1) / 20 * 1e3)

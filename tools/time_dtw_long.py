"""Config 4: long singing-transfer DTW, n pairs of 4096 x 4096 frames, unconstrained window."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kwiiyatta_b200 import fastdtw as kfd, synth
from kwiiyatta_b200.alignment import make_feature
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
base = []
for i in range(8):
    p, q = synth.make_pair(i, length=4096)
    base.append((make_feature(p, p.fs), make_feature(q, q.fs)))
rng = np.random.default_rng(0)
feats = [(base[i % 8][0] + rng.normal(0, 0.01, base[0][0].shape), base[i % 8][1]) for i in range(n)]
tx = np.full(n, 4096, dtype=np.int32); ty = np.full(n, 4096, dtype=np.int32)
xd = torch.from_numpy(np.concatenate([x for x, _ in feats])).cuda(); yd = torch.from_numpy(np.concatenate([y for _, y in feats])).cuda()
for prec in (0, 1):
    for _ in range(2): r = kfd.fastdtw_batch_device(xd, yd, tx, ty, -1, 2, precision=prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): r = kfd.fastdtw_batch_device(xd, yd, tx, ty, -1, 2, precision=prec)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f'config4 n={n} precision={prec}: {ms:.2f} ms  {int(r.cells.sum())/ms*1e3:.3e} cells/s')

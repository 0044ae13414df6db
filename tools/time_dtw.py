"""Time one batched FastDTW call (503 synthetic pairs) with CUDA events."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kwiiyatta_b200 import fastdtw as kfd, synth
from kwiiyatta_b200.alignment import make_feature
n = int(sys.argv[1]) if len(sys.argv) > 1 else 503
feats = []
for i in range(n):
    p, q = synth.make_padded_pair(i)
    feats.append((make_feature(p, p.fs), make_feature(q, q.fs)))
tx = np.array([len(x) for x, _ in feats], dtype=np.int32); ty = np.array([len(y) for _, y in feats], dtype=np.int32)
xd = torch.from_numpy(np.concatenate([x for x, _ in feats])).cuda(); yd = torch.from_numpy(np.concatenate([y for _, y in feats])).cuda()
for _ in range(3): r = kfd.fastdtw_batch_device(xd, yd, tx, ty, 32, 2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): r = kfd.fastdtw_batch_device(xd, yd, tx, ty, 32, 2)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f'KW_DTW_NT={os.environ.get("KW_DTW_NT")} {ms:.3f} ms  {int(r.cells.sum())/ms*1e3:.3e} cells/s')

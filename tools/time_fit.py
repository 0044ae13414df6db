import sys, time, warnings
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import kwiiyatta_b200 as kw
from kwiiyatta_b200 import synth, kmeans
kw.hooks.bind(pad_silence=lambda f, n: f, feature=synth.feature, resample=synth.resample)
padded = [synth.make_padded_pair(i) for i in range(503)]
x = kw.joint_array_from_pairs(padded, pad_silence=True, pad_len=synth.PAD_LEN)
xp = torch.from_numpy(x).pin_memory()
dev = torch.device('cuda', 0)
for rep in range(3):
    conv = kw.B200GMMFeatureConverter(components=64, random_state=0, verbose=0, device=dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        conv._train(xp)
    torch.cuda.synchronize(); print('fit', time.perf_counter() - t0, conv.gmm.n_iter_)
xd = xp.to(dev)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lab = kmeans.kmeans_labels(xd, 64, seed=0)
    torch.cuda.synchronize(); print('kmeans', time.perf_counter() - t0)

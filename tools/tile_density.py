"""How many (64-frame tile, component) pairs carry responsibility mass in the bench workload."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kwiiyatta_b200 as kw
from kwiiyatta_b200 import synth, kmeans
n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
kw.set_pad_silence(lambda f, n: f)
pairs = [synth.make_padded_pair(i) for i in range(n_pairs)]
x = kw.joint_array_from_pairs(pairs, pad_silence=True, pad_len=synth.PAD_LEN)
n = len(x); K = 64
xd = torch.from_numpy(x).cuda()
lab = kmeans.kmeans_labels(xd, K, seed=0, n_lloyd=5)
resp0 = torch.zeros((n, K), dtype=torch.float64, device='cuda'); resp0[torch.arange(n, device='cuda'), lab] = 1
gm = kw.GaussianMixture(n_components=K, max_iter=1, tol=0.0, resp_init=resp0)
xdev = gm.initialize(x)
for it in range(1, 31):
    gm.em_iteration(xdev)
    if it in (1, 2, 5, 10, 20, 30):
        gm._estep(torch, xdev)
        r = gm._resp[:, :n]
        nt = n // 64
        rt = r[:, :nt * 64].reshape(K, nt, 64).amax(dim=2)
        order = torch.argsort(lab, stable=True)
        rs = r[:, order][:, :nt * 64].reshape(K, nt, 64).amax(dim=2)
        order2 = torch.argsort(r.argmax(dim=0), stable=True)
        rs2 = r[:, order2][:, :nt * 64].reshape(K, nt, 64).amax(dim=2)
        print(f'   sorted by the INITIAL labels: {(rs > 1e-16).float().mean().item():.3f}; sorted by the current argmax: {(rs2 > 1e-16).float().mean().item():.3f}')
        print(f'iter {it}: frames {n}, mean comps/frame > 1e-16: {(r > 1e-16).sum().item() / n:.1f}; '
              f'(tile,k) with max r > 1e-16: {(rt > 1e-16).float().mean().item():.3f}, > 1e-10: {(rt > 1e-10).float().mean().item():.3f}')

"""Small end-to-end pass of every kernel for compute-sanitizer (memcheck)."""
import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kwiiyatta_b200 as kw
from kwiiyatta_b200 import synth
from kwiiyatta_b200.alignment import make_feature
rng = np.random.default_rng(0)
pairs = [(rng.standard_normal((tx, 26)), rng.standard_normal((ty, 26))) for tx, ty in [(300, 280), (129, 257), (70, 40), (5, 3)]]
for prec in (0, 1):
    kw.fastdtw.fastdtw_batch(pairs, radius=4, dist=2, precision=prec)
kw.fastdtw.fastdtw_batch(pairs[:2], radius=-1, dist=1)
x = rng.standard_normal((777, 40)) + rng.integers(0, 3, 777)[:, None]
for prec in ('fp64', 'tc'):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        gm = kw.GaussianMixture(n_components=3, max_iter=2, tol=0.0, random_state=0, precision=prec).fit(x)
    gm.predict_proba(x[:100])
w, m, c = synth.make_joint_gmm(5, seed=1)
model = type('M', (), dict(weights_=w, means_=m, covariances_=c, covariance_type='full'))
src = [kw.delta_features(s) for s in synth.make_source_utterances(2, frames=90)]
for prec in ('fp64', 'tc'):
    kw.MLPG(model, precision=prec).transform_many(src)
kw.MLPG(model, diff=True).transform(src[0])
w2, m2, c2 = synth.make_joint_gmm(3, dim_half=24, seed=2, static_dim=24)
model2 = type('M', (), dict(weights_=w2, means_=m2, covariances_=c2, covariance_type='full'))
kw.MLPG(model2, windows=kw.DELTA_WINDOWS[0:1]).transform(rng.standard_normal((50, 24)))
# round 2: tie modes / margins, device-side assembly, D = 144 tensor-core fit (corner warps,
# A in TMEM), packed host pipelines, mc2b
kw.fastdtw.fastdtw_batch(pairs, radius=2, dist=2, tie_mode='cython', with_margin=True)
kw.fastdtw.fastdtw_batch(pairs[:2], radius=-1, dist=2, with_margin=True)
kw.hooks.bind(pad_silence=lambda f, n: f, feature=synth.feature, resample=synth.resample)
padded = [synth.make_padded_pair(i) for i in range(3)]
xj = kw.joint_array_from_pairs(padded, pad_silence=True, pad_len=synth.PAD_LEN, device_resident=True)
xa = xj.cpu().numpy()
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    kw.GaussianMixture(n_components=2, max_iter=2, tol=0.0, random_state=0, precision='tc').fit(xa)
    kw.GaussianMixture(n_components=3, max_iter=2, tol=0.0, random_state=0, precision='tc').fit(x[:, :20])
import torch
pg = kw.MLPG(model, precision='tc')
pg.CHUNK_FRAMES = 60
lens = [len(s) for s in src]
pg.transform_packed(torch.from_numpy(np.concatenate(src)).pin_memory(), lens)
kw.fastdtw.fastdtw_batch_packed(torch.from_numpy(np.concatenate([p[0] for p in pairs])).pin_memory(),
                                torch.from_numpy(np.concatenate([p[1] for p in pairs])).pin_memory(),
                                [len(p[0]) for p in pairs], [len(p[1]) for p in pairs], radius=2,
                                n_chunks=2)
kw.mc2b(rng.standard_normal((30, 25)), 0.41)
print('sanitize smoke done')

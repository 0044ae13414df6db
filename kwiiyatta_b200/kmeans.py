"""K-means initial labels for the EM (sklearn ``init_params='kmeans'``,
sklearn/mixture/_base.py:119-128).

Scope note: this is the step *before* the hot path.  It runs on
the GPU with stock torch tensor ops (k-means++ seeding, Lloyd passes), not hand-written
kernels, and is not bit-compatible with any sklearn KMeans version (their seeding consumes the
RNG differently across versions).  Parity tests inject ``resp_init`` instead."""
import numpy as np


def _dist2(x, c):
    return (x * x).sum(1, keepdim=True) - 2.0 * (x @ c.t()) + (c * c).sum(1)[None, :]


def kmeans_labels(x, k, seed=None, group=None, n_lloyd=30):
    import torch
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else 0
    n, d = x.shape
    rs = np.random.RandomState(seed)
    centres = torch.empty((k, d), dtype=x.dtype, device=x.device)
    if rank == 0:
        if n < k:
            raise ValueError(f'need at least {k} frames on rank 0 to seed k-means, got {n}')
        first = int(rs.randint(n))
        centres[0] = x[first]
        closest = _dist2(x, centres[0:1]).squeeze(1).clamp_min_(0.0)
        for j in range(1, k):
            probs = (closest / closest.sum()).cpu().numpy()
            idx = int(np.searchsorted(np.cumsum(probs), rs.random_sample()))
            idx = min(idx, n - 1)
            centres[j] = x[idx]
            closest = torch.minimum(closest, _dist2(x, centres[j:j + 1]).squeeze(1).clamp_min_(0.0))
    if multi:
        dist.broadcast(centres, src=dist.get_global_rank(group, 0) if group is not None else 0,
                       group=group)
    labels = _dist2(x, centres).argmin(1)
    for _ in range(n_lloyd):
        sums = torch.zeros((k, d), dtype=x.dtype, device=x.device)
        sums.index_add_(0, labels, x)
        counts = torch.bincount(labels, minlength=k).to(x.dtype)
        if multi:
            dist.all_reduce(sums, group=group)
            dist.all_reduce(counts, group=group)
        nz = counts > 0
        centres[nz] = sums[nz] / counts[nz][:, None]
        new = _dist2(x, centres).argmin(1)
        changed = (new != labels).sum().to(torch.float64)
        if multi:
            dist.all_reduce(changed, group=group)
        labels = new
        if changed.item() == 0:
            break
    return labels

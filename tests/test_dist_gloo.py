"""CPU, world_size 2 over gloo: the multi-GPU host logic (pair sharding, the all-reduce of EM
sufficient statistics, identical finalisation on every rank).  The statistics themselves come
from the numpy oracle here; on GPUs they come from kw_gmm_mstep_accumulate."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kwiiyatta_b200 import dist as kdist
from oracle import gmm_ref


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _stats(x, resp, centres):
    k, d = centres.shape
    out = np.zeros(k * (1 + d + d * d) + 2)
    for i in range(k):
        xc = x - centres[i]
        blk = out[i * (1 + d + d * d):(i + 1) * (1 + d + d * d)]
        blk[0] = resp[:, i].sum()
        blk[1:1 + d] = resp[:, i] @ xc
        blk[1 + d:] = ((resp[:, i] * xc.T) @ xc).ravel()
    out[-1] = len(x)
    return out


def _finalize(stats, centres, reg):
    k, d = centres.shape
    sb = 1 + d + d * d
    nk = np.array([stats[i * sb] for i in range(k)]) + 10 * np.finfo(float).eps
    means = np.empty((k, d))
    cov = np.empty((k, d, d))
    for i in range(k):
        delta = stats[i * sb + 1:i * sb + 1 + d] / nk[i]
        means[i] = centres[i] + delta
        cov[i] = stats[i * sb + 1 + d:(i + 1) * sb].reshape(d, d) / nk[i] - np.outer(delta, delta)
        cov[i].flat[::d + 1] += reg
    return nk / nk.sum(), means, cov


def _worker(rank, world, port, sizes, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((600, 5)) + rng.integers(0, 3, 600)[:, None]
    resp = rng.uniform(size=(600, 3))
    resp /= resp.sum(1, keepdims=True)
    centres = rng.standard_normal((3, 5))
    mine = kdist.shard_indices(sizes, world)[rank]
    # frames follow their pair: pair p owns frames [bounds[p], bounds[p+1])
    bounds = np.linspace(0, 600, len(sizes) + 1).astype(int)
    rows = np.concatenate([np.arange(bounds[p], bounds[p + 1]) for p in mine])
    local = torch.from_numpy(_stats(x[rows], resp[rows], centres))
    kdist.allreduce_stats(local)
    w, m, c = _finalize(local.numpy(), centres, 1e-6)
    q.put((rank, sorted(int(p) for p in mine), w, m, c))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_statistics_allreduce_equals_single_process():
    sizes = [900 * 800, 500 * 450, 700 * 720, 300 * 310, 650 * 640, 810 * 790, 400 * 380]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort(key=lambda r: r[0])
    assert sorted(results[0][1] + results[1][1]) == list(range(len(sizes)))
    rng = np.random.default_rng(0)
    x = rng.standard_normal((600, 5)) + rng.integers(0, 3, 600)[:, None]
    resp = rng.uniform(size=(600, 3))
    resp /= resp.sum(1, keepdims=True)
    nk, means, cov = gmm_ref.estimate_parameters(x, resp, 1e-6)
    for _, _, w, m, c in results:
        assert np.abs(w - nk / nk.sum()).max() <= 1e-12
        assert np.abs(m - means).max() <= 1e-12
        assert np.abs(c - cov).max() <= 1e-12
    # every rank finalises bit-identical parameters (no broadcast needed)
    assert np.array_equal(results[0][3], results[1][3])
    assert np.array_equal(results[0][4], results[1][4])


def test_shard_indices_balance():
    rng = np.random.default_rng(1)
    sizes = rng.integers(300 * 300, 950 * 950, 503)
    for world in (1, 2, 4, 8):
        shards = kdist.shard_indices(sizes, world)
        assert sorted(np.concatenate(shards).tolist()) == list(range(503))
        loads = np.array([sizes[s].sum() for s in shards], dtype=float)
        assert loads.max() / loads.mean() <= 1.02

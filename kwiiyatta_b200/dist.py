"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The hot path has exactly one exchange: the all-reduce (sum, float64) of the EM sufficient
statistics per iteration, K*(1+D+D*D)+2 values.  DTW and conversion
shard by utterance (pair) with no data-path collective."""
import numpy as np


def shard_indices(sizes, world_size):
    """Longest-processing-time assignment of items (cost = sizes[i], e.g. tx*ty of a pair) to
    ranks.  Deterministic; returns a list of index arrays, one per rank."""
    sizes = np.asarray(sizes, dtype=np.float64)
    order = np.argsort(-sizes, kind='stable')
    loads = np.zeros(world_size)
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(loads))
        shards[r].append(int(i))
        loads[r] += sizes[i]
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def allreduce_stats(stats, group=None, n_components=None, dim=None, exchange_form=False):
    """In-place sum of a statistics vector over the ranks of ``group`` (no-op when
    torch.distributed is not initialised or the world has one rank).

    ``exchange_form`` (CUDA vector with its layout given by ``n_components``, ``dim``): only
    n_k, the first moments and the upper triangle of the symmetric second moments travel -- half
    the bytes (kw_gmm_stats_pack / kw_gmm_stats_unpack).  Off by default: over NVLink the full
    10.7 MB vector (K = 64, D = 144) is reduced in 48 us, the packed one in 56 us including the
    two copy kernels (tools/time_allreduce.py, 2 x B200) -- the exchange is latency bound there,
    so halving the bytes buys nothing; it is for links where bandwidth is the limit."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return stats
    if not exchange_form or n_components is None or dim is None or not stats.is_cuda:
        dist.all_reduce(stats, group=group)
        return stats
    import torch
    from . import _lib
    lib = _lib.lib()
    packed = torch.empty(lib.kw_gmm_stats_packed_len(n_components, dim), dtype=torch.float64,
                         device=stats.device)
    stream = torch.cuda.current_stream(stats.device).cuda_stream
    _lib.check(lib.kw_gmm_stats_pack(n_components, dim, stats.data_ptr(), packed.data_ptr(),
                                     stream), 'kw_gmm_stats_pack')
    dist.all_reduce(packed, group=group)
    _lib.check(lib.kw_gmm_stats_unpack(n_components, dim, packed.data_ptr(), stats.data_ptr(),
                                       stream), 'kw_gmm_stats_unpack')
    return stats


def init_from_env(backend='nccl'):
    """torchrun-style initialisation: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the
    environment; binds this process to its GPU."""
    import os

    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if backend == 'nccl':
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        kwargs = {}
        if backend == 'nccl':
            kwargs['device_id'] = torch.device('cuda', local_rank)
        dist.init_process_group(backend, **kwargs)
    return int(os.environ.get('RANK', '0')), world, local_rank

#!/usr/bin/env python
"""Benchmark of the kwiiyatta alignment + spectral-mapping hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Workload (BASELINE.json configs[1], per GPU; weak scaling over ranks):
  * DTW      503 synthetic ATR503-shaped padded pairs, 26-dim features, FastDTW radius 32
  * EM       one iteration of 64-mix full-covariance EM over the aligned (N, 144) joint frames
             (sufficient statistics all-reduced over ranks when N > 1)
  * convert  128-mix posterior + MLPG over 1 200 x 600 = 720 000 source frames (configs[4])
A "step" of the headline metric is one EM iteration; the DTW and conversion stages are timed
in their own K-step loops and reported under "stages".  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if '--impl' in sys.argv and 'reference' in sys.argv:
    # the CPU arm uses every host core whatever the launcher exported (torchrun sets
    # OMP_NUM_THREADS=1); must happen before numpy loads its BLAS
    for _var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_var] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PAIRS = 503
N_MIX_EM = 64
N_MIX_CONVERT = 128
N_UTTS = 1200
UTT_FRAMES = 600
RADIUS = 32
LONG_PAIRS = 256          # configs[3]: 256 pairs of 4096 x 4096 frames, unconstrained
LONG_FRAMES = 4096
CPU_EM_FRAMES = 12000     # frames of the CPU arms' EM sample
CPU_EM_ITERS = 10


def lloyd_labels(x, k, seed, passes=5):
    """Initial hard labels shared by the GPU arm and the CPU arms: ``k`` distinct frames drawn
    with ``default_rng(seed)`` as centres, ``passes`` Lloyd passes in numpy (the reference
    initialises with KMeans; sklearn's own is version dependent, so both arms get this one)."""
    rng = np.random.default_rng(seed)
    centres = x[rng.choice(len(x), size=k, replace=False)].copy()
    x2 = (x * x).sum(1)
    lab = None
    for _ in range(passes):
        d2 = x2[:, None] - 2.0 * (x @ centres.T) + (centres * centres).sum(1)[None]
        lab = d2.argmin(1)
        for j in range(k):
            m = lab == j
            if m.any():
                centres[j] = x[m].mean(0)
    return lab


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p['hbm_gbs'], bf16_tflops=p['bf16_tflops'],
                    bf16_tflops_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    source='measured')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0,
                source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.QUERY}',
                     '--format=csv,noheader,nounits'], capture_output=True, text=True,
                    timeout=5).stdout.strip().split(',')
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for name, val in zip(names, out[2:]):
                    if val.strip().lower().startswith('active'):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def ctypes_device_sms(device):
    import ctypes
    from kwiiyatta_b200 import _lib
    sms = ctypes.c_int(148)
    _lib.lib().kw_device_info(device, ctypes.byref(sms), None, None, None)
    return sms.value


def build_dtw_inputs(first_pair, n_pairs):
    from kwiiyatta_b200 import synth
    from kwiiyatta_b200.alignment import make_feature
    feats, padded = [], []
    for i in range(first_pair, first_pair + n_pairs):
        a, b = synth.make_padded_pair(i)
        padded.append((a, b))
        feats.append((make_feature(a, a.fs), make_feature(b, b.fs)))
    return padded, feats


# --------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------
_JSON_FD = None


def capture_stdout():
    """Rank 0 must print ONE JSON line on stdout: send everything else that lands on fd 1
    (NCCL's version banner, library chatter) to stderr and keep the real stdout for the line."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, line)


def run_b200(args):
    import torch
    import torch.distributed as dist
    import kwiiyatta_b200 as kw
    from kwiiyatta_b200 import _lib, synth
    from kwiiyatta_b200 import fastdtw as kfd
    from kwiiyatta_b200.mlpg import MLPG

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: there is no CPU fallback '
                         '(use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    _lib.lib()
    peaks = measured_peaks()
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def sum_over_ranks(v):
        if world > 1:
            t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            return float(t.item())
        return float(v)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed_loop(fn, flush):
        """W warm-up + K timed steps, CUDA events on the launching stream, barrier + synchronize
        on both sides; returns max-over-ranks total ms of the K steps.  With ``flush`` the L2 is
        overwritten between steps (outside the timed intervals, so each step is timed on its
        own); without it the K steps are enqueued back to back between one pair of events."""
        for _ in range(W):
            fn()
        barrier()
        if flush:
            total = 0.0
            for _ in range(K):
                flush_buf.fill_(1)
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                total += e0.elapsed_time(e1)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                fn()
            e1.record()
            e1.synchronize()
            total = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(total)

    # ---------------- inputs (this rank's shard: pairs [rank*503, (rank+1)*503)) -----------
    if args.total_pairs:
        # strong scaling (configs[2]: 5 030 pairs over the ranks): contiguous blocks of pairs
        lo = args.total_pairs * rank // world
        hi = args.total_pairs * (rank + 1) // world
        first_pair, n_pairs = lo, hi - lo
    else:
        first_pair, n_pairs = rank * args.pairs, args.pairs
    padded, feats = build_dtw_inputs(first_pair, n_pairs)
    tx = np.array([len(x) for x, _ in feats], dtype=np.int32)
    ty = np.array([len(y) for _, y in feats], dtype=np.int32)
    x_host = np.concatenate([x for x, _ in feats])
    y_host = np.concatenate([y for _, y in feats])
    x_dev = torch.from_numpy(x_host).to(dev)
    y_dev = torch.from_numpy(y_host).to(dev)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()

    # ---------------- stage 1: DTW ---------------------------------------------------------
    res = kfd.fastdtw_batch_device(x_dev, y_dev, tx, ty, radius=RADIUS, dist=2)
    cells_local = int(res.cells.sum().item())
    nominal_local = int((tx.astype(np.int64) * ty).sum())
    dtw_ms = timed_loop(lambda: kfd.fastdtw_batch_device(x_dev, y_dev, tx, ty, RADIUS, 2),
                        flush=True)
    cells_total = sum_over_ranks(cells_local)
    nominal_total = sum_over_ranks(nominal_local)
    dtw_cells_per_s = cells_total * K / (dtw_ms / 1e3)
    # end to end through the batched host API: features in pinned host memory in, host paths out
    x_pin, y_pin = torch.from_numpy(x_host).pin_memory(), torch.from_numpy(y_host).pin_memory()
    kfd.fastdtw_batch_packed(x_pin, y_pin, tx, ty, radius=RADIUS, dist=2, device=dev)  # warm-up
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    kfd.fastdtw_batch_packed(x_pin, y_pin, tx, ty, radius=RADIUS, dist=2, device=dev)
    torch.cuda.synchronize()
    dtw_e2e_s = time.perf_counter() - t0
    dtw_e2e = sum_over_ranks(cells_local) / max_over_ranks(dtw_e2e_s * 1e3) * 1e3
    # ... and from a list of separate pageable arrays (fastdtw_batch: host copies included)
    kfd.fastdtw_batch(feats, radius=RADIUS, dist=2, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kfd.fastdtw_batch(feats, radius=RADIUS, dist=2, device=dev)
    torch.cuda.synchronize()
    dtw_e2e_lists = cells_local / (time.perf_counter() - t0)
    del x_pin, y_pin

    # ---------------- joint frames for EM (align_even -> mcep -> delta -> hstack) ----------
    kw.set_pad_silence(lambda f, n: f)     # the synthetic features are already padded
    x_joint = kw.joint_array_from_pairs(padded, pad_silence=True, pad_len=synth.PAD_LEN)
    n_frames, dim = x_joint.shape
    # initial hard labels: five Lloyd passes (the reference initialises with KMeans), computed
    # outside every timed region by the function the CPU arms use too
    labels0 = lloyd_labels(x_joint, N_MIX_EM, seed=first_pair)

    def make_gm(max_iter):
        resp0 = torch.zeros((n_frames, N_MIX_EM), dtype=torch.float64, device=dev)
        resp0[torch.arange(n_frames, device=dev), torch.from_numpy(labels0).to(dev)] = 1.0
        return kw.GaussianMixture(n_components=N_MIX_EM, max_iter=max_iter, tol=0.0,
                                  resp_init=resp0, device=dev, precision=args.precision)

    gm = make_gm(1)
    xj_dev = gm.initialize(x_joint)
    lb_first = gm.em_iteration(xj_dev)
    gm._estep(torch, xj_dev)
    density = float((gm._resp[:, :n_frames] > 1e-16).sum().item()) / n_frames
    # the same first iteration on the FP64 kernels (not timed): the benchmarked path is checked
    # against it here as well as in tests/test_gpu_gmm_scale.py
    check = None
    if args.precision == 'tc' and world == 1:
        g64 = kw.GaussianMixture(n_components=N_MIX_EM, max_iter=1, tol=0.0, device=dev,
                                 precision='fp64', resp_init=make_gm(1).resp_init)
        x64 = g64.initialize(x_joint)
        lb64 = g64.em_iteration(x64)
        g64.em_iteration(x64)
        gm.em_iteration(xj_dev)
        mu_tc, mu_64 = gm._means[gm._cur], g64._means[g64._cur]
        per_comp = (mu_tc - mu_64).abs().max(dim=1).values / mu_64.abs().max()
        n_k = g64._weights * n_frames
        well = n_k >= 8 * dim
        check = {'lower_bound_tc': lb_first, 'lower_bound_fp64': lb64,
                 'lower_bound_rel_diff': abs(lb_first - lb64) / abs(lb64),
                 'means_rel_diff_after_2_iterations': {
                     'components_with_at_least_8_frames_per_dim':
                         float(per_comp[well].max()) if bool(well.any()) else None,
                     'all_components': float(per_comp.max()),
                     'n_components_below_8_frames_per_dim': int((~well).sum()),
                     'smallest_component_frames': float(n_k.min())},
                 'tolerance': 1e-5,
                 'note': 'tensor-core path vs the FP64 kernels from the same initial labels; '
                         'components the initialisation left with few frames per dimension '
                         'are ill-conditioned (see tests/test_gpu_gmm_scale.py)'}
        del g64, x64
        torch.cuda.empty_cache()

    # ---------------- stage 2 (headline): EM iteration -------------------------------------
    em_ms = timed_loop(lambda: gm.em_iteration_async(xj_dev), flush=False)
    lb_last = gm.last_lower_bound()
    frames_total = sum_over_ranks(n_frames)
    em_value = frames_total * K / (em_ms / 1e3)
    # per-entry-point timing for the roofline (E-step and M-step statistics)
    def time_call(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps
    estep_ms = time_call(lambda: gm._estep(torch, xj_dev, for_mstep=True))
    mstep_ms = time_call(lambda: gm._accumulate(torch, xj_dev, gm._means[gm._cur]))
    flops_half = 2.0 * n_frames * N_MIX_EM * dim * dim
    # end to end through the converter back-end API with host buffers
    x_pinned = torch.from_numpy(x_joint).pin_memory()
    e2e_iters = 100     # the reference's own default (GMMFeatureConverter(max_iter=100)), tol = 0
    import warnings

    def train_once():
        """One user-level training call: labels -> resp_init on the device, _train(host array)."""
        conv = kw.B200GMMFeatureConverter(components=N_MIX_EM, max_iter=e2e_iters, tol=0.0,
                                          verbose=0, device=dev, precision=args.precision)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            lab_dev = torch.from_numpy(labels0).to(dev)
            r0 = torch.zeros((n_frames, N_MIX_EM), dtype=torch.float64, device=dev)
            r0[torch.arange(n_frames, device=dev), lab_dev] = 1.0
            conv.gmm.resp_init = r0
            conv._train(x_pinned)
        torch.cuda.synchronize()
        return conv, time.perf_counter() - t0

    # the first call of a process also pays for device / pinned allocations and first launches
    # (0.29 s against 0.19 s, tools/time_e2e_em.py): one untimed warm-up call, like the W
    # warm-up steps of the device-timed loop, then the timed one
    conv, em_e2e_first_s = train_once()
    del conv
    conv, em_e2e_s = train_once()
    em_e2e_ms = max_over_ranks(em_e2e_s * 1e3)
    em_e2e = frames_total * e2e_iters / (em_e2e_ms / 1e3)
    model_bytes = sum(a.nbytes for a in (conv.gmm.weights_, conv.gmm.means_,
                                         conv.gmm.covariances_, conv.gmm.precisions_cholesky_))
    del conv
    # the fit a user gets: reference defaults (tol = 1e-3, max_iter = 100, KMeans initialisation,
    # here kwiiyatta_b200.kmeans on the device), timed whole with host buffers
    fit = None
    if world == 1:
        with warnings.catch_warnings():       # warm-up: the k-means path's one-time costs
            warnings.simplefilter('ignore')   # (library handles, first launches) on a small fit
            kw.B200GMMFeatureConverter(components=N_MIX_EM, random_state=0, verbose=0, max_iter=2,
                                       device=dev, precision=args.precision)._train(
                x_pinned[:32768])
        conv = kw.B200GMMFeatureConverter(components=N_MIX_EM, random_state=0, verbose=0,
                                          device=dev, precision=args.precision)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            conv._train(x_pinned)
        torch.cuda.synchronize()
        fit_s = time.perf_counter() - t0
        fit = {'seconds': fit_s, 'n_iter': int(conv.gmm.n_iter_),
               'converged': bool(conv.gmm.converged_),
               'lower_bound': float(conv.gmm.lower_bound_),
               'frames_per_s_per_iter': n_frames * conv.gmm.n_iter_ / fit_s,
               'note': "B200GMMFeatureConverter(components=64)._train(host array): "
                       "init_params='kmeans' on the device, tol=1e-3, max_iter=100"}
        del conv
    del x_pinned
    gm._resp = None
    gm._ws = None
    del gm, xj_dev
    torch.cuda.empty_cache()

    # ---------------- stage 3: conversion (configs[4]) -------------------------------------
    n_utts = args.utts
    w, m, c = synth.make_joint_gmm(N_MIX_CONVERT, seed=0)
    model = type('M', (), dict(weights_=w, means_=m, covariances_=c, covariance_type='full'))
    paramgen = MLPG(model, diff=False, device=dev, precision=args.precision)
    base = synth.make_source_utterances(8, frames=UTT_FRAMES, seed0=synth.SEED0 + rank)
    rng = np.random.default_rng(1000 + rank)
    statics = np.concatenate([base[i % 8] + rng.normal(0, 0.02, base[0].shape)
                              for i in range(n_utts)])
    off = torch.arange(0, (n_utts + 1) * UTT_FRAMES, UTT_FRAMES, dtype=torch.int64, device=dev)
    from kwiiyatta_b200.delta import delta_features_device
    src_dev = delta_features_device(torch.from_numpy(statics).to(dev), off, n_utts)
    conv_ms = timed_loop(lambda: paramgen.transform_device(src_dev, off, n_utts, UTT_FRAMES),
                         flush=True)
    conv_frames_total = sum_over_ranks(n_utts * UTT_FRAMES)
    conv_value = conv_frames_total * K / (conv_ms / 1e3)
    conv_flops = (2.0 * N_MIX_CONVERT + 2.0) * conv_frames_total * 72 * 72
    src_host = src_dev.cpu().numpy()
    src_list = [src_host[i * UTT_FRAMES:(i + 1) * UTT_FRAMES] for i in range(n_utts)]
    # end to end: source frames in one pinned host block -> converted frames in a pinned block
    src_pin = torch.from_numpy(src_host).pin_memory()
    out_pin = torch.empty((n_utts * UTT_FRAMES, 24), dtype=torch.float64).pin_memory()
    utt_lens = [UTT_FRAMES] * n_utts
    paramgen.transform_packed(src_pin, utt_lens, out=out_pin)     # warm-up (buffers, streams)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    paramgen.transform_packed(src_pin, utt_lens, out=out_pin)
    torch.cuda.synchronize()
    conv_e2e = conv_frames_total / (max_over_ranks((time.perf_counter() - t0) * 1e3) / 1e3)
    # ... and from / to lists of separate pageable arrays (host copies included)
    paramgen.transform_many(src_list)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    paramgen.transform_many(src_list)
    torch.cuda.synchronize()
    conv_e2e_lists = n_utts * UTT_FRAMES / (time.perf_counter() - t0)
    del src_pin, out_pin

    del src_dev, paramgen
    torch.cuda.empty_cache()

    # ---------------- configs[3]: long unconstrained DTW, 256 pairs of 4096 x 4096 ----------
    long_stage = None
    if args.long_pairs > 0:
        lx, ly = [], []
        for i in range(8):
            a, b = synth.make_pair(i, length=LONG_FRAMES)
            from kwiiyatta_b200.alignment import make_feature as _mf
            lx.append(_mf(a, a.fs))
            ly.append(_mf(b, b.fs))
        n_long = args.long_pairs
        ltx = np.full(n_long, LONG_FRAMES, dtype=np.int32)
        lrng = np.random.default_rng(77 + rank)
        lx_dev = torch.from_numpy(np.concatenate(
            [lx[i % 8] + lrng.normal(0, 0.01, lx[0].shape) for i in range(n_long)])).to(dev)
        ly_dev = torch.from_numpy(np.concatenate(
            [ly[i % 8] + lrng.normal(0, 0.01, ly[0].shape) for i in range(n_long)])).to(dev)
        long_ms = timed_loop(lambda: kfd.fastdtw_batch_device(lx_dev, ly_dev, ltx, ltx, -1, 2),
                             flush=True)
        long_cells = sum_over_ranks(float(n_long) * LONG_FRAMES * LONG_FRAMES)
        long_stage = {'ms': long_ms / K, 'cells': long_cells,
                      'value': long_cells * K / (long_ms / 1e3), 'pairs_per_gpu': n_long}
        del lx_dev, ly_dev
        torch.cuda.empty_cache()

    clock_info = clocks.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline (bounded samples, rank 0, N = 1 only) -------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baselines(feats, x_joint, (w, m, c), src_list)

    em_tflops = 2 * flops_half / (em_ms / K / 1e3) / 1e12
    tc = args.precision == 'tc'
    names = ('estep_tc_kernel', 'mstats_tc2_kernel') if tc else ('gmm_estep_kernel', 'gmm_mstats_kernel')
    dominant_ms, dominant = (estep_ms, names[0]) if estep_ms >= mstep_ms else (mstep_ms, names[1])
    dom_tflops = flops_half / (dominant_ms / 1e3) / 1e12
    peak = peaks['bf16_tflops_sustained']
    # DTW: FP64-pipe issue bound: 72 FP64-pipe instructions per cell,
    # 64 FP64 lanes per SM
    sm_count = ctypes_device_sms(local_rank)
    sm_hz = (clock_info['sm_max_mhz'] or 1965.0) * 1e6 if clock_info else 1965e6
    dtw_peak = sm_count * 64 * sm_hz / 72.0
    out = {
        'metric': 'GMM-EM frames/s/iter',
        'value': em_value,
        'unit': 'frames/s/iter',
        'n_gpus': world,
        'steps': K,
        'warmup': W,
        'ms_per_step': em_ms / K,
        'higher_is_better': True,
        'scaling': 'strong' if args.total_pairs else 'weak',
        'vs_baseline': None,
        'dtype': 'f64' if args.precision == 'fp64' else 'f16x2-split (fp32 accumulate) + f64',
        'data': 'synthetic',
        'config': {
            'workload': 'configs[1]: 503 synthetic ATR503-shaped pairs per GPU -> FastDTW r=32 '
                        '-> (N,144) joint frames -> 64-mix full-cov EM iteration; stages.convert '
                        'is configs[4] (128-mix, 720k frames)',
            'pairs_per_gpu': n_pairs, 'total_pairs': args.total_pairs or n_pairs * world,
            'frames_per_gpu': int(n_frames), 'dim': int(dim),
            'n_components': N_MIX_EM, 'precision': args.precision,
            'init': 'hard labels from 5 numpy Lloyd passes (bench.lloyd_labels, the same '
                    'function the CPU arms use; the reference initialises with KMeans)',
            'timing': 'EM: K iterations enqueued back to back between one pair of CUDA events '
                      '(em_iteration_async: no host read-back inside the timed region); DTW / '
                      'convert: per-step events with an L2 flush between steps',
            'mean_components_per_frame_above_1e-16': density,
            'l2': 'EM inputs (X 8*N*144 B + resp) exceed L2; DTW/convert stages flush L2 with a '
                  '256 MiB write between timed steps',
        },
        'e2e': {
            'value': em_e2e, 'unit': 'frames/s/iter',
            'h2d_bytes_per_step': int(x_joint.nbytes + labels0.nbytes),
            'd2h_bytes_per_step': int(model_bytes),
            'note': f'B200GMMFeatureConverter._train on a pinned host (N,144) array, '
                    f'{e2e_iters} iterations incl. H2D of X, initial M-step and D2H of the model; '
                    f'second call of the process (the first, with its allocations: '
                    f'{em_e2e_first_s:.3f} s)',
        },
        'gpu_launches': (11 if tc else 6) * K,
        'roofline': {
            'bound': 'tensor', 'kernel': dominant, 'achieved': dom_tflops, 'peak': peak,
            'unit': 'TFLOP/s', 'frac': dom_tflops / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel
            # in the ncu --set full capture of this workload (profiles/ncu_r2h_mstats_tc2.txt,
            # profiles/ncu_r2h_estep_tc.txt); algorithmic: packed X 113 MB + prepared weights
            'traffic': ({'mstats_tc2_kernel': 320.2e6 + 62.9e6, 'estep_tc_kernel': 116.5e6 + 53.9e6}
                        .get(dominant) if n_frames == 176323 else None),
            'traffic_unit': 'bytes per launch (ncu, 1 GPU)',
            'peak_source': f"{peaks['source']} bf16_tflops_sustained",
            'algorithmic': '2*N*K*D^2 flop per launch (half of the 4*N*K*D^2 EM iteration); '
                           'duration = CUDA events around the C-ABI entry that launches it',
            'mma_passes': 3 if tc else None,
            'frac_of_peak_over_passes': (dom_tflops / (peak / 3)) if tc else None,
            'estep_ms': estep_ms, 'mstep_accumulate_ms': mstep_ms,
            'em_iteration_tflops': em_tflops,
            'other_contraction': {
                'kernel': names[1] if dominant == names[0] else names[0],
                'achieved': flops_half / (min(estep_ms, mstep_ms) / 1e3) / 1e12,
                'frac': flops_half / (min(estep_ms, mstep_ms) / 1e3) / 1e12 / peak},
            'power': 'the M-step statistics kernel runs at the 1 kW board power cap with dense '
                     'data (tools/time_mstep.py: 984 W, sw_power_cap active); bf16_tflops_sustained '
                     'is the cuBLAS rate under the same cap',
        },
        'cpu_baseline': cpu['em'] if cpu else None,
        'clocks': clock_info,
        'parity_check': check,
        'lower_bound_after_timed_steps': lb_last,
        'e2e_fit': fit,
        # the other two stages' headline figures as flat keys (details under "stages")
        'dtw_cells_per_s': dtw_cells_per_s,
        'dtw_roofline_frac': dtw_cells_per_s / world / dtw_peak,
        'dtw_e2e_cells_per_s': dtw_e2e,
        'dtw_long_cells_per_s': long_stage['value'] if long_stage else None,
        'dtw_long_roofline_frac': (long_stage['value'] / world / dtw_peak) if long_stage else None,
        'convert_frames_per_s': conv_value,
        'convert_roofline_frac': conv_flops / (conv_ms / K / 1e3) / 1e12 / world / peak,
        'convert_e2e_frames_per_s': conv_e2e,
        'stages': {
            'dtw_long': None if long_stage is None else {
                'metric': 'DTW cells/s', 'value': long_stage['value'], 'unit': 'cells/s',
                'ms_per_step': long_stage['ms'], 'cells_per_step': long_stage['cells'],
                'workload': f"configs[3]: {long_stage['pairs_per_gpu']} pairs per GPU of "
                            f'{LONG_FRAMES} x {LONG_FRAMES} frames, 26-dim, unconstrained '
                            '(radius < 0), exact fp64 local distances',
                'roofline': {'bound': 'fp64 issue', 'achieved': long_stage['value'] / world,
                             'peak': dtw_peak, 'unit': 'cells/s per GPU',
                             'frac': long_stage['value'] / world / dtw_peak,
                             'model': 'SMs x 64 FP64 lanes x f_max / 72 FP64-pipe ops per cell'},
            },
            'dtw': {
                'metric': 'DTW cells/s', 'value': dtw_cells_per_s, 'unit': 'cells/s',
                'ms_per_step': dtw_ms / K, 'cells_per_step': cells_total,
                'nominal_cells_per_s': nominal_total * K / (dtw_ms / 1e3),
                'roofline': {'bound': 'fp64 issue', 'achieved': dtw_cells_per_s / world,
                             'peak': dtw_peak, 'unit': 'cells/s per GPU',
                             'frac': dtw_cells_per_s / world / dtw_peak,
                             'model': 'SMs x 64 FP64 lanes x f_max / 72 FP64-pipe ops per cell'},
                'e2e': {'value': dtw_e2e, 'unit': 'cells/s',
                        'h2d_bytes_per_step': int(x_host.nbytes + y_host.nbytes),
                        'd2h_bytes_per_step': int(8 * (tx.sum() + ty.sum()) + 20 * n_pairs),
                        'api': 'fastdtw_batch_packed (pinned host blocks in, host paths out)',
                        'from_lists_of_pageable_arrays': dtw_e2e_lists},
                'cpu_baseline': cpu['dtw'] if cpu else None,
            },
            'convert': {
                'metric': 'MLPG converted frames/s', 'value': conv_value, 'unit': 'frames/s',
                'ms_per_step': conv_ms / K, 'frames_per_step': conv_frames_total,
                'n_components': N_MIX_CONVERT,
                'hbm_boundary_gbs': conv_value * 768 / 1e9,
                'roofline': {'bound': 'tensor', 'kernel': 'estep_tc_kernel (posterior, 60 % of the stage)'
                             if tc else 'gmm_estep_kernel (posterior)',
                             'achieved': conv_flops / (conv_ms / K / 1e3) / 1e12 / world,
                             'peak': peak, 'unit': 'TFLOP/s per GPU',
                             'frac': conv_flops / (conv_ms / K / 1e3) / 1e12 / world / peak,
                             'model': '2*N*K*Dh^2 posterior + 2*N*Dh^2 conditional-mean flop over the '
                                      'WHOLE stage time (conservative); 768 B/frame at the converter '
                                      'boundary is only hbm_boundary_gbs, far from the HBM bound'},
                'e2e': {'value': conv_e2e, 'unit': 'frames/s',
                        'h2d_bytes_per_step': int(n_utts * UTT_FRAMES * 72 * 8),
                        'd2h_bytes_per_step': int(n_utts * UTT_FRAMES * 24 * 8),
                        'api': 'MLPG.transform_packed (pinned host block in, pinned block out)',
                        'from_lists_of_pageable_arrays': conv_e2e_lists},
                'cpu_baseline': cpu['convert'] if cpu else None,
            },
        },
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# CPU legs (oracle = test infrastructure; used here only as the reported baseline)
# --------------------------------------------------------------------------------------------
def _dtw_worker(pair):
    from oracle import dtw_c
    x, y = pair
    return dtw_c.fastdtw(x, y, radius=RADIUS, dist=2, return_cells=True)[2]


def cpu_dtw(feats, n_sample, procs):
    import multiprocessing as mp
    from oracle import dtw_c
    dtw_c.lib()
    sample = feats[:n_sample]
    t0 = time.perf_counter()
    if procs > 1:
        with mp.get_context('fork').Pool(procs) as pool:
            cells = sum(pool.map(_dtw_worker, sample))
    else:
        cells = sum(_dtw_worker(p) for p in sample)
    dt = time.perf_counter() - t0
    return cells / dt, cells, dt


def cpu_em(x, n_sample, iters):
    from oracle import gmm_ref
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:
        threadpool_limits = None
    xs = np.ascontiguousarray(x[:n_sample])
    labels = lloyd_labels(xs, N_MIX_EM, seed=0)
    resp0 = np.zeros((len(xs), N_MIX_EM))
    resp0[np.arange(len(xs)), labels] = 1.0
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    if threadpool_limits is not None:
        with threadpool_limits(limits=cores):
            gmm_ref.sklearn_em(xs, resp0, max_iter=iters, tol=0.0)
    else:
        gmm_ref.sklearn_em(xs, resp0, max_iter=iters, tol=0.0)
    dt = time.perf_counter() - t0
    return len(xs) * iters / dt, dt


def _convert_worker(job):
    from oracle import mlpg_ref
    (w, m, c), srcs = job
    try:
        from threadpoolctl import threadpool_limits
        limit = threadpool_limits(limits=1)      # one BLAS thread per worker process
    except ImportError:
        limit = None
    model = mlpg_ref.split_joint(w, m, c, False)
    for s in srcs:
        mlpg_ref.transform_vectorised(s, w, m, c, model=model)
    del limit
    return sum(len(s) for s in srcs)


def cpu_convert(model, src_list, n_sample):
    """Faithful per-frame restatement on one core, and the vectorised variant over all cores."""
    import multiprocessing as mp
    from oracle import mlpg_ref
    w, m, c = model
    t0 = time.perf_counter()
    frames = 0
    for s in src_list[:n_sample]:
        mlpg_ref.transform(s, w, m, c)
        frames += len(s)
    dt = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    per = 2
    jobs = [(model, src_list[i * per:(i + 1) * per]) for i in range(cores)
            if src_list[i * per:(i + 1) * per]]
    t1 = time.perf_counter()
    with mp.get_context('fork').Pool(len(jobs)) as pool:
        vframes = sum(pool.map(_convert_worker, jobs))
    vdt = time.perf_counter() - t1
    return frames / dt, dt, vframes / vdt, vdt, len(jobs)


def cpu_baselines(feats, x_joint, model, src_list):
    cores = os.cpu_count() or 1
    n_em = min(len(x_joint), CPU_EM_FRAMES)
    em_v, em_dt = cpu_em(x_joint, n_em, CPU_EM_ITERS)
    n_dtw = min(len(feats), 4 * cores)
    dtw_v, dtw_cells, dtw_dt = cpu_dtw(feats, n_dtw, cores)
    cv_v, cv_dt, cvv_v, cvv_dt, cv_procs = cpu_convert(model, src_list, 2)
    return {
        'em': {'value': em_v, 'unit': 'frames/s/iter', 'cores': cores, 'kind': 'reference',
               'sample': f'sklearn {__import__("sklearn").__version__} GaussianMixture.fit '
                         f'(the library the reference calls), first {n_em} frames, K={N_MIX_EM}, '
                         f'{CPU_EM_ITERS} iterations, labels from bench.lloyd_labels, BLAS '
                         f'threads = all cores, {em_dt:.1f} s'},
        'dtw': {'value': dtw_v, 'unit': 'cells/s', 'cores': cores, 'kind': 'port',
                'sample': f'C restatement of FastDTW (oracle/dtw_c.c), {n_dtw} pairs over '
                          f'{cores} processes, {dtw_dt:.2f} s'},
        'convert': {'value': cvv_v, 'unit': 'frames/s', 'cores': cv_procs, 'kind': 'port',
                    'sample': f'vectorised numpy MLPG restatement '
                              f'(oracle/mlpg_ref.transform_vectorised), {cv_procs} processes x 2 '
                              f'utterances x {UTT_FRAMES} frames, K={N_MIX_CONVERT}, {cvv_dt:.1f} s',
                    'faithful_per_frame': {
                        'value': cv_v, 'cores': 1,
                        'sample': f'per-frame restatement (oracle/mlpg_ref.transform), 2 '
                                  f'utterances, {cv_dt:.1f} s'}},
    }


def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores: sklearn's
    GaussianMixture.fit for EM (what kwiiyatta/converter/gmm.py:25-26 calls) with all host
    threads, and the oracle restatements of fastdtw / nnmnkwii MLPG (absent third-party
    packages) for the other two stages.  A step = one EM iteration over a bounded sample of the
    configs[1] workload: the first CPU_EM_FRAMES joint frames, same initialisation function as
    the GPU arm."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from kwiiyatta_b200 import synth
    from oracle import align_ref, delta_ref, dtw_c
    world = int(os.environ.get('WORLD_SIZE', str(args.gpus)))
    cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    # joint frames through the oracle chain until the sample is full
    chunks, feats, total, i = [], [], 0, 0
    while total < CPU_EM_FRAMES and i < N_PAIRS:
        (a, b), = build_dtw_inputs(i, 1)[0]
        xf, yf = align_ref.make_feature(a.mel_cepstrum.data, a.f0, a.is_voiced), \
            align_ref.make_feature(b.mel_cepstrum.data, b.f0, b.is_voiced)
        feats.append((xf, yf))
        _, path = dtw_c.fastdtw(xf, yf, radius=RADIUS, dist=2)
        p = align_ref.trim_even_path(align_ref.strict_filter(path, xf, yf), a.frame_len,
                                     b.frame_len, synth.PAD_LEN)
        src = delta_ref.delta_features(a.mel_cepstrum.data[p[0]][:, 1:])
        tgt = delta_ref.delta_features(b.mel_cepstrum.data[p[1]][:, 1:])
        chunks.append(delta_ref.remove_zeros_frames(np.hstack((src, tgt))))
        total += len(chunks[-1])
        i += 1
    x = np.concatenate(chunks)
    n_em = min(len(x), CPU_EM_FRAMES)
    if W > 0:
        cpu_em(x, n_em, W)
    em_v, em_dt = cpu_em(x, n_em, K)
    dtw_v, _, dtw_dt = cpu_dtw(feats, len(feats), cores)
    w, m, c = synth.make_joint_gmm(N_MIX_CONVERT, seed=0)
    srcs = [delta_ref.delta_features(s) for s in
            synth.make_source_utterances(2 * cores, frames=UTT_FRAMES)]
    cv_v, cv_dt, cvv_v, cvv_dt, cv_procs = cpu_convert((w, m, c), srcs, 2)
    import sklearn
    sample = (f'sklearn {sklearn.__version__} GaussianMixture.fit on the first {n_em} joint '
              f'frames of the corpus ({len(feats)} pairs), K={N_MIX_EM}, {K} iterations after '
              f'{W} warm-up iterations, labels from bench.lloyd_labels, BLAS threads = {cores}')
    out = {
        'impl': 'reference',
        'metric': 'GMM-EM frames/s/iter', 'value': em_v, 'unit': 'frames/s/iter',
        'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': em_dt / K * 1e3,
        'higher_is_better': True, 'scaling': 'strong' if args.total_pairs else 'weak',
        'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': 'configs[1] bounded sample: ' + sample,
                   'n_components': N_MIX_EM, 'dim': int(x.shape[1]), 'frames': int(n_em)},
        'cpu_baseline': {'value': em_v, 'unit': 'frames/s/iter', 'cores': cores,
                         'kind': 'reference', 'sample': sample},
        'e2e': {'value': em_v, 'unit': 'frames/s/iter', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'dtw_cells_per_s': dtw_v,
        'convert_frames_per_s': cvv_v,
        'stages': {
            'dtw': {'metric': 'DTW cells/s', 'value': dtw_v, 'unit': 'cells/s', 'kind': 'port',
                    'cores': cores,
                    'sample': f'oracle/dtw_c.c FastDTW r={RADIUS}, {len(feats)} pairs over '
                              f'{cores} processes, {dtw_dt:.2f} s'},
            'convert': {'metric': 'MLPG converted frames/s', 'value': cvv_v, 'unit': 'frames/s',
                        'kind': 'port', 'cores': cv_procs,
                        'sample': f'oracle/mlpg_ref.transform_vectorised, {cv_procs} processes x '
                                  f'2 x {UTT_FRAMES} frames, K={N_MIX_CONVERT}, {cvv_dt:.1f} s',
                        'faithful_per_frame': {'value': cv_v, 'cores': 1,
                                               'sample': f'oracle/mlpg_ref.transform, 2 x '
                                                         f'{UTT_FRAMES} frames, {cv_dt:.1f} s'}},
        },
    }
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default='tc', choices=['fp64', 'tc'],
                    help='tc = split-fp16 tcgen05 contractions (default), fp64 = CUDA-core DFMA')
    ap.add_argument('--pairs', type=int, default=N_PAIRS, help='pairs per GPU (weak scaling)')
    ap.add_argument('--total-pairs', type=int, default=0,
                    help='strong scaling: this many pairs split over the ranks (configs[2]: 5030)')
    ap.add_argument('--long-pairs', type=int, default=LONG_PAIRS,
                    help='configs[3] stage: pairs of 4096 x 4096 frames per GPU (0 = skip)')
    ap.add_argument('--utts', type=int, default=N_UTTS, help='conversion utterances per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    # rank 0 prints ONE JSON line on stdout: keep NCCL's own banner / debug output on stderr
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
    capture_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()

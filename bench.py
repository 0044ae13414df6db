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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PAIRS = 503
N_MIX_EM = 64
N_MIX_CONVERT = 128
N_UTTS = 1200
UTT_FRAMES = 600
RADIUS = 32


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
        """W warm-up + K timed steps; per-step CUDA events on the launching stream, L2 flushed
        between steps when asked; returns max-over-ranks total ms of the K steps."""
        for _ in range(W):
            fn()
        barrier()
        total = 0.0
        for _ in range(K):
            if flush:
                flush_buf.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(total)

    # ---------------- inputs (this rank's shard: pairs [rank*503, (rank+1)*503)) -----------
    n_pairs = args.pairs
    padded, feats = build_dtw_inputs(rank * n_pairs, n_pairs)
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
    # end to end through the fastdtw-compatible API: host features in, host paths out
    kfd.fastdtw_batch(feats, radius=RADIUS, dist=2, device=dev)     # warm-up (pins the staging)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    kfd.fastdtw_batch(feats, radius=RADIUS, dist=2, device=dev)
    torch.cuda.synchronize()
    dtw_e2e_s = time.perf_counter() - t0
    dtw_e2e = sum_over_ranks(cells_local) / max_over_ranks(dtw_e2e_s * 1e3) * 1e3

    # ---------------- joint frames for EM (align_even -> mcep -> delta -> hstack) ----------
    kw.set_pad_silence(lambda f, n: f)     # the synthetic features are already padded
    x_joint = kw.joint_array_from_pairs(padded, pad_silence=True, pad_len=synth.PAD_LEN)
    n_frames, dim = x_joint.shape
    # initial hard labels: a few Lloyd passes (the reference initialises with KMeans), computed
    # outside every timed region
    from kwiiyatta_b200 import kmeans
    labels0 = kmeans.kmeans_labels(torch.from_numpy(x_joint).to(dev), N_MIX_EM, seed=rank,
                                   n_lloyd=5).cpu().numpy()

    def make_gm(max_iter):
        resp0 = torch.zeros((n_frames, N_MIX_EM), dtype=torch.float64, device=dev)
        resp0[torch.arange(n_frames, device=dev), torch.from_numpy(labels0).to(dev)] = 1.0
        return kw.GaussianMixture(n_components=N_MIX_EM, max_iter=max_iter, tol=0.0,
                                  resp_init=resp0, device=dev, precision=args.precision)

    gm = make_gm(1)
    xj_dev = gm.initialize(x_joint)
    gm.em_iteration(xj_dev)
    gm._estep(torch, xj_dev)
    density = float((gm._resp[:, :n_frames] > 1e-16).sum().item()) / n_frames

    # ---------------- stage 2 (headline): EM iteration -------------------------------------
    em_ms = timed_loop(lambda: gm.em_iteration(xj_dev), flush=False)
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
    estep_ms = time_call(lambda: gm._estep(torch, xj_dev))
    mstep_ms = time_call(lambda: gm._accumulate(torch, xj_dev, gm._means[gm._cur]))
    flops_half = 2.0 * n_frames * N_MIX_EM * dim * dim
    # end to end through the converter back-end API with host buffers
    x_pinned = torch.from_numpy(x_joint).pin_memory()
    e2e_iters = 100     # the reference's own default (GMMFeatureConverter(max_iter=100)), tol = 0
    conv = kw.B200GMMFeatureConverter(components=N_MIX_EM, max_iter=e2e_iters, tol=0.0,
                                      verbose=0, device=dev, precision=args.precision)
    barrier()
    t0 = time.perf_counter()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        lab_dev = torch.from_numpy(labels0).to(dev)
        r0 = torch.zeros((n_frames, N_MIX_EM), dtype=torch.float64, device=dev)
        r0[torch.arange(n_frames, device=dev), lab_dev] = 1.0
        conv.gmm.resp_init = r0
        conv._train(x_pinned)
    torch.cuda.synchronize()
    em_e2e_s = time.perf_counter() - t0
    em_e2e_ms = max_over_ranks(em_e2e_s * 1e3)
    em_e2e = frames_total * e2e_iters / (em_e2e_ms / 1e3)
    model_bytes = sum(a.nbytes for a in (conv.gmm.weights_, conv.gmm.means_,
                                         conv.gmm.covariances_, conv.gmm.precisions_cholesky_))
    del conv, x_pinned, r0
    gm._resp = None
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
    paramgen.transform_many(src_list[:max(1, n_utts // 8)])     # warm-up (pins the staging)
    paramgen.transform_many(src_list)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    paramgen.transform_many(src_list)
    torch.cuda.synchronize()
    conv_e2e = conv_frames_total / (max_over_ranks((time.perf_counter() - t0) * 1e3) / 1e3)

    clock_info = clocks.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline (bounded samples, rank 0, N = 1 only) -------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baselines(feats, x_joint, labels0, (w, m, c), src_list, em_iters=2)

    em_tflops = 2 * flops_half / (em_ms / K / 1e3) / 1e12
    tc = args.precision == 'tc'
    names = ('estep_tc_kernel', 'mstats_tc_kernel') if tc else ('gmm_estep_kernel', 'gmm_mstats_kernel')
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
        'scaling': 'weak',
        'vs_baseline': None,
        'dtype': 'f64' if args.precision == 'fp64' else 'f16x2-split (fp32 accumulate) + f64',
        'data': 'synthetic',
        'config': {
            'workload': 'configs[1]: 503 synthetic ATR503-shaped pairs per GPU -> FastDTW r=32 '
                        '-> (N,144) joint frames -> 64-mix full-cov EM iteration; stages.convert '
                        'is configs[4] (128-mix, 720k frames)',
            'pairs_per_gpu': n_pairs, 'frames_per_gpu': int(n_frames), 'dim': int(dim),
            'n_components': N_MIX_EM, 'precision': args.precision,
            'init': 'hard labels from 5 Lloyd passes (KMeans-style, as the reference initialises)',
            'mean_components_per_frame_above_1e-16': density,
            'l2': 'EM inputs (X 8*N*144 B + resp) exceed L2; DTW/convert stages flush L2 with a '
                  '256 MiB write between timed steps',
        },
        'e2e': {
            'value': em_e2e, 'unit': 'frames/s/iter',
            'h2d_bytes_per_step': int(x_joint.nbytes + labels0.nbytes),
            'd2h_bytes_per_step': int(model_bytes),
            'note': f'B200GMMFeatureConverter._train on a pinned host (N,144) array, '
                    f'{e2e_iters} iterations incl. H2D of X, initial M-step and D2H of the model',
        },
        'gpu_launches': (10 if tc else 6) * K,
        'roofline': {
            'bound': 'tensor', 'kernel': dominant, 'achieved': dom_tflops, 'peak': peak,
            'unit': 'TFLOP/s', 'frac': dom_tflops / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel
            # in the ncu --set full capture of this workload (profiles/ncu_r1d_mstats_tc.txt,
            # profiles/ncu_r1c_estep_tc.txt); algorithmic: packed X 113 MB + resp 90 MB in
            'traffic': ({'mstats_tc_kernel': 204.4e6 + 75.6e6, 'estep_tc_kernel': 116.5e6 + 54.3e6}
                        .get(dominant) if n_frames == 176323 else None),
            'traffic_unit': 'bytes per launch (ncu, 1 GPU)',
            'peak_source': f"{peaks['source']} bf16_tflops_sustained",
            'algorithmic': '2*N*K*D^2 flop per launch (half of the 4*N*K*D^2 EM iteration); '
                           'duration = CUDA events around the C-ABI entry that launches it',
            'mma_passes': 3 if tc else None,
            'frac_of_peak_over_passes': (dom_tflops / (peak / 3)) if tc else None,
            'estep_ms': estep_ms, 'mstep_accumulate_ms': mstep_ms,
            'em_iteration_tflops': em_tflops,
        },
        'cpu_baseline': cpu['em'] if cpu else None,
        'clocks': clock_info,
        'stages': {
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
                        'd2h_bytes_per_step': int(8 * (tx.sum() + ty.sum()) + 20 * n_pairs)},
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
                        'd2h_bytes_per_step': int(n_utts * UTT_FRAMES * 24 * 8)},
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


def cpu_em(x, labels, n_sample, iters):
    from oracle import gmm_ref
    xs = np.ascontiguousarray(x[:n_sample])
    resp0 = np.zeros((len(xs), N_MIX_EM))
    resp0[np.arange(len(xs)), labels[:len(xs)]] = 1.0
    t0 = time.perf_counter()
    gmm_ref.sklearn_em(xs, resp0, max_iter=iters, tol=0.0)
    dt = time.perf_counter() - t0
    return len(xs) * iters / dt, dt


def cpu_convert(model, src_list, n_sample):
    from oracle import mlpg_ref
    w, m, c = model
    t0 = time.perf_counter()
    frames = 0
    for s in src_list[:n_sample]:
        mlpg_ref.transform(s, w, m, c)
        frames += len(s)
    dt = time.perf_counter() - t0
    return frames / dt, dt


def cpu_baselines(feats, x_joint, labels0, model, src_list, em_iters):
    cores = os.cpu_count() or 1
    n_em = min(len(x_joint), 12000)
    em_v, em_dt = cpu_em(x_joint, labels0, n_em, em_iters)
    n_dtw = min(len(feats), 4 * cores)
    dtw_v, dtw_cells, dtw_dt = cpu_dtw(feats, n_dtw, cores)
    cv_v, cv_dt = cpu_convert(model, src_list, 2)
    return {
        'em': {'value': em_v, 'unit': 'frames/s/iter', 'cores': cores, 'kind': 'reference',
               'sample': f'sklearn {__import__("sklearn").__version__} GaussianMixture.fit '
                         f'(the library the reference calls), first {n_em} frames, K={N_MIX_EM}, '
                         f'{em_iters} iterations, injected init, BLAS threads = all cores, '
                         f'{em_dt:.1f} s'},
        'dtw': {'value': dtw_v, 'unit': 'cells/s', 'cores': cores, 'kind': 'port',
                'sample': f'C restatement of FastDTW (oracle/dtw_c.c), {n_dtw} pairs over '
                          f'{cores} processes, {dtw_dt:.2f} s'},
        'convert': {'value': cv_v, 'unit': 'frames/s', 'cores': 1, 'kind': 'port',
                    'sample': f'faithful per-frame MLPG restatement (oracle/mlpg_ref.py), '
                              f'2 utterances x {UTT_FRAMES} frames, K={N_MIX_CONVERT}, '
                              f'{cv_dt:.1f} s'},
    }


def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores: sklearn's
    GaussianMixture.fit for EM (what kwiiyatta/converter/gmm.py:25-26 calls), and the oracle
    restatements of fastdtw / nnmnkwii MLPG (absent third-party packages) for the stages."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from kwiiyatta_b200 import synth
    from oracle import align_ref, delta_ref, dtw_c
    world = int(os.environ.get('WORLD_SIZE', str(args.gpus)))
    cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    n_pairs_sample = max(cores, 16)
    padded, feats = build_dtw_inputs(0, n_pairs_sample)
    # joint frames of the sample through the oracle chain
    chunks = []
    for (a, b), (xf, yf) in zip(padded, feats):
        _, path = dtw_c.fastdtw(xf, yf, radius=RADIUS, dist=2)
        p = align_ref.trim_even_path(align_ref.strict_filter(path, xf, yf), a.frame_len,
                                     b.frame_len, synth.PAD_LEN)
        src = delta_ref.delta_features(a.mel_cepstrum.data[p[0]][:, 1:])
        tgt = delta_ref.delta_features(b.mel_cepstrum.data[p[1]][:, 1:])
        chunks.append(delta_ref.remove_zeros_frames(np.hstack((src, tgt))))
    x = np.concatenate(chunks)
    n_em = min(len(x), 12000)
    labels0 = np.random.default_rng(0).integers(0, N_MIX_EM, len(x))
    if W > 0:
        cpu_em(x, labels0, n_em, 1)
    em_v, em_dt = cpu_em(x, labels0, n_em, K)
    dtw_v, _, dtw_dt = cpu_dtw(feats, len(feats), cores)
    w, m, c = synth.make_joint_gmm(N_MIX_CONVERT, seed=0)
    srcs = [delta_ref.delta_features(s) for s in
            synth.make_source_utterances(2, frames=UTT_FRAMES)]
    cv_v, cv_dt = cpu_convert((w, m, c), srcs, 2)
    import sklearn
    sample = (f'sklearn {sklearn.__version__} GaussianMixture.fit on the first {n_em} joint '
              f'frames of {n_pairs_sample} synthetic pairs, K={N_MIX_EM}, {K} iterations, '
              f'injected init, BLAS threads = all cores')
    out = {
        'impl': 'reference',
        'metric': 'GMM-EM frames/s/iter', 'value': em_v, 'unit': 'frames/s/iter',
        'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': em_dt / K * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': 'configs[1] bounded sample: ' + sample,
                   'n_components': N_MIX_EM, 'dim': int(x.shape[1])},
        'cpu_baseline': {'value': em_v, 'unit': 'frames/s/iter', 'cores': cores,
                         'kind': 'reference', 'sample': sample},
        'e2e': {'value': em_v, 'unit': 'frames/s/iter', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'stages': {
            'dtw': {'metric': 'DTW cells/s', 'value': dtw_v, 'unit': 'cells/s', 'kind': 'port',
                    'cores': cores,
                    'sample': f'oracle/dtw_c.c FastDTW r={RADIUS}, {len(feats)} pairs over '
                              f'{cores} processes, {dtw_dt:.2f} s'},
            'convert': {'metric': 'MLPG converted frames/s', 'value': cv_v, 'unit': 'frames/s',
                        'kind': 'port', 'cores': 1,
                        'sample': f'oracle/mlpg_ref.py per-frame restatement, 2 x {UTT_FRAMES} '
                                  f'frames, K={N_MIX_CONVERT}, {cv_dt:.1f} s'},
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
    ap.add_argument('--pairs', type=int, default=N_PAIRS, help='pairs per GPU')
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

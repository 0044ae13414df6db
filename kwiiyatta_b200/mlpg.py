"""Posterior + MLPG on the B200 (nnmnkwii.baseline.gmm.MLPG as used at
kwiiyatta/converter/gmm.py:28-34)."""
import numpy as np

from . import _lib
from .delta import DELTA_WINDOWS, check_windows


_on_own_device = _lib.device_guard(lambda self, *a: getattr(self, '_dev', None))


class MLPG:
    """``MLPG(gmm, windows, diff).transform(src)``.  ``gmm`` is anything with ``weights_``,
    ``means_`` and ``covariances_`` (our GaussianMixture or sklearn's).  The sliced model
    (marginal precision Cholesky, regression matrices, variance table) is prepared once
    here; the reference rebuilds it on every convert call (kwiiyatta/converter/gmm.py:32)."""

    def __init__(self, gmm, windows=None, swap=False, diff=False, precision='auto',
                 device=None):
        torch = _lib.require_cuda()
        self._dev = torch.device('cuda' if device is None else device)
        if self._dev.index is not None:
            with torch.cuda.device(self._dev):
                self._setup(torch, gmm, windows, swap, diff, precision)
        else:
            self._setup(torch, gmm, windows, swap, diff, precision)

    def _setup(self, torch, gmm, windows, swap, diff, precision):
        if windows is None:
            windows = DELTA_WINDOWS
        if len(windows) > 1:
            check_windows(windows)
        elif not (windows[0][0] == 0 and windows[0][1] == 0 and
                  np.array_equal(windows[0][2], [1.0])):
            raise NotImplementedError('a single window must be the static window (0, 0, [1.0])')
        if swap:
            raise NotImplementedError('swap=True is not built')
        if getattr(gmm, 'covariance_type', 'full') != 'full':
            raise AssertionError("covariance_type must be 'full'")
        self.windows = windows
        self.diff = bool(diff)
        dev = self._dev
        means = np.ascontiguousarray(gmm.means_, dtype=np.float64)
        self.num_mixtures, d = means.shape
        self.dim_half = d // 2
        self.static_dim = d // 2 // len(windows)
        # 'auto' = the tcgen05 posterior whenever the kernels hold the frame width: its hard
        # labels are re-checked in FP64 where the two best mixtures are close, so the mixture
        # sequence and the (FP64) MLPG output are those of the FP64 path
        self.precision = {'fp64': 0, 'tc': 1, 0: 0, 1: 1,
                          'auto': 1 if self.dim_half <= 144 else 0}[precision]
        lib = _lib.lib()
        k, dh = self.num_mixtures, self.dim_half
        w = torch.from_numpy(np.ascontiguousarray(gmm.weights_, dtype=np.float64)).to(dev)
        m = torch.from_numpy(means).to(dev)
        c = torch.from_numpy(np.ascontiguousarray(gmm.covariances_, dtype=np.float64)).to(dev)
        self._prepared = torch.empty(lib.kw_convert_prepared_len(k, dh), dtype=torch.float64,
                                     device=dev)
        info = torch.zeros(k, dtype=torch.int32, device=dev)
        rc = lib.kw_convert_prepare(k, dh, int(self.diff), w.data_ptr(), m.data_ptr(),
                                    c.data_ptr(), self._prepared.data_ptr(), info.data_ptr(),
                                    _lib.stream_ptr(torch))
        _lib.check(rc, 'kw_convert_prepare')
        if info.cpu().numpy().any():
            raise np.linalg.LinAlgError('source covariance of some mixture is not positive '
                                        'definite')

    @_on_own_device
    def transform_device(self, src_dev, offsets_dev, n_utts, max_frames, return_mix=False):
        """src_dev (sum T, dim_half) float64 CUDA tensor, offsets int64 (n_utts + 1)."""
        torch = _lib.require_cuda()
        lib = _lib.lib()
        total = src_dev.shape[0]
        out = torch.empty((total, self.static_dim), dtype=torch.float64, device=src_dev.device)
        mix = torch.empty(total, dtype=torch.int32, device=src_dev.device) if return_mix else None
        ws_bytes = lib.kw_convert_workspace_bytes(total, self.num_mixtures, self.dim_half,
                                                  self.precision)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=src_dev.device)
        rc = lib.kw_convert_batch(n_utts, offsets_dev.data_ptr(), total, int(max_frames),
                                  src_dev.data_ptr(), self.num_mixtures, self.dim_half,
                                  self._prepared.data_ptr(), out.data_ptr(), _lib.ptr(mix),
                                  self.precision, ws.data_ptr(), ws_bytes,
                                  _lib.stream_ptr(torch))
        _lib.check(rc, 'kw_convert_batch')
        return (out, mix) if return_mix else out

    CHUNK_FRAMES = 65536       # frames per pipeline chunk (36 MB in, 12 MB out at Dh = 72)

    @_on_own_device
    def transform_many(self, features):
        """Batched ``transform`` from a list of host arrays to a list of host arrays: one
        multi-threaded gather into pinned memory, ``transform_packed``, one scatter.  (With
        separate pageable arrays on both sides the host copies bound this call; callers that
        keep their frames in one pinned block use ``transform_packed`` directly.)"""
        torch = _lib.require_cuda()
        feats = [np.ascontiguousarray(f, dtype=np.float64) for f in features]
        for f in feats:
            if f.ndim != 2 or f.shape[1] != self.dim_half:
                raise ValueError(f'expected (T, {self.dim_half}) features, got {f.shape}')
        lens = np.array([len(f) for f in feats], dtype=np.int64)
        total = int(lens.sum())
        if total == 0:
            return [np.zeros((0, self.static_dim)) for _ in feats]
        off = np.concatenate(([0], np.cumsum(lens)))
        key = ('mlpg_many', str(self._dev))
        bufs = _lib._STAGING.get(key)
        if bufs is None or bufs[0].shape[0] < total or bufs[0].shape[1] != self.dim_half \
                or bufs[1].shape[1] != self.static_dim:
            bufs = (torch.empty((total, self.dim_half), dtype=torch.float64).pin_memory(),
                    torch.empty((total, self.static_dim), dtype=torch.float64).pin_memory())
            _lib._STAGING[key] = bufs
        hin, hout = bufs[0][:total], bufs[1][:total]
        host = hin.numpy()
        _lib.parallel_copy([(host[off[i]:off[i + 1]], feats[i]) for i in range(len(feats))])
        self.transform_packed(hin, lens, out=hout)
        res = hout.numpy()
        outputs = [np.empty((int(t), self.static_dim)) for t in lens]
        _lib.parallel_copy([(outputs[i], res[off[i]:off[i + 1]]) for i in range(len(feats))])
        return outputs

    @_on_own_device
    def transform_packed(self, src, lens, out=None):
        """Conversion of utterances packed row-wise in ONE host tensor ``src`` (sum T, dim_half)
        float64 -- pinned memory for full speed -- with lengths ``lens``; returns (and, when
        given, fills) the host tensor ``out`` (sum T, static_dim).  Runs as a three-stage
        pipeline over chunks of whole utterances: while the kernels of chunk i run, chunk i + 1
        is copied up and the result of chunk i - 1 is copied down (a copy-in and a copy-out
        stream beside the compute stream, two device buffers per direction); synchronised on
        return."""
        torch = _lib.require_cuda()
        dev, dh, sd = self._dev, self.dim_half, self.static_dim
        lens = np.asarray(lens, dtype=np.int64)
        total = int(lens.sum())
        if tuple(src.shape) != (total, dh) or src.dtype != torch.float64 or src.is_cuda:
            raise ValueError(f'src must be a host float64 tensor of shape ({total}, {dh})')
        if out is None:
            out = torch.empty((total, sd), dtype=torch.float64).pin_memory()
        elif tuple(out.shape) != (total, sd) or out.dtype != torch.float64 or out.is_cuda:
            raise ValueError(f'out must be a host float64 tensor of shape ({total}, {sd})')
        if total == 0:
            return out
        rows = np.concatenate(([0], np.cumsum(lens)))
        chunks, lo, acc = [], 0, 0
        for i, t in enumerate(lens):
            acc += int(t)
            if acc >= self.CHUNK_FRAMES or i == len(lens) - 1:
                chunks.append((lo, i + 1))
                lo, acc = i + 1, 0
        cap = max(int(rows[b] - rows[a]) for a, b in chunks)
        key = ('mlpg_pipe', str(dev), dh, sd)
        bufs = _lib._STAGING.get(key)
        if bufs is None or bufs['cap'] < cap:
            bufs = {'cap': cap,
                    'din': [torch.empty((cap, dh), dtype=torch.float64, device=dev)
                            for _ in range(2)],
                    'dout': [torch.empty((cap, sd), dtype=torch.float64, device=dev)
                             for _ in range(2)],
                    'streams': (torch.cuda.Stream(dev), torch.cuda.Stream(dev))}
            _lib._STAGING[key] = bufs
        s_in, s_out = bufs['streams']
        compute = torch.cuda.current_stream(dev)
        start = torch.cuda.Event()
        start.record(compute)
        s_in.wait_event(start)            # buffers may still be in use by the previous call
        s_out.wait_event(start)
        ev_comp, ev_d2h = {}, {}
        lib = _lib.lib()
        keep = []
        for c, (a, b) in enumerate(chunks):
            r0, r1 = int(rows[a]), int(rows[b])
            n_c = r1 - r0
            off = rows[a:b + 1] - r0
            with torch.cuda.stream(s_in):
                if c >= 2:
                    s_in.wait_event(ev_comp[c - 2])      # device input buffer consumed
                bufs['din'][c & 1][:n_c].copy_(src[r0:r1], non_blocking=True)
                off_dev = torch.from_numpy(off).to(dev, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            compute.wait_event(ev_in)
            if c >= 2:
                compute.wait_event(ev_d2h[c - 2])        # device output buffer copied down
            chunk_in, chunk_out = bufs['din'][c & 1][:n_c], bufs['dout'][c & 1][:n_c]
            ws_bytes = lib.kw_convert_workspace_bytes(n_c, self.num_mixtures, dh, self.precision)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            rc = lib.kw_convert_batch(b - a, off_dev.data_ptr(), n_c, int(lens[a:b].max()),
                                      chunk_in.data_ptr(), self.num_mixtures, dh,
                                      self._prepared.data_ptr(), chunk_out.data_ptr(), None,
                                      self.precision, ws.data_ptr(), ws_bytes,
                                      compute.cuda_stream)
            _lib.check(rc, 'kw_convert_batch')
            off_dev.record_stream(compute)
            keep.append(off)
            ev_comp[c] = torch.cuda.Event()
            ev_comp[c].record(compute)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_comp[c])
                out[r0:r1].copy_(chunk_out, non_blocking=True)
                ev_d2h[c] = torch.cuda.Event()
                ev_d2h[c].record(s_out)
        ev_d2h[len(chunks) - 1].synchronize()
        compute.wait_event(ev_d2h[len(chunks) - 1])
        return out

    @_on_own_device
    def transform_soft(self, src):
        """MLPGBase.transform: per-frame soft-posterior conditional mean, (T, dim_half)."""
        torch = _lib.require_cuda()
        lib = _lib.lib()
        src = np.ascontiguousarray(src, dtype=np.float64)
        if src.ndim != 2 or src.shape[1] != self.dim_half:
            raise ValueError(f'expected (T, {self.dim_half}) features, got {src.shape}')
        total = len(src)
        if total == 0:
            return np.zeros((0, self.dim_half))
        x = torch.from_numpy(src).to(self._dev)
        out = torch.empty((total, self.dim_half), dtype=torch.float64, device=self._dev)
        ws_bytes = lib.kw_convert_soft_workspace_bytes(total, self.num_mixtures, self.dim_half)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=self._dev)
        rc = lib.kw_convert_soft_batch(total, x.data_ptr(), self.num_mixtures, self.dim_half,
                                       self._prepared.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                       ws_bytes, _lib.stream_ptr(torch))
        _lib.check(rc, 'kw_convert_soft_batch')
        return out.cpu().numpy()

    def transform(self, src):
        src = np.asarray(src, dtype=np.float64)
        if src.ndim != 2:
            raise ValueError('MLPG.transform expects a (T, dim) array')
        if src.shape[1] == self.static_dim:
            # nnmnkwii: feature_dim == static_dim -> MLPGBase.transform (no trajectory smoothing)
            return self.transform_soft(src)
        return self.transform_many([src])[0]

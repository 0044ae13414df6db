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
        """Batched ``transform`` from host arrays to host arrays.  Large batches run as a
        three-stage pipeline over chunks of utterances: while the kernels of chunk i run, chunk
        i + 1 is gathered into pinned memory and copied up and the result of chunk i - 1 is
        copied down and scattered to the output arrays (two pinned buffers per direction, a
        copy-in and a copy-out stream beside the compute stream)."""
        torch = _lib.require_cuda()
        feats = [np.ascontiguousarray(f, dtype=np.float64) for f in features]
        for f in feats:
            if f.ndim != 2 or f.shape[1] != self.dim_half:
                raise ValueError(f'expected (T, {self.dim_half}) features, got {f.shape}')
        lens = np.array([len(f) for f in feats], dtype=np.int64)
        if lens.sum() == 0:
            return [np.zeros((0, self.static_dim)) for _ in feats]
        if lens.sum() <= 2 * self.CHUNK_FRAMES:
            off = np.concatenate(([0], np.cumsum(lens)))
            src = _lib.gather_to_device(torch, feats, self._dev, 'mlpg_in')
            off_dev = torch.from_numpy(off).to(self._dev, non_blocking=True)
            out = self.transform_device(src, off_dev, len(feats), int(lens.max()))
            return _lib.scatter_to_host(torch, out, off, 'mlpg_out')
        return self._transform_pipelined(torch, feats, lens)

    def _transform_pipelined(self, torch, feats, lens):
        dev, dh, sd = self._dev, self.dim_half, self.static_dim
        # chunks of whole utterances
        chunks, lo, acc = [], 0, 0
        for i, t in enumerate(lens):
            acc += int(t)
            if acc >= self.CHUNK_FRAMES or i == len(lens) - 1:
                chunks.append((lo, i + 1))
                lo, acc = i + 1, 0
        cap = max(int(lens[a:b].sum()) for a, b in chunks)
        key = ('mlpg_pipe', str(dev), dh, sd)
        bufs = _lib._STAGING.get(key)
        if bufs is None or bufs['cap'] < cap:
            bufs = {'cap': cap,
                    'hin': [torch.empty((cap, dh), dtype=torch.float64).pin_memory()
                            for _ in range(2)],
                    'hout': [torch.empty((cap, sd), dtype=torch.float64).pin_memory()
                             for _ in range(2)],
                    'din': [torch.empty((cap, dh), dtype=torch.float64, device=dev)
                            for _ in range(2)],
                    'dout': [torch.empty((cap, sd), dtype=torch.float64, device=dev)
                             for _ in range(2)],
                    'streams': (torch.cuda.Stream(dev), torch.cuda.Stream(dev))}
            _lib._STAGING[key] = bufs
        s_in, s_out = bufs['streams']
        compute = torch.cuda.current_stream(dev)
        outputs = [None] * len(feats)
        ev_h2d, ev_comp, ev_d2h = {}, {}, {}
        lib = _lib.lib()

        def scatter(c):
            a, b = chunks[c]
            ev_d2h[c].synchronize()
            host = bufs['hout'][c & 1].numpy()
            off = np.concatenate(([0], np.cumsum(lens[a:b])))
            jobs = []
            for j in range(a, b):
                outputs[j] = np.empty((int(lens[j]), sd))
                jobs.append((outputs[j], host[off[j - a]:off[j - a + 1]]))
            _lib.parallel_copy(jobs)

        for c, (a, b) in enumerate(chunks):
            n_c = int(lens[a:b].sum())
            off = np.concatenate(([0], np.cumsum(lens[a:b])))
            if c >= 2:
                ev_h2d[c - 2].synchronize()          # pinned input buffer free again
            host = bufs['hin'][c & 1].numpy()
            _lib.parallel_copy([(host[off[j - a]:off[j - a + 1]], feats[j]) for j in range(a, b)])
            with torch.cuda.stream(s_in):
                if c >= 2:
                    s_in.wait_event(ev_comp[c - 2])  # device input buffer consumed
                bufs['din'][c & 1][:n_c].copy_(bufs['hin'][c & 1][:n_c], non_blocking=True)
                off_dev = torch.from_numpy(off).to(dev, non_blocking=True)
                ev_h2d[c] = torch.cuda.Event()
                ev_h2d[c].record(s_in)
            compute.wait_event(ev_h2d[c])
            if c >= 2:
                compute.wait_event(ev_d2h[c - 2])    # device output buffer copied down
            src = bufs['din'][c & 1][:n_c]
            out = bufs['dout'][c & 1][:n_c]
            ws_bytes = lib.kw_convert_workspace_bytes(n_c, self.num_mixtures, dh, self.precision)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            rc = lib.kw_convert_batch(b - a, off_dev.data_ptr(), n_c, int(lens[a:b].max()),
                                      src.data_ptr(), self.num_mixtures, dh,
                                      self._prepared.data_ptr(), out.data_ptr(), None,
                                      self.precision, ws.data_ptr(), ws_bytes,
                                      compute.cuda_stream)
            _lib.check(rc, 'kw_convert_batch')
            off_dev.record_stream(compute)
            ev_comp[c] = torch.cuda.Event()
            ev_comp[c].record(compute)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_comp[c])
                if c >= 2:
                    pass                              # hout[c & 1] was scattered before (below)
                bufs['hout'][c & 1][:n_c].copy_(out, non_blocking=True)
                ev_d2h[c] = torch.cuda.Event()
                ev_d2h[c].record(s_out)
            if c >= 1:
                scatter(c - 1)                        # overlaps the kernels of chunk c
        scatter(len(chunks) - 1)
        return outputs

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

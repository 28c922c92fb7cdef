"""Whole-cloud MuPS extraction with host buffers: the call a user of the reference's inference
loop makes (test_n_est_w_experts.py:129-148 feeds host arrays per batch; here one cloud at a
time): host cloud in -> index build -> per chunk of query points both halves on the GPU -> MuPS
chunks streamed back into pinned host memory, copies overlapped with compute on a second stream.
"""
import numpy as np
import torch

from . import mups as _m


class MuPSPipeline(object):
    def __init__(self, gmm, patch_radius, points_per_patch, seed=3627473, chunk=8192, device=None):
        """gmm: GridGMM-like object (weights_, means_, covariances_) or a GMMHandle."""
        _m._require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if isinstance(gmm, _m.GMMHandle):
            self.gmm = gmm
        else:
            self.gmm = _m.gmm_handle(gmm.weights_, gmm.means_, np.sqrt(gmm.covariances_), self.device)
        self.patch_radius = list(patch_radius)
        self.P = int(points_per_patch)
        self.seed = seed
        self.chunk = int(chunk)
        G, S = self.gmm.G, len(self.patch_radius)
        res = int(round(G ** (1.0 / 3.0)))
        self.feat_shape = (res, res, res, 20 * S) if res ** 3 == G else (G, 20 * S)
        self.row_floats = 20 * S * G
        self._dev = [torch.empty((self.chunk, self.row_floats), dtype=torch.float32, device=self.device) for _ in range(2)]
        self._host = [torch.empty((self.chunk, self.row_floats), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(self.device)
        self._copied = [torch.cuda.Event() for _ in range(2)]
        self._ready = [torch.cuda.Event() for _ in range(2)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def features_on_device(self, pts, query_idx=None, out=None):
        """MuPS [B, res, res, res, 20*S] on the GPU for one cloud (pts: host or device [N,3])."""
        index = _m.PointIndex(pts, cell_frac=max(self.patch_radius), device=self.device)
        radii = index.absolute_radii(self.patch_radius)
        if query_idx is None:
            query_idx = torch.arange(index.n, dtype=torch.int64, device=self.device)
        return _m.mups_features(index, self.gmm, query_idx, radii, self.P, seed=self.seed, out=out)

    def _stage_in(self, pts_host, query_idx_host):
        """Host cloud (and query list) to the device, index build, radii: the front end shared by both streaming calls."""
        dev = self.device
        pts_host = torch.as_tensor(pts_host, dtype=torch.float32)
        xyz = pts_host.to(dev, non_blocking=True)
        self.h2d_bytes += pts_host.numel() * 4
        index = _m.PointIndex(xyz, cell_frac=max(self.patch_radius), device=dev)
        radii = index.absolute_radii(self.patch_radius)          # the one small synchronising read (bbox)
        if query_idx_host is None:
            q = torch.arange(index.n, dtype=torch.int64, device=dev)
        else:
            qh = torch.as_tensor(query_idx_host, dtype=torch.int64)
            q = qh.to(dev, non_blocking=True)
            self.h2d_bytes += qh.numel() * 8
        return index, radii, q

    def features_to_consumer(self, pts_host, query_idx_host, consume):
        """Host cloud in, MuPS consumed ON THE DEVICE chunk by chunk -- the flow of the reference's inference
        loop, where MuPS is an intermediate of the graph that feeds the 3D-CNN and never visits the host
        (models/experts_n_est.py:59-108).  `consume(lo, hi, rows)` gets a CUDA view [hi-lo, 20*S*G] on the current
        stream (valid until the next chunk is enqueued on that stream).  Returns the number of query points."""
        index, radii, q = self._stage_in(pts_host, query_idx_host)
        B = int(q.shape[0])
        for c, lo in enumerate(range(0, B, self.chunk)):
            hi = min(B, lo + self.chunk)
            dbuf = self._dev[c & 1][: hi - lo]
            _m.mups_features(index, self.gmm, q[lo:hi], radii, self.P, seed=self.seed, out=dbuf)
            consume(lo, hi, dbuf)
        return B

    def features_to_host(self, pts_host, query_idx_host=None, consume=None):
        """Streams the MuPS rows of every query of one cloud to the host.  `consume(lo, hi, rows)`
        is called with a pinned numpy view [hi-lo, 20*S*G] per chunk, in ascending order of `lo` (valid only during the call).
        Returns the number of query points processed."""
        compute = torch.cuda.current_stream(self.device)
        index, radii, q = self._stage_in(pts_host, query_idx_host)
        B = int(q.shape[0])
        pending = [None, None]

        def drain(k):
            if pending[k] is not None:                            # wait for the copy out of buffer k, hand it over
                self._copied[k].synchronize()
                if consume is not None:
                    plo, phi = pending[k]
                    consume(plo, phi, self._host[k][: phi - plo].numpy())
                pending[k] = None

        for c, lo in enumerate(range(0, B, self.chunk)):
            hi = min(B, lo + self.chunk)
            k = c & 1
            drain(k)                                              # chunk c-2 has left _dev[k] / _host[k]; c-1 is computing
            dbuf = self._dev[k][: hi - lo]
            _m.mups_features(index, self.gmm, q[lo:hi], radii, self.P, seed=self.seed, out=dbuf)
            self._ready[k].record(compute)
            self._copy_stream.wait_event(self._ready[k])
            with torch.cuda.stream(self._copy_stream):
                self._host[k][: hi - lo].copy_(dbuf, non_blocking=True)
                self._copied[k].record(self._copy_stream)
            pending[k] = (lo, hi)
            self.d2h_bytes += (hi - lo) * self.row_floats * 4
        last = (len(range(0, B, self.chunk)) - 1) & 1      # buffer of the newest chunk: the other one holds the older chunk
        drain(last ^ 1)
        drain(last)
        return B

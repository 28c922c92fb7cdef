"""Host side of the MuPS hot path: thin Python over the C ABI (include/mups.h).

torch is used for device memory, streams and (in ``dist.py``) torch.distributed only; every
computation is a kernel of libmups_b200.so.  Names follow the reference's domain: clouds,
patches, scales, Gaussians.
"""
from __future__ import annotations

import ctypes
import threading

import numpy as np
import torch

from . import _lib

_F32P = ctypes.POINTER(ctypes.c_float)
_F64P = ctypes.POINTER(ctypes.c_double)


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("MuPS needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _as_device(x, dtype, device):
    """numpy / torch (any device) -> contiguous torch tensor of `dtype` on `device`."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(device, non_blocking=True)


# --------------------------------------------------------------------------------------------------
# GMM (reference utils/utils.py:70-95 get_3d_grid_gmm; fed as w, mu, sqrt(cov) in float32,
# train_n_est_w_experts.py:284-286)
# --------------------------------------------------------------------------------------------------

class GridGMM(object):
    """The three arrays the reference keeps in an sklearn GaussianMixture used as a struct."""

    def __init__(self, weights, means, covariances):
        self.weights_ = weights
        self.means_ = means
        self.covariances_ = covariances
        self.n_components = len(weights)
        self.covariance_type = "diag"


def get_3d_grid_gmm(subdivisions=[5, 5, 5], variance=0.04):
    """Same signature and values as reference utils/utils.py:70-95: Gaussians on a regular
    lattice over [-1, 1]^3 (cell centres), isotropic covariance `variance`, uniform weights."""
    nx, ny, nz = (int(v) for v in subdivisions)
    n_gaussians = nx * ny * nz
    # np.mgrid[a:b:n*1j] evaluates arange(n) * ((b - a) / (n - 1)) + a; reproduce that rounding
    axes = []
    for n in (nx, ny, nz):
        lo, hi = 1.0 / n - 1.0, 1.0 - 1.0 / n
        axes.append(np.arange(n, dtype=np.float64) * ((hi - lo) / float(n - 1)) + lo if n > 1 else np.array([lo]))
    gx, gy, gz = np.meshgrid(*axes, indexing="ij")          # x slowest, z fastest (np.mgrid order)
    means = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)
    covariances = variance * np.ones_like(means)
    weights = np.full(n_gaussians, 1.0 / n_gaussians)
    return GridGMM(weights, means, covariances)


class GMMHandle(object):
    """Device copy of (w, mu, sigma) + derived constants (mups_gmm_*)."""

    def __init__(self, w, mu, sigma, device=None):
        _require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        w = np.ascontiguousarray(np.asarray(w, dtype=np.float32).reshape(-1))
        mu = np.ascontiguousarray(np.asarray(mu, dtype=np.float32))
        sigma = np.ascontiguousarray(np.asarray(sigma, dtype=np.float32))
        if mu.ndim != 2 or mu.shape[1] != 3 or sigma.shape != mu.shape or w.shape[0] != mu.shape[0]:
            raise ValueError("GMM shapes: w [G], mu [G,3], sigma [G,3]; got %s %s %s" % (w.shape, mu.shape, sigma.shape))
        self.G = int(w.shape[0])
        self._h = ctypes.c_void_p()
        L = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(L.mups_gmm_create(ctypes.byref(self._h), w.ctypes.data_as(_F32P), mu.ctypes.data_as(_F32P),
                                         sigma.ctypes.data_as(_F32P), self.G), "mups_gmm_create")
        self.separable = bool(L.mups_gmm_is_separable(self._h))

    @property
    def handle(self):
        return self._h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().mups_gmm_destroy(h)
            except Exception:
                pass


_gmm_cache = {}
_gmm_lock = threading.Lock()


def gmm_handle(w, mu, sigma, device=None):
    """Cached GMMHandle for constant (w, mu, sigma) -- the reference re-feeds the same arrays on
    every sess.run (test_n_est_w_experts.py:142-147)."""
    _require_cuda()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def host(x):
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        return np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    w, mu, sigma = host(w).reshape(-1), host(mu), host(sigma)
    key = (str(device), w.tobytes(), mu.tobytes(), sigma.tobytes())
    with _gmm_lock:
        h = _gmm_cache.get(key)
        if h is None:
            if len(_gmm_cache) > 16:
                _gmm_cache.clear()
            h = _gmm_cache[key] = GMMHandle(w, mu, sigma, device)
    return h


# --------------------------------------------------------------------------------------------------
# Spatial index (replaces load_shape -> cKDTree, reference utils/pcpnet_dataset.py:13-39)
# --------------------------------------------------------------------------------------------------

def grid_cell_scale(n_points):
    """Grid cell edge as a multiple of the largest query radius.  One cell per radius (27 candidate cells, flat scan)
    is the fastest for PCPNet-size clouds.  Dense clouds get fine cells (1/8 of the radius up to 4 M points, 1/16
    above): the library then answers with the hierarchical kernel, which accepts cells wholly inside a ball without
    looking at their points and tests only the cells that straddle a sphere (profiles/r02_ball_query_dense.md).
    Results never depend on the grid."""
    if n_points >= 4000000:
        return 1.0 / 16.0
    if n_points >= 400000:
        return 0.125
    return 1.0


class PointIndex(object):
    """Uniform-grid spatial hash of one cloud on the GPU (mups_index_*).  ``cell_frac`` is the largest query
    radius as a fraction of the bounding-box diagonal; the cell edge is ``cell_frac * cell_scale`` of the diagonal
    (``cell_scale=None``: chosen from the cloud size, see grid_cell_scale)."""

    def __init__(self, pts, cell_frac=0.07, device=None, cell_scale=None):
        _require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        xyz = _as_device(pts, torch.float32, self.device)
        if xyz.ndim != 2 or xyz.shape[1] != 3 or xyz.shape[0] < 1:
            raise ValueError("points must be [N,3] with N >= 1, got %s" % (tuple(xyz.shape),))
        self.n = int(xyz.shape[0])
        self._h = ctypes.c_void_p()
        L = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(L.mups_index_create(ctypes.byref(self._h), ctypes.c_void_p(xyz.data_ptr()), self.n,
                                           float(cell_frac) * float(grid_cell_scale(self.n) if cell_scale is None else cell_scale),
                                           _stream_ptr(self.device)), "mups_index_create")
        # the library copies the cloud during the (asynchronous) build on this stream
        xyz.record_stream(torch.cuda.current_stream(self.device))
        self._bbox = None

    @property
    def handle(self):
        return self._h

    def bbox(self):
        """(min [3], max [3]) float32 -- pts.min(0), pts.max(0)."""
        if self._bbox is None:
            mn = np.empty(3, np.float32)
            mx = np.empty(3, np.float32)
            with torch.cuda.device(self.device):
                _lib.check(_lib.load().mups_index_bbox(self._h, mn.ctypes.data_as(_F32P), mx.ctypes.data_as(_F32P)),
                           "mups_index_bbox")
            self._bbox = (mn, mx)
        return self._bbox

    def bbdiag(self):
        """float(np.linalg.norm(pts.max(0) - pts.min(0), 2)) in float32 (pcpnet_dataset.py:281)."""
        mn, mx = self.bbox()
        return float(np.linalg.norm(mx - mn, 2))

    def absolute_radii(self, patch_radius):
        """[bbdiag * rad for rad in patch_radius] (pcpnet_dataset.py:282), Python floats."""
        bbdiag = self.bbdiag()
        return [bbdiag * float(rad) for rad in patch_radius]

    def ball_query(self, query_idx, radii_abs, points_per_patch, seed=3627473, return_indices=False,
                   return_patches=True):
        """Half 1 for a batch of centre indices (pcpnet_dataset.py:286-343, center='point').

        Returns (patches [B,S*P,3] f32, n_eff [B,S] i32, nbr_total [B,S] i32[, nbr_idx [B,S,P] i32]),
        all CUDA tensors on the current stream."""
        dev = self.device
        q = _as_device(query_idx, torch.int64, dev).reshape(-1)
        B, S, P = int(q.shape[0]), len(radii_abs), int(points_per_patch)
        r = np.ascontiguousarray(np.asarray(radii_abs, dtype=np.float64))
        patches = torch.empty((B, S * P, 3), dtype=torch.float32, device=dev) if return_patches else None
        n_eff = torch.empty((B, S), dtype=torch.int32, device=dev)
        total = torch.empty((B, S), dtype=torch.int32, device=dev)
        nbr = torch.empty((B, S, P), dtype=torch.int32, device=dev) if return_indices else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().mups_ball_query(
                self._h, ctypes.c_void_p(q.data_ptr()), B, r.ctypes.data_as(_F64P), S, P, int(seed) & (2 ** 64 - 1),
                ctypes.c_void_p(nbr.data_ptr() if nbr is not None else 0), ctypes.c_void_p(total.data_ptr()),
                ctypes.c_void_p(patches.data_ptr() if patches is not None else 0), ctypes.c_void_p(n_eff.data_ptr()),
                _stream_ptr(dev)), "mups_ball_query")
        if return_indices:
            return patches, n_eff, total, nbr
        return patches, n_eff, total

    def select(self, query_idx, radii_abs, points_per_patch, seed=3627473):
        """Half 1 without the patch tensor (mups_ball_query_select): (nbr_pos [B,S,P] i32 -- opaque positions for
        ``stats_3dmfv_selected`` --, n_eff [B,S] i32, nbr_total [B,S] i32)."""
        dev = self.device
        q = _as_device(query_idx, torch.int64, dev).reshape(-1)
        B, S, P = int(q.shape[0]), len(radii_abs), int(points_per_patch)
        r = np.ascontiguousarray(np.asarray(radii_abs, dtype=np.float64))
        pos = torch.empty((B, S, P), dtype=torch.int32, device=dev)
        n_eff = torch.empty((B, S), dtype=torch.int32, device=dev)
        total = torch.empty((B, S), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().mups_ball_query_select(
                self._h, ctypes.c_void_p(q.data_ptr()), B, r.ctypes.data_as(_F64P), S, P, int(seed) & (2 ** 64 - 1),
                ctypes.c_void_p(pos.data_ptr()), ctypes.c_void_p(total.data_ptr()), ctypes.c_void_p(n_eff.data_ptr()),
                _stream_ptr(dev)), "mups_ball_query_select")
        return pos, n_eff, total

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().mups_index_destroy(h)
            except Exception:
                pass


# --------------------------------------------------------------------------------------------------
# Half 2
# --------------------------------------------------------------------------------------------------

def stats_3dmfv(patches, n_eff, gmm, n_scales, masked=True, layout="mups", out=None, fastpath=True, wide_stores=False):
    """3DmFV statistics of patches [B, S*P, 3] for all S scales in one launch (mups_3dmfv).

    layout 'mups'    -> [B, res, res, res, 20*S] (models/experts_n_est.py:71-76; [B, G, 20*S] when
                        G is not a cube)
    layout 'channel' -> [B, S, 20, G]            (per scale the flatten=True / flatten=False memory order)
    wide_stores: write each (query, scale) result as 80-byte runs (for `out` in a peer GPU's memory, dist.PeerSlabGather)
    """
    dev = gmm.device
    pts = _as_device(patches, torch.float32, dev)
    if pts.ndim != 3 or pts.shape[2] != 3:
        raise ValueError("patches must be [B, S*P, 3], got %s" % (tuple(pts.shape),))
    B, S = int(pts.shape[0]), int(n_scales)
    if S < 1 or pts.shape[1] % S:
        raise ValueError("patch length %d is not a multiple of the %d scales" % (pts.shape[1], S))
    P = int(pts.shape[1]) // S
    G = gmm.G
    flags = 0
    ne = None
    if masked:
        if n_eff is None:
            raise ValueError("n_original_points is required (the reference fails on None, tf_util.py:665)")
        ne = _as_device(n_eff, torch.int32, dev).reshape(B, S)
        flags |= _lib.FLAG_MASKED
    if not fastpath:
        flags |= _lib.FLAG_NO_FASTPATH
    if wide_stores:
        flags |= _lib.FLAG_WIDE_STORES
    if layout == "mups":
        res = int(round(G ** (1.0 / 3.0)))
        shape = (B, res, res, res, 20 * S) if res ** 3 == G else (B, G, 20 * S)
    elif layout == "channel":
        flags |= _lib.LAYOUT_CHANNEL
        shape = (B, S, 20, G)
    else:
        raise ValueError("Unknown layout: %s" % (layout,))
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=dev)
    elif out.numel() != B * S * 20 * G or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
        raise ValueError("out must be a contiguous float32 CUDA tensor of %d elements" % (B * S * 20 * G))
    with torch.cuda.device(dev):
        _lib.check(_lib.load().mups_3dmfv(
            gmm.handle, ctypes.c_void_p(pts.data_ptr()), ctypes.c_void_p(ne.data_ptr() if ne is not None else 0),
            B, S, P, flags, ctypes.c_void_p(out.data_ptr()), _stream_ptr(dev)), "mups_3dmfv")
    return out


def stats_3dmfv_selected(index, gmm, query_idx, radii_abs, nbr_pos, n_eff, out=None, fastpath=True):
    """3DmFV statistics of the patches a selection (PointIndex.select) describes, gathered from the index by the
    statistics kernel itself (mups_3dmfv_selected): MuPS [B,res,res,res,20*S]."""
    dev = index.device
    q = _as_device(query_idx, torch.int64, dev).reshape(-1)
    B, S, P, G = int(nbr_pos.shape[0]), int(nbr_pos.shape[1]), int(nbr_pos.shape[2]), gmm.G
    r = np.ascontiguousarray(np.asarray(radii_abs, dtype=np.float64))
    res = int(round(G ** (1.0 / 3.0)))
    shape = (B, res, res, res, 20 * S) if res ** 3 == G else (B, G, 20 * S)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=dev)
    elif out.numel() != B * S * 20 * G or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 CUDA tensor of %d elements" % (B * S * 20 * G))
    with torch.cuda.device(dev):
        _lib.check(_lib.load().mups_3dmfv_selected(
            gmm.handle, index.handle, ctypes.c_void_p(q.data_ptr()), B, r.ctypes.data_as(_F64P), S, P,
            ctypes.c_void_p(nbr_pos.data_ptr()), ctypes.c_void_p(n_eff.data_ptr()), 0 if fastpath else _lib.FLAG_NO_FASTPATH,
            ctypes.c_void_p(out.data_ptr()), _stream_ptr(dev)), "mups_3dmfv_selected")
    return out


def mups_features(index, gmm, query_idx, radii_abs, points_per_patch, seed=3627473, out=None,
                  return_patches=False, fastpath=True):
    """Both halves for a batch of centres: MuPS [B,res,res,res,20*S] (mups_features).  Without ``return_patches`` no
    patch tensor exists anywhere (K6: the ball query hands the statistics kernel the positions of the selected
    neighbours); the features are the same bit for bit."""
    dev = index.device
    q = _as_device(query_idx, torch.int64, dev).reshape(-1)
    B, S, P, G = int(q.shape[0]), len(radii_abs), int(points_per_patch), gmm.G
    r = np.ascontiguousarray(np.asarray(radii_abs, dtype=np.float64))
    patches = torch.empty((B, S * P, 3), dtype=torch.float32, device=dev) if return_patches else None
    n_eff = torch.empty((B, S), dtype=torch.int32, device=dev)
    total = torch.empty((B, S), dtype=torch.int32, device=dev)
    res = int(round(G ** (1.0 / 3.0)))
    shape = (B, res, res, res, 20 * S) if res ** 3 == G else (B, G, 20 * S)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=dev)
    elif out.numel() != B * S * 20 * G or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 CUDA tensor of %d elements" % (B * S * 20 * G))
    flags = 0 if fastpath else _lib.FLAG_NO_FASTPATH
    with torch.cuda.device(dev):
        _lib.check(_lib.load().mups_features(
            index.handle, gmm.handle, ctypes.c_void_p(q.data_ptr()), B, r.ctypes.data_as(_F64P), S, P,
            int(seed) & (2 ** 64 - 1), flags, ctypes.c_void_p(patches.data_ptr() if patches is not None else 0),
            ctypes.c_void_p(n_eff.data_ptr()), ctypes.c_void_p(total.data_ptr()), ctypes.c_void_p(out.data_ptr()),
            _stream_ptr(dev)), "mups_features")
    if return_patches:
        return out, patches, n_eff, total
    return out

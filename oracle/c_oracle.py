"""ctypes loader for oracle/libmups_oracle.so (the plain-C port of half 2).

TEST INFRASTRUCTURE ONLY (see oracle/mups_oracle.c).  Built by
``__graft_entry__.build()`` / ``make -C oracle``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_TUNED = None

# constants of the fp32 forward-error model of oracle_mups_f64: per-term relative error
# eps (C0 + C1 (ss + ss_min)), accumulation eps CSUM sqrt(m) sum|term|
BOUND_C0, BOUND_C1, BOUND_CSUM = 8.0, 4.0, 1.0


def _stale(so, srcs):
    return not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs)


def build(force=False):
    so = os.path.join(_HERE, "libmups_oracle.so")
    mk = os.path.join(_HERE, "Makefile")
    if force or _stale(so, [os.path.join(_HERE, "mups_oracle.c"), mk]):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmups_oracle.so"], stdout=subprocess.DEVNULL)
    tuned = os.path.join(_HERE, "libmups_oracle_tuned.so")
    if force or _stale(tuned, [os.path.join(_HERE, "mups_oracle_tuned.c"), mk]):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmups_oracle_tuned.so"], stdout=subprocess.DEVNULL)
    return so


def build_tuned_native():
    """Rebuild the tuned CPU implementation with -march=native on the host that is about to time it (bench.py's CPU
    legs); keeps the shipped x86-64-v3 build when the compiler is missing or fails.  Returns the -march used."""
    global _TUNED
    native = os.path.join(_HERE, "libmups_oracle_tuned_native.so")
    try:
        subprocess.check_call(["/usr/bin/gcc", "-O3", "-march=native", "-ffast-math", "-fopenmp", "-fPIC", "-shared", "-o",
                               native, os.path.join(_HERE, "mups_oracle_tuned.c"), "-lm"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _TUNED = _load_tuned(native)
        return "native"
    except Exception:
        _TUNED = None
        return "x86-64-v3"


def _load_tuned(path):
    L = ctypes.CDLL(path)
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int32)
    L.oracle_mups_tuned.argtypes = [fp, ip, fp, fp, fp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
    L.oracle_mups_tuned.restype = ctypes.c_int
    L.oracle_tuned_set_num_threads.argtypes = [ctypes.c_int]
    L.oracle_tuned_set_num_threads.restype = None
    return L


def tuned_lib():
    global _TUNED
    if _TUNED is None:
        so = os.path.join(_HERE, "libmups_oracle_tuned.so")
        if not os.path.exists(so):
            build()
        _TUNED = _load_tuned(so)
    return _TUNED


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libmups_oracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.oracle_3dmfv.argtypes = [fp, ip, fp, fp, fp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
        L.oracle_3dmfv.restype = ctypes.c_int
        L.oracle_mups.argtypes = [fp, ip, fp, fp, fp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
        L.oracle_mups.restype = ctypes.c_int
        L.oracle_selection_keys.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ip, ctypes.c_int64,
                                            ctypes.POINTER(ctypes.c_uint32)]
        L.oracle_selection_keys.restype = None
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_mups_f64.argtypes = [fp, ip, fp, fp, fp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_double, ctypes.c_double, ctypes.c_double, dp, dp]
        L.oracle_mups_f64.restype = ctypes.c_int
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.oracle_half1_gather.argtypes = [fp, i64p, ctypes.c_int64, ctypes.c_float, i64p, i64p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_uint64, fp, ip]
        L.oracle_half1_gather.restype = ctypes.c_int
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        L.oracle_set_num_threads.restype = None
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core."""
    lib().oracle_set_num_threads(int(n))
    tuned_lib().oracle_tuned_set_num_threads(int(n))


def get_3dmfv(points, w, mu, sigma, n_eff=None, masked=True):
    """[B,P,3] -> [B,20,G] (tf_util.py:655-753 when masked else :578-652)."""
    points = np.ascontiguousarray(points, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    mu = np.ascontiguousarray(mu, np.float32)
    sigma = np.ascontiguousarray(sigma, np.float32)
    B, P, _ = points.shape
    G = mu.shape[0]
    ne = np.ascontiguousarray(n_eff if n_eff is not None else np.full(B, P), np.int32)
    out = np.empty((B, 20, G), np.float32)
    rc = lib().oracle_3dmfv(_f(points), _i(ne), _f(w), _f(mu), _f(sigma), B, P, G, 1 if masked else 0, _f(out))
    if rc:
        raise MemoryError("oracle_3dmfv")
    return out


def mups(points, n_eff, w, mu, sigma, S):
    """[B,S*P,3], [B,S] -> MuPS [B,res,res,res,20*S] (models/experts_n_est.py:59-76)."""
    points = np.ascontiguousarray(points, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    mu = np.ascontiguousarray(mu, np.float32)
    sigma = np.ascontiguousarray(sigma, np.float32)
    B = points.shape[0]
    P = points.shape[1] // S
    G = mu.shape[0]
    res = int(round(G ** (1.0 / 3.0)))
    ne = np.ascontiguousarray(n_eff, np.int32)
    out = np.empty((B, res, res, res, 20 * S) if res ** 3 == G else (B, G, 20 * S), np.float32)
    rc = lib().oracle_mups(_f(points), _i(ne), _f(w), _f(mu), _f(sigma), B, S, P, G, _f(out))
    if rc:
        raise MemoryError("oracle_mups")
    return out


def selection_keys(seed, center, scale, nbr):
    nbr = np.ascontiguousarray(nbr, np.int32)
    out = np.empty(len(nbr), np.uint32)
    lib().oracle_selection_keys(int(seed), int(center), int(scale), _i(nbr), len(nbr),
                                out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    return out


def mups_tuned(points, n_eff, w, mu, sigma, S):
    """The tuned CPU implementation (mups_oracle_tuned.c): same signature and layout as mups()."""
    points = np.ascontiguousarray(points, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    mu = np.ascontiguousarray(mu, np.float32)
    sigma = np.ascontiguousarray(sigma, np.float32)
    B = points.shape[0]
    P = points.shape[1] // S
    G = mu.shape[0]
    res = int(round(G ** (1.0 / 3.0)))
    ne = np.ascontiguousarray(n_eff, np.int32)
    out = np.empty((B, res, res, res, 20 * S) if res ** 3 == G else (B, G, 20 * S), np.float32)
    rc = tuned_lib().oracle_mups_tuned(_f(points), _i(ne), _f(w), _f(mu), _f(sigma), B, S, P, G, _f(out))
    if rc:
        raise MemoryError("oracle_mups_tuned")
    return out


def mups_f64(points, n_eff, w, mu, sigma, S, c0=BOUND_C0, c1=BOUND_C1, csum=BOUND_CSUM, masked=True):
    """float64 evaluation of the reference formula in the MuPS layout and the per-element forward-error bound of an
    fp32 evaluation of it (oracle_mups_f64): (truth, bound), both float64 [B,res,res,res,20*S]."""
    points = np.ascontiguousarray(points, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    mu = np.ascontiguousarray(mu, np.float32)
    sigma = np.ascontiguousarray(sigma, np.float32)
    B = points.shape[0]
    P = points.shape[1] // S
    G = mu.shape[0]
    res = int(round(G ** (1.0 / 3.0)))
    ne = np.ascontiguousarray(n_eff if n_eff is not None else np.full((B, S), P), np.int32).reshape(B, S)
    shape = (B, res, res, res, 20 * S) if res ** 3 == G else (B, G, 20 * S)
    out = np.empty(shape, np.float64)
    bound = np.empty(shape, np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    rc = lib().oracle_mups_f64(_f(points), _i(ne), _f(w), _f(mu), _f(sigma), B, S, P, G, 1 if masked else 0, float(c0), float(c1),
                               float(csum),
                               out.ctypes.data_as(dp), bound.ctypes.data_as(dp))
    if rc:
        raise MemoryError("oracle_mups_f64")
    return out, bound


def half1_gather(pts, query_idx, rad, lists, P, S, scale, seed, patches, n_eff):
    """Selection + gather + centre + normalise of one radius for a batch (oracle_half1_gather, OpenMP): `lists` are the
    neighbour lists cKDTree.query_ball_point returned for the batch; fills patches[:, scale*P:(scale+1)*P] and n_eff[:, scale]."""
    pts = np.ascontiguousarray(pts, np.float32)
    q = np.ascontiguousarray(query_idx, np.int64)
    lens = np.fromiter((len(l) for l in lists), np.int64, len(lists))
    off = np.zeros(len(lists) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    flat = np.fromiter((j for l in lists for j in l), np.int64, int(off[-1])) if off[-1] < 200000 else \
        np.concatenate([np.asarray(l, np.int64) for l in lists])
    i64p = ctypes.POINTER(ctypes.c_int64)
    rc = lib().oracle_half1_gather(_f(pts), q.ctypes.data_as(i64p), len(q), ctypes.c_float(np.float32(rad)),
                                   flat.ctypes.data_as(i64p), off.ctypes.data_as(i64p), int(P), int(S), int(scale),
                                   int(seed) & (2 ** 64 - 1), _f(patches), _i(n_eff))
    if rc:
        raise MemoryError("oracle_half1_gather")

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


def build(force=False):
    so = os.path.join(_HERE, "libmups_oracle.so")
    src = os.path.join(_HERE, "mups_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmups_oracle.so"], stdout=subprocess.DEVNULL)
    return so


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

"""TEST INFRASTRUCTURE ONLY -- ctypes view of oracle/_build/libtes_oracle.so (oracle/tes_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtes_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    src = os.path.join(_HERE, "tes_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.tes_oracle_cospi.restype = ctypes.c_double
        _lib.tes_oracle_cospi.argtypes = [ctypes.c_double]
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def cospi(h):
    L = lib()
    return np.array([L.tes_oracle_cospi(float(v)) for v in np.ravel(h)]).reshape(np.shape(h))


def _shapes(cand, W, b, theta):
    cand = np.ascontiguousarray(cand, dtype=np.float64)
    N, d = cand.shape
    theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(theta.shape[0], -1)
    S, F = theta.shape
    W = np.ascontiguousarray(W, dtype=np.float64).reshape(-1, F, d)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1, F)
    w_shared = int(W.shape[0] == 1 and S > 1) if W.shape[0] != S else 0
    assert W.shape[0] in (1, S) and b.shape[0] == W.shape[0]
    return cand, W, b, theta, N, d, S, F, w_shared


def rff_eval(cand, W, b, theta, sigma):
    cand, W, b, theta, N, d, S, F, ws = _shapes(cand, W, b, theta)
    out = np.empty((S, N))
    rc = lib().tes_oracle_rff_eval(
        cand.ctypes.data_as(_dp), ctypes.c_int64(N), ctypes.c_int64(d), W.ctypes.data_as(_dp),
        b.ctypes.data_as(_dp), theta.ctypes.data_as(_dp), ctypes.c_int64(S), ctypes.c_int64(F),
        ctypes.c_int(ws), ctypes.c_double(sigma), out.ctypes.data_as(_dp))
    assert rc == 0, rc
    return out


def rff_topk(cand, W, b, theta, sigma, k, idx_offset=0):
    cand, W, b, theta, N, d, S, F, ws = _shapes(cand, W, b, theta)
    idx = np.empty((S, k), dtype=np.int64)
    val = np.empty((S, k))
    rc = lib().tes_oracle_rff_topk(
        cand.ctypes.data_as(_dp), ctypes.c_int64(N), ctypes.c_int64(d), W.ctypes.data_as(_dp),
        b.ctypes.data_as(_dp), theta.ctypes.data_as(_dp), ctypes.c_int64(S), ctypes.c_int64(F),
        ctypes.c_int(ws), ctypes.c_double(sigma), ctypes.c_int(k), ctypes.c_int64(idx_offset),
        idx.ctypes.data_as(_ip), val.ctypes.data_as(_dp))
    assert rc == 0, rc
    return idx, val


def normals(seed, iteration, stream, first_elem, count):
    out = np.empty(int(count))
    rc = lib().tes_oracle_normals(ctypes.c_uint64(seed), ctypes.c_uint32(iteration),
                                  ctypes.c_uint32(stream), ctypes.c_uint64(first_elem),
                                  ctypes.c_int64(count), out.ctypes.data_as(_dp))
    assert rc == 0
    return out


def philox_raw(seed, iteration, stream, pair):
    out = (ctypes.c_uint32 * 4)()
    lib().tes_oracle_philox_raw(ctypes.c_uint64(seed), ctypes.c_uint32(iteration),
                                ctypes.c_uint32(stream), ctypes.c_uint64(pair), out)
    return [int(v) for v in out]

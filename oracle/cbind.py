"""ctypes binding of oracle/libzkoracle.so (test infrastructure only).

``zo`` = fast result-exact restatement (large-size checker), ``zr`` = faithful-
algorithm port (timed CPU baseline).  Values travel as numpy uint64 arrays of shape
(n, 2): little-endian u128 = (lo, hi), the same bytes as Rust's in-memory u128.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libzkoracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def to_arr(vals):
    a = np.empty((len(vals), 2), dtype=np.uint64)
    m = (1 << 64) - 1
    for i, v in enumerate(vals):
        a[i, 0] = v & m
        a[i, 1] = v >> 64
    return a


def from_arr(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 2)
    return [int(lo) | (int(hi) << 64) for lo, hi in a.tolist()]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _one(v):
    return to_arr([v])


def next_pow2(n):
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


def ntt(root, arr, inverse=False, faithful=False):
    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 2)
    out = np.empty((next_pow2(len(arr)), 2), dtype=np.uint64)
    name = ("zr_ntt" if faithful else ("zo_intt" if inverse else "zo_ntt"))
    rc = getattr(lib(), name)(_p(_one(root)), _p(arr), ctypes.c_size_t(len(arr)), _p(out))
    assert rc == 0, rc
    return out


def coset_lde(omega, order, offset, coeffs, faithful=False):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 2)
    out = np.empty((order, 2), dtype=np.uint64)
    fn = lib().zr_coset_lde if faithful else lib().zo_coset_lde
    rc = fn(_p(_one(omega)), ctypes.c_size_t(order), _p(_one(offset)), _p(coeffs),
            ctypes.c_size_t(len(coeffs)), _p(out))
    assert rc == 0, rc
    return out


def merkle(vals, want_nodes=False, faithful=False):
    vals = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 2)
    n = len(vals)
    root = np.empty(64, dtype=np.uint8)
    if faithful:
        rc = lib().zr_merkle_commit(_p(vals), ctypes.c_size_t(n), _p(root))
        assert rc == 0, rc
        return root.tobytes()
    nodes = np.empty((2 * n - 1, 64), dtype=np.uint8) if want_nodes else None
    rc = lib().zo_merkle(_p(vals), ctypes.c_size_t(n), _p(root),
                         _p(nodes) if want_nodes else None)
    assert rc == 0, rc
    return (root.tobytes(), nodes) if want_nodes else root.tobytes()


def fri_fold(cw, alpha, offset, omega, faithful=False):
    cw = np.ascontiguousarray(cw, dtype=np.uint64).reshape(-1, 2)
    out = np.empty((len(cw) // 2, 2), dtype=np.uint64)
    fn = lib().zr_fri_fold if faithful else lib().zo_fri_fold
    fn(_p(cw), ctypes.c_size_t(len(cw)), _p(_one(alpha)), _p(_one(offset)), _p(_one(omega)), _p(out))
    return out


def scalar(name, *vals):
    out = np.empty((1, 2), dtype=np.uint64)
    getattr(lib(), name)(*[_p(_one(v)) for v in vals], _p(out))
    return from_arr(out)[0]


def synth(seed, n, start=0):
    """Vectorised splitmix64 synthetic elements (same stream as oracle.field.synth_elements)."""
    from .field import P
    j = np.arange(start, start + n, dtype=np.uint64)

    def sm(idx):
        with np.errstate(over="ignore"):
            z = np.uint64(seed) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))
    hi = sm(j * np.uint64(2))
    lo = sm(j * np.uint64(2) + np.uint64(1))
    # reduce mod p: value < 2^128 < 2p, so one conditional subtract (p = 0xCB80..00:00..01)
    p_hi, p_lo = np.uint64(P >> 64), np.uint64(P & ((1 << 64) - 1))
    ge = (hi > p_hi) | ((hi == p_hi) & (lo >= p_lo))
    with np.errstate(over="ignore"):
        borrow = (lo < p_lo) & ge
        lo2 = np.where(ge, lo - p_lo, lo)
        hi2 = np.where(ge, hi - p_hi - borrow.astype(np.uint64), hi)
    return np.stack([lo2, hi2], axis=1)

"""Synthetic field elements for benchmarks: element j of stream `seed` is
((splitmix64(seed, 2j) << 64) | splitmix64(seed, 2j+1)) mod p, as (n, 2) uint64 (lo, hi)."""
import numpy as np

from .context import P


def _splitmix64(seed, idx):
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def elements(seed, n, start=0, step=1):
    """n elements j = start, start + step, ... of stream `seed` (a strided slice of the stream: a rank's cyclic share)"""
    j = np.uint64(start) + np.arange(n, dtype=np.uint64) * np.uint64(step)
    hi = _splitmix64(seed, j * np.uint64(2))
    lo = _splitmix64(seed, j * np.uint64(2) + np.uint64(1))
    p_hi, p_lo = np.uint64(P >> 64), np.uint64(P & ((1 << 64) - 1))
    ge = (hi > p_hi) | ((hi == p_hi) & (lo >= p_lo))        # value < 2^128 < 2p: one conditional subtract
    with np.errstate(over="ignore"):
        borrow = (lo < p_lo) & ge
        lo2 = np.where(ge, lo - p_lo, lo)
        hi2 = np.where(ge, hi - p_hi - borrow.astype(np.uint64), hi)
    return np.stack([lo2, hi2], axis=1)

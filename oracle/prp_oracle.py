"""numpy restatement of the keyed permutation behind the seeded randomization entry points.

TEST INFRASTRUCTURE ONLY (see oracle/haystack_oracle.py): only tests/, smoke() and bench.py's CPU legs
import this.  The product's definition is singlecellhaystack_b200/csrc/hs_prp.cuh; this file restates it
with numpy uint32/uint64 arithmetic so that the ids the device path draws can be fed to the fp64 oracle.

The draw stands in for `sample(x = count.cells, size = nnz)` (R/haystack_continuous.R:275): the i-th
stored value of the gene is paired with cell pi(i), pi a bijection of [0, n) keyed by (seed, stream).
"""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_C1 = np.uint32(0x9E3779B1)
_C2 = np.uint32(0x85EBCA77)


def _mix64(z: np.uint64) -> np.uint64:
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def prp_params(seed: int, stream: int, n: int):
    """(keys[4] uint32, half, mask) of the permutation of [0, n) for (seed, stream)."""
    if not 1 <= n < 2 ** 31:
        raise ValueError("n must be in [1, 2^31)")
    bits = 2
    while bits < 32 and (1 << bits) < n:
        bits += 1
    bits += bits & 1
    half = bits // 2
    with np.errstate(over="ignore"):
        z = np.uint64(seed % (1 << 64)) + (np.uint64(stream % (1 << 64)) + np.uint64(1)) * _GOLD
        k01 = _mix64(z)
        k23 = _mix64(z + _GOLD)
    keys = np.array([int(k01) & 0xFFFFFFFF, int(k01) >> 32, int(k23) & 0xFFFFFFFF, int(k23) >> 32], dtype=np.uint32)
    return keys, half, np.uint32((1 << half) - 1)


def _round(x: np.ndarray, key: np.uint32, half: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        h = x * _C1 + key
        h = h ^ (h >> np.uint32(15))
        h = h * _C2
    return h >> np.uint32(32 - half)


def _enc1(x, keys, half, mask):
    l, r = x >> np.uint32(half), x & mask
    for i in range(4):
        l, r = r, l ^ _round(r, keys[i], half)
    return (l << np.uint32(half)) | r


def _dec1(y, keys, half, mask):
    l, r = y >> np.uint32(half), y & mask
    for i in (3, 2, 1, 0):
        l, r = r ^ _round(l, keys[i], half), l
    return (l << np.uint32(half)) | r


def _walk(step, x, n, keys, half, mask):
    x = step(x, keys, half, mask)
    while True:
        bad = x >= np.uint32(n)
        if not bad.any():
            return x
        x = np.where(bad, step(x, keys, half, mask), x)


def prp_forward(seed: int, stream: int, n: int, idx) -> np.ndarray:
    """pi(idx) for idx in [0, n) (uint32 array)."""
    keys, half, mask = prp_params(seed, stream, n)
    return _walk(_enc1, np.asarray(idx, dtype=np.uint32), n, keys, half, mask)


def prp_inverse(seed: int, stream: int, n: int, cells) -> np.ndarray:
    keys, half, mask = prp_params(seed, stream, n)
    return _walk(_dec1, np.asarray(cells, dtype=np.uint32), n, keys, half, mask)


def prp_sample(seed: int, stream: int, n: int, size: int, index_base: int = 1) -> np.ndarray:
    """The `size` cells drawn for (seed, stream): pi(0..size-1) + index_base (1-based like sample())."""
    if size > n:
        raise ValueError("cannot take a sample larger than the population")
    return (prp_forward(seed, stream, n, np.arange(size, dtype=np.uint32)).astype(np.int64) + index_base).astype(np.int32)

"""TEST INFRASTRUCTURE (oracle): base-R behaviour the reference's continuous path depends on and that is not
vendored under /root/reference -- set.seed / Mersenne-Twister / sample() / sample.int() and stats::kmeans
(Hartigan-Wong, nstart restarts) -- restated in C (oracle/r_base.c, compiled here with gcc) behind ctypes, plus
the reference's own get_grid_points (R/haystack_highD.R:72-134) on top of them.

Pinned by tests/test_r_parity_cpu.py: `set.seed(1234); haystack(dat.tsne, dat.expression)` reproduces the table
printed in the reference's rendered vignette (docs/articles/a01_toy_example.html:129, :168) to the printed digits.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_SRC = HERE / "r_base.c"
_LIB = HERE / "_build" / "libr_base.so"
_lib = None


def build(force: bool = False) -> Path:
    """gcc -O2 -shared oracle/r_base.c -> oracle/_build/libr_base.so (git-ignored)."""
    if not force and _LIB.exists() and _LIB.stat().st_mtime >= _SRC.stat().st_mtime:
        return _LIB
    _LIB.parent.mkdir(exist_ok=True)
    # -ffp-contract=off: R is built without fused multiply-adds on x86-64; keep the k-means arithmetic identical
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(_LIB), str(_SRC), "-lm"], check=True)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(build()))
        lib.r_rng_sizeof.restype = C.c_size_t
        lib.r_set_seed.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        lib.r_unif_rand.argtypes = [C.c_void_p]
        lib.r_unif_rand.restype = C.c_double
        lib.r_unif_index.argtypes = [C.c_void_p, C.c_double]
        lib.r_unif_index.restype = C.c_double
        lib.r_sample_int_many.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.r_sample_prob_one.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.r_sample_prob_one.restype = C.c_int
        lib.r_kmeans_hartigan_wong.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


class RRng:
    """R's default generator: RNGkind("Mersenne-Twister", "Inversion", sample.kind) after set.seed(seed)."""

    def __init__(self, seed: int, sample_kind: str = "Rejection"):
        if sample_kind not in ("Rejection", "Rounding"):
            raise ValueError("sample.kind must be 'Rejection' (R >= 3.6.0) or 'Rounding'")
        self._l = _load()
        self._state = C.create_string_buffer(self._l.r_rng_sizeof())
        self._l.r_set_seed(self._state, C.c_uint32(int(seed) & 0xFFFFFFFF), int(sample_kind == "Rounding"))

    def unif_rand(self) -> float:
        return float(self._l.r_unif_rand(self._state))

    def sample_int(self, n: int, size: int | None = None, count: int | None = None) -> np.ndarray:
        """sample.int(n, size): `size` distinct 1-based ids.  With `count`, that many consecutive draws as rows."""
        size = int(n) if size is None else int(size)
        if size > n:
            raise ValueError("cannot take a sample larger than the population when 'replace = FALSE'")
        if n > 1e7 and size <= n / 2:
            raise NotImplementedError("sample.int switches to its hashing algorithm for n > 1e7; not restated")
        k = 1 if count is None else int(count)
        out = np.empty((k, size), dtype=np.int32)
        self._l.r_sample_int_many(self._state, int(n), size, k, out.ctypes.data_as(C.c_void_p))
        return out[0] if count is None else out

    def sample(self, x) -> np.ndarray:
        """sample(x) for a vector x of length > 1: x[sample.int(length(x))]."""
        x = np.asarray(x)
        if x.ndim != 1 or x.size < 2:
            raise ValueError("sample(x) restated for vectors of length >= 2 only")
        return x[self.sample_int(x.size) - 1]

    def sample_prob_one(self, prob) -> int:
        """sample(length(prob), 1, prob = prob): one 1-based id."""
        p = np.ascontiguousarray(prob, dtype=np.float64)
        return int(self._l.r_sample_prob_one(self._state, p.size, p.ctypes.data_as(C.c_void_p)))


def _unique_rows(x: np.ndarray) -> np.ndarray:
    """unique(x) for a matrix: first occurrences, original order."""
    _, first = np.unique(x, axis=0, return_index=True)
    return x[np.sort(first)]


def kmeans_hartigan_wong(x: np.ndarray, centers: np.ndarray, iter_max: int = 10):
    """One run of kmeans(x, centers, algorithm = "Hartigan-Wong") from given initial centres."""
    lib = _load()
    a = np.asfortranarray(x, dtype=np.float64)
    c = np.asfortranarray(centers, dtype=np.float64).copy(order="F")
    m, n = a.shape
    k = c.shape[0]
    ic1 = np.zeros(m, dtype=np.int32)
    nc = np.zeros(k, dtype=np.int32)
    wss = np.zeros(k, dtype=np.float64)
    it = C.c_int(int(iter_max))
    ifault = C.c_int(0)
    imaxqtr = int(min(2 ** 31 - 1, 50 * m))
    lib.r_kmeans_hartigan_wong(a.ctypes.data_as(C.c_void_p), m, n, c.ctypes.data_as(C.c_void_p), k,
                               ic1.ctypes.data_as(C.c_void_p), nc.ctypes.data_as(C.c_void_p), C.byref(it),
                               wss.ctypes.data_as(C.c_void_p), C.byref(ifault), imaxqtr)
    if ifault.value == 1:
        raise ValueError("empty cluster: try a better set of initial centers")
    if ifault.value == 3:
        raise ValueError("number of cluster centres must lie between 1 and nrow(x)")
    return dict(centers=np.array(c), cluster=ic1, size=nc, withinss=wss, iter=it.value, ifault=ifault.value)


def kmeans(x: np.ndarray, k: int, rng: RRng, iter_max: int = 10, nstart: int = 1):
    """stats::kmeans(x, centers = k, iter.max, nstart): initial centres are rows sample.int() picks from the
    distinct rows; the run with the smallest total within-cluster sum of squares wins."""
    x = np.asarray(x, dtype=np.float64)
    m = x.shape[0]
    cn = None
    if nstart == 1:
        centers = x[rng.sample_int(m, k) - 1]
        dup = np.unique(centers, axis=0).shape[0] != k
    else:
        dup = False
    if nstart >= 2 or dup:
        cn = _unique_rows(x)
        mm = cn.shape[0]
        if mm < k:
            raise ValueError("more cluster centers than distinct data points.")
        centers = cn[rng.sample_int(mm, k) - 1]
    best = kmeans_hartigan_wong(x, centers, iter_max)
    best_ss = best["withinss"].sum()
    if nstart >= 2 and cn is not None:
        for _ in range(2, nstart + 1):
            centers = cn[rng.sample_int(cn.shape[0], k) - 1]
            z = kmeans_hartigan_wong(x, centers, iter_max)
            if z["withinss"].sum() < best_ss:
                best, best_ss = z, z["withinss"].sum()
    return best


def get_grid_points(x: np.ndarray, rng: RRng, method: str = "centroid", grid_points: int = 100) -> np.ndarray:
    """R/haystack_highD.R:72-134 on R's own generator."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    g = min(int(grid_points), n)                                                    # :74-77
    if method == "centroid":
        return kmeans(x, g, rng, iter_max=10, nstart=10)["centers"]                 # :82-85
    if method != "seeding":
        raise ValueError("method should be 'centroid' or 'seeding'")
    grid = np.empty((g, x.shape[1]))
    index = int(rng.sample_int(n, 1)[0])                                            # :95
    grid[0] = x[index - 1]
    dist = np.empty((n, g))
    dist[:, 0] = np.sqrt(((x - grid[0]) ** 2).sum(axis=1))
    min_dist = dist[:, 0].copy()
    for i in range(1, g):
        index = rng.sample_prob_one(min_dist ** 2)                                  # :107-110
        grid[i] = x[index - 1]
        dist[:, i] = np.sqrt(((x - grid[i]) ** 2).sum(axis=1))
        min_dist = np.minimum(min_dist, dist[:, i])                                 # :114-116
    w = np.array([2.0, 1.0, 1.0])
    for i in range(g):                                                              # :124-128
        o = np.argsort(dist[:, i], kind="stable")[:3]
        grid[i] = (x[o] * w[:, None]).sum(axis=0) / w.sum()
    return grid

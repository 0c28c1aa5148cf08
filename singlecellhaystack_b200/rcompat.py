"""R's random stream for hosts without R: `RRandom(seed)` is `set.seed(seed)` on R's default generator
(Mersenne-Twister, sample.kind "Rejection").  Pass it as `rng=` to haystack() / haystack_continuous_highD() and the
grid (kmeans, Hartigan-Wong, nstart = 10), the randomizations (sample() in the reference's gene-major order) and the
cross-validation folds are the ones R would draw, so the B200 path reproduces an R session's numbers -- e.g. the table
of the reference's vignette (tests/test_gpu_r_parity.py).  Host-only helpers of libhaystack_b200.so (hs_r_*)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

__all__ = ["RRandom"]


class RRandom:
    def __init__(self, seed: int, sample_kind: str = "Rejection"):
        if sample_kind not in ("Rejection", "Rounding"):
            raise ValueError("sample.kind must be 'Rejection' or 'Rounding'")
        self._lib = _lib.load()
        h = C.c_void_p(None)
        rc = self._lib.hs_r_rng_create(C.c_uint32(int(seed) & 0xFFFFFFFF), int(sample_kind == "Rounding"), C.byref(h))
        if rc != _lib.HS_OK:
            raise MemoryError("hs_r_rng_create failed")
        self._h = h

    def __del__(self):  # pragma: no cover
        try:
            if getattr(self, "_h", None):
                self._lib.hs_r_rng_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def unif_rand(self) -> float:
        return float(self._lib.hs_r_unif_rand(self._h))

    def sample_int(self, n: int, size: int | None = None, count: int | None = None) -> np.ndarray:
        """sample.int(n, size): distinct 1-based ids; `count` consecutive draws come back as rows."""
        size = int(n) if size is None else int(size)
        k = 1 if count is None else int(count)
        out = np.empty((k, size), dtype=np.int32)
        rc = self._lib.hs_r_sample_int(self._h, int(n), size, k, C.c_void_p(out.ctypes.data))
        if rc != _lib.HS_OK:
            raise ValueError(_lib.last_error(None))
        return out[0] if count is None else out

    def sample(self, x) -> np.ndarray:
        """sample(x) for a vector of length >= 2."""
        x = np.asarray(x)
        if x.ndim != 1 or x.size < 2:
            raise ValueError("sample(x) is provided for vectors of length >= 2")
        return x[self.sample_int(x.size) - 1]

    def sample_prob_one(self, prob) -> int:
        p = np.ascontiguousarray(prob, dtype=np.float64)
        out = C.c_int32(0)
        rc = self._lib.hs_r_sample_prob_one(self._h, p.size, C.c_void_p(p.ctypes.data), C.byref(out))
        if rc != _lib.HS_OK:
            raise ValueError(_lib.last_error(None))
        return int(out.value)

    def kmeans(self, x, k: int, iter_max: int = 10, nstart: int = 1) -> np.ndarray:
        """kmeans(x, centers = k, iter.max, nstart)$centers (Hartigan-Wong)."""
        x = np.asfortranarray(x, dtype=np.float64)
        m, d = x.shape
        centers = np.empty((int(k), d), dtype=np.float64, order="F")
        rc = self._lib.hs_r_kmeans(self._h, C.c_void_p(x.ctypes.data), m, d, int(k), int(iter_max), int(nstart),
                                   C.c_void_p(centers.ctypes.data), None, None)
        if rc != _lib.HS_OK:
            raise ValueError(_lib.last_error(None))
        return centers

    # ---- the two hooks haystack_continuous_highD() takes ----
    def perm_source(self, kind, gene, n, size, count):
        """sample(x = row) (dense, R/haystack_continuous.R:259) / sample(x = n, size = nnz) (sparse, :275)."""
        return self.sample_int(n, n if kind == "dense" else size, count=count)

    def folds(self, n: int, n_cv: int = 10) -> np.ndarray:
        """sample(rep(1:n_cv, length.out = n))  (R/haystack_continuous.R:437, :505)."""
        return self.sample(np.resize(np.arange(1, n_cv + 1), n))

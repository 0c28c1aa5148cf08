"""Sparse containers mirroring the Matrix classes the reference accepts.

`dgRMatrix` (row-compressed: slots p, j, x, Dim) is what the scored path consumes
(R/haystack_continuous.R:207-212, R/sparse.R:39-49); `dgCMatrix` (column-compressed: p, i, x) is
what users usually hold and is converted with `as_RsparseMatrix`, the counterpart of
`as(expression, "RsparseMatrix")` at R/haystack_continuous.R:113-116.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

__all__ = ["dgRMatrix", "dgCMatrix", "as_RsparseMatrix", "as_CsparseMatrix", "is_sparse",
           "extract_row_dgRMatrix", "extract_row_dgRMatrix_as_sparse"]


@dataclass
class dgRMatrix:
    """Row-compressed numeric matrix: p (M+1), j (0-based column ids), x (values), Dim (M, N)."""
    p: np.ndarray
    j: np.ndarray
    x: np.ndarray
    Dim: tuple
    rownames: list | None = None
    colnames: list | None = None

    def __post_init__(self):
        self.p = np.ascontiguousarray(self.p, dtype=np.int64 if np.asarray(self.p).dtype == np.int64 else np.int32)
        self.j = np.ascontiguousarray(self.j, dtype=np.int32)
        # R's @x is double; float32 hosts (AnnData) keep their width -- the device multiplies in fp32 anyway
        xa = np.asarray(self.x)
        self.x = np.ascontiguousarray(xa, dtype=np.float32 if xa.dtype == np.float32 else np.float64)
        self.Dim = (int(self.Dim[0]), int(self.Dim[1]))
        if self.p.size != self.Dim[0] + 1:
            raise ValueError("slot p must have nrow+1 entries")

    @property
    def shape(self):
        return self.Dim

    def row_subset(self, keep: np.ndarray) -> "dgRMatrix":
        """expression[!sel.bad, ] (R/haystack_continuous.R:77)."""
        keep = np.asarray(keep, dtype=bool)
        rows = np.nonzero(keep)[0]
        counts = (self.p[1:] - self.p[:-1])[rows]
        p = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        if rows.size == self.Dim[0]:
            return self
        idx = np.concatenate([np.arange(self.p[r], self.p[r + 1]) for r in rows]) if rows.size else np.zeros(0, np.int64)
        names = [self.rownames[r] for r in rows] if self.rownames is not None else None
        return dgRMatrix(p=p, j=self.j[idx], x=self.x[idx], Dim=(rows.size, self.Dim[1]), rownames=names,
                         colnames=self.colnames)

    def toarray(self) -> np.ndarray:
        out = np.zeros(self.Dim, dtype=np.float64)
        rows = np.repeat(np.arange(self.Dim[0]), np.diff(self.p))
        out[rows, self.j] = self.x
        return out


@dataclass
class dgCMatrix:
    """Column-compressed numeric matrix: p (N+1), i (0-based row ids), x, Dim (M, N)."""
    p: np.ndarray
    i: np.ndarray
    x: np.ndarray
    Dim: tuple
    rownames: list | None = None
    colnames: list | None = None

    def __post_init__(self):
        self.p = np.ascontiguousarray(self.p)
        self.i = np.ascontiguousarray(self.i, dtype=np.int32)
        self.x = np.ascontiguousarray(self.x, dtype=np.float64)
        self.Dim = (int(self.Dim[0]), int(self.Dim[1]))
        if self.p.size != self.Dim[1] + 1:
            raise ValueError("slot p must have ncol+1 entries")

    @property
    def shape(self):
        return self.Dim


def is_sparse(m) -> bool:
    if isinstance(m, (dgRMatrix, dgCMatrix)):
        return True
    try:
        import scipy.sparse as sp
        return sp.issparse(m)
    except Exception:  # pragma: no cover
        return False


def as_CsparseMatrix(m, rownames=None, colnames=None) -> dgCMatrix:
    """as(m, "CsparseMatrix") for a dense array or scipy matrix (explicit zeros dropped for dense)."""
    import scipy.sparse as sp
    if isinstance(m, dgCMatrix):
        return m
    c = sp.csc_matrix(m.toarray() if isinstance(m, dgRMatrix) else m)
    c.sort_indices()
    return dgCMatrix(p=c.indptr, i=c.indices, x=c.data, Dim=c.shape, rownames=rownames, colnames=colnames)


def as_RsparseMatrix(m, rownames=None, colnames=None, ctx=None) -> dgRMatrix:
    """as(m, "RsparseMatrix"): dgCMatrix / scipy sparse / dense -> dgRMatrix with ascending j per row
    (counting sort over the stored entries; stored zeros of a sparse input are kept, as Matrix does).
    With a DeviceContext `ctx`, a dgCMatrix is converted on the GPU (hs_csc_to_csr): same slots, bit for bit."""
    if isinstance(m, dgRMatrix):
        return m
    if isinstance(m, dgCMatrix) and ctx is not None:
        p, j, x = ctx.csc_to_csr(m.p, m.i, m.x, m.Dim[0], m.Dim[1])
        return dgRMatrix(p=p, j=j, x=x, Dim=m.Dim, rownames=rownames or m.rownames, colnames=colnames or m.colnames)
    if isinstance(m, dgCMatrix):
        nrow, ncol = m.Dim
        counts = np.bincount(m.i, minlength=nrow)
        p = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        order = np.argsort(m.i, kind="stable")          # stable: column ids stay ascending within a row
        cols = np.repeat(np.arange(ncol, dtype=np.int32), np.diff(m.p))
        return dgRMatrix(p=p, j=cols[order], x=m.x[order], Dim=m.Dim,
                         rownames=rownames or m.rownames, colnames=colnames or m.colnames)
    import scipy.sparse as sp
    if sp.issparse(m):
        r = sp.csr_matrix(m)
        r.sort_indices()
        return dgRMatrix(p=r.indptr.astype(np.int64), j=r.indices, x=r.data, Dim=r.shape, rownames=rownames,
                         colnames=colnames)
    a = np.asarray(m, dtype=np.float64)
    nz = a != 0
    p = np.concatenate([[0], np.cumsum(nz.sum(axis=1))]).astype(np.int64)
    rows, cols = np.nonzero(nz)
    return dgRMatrix(p=p, j=cols.astype(np.int32), x=a[rows, cols], Dim=a.shape, rownames=rownames, colnames=colnames)


def extract_row_dgRMatrix(m: dgRMatrix, i: int = 1) -> np.ndarray:
    """Row i (1-based, as in R) as a dense numeric vector -- R/sparse.R:28-37."""
    r = np.zeros(m.Dim[1], dtype=np.float64)
    lo, hi = int(m.p[i - 1]), int(m.p[i])
    r[m.j[lo:hi]] = m.x[lo:hi]
    return r


def extract_row_dgRMatrix_as_sparse(m: dgRMatrix, i: int):
    """Row i (1-based) as dict(ind 1-based, val) -- R/sparse.R:39-49."""
    lo, hi = int(m.p[i - 1]), int(m.p[i])
    if lo < hi:
        return dict(ind=m.j[lo:hi].astype(np.int64) + 1, val=m.x[lo:hi].copy())
    return dict(ind=np.zeros(0, dtype=np.int64), val=np.zeros(0, dtype=np.float64))

"""Minimal reader for R's `save()` format (gzip'd "RDX2" + XDR serialization v2).

TEST INFRASTRUCTURE ONLY (see oracle/README.md): used by `oracle/make_golden.py` to turn the
reference's one in-tree fixture, `/root/reference/data/toy.rda`, into `tests/golden/toy.npz`.
Nothing in the product package imports this.

Only the SEXP types that occur in `toy.rda` are handled (pairlist, symbol, character, integer,
real, string vector, generic vector, attributes).  The format is R's own (`R_Serialize` in base R,
not vendored under /root/reference); the layout facts used here are the documented ones:
big-endian XDR, a 32-bit flags word per item with the SEXPTYPE in the low byte, bit 8 = is-object,
bit 9 = has-attributes, bit 10 = has-tag.
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

NILVALUE_SXP = 254
REFSXP = 255
SYMSXP, LISTSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP = 1, 2, 9, 10, 13, 14, 16, 19


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.o = 0
        self.refs: list = []

    def i32(self) -> int:
        (v,) = struct.unpack_from(">i", self.b, self.o)
        self.o += 4
        return v

    def take(self, n: int) -> bytes:
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def length(self) -> int:
        n = self.i32()
        if n == -1:  # long vector: two more words
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def item(self):
        flags = self.i32()
        ty = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if ty == NILVALUE_SXP:
            return None
        if ty == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if ty == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if ty == LISTSXP:
            out = []
            # iterative walk down the CDR chain
            while True:
                attr = self.item() if has_attr else None  # noqa: F841 (pairlist attrs unused)
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                ty = flags & 0xFF
                if ty == NILVALUE_SXP:
                    break
                if ty != LISTSXP:
                    raise ValueError(f"unexpected CDR type {ty}")
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if ty == CHARSXP:
            n = self.i32()
            return None if n == -1 else self.take(n).decode("utf-8", "replace")
        if ty in (INTSXP, LGLSXP):
            n = self.length()
            val = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
        elif ty == REALSXP:
            n = self.length()
            val = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif ty == STRSXP:
            n = self.length()
            val = [self.item() for _ in range(n)]
        elif ty == VECSXP:
            n = self.length()
            val = [self.item() for _ in range(n)]
        else:
            raise ValueError(f"unsupported SEXPTYPE {ty} at offset {self.o}")
        attrs = dict(self.item()) if has_attr else {}
        return {"value": val, "attr": attrs}


def read_rda(path: str) -> dict:
    """Return {object name: decoded object} for an RDX2 `.rda` file."""
    raw = gzip.open(path, "rb").read()
    if raw[:5] != b"RDX2\n":
        raise ValueError("not an RDX2 file")
    if raw[5:7] != b"X\n":
        raise ValueError("only XDR serialization is handled")
    r = _Reader(raw[7:])
    version, _writer, _minver = r.i32(), r.i32(), r.i32()
    if version != 2:
        raise ValueError(f"serialization version {version} not handled")
    top = r.item()
    return {tag: val for tag, val in top}


def _rownames(attr) -> list | None:
    rn = attr.get("row.names")
    if rn is None:
        return None
    v = rn["value"]
    if isinstance(v, list):
        return v
    # compact form c(NA_integer_, -n)
    n = abs(int(v[1]))
    return [str(i + 1) for i in range(n)]


def load_toy(path: str) -> dict:
    """Decode `toy.rda` into plain arrays.

    Returns dict with
      coord       (601, 2) float64   -- dat.tsne (data.frame tSNE1,tSNE2)
      coord_names list[str]
      cell_names  list[str]
      expression  (500, 601) float64 -- dat.expression (genes x cells), R column-major respected
      gene_names  list[str]
    """
    objs = read_rda(path)
    tsne = objs["dat.tsne"]
    cols = [np.asarray(c["value"], dtype=np.float64) for c in tsne["value"]]
    coord = np.stack(cols, axis=1)
    coord_names = tsne["attr"]["names"]["value"]
    cell_names = _rownames(tsne["attr"])
    ex = objs["dat.expression"]
    dim = ex["attr"]["dim"]["value"]
    expr = np.asarray(ex["value"], dtype=np.float64).reshape(int(dim[0]), int(dim[1]), order="F")
    dn = ex["attr"].get("dimnames")
    gene_names = dn["value"][0]["value"] if dn and dn["value"][0] is not None else None
    return dict(coord=coord, coord_names=coord_names, cell_names=cell_names,
                expression=expr, gene_names=gene_names)


if __name__ == "__main__":
    import sys
    t = load_toy(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/toy.rda")
    e = t["expression"]
    print("coord", t["coord"].shape, t["coord_names"], t["coord"].mean(0), t["coord"].std(0, ddof=1))
    print("expr", e.shape, e.max(), (e != 0).mean(), t["gene_names"][:3], t["cell_names"][:3])

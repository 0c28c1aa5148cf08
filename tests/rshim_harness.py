"""TEST INFRASTRUCTURE: build rshim/src/haystack_b200_shim.c against the mock of R's C API (tests/rmock) and drive
its registered .Call routines from Python, with R's conventions (column-major matrices, 1-based ids, R errors)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
MOCK = ROOT / "tests" / "rmock"
OUT = ROOT / "tests" / "rmock" / "_build" / "haystackB200.so"


def build() -> Path:
    OUT.parent.mkdir(exist_ok=True)
    libdir = ROOT / "singlecellhaystack_b200" / "lib"
    cmd = ["gcc", "-std=c11", "-O1", "-fPIC", "-shared", f"-I{MOCK}", f"-I{ROOT / 'include'}", "-o", str(OUT),
           str(ROOT / "rshim" / "src" / "haystack_b200_shim.c"), str(MOCK / "rmock.c"),
           f"-L{libdir}", "-lhaystack_b200", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("the .Call shim does not compile against the R API mock:\n" + res.stderr)
    return OUT


class CallMethod(C.Structure):
    _fields_ = [("name", C.c_char_p), ("fun", C.c_void_p), ("numArgs", C.c_int)]


class Sexp(C.Structure):
    _fields_ = [("type", C.c_int), ("length", C.c_ssize_t), ("nrow", C.c_int), ("ncol", C.c_int), ("data", C.c_void_p),
                ("finalizer", C.c_void_p)]


class RError(RuntimeError):
    pass


class Shim:
    def __init__(self):
        self.lib = C.CDLL(str(build()))
        L = self.lib
        L.rmock_load.restype = C.POINTER(CallMethod)
        L.rmock_last_error.restype = C.c_char_p
        for f in ("rmock_real", "rmock_int"):
            getattr(L, f).restype = C.POINTER(Sexp)
            getattr(L, f).argtypes = [C.c_void_p, C.c_ssize_t, C.c_int, C.c_int]
        L.rmock_lgl.restype = C.POINTER(Sexp)
        L.rmock_call.restype = C.POINTER(Sexp)
        L.rmock_call.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.POINTER(Sexp))]
        L.rmock_finalize.argtypes = [C.POINTER(Sexp)]
        self.nil = C.cast(C.addressof(C.c_void_p.in_dll(L, "R_NilValue")), C.POINTER(C.POINTER(Sexp)))[0]
        tab = L.rmock_load()
        self.registered = {}
        i = 0
        while tab[i].name:
            self.registered[tab[i].name.decode()] = tab[i].numArgs
            i += 1

    # ---- R values ----
    def real(self, a):
        a = np.asfortranarray(a, dtype=np.float64)
        nr, nc = (a.shape if a.ndim == 2 else (0, 0))
        flat = np.ascontiguousarray(a.ravel(order="F"))
        return self.lib.rmock_real(flat.ctypes.data, flat.size, nr, nc)

    def integer(self, a):
        a = np.asfortranarray(a, dtype=np.int32)
        nr, nc = (a.shape if a.ndim == 2 else (0, 0))
        flat = np.ascontiguousarray(a.ravel(order="F"))
        return self.lib.rmock_int(flat.ctypes.data, flat.size, nr, nc)

    def logical(self, v):
        return self.lib.rmock_lgl(int(bool(v)))

    def call(self, name, *args):
        arr = (C.POINTER(Sexp) * len(args))(*args)
        out = self.lib.rmock_call(name.encode(), len(args), arr)
        if not out:
            raise RError(self.lib.rmock_last_error().decode())
        return out

    def to_numpy(self, s):
        s = s.contents
        if s.type == 0:
            return None
        if s.type == 19:      # VECSXP
            ptrs = C.cast(s.data, C.POINTER(C.POINTER(Sexp)))
            return [self.to_numpy(ptrs[i]) for i in range(s.length)]
        dt = np.float64 if s.type == 14 else np.int32
        a = np.ctypeslib.as_array(C.cast(s.data, C.POINTER(C.c_double if s.type == 14 else C.c_int)), shape=(s.length,)).astype(dt)
        return a.reshape((s.nrow, s.ncol), order="F") if s.nrow else a


def dot_calls(r_source: str):
    """[(routine, n_args)] of every .Call("routine", ..., PACKAGE = ...) in an R source text."""
    out = []
    for m in re.finditer(r'\.Call\(\s*"(\w+)"', r_source):
        i = m.end()
        depth, n_args, j = 1, 0, i
        while depth:
            ch = r_source[j]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                n_args += 1
            j += 1
        body = r_source[i:j]
        if "PACKAGE" in body:
            n_args -= 1
        out.append((m.group(1), n_args))
    return out

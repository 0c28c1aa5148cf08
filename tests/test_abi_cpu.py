"""The C-ABI library loads without a GPU and exports every symbol include/haystack_b200.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "haystack_b200.h"


def _declared():
    txt = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(hs_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound():
    from singlecellhaystack_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.hs_abi_version() == _lib.ABI_VERSION


def test_only_abi_symbols_are_visible():
    import subprocess
    from singlecellhaystack_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    ours = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert sorted(ours) == _declared()


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from singlecellhaystack_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p(None)
    rc = lib.hs_create(0, C.byref(h))
    assert rc == _lib.HS_ERR_CUDA and not h.value
    assert "no CPU fallback" in _lib.last_error(None)
    assert lib.hs_launch_count(None) == 0
    lib.hs_destroy(None)      # must be a no-op


def test_built_for_sm_100a_with_lineinfo():
    import shutil
    import subprocess
    from singlecellhaystack_b200 import _lib
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert all("sm_100a" in l for l in out.splitlines() if "ELF file" in l)

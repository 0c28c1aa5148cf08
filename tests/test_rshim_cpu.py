"""The R drop-in (rshim/) checked without R: the .Call shim compiles, links against libhaystack_b200.so and registers
its routines through a mock of R's C API; every .Call in the R sources names a registered routine with the right
number of arguments; the package skeleton declares what R needs; the replacement function keeps the reference's
signature (R/haystack_continuous.R:27-31) and fills every field of the result object (:318-338)."""
import re
from pathlib import Path

import pytest

from rshim_harness import ROOT, Shim, dot_calls

RSHIM = ROOT / "rshim"
REF_ARGS = ["x", "expression", "grid.points", "weights.advanced.Q", "dir.randomization", "scale", "grid.method",
            "randomization.count", "n.genes.to.randomize", "selection.method.genes.to.randomize", "grid.coord",
            "spline.method"]          # R/haystack_continuous.R:27-31, man/haystack_continuous_highD.Rd


@pytest.fixture(scope="module")
def shim():
    return Shim()


def test_shim_compiles_links_and_registers(shim):
    assert set(shim.registered) >= {"hsb_create", "hsb_density", "hsb_dkl_dense", "hsb_dkl_csr", "hsb_dkl_perm_dense",
                                    "hsb_dkl_perm_csr", "hsb_dkl_csr_randomized", "hsb_perm_sample"}
    with pytest.raises(Exception):            # arity is enforced like .Call with registration does
        shim.call("hsb_create")
    # no GPU here: the library refuses loudly and the shim turns that into an R error, not a crash
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(Exception, match="no CUDA device|hs_create failed"):
            shim.call("hsb_create", shim.integer([0]))


def test_every_dot_call_is_registered_with_matching_arity(shim):
    seen = set()
    for f in sorted((RSHIM / "R").glob("*.R")):
        for name, n_args in dot_calls(f.read_text()):
            assert name in shim.registered, f"{f.name}: .Call target {name} is not registered in the shim"
            assert shim.registered[name] == n_args, f"{f.name}: {name} called with {n_args} args, registered with {shim.registered[name]}"
            seen.add(name)
    assert seen == set(shim.registered), f"registered but never called from R: {set(shim.registered) - seen}"


def test_package_skeleton():
    ns = (RSHIM / "NAMESPACE").read_text()
    assert re.search(r"useDynLib\(haystackB200,\s*\.registration\s*=\s*TRUE\)", ns)
    desc = (RSHIM / "DESCRIPTION").read_text()
    for field in ("Package: haystackB200", "Imports:", "singleCellHaystack", "NeedsCompilation: yes"):
        assert field in desc
    exported = set(re.findall(r"[\w.]+", " ".join(re.findall(r"export\(([^)]*)\)", ns))))
    defined = set()
    for f in (RSHIM / "R").glob("*.R"):
        defined |= set(re.findall(r"^([\w.]+)\s*<-\s*function", f.read_text(), flags=re.M))
    assert exported <= defined, exported - defined
    assert "R_init_haystackB200" in (RSHIM / "src" / "haystack_b200_shim.c").read_text()


def test_drop_in_function_keeps_signature_and_result_object():
    src = (RSHIM / "R" / "haystack_continuous_highD.R").read_text()
    m = re.search(r"haystack_continuous_highD <- function\((.*?)\)\s*\{", src, flags=re.S)
    args = [a.split("=")[0].strip() for a in re.split(r",(?![^()]*\))", m.group(1))]
    assert args == REF_ARGS
    ref = Path("/root/reference/R/haystack_continuous.R")
    if ref.exists():          # the reference's own text, where it is mounted
        rm = re.search(r"haystack_continuous_highD = function\((.*?)\)\{", ref.read_text(), flags=re.S)
        ref_args = [a.split("=")[0].strip() for a in re.split(r",(?![^()]*\))", rm.group(1))]
        assert args == ref_args
        defaults = lambda text: {a.split("=")[0].strip(): a.split("=", 1)[1].strip().replace(" ", "") for a in re.split(r",(?![^()]*\))", text) if "=" in a}
        assert defaults(m.group(1)) == defaults(rm.group(1))
    for field in ("results", "D_KL", "log.p.vals", "log.p.adj", "method", "randomization", "grid.coordinates", "coord_mean",
                  "coord_std", "densities", "cv", "genes_to_randomize", "KLD_rand", 'class(res) <- "haystack"'):
        assert field in src, field
    # the random draws stay R's: grid through the reference's get_grid_points, fit through its get_log_p_D_KL_continuous
    assert "sch$get_grid_points" in src and "sch$get_log_p_D_KL_continuous" in src


def _strip_r(src: str) -> str:
    """R source without string literals and comments (enough of a lexer for bracket matching)."""
    out, i, n = [], 0, len(src)
    while i < n:
        ch = src[i]
        if ch == "#":
            while i < n and src[i] != "\n":
                i += 1
        elif ch in "\"'":
            q = ch
            i += 1
            while i < n and src[i] != q:
                i += 2 if src[i] == "\\" else 1
            i += 1
        elif ch == "`":
            i = src.index("`", i + 1) + 1
        else:
            out.append(ch)
            i += 1
    return "".join(out)


def test_r_sources_have_balanced_brackets_and_known_calls():
    """No R interpreter here: at least every bracket of the R sources closes in order, and every b200_* function the
    drop-in calls is defined in the package."""
    defined, called = set(), set()
    for f in sorted((RSHIM / "R").glob("*.R")):
        code = _strip_r(f.read_text())
        stack = []
        pairs = {")": "(", "]": "[", "}": "{"}
        for pos, ch in enumerate(code):
            if ch in "([{":
                stack.append(ch)
            elif ch in ")]}":
                assert stack and stack[-1] == pairs[ch], f"{f.name}: unbalanced '{ch}' near ...{code[max(0, pos - 60):pos + 1]!r}"
                stack.pop()
        assert not stack, f"{f.name}: {len(stack)} unclosed bracket(s)"
        defined |= set(re.findall(r"^\s*([.\w]+)\s*<-\s*function\b", code, flags=re.M))
        called |= set(re.findall(r"(?<![\w.])(\.?b200_\w+)\s*\(", code))
    assert called <= defined, f"called but not defined: {sorted(called - defined)}"

"""GPU parity on the reference's toy fixture: CUDA path (through the C ABI) vs the fp64 oracle.

Tolerances (north_star): grid / permutation indices bit-exact; D_KL 1e-5 relative; log10 p 1e-3 absolute.
"""
import numpy as np
import pytest

from oracle import haystack_oracle as ho
from singlecellhaystack_b200 import synth

pytestmark = pytest.mark.gpu

RTOL_DKL = 1e-5


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


@pytest.mark.parametrize("g", [60, 100])
def test_density_stage(ctx, toy, g):
    xs, _, _ = ho.scale_coords(toy["coord"])
    grid = toy[f"grid{g}"]
    ref = ho.density_stage(xs, grid)
    out = ctx.density(xs, grid, want_densities=True, want_nearest=True)
    assert abs(out["bandwidth"] - ref["bandwidth"]) <= 1e-12 * ref["bandwidth"]
    assert abs(out["bandwidth"] - float(toy[f"bandwidth{g}"])) <= 1e-12 * ref["bandwidth"]
    np.testing.assert_array_equal(out["nearest"], ref["nearest"])          # grid indices: bit-exact
    np.testing.assert_array_equal(out["nearest"], toy[f"nearest{g}"])
    np.testing.assert_allclose(out["Q"], ref["Q"], rtol=1e-10, atol=0)
    np.testing.assert_allclose(out["densities"], ref["density"], rtol=1e-9, atol=1e-300)
    assert ctx.shape == (601, g)


def test_density_weights_advanced_q(ctx, toy):
    xs, _, _ = ho.scale_coords(toy["coord"])
    grid = toy["grid60"]
    w = 0.5 + (np.arange(601) % 7) / 3.0
    ref = ho.density_stage(xs, grid, w)
    out = ctx.density(xs, grid, w_q=w)
    np.testing.assert_allclose(out["Q"], ref["Q"], rtol=1e-10, atol=0)


@pytest.mark.parametrize("g", [60, 100])
def test_dkl_observed_dense_and_csr(ctx, toy, g):
    xs, _, _ = ho.scale_coords(toy["coord"])
    ctx.density(xs, toy[f"grid{g}"])
    expr = toy["expression"]
    gold = toy[f"dkl{g}"]
    d_dense = ctx.dkl_dense(expr)
    assert _rel(d_dense, gold) < RTOL_DKL
    sp = ho.RowSparse.from_dense(expr)
    d_csr = ctx.dkl_csr(sp.p, sp.j, sp.x)
    assert _rel(d_csr, gold) < RTOL_DKL
    d_csr32 = ctx.dkl_csr(sp.p.astype(np.int32), sp.j, sp.x)
    np.testing.assert_array_equal(d_csr, d_csr32)


def test_randomized_sparse_and_dense(ctx, toy):
    xs, _, _ = ho.scale_coords(toy["coord"])
    grid = toy["grid100"]
    ref = ho.density_stage(xs, grid)
    ctx.density(xs, grid)
    expr = toy["expression"]
    sp = ho.RowSparse.from_dense(expr)
    genes = toy["full_genes_to_randomize"][:12]
    src = synth.perm_source(int(toy["seed"]))
    n = expr.shape[1]
    vals, inds_o, inds_d = [], [], []
    for gidx in genes:
        _, val = ho.extract_row_dgRMatrix_as_sparse(sp, int(gidx) + 1)
        ids = src("sparse", int(gidx), n, val.size, 100)          # (R, nnz) 1-based
        vals.append(val)
        inds_o.append(ids)
        inds_d.append(np.asfortranarray(ids.T))                    # (nnz, R) column r = one sample()
    want = ho.dkl_randomized_sparse(vals, inds_o, ref["density"], ref["Q"])
    got = ctx.dkl_perm_csr(vals, inds_d, index_base=1)
    assert got.shape == (12, 100)
    assert _rel(got, want) < RTOL_DKL
    assert _rel(got, toy["full_KLD_rand"][:12]) < RTOL_DKL
    # dense form: full permutations of the expression row
    for k, gidx in enumerate(genes[:3]):
        perms = src("dense", int(gidx), n, n, 100)                 # (R, N)
        want_d = ho.dkl_randomized_dense(expr[[gidx]], perms[None], ref["density"], ref["Q"])[0]
        got_d = ctx.dkl_perm_dense(expr[gidx], np.asfortranarray(perms.T), index_base=1)
        assert _rel(got_d, want_d) < RTOL_DKL
        assert _rel(got_d, toy["fulld_KLD_rand"][k]) < RTOL_DKL


def test_state_errors(toy):
    from singlecellhaystack_b200.device import DeviceContext, HaystackLibraryError
    with DeviceContext(0) as c:
        with pytest.raises(HaystackLibraryError, match="hs_density first"):
            c.dkl_csr(np.array([0, 1]), np.array([0], dtype=np.int32), np.array([1.0]))
        with pytest.raises(ValueError, match="different number of columns"):
            c.density(np.zeros((10, 2)), np.zeros((3, 3)))


def test_edge_rows(ctx, toy):
    """Empty rows, a single-entry row, unsorted ids, and out-of-range ids (-> NaN, never a fault)."""
    xs, _, _ = ho.scale_coords(toy["coord"])
    ref = ho.density_stage(xs, toy["grid60"])
    ctx.density(xs, toy["grid60"])
    p = np.array([0, 0, 1, 4, 4, 5], dtype=np.int64)
    j = np.array([5, 300, 7, 600, 601], dtype=np.int32)
    x = np.array([2.0, 1.0, 3.0, 0.5, 1.0])
    got = ctx.dkl_csr(p, j, x)
    sp = ho.RowSparse(p=p[:5], j=j[:4], x=x[:4], shape=(4, 601))
    want = ho.dkl_observed(sp, ref["density"], ref["Q"])
    assert _rel(got[:4], want) < RTOL_DKL
    assert np.isnan(got[4])
    # ascending rows (the slab kernel keeps them): an id >= n_cells poisons only its own row
    p2 = np.array([0, 2, 4, 4, 7], dtype=np.int64)
    j2 = np.array([3, 599, 10, 601, 0, 1, 600], dtype=np.int32)
    x2 = np.array([1.0, 2.0, 3.0, 4.0, 0.5, 0.25, 2.0])
    got2 = ctx.dkl_csr(p2, j2, x2)
    sp2 = ho.RowSparse(p=np.array([0, 2, 2, 2, 5]), j=np.array([3, 599, 0, 1, 600], np.int32),
                       x=np.array([1.0, 2.0, 0.5, 0.25, 2.0]), shape=(4, 601))
    want2 = ho.dkl_observed(sp2, ref["density"], ref["Q"])
    assert _rel(got2[[0, 3]], want2[[0, 3]]) < RTOL_DKL
    assert np.isnan(got2[1])
    assert abs(got2[2] - want2[2]) < 1e-9          # empty row: P = pseudo everywhere, as the reference
    # a negative id cannot be dereferenced either
    got3 = ctx.dkl_csr(np.array([0, 2, 3], dtype=np.int64), np.array([-1, 5, 9], np.int32), np.array([1.0, 1.0, 1.0]))
    assert np.isnan(got3[0]) and np.isfinite(got3[1])
    assert ctx.dkl_csr(np.array([0], dtype=np.int64), np.zeros(0, np.int32), np.zeros(0)).shape == (0,)

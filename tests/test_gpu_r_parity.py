"""The device path against the reference's own printed results.

`set.seed(1234); haystack(dat.tsne, dat.expression)` (docs/articles/a01_toy_example.html:168) replayed through the
C ABI: the grid R's kmeans() picked, R's sample() permutations and R's cross-validation folds are regenerated from the
seed by the restated R generator (oracle/r_base, pinned in tests/test_r_parity_cpu.py) and fed to haystack() on the
B200.  north_star's tolerances, against R's printed numbers: D_KL 1e-5 relative, log10 p 1e-3 absolute; the
integer outputs (nearest grid point, selected genes) bit-exact against the oracle."""
import warnings
from pathlib import Path

import numpy as np
import pytest

from oracle import haystack_oracle as ho
from oracle import r_base as rb
from singlecellhaystack_b200 import haystack as hh
from vignette_table import VIGNETTE_TOP10

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "toy_r1234.npz"


def _r_stream(toy):
    """R's draws for the vignette call, in R's order (grid -> permutations gene-major -> folds)."""
    z = np.load(GOLD)
    coord = toy["coord"].astype(np.float64)
    rng = rb.RRng(1234)
    xs, _, _ = ho.scale_coords(coord)
    grid = rb.get_grid_points(xs, rng, "centroid", 100)
    np.testing.assert_array_equal(grid, z["grid_scaled"])
    perms = {int(g): rng.sample_int(601, 601, count=100) for g in z["genes_to_randomize"]}
    folds = (rng.sample(np.resize(np.arange(1, 11), 100)), rng.sample(np.resize(np.arange(1, 11), 100)))
    return z, grid, perms, folds


@pytest.mark.parametrize("sparse", [False, True])
def test_vignette_call_on_the_device(ctx, toy, sparse):
    z, grid, perms, folds = _r_stream(toy)
    coord, expr = toy["coord"].astype(np.float64), toy["expression"].astype(np.float64)
    names = [str(s) for s in toy["gene_names"]]
    if not sparse:
        def src(kind, gene, n, size, count):
            assert kind == "dense"
            return perms[gene]
        expression = expr
    else:
        # the dgCMatrix route of tests/testthat/test-haystack_sparse.R: values keep their order, the cells are
        # redrawn (R/haystack_continuous.R:275) -- a different use of the stream, so only the observed D_KL are R's
        import scipy.sparse as ssp
        rng = rb.RRng(99)

        def src(kind, gene, n, size, count):
            assert kind == "sparse"
            return rng.sample_int(n, size, count=count)
        expression = ssp.csc_matrix(expr)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = hh.haystack(coord, expression, grid_coord=grid, perm_source=src, cv_folds=folds, ctx=ctx,
                          gene_names=names, verbose=False)
    r = res["results"]
    for gene, dkl, logp, padj in VIGNETTE_TOP10:
        assert abs(r.loc[gene, "D_KL"] - dkl) <= 1e-5 * dkl + 5.1e-7, (gene, r.loc[gene, "D_KL"], dkl)
        if not sparse:
            assert abs(r.loc[gene, "log.p.vals"] - logp) <= 1e-3, (gene, r.loc[gene, "log.p.vals"], logp)
            assert abs(r.loc[gene, "log.p.adj"] - padj) <= 1e-3
    np.testing.assert_allclose(r["D_KL"].values, z["D_KL"], rtol=1e-5)
    np.testing.assert_allclose(res["info"]["grid_coordinates"], z["grid_coordinates"], rtol=1e-12)
    if not sparse:
        # (the sparse route computes the row sds by another formula; genes with tied coefficients of variation --
        # rows that are permutations of each other -- may then swap places in the selection, in R as well)
        np.testing.assert_array_equal(res["info"]["randomization"]["genes_to_randomize"] - 1, z["genes_to_randomize"])
        np.testing.assert_allclose(res["info"]["randomization"]["KLD_rand"], z["KLD_rand"], rtol=1e-5)
        np.testing.assert_allclose(r["log.p.vals"].values, z["log_p_vals"], atol=1e-3)
        top = hh.show_result_haystack(res, n=10)
        assert list(top.index) == [t[0] for t in VIGNETTE_TOP10]


def test_product_alone_replays_the_r_session(ctx, toy):
    """No oracle in the loop: haystack(rng = RRandom(1234)) draws the grid (Hartigan-Wong k-means), the 100 x 100
    permutations and the folds from R's stream inside the product and scores on the B200; the top-10 table is the one
    the reference printed."""
    from singlecellhaystack_b200.rcompat import RRandom
    coord, expr = toy["coord"].astype(np.float64), toy["expression"].astype(np.float64)
    names = [str(s) for s in toy["gene_names"]]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = hh.haystack(coord, expr, rng=RRandom(1234), ctx=ctx, gene_names=names, verbose=False)
    r = res["results"]
    for gene, dkl, logp, padj in VIGNETTE_TOP10:
        assert abs(r.loc[gene, "D_KL"] - dkl) <= 1e-5 * dkl + 5.1e-7
        assert abs(r.loc[gene, "log.p.vals"] - logp) <= 1e-3
        assert abs(r.loc[gene, "log.p.adj"] - padj) <= 1e-3
    assert list(hh.show_result_haystack(res, n=10).index) == [t[0] for t in VIGNETTE_TOP10]
    cv = res["info"]["randomization"]
    assert (cv["mean"]["cv.selected"]["df"], cv["sd"]["cv.selected"]["df"]) == (3, 5) if "cv.selected" in cv.get("mean", {}) else True

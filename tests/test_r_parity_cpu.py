"""The oracle pinned to the reference's own output.

The only numbers the reference prints for its fixture are in the rendered vignette
(/root/reference/docs/articles/a01_toy_example.html:129 "best df 3 / 5", :168 the top-10 table), produced by

    set.seed(1234); res <- haystack(dat.tsne, dat.expression); show_result_haystack(res, n = 10)

Every step of that call that draws random numbers lives in base R, not in the reference: kmeans() for the grid
(R/haystack_highD.R:83; Hartigan-Wong, nstart = 10), sample() for the randomizations (R/haystack_continuous.R:259) and
for the cross-validation folds (:437, twice).  oracle/r_base.{c,py} restate set.seed / Mersenne-Twister / sample.int /
Hartigan-Wong; with them the fp64 oracle reproduces the table to every printed digit, which pins the whole chain
RNG -> grid -> density -> D_KL -> randomization -> spline fit -> pnorm against a real R run (R >= 3.6 "Rejection"
sampling; the page was built 2023-08).  The table below is copied from the rendered page, not computed.
"""
import math

import numpy as np
import pytest

from oracle import haystack_oracle as ho
from oracle import r_base as rb

from vignette_table import VIGNETTE_TOP10  # noqa: E402  (tests/ is on sys.path: conftest.py)


def test_r_generator_known_answers():
    """set.seed(1234); runif(3) and set.seed(1234); sample.int(10) as any R >= 3.6 prints them; set.seed(42)
    likewise; "Rounding" (R < 3.6) differs."""
    g = rb.RRng(1234)
    np.testing.assert_allclose([g.unif_rand() for _ in range(3)], [0.1137034, 0.6222994, 0.6092747], atol=5e-8)
    assert rb.RRng(1234).sample_int(10).tolist() == [10, 6, 5, 4, 1, 8, 2, 7, 9, 3]
    assert rb.RRng(42).sample_int(10).tolist() == [1, 5, 10, 8, 2, 4, 6, 9, 7, 3]
    np.testing.assert_allclose(rb.RRng(42).unif_rand(), 0.914806, atol=5e-7)
    assert rb.RRng(1234, "Rounding").sample_int(10).tolist() != [10, 6, 5, 4, 1, 8, 2, 7, 9, 3]
    # a full permutation is a permutation; partial draws are distinct
    p = rb.RRng(7).sample_int(601)
    assert sorted(p.tolist()) == list(range(1, 602))
    q = rb.RRng(7).sample_int(1000, 100, count=5)
    assert all(len(set(r.tolist())) == 100 for r in q)


def run_reference_call(toy, seed=1234):
    """haystack(dat.tsne, dat.expression) after set.seed(seed), R's draws in R's order:
    grid (10 x sample.int(601, 100)) -> randomizations gene-major (100 x 100 x sample.int(601)) -> folds of the
    mean model -> folds of the sd model."""
    coord, expr = toy["coord"].astype(np.float64), toy["expression"].astype(np.float64)
    rng = rb.RRng(seed)
    xs, center, sc = ho.scale_coords(coord)                                   # R/haystack_continuous.R:139-145
    grid = rb.get_grid_points(xs, rng, "centroid", 100)                       # :157-166
    # the oracle takes the draws as inputs; make them in R's stream order first
    mean, sd = ho.row_mean_sd(expr)
    cv = (sd + 1e-300) / (mean + 1e-300)                                      # :82-83, :222
    genes = ho.select_genes_to_randomize(cv, 100, "heavytails")              # :227-236
    perms = {int(g): rng.sample_int(601, 601, count=100) for g in genes}     # :259 sample(x = row), gene-major
    folds_mean = rng.sample(np.resize(np.arange(1, 11), 100))                 # :437, mean model
    folds_sd = rng.sample(np.resize(np.arange(1, 11), 100))                   # :437, sd model
    res = ho.haystack_continuous_highD(coord, expr, grid, lambda kind, gene, n, size, count: perms[gene],
                                       folds_mean, folds_sd)
    return res, dict(grid=grid, perms=perms, genes=genes, folds_mean=folds_mean, folds_sd=folds_sd)


@pytest.fixture(scope="module")
def r1234(toy):
    return run_reference_call(toy)


def test_vignette_table_reproduced_to_printed_digits(toy, r1234):
    res, _ = r1234
    names = [str(s) for s in toy["gene_names"]]
    idx = {n: i for i, n in enumerate(names)}
    for gene, dkl, logp, padj in VIGNETTE_TOP10:
        i = idx[gene]
        assert abs(res["D_KL"][i] - dkl) <= 5.1e-7 * max(1.0, abs(dkl)), (gene, res["D_KL"][i], dkl)     # 7 sig. digits
        assert abs(res["log_p_vals"][i] - logp) <= 5.1e-6, (gene, res["log_p_vals"][i], logp)           # 5 decimals
        assert abs(res["log_p_adj"][i] - padj) <= 5.1e-6, (gene, res["log_p_adj"][i], padj)
    order = np.argsort(res["log_p_vals"], kind="stable")[:10]
    assert [names[i] for i in order] == [t[0] for t in VIGNETTE_TOP10]
    # docs/articles/a01_toy_example.html:129: "best RMSD : 0.09 / best df : 3" and "0.02 / 5"
    fit = res["fit"]
    assert fit["mean_model"]["best_df"] == 3 and round(fit["mean_model"]["best_rmsd"], 3) == 0.09
    assert fit["sd_model"]["best_df"] == 5 and round(fit["sd_model"]["best_rmsd"], 3) == 0.02
    assert abs((res["log_p_adj"][idx["gene_79"]] - res["log_p_vals"][idx["gene_79"]]) - math.log10(500)) < 1e-12


def test_committed_r_golden_matches_regeneration(toy, r1234):
    """tests/golden/toy_r1234.npz (oracle/make_golden.py) holds the R-stream inputs and the oracle's results for
    the vignette call; the GPU tests load it.  It must be what the restatement regenerates today."""
    from pathlib import Path
    res, inp = r1234
    z = np.load(Path(__file__).resolve().parent / "golden" / "toy_r1234.npz")
    np.testing.assert_array_equal(z["grid_scaled"], inp["grid"])
    np.testing.assert_array_equal(z["genes_to_randomize"], inp["genes"])
    np.testing.assert_array_equal(z["folds_mean"], inp["folds_mean"])
    np.testing.assert_array_equal(z["folds_sd"], inp["folds_sd"])
    np.testing.assert_allclose(z["D_KL"], res["D_KL"], rtol=1e-13)
    np.testing.assert_allclose(z["log_p_vals"], res["log_p_vals"], rtol=1e-10)
    chk = np.array([int(inp["perms"][int(g)].astype(np.int64).dot(np.arange(1, 602)).sum() % (2 ** 31)) for g in inp["genes"]])
    np.testing.assert_array_equal(z["perm_checksums"], chk)


def test_kmeans_hartigan_wong_properties():
    """Hartigan-Wong on separated blobs: recovers them, never leaves a cluster empty, total within-SS no worse
    than the initial assignment's, and the nstart logic keeps the best run."""
    rs = np.random.default_rng(0)
    centres = np.array([[0, 0], [10, 0], [0, 10], [10, 10], [5, 5.0]])
    x = np.concatenate([c + rs.normal(scale=0.3, size=(40, 2)) for c in centres])
    out = rb.kmeans(x, 5, rb.RRng(1), iter_max=10, nstart=10)
    assert sorted(out["size"].tolist()) == [40] * 5 and out["ifault"] == 0
    for c in centres:      # every blob has a centre on it
        assert np.min(np.abs(out["centers"] - c).max(axis=1)) < 0.2
    one = rb.kmeans(x, 5, rb.RRng(1), iter_max=10, nstart=1)
    assert out["withinss"].sum() <= one["withinss"].sum() + 1e-9
    # centres are the means of their clusters
    for l in range(5):
        np.testing.assert_allclose(out["centers"][l], x[out["cluster"] == l + 1].mean(axis=0), rtol=1e-12, atol=1e-12)

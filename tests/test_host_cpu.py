"""CPU tests of the host layer (everything above the C ABI that needs no GPU)."""
import math
import warnings

import numpy as np
import pytest

from oracle import haystack_oracle as ho
from singlecellhaystack_b200 import haystack as hh
from singlecellhaystack_b200 import sparse as sp
from singlecellhaystack_b200 import splinefit as sf
from singlecellhaystack_b200 import synth


def test_select_genes_matches_oracle():
    rng = np.random.default_rng(3)
    for m in (118, 500, 12382, 16811, 30000):
        cv = rng.gamma(2.0, size=m)
        cv[rng.integers(0, m, 20)] = cv[0]            # ties -> stable order matters
        for method in ("heavytails", "uniform"):
            np.testing.assert_array_equal(hh.select_genes_to_randomize(cv, 100, method),
                                          ho.select_genes_to_randomize(cv, 100, method))
    with pytest.raises(ValueError):
        hh.select_genes_to_randomize(np.arange(200.0), 100, "nope")


def test_row_mean_sd_sparse_equals_dense(toy):
    expr = toy["expression"]
    m0, s0 = hh.row_mean_sd(expr)
    m1, s1 = hh.row_mean_sd(sp.as_RsparseMatrix(expr))
    np.testing.assert_allclose(m1, m0, rtol=1e-13)
    np.testing.assert_allclose(s1, s0, rtol=1e-12)
    m2, s2 = ho.row_mean_sd(ho.RowSparse.from_dense(expr))
    np.testing.assert_allclose(m1, m2, rtol=1e-13)
    np.testing.assert_allclose(s1, s2, rtol=1e-12)


def test_scale_matches_oracle(toy):
    a, c, s = hh.scale(toy["coord"])
    b, c2, s2 = ho.scale_coords(toy["coord"])
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(c, c2)
    np.testing.assert_array_equal(s, s2)


def test_sparse_conversions(toy):
    import scipy.sparse as ssp
    expr = toy["expression"]
    r = sp.as_RsparseMatrix(expr)
    c = sp.as_CsparseMatrix(expr)
    r2 = sp.as_RsparseMatrix(c)                         # as(dgCMatrix, "RsparseMatrix"), :113-116
    np.testing.assert_array_equal(r.p, r2.p)
    np.testing.assert_array_equal(r.j, r2.j)
    np.testing.assert_array_equal(r.x, r2.x)
    r3 = sp.as_RsparseMatrix(ssp.csc_matrix(expr))
    np.testing.assert_array_equal(r.j, r3.j)
    np.testing.assert_array_equal(r.toarray(), expr)
    # reference known answers, tests/testthat/test-extract_row.R:13-20
    m = sp.as_RsparseMatrix(np.array([[1.0, 3.0], [2.0, 4.0]]))
    assert sp.extract_row_dgRMatrix(m, 1).tolist() == [1.0, 3.0]
    row = sp.extract_row_dgRMatrix_as_sparse(m, 2)
    assert row["ind"].tolist() == [1, 2] and row["val"].tolist() == [2.0, 4.0]
    keep = np.ones(500, bool)
    keep[[0, 17, 499]] = False
    sub = r.row_subset(keep)
    np.testing.assert_array_equal(sub.toarray(), expr[keep])


def test_ns_fit_matches_oracle():
    rng = np.random.default_rng(0)
    x = np.sort(rng.normal(size=100))
    y = np.sin(x) + 0.1 * rng.normal(size=100)
    sets = synth.cv_folds(1, 0, 100)
    o = ho.get_model_cv_ns(x, y, sets)
    m = sf.get_model_cv_ns(x, y, sets=sets)
    assert m["info"]["cv_selected"]["df"] == o["best_df"]
    xn = np.linspace(x.min() - 1, x.max() + 1, 57)           # includes linear extrapolation on both sides
    np.testing.assert_allclose(m["model"].predict(xn), o["predict"](xn), rtol=0, atol=1e-10)


def test_bs_design_properties():
    x = np.linspace(-1, 2, 41)
    b = np.array([-1.03, 2.03])
    for degree in (1, 2, 3, 5):
        knots = np.array([0.0, 0.7, 1.1])
        full = sf.bs_design(x, knots, b, degree)
        assert full.shape == (41, knots.size + degree)        # df = length(knots) + degree, intercept dropped
        assert (full >= -1e-12).all() and (full.sum(axis=1) <= 1 + 1e-12).all()   # partition of unity minus col 1
    # a cubic is reproduced exactly by lm(y ~ bs(x, df = 5))
    y = 1 + 2 * x - 0.5 * x ** 2 + 0.25 * x ** 3
    m = sf.SplineModel(x, y, 5, b, "bs", 3)
    np.testing.assert_allclose(m.predict(x), y, atol=1e-10)
    m2 = sf.get_model_cv_bs(x, y + 0.0, sets=(np.arange(41) % 10) + 1)
    assert m2["info"]["cv_selected"]["rmsd"] < 1e-8


def test_log_p_matches_golden(toy):
    """The host fit on the oracle's own D_KL inputs reproduces the golden log10 p-values (tolerance of
    north_star: 1e-3 absolute)."""
    seed = int(toy["seed"])
    expr = toy["expression"]
    # the goldens were generated from the sparse form of the fixture; the toy counts contain genes with
    # (nearly) tied coefficients of variation, so the selection is only reproducible from the same row
    # statistics -- in the drop-in those are R's own rowMeans / rowSds
    mean, sd = ho.row_mean_sd(ho.RowSparse.from_dense(expr))
    cv = (sd + 1e-300) / (mean + 1e-300)
    m1, s1 = hh.row_mean_sd(sp.as_RsparseMatrix(expr))
    np.testing.assert_allclose((s1 + 1e-300) / (m1 + 1e-300), cv, rtol=1e-12)
    genes = hh.select_genes_to_randomize(cv, 100)
    np.testing.assert_array_equal(genes, toy["full_genes_to_randomize"])       # integer output: bit-exact
    pv = sf.get_log_p_D_KL_continuous(toy["dkl100"], toy["full_KLD_rand"], cv, cv[genes], "ns",
                                      sets_mean=synth.cv_folds(seed, 0, 100), sets_sd=synth.cv_folds(seed, 1, 100))
    assert np.max(np.abs(pv["fitted"] - toy["full_log_p"])) < 1e-6
    assert [pv["info"]["mean"]["cv_selected"]["df"], pv["info"]["sd"]["cv_selected"]["df"]] == toy["full_best_df"].tolist()


def test_show_result_haystack():
    import pandas as pd
    res = hh.Haystack(results=pd.DataFrame({"D_KL": [1.0, 2.0, 3.0, 0.5], "log.p.vals": [-3.0, -10.0, -1.0, -10.0],
                                            "log.p.adj": [-2.0, -9.0, 0.0, -9.0]}, index=["a", "b", "c", "d"]), info={})
    assert hh.show_result_haystack(res).index.tolist() == ["b", "d", "a", "c"]        # stable order()
    assert hh.show_result_haystack(res, n=2).index.tolist() == ["b", "d"]
    assert hh.show_result_haystack(res, p_value_threshold=1e-3).index.tolist() == ["b", "d", "a"]
    assert hh.show_result_haystack(res, gene=["c", "a"]).index.tolist() == ["a", "c"]
    with pytest.warns(UserWarning, match="larger than the number of rows"):
        assert len(hh.show_result_haystack(res, n=10)) == 4
    with pytest.raises(ValueError, match="between 0 and 1"):
        hh.show_result_haystack(res, p_value_threshold=2)
    with pytest.raises(TypeError, match="should be an integer"):
        hh.show_result_haystack(res, n="3")
    with pytest.raises(ValueError, match="Results seem to be missing"):
        hh.show_result_haystack(hh.Haystack(info={}))


def test_validation_messages_before_any_device_call(toy):
    """The reference's stop() texts (R/haystack_continuous.R:37-46, 68-69, 89-111), raised on the host."""
    coord, expr = toy["coord"], toy["expression"]
    f = lambda *a, **k: hh.haystack_continuous_highD(*a, verbose=False, **k)
    with pytest.raises(ValueError, match="different number of columns"):
        f(coord, expr, grid_coord=np.zeros((10, 3)))
    with pytest.raises(ValueError, match="average signal < 0"):
        f(coord, -expr)
    with pytest.raises(ValueError, match="number of columns in 'expression'"):
        f(coord[:100], expr)
    with pytest.raises(ValueError, match="very high"):
        f(coord, expr, grid_points=601)
    with pytest.raises(ValueError, match="'weights.advanced.Q' must be the same"):
        f(coord, expr, weights_advanced_Q=np.ones(5))
    with pytest.raises(ValueError, match="'ns' or 'bs'"):
        f(coord, expr, spline_method="cubic")
    with pytest.raises(TypeError, match="must be a numeric matrix"):
        f(np.array([["a", "b"]] * 601), expr)
    with pytest.raises(TypeError, match="'expression' must be a matrix"):
        f(coord, np.array([["x"] * 601] * 3))
    with pytest.raises(TypeError, match="missing"):
        hh.haystack(coord)


def test_no_cpu_fallback(toy):
    """On a machine without a B200 the product path fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from singlecellhaystack_b200.device import HaystackLibraryError
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(HaystackLibraryError, match="no CPU fallback"):
            hh.haystack(toy["coord"], toy["expression"], grid_points=60, verbose=False, rng=1)


def test_package_never_imports_oracle():
    import pathlib
    import re
    pkg = pathlib.Path(hh.__file__).parent
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        txt = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
        assert "oracle/" not in txt and "haystack_oracle" not in txt, f


def test_selection_ranks_are_the_positions_the_selection_takes():
    """selection_ranks (what the device selection hs_select_by_cv_dev is given) + a stable argsort == the host selection,
    including tied and NaN coefficients; the heavytails guard of R/haystack_continuous.R:233-235."""
    from singlecellhaystack_b200 import haystack as hk
    rng = np.random.default_rng(1)
    for m in (18, 500, 12382, 30000):
        cv = rng.gamma(2.0, 1.0, size=m)
        cv[rng.choice(m, m // 7, replace=False)] = cv[0]          # ties
        if m > 100:
            cv[rng.choice(m, 5, replace=False)] = np.nan          # order(): NA last
        for method in ("heavytails", "uniform"):
            s_ = min(100, m)
            ranks = hk.selection_ranks(m, s_, method)
            assert ranks.min() >= 1 and ranks.max() <= m and ranks.size == s_
            np.testing.assert_array_equal(np.argsort(cv, kind="stable")[ranks - 1], hk.select_genes_to_randomize(cv, s_, method))
    with pytest.raises(ValueError, match="at least 18"):
        hk.selection_ranks(500, 17, "heavytails")
    with pytest.raises(ValueError, match="uniform' or 'heavytails"):
        hk.selection_ranks(500, 100, "other")


def test_seeding_grid_host_path_equals_the_oracle(toy):
    """get_grid_points(method = "seeding") on R's stream: the library's draws (hs_r_sample_int / hs_r_sample_prob_one)
    and the host distance passes give the grid of the oracle's restatement of R/haystack_highD.R:87-129."""
    from oracle import haystack_oracle as ho
    from oracle import r_base
    from singlecellhaystack_b200 import grid as gr
    from singlecellhaystack_b200.rcompat import RRandom
    xs, _, _ = ho.scale_coords(toy["coord"])
    for g in (5, 40):
        got = gr.get_grid_points(xs, "seeding", g, rng=RRandom(42))
        want = r_base.get_grid_points(xs, r_base.RRng(42), "seeding", g)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)

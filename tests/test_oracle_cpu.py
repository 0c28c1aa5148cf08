"""CPU tests of the oracle (the parity checker) against everything the reference tree pins:
the decoded toy fixture, the two exact-value tests of tests/testthat/test-extract_row.R, the rendered
vignette table (loose: it depends on R's RNG + Hartigan-Wong k-means), and the committed goldens."""
import math
import warnings

import numpy as np
import pytest

from oracle import haystack_oracle as ho
from singlecellhaystack_b200 import synth

# docs/articles/a01_toy_example.html:168 (set.seed(1234); haystack(dat.tsne, dat.expression))
VIGNETTE = [("gene_79", 2.447641, -39.94871), ("gene_497", 2.271242, -39.84546), ("gene_242", 1.742783, -35.91184),
            ("gene_275", 1.819669, -35.88741), ("gene_62", 2.174074, -35.63854), ("gene_71", 2.546493, -34.83522),
            ("gene_381", 2.733446, -34.38544), ("gene_351", 1.844673, -33.45724), ("gene_479", 2.343509, -31.98940),
            ("gene_300", 2.097546, -30.42637)]


def test_toy_fixture_decode_facts(toy):
    """dat.tsne 601x2 (mean 0, sd 16.8356 / 14.8677), dat.expression 500x601, max 44,
    16.163 % non-zero."""
    coord, expr = toy["coord"], toy["expression"]
    assert coord.shape == (601, 2) and expr.shape == (500, 601)
    assert np.allclose(coord.mean(axis=0), 0, atol=1e-10)
    assert np.allclose(coord.std(axis=0, ddof=1), [16.8356, 14.8677], atol=5e-4)
    assert expr.max() == 44 and expr.min() == 0
    assert abs((expr != 0).mean() * 100 - 16.163) < 5e-3
    assert toy["gene_names"][0] == "gene_1" and toy["gene_names"][-1] == "gene_500"
    assert (expr.std(axis=1, ddof=1) > 0).all()          # no zero-variance genes


def test_reference_extract_row_known_answers():
    """tests/testthat/test-extract_row.R:3-10 and :13-20 -- the only exact-value assertions of the reference."""
    lg = ho.RowSparse.from_dense(np.array([[1.0, 0.0], [0.0, 1.0]]))          # matrix(c(TRUE,FALSE,FALSE,TRUE), 2)
    assert ho.extract_row_lgRMatrix(lg, 1).tolist() == [True, False]
    dg = ho.RowSparse.from_dense(np.array([[1.0, 3.0], [2.0, 4.0]]))          # matrix(1:4, 2) is column-major
    assert ho.extract_row_dgRMatrix(dg, 1).tolist() == [1.0, 3.0]
    ind, val = ho.extract_row_dgRMatrix_as_sparse(dg, 2)
    assert ind.tolist() == [1, 2] and val.tolist() == [2.0, 4.0]              # 1-based, R/sparse.R:39-49
    empty = ho.RowSparse(p=np.array([0, 0, 1]), j=np.array([1], np.int32), x=np.array([5.0]), shape=(2, 3))
    ind, val = ho.extract_row_dgRMatrix_as_sparse(empty, 1)
    assert ind.size == 0 and val.size == 0


def test_oracle_reproduces_committed_goldens(toy):
    xs, _, _ = ho.scale_coords(toy["coord"])
    for g in (60, 100):
        ds = ho.density_stage(xs, toy[f"grid{g}"])
        assert ds["bandwidth"] == float(toy[f"bandwidth{g}"])
        np.testing.assert_array_equal(ds["nearest"], toy[f"nearest{g}"])
        np.testing.assert_allclose(ds["Q"], toy[f"Q{g}"], rtol=1e-13)
        dkl = ho.dkl_observed(toy["expression"], ds["density"], ds["Q"])
        np.testing.assert_allclose(dkl, toy[f"dkl{g}"], rtol=1e-12)


def test_dense_and_sparse_branches_agree(toy):
    xs, _, _ = ho.scale_coords(toy["coord"])
    ds = ho.density_stage(xs, toy["grid60"])
    sp = ho.RowSparse.from_dense(toy["expression"])
    np.testing.assert_allclose(ho.dkl_observed(sp, ds["density"], ds["Q"]), toy["dkl60"], rtol=1e-12)
    np.testing.assert_array_equal(sp.to_dense(), toy["expression"])


def test_vignette_table_loose(toy):
    """Loose bound on the printed D_KL values with a grid that is NOT R's (the package's Lloyd k-means on a numpy
    stream): the scores move by ~10 % with the grid.  The exact pin -- R's own grid and draws, every printed digit --
    is tests/test_r_parity_cpu.py."""
    from singlecellhaystack_b200 import grid as G
    xs, _, _ = ho.scale_coords(toy["coord"])
    names = list(toy["gene_names"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        grid = G.get_grid_points(xs, "centroid", 100, rng=1234)
    ds = ho.density_stage(xs, grid)
    dkl = ho.dkl_observed(toy["expression"], ds["density"], ds["Q"])
    dev = np.array([dkl[names.index(n)] / v - 1 for n, v, _ in VIGNETTE])
    assert np.mean(np.abs(dev)) < 0.12 and np.max(np.abs(dev)) < 0.25
    # the vignette's top genes are strongly structured here too: all in our top 15 % by D_KL
    rank = np.argsort(np.argsort(-dkl))
    assert all(rank[names.index(n)] < 75 for n, _, _ in VIGNETTE)


def test_bonferroni_identity_from_vignette(toy):
    """log.p.adj - log.p.vals = log10(500) = 2.69897 (R/haystack_continuous.R:301-302; vignette rows:
    -39.94871 -> -37.24974)."""
    assert abs((-37.24974 - -39.94871) - math.log10(500)) < 1e-5
    np.testing.assert_allclose(toy["full_log_p_adj"], np.minimum(0, toy["full_log_p"] + math.log10(500)), rtol=1e-12)
    # our goldens call the same genes extreme: vignette's top 10 all have log10 p < -15 here
    names = list(toy["gene_names"])
    assert all(toy["full_log_p"][names.index(n)] < -15 for n, _, _ in VIGNETTE)


def test_kl_invariants(toy):
    xs, _, _ = ho.scale_coords(toy["coord"])
    ds = ho.density_stage(xs, toy["grid60"])
    dens, q = ds["density"], ds["Q"]
    rng = np.random.default_rng(0)
    w = rng.gamma(1.0, size=601) * (rng.random(601) < 0.3)
    d0 = ho.get_D_KL_continuous_highD(w, dens, q, ho.PSEUDO)
    assert d0 >= 0
    assert abs(ho.get_D_KL_continuous_highD(7.5 * w, dens, q, ho.PSEUDO) - d0) < 1e-12      # scale invariance
    ind = np.nonzero(w)[0] + 1
    perm = rng.permutation(ind.size)
    d1 = ho.get_D_KL_continuous_highD_SPARSE(ind[perm], w[ind - 1][perm], dens, q, ho.PSEUDO)  # order invariance
    assert abs(d1 - d0) < 1e-12
    assert abs(ho.get_D_KL_continuous_highD(np.ones(601), dens, q, ho.PSEUDO)) < 1e-12        # P == Q  ->  0
    # weights.advanced.Q only reweights Q (R/haystack_continuous.R:182)
    wq = 0.5 + rng.random(601)
    dsw = ho.density_stage(xs, toy["grid60"], wq)
    assert abs(ho.get_D_KL_continuous_highD(wq, dsw["density"], dsw["Q"], ho.PSEUDO)) < 1e-12


def test_fp32_traps_documented_in_survey(toy):
    """the pseudocount underflows in fp32 and many toy densities are below fp32 range."""
    assert np.float32(ho.PSEUDO) == 0
    xs, _, _ = ho.scale_coords(toy["coord"])
    ds = ho.density_stage(xs, toy["grid100"])
    assert (ds["density"] < np.finfo(np.float32).tiny).mean() > 0.5


def test_gene_selection_integer_exact():
    """R/haystack_continuous.R:227-236 at the sizes of the configs: 100 distinct indices, tails included."""
    for m in (500, 12382, 16811, 30000):
        cv = synth.u24(synth.hash3(5, np.int64(3), np.arange(m, dtype=np.int64))).astype(np.float64)
        sel = ho.select_genes_to_randomize(cv, 100, "heavytails")
        o = np.argsort(cv, kind="stable")
        assert len(np.unique(sel)) == 100
        assert sel[:9].tolist() == o[:9].tolist() and sel[-9:].tolist() == o[-9:].tolist()
        pos = [int(10 + i * ((m - 9 - 10) / 81)) for i in range(82)]
        pos[-1] = m - 9
        assert sel[9:91].tolist() == o[np.array(pos) - 1].tolist()
        selu = ho.select_genes_to_randomize(cv, 100, "uniform")
        assert selu[0] == o[0] and selu[-1] == o[-1]


def test_randomization_semantics(toy):
    """:250-283: dense = permute the whole vector; sparse = redraw distinct cells, values keep order."""
    xs, _, _ = ho.scale_coords(toy["coord"])
    ds = ho.density_stage(xs, toy["grid60"])
    expr = toy["expression"]
    perm = synth.sample_int(1, 2, 601)
    d_dense = ho.dkl_randomized_dense(expr[[3]], perm[None, None, :], ds["density"], ds["Q"])[0, 0]
    assert abs(d_dense - ho.get_D_KL_continuous_highD(expr[3][perm - 1], ds["density"], ds["Q"], ho.PSEUDO)) == 0
    sp = ho.RowSparse.from_dense(expr)
    _, val = ho.extract_row_dgRMatrix_as_sparse(sp, 4)
    ids = synth.sample_int(1, 3, 601, val.size)
    d_sp = ho.dkl_randomized_sparse([val], [ids[None, :]], ds["density"], ds["Q"])[0, 0]
    w = np.zeros(601)
    w[ids - 1] = val
    assert abs(d_sp - ho.get_D_KL_continuous_highD(w, ds["density"], ds["Q"], ho.PSEUDO)) < 1e-12

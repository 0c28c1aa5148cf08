"""GPU parity of the whole drop-in path (host layer -> C ABI -> CUDA) against the committed goldens and
the oracle.  Tolerances are north_star's: integer outputs bit-exact, D_KL 1e-5 relative, log10 p 1e-3."""
import warnings

import numpy as np
import pytest

from oracle import haystack_oracle as ho
from singlecellhaystack_b200 import haystack as hh
from singlecellhaystack_b200 import sparse as sp
from singlecellhaystack_b200 import synth

pytestmark = pytest.mark.gpu

RTOL_DKL = 1e-5
ATOL_LOGP = 1e-3


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def _oracle_stats(expr_sparse):
    return ho.row_mean_sd(expr_sparse)


@pytest.mark.parametrize("form", ["sparse", "dense"])
def test_toy_full_path_matches_golden(ctx, toy, form):
    seed = int(toy["seed"])
    coord, expr = toy["coord"], toy["expression"]
    xs, center, sc = ho.scale_coords(coord)
    stats = ho.row_mean_sd(ho.RowSparse.from_dense(expr)) if form == "sparse" else ho.row_mean_sd(expr)
    e = sp.as_CsparseMatrix(expr) if form == "sparse" else expr      # dgCMatrix in, as users hold it
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = hh.haystack(coord, e, grid_coord=toy["grid100"], gene_names=list(toy["gene_names"]),
                          perm_source=synth.perm_source(seed), row_stats=stats,
                          cv_folds=(synth.cv_folds(seed, 0, 100), synth.cv_folds(seed, 1, 100)), ctx=ctx,
                          verbose=False)
    r = res["results"]
    assert list(r.columns) == ["D_KL", "log.p.vals", "log.p.adj"] and r.shape == (500, 3)
    assert _rel(r["D_KL"].values, toy["dkl100"]) < RTOL_DKL
    info = res["info"]
    assert info["method"] == "continuous_highD"
    gold_rand = toy["full_KLD_rand"] if form == "sparse" else toy["fulld_KLD_rand"]
    gold_logp = toy["full_log_p"] if form == "sparse" else toy["fulld_log_p"]
    if form == "sparse":
        np.testing.assert_array_equal(info["randomization"]["genes_to_randomize"], toy["full_genes_to_randomize"] + 1)
    assert _rel(info["randomization"]["KLD_rand"], gold_rand) < RTOL_DKL
    assert np.max(np.abs(r["log.p.vals"].values - gold_logp)) < ATOL_LOGP
    np.testing.assert_allclose(r["log.p.adj"].values, np.minimum(0, r["log.p.vals"].values + np.log10(500)))
    np.testing.assert_allclose(info["grid_coordinates"], toy["grid100"] * sc + center, rtol=1e-14)
    np.testing.assert_allclose(info["coord_mean"], center)
    assert info["densities"].shape == (601, 100)
    top = hh.show_result_haystack(res, n=10)
    assert top["log.p.vals"].is_monotonic_increasing and len(top) == 10


def test_reference_structural_assertions(ctx, toy):
    """tests/testthat/test-haystack_continuous.R:6-15 and test-haystack_sparse.R:4-15: grid.points = 60,
    own grid selection and random draws; only structure is asserted, as in the reference."""
    for method in ("centroid", "seeding"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = hh.haystack_continuous_highD(toy["coord"], sp.as_CsparseMatrix(toy["expression"]), grid_points=60,
                                               grid_method=method, rng=123, ctx=ctx, verbose=False,
                                               gene_names=list(toy["gene_names"]))
        assert isinstance(res, dict) and isinstance(res, hh.Haystack)
        assert list(res.keys()) == ["results", "info"]
        assert res["results"].shape == (500, 3)
        assert res["info"]["grid_coordinates"].shape == (60, 2)
        assert np.isfinite(res["results"]["D_KL"].values).all()
        # the vignette's strongest genes come out on top with any reasonable grid
        top = set(hh.show_result_haystack(res, n=50).index)
        assert len(top & {"gene_79", "gene_497", "gene_242", "gene_275", "gene_62"}) >= 4


@pytest.mark.parametrize("n,d,g,m,rho", [(20000, 10, 100, 256, 0.08), (5000, 1, 37, 128, 0.3), (3000, 50, 130, 64, 0.02)])
def test_synthetic_csr_parity(ctx, n, d, g, m, rho):
    x, labels = synth.make_embedding(n, d)
    xs, _, _ = synth.scale_coords(x)
    grid = synth.make_grid(xs, g)
    p, j, v = synth.make_csr_numpy(m, n, rho, labels)
    ref = ho.density_stage(xs, grid)
    out = ctx.density(xs, grid, want_nearest=True)
    assert abs(out["bandwidth"] - ref["bandwidth"]) <= 1e-12 * ref["bandwidth"]
    np.testing.assert_array_equal(out["nearest"], ref["nearest"])
    np.testing.assert_allclose(out["Q"], ref["Q"], rtol=1e-9)
    want = ho.dkl_observed(ho.RowSparse(p=p, j=j, x=v, shape=(m, n)), ref["density"], ref["Q"])
    got = ctx.dkl_csr(p, j, v)
    assert _rel(got, want) < RTOL_DKL
    # float32 host values (hs_dkl_csr_f32): the device rounds doubles to fp32 the same way -> identical bits
    np.testing.assert_array_equal(ctx.dkl_csr(p, j, v.astype(np.float32)), got)
    # dense form of the same rows
    dense = ho.RowSparse(p=p, j=j, x=v, shape=(m, n)).to_dense()
    got_d = ctx.dkl_dense(dense[:32])
    assert _rel(got_d, want[:32]) < RTOL_DKL
    # randomized, a few genes
    src = synth.perm_source(7)
    vals, ids_o, ids_d = [], [], []
    for gi in range(4):
        val = v[p[gi]:p[gi + 1]]
        ids = src("sparse", gi, n, val.size, 10)
        vals.append(val)
        ids_o.append(ids)
        ids_d.append(np.asfortranarray(ids.T))
    want_r = ho.dkl_randomized_sparse(vals, ids_o, ref["density"], ref["Q"])
    got_r = ctx.dkl_perm_csr(vals, ids_d, index_base=1)
    assert _rel(got_r, want_r) < RTOL_DKL
    gp = np.concatenate([[0], np.cumsum([len(a_) for a_ in vals])]).astype(np.int64)
    flat32 = ctx.dkl_perm_csr_flat(gp, np.concatenate(vals).astype(np.float32),
                                   np.concatenate([a_.ravel(order="F") for a_ in ids_d]), 10, index_base=1)
    assert _rel(flat32, want_r) < RTOL_DKL


def test_sharded_layer_on_device(ctx, toy):
    """The gene-sharded host layer with the real DeviceContext as backend (single rank here; the two-rank
    logic is covered on CPU over gloo, the NCCL broadcast by bench.py --gpus N)."""
    from singlecellhaystack_b200.distributed import haystack_continuous_highD_sharded
    seed = int(toy["seed"])
    expr = toy["expression"]
    stats = ho.row_mean_sd(ho.RowSparse.from_dense(expr))
    res = haystack_continuous_highD_sharded(
        toy["coord"], sp.as_CsparseMatrix(expr), ctx, toy["grid100"], perm_source=synth.perm_source(seed),
        row_stats=stats, cv_folds=(synth.cv_folds(seed, 0, 100), synth.cv_folds(seed, 1, 100)))
    assert _rel(res["results"]["D_KL"].values, toy["dkl100"]) < RTOL_DKL
    assert _rel(res["info"]["randomization"]["KLD_rand"], toy["full_KLD_rand"]) < RTOL_DKL
    assert np.max(np.abs(res["results"]["log.p.vals"].values - toy["full_log_p"])) < ATOL_LOGP


def test_density_buffers_roundtrip(ctx, toy):
    """hs_density_buffers / hs_density_commit: a second context adopts the first one's resident state
    (what the NCCL broadcast does between ranks) and scores identically."""
    import torch
    from singlecellhaystack_b200.device import DeviceContext, _wrap_device_ptr
    xs, _, _ = ho.scale_coords(toy["coord"])
    ctx.density(xs, toy["grid60"])
    sp_ = ho.RowSparse.from_dense(toy["expression"])
    want = ctx.dkl_csr(sp_.p, sp_.j, sp_.x)
    with DeviceContext(0) as other:
        d0, db, q0, qb = ctx.density_buffers(601, 60)
        d1, db1, q1, qb1 = other.density_buffers(601, 60)
        assert (db, qb) == (db1, qb1) and d0 != d1
        dev = torch.device("cuda", 0)
        ctx.synchronize()
        _wrap_device_ptr(d1, db, dev).copy_(_wrap_device_ptr(d0, db, dev))
        _wrap_device_ptr(q1, qb, dev).copy_(_wrap_device_ptr(q0, qb, dev))
        torch.cuda.synchronize()
        other.density_commit(601, 60)
        assert other.shape == (601, 60)
        np.testing.assert_array_equal(other.dkl_csr(sp_.p, sp_.j, sp_.x), want)


def test_tile_path_is_the_default_csr_path(monkeypatch):
    """The CSR contraction runs on the tensor-core tile kernels by default (csrc/hs_tiles.cu); HS_TILE_DISABLE=1
    selects the SIMT slab kernel.  Same tolerance, different summation order."""
    from singlecellhaystack_b200.device import DeviceContext
    n, d, g, m = 20000, 10, 100, 600
    x, labels = synth.make_embedding(n, d)
    xs, _, _ = synth.scale_coords(x)
    grid = synth.make_grid(xs, g)
    p, j, v = synth.make_csr_numpy(m, n, 0.15, labels)
    ref = ho.density_stage(xs, grid)
    want = ho.dkl_observed(ho.RowSparse(p=p, j=j, x=v, shape=(m, n)), ref["density"], ref["Q"])
    res = {}
    for name in ("tiles", "slab"):
        if name == "slab":
            monkeypatch.setenv("HS_TILE_DISABLE", "1")
        else:
            monkeypatch.delenv("HS_TILE_DISABLE", raising=False)
        with DeviceContext(0) as c:
            c.density(xs, grid)
            res[name] = c.dkl_csr(p, j, v)
    assert _rel(res["tiles"], want) < RTOL_DKL and _rel(res["slab"], want) < RTOL_DKL
    assert not np.array_equal(res["tiles"], res["slab"])          # two different kernels really ran


def test_front_door_dispatch_variants(ctx, toy, tmp_path):
    """haystack(): data.frame coordinates, an AnnData-like object (the Seurat / SingleCellExperiment methods'
    role, R/s3.R:46-87), `dims`, spline.method = "bs", dir.randomization output, weights.advanced.Q."""
    import pandas as pd
    import scipy.sparse as ssp
    coord, expr = toy["coord"], toy["expression"]
    names = list(toy["gene_names"])
    common = dict(grid_coord=toy["grid60"], ctx=ctx, verbose=False, rng=7, gene_names=names)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        base = hh.haystack(coord, expr, **common)
        # data.frame in, same numbers (to rounding: DataFrame.values is column-major, so scale() sums in
        # another order) -- and the same call twice is bit-identical
        df = pd.DataFrame(coord, columns=["tSNE1", "tSNE2"])
        r_df = hh.haystack(df, expr, **common)
        np.testing.assert_allclose(r_df["results"]["D_KL"].values, base["results"]["D_KL"].values, rtol=1e-12)
        again = hh.haystack(coord, expr, **common)
        np.testing.assert_array_equal(again["results"]["D_KL"].values, base["results"]["D_KL"].values)
        np.testing.assert_array_equal(again["results"]["log.p.vals"].values, base["results"]["log.p.vals"].values)

        class FakeAnnData:           # cells x genes, embedding in .obsm
            X = ssp.csr_matrix(expr.T)
            obsm = {"X_tsne": np.column_stack([coord, np.zeros(601)])}
            layers = {}
            var_names = pd.Index(names)
        r_ad = hh.haystack(FakeAnnData(), coord="X_tsne", dims=[0, 1], grid_coord=toy["grid60"], ctx=ctx,
                           verbose=False, rng=7)
        assert list(r_ad["results"].index) == names
        assert _rel(r_ad["results"]["D_KL"].values, base["results"]["D_KL"].values) < RTOL_DKL
        with pytest.raises(ValueError, match="Please specify an embedding"):
            hh.haystack(FakeAnnData())
        # B-splines and the randomization dump
        out = tmp_path / "rand"
        r_bs = hh.haystack(coord, expr, spline_method="bs", dir_randomization=str(out), **common)
        assert r_bs["info"]["randomization"]["method"] == "bs"
        assert np.isfinite(r_bs["results"]["log.p.vals"].values).all()
        dumped = np.loadtxt(out / "random.D_KL.csv", delimiter=",")
        np.testing.assert_allclose(dumped, r_bs["info"]["randomization"]["KLD_rand"], rtol=1e-12)
        # weights.advanced.Q changes Q (and therefore D_KL), never P (R/haystack_continuous.R:182)
        w = 0.5 + (np.arange(601) % 5) / 2.0
        r_w = hh.haystack(coord, expr, weights_advanced_Q=w, **common)
    xs, _, _ = ho.scale_coords(coord)
    dsw = ho.density_stage(xs, toy["grid60"], w)
    want = ho.dkl_observed(expr, dsw["density"], dsw["Q"])
    assert _rel(r_w["results"]["D_KL"].values, want) < RTOL_DKL
    # strongly structured genes keep extreme p-values under either spline family
    top_ns = set(hh.show_result_haystack(base, n=25).index)
    top_bs = set(hh.show_result_haystack(r_bs, n=25).index)
    assert len(top_ns & top_bs) >= 15


def test_zero_variance_genes_are_filtered(ctx, toy):
    """R/haystack_continuous.R:76-80: rows with sd == 0 are dropped before scoring."""
    expr = toy["expression"].copy()
    expr[3, :] = 0.0
    expr[10, :] = 2.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = hh.haystack(toy["coord"], expr, grid_coord=toy["grid60"], ctx=ctx, verbose=False, rng=1,
                          gene_names=list(toy["gene_names"]))
    assert res["results"].shape == (498, 3)
    assert "gene_4" not in res["results"].index and "gene_11" not in res["results"].index


def test_toy_full_path_with_device_draws(ctx, toy):
    """perm_source="device": the randomization draws are made on the GPU from a seed.  The whole result must
    equal the oracle's when the oracle is fed the very ids that seed yields (oracle/prp_oracle.py), and the
    p-values must agree with the host-drawn run within what 100 randomizations can resolve."""
    from oracle import prp_oracle as po
    seed = int(toy["seed"])
    coord, expr = toy["coord"], toy["expression"]
    stats = ho.row_mean_sd(ho.RowSparse.from_dense(expr))
    folds = (synth.cv_folds(seed, 0, 100), synth.cv_folds(seed, 1, 100))
    kw = dict(grid_coord=toy["grid100"], gene_names=list(toy["gene_names"]), row_stats=stats, cv_folds=folds,
              ctx=ctx, verbose=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = hh.haystack(coord, sp.as_CsparseMatrix(expr), perm_source="device", rng=123, **kw)
        res2 = hh.haystack(coord, sp.as_CsparseMatrix(expr), perm_source="device", rng=123, **kw)
    rnd = res["info"]["randomization"]
    perm_seed = rnd["perm_seed"]
    np.testing.assert_array_equal(rnd["KLD_rand"], res2["info"]["randomization"]["KLD_rand"])   # reproducible
    np.testing.assert_array_equal(rnd["genes_to_randomize"], toy["full_genes_to_randomize"] + 1)

    def oracle_src(kind, gene, n, size, count):
        i = int(np.nonzero(toy["full_genes_to_randomize"] == gene)[0][0])
        return np.stack([po.prp_sample(perm_seed, i * count + r, n, size) for r in range(count)])

    want = ho.haystack_continuous_highD(coord, ho.RowSparse.from_dense(expr), toy["grid100"], oracle_src, folds[0], folds[1])
    assert _rel(rnd["KLD_rand"], want["KLD_rand"]) < RTOL_DKL
    assert np.max(np.abs(res["results"]["log.p.vals"].values - want["log_p_vals"])) < ATOL_LOGP
    # same null distribution as the host-drawn golden run: per-gene means within a few standard errors
    gold = toy["full_KLD_rand"]
    z = (rnd["KLD_rand"].mean(axis=1) - gold.mean(axis=1)) / (gold.std(axis=1, ddof=1) * np.sqrt(2 / 100))
    assert np.abs(z).max() < 5 and abs(z.mean()) < 0.5


def test_convert_on_device_gives_identical_results(ctx, toy):
    """haystack(convert_on_device=True): the dgCMatrix -> dgRMatrix step on the GPU (hs_csc_to_csr) yields the
    same slots as the host conversion, hence identical results."""
    seed = int(toy["seed"])
    expr = toy["expression"]
    kw = dict(grid_coord=toy["grid100"], perm_source=synth.perm_source(seed),
              row_stats=ho.row_mean_sd(ho.RowSparse.from_dense(expr)),
              cv_folds=(synth.cv_folds(seed, 0, 100), synth.cv_folds(seed, 1, 100)), ctx=ctx, verbose=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = hh.haystack(toy["coord"], sp.as_CsparseMatrix(expr), **kw)
        b = hh.haystack(toy["coord"], sp.as_CsparseMatrix(expr), convert_on_device=True, **kw)
    np.testing.assert_array_equal(a["results"]["D_KL"].values, b["results"]["D_KL"].values)
    # (the ids-given randomization sums in atomics order: its scores repeat to ~1e-8, not bit for bit)
    np.testing.assert_allclose(a["results"]["log.p.vals"].values, b["results"]["log.p.vals"].values, rtol=0, atol=1e-6)

"""Shape sweep of the device path against the oracle: grid sizes that select every kernel variant
(1, 2, 4, 8 float4 per lane; tensor-core and SIMT dense kernels), tiny and ragged inputs."""
import os

import numpy as np
import pytest

from oracle import haystack_oracle as ho
from singlecellhaystack_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300), initial=0.0)


CASES = [
    # n_cells, dims, grid, genes, density
    (40, 2, 3, 5, 0.5),          # fewer cells than a slab, grid of 3
    (257, 1, 1, 9, 0.3),         # a single grid point: every D_KL is 0
    (1000, 3, 31, 64, 0.05),
    (2048, 7, 128, 33, 0.2),     # exactly one tensor-core column tile
    (3000, 4, 129, 40, 0.1),     # first size past it (SIMT dense kernel, 2 float4 per lane)
    (1500, 5, 300, 24, 0.1),     # 4 float4 per lane
    (1200, 3, 600, 12, 0.15),    # 8 float4 per lane
    (900, 2, 1024, 6, 0.2),      # the supported maximum
    (70000, 6, 100, 20, 0.002),  # very sparse rows, many slabs
]


@pytest.mark.parametrize("n,d,g,m,rho", CASES)
def test_shape_sweep(ctx, n, d, g, m, rho):
    rng = np.random.default_rng(n + g)
    x = rng.normal(size=(n, d)) * (1 + np.arange(d))[None, :]
    xs, _, _ = ho.scale_coords(x) if n > 1 else (x, None, None)
    grid = xs[rng.choice(n, size=g, replace=g > n)] + 0.01 * rng.normal(size=(g, d))
    dense = np.where(rng.random((m, n)) < rho, rng.gamma(2.0, size=(m, n)), 0.0)
    dense[0, :] = 0.0                       # an empty row
    dense[-1, :] = 1.5                      # a constant row (P == Q)
    if m > 3:
        dense[1, :] = 0.0
        dense[1, n // 2] = 3.0              # a single stored entry
    sp = ho.RowSparse.from_dense(dense)
    ref = ho.density_stage(xs, grid)
    want = ho.dkl_observed(sp, ref["density"], ref["Q"])
    out = ctx.density(xs, grid, want_nearest=True, want_densities=True)
    assert abs(out["bandwidth"] - ref["bandwidth"]) <= 1e-12 * max(ref["bandwidth"], 1e-300)
    np.testing.assert_array_equal(out["nearest"], ref["nearest"])
    np.testing.assert_allclose(out["Q"], ref["Q"], rtol=1e-9)
    np.testing.assert_allclose(out["densities"], ref["density"], rtol=1e-9, atol=1e-300)
    got_csr = ctx.dkl_csr(sp.p, sp.j, sp.x)
    got_dense = ctx.dkl_dense(dense)
    for got in (got_csr, got_dense):
        assert got.shape == (m,)
        big = np.abs(want) > 1e-9
        assert _rel(got[big], want[big]) < RTOL
        assert np.max(np.abs(got[~big] - want[~big]), initial=0.0) < 1e-9
    # randomized rows for two genes, both forms
    src = synth.perm_source(11)
    rows = [2 % m, m // 2]
    vals = [sp.x[sp.p[r]:sp.p[r + 1]] for r in rows]
    ids = [src("sparse", r, n, v.size, 5) for r, v in zip(rows, vals)]
    want_r = ho.dkl_randomized_sparse(vals, ids, ref["density"], ref["Q"])
    got_r = ctx.dkl_perm_csr(vals, [np.asfortranarray(a.T) for a in ids], index_base=1)
    ok = np.abs(want_r) > 1e-9
    assert _rel(got_r[ok], want_r[ok]) < RTOL and np.max(np.abs(got_r[~ok] - want_r[~ok]), initial=0) < 1e-9
    perms = src("dense", rows[1], n, n, 4)
    want_d = ho.dkl_randomized_dense(dense[[rows[1]]], perms[None], ref["density"], ref["Q"])[0]
    got_d = ctx.dkl_perm_dense(dense[rows[1]], np.asfortranarray(perms.T), index_base=1)
    okd = np.abs(want_d) > 1e-9
    assert _rel(got_d[okd], want_d[okd]) < RTOL and np.max(np.abs(got_d[~okd] - want_d[~okd]), initial=0) < 1e-9


def test_row_stats_csr_match_oracle_and_selection(ctx, toy, monkeypatch):
    """hs_row_stats_csr (R/haystack_continuous.R:57-74): fp64 two-pass mean / sd per CSR row on the device;
    the integer gene selection computed from them must be the oracle's."""
    from oracle import haystack_oracle as ho
    from singlecellhaystack_b200 import haystack as hh
    sp = ho.RowSparse.from_dense(toy["expression"].astype(np.float64))
    want_mean, want_sd = ho.row_mean_sd(sp)
    mean, sd = ctx.row_stats_csr(sp.p, sp.x, sp.shape[1])
    np.testing.assert_allclose(mean, want_mean, rtol=1e-14, atol=0)
    np.testing.assert_allclose(sd, want_sd, rtol=1e-13, atol=0)
    keep = want_sd > 0
    cv_dev, cv_ref = (sd[keep] + 1e-300) / (mean[keep] + 1e-300), (want_sd[keep] + 1e-300) / (want_mean[keep] + 1e-300)
    # the selection picks the same coefficient-of-variation ranks (genes with tied CVs -- the toy data has
    # rows that are permutations of each other -- may swap, in R as well: their order hangs on the last bit)
    sel_dev = hh.select_genes_to_randomize(cv_dev, 100, "heavytails")
    sel_ref = ho.select_genes_to_randomize(cv_ref, 100, "heavytails")
    np.testing.assert_allclose(cv_ref[sel_dev], cv_ref[sel_ref], rtol=1e-12, atol=0)
    # int32 row pointers, several upload pieces, empty rows and a constant row (sd == 0 exactly)
    monkeypatch.setenv("HS_STAGE_VALUES", "4096")
    m2, s2 = ctx.row_stats_csr(sp.p.astype(np.int32), sp.x, sp.shape[1])
    np.testing.assert_array_equal(m2, mean)
    np.testing.assert_array_equal(s2, sd)
    p = np.array([0, 0, 5, 5, 6], dtype=np.int64)
    x = np.array([2.0, 2.0, 2.0, 2.0, 2.0, 7.5])
    m3, s3 = ctx.row_stats_csr(p, x, 5)
    np.testing.assert_array_equal(m3, [0.0, 2.0, 0.0, 1.5])
    assert s3[0] == 0 and s3[1] == 0 and s3[2] == 0
    assert abs(s3[3] - np.std([7.5, 0, 0, 0, 0], ddof=1)) < 1e-15


def test_front_door_with_device_row_stats(ctx, toy):
    import warnings
    from oracle import haystack_oracle as ho
    from singlecellhaystack_b200 import haystack as hh, sparse as sp_, synth
    seed = int(toy["seed"])
    folds = (synth.cv_folds(seed, 0, 100), synth.cv_folds(seed, 1, 100))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = hh.haystack(toy["coord"], sp_.as_CsparseMatrix(toy["expression"].astype(np.float64)), grid_coord=toy["grid100"],
                          perm_source=synth.perm_source(seed), row_stats="device", cv_folds=folds, ctx=ctx, verbose=False)
    assert np.max(np.abs(res["results"]["D_KL"].values - toy["dkl100"]) / toy["dkl100"]) < RTOL
    same = np.array_equal(res["info"]["randomization"]["genes_to_randomize"], toy["full_genes_to_randomize"] + 1)
    # identical selection -> identical null sample -> identical p-values; tied genes swapped -> another null sample
    assert np.max(np.abs(res["results"]["log.p.vals"].values - toy["full_log_p"])) < (1e-3 if same else 1.5)


def test_csc_to_csr_matches_host_conversion(ctx, toy):
    """hs_csc_to_csr (as(expression, "RsparseMatrix"), R/haystack_continuous.R:113-116): the dgRMatrix slots
    must equal the host conversion bit for bit -- row pointers, ascending column ids, values."""
    from singlecellhaystack_b200 import sparse as sp_
    cases = [sp_.as_CsparseMatrix(toy["expression"].astype(np.float64))]
    rng = np.random.default_rng(12)
    a = rng.random((300, 1500)) * (rng.random((300, 1500)) < 0.07)
    a[17] = 0.0                      # empty row
    a[:, 40:90] = 0.0                # a run of empty columns
    a[250] = 1.5                     # a full row
    cases.append(sp_.as_CsparseMatrix(a))
    cases.append(sp_.as_CsparseMatrix(np.zeros((5, 7))))          # nothing stored
    tall = rng.random((130_000, 24)) * (rng.random((130_000, 24)) < 0.06)     # more rows than one band of counters holds
    cases.append(sp_.as_CsparseMatrix(tall))
    for m in cases:
        want = sp_.as_RsparseMatrix(m)
        got = sp_.as_RsparseMatrix(m, ctx=ctx)
        np.testing.assert_array_equal(got.p, want.p)
        np.testing.assert_array_equal(got.j, want.j)
        np.testing.assert_array_equal(got.x, want.x)
        # large inputs are placed by a stable radix sort, chunk of column blocks by chunk: forced here, small chunks
        for chunk in ("1000000000", "5000", "1024"):
            os.environ["HS_CONVERT_SORT"], os.environ["HS_CONVERT_CHUNK"] = "1", chunk
            try:
                srt = ctx.csc_to_csr(m.p, m.i, m.x, m.Dim[0], m.Dim[1])
            finally:
                del os.environ["HS_CONVERT_SORT"], os.environ["HS_CONVERT_CHUNK"]
            np.testing.assert_array_equal(srt[0], want.p)
            np.testing.assert_array_equal(srt[1], want.j)
            np.testing.assert_array_equal(srt[2], want.x)
        again = ctx.csc_to_csr(m.p.astype(np.int32), m.i, m.x, m.Dim[0], m.Dim[1])     # int32 pointers, repeat
        np.testing.assert_array_equal(again[1], want.j)
    from singlecellhaystack_b200.device import HaystackLibraryError
    m = cases[1]
    bad_i = m.i.copy()
    bad_i[3] = 300
    with pytest.raises(HaystackLibraryError, match="row ids outside"):
        ctx.csc_to_csr(m.p, bad_i, m.x, 300, 1500)


def test_csc_to_csr_device_form_feeds_the_scoring(ctx, toy):
    import torch
    from oracle import haystack_oracle as ho
    from singlecellhaystack_b200 import sparse as sp_
    xs, _, _ = ho.scale_coords(toy["coord"])
    ctx.density(xs, toy["grid100"])
    m = sp_.as_CsparseMatrix(toy["expression"].astype(np.float64))
    dev = torch.device("cuda:0")
    cp = torch.as_tensor(m.p.astype(np.int64), device=dev)
    ri = torch.as_tensor(m.i, device=dev)
    x = torch.as_tensor(m.x.astype(np.float32), device=dev)
    p_out = torch.empty(501, dtype=torch.int64, device=dev)
    j_out = torch.empty(ri.numel(), dtype=torch.int32, device=dev)
    x_out = torch.empty(ri.numel(), dtype=torch.float32, device=dev)
    ctx.use_torch_stream(torch.cuda.current_stream())
    try:
        ctx.csc_to_csr_dev(cp, ri, x, 500, 601, p_out, j_out, x_out)
        dkl = torch.empty(500, dtype=torch.float64, device=dev)
        ctx.dkl_csr_dev(p_out, j_out, x_out, dkl)
        torch.cuda.synchronize()
    finally:
        ctx.set_stream(None)
    assert _rel(dkl.cpu().numpy(), toy["dkl100"]) < RTOL


def test_gene_selection_on_the_device(ctx):
    """hs_select_by_cv_dev against the host restatement of R/haystack_continuous.R:222-236: ties (order() is stable),
    NaN last, negative and zero coefficients, both selection methods, sizes of BASELINE.json's configs."""
    from singlecellhaystack_b200 import haystack as hk
    rng = np.random.default_rng(5)
    for m in (500, 12382, 16811, 30000):
        mean = rng.gamma(2.0, 1.0, size=m)
        sd = rng.gamma(2.0, 1.0, size=m)
        sd[rng.choice(m, m // 10, replace=False)] = sd[0]            # ties in sd ...
        mean[rng.choice(m, m // 10, replace=False)] = mean[1]        # ... and in mean: tied coefficients
        dup = rng.choice(m, m // 5, replace=False)
        mean[dup], sd[dup] = mean[dup[0]], sd[dup[0]]
        cv = (sd + 1e-300) / (mean + 1e-300)
        for method in ("heavytails", "uniform"):
            want = hk.select_genes_to_randomize(cv, 100, method)
            ranks = hk.selection_ranks(m, 100, method) - 1
            got, cv_dev = ctx.select_by_cv(mean, sd, ranks, return_cv=True)
            np.testing.assert_array_equal(got, want)
            np.testing.assert_array_equal(cv_dev.cpu().numpy(), cv)
    # the whole order, with NaN, infinities, zeros of both signs and negative coefficients
    mean = np.array([1.0, 0.0, 2.0, -1.0, np.nan, 1.0, -0.0, 4.0, 1e300, 3.0])
    sd = np.array([1.0, 1.0, 0.0, 1.0, 1.0, 1.0, 0.0, np.nan, 1.0, 0.0])
    cv = (sd + 1e-300) / (mean + 1e-300)
    got = ctx.select_by_cv(mean, sd, np.arange(10))
    np.testing.assert_array_equal(got, np.argsort(cv, kind="stable"))
    from singlecellhaystack_b200.device import HaystackLibraryError
    with pytest.raises(HaystackLibraryError, match="outside"):
        ctx.select_by_cv(mean, sd, np.array([10]))


def test_seeding_grid_with_device_distances(ctx, toy):
    """get_grid_points(method = "seeding") (R/haystack_highD.R:87-129) with the distance passes on the device
    (hs_grid_seed_*): the same picks and the same grid as the host computation, on R's stream and on numpy's; and as
    the oracle's restatement of the R procedure."""
    from oracle import r_base
    from singlecellhaystack_b200 import grid as gr
    from singlecellhaystack_b200.rcompat import RRandom
    xs, _, _ = ho.scale_coords(toy["coord"])
    for g in (7, 60):
        host = gr.get_grid_points(xs, "seeding", g, rng=RRandom(1234))
        dev = gr.get_grid_points(xs, "seeding", g, rng=RRandom(1234), ctx=ctx)
        np.testing.assert_allclose(dev, host, rtol=0, atol=1e-12)
        np.testing.assert_allclose(dev, r_base.get_grid_points(xs, r_base.RRng(1234), "seeding", g), rtol=0, atol=1e-12)
        host_np = gr.get_grid_points(xs, "seeding", g, rng=5)
        dev_np = gr.get_grid_points(xs, "seeding", g, rng=5, ctx=ctx)
        np.testing.assert_allclose(dev_np, host_np, rtol=0, atol=1e-12)
    # a bigger cloud in 20 dimensions, duplicated cells (tied distances: the lower cell id wins, as order() does)
    rng = np.random.default_rng(3)
    x = rng.normal(size=(20000, 20))
    x[100:110] = x[50]
    np.testing.assert_allclose(gr.get_grid_points(x, "seeding", 25, rng=11, ctx=ctx),
                               gr.get_grid_points(x, "seeding", 25, rng=11), rtol=0, atol=1e-12)
    # fewer than three cells
    tiny = rng.normal(size=(2, 3))
    np.testing.assert_allclose(gr.get_grid_points(tiny, "seeding", 2, rng=1, ctx=ctx), gr.get_grid_points(tiny, "seeding", 2, rng=1))

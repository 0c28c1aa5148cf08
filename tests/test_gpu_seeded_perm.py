"""Seeded randomization batch (hs_dkl_perm_csr_seeded): the device draws sample(n_cells, nnz)
(R/haystack_continuous.R:275) from the keyed permutation; parity against the fp64 oracle fed with the same
ids (oracle/prp_oracle.py), and against the ids-given entry point."""
import numpy as np
import pytest

from oracle import haystack_oracle as ho
from oracle import prp_oracle as po
from singlecellhaystack_b200 import synth

pytestmark = pytest.mark.gpu

RTOL_DKL = 1e-5


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def _toy_rows(toy, k):
    expr = toy["expression"]
    sp = ho.RowSparse.from_dense(expr)
    genes = toy["full_genes_to_randomize"][:k]
    return [ho.extract_row_dgRMatrix_as_sparse(sp, int(g) + 1)[1] for g in genes], expr.shape[1]


@pytest.mark.parametrize("g", [60, 100])
def test_seeded_matches_oracle_on_toy(ctx, toy, g):
    xs, _, _ = ho.scale_coords(toy["coord"])
    ref = ho.density_stage(xs, toy[f"grid{g}"])
    ctx.density(xs, toy[f"grid{g}"])
    vals, n = _toy_rows(toy, 10)
    n_perm, seed, base = 40, 424242, 1000
    inds = [np.stack([po.prp_sample(seed, base + s * n_perm + r, n, v.size) for r in range(n_perm)])
            for s, v in enumerate(vals)]                                          # (R, nnz) 1-based
    want = ho.dkl_randomized_sparse(vals, inds, ref["density"], ref["Q"])
    got = ctx.dkl_perm_csr_seeded(vals, n_perm, seed, stream_base=base)
    assert got.shape == (10, n_perm)
    assert _rel(got, want) < RTOL_DKL
    # the ids-given entry point on the very same draw
    via_ids = ctx.dkl_perm_csr(vals, [np.asfortranarray(a.T) for a in inds], index_base=1)
    assert _rel(got, via_ids) < 1e-6
    # bit-identical when repeated, and when the genes are scored in two calls with shifted streams
    again = ctx.dkl_perm_csr_seeded(vals, n_perm, seed, stream_base=base)
    np.testing.assert_array_equal(got, again)
    lo = ctx.dkl_perm_csr_seeded(vals[:4], n_perm, seed, stream_base=base)
    hi = ctx.dkl_perm_csr_seeded(vals[4:], n_perm, seed, stream_base=base + 4 * n_perm)
    np.testing.assert_array_equal(got, np.vstack([lo, hi]))


def test_seeded_edge_rows(ctx, toy):
    """Empty gene, single stored value, and a gene stored in every cell (the draw is a full permutation)."""
    xs, _, _ = ho.scale_coords(toy["coord"])
    ref = ho.density_stage(xs, toy["grid60"])
    ctx.density(xs, toy["grid60"])
    n = 601
    rng = np.random.default_rng(3)
    vals = [np.zeros(0), np.array([2.5]), rng.random(n) + 0.1, rng.random(17)]
    n_perm, seed = 7, 5
    got = ctx.dkl_perm_csr_seeded(vals, n_perm, seed)
    inds = [np.stack([po.prp_sample(seed, s * n_perm + r, n, v.size) for r in range(n_perm)]) if v.size
            else np.zeros((n_perm, 0), dtype=np.int32) for s, v in enumerate(vals)]
    want = ho.dkl_randomized_sparse(vals[1:], inds[1:], ref["density"], ref["Q"])
    assert _rel(got[1:], want) < RTOL_DKL
    # a gene without stored values: P is the pseudocount only, i.e. uniform over the grid points
    q = ref["Q"]
    assert np.allclose(got[0], np.sum(np.full(60, 1 / 60) * np.log(np.full(60, 1 / 60) / q)), rtol=1e-9)
    # every cell drawn: any permutation of a full row of constant values scores the same
    full = ctx.dkl_perm_csr_seeded([np.ones(n)], 5, 11)
    assert np.ptp(full) < 1e-12 * np.abs(full).max()


def test_seeded_oversized_sample_is_rejected(ctx, toy):
    from singlecellhaystack_b200.device import HaystackLibraryError
    xs, _, _ = ho.scale_coords(toy["coord"])
    ctx.density(xs, toy["grid60"])
    with pytest.raises(HaystackLibraryError, match="larger than the population"):
        ctx.dkl_perm_csr_seeded([np.ones(602)], 3, 1)


def test_seeded_batches_and_large_population(ctx):
    """A population that is not a power of two (cycle walking), genes split into memory batches."""
    import os
    n, d, g = 70001, 6, 48
    x, _ = synth.make_embedding(n, d, seed=77)
    xs, _, _ = ho.scale_coords(x)
    grid = synth.make_grid(xs, g, seed=77)
    ref = ho.density_stage(xs, grid)
    ctx.density(xs, grid)
    rng = np.random.default_rng(8)
    vals = [rng.gamma(2.0, 1.0, size=int(k)) for k in (5, 900, 23000, 70001, 3100, 1)]
    n_perm, seed = 6, 2026
    got = ctx.dkl_perm_csr_seeded(vals, n_perm, seed)
    inds = [np.stack([po.prp_sample(seed, s * n_perm + r, n, v.size) for r in range(n_perm)]) for s, v in enumerate(vals)]
    want = ho.dkl_randomized_sparse(vals, inds, ref["density"], ref["Q"])
    assert _rel(got, want) < RTOL_DKL
    os.environ["HS_UPLOAD_BUDGET"] = str(2_000_000)      # forces several batches
    try:
        split = ctx.dkl_perm_csr_seeded(vals, n_perm, seed)
    finally:
        del os.environ["HS_UPLOAD_BUDGET"]
    np.testing.assert_array_equal(got, split)


def test_short_and_long_row_kernels_agree(ctx, toy):
    """Rows with few stored values are drawn forward and sorted; the others walk the cells through the
    inverse permutation.  Both emit the same pairs in the same order, so the scores are bit-identical."""
    import os
    xs, _, _ = ho.scale_coords(toy["coord"])
    ctx.density(xs, toy["grid100"])
    vals, _ = _toy_rows(toy, 12)
    os.environ["HS_PERM_FUSED"] = "0"                # the pair generators, not the draw into the tiles
    try:
        a = ctx.dkl_perm_csr_seeded(vals, 9, 77)
        os.environ["HS_PERM_SMALL_ROW"] = "0"        # everything through the cell walk
        try:
            b = ctx.dkl_perm_csr_seeded(vals, 9, 77)
        finally:
            del os.environ["HS_PERM_SMALL_ROW"]
    finally:
        del os.environ["HS_PERM_FUSED"]
    np.testing.assert_array_equal(a, b)


def _with_env(name, value, fn):
    import os
    os.environ[name] = value
    try:
        return fn()
    finally:
        del os.environ[name]


def test_draw_into_tiles_matches_pair_generators(ctx):
    """The default seeded path draws pi(i) forward, straight into the tensor-core kernel's tiles; the earlier one walks the
    cells through the inverse permutation, emits sorted pairs and buckets them.  Same rows in the same tiles: the scores
    are bit-identical, whatever the number of cell ranges the draw is cut into."""
    n, d, g = 150_000, 5, 100
    x, _ = synth.make_embedding(n, d, seed=5)
    xs, _, _ = ho.scale_coords(x)
    grid = synth.make_grid(xs, g, seed=5)
    ctx.density(xs, grid)
    rng = np.random.default_rng(12)
    vals = [rng.gamma(2.0, 1.0, size=int(k)) for k in (40_000, 1, 0, 15_000, 150_000, 777, 9_000, 31)]
    n_perm, seed = 40, 99                                  # 320 rows: three tile rows
    fused = ctx.dkl_perm_csr_seeded(vals, n_perm, seed)
    pairs = _with_env("HS_PERM_FUSED", "0", lambda: ctx.dkl_perm_csr_seeded(vals, n_perm, seed))
    np.testing.assert_array_equal(fused, pairs)
    for parts in ("1", "3", "7"):
        cut = _with_env("HS_PERM_TILE_PARTS", parts, lambda: ctx.dkl_perm_csr_seeded(vals, n_perm, seed))
        np.testing.assert_array_equal(fused, cut)
    # and against the fp64 oracle on a few rows
    ref = ho.density_stage(xs, grid)
    pick = [0, 3, 5]
    inds = [np.stack([po.prp_sample(seed, s * n_perm + r, n, vals[s].size) for r in range(3)]) for s in pick]
    want = ho.dkl_randomized_sparse([vals[s] for s in pick], inds, ref["density"], ref["Q"])
    assert _rel(fused[pick][:, :3], want) < RTOL_DKL


def test_draw_into_tiles_fallbacks(ctx, toy):
    """A tile window that overflows, or a value beyond fp16, hands the call to the chained gather kernel on the device."""
    xs, _, _ = ho.scale_coords(toy["coord"])
    ref = ho.density_stage(xs, toy["grid100"])
    ctx.density(xs, toy["grid100"])
    vals, n = _toy_rows(toy, 6)
    n_perm, seed = 12, 31
    inds = [np.stack([po.prp_sample(seed, s * n_perm + r, n, v.size) for r in range(n_perm)]) for s, v in enumerate(vals)]
    want = ho.dkl_randomized_sparse(vals, inds, ref["density"], ref["Q"])
    normal = ctx.dkl_perm_csr_seeded(vals, n_perm, seed)
    assert _rel(normal, want) < RTOL_DKL
    tiny = _with_env("HS_PERM_TILE_SIGMAS", "-1000", lambda: ctx.dkl_perm_csr_seeded(vals, n_perm, seed))   # 2-entry windows
    assert _rel(tiny, want) < RTOL_DKL
    assert not np.array_equal(tiny, normal)                # it really took the other kernel
    big = [v.copy() for v in vals]
    big[2][0] = 1.0e6
    want_big = ho.dkl_randomized_sparse(big, inds, ref["density"], ref["Q"])
    assert _rel(ctx.dkl_perm_csr_seeded(big, n_perm, seed), want_big) < RTOL_DKL


@pytest.mark.parametrize("g", [60, 100])
def test_seeded_dense_matches_oracle_on_toy(ctx, toy, g):
    """Dense form: sample(x = row) == row[sample.int(N)] (R/haystack_continuous.R:259) drawn on the device."""
    xs, _, _ = ho.scale_coords(toy["coord"])
    ref = ho.density_stage(xs, toy[f"grid{g}"])
    ctx.density(xs, toy[f"grid{g}"])
    expr = toy["expression"]
    n = expr.shape[1]
    n_perm, seed = 30, 99
    for k, gidx in enumerate(toy["full_genes_to_randomize"][:4]):
        base = k * n_perm
        perms = np.stack([po.prp_sample(seed, base + r, n, n) for r in range(n_perm)])          # (R, N) 1-based
        want = ho.dkl_randomized_dense(expr[[gidx]], perms[None], ref["density"], ref["Q"])[0]
        got = ctx.dkl_perm_dense_seeded(expr[gidx], n_perm, seed, stream_base=base)
        assert _rel(got, want) < RTOL_DKL
        via_ids = ctx.dkl_perm_dense(expr[gidx], np.asfortranarray(perms.T), index_base=1)
        np.testing.assert_array_equal(got, via_ids)          # same ids, same kernels
        np.testing.assert_array_equal(got, ctx.dkl_perm_dense_seeded(expr[gidx], n_perm, seed, stream_base=base))


def test_dense_front_door_with_device_draws(ctx, toy):
    import warnings
    from singlecellhaystack_b200 import haystack as hh
    seed = int(toy["seed"])
    coord, expr = toy["coord"], toy["expression"]
    stats = ho.row_mean_sd(expr)
    folds = (synth.cv_folds(seed, 0, 20), synth.cv_folds(seed, 1, 20))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = hh.haystack(coord, expr, grid_coord=toy["grid100"], perm_source="device", rng=5, row_stats=stats,
                          cv_folds=folds, ctx=ctx, verbose=False, n_genes_to_randomize=20, randomization_count=25)
    rnd = res["info"]["randomization"]
    ps = rnd["perm_seed"]
    ref = ho.density_stage(ho.scale_coords(coord)[0], toy["grid100"])
    g0 = int(rnd["genes_to_randomize"][3]) - 1
    perms = np.stack([po.prp_sample(ps, 3 * 25 + r, 601, 601) for r in range(25)])
    want = ho.dkl_randomized_dense(expr[[g0]], perms[None], ref["density"], ref["Q"])[0]
    assert _rel(rnd["KLD_rand"][3], want) < RTOL_DKL


def test_combined_call_equals_separate_calls(ctx, toy):
    """hs_dkl_csr_randomized launches the randomization batch before the matrix upload; scores and randomized
    scores must be bit-identical to hs_dkl_csr followed by hs_dkl_perm_csr_seeded."""
    xs, _, _ = ho.scale_coords(toy["coord"])
    ctx.density(xs, toy["grid100"])
    sp = ho.RowSparse.from_dense(toy["expression"])
    sel = toy["full_genes_to_randomize"][:15].astype(np.int64)
    dkl, rnd = ctx.dkl_csr_randomized(sp.p, sp.j, sp.x, sel, 12, 31337, stream_base=7)
    np.testing.assert_array_equal(dkl, ctx.dkl_csr(sp.p, sp.j, sp.x))
    vals = [sp.x[sp.p[g]:sp.p[g + 1]] for g in sel]
    np.testing.assert_array_equal(rnd, ctx.dkl_perm_csr_seeded(vals, 12, 31337, stream_base=7))
    assert _rel(dkl, toy["dkl100"]) < RTOL_DKL
    # no selected rows: plain scoring
    dkl0, rnd0 = ctx.dkl_csr_randomized(sp.p, sp.j, sp.x, np.zeros(0, dtype=np.int64), 5, 1)
    np.testing.assert_array_equal(dkl0, dkl)
    assert rnd0.shape == (0, 5)
    from singlecellhaystack_b200.device import HaystackLibraryError
    with pytest.raises(HaystackLibraryError, match="out of range"):
        ctx.dkl_csr_randomized(sp.p, sp.j, sp.x, np.array([500], dtype=np.int64), 5, 1)


def test_dense_seeded_batch_equals_single_calls(ctx, toy):
    """hs_dkl_perm_dense_seeded_batch = the loop R/haystack_continuous.R:250-265 in one call: same streams, same
    kernels, so bit-identical to n_sel calls of hs_dkl_perm_dense_seeded; and against the oracle fed the same ids."""
    xs, _, _ = ho.scale_coords(toy["coord"])
    ref = ho.density_stage(xs, toy["grid100"])
    ctx.density(xs, toy["grid100"])
    expr = np.asfortranarray(toy["expression"].astype(np.float64))
    rows = np.array([3, 499, 0, 77, 250])
    n_perm, seed, base = 9, 1234567, 40
    got = ctx.dkl_perm_dense_seeded_batch(expr, rows, n_perm, seed, stream_base=base)
    assert got.shape == (rows.size, n_perm)
    one = np.stack([ctx.dkl_perm_dense_seeded(expr[r], n_perm, seed, stream_base=base + s * n_perm) for s, r in enumerate(rows)])
    np.testing.assert_array_equal(got, one)
    n = expr.shape[1]
    for s in (0, 4):
        perms = np.stack([po.prp_sample(seed, base + s * n_perm + r, n, n) for r in range(n_perm)])     # 1-based
        want = ho.dkl_randomized_dense(expr[[rows[s]]], perms[None], ref["density"], ref["Q"])[0]
        assert _rel(got[s], want) < RTOL_DKL

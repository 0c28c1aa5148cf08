"""The tensor-core tile path of the sparse contraction (csrc/hs_tiles.cu) against the fp64 oracle
(R/haystack_continuous.R:586-602): tile-row / segment / chunk boundaries, the fallbacks chained on its flag
(ids not ascending, ids outside [0, N), values beyond fp16), both MMA widths, the randomization batch through
the same kernels, and agreement with the SIMT slab kernel it replaces."""
import os

import numpy as np
import pytest

from oracle import haystack_oracle as ho
from singlecellhaystack_b200 import synth
from singlecellhaystack_b200.device import DeviceContext

pytestmark = pytest.mark.gpu
RTOL = 1e-5          # north_star: D_KL within 1e-5 relative (fp32 accumulate)


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300), initial=0.0)


def _check(got, want):
    big = np.abs(want) > 1e-9
    assert _rel(got[big], want[big]) < RTOL, _rel(got[big], want[big])
    assert np.max(np.abs(got[~big] - want[~big]), initial=0.0) < 1e-9


def _problem(n, d, g, m, rho_lo, rho_hi, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, d)) * (1 + np.arange(d))[None, :]
    xs, _, _ = ho.scale_coords(x)
    grid = xs[rng.choice(n, size=g, replace=False)] + 0.01 * rng.normal(size=(g, d))
    rho = np.exp(rng.uniform(np.log(rho_lo), np.log(rho_hi), size=m))
    p = [0]
    js, vs = [], []
    for r in range(m):
        k = int(round(rho[r] * n))
        j = np.sort(rng.choice(n, size=k, replace=False)).astype(np.int32)
        js.append(j)
        vs.append(np.log1p(1 + rng.poisson(2.0, size=k)).astype(np.float64))
        p.append(p[-1] + k)
    sp = ho.RowSparse(p=np.array(p, dtype=np.int64), j=np.concatenate(js) if js else np.zeros(0, np.int32),
                      x=np.concatenate(vs) if vs else np.zeros(0), shape=(m, n))
    return xs, grid, sp


# cells, dims, grid, genes: tile rows 1..5 (odd counts leave the second tile of the last pair empty), cells on
# both sides of the 64-cell slab, 2048-cell flush window and 4096-cell chunk boundaries, every MMA width
CASES = [
    (63, 2, 5, 3),            # less than one slab, one partial tile
    (64, 2, 16, 130),         # exactly one slab; two tile rows, the second nearly empty
    (2049, 3, 100, 257),      # one cell past a flush window; three tile rows (odd)
    (4097, 2, 37, 300),       # one cell past a chunk
    (9000, 4, 128, 513),      # widest MMA (N = 128), five tile rows
    (20000, 5, 100, 256),     # several segments
    (12345, 1, 60, 129),      # 1-D coordinates, N = 64
]


@pytest.mark.parametrize("n,d,g,m", CASES)
def test_tile_path_matches_oracle(ctx, n, d, g, m):
    xs, grid, sp = _problem(n, d, g, m, 0.002, 0.6, seed=n + m)
    ref = ho.density_stage(xs, grid)
    want = ho.dkl_observed(sp, ref["density"], ref["Q"])
    ctx.density(xs, grid)
    launches0 = ctx.launch_count
    got = ctx.dkl_csr(sp.p, sp.j, sp.x)
    assert ctx.launch_count > launches0
    _check(got, want)
    # bit-identical repeats: every sum has a single writer and a fixed order
    again = ctx.dkl_csr(sp.p, sp.j, sp.x)
    np.testing.assert_array_equal(got, again)


def test_tile_path_agrees_with_slab_kernel_and_both_widths(monkeypatch):
    """Same rows through the tile path (N rounded to 16 and forced to 128) and the SIMT slab kernel."""
    xs, grid, sp = _problem(30000, 6, 100, 400, 0.001, 0.5, seed=7)
    ref = ho.density_stage(xs, grid)
    want = ho.dkl_observed(sp, ref["density"], ref["Q"])
    res = {}
    for name, env in (("tiles", {}), ("tiles128", {"HS_TILE_N128": "1"}), ("slab", {"HS_TILE_DISABLE": "1"})):
        for k in ("HS_TILE_N128", "HS_TILE_DISABLE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with DeviceContext(0) as c:
            c.density(xs, grid)
            res[name] = c.dkl_csr(sp.p, sp.j, sp.x)
        _check(res[name], want)
    assert _rel(res["tiles"], res["slab"]) < 2e-5


def test_tile_path_fallbacks(ctx):
    """Rows the tile kernels cannot take are handed on: ids not ascending and a value beyond fp16's range are
    still scored (slab / gather kernels); an id >= N poisons exactly its own row with NaN, like the reference's
    D[ind, ] would fail (R/haystack_continuous.R:596)."""
    xs, grid, sp = _problem(9000, 3, 50, 140, 0.01, 0.3, seed=3)
    ref = ho.density_stage(xs, grid)
    want = ho.dkl_observed(sp, ref["density"], ref["Q"])
    ctx.density(xs, grid)
    # (a) one row with two ids swapped across a chunk boundary
    j = sp.j.copy()
    x = sp.x.copy()
    a, b = int(sp.p[5]), int(sp.p[6])
    j[a], j[b - 1] = j[b - 1], j[a]
    x[a], x[b - 1] = x[b - 1], x[a]
    _check(ctx.dkl_csr(sp.p, j, x), want)
    # (b) a value fp16 cannot hold
    x2 = sp.x.copy()
    x2[int(sp.p[7])] = 1.0e6
    sp2 = ho.RowSparse(p=sp.p, j=sp.j, x=x2, shape=sp.shape)
    _check(ctx.dkl_csr(sp.p, sp.j, x2), ho.dkl_observed(sp2, ref["density"], ref["Q"]))
    # (c) an id outside the cells
    j3 = sp.j.copy()
    j3[int(sp.p[9]) - 1] = 9000
    got = ctx.dkl_csr(sp.p, j3, sp.x)
    assert np.isnan(got[8])
    ok = np.ones(got.size, bool)
    ok[8] = False
    _check(got[ok], want[ok])


def test_randomization_batch_through_tiles(ctx):
    """K4-sparse, ids given (R/haystack_continuous.R:266-283) and seeded: 3 genes x 7 permutations against the
    oracle fed the same ids."""
    from singlecellhaystack_b200.device import perm_sample
    n = 15000
    xs, grid, sp = _problem(n, 4, 100, 40, 0.005, 0.4, seed=11)
    ref = ho.density_stage(xs, grid)
    ctx.density(xs, grid)
    rows = [0, 17, 39]
    vals = [sp.x[sp.p[r]:sp.p[r + 1]] for r in rows]
    src = synth.perm_source(5)
    ids = [src("sparse", r, n, v.size, 7) for r, v in zip(rows, vals)]
    want = ho.dkl_randomized_sparse(vals, ids, ref["density"], ref["Q"])
    got = ctx.dkl_perm_csr(vals, [np.asfortranarray(a.T) for a in ids], index_base=1)
    _check(got, want)
    seed = 99
    got_s = ctx.dkl_perm_csr_seeded(vals, 7, seed)
    ids_s = [np.stack([perm_sample(seed, s * 7 + r, n, v.size, 1) for r in range(7)]) for s, v in enumerate(vals)]
    want_s = ho.dkl_randomized_sparse(vals, ids_s, ref["density"], ref["Q"])
    _check(got_s, want_s)
    # an id outside [1, N] in one permutation: that entry of the batch is NaN, the others unchanged
    bad = [a.copy() for a in ids]
    bad[1][3, 0] = n + 5
    got_b = ctx.dkl_perm_csr(vals, [np.asfortranarray(a.T) for a in bad], index_base=1)
    assert np.isnan(got_b[1, 3])
    mask = np.ones_like(got_b, bool)
    mask[1, 3] = False
    _check(got_b[mask], want[mask])


def test_packed_id_upload_is_transparent(ctx, monkeypatch):
    """Rows with more than ~4 % stored values cross PCIe as bitmaps over the cells and are expanded on the device
    (hs_api.cu: pack_ids, hs_convert.cu: k_bitmap_to_ids).  Same ids on the device, so the scores are bit-identical to
    the plain upload -- over several pieces of rows, several ring buffers per piece, pinned or pageable, and when one
    piece holds an unsorted row (that piece then goes up as it is)."""
    n, m = 3001, 6000
    x, labels = synth.make_embedding(n, 4, seed=21)
    xs, _, _ = ho.scale_coords(x)
    grid = synth.make_grid(xs, 40, seed=21)
    p, j, v = synth.make_csr_numpy(m, n, 0.10, labels, seed=21)
    ctx.density(xs, grid)
    monkeypatch.setenv("HS_PACK_IDS", "0")
    plain = ctx.dkl_csr(p, j, v)
    monkeypatch.setenv("HS_PACK_IDS", "1")
    monkeypatch.setenv("HS_PIECE_ROWS", "700")
    monkeypatch.setenv("HS_PIECE_VALUES", "200000")
    monkeypatch.setenv("HS_STAGE_VALUES", "30000")       # a ring buffer holds ~300 rows' bitmaps
    packed = ctx.dkl_csr(p, j, v)
    np.testing.assert_array_equal(plain, packed)
    # the oracle on a few rows, for good measure
    ref = ho.density_stage(xs, grid)
    rows = [0, 699, 700, 5999]
    sub = ho.RowSparse(p=np.concatenate([[0], np.cumsum([p[r + 1] - p[r] for r in rows])]).astype(np.int64),
                       j=np.concatenate([j[p[r]:p[r + 1]] for r in rows]), x=np.concatenate([v[p[r]:p[r + 1]] for r in rows]),
                       shape=(len(rows), n))
    _check(packed[rows], ho.dkl_observed(sub, ref["density"], ref["Q"]))
    # an unsorted row in the third piece
    j2, v2 = j.copy(), v.copy()
    a, b = int(p[1500]), int(p[1501])
    j2[a], j2[b - 1] = j2[b - 1], j2[a]
    v2[a], v2[b - 1] = v2[b - 1], v2[a]
    swapped = ctx.dkl_csr(p, j2, v2)
    ok = np.ones(m, bool)
    ok[1500] = False
    np.testing.assert_array_equal(swapped[ok], plain[ok])
    assert abs(swapped[1500] - plain[1500]) <= 2e-5 * abs(plain[1500])      # scored by the gather kernel


def test_begin_wait_pair_overlaps_host_work(ctx, toy):
    """hs_dkl_perm_csr_begin / _wait: the same scores as the blocking call; the context refuses other work while a call
    is in flight; batches can be pipelined (draw the next ids while the previous batch is scored)."""
    from singlecellhaystack_b200.device import HaystackLibraryError
    xs, _, _ = ho.scale_coords(toy["coord"])
    ctx.density(xs, toy["grid100"])
    rng = np.random.default_rng(9)
    n = 601

    def batch(k):
        vals = [rng.gamma(2.0, 1.0, size=int(s)) for s in rng.integers(1, 400, size=k)]
        inds = [np.asfortranarray(np.stack([rng.permutation(n)[:v.size] + 1 for _ in range(6)]).T.astype(np.int32)) for v in vals]
        return vals, inds

    b0, b1 = batch(5), batch(7)
    want0, want1 = ctx.dkl_perm_csr(*b0), ctx.dkl_perm_csr(*b1)
    ctx.dkl_perm_csr_begin(*b0)
    with pytest.raises(HaystackLibraryError):              # single user: nothing else may run on the context now
        ctx.dkl_perm_csr(*b1)
    got0 = ctx.dkl_perm_csr_wait()
    np.testing.assert_array_equal(got0, want0)
    # pipelined: begin(b1), prepare the next batch on this thread, wait
    ctx.dkl_perm_csr_begin(*b1)
    b2 = batch(3)
    got1 = ctx.dkl_perm_csr_wait()
    np.testing.assert_array_equal(got1, want1)
    np.testing.assert_array_equal(ctx.dkl_perm_csr(*b2), ctx.dkl_perm_csr(*b2))
    with pytest.raises(HaystackLibraryError, match="no asynchronous call"):
        ctx._check(ctx._lib.hs_dkl_perm_csr_wait(ctx._h, None), "hs_dkl_perm_csr_wait")

"""Parity at the exact shapes BASELINE.json names: MOCA-like 100k x 16 811 (50-D), Visium-like
2 696 x 12 382 and x 30 000 (2-D), pseudotime 1 399 x 23 341 (1-D), Slide-seqV2-like 53 173 x 11 390 (2-D) -- the
synthetic stand-ins bench.py times.  Per shape: density stage against the oracle (bandwidth 1e-12, nearest grid point
bit-exact, Q 1e-9), D_KL of 64 genes spread over the matrix and of 2 genes x 10 permutations within 1e-5 relative
(R/haystack_continuous.R:199-213, :266-283)."""
import argparse

import numpy as np
import pytest

from oracle import haystack_oracle as ho
from singlecellhaystack_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c2", "c3", "c3b", "c4", "c4b"])
def test_config_shape_parity(ctx, name):
    import torch
    import bench
    a = argparse.Namespace(cells=None, genes=None, density=None, dims=None, grid_points=100, rand_genes=100, rand_count=100)
    bench.apply_config(a, name)
    dev = torch.device("cuda", 0)
    xs, labels, grid = bench.make_coords(a)
    p, cols, vals = bench.make_csr_device(a, labels, 0, dev)
    assert p.numel() == a.genes + 1
    # ---- density stage ----
    ref = ho.density_stage(xs, grid)
    out = ctx.density(xs, grid, want_nearest=True)
    assert abs(out["bandwidth"] - ref["bandwidth"]) <= 1e-12 * ref["bandwidth"]
    np.testing.assert_array_equal(out["nearest"], ref["nearest"])
    np.testing.assert_allclose(out["Q"], ref["Q"], rtol=1e-9)
    # ---- observed D_KL: all genes on the device, 64 of them against the oracle ----
    dkl = torch.empty(a.genes, dtype=torch.float64, device=dev)
    ctx.dkl_csr_dev(p, cols, vals, dkl)
    ctx.synchronize()
    got = dkl.cpu().numpy()
    assert np.isfinite(got).all()
    p_h = p.cpu().numpy()
    pick = np.unique(np.linspace(0, a.genes - 1, 64).astype(np.int64))
    for g in pick:
        lo, hi = int(p_h[g]), int(p_h[g + 1])
        if hi == lo:
            continue
        ind1 = cols[lo:hi].cpu().numpy().astype(np.int64) + 1
        val = vals[lo:hi].double().cpu().numpy()
        want = ho.get_D_KL_continuous_highD_SPARSE(ind1, val, ref["density"], ref["Q"], ho.PSEUDO)
        assert abs(got[g] - want) <= 1e-5 * abs(want) + 1e-12, (name, int(g), got[g], want)
    # ---- randomized D_KL: 2 genes x 10 permutations, ids given (R's sample() stand-in) ----
    nnz = np.diff(p_h)
    rows = [int(np.argmax(nnz)), int(np.argsort(nnz)[len(nnz) // 2])]
    rows = [r for r in rows if nnz[r] > 0]
    vals_h = [vals[int(p_h[r]):int(p_h[r + 1])].double().cpu().numpy() for r in rows]
    src = synth.perm_source(7)
    ids = [src("sparse", r, a.cells, v.size, 10) for r, v in zip(rows, vals_h)]
    want_r = ho.dkl_randomized_sparse(vals_h, ids, ref["density"], ref["Q"])
    got_r = ctx.dkl_perm_csr(vals_h, [np.asfortranarray(i.T) for i in ids], index_base=1)
    np.testing.assert_allclose(got_r, want_r, rtol=1e-5, atol=1e-12)

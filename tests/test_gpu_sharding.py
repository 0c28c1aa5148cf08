"""Multi-GPU pieces exercised on one GPU: the cell-sharded density stage (hs_density_shard_*) reproduces hs_density_dev
bit for bit when several ranks' parts are driven through one context, and the per-unit streams of
hs_dkl_perm_csr_seeded_units_dev make the randomized matrix independent of how the (gene, permutation block) units
are dealt (distributed.deal_permutation_units)."""
import ctypes as C

import numpy as np
import pytest

from oracle import haystack_oracle as ho
from singlecellhaystack_b200 import synth
from singlecellhaystack_b200.distributed import deal_permutation_units

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,world", [(5000, 3), (1000, 8), (129, 2), (70000, 4)])
def test_cell_sharded_density_is_bit_identical(ctx, n, world):
    import torch
    d, g = 6, 40
    rng = np.random.default_rng(n)
    xs, _, _ = ho.scale_coords(rng.normal(size=(n, d)))
    grid = xs[rng.choice(n, size=g, replace=False)] + 0.01
    dev = torch.device("cuda", 0)
    x_dev = torch.as_tensor(np.ascontiguousarray(xs.T), device=dev)
    g_dev = torch.as_tensor(np.ascontiguousarray(grid.T), device=dev)
    ref = ctx.density_dev(x_dev, g_dev)
    d_ptr, d_bytes, q_ptr, q_bytes = ctx.density_buffers(n, g)
    from singlecellhaystack_b200.device import _wrap_device_ptr
    want_d = _wrap_device_ptr(d_ptr, d_bytes, dev).clone()
    lib, h = ctx._lib, ctx._h
    vp = lambda t: C.c_void_p(t.data_ptr())
    # pass 1: every rank's distances -> the shared row-minimum buffer is complete (what the all-gather achieves)
    for r in range(world):
        assert lib.hs_density_shard_dist_dev(h, vp(x_dev), n, d, vp(g_dev), g, r, world) == 0
    # pass 2: each rank's distances again (its scratch was overwritten by the others), then its densities
    for r in range(world):
        assert lib.hs_density_shard_dist_dev(h, vp(x_dev), n, d, vp(g_dev), g, r, world) == 0
        assert lib.hs_density_shard_dens_dev(h, n, g, None, r, world) == 0
    bw = C.c_double(0.0)
    q = np.empty(g)
    assert lib.hs_density_shard_finish(h, n, g, C.byref(bw), C.c_void_p(q.ctypes.data)) == 0
    assert bw.value == ref["bandwidth"]
    np.testing.assert_array_equal(q, ref["Q"])
    d_ptr2, d_bytes2, _, _ = ctx.density_buffers(n, g)
    got_d = _wrap_device_ptr(d_ptr2, d_bytes2, dev)
    assert torch.equal(got_d, want_d)
    # and the state scores like the unsharded one
    p, j, v = synth.make_csr_numpy(20, n, 0.05, np.zeros(n, dtype=np.int64))
    dens = ho.density_stage(xs, grid)
    want = ho.dkl_observed(ho.RowSparse(p=p, j=j, x=v, shape=(20, n)), dens["density"], dens["Q"])
    np.testing.assert_allclose(ctx.dkl_csr(p, j, v), want, rtol=1e-5)


def test_unit_streams_make_the_randomization_independent_of_the_deal(ctx):
    import torch
    n, g, R = 6000, 30, 8
    rng = np.random.default_rng(3)
    xs, _, _ = ho.scale_coords(rng.normal(size=(n, 3)))
    grid = xs[rng.choice(n, size=g, replace=False)]
    ctx.density(xs, grid)
    dev = torch.device("cuda", 0)
    nnz = [500, 40, 3000, 7, 1200]
    vals = [np.log1p(1 + rng.poisson(2.0, size=k)).astype(np.float32) for k in nnz]
    whole = ctx.dkl_perm_csr_seeded([v.astype(np.float64) for v in vals], R, 77)           # (S, R), stream = s * R + r
    for world in (1, 2, 3):
        got = np.full((len(nnz), R), np.nan)
        for units in deal_permutation_units(nnz, R, world, block=4):
            if not units:
                continue
            gp = np.concatenate([[0], np.cumsum([nnz[s] for s, _, _ in units])]).astype(np.int64)
            val = np.concatenate([vals[s] for s, _, _ in units])
            streams = np.array([s * R + r_lo for s, r_lo, _ in units], dtype=np.int64)
            out = torch.empty(len(units) * 4, dtype=torch.float64, device=dev)
            ctx.dkl_perm_csr_seeded_units_dev(torch.as_tensor(gp, device=dev), torch.as_tensor(val, device=dev), 4, 77,
                                              torch.as_tensor(streams, device=dev), out)
            ctx.synchronize()            # *_dev calls are asynchronous on the context's stream
            res = out.cpu().numpy().reshape(4, len(units))                                   # units x 4, column-major
            for u, (s, r_lo, r_hi) in enumerate(units):
                got[s, r_lo:r_hi] = res[:r_hi - r_lo, u]
        # the same draws whoever makes them; the cell segments of the contraction depend on the batch size, so the
        # fp64 sums are grouped differently: equal to rounding, not bit for bit
        np.testing.assert_allclose(got, whole, rtol=2e-6)

"""Shape sweep of the device path against the oracle: grid sizes that select every kernel variant
(1, 2, 4, 8 float4 per lane; tensor-core and SIMT dense kernels), tiny and ragged inputs."""
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

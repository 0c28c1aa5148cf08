"""The R .Call shim (rshim/src/haystack_b200_shim.c) RUN on the B200 through a mock of R's C API: the routines the
patched haystack_continuous_highD() calls return what the C ABI returns for the same inputs, with R's conventions
(column-major matrices, dgRMatrix slots, 1-based ids from sample()), and library errors come back as R errors."""
import numpy as np
import pytest

from oracle import haystack_oracle as ho
from rshim_harness import RError, Shim

pytestmark = pytest.mark.gpu


def test_shim_routines_match_the_c_abi(ctx, toy):
    sh = Shim()
    coord, expr = toy["coord"].astype(np.float64), toy["expression"].astype(np.float64)
    xs, _, _ = ho.scale_coords(coord)
    grid = toy["grid100"]
    dev = sh.call("hsb_create", sh.integer([0]))
    with pytest.raises(RError, match="no density stage resident"):          # library error -> R error, not a crash
        sh.call("hsb_dkl_dense", dev, sh.real(expr))
    res = sh.to_numpy(sh.call("hsb_density", dev, sh.real(xs), sh.real(grid), sh.nil, sh.logical(True)))
    want = ctx.density(xs, grid, want_densities=True)
    assert res[0][0] == want["bandwidth"]
    np.testing.assert_array_equal(res[1], want["Q"])
    np.testing.assert_array_equal(res[2], want["densities"])
    np.testing.assert_array_equal(sh.to_numpy(sh.call("hsb_dkl_dense", dev, sh.real(expr))), ctx.dkl_dense(expr))
    sp = ho.RowSparse.from_dense(expr)
    got = sh.to_numpy(sh.call("hsb_dkl_csr", dev, sh.integer(sp.p), sh.integer(sp.j), sh.real(sp.x)))
    np.testing.assert_array_equal(got, ctx.dkl_csr(sp.p.astype(np.int32), sp.j, sp.x))
    np.testing.assert_allclose(got, toy["dkl100"], rtol=1e-5)
    # randomization, dense form: perm = N x R integer matrix of sample.int(N) columns (1-based)
    rng = np.random.default_rng(0)
    perm = np.stack([rng.permutation(601) + 1 for _ in range(5)], axis=1)
    got = sh.to_numpy(sh.call("hsb_dkl_perm_dense", dev, sh.real(expr[7]), sh.integer(perm)))
    np.testing.assert_array_equal(got, ctx.dkl_perm_dense(expr[7], perm, index_base=1))
    # sparse form: gene.ptr as doubles, ids gene by gene as nnz x R column-major blocks
    rows = [3, 40]
    vals = [sp.x[sp.p[r]:sp.p[r + 1]] for r in rows]
    inds = [np.stack([rng.choice(601, size=v.size, replace=False) + 1 for _ in range(4)], axis=1) for v in vals]
    gp = np.concatenate([[0], np.cumsum([v.size for v in vals])]).astype(np.float64)
    got = sh.to_numpy(sh.call("hsb_dkl_perm_csr", dev, sh.real(gp), sh.real(np.concatenate(vals)),
                              sh.integer(np.concatenate([a.ravel(order="F") for a in inds])), sh.integer([4])))
    np.testing.assert_array_equal(got, ctx.dkl_perm_csr(vals, inds, index_base=1))
    want_r = ho.dkl_randomized_sparse(vals, [a.T for a in inds], ho.density_stage(xs, grid)["density"], want["Q"])
    np.testing.assert_allclose(got, want_r, rtol=1e-5)
    # the asynchronous pair: same numbers; a second begin before the wait and a wait without a begin are R errors
    args = (dev, sh.real(gp), sh.real(np.concatenate(vals)),
            sh.integer(np.concatenate([a.ravel(order="F") for a in inds])), sh.integer([4]))
    sh.call("hsb_dkl_perm_csr_begin", *args)
    with pytest.raises(RError, match="already in flight"):
        sh.call("hsb_dkl_perm_csr_begin", *args)
    np.testing.assert_array_equal(sh.to_numpy(sh.call("hsb_dkl_perm_csr_wait", dev)), got)
    with pytest.raises(RError, match="no batch in flight"):
        sh.call("hsb_dkl_perm_csr_wait", dev)
    # one call for scores + seeded randomization; ids reproducible on the R side
    out = sh.to_numpy(sh.call("hsb_dkl_csr_randomized", dev, sh.integer(sp.p), sh.integer(sp.j), sh.real(sp.x),
                              sh.integer(np.array(rows) + 1), sh.integer([4]), sh.real([123.0])))
    d2, r2 = ctx.dkl_csr_randomized(sp.p.astype(np.int32), sp.j, sp.x, rows, 4, 123)
    np.testing.assert_array_equal(out[0], d2)
    np.testing.assert_array_equal(out[1], r2)
    ids = sh.to_numpy(sh.call("hsb_perm_sample", sh.real([123.0]), sh.real([5.0]), sh.real([601.0]), sh.real([17.0])))
    assert ids.min() >= 1 and ids.max() <= 601 and len(set(ids.tolist())) == 17
    sh.lib.rmock_finalize(dev)          # the external pointer's finalizer destroys the context
    with pytest.raises(RError, match="destroyed"):
        sh.call("hsb_dkl_dense", dev, sh.real(expr))

"""prepare_clustering / hclust_haystack / kmeans_haystack (R/haystack_continuous.R:604-697, R/s3.R:213-253): the
gene x grid profiles from the device against the oracle's restatement, dense and sparse input, and the two clustering
front ends on top of them."""
import numpy as np
import pytest

from oracle import haystack_oracle as ho
from singlecellhaystack_b200 import clustering as cl
from singlecellhaystack_b200 import sparse as sp

pytestmark = pytest.mark.gpu


def test_profiles_match_the_oracle(ctx, toy):
    coord, expr = toy["coord"].astype(np.float64), toy["expression"].astype(np.float64)[:120]
    xs, center, sc = ho.scale_coords(coord)
    grid_unscaled = toy["grid60"] * sc + center            # what haystack() returns as info$grid.coordinates
    want = ho.prepare_clustering(coord, expr, grid_unscaled, scale=True)
    got_d = cl.prepare_clustering(coord, expr, grid_unscaled, ctx=ctx)
    got_s = cl.prepare_clustering(coord, sp.as_CsparseMatrix(expr), grid_unscaled, ctx=ctx)
    for got in (got_d, got_s):
        assert got.shape == (120, 60)
        np.testing.assert_allclose(got.sum(axis=1), 1.0, rtol=1e-12)
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-10)     # rows sum to 1: entries ~1e-2; the density matrix is fp32
    # dense and sparse rows: the same sums, folded into fp64 in different groups of 32 fp32 products
    np.testing.assert_allclose(got_d, got_s, rtol=2e-6, atol=1e-12)
    # scale = FALSE: coordinates and grid used as they are
    want_u = ho.prepare_clustering(xs, expr, toy["grid60"], scale=False)
    np.testing.assert_allclose(cl.prepare_clustering(xs, expr, toy["grid60"], scale=False, ctx=ctx), want_u, rtol=1e-5, atol=1e-10)


def test_clustering_front_ends(ctx, toy):
    coord, expr = toy["coord"].astype(np.float64), toy["expression"].astype(np.float64)
    xs, center, sc = ho.scale_coords(coord)
    grid_unscaled = toy["grid60"] * sc + center
    top = np.argsort(toy["dkl60"])[-60:]                    # the most structured genes
    link = cl.hclust_haystack(coord, expr[top], grid_unscaled, ctx=ctx)
    assert link.shape == (59, 4) and np.all(np.diff(link[:, 2]) >= -1e-12)       # a valid, monotone dendrogram
    km = cl.kmeans_haystack(coord, expr[top], grid_unscaled, k=4, random_state=0, ctx=ctx)
    assert km.cluster_centers_.shape == (4, 60) and len(set(km.labels_.tolist())) == 4
    # genes in one k-means cluster have more similar profiles than genes in different clusters
    prof = cl.prepare_clustering(coord, expr[top], grid_unscaled, ctx=ctx)
    c = np.corrcoef(prof)
    same = km.labels_[:, None] == km.labels_[None, :]
    off = ~np.eye(60, dtype=bool)
    assert c[same & off].mean() > c[~same].mean()

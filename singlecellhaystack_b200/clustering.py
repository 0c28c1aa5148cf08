"""The step after the path: gene clustering on the grid profiles.

Mirrors (R dots become underscores)
  prepare_clustering            R/haystack_continuous.R:604-666
  hclust_haystack(_continuous)  R/s3.R:213-227, R/haystack_continuous.R:668-682
  kmeans_haystack(_continuous)  R/s3.R:239-253, R/haystack_continuous.R:684-697
The gene x grid profile matrix comes from the B200 (hs_density + hs_profile_csr / hs_profile_dense: the same density
stage and contraction as haystack(), without the KL reduction); the clustering itself (cor / hclust / kmeans in the
reference) stays on the host, here with scipy / scikit-learn.  No CPU fallback for the profiles.
"""
from __future__ import annotations

import numpy as np

from . import haystack as _hs
from . import sparse as _sparse

__all__ = ["prepare_clustering", "hclust_haystack", "kmeans_haystack"]


def prepare_clustering(x, expression, grid_coordinates, scale=True, *, ctx=None, device=0):
    """(genes, grid points) matrix of per-gene densities, every row rescaled to sum to 1.  `grid_coordinates` are in
    the ORIGINAL space (res["info"]["grid_coordinates"] of haystack()); with scale=True they are moved into the
    scaled space with x's own centre and scale (:606-612)."""
    from .device import DeviceContext
    x = np.asarray(x.values if hasattr(x, "values") and hasattr(x, "columns") else x, dtype=np.float64)
    grid = np.asarray(grid_coordinates, dtype=np.float64)
    if x.ndim != 2 or grid.ndim != 2 or x.shape[1] != grid.shape[1]:
        raise ValueError("Coordinates and grid points have different number of columns.")
    if scale:
        x, center, sc = _hs.scale(x)
        grid = (grid - center) / sc
    sparse_in = _sparse.is_sparse(expression)
    if sparse_in:
        expression = _sparse.as_RsparseMatrix(expression)          # :615-618
    else:
        if hasattr(expression, "values") and hasattr(expression, "index"):
            expression = expression.values
        expression = np.asarray(expression, dtype=np.float64)
        if expression.ndim != 2:
            raise TypeError("'expression' must be a matrix or dgRMatrix")
    if expression.shape[1] != x.shape[0]:
        raise ValueError("The number of columns in 'expression' must be the same as the rows in 'x'")
    own = ctx is None
    if own:
        ctx = DeviceContext(device)
    try:
        ctx.density(x, grid)                                         # :624-632
        if sparse_in:
            return ctx.profile_csr(expression.p, expression.j, expression.x)      # :651-657, :664
        return ctx.profile_dense(expression)                         # :643-649, :664
    finally:
        if own:
            ctx.close()


def _rank_rows(a):
    from scipy.stats import rankdata
    return np.apply_along_axis(rankdata, 1, a)


def hclust_haystack(x, expression, grid_coordinates, hclust_method="ward.D", cor_method="spearman", scale=True, **kw):
    """hclust(as.dist(1 - cor(t(densities), method = cor.method)), method = hclust.method)  (:676-677).
    Returns a scipy linkage matrix.  R's "ward.D" applies Ward's update to the unsquared dissimilarities, which scipy
    does not offer; "ward.D2", "average", "complete", "single" map directly (ward.D is served by ward.D2's linkage on
    the same dissimilarities -- documented difference)."""
    from scipy.cluster.hierarchy import linkage
    from scipy.spatial.distance import squareform
    dens = prepare_clustering(x, expression, grid_coordinates, scale=scale, **kw)
    if cor_method == "spearman":
        dens = _rank_rows(dens)
    elif cor_method != "pearson":
        raise ValueError("cor.method should be 'spearman' or 'pearson'")
    d = 1.0 - np.corrcoef(dens)
    np.fill_diagonal(d, 0.0)
    d = np.clip((d + d.T) / 2, 0.0, None)
    method = {"ward.D": "ward", "ward.D2": "ward", "average": "average", "complete": "complete", "single": "single"}.get(hclust_method)
    if method is None:
        raise ValueError(f"hclust.method '{hclust_method}' is not supported")
    return linkage(squareform(d, checks=False), method=method)


def kmeans_haystack(x, expression, grid_coordinates, k, scale=True, random_state=None, **kw):
    """kmeans(x = densities, centers = k)  (:695).  Returns the fitted sklearn KMeans object (labels_ are 0-based)."""
    from sklearn.cluster import KMeans
    dens = prepare_clustering(x, expression, grid_coordinates, scale=scale, **kw)
    return KMeans(n_clusters=int(k), n_init=10, random_state=random_state).fit(dens)

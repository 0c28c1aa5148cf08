"""Host layer above the C ABI: the reference's front door for the continuous high-D path.

Mirrors, name for name and argument for argument (R dots become underscores):
  haystack()                    R/s3.R:21-87   (matrix / data.frame / sparse / AnnData-like dispatch)
  haystack_continuous_highD()   R/haystack_continuous.R:27-339
  show_result_haystack()        R/haystack.R:587-634
Everything numeric between "grid decided" and "spline fit" runs on the B200 through
libhaystack_b200.so (density stage, observed D_KL, randomized D_KL); validation, row statistics,
scaling, grid selection, gene selection, the random draws and the spline fit stay on the host, as
north_star prescribes for the R drop-in.  The validation messages are the reference's, verbatim.

There is no CPU fallback: without the CUDA library / a B200 the DeviceContext constructor raises.
"""
from __future__ import annotations

import math
import os
import sys
import warnings

import numpy as np

from . import grid as _grid
from . import sparse as _sparse
from . import splinefit as _fit

__all__ = ["haystack", "haystack_continuous_highD", "show_result_haystack", "Haystack",
           "select_genes_to_randomize", "row_mean_sd", "scale"]


def _message(verbose, *parts):
    if verbose:
        print("".join(str(p) for p in parts), file=sys.stderr)


class Haystack(dict):
    """The S3 object of class "haystack": res["results"] (pandas DataFrame with columns D_KL,
    log.p.vals, log.p.adj and gene row names) and res["info"] (R/haystack_continuous.R:318-337)."""

    @property
    def results(self):
        return self["results"]

    @property
    def info(self):
        return self["info"]

    def __repr__(self):
        r = self.get("results")
        n = 0 if r is None else len(r)
        return f"<haystack result: {n} features, method {self.get('info', {}).get('method')}>"


# --------------------------------------------------------------------------------------------
# small host helpers restating base-R behaviour on the path
# --------------------------------------------------------------------------------------------
def scale(x):
    """base::scale(x): returns (scaled, scaled:center, scaled:scale) -- R/haystack_continuous.R:139-145."""
    x = np.asarray(x, dtype=np.float64)
    center = x.mean(axis=0)
    xc = x - center
    sc = np.sqrt((xc * xc).sum(axis=0) / (x.shape[0] - 1))
    return xc / sc, center, sc


_scale = scale      # the orchestrator's `scale` argument shadows the function


def row_mean_sd(expression):
    """Matrix::rowMeans and the per-row sample sd (n-1, zeros included) -- R/haystack_continuous.R:57-74.
    Two-pass, fp64."""
    if isinstance(expression, _sparse.dgRMatrix):
        m, n = expression.Dim
        p = expression.p.astype(np.int64)
        nnz = np.diff(p)
        rows = np.repeat(np.arange(m), nnz)
        s = np.bincount(rows, weights=expression.x, minlength=m)
        mean = s / n
        dev = expression.x - mean[rows]
        ss = np.bincount(rows, weights=dev * dev, minlength=m) + (n - nnz) * mean * mean
        with np.errstate(invalid="ignore", divide="ignore"):
            sd = np.sqrt(ss / (n - 1))
        return mean, sd
    e = np.asarray(expression, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        return e.mean(axis=1), e.std(axis=1, ddof=1)


def _seq_length_out(frm, to, n):
    """seq(from, to, length.out = n) of base R."""
    n = int(n)
    if n <= 0:
        return np.zeros(0)
    if n == 1:
        return np.array([float(frm)])
    by = (float(to) - float(frm)) / (n - 1)
    out = float(frm) + np.arange(n, dtype=np.float64) * by
    out[-1] = float(to)
    return out


def selection_ranks(count_genes, n_genes_to_randomize, method="heavytails"):
    """The 1-based positions in order(coeffVar) that R/haystack_continuous.R:227-236 picks: integer arithmetic exactly
    as R (as.integer() truncates, floor() floors)."""
    s = int(n_genes_to_randomize)
    if method == "uniform":
        pos = np.floor(_seq_length_out(1, count_genes, s)).astype(np.int64)
    elif method == "heavytails":
        if s < 18 or count_genes < 18:
            # R: seq(10, M - 9, length.out = S - 18) stops with "'length.out' must be a non-negative number"
            # (R/haystack_continuous.R:233-235); fail as clearly instead of returning 18 (possibly repeated) rows
            raise ValueError("'heavytails' selection needs at least 18 genes to randomize "
                             f"(n.genes.to.randomize = {s}, genes = {count_genes}); use "
                             "selection.method.genes.to.randomize = 'uniform'")
        mid = np.trunc(_seq_length_out(10, count_genes - 9, s - 18)).astype(np.int64)
        pos = np.concatenate([np.arange(1, 10), mid, count_genes - np.arange(8, -1, -1)])
    else:
        # the reference silently leaves genes.to.randomize undefined here and fails later
        raise ValueError("'selection.method.genes.to.randomize' should be either 'uniform' or 'heavytails'")
    return pos


def select_genes_to_randomize(coeffVar, n_genes_to_randomize, method="heavytails"):
    """R/haystack_continuous.R:227-236: order() is stable with NA last.  Returns 0-based row indices."""
    cv = np.asarray(coeffVar, dtype=np.float64)
    pos = selection_ranks(cv.size, n_genes_to_randomize, method)
    o = np.argsort(cv, kind="stable")
    return o[pos - 1]


# --------------------------------------------------------------------------------------------
# random draws (R: sample()); injectable so tests can feed oracle and device the same indices
# --------------------------------------------------------------------------------------------
class _NumpyPermSource:
    """sample(x = N, size = k) / sample.int(N) stand-in on a numpy Generator (1-based ids).
    Under R the shim calls R's own sample(), so set.seed() reproducibility is R's; here it is numpy's."""

    def __init__(self, rng):
        self.rng = np.random.default_rng(rng)

    def __call__(self, kind, gene, n, size, count):
        out = np.empty((count, size), dtype=np.int32)
        for r in range(count):
            if kind == "dense" or size == n:
                out[r] = self.rng.permutation(n)[:size] + 1
            else:
                out[r] = self.rng.choice(n, size=size, replace=False, shuffle=True) + 1
        return out


# --------------------------------------------------------------------------------------------
# the orchestrator
# --------------------------------------------------------------------------------------------
def haystack_continuous_highD(x, expression, grid_points=100, weights_advanced_Q=None, dir_randomization=None,
                              scale=True, grid_method="centroid", randomization_count=100,
                              n_genes_to_randomize=100, selection_method_genes_to_randomize="heavytails",
                              grid_coord=None, spline_method="ns", *, gene_names=None, rng=None,
                              perm_source=None, cv_folds=None, row_stats=None, ctx=None, device=0,
                              verbose=True, return_densities=True, convert_on_device=False):
    """The main Haystack function for continuous expression in a high-dimensional space
    (R/haystack_continuous.R:27-339).  Positional / keyword arguments up to `spline_method` are the
    reference's; the keyword-only ones are host plumbing:
      gene_names   row names of `expression` (R takes them from rownames())
      rng          seed or numpy Generator for every random draw (R: set.seed()); or rcompat.RRandom(seed): R's own
                   stream -- grid, randomizations and folds are then what R draws after set.seed(seed)
      perm_source  callable(kind, gene0, n, size, count) -> (count, size) 1-based ids, replaces sample();
                   or the string "device": the draws are made on the GPU from a seed taken from `rng`
                   (hs_dkl_perm_csr_seeded / hs_dkl_perm_dense_seeded; stream = position in
                   genes_to_randomize * randomization_count + r, reproducible on the host with
                   device.perm_sample)
      cv_folds     (folds_for_mean_model, folds_for_sd_model) replacing sample(rep(1:10, ...))
      row_stats    (mean, sd) per gene computed by the caller (the R shim passes R's own
                   Matrix::rowMeans / rowSds, so the integer gene selection is R's even for tied genes);
                   or "device": sparse input is reduced by hs_row_stats_csr (fp64, two-pass) on the GPU
      ctx/device   an existing DeviceContext, or the CUDA device ordinal to create one on
      convert_on_device  dgCMatrix -> dgRMatrix (:113-116) with hs_csc_to_csr on the GPU instead of on the host
      return_densities  fill info["densities"] (N x G fp64, 800 MB at 1M cells) as the reference does
    """
    from .device import DeviceContext

    _message(verbose, "### calling haystack_continuous_highD()...")
    x_in = x
    x = np.asarray(x, dtype=None)
    if grid_coord is not None:
        grid_coord = np.asarray(grid_coord, dtype=np.float64)
        if grid_coord.ndim == 1:
            grid_coord = grid_coord.reshape(-1, 1)
        grid_points = grid_coord.shape[0]
        if x.ndim == 2 and x.shape[1] != grid_coord.shape[1]:
            raise ValueError("Coordinates and grid points have different number of columns.")
        xc, gc = getattr(x_in, "columns", None), getattr(grid_coord, "columns", None)
        if (xc is not None or gc is not None) and \
                (list(xc) if xc is not None else []) != (list(gc) if gc is not None else []):
            raise ValueError("Coordinates and grid points have different column names.")

    # ---- expression container -------------------------------------------------------------
    sparse_in = _sparse.is_sparse(expression)
    if not sparse_in:
        if hasattr(expression, "index") and gene_names is None and hasattr(expression, "values"):
            gene_names = list(expression.index)
            expression = expression.values
        expression = np.asarray(expression)
        if expression.ndim != 2 or not np.issubdtype(expression.dtype, np.number):
            raise TypeError("'expression' must be a matrix, dgCMatrix, or dgRMatrix")
    else:
        if gene_names is None:
            gene_names = getattr(expression, "rownames", None)

    _message(verbose, "### Calculating row-wise mean and SD... ")
    if sparse_in:
        # the reference computes the statistics first and converts afterwards (:113-116); the order is
        # immaterial for the numbers, and one CSR pass serves both
        own_ctx = False
        if isinstance(expression, _sparse.dgCMatrix) or not isinstance(expression, _sparse.dgRMatrix):
            _message(verbose, "### converting expression data from dgCMatrix to dgRMatrix")
            if convert_on_device:             # hs_csc_to_csr instead of the host counting sort
                if ctx is None:
                    ctx, own_ctx = DeviceContext(device), True
                expression = _sparse.as_CsparseMatrix(expression)
        expression = _sparse.as_RsparseMatrix(expression, ctx=ctx if convert_on_device else None)
    else:
        own_ctx = False
    if isinstance(row_stats, str):
        if row_stats != "device":
            raise ValueError("row_stats must be a (mean, sd) pair, None or 'device'")
        if sparse_in:                  # one fp64 two-pass kernel instead of a host pass over the matrix
            if ctx is None:
                ctx, own_ctx = DeviceContext(device), True
            expr_mean, expr_sd = ctx.row_stats_csr(expression.p, expression.x, expression.shape[1])
        else:
            expr_mean, expr_sd = row_mean_sd(expression)
    elif row_stats is not None:
        expr_mean = np.asarray(row_stats[0], dtype=np.float64).copy()
        expr_sd = np.asarray(row_stats[1], dtype=np.float64).copy()
        if expr_mean.shape != (expression.shape[0],) or expr_sd.shape != expr_mean.shape:
            raise ValueError("row_stats must hold one mean and one sd per row of 'expression'")
    else:
        expr_mean, expr_sd = row_mean_sd(expression)
    if expr_mean.size and np.nanmin(expr_mean) < 0:
        raise ValueError("Some features have an average signal < 0. Expect average signal >= 0.")
    sel_bad = expr_sd == 0
    if sel_bad.any():
        expression = expression.row_subset(~sel_bad) if sparse_in else expression[~sel_bad]
    if gene_names is not None:
        gene_names = [g for g, bad in zip(gene_names, sel_bad) if not bad]
    stats_raw = (expr_mean[~sel_bad], expr_sd[~sel_bad])      # what hs_select_by_cv_dev takes (it adds the 1e-300 itself)
    expr_mean = expr_mean[~sel_bad] + 1e-300
    expr_sd = expr_sd[~sel_bad] + 1e-300
    _message(verbose, "### Filtered ", int(sel_bad.sum()), " genes with zero variance...")
    _message(verbose, "### Using ", randomization_count, " randomizations...")
    _message(verbose, "### Using ", n_genes_to_randomize, " genes to randomize...")

    # ---- input checks, in the reference's order (:89-111) ----------------------------------
    if x.dtype == object or not np.issubdtype(x.dtype, np.number):
        raise TypeError("'x' must be a numeric matrix or data.frame")
    if x.ndim != 2:
        raise TypeError("'x' must be a numeric matrix or data.frame")
    x = np.asarray(x, dtype=np.float64)
    n_expr_cols = expression.shape[1]
    if n_expr_cols != x.shape[0]:
        raise ValueError("The number of columns in 'expression' must be the same as the rows in 'x'")
    if not isinstance(grid_points, (int, float, np.integer, np.floating)) or isinstance(grid_points, bool):
        raise TypeError("The value of 'grid.points' must be a numeric")
    if grid_points >= n_expr_cols:
        raise ValueError("The number of grid points appears to be very high (higher than the number of cells). "
                         "You can set the number of grid points using the 'grid.points' parameter.")
    if weights_advanced_Q is not None:
        weights_advanced_Q = np.asarray(weights_advanced_Q)
        if not np.issubdtype(weights_advanced_Q.dtype, np.number):
            raise TypeError("'weights.advanced.Q' must either be NULL or a numeric vector")
        if weights_advanced_Q.size != x.shape[0]:
            raise ValueError("The length of 'weights.advanced.Q' must be the same as the number of rows in 'x'")
        weights_advanced_Q = weights_advanced_Q.astype(np.float64).ravel()
    if spline_method not in ("ns", "bs"):
        raise ValueError("'spline.method' should be either 'ns' or 'bs'")

    count_cells = n_expr_cols
    count_genes = expression.shape[0]
    if n_genes_to_randomize > count_genes:
        warnings.warn(f"Number of genes to randomize ({n_genes_to_randomize}) is higher than the number of genes in the data.")
        warnings.warn(f"Setting number of genes to randomize to {count_genes}")
        n_genes_to_randomize = count_genes
    if x.shape[0] < 50:
        warnings.warn(f"The number of cells seems very low ({x.shape[0]}). Check your input.")
    if count_genes < 100:
        warnings.warn(f"The number of genes seems very low ({count_genes}). Check your input.")
    if grid_points < 10:
        warnings.warn("The value of 'grid.points' appears to be very low (<10). You can set the number of grid "
                      "points using the 'grid.points' parameter.")
    if grid_points > count_cells / 10:
        warnings.warn("The value of 'grid.points' appears to be very high (> No. of cells / 10). You can set the "
                      "number of grid points using the 'grid.points' parameter.")

    x_scale_center = x_scale_scale = None
    if scale:
        _message(verbose, "### scaling input data...")
        x, x_scale_center, x_scale_scale = _scale(x)
    if dir_randomization is not None:
        os.makedirs(dir_randomization, exist_ok=True)

    # rng = rcompat.RRandom(seed): every draw below is R's, in R's order (grid -> randomizations gene-major -> folds of
    # the mean model -> folds of the sd model): the call replays set.seed(seed); haystack(...)
    r_stream = rng if hasattr(rng, "perm_source") and hasattr(rng, "folds") else None
    if r_stream is None:
        rng = np.random.default_rng(rng)
    elif perm_source is None:
        perm_source = r_stream.perm_source
    if ctx is None:
        ctx, own_ctx = DeviceContext(device), True
    try:
        if grid_coord is None:
            _message(verbose, "### deciding grid points...")
            # "seeding": the distance passes run on the device (hs_grid_seed_*), the draws stay on `rng`;
            # "centroid" is kmeans() on the host (Hartigan-Wong with rcompat.RRandom)
            grid_coord = _grid.get_grid_points(x, method=grid_method, grid_points=int(grid_points), rng=rng,
                                               ctx=ctx if grid_method == "seeding" else None)
            if grid_coord.shape[0] != grid_points:
                warnings.warn(f"The number of grid points was changed from {grid_points} to {grid_coord.shape[0]}")
                grid_points = grid_coord.shape[0]

        # ---- K1: density stage on the device (:168-186) -----------------------------------
        dens = ctx.density(x, grid_coord, w_q=weights_advanced_Q, want_densities=return_densities)

        # ---- K2 / K3: observed D_KL (:199-213) --------------------------------------------
        _message(verbose, "### calculating Kullback-Leibler divergences...")
        # The gene selection (:222-236) needs only the row statistics, so it can precede the scoring; that lets
        # the sparse path with device draws score the genes and the randomization batch in one call.
        coeffVar = expr_sd / expr_mean                                                   # :222
        if isinstance(row_stats, str):
            # statistics from the device: the order(coeffVar) of the selection is taken there as well
            ranks = selection_ranks(coeffVar.size, n_genes_to_randomize, selection_method_genes_to_randomize)
            genes_to_randomize = ctx.select_by_cv(stats_raw[0], stats_raw[1], ranks - 1)
        else:
            genes_to_randomize = select_genes_to_randomize(coeffVar, n_genes_to_randomize,
                                                           selection_method_genes_to_randomize)
        info_perm_seed = None
        device_draw = isinstance(perm_source, str)
        if device_draw and perm_source != "device":
            raise ValueError("perm_source must be a callable, None or 'device'")
        src = None if device_draw else (perm_source if perm_source is not None else _NumpyPermSource(rng))
        all_D_KL_randomized = np.full((n_genes_to_randomize, randomization_count), np.nan)

        if sparse_in and device_draw:
            perm_seed = int(r_stream.unif_rand() * 2 ** 53) if r_stream is not None else int(rng.integers(0, 2 ** 63))
            _message(verbose, "### performing randomizations...")
            D_KL_observed, rand = ctx.dkl_csr_randomized(expression.p, expression.j, expression.x, genes_to_randomize,
                                                         randomization_count, perm_seed)
            all_D_KL_randomized[:, :] = rand
            info_perm_seed = perm_seed
        elif sparse_in:
            D_KL_observed = ctx.dkl_csr(expression.p, expression.j, expression.x)
        else:
            D_KL_observed = ctx.dkl_dense(expression)
        if not (sparse_in and device_draw):
            _message(verbose, "### performing randomizations...")

        # ---- K4: randomized D_KL (:245-284); draws in the reference's order (gene-major) ---
        if sparse_in and device_draw:
            pass
        elif sparse_in:
            budget = 64 << 20          # ids per device call
            batch_vals, batch_inds, batch_rows, used = [], [], [], 0

            def flush():
                nonlocal batch_vals, batch_inds, batch_rows, used
                if batch_rows:
                    out = ctx.dkl_perm_csr(batch_vals, batch_inds, index_base=1)
                    all_D_KL_randomized[batch_rows, :] = out
                batch_vals, batch_inds, batch_rows, used = [], [], [], 0

            for i, g in enumerate(genes_to_randomize):
                row = _sparse.extract_row_dgRMatrix_as_sparse(expression, int(g) + 1)
                ids = src("sparse", int(g), count_cells, row["val"].size, randomization_count)   # (R, nnz)
                batch_vals.append(row["val"])
                batch_inds.append(np.asfortranarray(np.asarray(ids, dtype=np.int32).T))
                batch_rows.append(i)
                used += ids.size
                if used >= budget:
                    flush()
            flush()
        elif device_draw:
            perm_seed = int(r_stream.unif_rand() * 2 ** 53) if r_stream is not None else int(rng.integers(0, 2 ** 63))
            # the whole loop :250-265 in one call: only the selected rows cross the link
            all_D_KL_randomized[:, :] = ctx.dkl_perm_dense_seeded_batch(expression, genes_to_randomize,
                                                                        randomization_count, perm_seed)
            info_perm_seed = perm_seed
        else:
            for i, g in enumerate(genes_to_randomize):
                perms = src("dense", int(g), count_cells, count_cells, randomization_count)       # (R, N)
                all_D_KL_randomized[i, :] = ctx.dkl_perm_dense(expression[int(g)],
                                                               np.asfortranarray(np.asarray(perms, dtype=np.int32).T),
                                                               index_base=1)
    finally:
        if own_ctx:
            ctx.close()

    # ---- host: significance fit (:288-302) -------------------------------------------------
    _message(verbose, "### estimating p-values...")
    names_rand = [gene_names[int(g)] for g in genes_to_randomize] if gene_names is not None else None
    sets_mean, sets_sd = cv_folds if cv_folds is not None else (None, None)
    if r_stream is not None and cv_folds is None:
        sets_mean = r_stream.folds(n_genes_to_randomize)
        sets_sd = r_stream.folds(n_genes_to_randomize)
    pv = _fit.get_log_p_D_KL_continuous(D_KL_observed, all_D_KL_randomized, coeffVar, coeffVar[genes_to_randomize],
                                        spline_method=spline_method, sets_mean=sets_mean, sets_sd=sets_sd,
                                        rng=None if r_stream is not None else rng, names=names_rand, verbose=verbose)
    rand_info = pv["info"]
    rand_info["method"] = spline_method
    rand_info["genes_to_randomize"] = genes_to_randomize + 1          # 1-based, as R stores them
    rand_info["KLD_rand"] = all_D_KL_randomized
    if info_perm_seed is not None:
        rand_info["perm_seed"] = info_perm_seed      # device.perm_sample(perm_seed, i * R + r, N, nnz) reproduces a draw
    p_vals = pv["fitted"]
    p_adjs = p_vals + math.log10(p_vals.size) if p_vals.size else p_vals.copy()
    p_adjs = np.where(p_adjs > 0, 0.0, p_adjs)

    if dir_randomization is not None:
        _message(verbose, "### writing randomized Kullback-Leibler divergences to file...")
        np.savetxt(os.path.join(dir_randomization, "random.D_KL.csv"), all_D_KL_randomized, delimiter=",")

    grid_out = grid_coord * x_scale_scale + x_scale_center if scale else grid_coord      # :313-314
    _message(verbose, "### returning result...")
    import pandas as pd
    results = pd.DataFrame({"D_KL": D_KL_observed, "log.p.vals": p_vals, "log.p.adj": p_adjs},
                           index=gene_names if gene_names is not None else None)
    res = Haystack(results=results,
                   info=dict(method="continuous_highD", randomization=rand_info, grid_coordinates=grid_out,
                             coord_mean=x_scale_center, coord_std=x_scale_scale, densities=dens["densities"],
                             cv=coeffVar, bandwidth=dens["bandwidth"], Q=dens["Q"]))
    return res


# --------------------------------------------------------------------------------------------
# S3-style front door (R/s3.R:21-87)
# --------------------------------------------------------------------------------------------
def haystack(x, expression=None, *, coord=None, assay=None, slot="data", dims=None, cutoff=1, method=None,
             weights_advanced_Q=None, dir_randomization=None, scale=True, grid_points=100,
             grid_method="centroid", **kwargs):
    """haystack(x, ...): dispatch on the type of `x`.

      numpy matrix / pandas DataFrame of coordinates + `expression`   (haystack.matrix / .data.frame)
      AnnData-like object (has .obsm and .X / .layers)                (the Seurat / SingleCellExperiment
          methods' role: `coord` names the embedding, `assay` a layer, `dims` selects columns;
          AnnData stores cells x genes, so the matrix is transposed to the reference's genes x cells)
    """
    if hasattr(x, "obsm") and hasattr(x, "X"):
        if coord is None:
            raise ValueError("Please specify an embedding. One of: " + ", ".join(map(str, x.obsm.keys())))
        mat = x.X if assay in (None, "X") else x.layers[assay]
        emb = np.asarray(x.obsm[coord], dtype=np.float64)
        if dims is not None:
            emb = emb[:, list(dims)]
        vn = getattr(x, "var_names", None)
        names = list(vn) if vn is not None and len(vn) else None
        import scipy.sparse as sp
        expr = _sparse.as_RsparseMatrix(sp.csr_matrix(mat).T.tocsr()) if sp.issparse(mat) else np.asarray(mat).T
        return haystack(emb, expr, weights_advanced_Q=weights_advanced_Q, dir_randomization=dir_randomization,
                        scale=scale, grid_points=grid_points, grid_method=grid_method,
                        gene_names=kwargs.pop("gene_names", names), **kwargs)
    if expression is None:
        raise TypeError("argument \"expression\" is missing, with no default")
    if hasattr(x, "values") and hasattr(x, "columns"):        # data.frame -> as.matrix
        x = x.values
    if dims is not None:
        x = np.asarray(x)[:, list(dims)]
    return haystack_continuous_highD(x, expression=expression, weights_advanced_Q=weights_advanced_Q,
                                     dir_randomization=dir_randomization, scale=scale, grid_points=grid_points,
                                     grid_method=grid_method, **kwargs)


def show_result_haystack(res_haystack, n=None, p_value_threshold=None, gene=None):
    """R/haystack.R:587-634: order by log.p.vals, then filter gene -> p.value.threshold -> head(n)."""
    result = res_haystack.get("results") if isinstance(res_haystack, dict) else None
    if result is None:
        raise ValueError("Results seem to be missing from 'haystack' object. Is 'res.haystack' a valid 'haystack' result?")
    if n is not None and (isinstance(n, bool) or not isinstance(n, (int, float, np.integer, np.floating))):
        raise TypeError("The value of 'n' should be an integer")
    if n is not None and n > len(result):
        warnings.warn("Integer value of 'n' is larger than the number of rows in the 'haystack' results. "
                      "Will return all results sorted.")
    if p_value_threshold is not None and (p_value_threshold < 0 or p_value_threshold > 1):
        raise ValueError("If 'p.value.threshold' is given as input, it should be between 0 and 1")
    result = result.iloc[np.argsort(result["log.p.vals"].values, kind="stable")]
    if gene is not None:
        gene = [gene] if isinstance(gene, str) else list(gene)
        result = result[result.index.isin(gene)]
    if p_value_threshold is not None:
        with np.errstate(divide="ignore"):
            result = result[result["log.p.vals"] <= np.log10(p_value_threshold)]
    if n is not None:
        result = result.head(int(n))
    return result

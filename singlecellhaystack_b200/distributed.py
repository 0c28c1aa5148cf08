"""Multi-GPU host layer: one process per GPU (torch.distributed), genes sharded across ranks.

The path shards naturally: every gene's D_KL depends only on its own row plus the
shared density state, so there is exactly one exchange step on each side of the scoring --
  the density stage is sharded over the CELLS (DeviceShard, density_mode="cells": every rank computes its cells'
  distances and densities, three in-place NCCL all-gathers over NVLink complete the row minima, the density matrix and
  Q on every rank -- bit-identical to the single-GPU stage; "broadcast": rank 0 computes and broadcasts),
  every rank scores a contiguous, nnz-balanced gene slice and its share of the (gene, 25-permutation) units of the
  randomization (dealt longest first; a row's stream does not depend on the rank count), and D_KL / the randomized
  matrix are all-gathered.
No collective sits inside a kernel.  Everything here is SPMD: all ranks call with the same host inputs
(each rank only uploads its own slice) and all ranks return the same result object.

The compute calls go through a small backend protocol (density / density_broadcast / dkl_csr /
dkl_dense / dkl_perm_csr / dkl_perm_dense) implemented by `DeviceContext`; the CPU tests plug in a
checker-backed stand-in to exercise the sharding and gathering logic over gloo.
"""
from __future__ import annotations

import math
import warnings

import numpy as np

from . import haystack as _hs
from . import sparse as _sparse
from . import splinefit as _fit

__all__ = ["plan_gene_shards", "owner_of", "deal_permutation_units", "DeviceShard", "haystack_continuous_highD_sharded"]


def plan_gene_shards(work: np.ndarray, world: int) -> np.ndarray:
    """Contiguous gene ranges with (nearly) equal work.  `work[g]` is the cost of gene g (stored
    entries for a dgRMatrix; 1 per gene for a dense matrix).  Returns `world + 1` boundaries;
    rank r owns genes [b[r], b[r+1])."""
    work = np.asarray(work, dtype=np.float64)
    m = work.size
    if world <= 1:
        return np.array([0, m], dtype=np.int64)
    cum = np.concatenate([[0.0], np.cumsum(work + 1e-9)])
    targets = cum[-1] * np.arange(1, world) / world
    inner = np.searchsorted(cum, targets, side="left")
    b = np.concatenate([[0], inner, [m]]).astype(np.int64)
    return np.maximum.accumulate(np.minimum(b, m))


def owner_of(genes: np.ndarray, bounds: np.ndarray) -> np.ndarray:
    """Rank owning each gene index under `bounds`."""
    return (np.searchsorted(bounds, np.asarray(genes), side="right") - 1).astype(np.int64)


def deal_permutation_units(costs, n_perm: int, world: int, block: int = 25):
    """Split the randomization batch (R/haystack_continuous.R:245-284: S selected genes x R permutations) into
    (gene position s, permutations [r_lo, r_hi)) units of `block` permutations and deal them to `world` ranks by
    descending cost (longest processing time first), so that the fixed total of S x R permuted rows is balanced
    however the selected genes are spread over the gene slices.  `costs[s]` = stored entries of selected gene s.
    Returns one list of (s, r_lo, r_hi) per rank; deterministic, identical on every rank."""
    costs = np.asarray(costs, dtype=np.float64)
    block = max(1, min(int(block), int(n_perm)))
    units = []
    for s, c in enumerate(costs):
        for r_lo in range(0, int(n_perm), block):
            r_hi = min(int(n_perm), r_lo + block)
            units.append((float(c) * (r_hi - r_lo), s, r_lo, r_hi))
    units.sort(key=lambda u: (-u[0], u[1], u[2]))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for cost, s, r_lo, r_hi in units:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += cost
        out[r].append((s, r_lo, r_hi))
    return out


class DeviceShard:
    """One rank's part of the device-resident path, SPMD over the ranks of `group` (one process per GPU).

    The three spans of the path and their exchange steps:
      density()    K1.  "cells" (default with several ranks): every rank computes 1 / world of the cells and three
                   in-place all-gathers rebuild the full state on every rank, bit-identical to one GPU
                   (DeviceContext.density_sharded_dev); "broadcast": rank 0 computes, NCCL broadcasts;
                   "replicate": every rank computes everything.
      score()      K3 on this rank's gene slice (plan_gene_shards: contiguous, nnz-balanced).
      randomize()  K4 on this rank's (gene, permutation block) units (deal_permutation_units), ONE batched call;
                   unit u draws from the streams s * R + r of the whole job, so the randomized matrix does not
                   depend on the number of ranks.
      gather()     D_KL slices -> every rank (all-gather, padded to the longest slice).
    """

    def __init__(self, ctx, group=None, density_mode: str = "cells"):
        import torch.distributed as dist
        self.ctx = ctx
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if density_mode not in ("cells", "broadcast", "replicate"):
            raise ValueError("density_mode must be 'cells', 'broadcast' or 'replicate'")
        self.density_mode = density_mode

    def density(self, x_dev, grid_dev, w_q_dev=None):
        if self.world == 1 or self.density_mode == "replicate":
            return self.ctx.density_dev(x_dev, grid_dev, w_q_dev)
        if self.density_mode == "cells":
            return self.ctx.density_sharded_dev(x_dev, grid_dev, w_q_dev, group=self.group)
        out = self.ctx.density_dev(x_dev, grid_dev, w_q_dev) if self.rank == 0 else None
        self.ctx.density_broadcast(x_dev.shape[1], grid_dev.shape[1], src=0, group=self.group)
        return out

    def score(self, p_dev, j_dev, x_dev, out_dev):
        self.ctx.dkl_csr_dev(p_dev, j_dev, x_dev, out_dev)

    def randomize(self, gene_ptr_dev, val_dev, n_perm_unit: int, seed: int, unit_stream_dev, out_dev):
        if gene_ptr_dev.numel() > 1:
            self.ctx.dkl_perm_csr_seeded_units_dev(gene_ptr_dev, val_dev, n_perm_unit, seed, unit_stream_dev, out_dev)

    def gather(self, local_dev, bounds):
        """All-gather the ranks' D_KL slices (genes [bounds[r], bounds[r+1]) on rank r) into one device vector."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return local_dev
        sizes = [int(bounds[r + 1] - bounds[r]) for r in range(self.world)]
        pad = max(sizes)
        send = torch.zeros(pad, dtype=local_dev.dtype, device=local_dev.device)
        send[:sizes[self.rank]] = local_dev[:sizes[self.rank]]
        recv = torch.empty(pad * self.world, dtype=local_dev.dtype, device=local_dev.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        return torch.cat([recv[r * pad:r * pad + sizes[r]] for r in range(self.world)])


def _all_gather_rows(local: np.ndarray, group):
    import torch.distributed as dist
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    return parts


class _GeneStreams:
    """sample() stand-in whose draws depend on (seed, gene) only, not on which rank owns the gene."""

    def __init__(self, seed):
        self.seed = int(seed) & 0x7fffffff

    def __call__(self, kind, gene, n, size, count):
        rng = np.random.default_rng([self.seed, int(gene)])
        out = np.empty((count, size), dtype=np.int32)
        for r in range(count):
            if kind == "dense" or size == n:
                out[r] = rng.permutation(n)[:size] + 1
            else:
                out[r] = rng.choice(n, size=size, replace=False, shuffle=True) + 1
        return out


def haystack_continuous_highD_sharded(x, expression, backend, grid_coord, *, group=None, scale=True,
                                      weights_advanced_Q=None, randomization_count=100,
                                      n_genes_to_randomize=100,
                                      selection_method_genes_to_randomize="heavytails", spline_method="ns",
                                      gene_names=None, seed=0, perm_source=None, cv_folds=None, row_stats=None,
                                      src_rank=0):
    """Gene-sharded haystack_continuous_highD (R/haystack_continuous.R:27-339) with a supplied grid.

    backend: this rank's DeviceContext (or an object with the same compute methods).
    perm_source: callable / None as in haystack_continuous_highD, or "device": the draws are made by the
    backend from `seed` (hs_dkl_perm_csr_seeded / hs_dkl_perm_dense_seeded).
    Returns the same Haystack object on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0

    sparse_in = _sparse.is_sparse(expression)
    if sparse_in:
        expression = _sparse.as_RsparseMatrix(expression)
    else:
        expression = np.asarray(expression, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    if expression.shape[1] != x.shape[0]:
        raise ValueError("The number of columns in 'expression' must be the same as the rows in 'x'")
    if row_stats is not None:
        expr_mean, expr_sd = (np.asarray(a, dtype=np.float64).copy() for a in row_stats)
    else:
        expr_mean, expr_sd = _hs.row_mean_sd(expression)
    if expr_mean.size and np.nanmin(expr_mean) < 0:
        raise ValueError("Some features have an average signal < 0. Expect average signal >= 0.")
    keep = ~(expr_sd == 0)
    if not keep.all():
        expression = expression.row_subset(keep) if sparse_in else expression[keep]
        if gene_names is not None:
            gene_names = [g for g, k in zip(gene_names, keep) if k]
    expr_mean = expr_mean[keep] + 1e-300
    expr_sd = expr_sd[keep] + 1e-300
    m, n = expression.shape
    n_sel = min(int(n_genes_to_randomize), m)

    center = sc = None
    if scale:
        x, center, sc = _hs.scale(x)
    grid_coord = np.asarray(grid_coord, dtype=np.float64)

    # ---- exchange step 1: density state from src_rank to everyone ----
    dens = None
    if rank == src_rank or world == 1:
        dens = backend.density(x, grid_coord, w_q=weights_advanced_Q)
    if world > 1:
        backend.density_broadcast(n, grid_coord.shape[0], src=src_rank, group=group)

    # ---- this rank's genes ----
    work = np.diff(expression.p).astype(np.float64) if sparse_in else np.ones(m)
    bounds = plan_gene_shards(work, world)
    g0, g1 = int(bounds[rank]), int(bounds[rank + 1])
    if sparse_in:
        p = expression.p.astype(np.int64)
        local = backend.dkl_csr(p[g0:g1 + 1] - p[g0], expression.j[p[g0]:p[g1]], expression.x[p[g0]:p[g1]])
    else:
        local = backend.dkl_dense(expression[g0:g1]) if g1 > g0 else np.zeros(0)
    # ---- exchange step 2: gather D_KL ----
    if world > 1:
        d_kl = np.concatenate(_all_gather_rows(np.asarray(local, dtype=np.float64), group))
    else:
        d_kl = np.asarray(local, dtype=np.float64)

    coeff_var = expr_sd / expr_mean
    genes = _hs.select_genes_to_randomize(coeff_var, n_sel, selection_method_genes_to_randomize)
    device_draw = isinstance(perm_source, str)
    if device_draw and perm_source != "device":
        raise ValueError("perm_source must be a callable, None or 'device'")
    src = None if device_draw else (perm_source if perm_source is not None else _GeneStreams(seed))
    owners = owner_of(genes, bounds)
    mine = np.nonzero(owners == rank)[0]
    rows_local = np.zeros((mine.size, randomization_count))
    if mine.size and device_draw:
        # draws made on the device; the stream of (gene, r) is its position in `genes` * R + r, whoever owns
        # the gene, so the result does not depend on the number of ranks
        for k, i in enumerate(mine):
            base = int(i) * int(randomization_count)
            if sparse_in:
                row = _sparse.extract_row_dgRMatrix_as_sparse(expression, int(genes[i]) + 1)
                rows_local[k] = np.asarray(backend.dkl_perm_csr_seeded([row["val"]], randomization_count, seed,
                                                                       stream_base=base))[0]
            else:
                rows_local[k] = backend.dkl_perm_dense_seeded(expression[int(genes[i])], randomization_count, seed,
                                                              stream_base=base)
    elif mine.size:
        if sparse_in:
            vals, inds = [], []
            for i in mine:
                row = _sparse.extract_row_dgRMatrix_as_sparse(expression, int(genes[i]) + 1)
                ids = src("sparse", int(genes[i]), n, row["val"].size, randomization_count)
                vals.append(row["val"])
                inds.append(np.asfortranarray(np.asarray(ids, dtype=np.int32).T))
            rows_local = np.asarray(backend.dkl_perm_csr(vals, inds, index_base=1))
        else:
            for k, i in enumerate(mine):
                perms = src("dense", int(genes[i]), n, n, randomization_count)
                rows_local[k] = backend.dkl_perm_dense(expression[int(genes[i])],
                                                       np.asfortranarray(np.asarray(perms, dtype=np.int32).T),
                                                       index_base=1)
    kld_rand = np.full((n_sel, randomization_count), np.nan)
    if world > 1:
        for idx, rows in _all_gather_rows((mine, rows_local), group):
            kld_rand[idx] = rows
    else:
        kld_rand[mine] = rows_local

    sets_mean, sets_sd = cv_folds if cv_folds is not None else (None, None)
    pv = _fit.get_log_p_D_KL_continuous(d_kl, kld_rand, coeff_var, coeff_var[genes], spline_method=spline_method,
                                        sets_mean=sets_mean, sets_sd=sets_sd, rng=np.random.default_rng(seed))
    p_vals = pv["fitted"]
    p_adjs = np.where(p_vals + math.log10(max(p_vals.size, 1)) > 0, 0.0, p_vals + math.log10(max(p_vals.size, 1)))
    import pandas as pd
    info = pv["info"]
    info.update(method=spline_method, genes_to_randomize=genes + 1, KLD_rand=kld_rand)
    results = pd.DataFrame({"D_KL": d_kl, "log.p.vals": p_vals, "log.p.adj": p_adjs},
                           index=gene_names if gene_names is not None else None)
    grid_out = grid_coord * sc + center if scale else grid_coord
    return _hs.Haystack(results=results,
                        info=dict(method="continuous_highD", randomization=info, grid_coordinates=grid_out,
                                  coord_mean=center, coord_std=sc, cv=coeff_var, shard_bounds=bounds,
                                  bandwidth=None if dens is None else dens["bandwidth"]))

"""Deterministic synthetic inputs for tests and bench.py.

Everything is a pure function of (seed, gene, cell, ...) through a splitmix64-style counter
hash written once over int64 arithmetic, so the same code runs on numpy arrays (CPU, small
cases, the oracle's inputs) and on torch CUDA tensors (bench-size inputs generated directly in
HBM) and yields identical values.  No library RNG stream is involved, so fixtures do not depend
on numpy/torch versions.

This module creates INPUTS only (embedding, grid, expression, stand-ins for R's sample()); it
contains no part of the scored path.
"""
from __future__ import annotations

import math

import numpy as np

BASE_SEED = 20261019

_M1 = np.int64(np.uint64(0xBF58476D1CE4E5B9).astype(np.int64))
_M2 = np.int64(np.uint64(0x94D049BB133111EB).astype(np.int64))
_GOLD = np.int64(np.uint64(0x9E3779B97F4A7C15).astype(np.int64))


def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


def _shr(z, k: int):
    """Logical right shift on int64 (torch only has the arithmetic one)."""
    return (z >> k) & ((1 << (64 - k)) - 1)


def _mix(z):
    """splitmix64 finaliser on int64 arrays/tensors with wrap-around arithmetic."""
    m1, m2 = int(_M1), int(_M2)
    z = (z ^ _shr(z, 30)) * m1
    z = (z ^ _shr(z, 27)) * m2
    return z ^ _shr(z, 31)


def hash3(seed: int, a, b):
    """64-bit hash of (seed, a, b); a and b are int64 arrays/tensors (broadcastable) or ints."""
    with np.errstate(over="ignore"):
        s = int(np.int64(seed) * _GOLD) if not isinstance(seed, np.ndarray) else seed
        return _mix(_mix(a * int(_GOLD) + s) + b)


def u24(h):
    """Top 24 bits as a non-negative integer in [0, 2^24)."""
    return _shr(h, 40)


# --------------------------------------------------------------------------------------------
# stand-ins for R's sample()
# --------------------------------------------------------------------------------------------
def sample_int(seed: int, stream: int, n: int, size: int | None = None) -> np.ndarray:
    """`size` distinct ids from 1..n in pseudo-random order (size=None: a full permutation).
    1-based like sample.int().  O(n log n): for tests and oracle-sized inputs."""
    with np.errstate(over="ignore"):
        keys = hash3(seed, np.int64(stream), np.arange(n, dtype=np.int64))
    order = np.argsort(keys.view(np.uint64), kind="stable")
    if size is not None:
        order = order[:size]
    return (order + 1).astype(np.int32)


def perm_source(seed: int):
    """Callable with the signature the oracle's `haystack_continuous_highD` expects:
    (kind, gene_index0, n, size, count) -> (count, size) 1-based ids."""
    def f(kind, gene, n, size, count):
        out = np.empty((count, size), dtype=np.int32)
        for r in range(count):
            out[r] = sample_int(seed, gene * 100003 + r, n, None if kind == "dense" else size)[:size]
        return out
    return f


def cv_folds(seed: int, stream: int, n: int, n_cv: int = 10) -> np.ndarray:
    """Stand-in for sample(rep(1:10, length.out=n)) (R/haystack_continuous.R:437)."""
    base = (np.arange(n) % n_cv) + 1
    return base[sample_int(seed, 7919 + stream, n) - 1]


def feistel_perm(seed: int, stream, idx, n: int):
    """Bijection of [0, n) applied to idx: a 4-round Feistel network on an even number of bits
    >= log2(n) with cycle walking (re-encrypt values that fall outside the domain).  stream/idx:
    int64 numpy arrays or torch tensors.  For distinct idx it yields distinct pseudo-random cell
    ids without materialising a permutation -- the bench-size stand-in for sample(n, size)."""
    bits = max(2, math.ceil(math.log2(max(n, 2))))
    if bits % 2:
        bits += 1
    half = bits // 2
    mask = (1 << half) - 1
    if _is_torch(idx):
        import torch
        where = torch.where
    else:
        where = np.where

    def rounds(x):
        left, right = _shr(x, half) & mask, x & mask
        for rnd in range(4):
            f = u24(hash3(seed + 1000 * (rnd + 1), stream, right)) & mask
            left, right = right, left ^ f
        return (left << half) | right

    x = rounds(idx)
    for _ in range(64):  # domain < 4n, so the expected number of walks is < 4
        bad = x >= n
        if not bool(bad.any()):
            break
        x = where(bad, rounds(x), x)
    return x


# --------------------------------------------------------------------------------------------
# embedding and grid
# --------------------------------------------------------------------------------------------
def make_embedding(n_cells: int, n_dims: int, seed: int = BASE_SEED, n_components: int = 32):
    """Mixture of `n_components` isotropic Gaussians with PC-like per-dimension decay.
    Returns (x (N,d) float64 C-order, labels (N,) int32).  Host-side (numpy Generator seeded
    from `seed`; ~2 s at 1M x 50)."""
    rng = np.random.default_rng(seed)
    decay = 1.0 / np.sqrt(1.0 + np.arange(n_dims))
    means = rng.normal(0.0, 4.0, size=(n_components, n_dims)) * decay
    labels = rng.integers(0, n_components, size=n_cells).astype(np.int32)
    x = means[labels] + rng.normal(0.0, 1.0, size=(n_cells, n_dims)) * decay
    return x, labels


def scale_coords(x: np.ndarray):
    """R's scale(): centre by column mean, divide by column sd (n-1)."""
    x = np.asarray(x, dtype=np.float64)
    center = x.mean(axis=0)
    xc = x - center
    sc = np.sqrt((xc * xc).sum(axis=0) / (x.shape[0] - 1))
    return xc / sc, center, sc


def make_grid(x_scaled: np.ndarray, n_grid: int, seed: int = BASE_SEED, subsample: int = 10000,
              lloyd_steps: int = 3) -> np.ndarray:
    """Deterministic grid: k-means++-style seeding on a subsample, then a few Lloyd steps.
    The grid is an INPUT of the scored path (R/haystack_continuous.R:30,157)."""
    n = x_scaled.shape[0]
    if n > subsample:
        pick = sample_int(seed, 11, n, subsample) - 1
        xs = x_scaled[np.sort(pick)]
    else:
        xs = x_scaled
    m = xs.shape[0]
    first = int(sample_int(seed, 12, m, 1)[0] - 1)
    centers = [xs[first]]
    d2 = ((xs - centers[0]) ** 2).sum(axis=1)
    for i in range(1, n_grid):
        # farthest-point flavour of the seeding: deterministic and well spread
        u = (u24(hash3(seed, np.int64(13 + i), np.arange(m, dtype=np.int64))).astype(np.float64) + 0.5) / (1 << 24)
        score = d2 * (0.5 + u)  # randomised farthest point
        nxt = int(np.argmax(score))
        centers.append(xs[nxt])
        d2 = np.minimum(d2, ((xs - xs[nxt]) ** 2).sum(axis=1))
    c = np.stack(centers)
    for _ in range(lloyd_steps):
        dist = ((xs[:, None, :] - c[None, :, :]) ** 2).sum(axis=2) if m * n_grid * xs.shape[1] < 5e7 else None
        if dist is None:
            dist = (xs * xs).sum(1)[:, None] - 2 * xs @ c.T + (c * c).sum(1)[None, :]
        lab = dist.argmin(axis=1)
        for k in range(n_grid):
            sel = lab == k
            if sel.any():
                c[k] = xs[sel].mean(axis=0)
    return np.ascontiguousarray(c)


# --------------------------------------------------------------------------------------------
# sparse expression
# --------------------------------------------------------------------------------------------
_VAL_TABLE = np.log1p(1.0 + np.arange(8, dtype=np.float64))  # log1p(1 + k), k = 0..7
# cumulative 8-bit thresholds of a Poisson(2) truncated at 7
_POIS_CUM = np.cumsum([math.exp(-2) * 2 ** k / math.factorial(k) for k in range(8)])
_POIS_CUM = np.minimum(255, np.floor(_POIS_CUM / _POIS_CUM[-1] * 256)).astype(np.int64)


def gene_densities(n_genes: int, mean_density: float, seed: int = BASE_SEED) -> np.ndarray:
    """rho_g log-uniform in [1e-4, 0.3], rescaled so that mean(rho) == mean_density, capped at 0.9."""
    u = (u24(hash3(seed, np.int64(21), np.arange(n_genes, dtype=np.int64))).astype(np.float64) + 0.5) / (1 << 24)
    rho = np.exp(np.log(1e-4) + u * (np.log(0.3) - np.log(1e-4)))
    for _ in range(8):
        rho = np.minimum(0.9, rho * (mean_density / rho.mean()))
    return rho


def gene_plant(n_genes: int, n_components: int, seed: int = BASE_SEED, frac: float = 0.10):
    """10 % of genes are 'planted': enriched 6x in one mixture component. Returns component id or -1."""
    h = hash3(seed, np.int64(22), np.arange(n_genes, dtype=np.int64))
    planted = (u24(h) % 1000) < int(frac * 1000)
    comp = (_shr(h, 8) % n_components).astype(np.int64)
    return np.where(planted, comp, -1)


def expression_block(g0: int, g1: int, n_cells: int, rho, plant, labels, seed: int = BASE_SEED):
    """Rows g0..g1 of the synthetic expression matrix as CSR pieces.

    rho/plant: per-gene arrays (numpy) for ALL genes; labels: per-cell component ids as numpy
    array or torch tensor -- its type selects the backend.  Returns (row_counts (g1-g0,),
    cols int32, vals float32/float64) with cols ascending within each row.
    """
    torch_mode = _is_torch(labels)
    ng = g1 - g0
    n_comp_boost = 6.0
    rho_b = np.asarray(rho[g0:g1], dtype=np.float64)
    pl_b = np.asarray(plant[g0:g1], dtype=np.int64)
    # thresholds on a 24-bit uniform: base rate outside the planted component, boosted inside
    # (approximately rho overall when the component holds ~1/32 of the cells)
    base = np.where(pl_b >= 0, rho_b / (1.0 + (n_comp_boost - 1.0) / 32.0), rho_b)
    thr_out = np.minimum((1 << 24) - 1, np.floor(base * (1 << 24))).astype(np.int64)
    thr_in = np.minimum((1 << 24) - 1, np.floor(np.minimum(0.95, base * n_comp_boost) * (1 << 24))).astype(np.int64)
    if torch_mode:
        import torch
        dev = labels.device
        genes = torch.arange(g0, g1, device=dev, dtype=torch.int64)[:, None]
        cells = torch.arange(n_cells, device=dev, dtype=torch.int64)[None, :]
        h = hash3(seed, genes, cells)
        t_out = torch.as_tensor(thr_out, device=dev)[:, None]
        t_in = torch.as_tensor(thr_in, device=dev)[:, None]
        pl = torch.as_tensor(pl_b, device=dev)[:, None]
        thr = torch.where(labels.to(torch.int64)[None, :] == pl, t_in, t_out)
        mask = u24(h) < thr
        counts = mask.sum(dim=1)
        rows, cols = torch.nonzero(mask, as_tuple=True)
        hv = h[rows, cols]
        k = (torch.bucketize(_shr(hv, 8) & 255, torch.as_tensor(_POIS_CUM, device=dev), right=True)).clamp_(max=7)
        vals = torch.as_tensor(_VAL_TABLE, device=dev)[k].to(torch.float32)
        return counts, cols.to(torch.int32), vals
    genes = np.arange(g0, g1, dtype=np.int64)[:, None]
    cells = np.arange(n_cells, dtype=np.int64)[None, :]
    with np.errstate(over="ignore"):
        h = hash3(seed, genes, cells)
    thr = np.where(np.asarray(labels, dtype=np.int64)[None, :] == pl_b[:, None], thr_in[:, None], thr_out[:, None])
    mask = u24(h) < thr
    counts = mask.sum(axis=1)
    rows, cols = np.nonzero(mask)
    hv = h[rows, cols]
    k = np.minimum(7, np.searchsorted(_POIS_CUM, _shr(hv, 8) & 255, side="right"))
    vals = _VAL_TABLE[k]
    return counts.astype(np.int64), cols.astype(np.int32), vals


def make_csr_numpy(n_genes: int, n_cells: int, mean_density: float, labels: np.ndarray,
                   seed: int = BASE_SEED, block: int = 64):
    """Whole synthetic matrix on the host as dgRMatrix slots (p int64, j int32, x float64)."""
    rho = gene_densities(n_genes, mean_density, seed)
    plant = gene_plant(n_genes, int(labels.max()) + 1 if labels.size else 1, seed)
    ps, js, xs = [np.zeros(1, dtype=np.int64)], [], []
    for g0 in range(0, n_genes, block):
        g1 = min(n_genes, g0 + block)
        counts, cols, vals = expression_block(g0, g1, n_cells, rho, plant, labels, seed)
        ps.append(counts)
        js.append(cols)
        xs.append(vals)
    p = np.cumsum(np.concatenate(ps)).astype(np.int64)
    return p, np.concatenate(js).astype(np.int32), np.concatenate(xs).astype(np.float64)

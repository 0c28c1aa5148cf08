"""Grid-point selection, the step before the scored path (R/haystack_highD.R:72-134).

This stays on the HOST in the drop-in design: it draws random numbers, and
under R it is R's own `kmeans()` / `sample()` that run, so `set.seed()` reproducibility is R's.
This Python mirror exists so the package is usable without R.  With `rng=rcompat.RRandom(seed)` it is R's
own procedure -- kmeans() as Hartigan-Wong with nstart restarts on R's Mersenne-Twister stream, sample() for the
seeding method -- and returns the grid an R session with set.seed(seed) gets.  With a numpy seed / Generator the
random stream is numpy's and "centroid" is Lloyd's algorithm: same distribution, different numbers.
"""
from __future__ import annotations

import warnings

import numpy as np

__all__ = ["get_grid_points", "get_dist_two_sets", "get_euclidean_distance"]


def get_euclidean_distance(x, y) -> float:
    """R/haystack_highD.R:7-9."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return float(np.sqrt(((x - y) ** 2).sum()))


def get_dist_two_sets(set1, set2) -> np.ndarray:
    """R/haystack_highD.R:17-24 (host helper used by the seeding method only; the scored path
    computes its distances on the device)."""
    set1 = np.asarray(set1, dtype=np.float64)
    set2 = np.asarray(set2, dtype=np.float64)
    out = np.empty((set1.shape[0], set2.shape[0]))
    for s2 in range(set2.shape[0]):
        d = set1 - set2[s2]
        out[:, s2] = np.sqrt((d * d).sum(axis=1))
    return out


def _kmeans_lloyd(x: np.ndarray, k: int, iter_max: int, nstart: int, rng: np.random.Generator) -> np.ndarray:
    """kmeans(x, centers=k, iter.max, nstart)$centers with Lloyd iterations: each start picks k
    distinct rows as initial centres (as R does), the start with the lowest within-SS wins."""
    n = x.shape[0]
    x2 = (x * x).sum(axis=1)
    best, best_ss = None, np.inf
    for _ in range(nstart):
        c = x[rng.choice(n, size=k, replace=False)].copy()
        for _it in range(iter_max):
            d = x2[:, None] - 2.0 * (x @ c.T) + (c * c).sum(axis=1)[None, :]
            lab = d.argmin(axis=1)
            newc = c.copy()
            sums = np.zeros_like(c)
            np.add.at(sums, lab, x)
            cnt = np.bincount(lab, minlength=k)
            nz = cnt > 0
            newc[nz] = sums[nz] / cnt[nz][:, None]
            if np.array_equal(newc, c):
                break
            c = newc
        d = x2[:, None] - 2.0 * (x @ c.T) + (c * c).sum(axis=1)[None, :]
        ss = float(np.maximum(d.min(axis=1), 0).sum())
        if ss < best_ss:
            best, best_ss = c, ss
    return best


def get_grid_points(input, method: str = "centroid", grid_points: int = 100, rng=None, ctx=None) -> np.ndarray:
    """R/haystack_highD.R:72-134.  Returns (grid.points, d) coordinates.  With a device context `ctx` the distance
    passes of the "seeding" method run on the GPU (hs_grid_seed_*); the draws stay here, on `rng`."""
    x = np.asarray(input, dtype=np.float64)
    r_stream = rng if hasattr(rng, "kmeans") and hasattr(rng, "sample_prob_one") else None     # rcompat.RRandom
    if r_stream is None:
        rng = np.random.default_rng(rng)
    if x.shape[0] < grid_points:                                                      # :74-77
        warnings.warn(f"Fewer input points than grid.points. Adjusting value of grid.points to {x.shape[0]}")
        grid_points = x.shape[0]
    if method == "centroid":                                                          # :79-85
        if r_stream is not None:           # R's own kmeans(): Hartigan-Wong on R's random stream
            return r_stream.kmeans(x, int(grid_points), iter_max=10, nstart=10)
        return _kmeans_lloyd(x, int(grid_points), iter_max=10, nstart=10, rng=rng)
    if method == "seeding" and ctx is not None:                                       # :87-129, distances on the device
        n, d = x.shape
        w = np.array([2.0, 1.0, 1.0])                                                 # :124
        grid = np.full((grid_points, d), np.nan)
        ctx.grid_seed_begin(x)
        index = int(r_stream.sample_int(n, 1)[0]) - 1 if r_stream is not None else int(rng.integers(n))   # :95
        for i in range(grid_points):
            probs, o = ctx.grid_seed_step(index, want_probs=i + 1 < grid_points)
            ww = w[:o.size]
            grid[i] = (x[o] * ww[:, None]).sum(axis=0) / ww.sum()                     # :127 weighted.mean of order(tmp)[1:3]
            if i + 1 < grid_points:                                                   # :107-110
                index = (r_stream.sample_prob_one(probs) - 1 if r_stream is not None
                         else int(rng.choice(n, p=probs / probs.sum())))
        return grid
    if method == "seeding":                                                           # :87-129
        n, d = x.shape
        grid = np.full((grid_points, d), np.nan)
        index = int(r_stream.sample_int(n, 1)[0]) - 1 if r_stream is not None else int(rng.integers(n))   # :95
        grid[0] = x[index]
        min_dist = get_dist_two_sets(x, grid[0:1])[:, 0]
        dist_to_grid = np.full((n, grid_points), np.nan)
        dist_to_grid[:, 0] = min_dist
        for i in range(1, grid_points):
            probs = min_dist ** 2                                                     # :107
            index = (r_stream.sample_prob_one(probs) - 1 if r_stream is not None
                     else int(rng.choice(n, p=probs / probs.sum())))                  # :110
            grid[i] = x[index]
            tmp = get_dist_two_sets(x, grid[i:i + 1])[:, 0]
            dist_to_grid[:, i] = tmp
            min_dist = np.minimum(min_dist, tmp)                                      # :116
        w = np.array([2.0, 1.0, 1.0])                                                 # :124
        for i in range(grid_points):
            o = np.argsort(dist_to_grid[:, i], kind="stable")[:3]
            ww = w[:o.size]
            grid[i] = (x[o] * ww[:, None]).sum(axis=0) / ww.sum()                     # :127 weighted.mean
        return grid
    raise ValueError("grid.method should be 'centroid' or 'seeding'")

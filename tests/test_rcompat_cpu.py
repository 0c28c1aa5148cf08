"""The product's R-compatible host helpers (hs_r_*, singlecellhaystack_b200/rcompat.py) against the oracle's separate
restatement (oracle/r_base.c) and against values any R >= 3.6 prints: the same Mersenne-Twister stream, the same
sample.int() draws, the same Hartigan-Wong k-means -- bit for bit -- and the grid of the reference's vignette call."""
from pathlib import Path

import numpy as np

from oracle import haystack_oracle as ho
from oracle import r_base as rb
from singlecellhaystack_b200 import grid as grid_mod
from singlecellhaystack_b200.rcompat import RRandom


def test_stream_and_draws_match_r_and_the_oracle():
    a = RRandom(1234)
    np.testing.assert_allclose([a.unif_rand() for _ in range(3)], [0.1137034, 0.6222994, 0.6092747], atol=5e-8)
    assert RRandom(1234).sample_int(10).tolist() == [10, 6, 5, 4, 1, 8, 2, 7, 9, 3]      # set.seed(1234); sample.int(10)
    assert RRandom(42).sample_int(10).tolist() == [1, 5, 10, 8, 2, 4, 6, 9, 7, 3]
    for seed, kind in ((7, "Rejection"), (123, "Rounding"), (2 ** 31 + 5, "Rejection")):
        p, o = RRandom(seed, kind), rb.RRng(seed, kind)
        assert [p.unif_rand() for _ in range(700)] == [o.unif_rand() for _ in range(700)]            # across a refill
        np.testing.assert_array_equal(p.sample_int(601, 601, count=4), o.sample_int(601, 601, count=4))
        np.testing.assert_array_equal(p.sample_int(100000, 137, count=3), o.sample_int(100000, 137, count=3))
        np.testing.assert_array_equal(p.sample_int(5, 1, count=6), o.sample_int(5, 1, count=6))       # k < 2: direct draws
        np.testing.assert_array_equal(p.folds(100), o.sample(np.resize(np.arange(1, 11), 100)))
        w = np.abs(np.sin(np.arange(1, 400))) + 0.01
        assert [p.sample_prob_one(w) for _ in range(20)] == [o.sample_prob_one(w) for _ in range(20)]     # Walker (> 200 large)
        w2 = np.array([5.0, 1.0, 1.0, 3.0, 0.0, 2.0])
        assert [p.sample_prob_one(w2) for _ in range(20)] == [o.sample_prob_one(w2) for _ in range(20)]   # sorted search


def test_kmeans_and_grid_match_the_oracle_and_the_committed_r_grid(toy):
    xs, _, _ = ho.scale_coords(toy["coord"].astype(np.float64))
    got = grid_mod.get_grid_points(xs, "centroid", 100, rng=RRandom(1234))
    np.testing.assert_array_equal(got, rb.get_grid_points(xs, rb.RRng(1234), "centroid", 100))
    z = np.load(Path(__file__).resolve().parent / "golden" / "toy_r1234.npz")
    np.testing.assert_array_equal(got, z["grid_scaled"])
    np.testing.assert_array_equal(grid_mod.get_grid_points(xs, "seeding", 40, rng=RRandom(9)),
                                  rb.get_grid_points(xs, rb.RRng(9), "seeding", 40))
    # duplicated rows: nstart = 1 falls back to the distinct rows, as kmeans() does
    x = np.repeat(np.arange(12.0).reshape(6, 2), 3, axis=0)
    c = RRandom(3).kmeans(x, 5, iter_max=10, nstart=1)
    o = rb.kmeans(x, 5, rb.RRng(3), iter_max=10, nstart=1)["centers"]
    np.testing.assert_array_equal(c, o)

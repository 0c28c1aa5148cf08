"""The keyed permutation behind hs_dkl_perm_csr_seeded: the library's host helper (csrc/hs_prp.cuh through
hs_perm_sample) against the numpy restatement in oracle/prp_oracle.py, bit for bit; no GPU involved."""
import numpy as np
import pytest

from oracle import prp_oracle as po


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 16, 17, 601, 1024, 1025, 4097, 65536, 65537, 1_000_003])
def test_library_matches_oracle(n):
    from singlecellhaystack_b200.device import perm_sample
    for seed, stream in [(0, 0), (20261019, 7), (2**63 + 5, 2**40 + 3), (2**64 - 1, 2**64 - 1)]:
        size = n if n <= 70000 else 50000
        got = perm_sample(seed, stream, n, size, index_base=1)
        want = po.prp_sample(seed, stream, n, size, index_base=1)
        np.testing.assert_array_equal(got, want)
        assert got.min() >= 1 and got.max() <= n
        assert np.unique(got).size == size            # distinct: a sample without replacement


def test_full_draw_is_a_permutation_and_inverts():
    n = 12345
    fwd = po.prp_forward(99, 3, n, np.arange(n))
    assert np.array_equal(np.sort(fwd), np.arange(n))
    assert np.array_equal(po.prp_inverse(99, 3, n, fwd), np.arange(n))


def test_streams_and_seeds_differ():
    a = po.prp_sample(1, 0, 100000, 1000)
    assert not np.array_equal(a, po.prp_sample(1, 1, 100000, 1000))
    assert not np.array_equal(a, po.prp_sample(2, 0, 100000, 1000))
    assert np.array_equal(a, po.prp_sample(1, 0, 100000, 1000))


def test_draw_is_uniform_enough():
    """Block occupancy of the drawn cells behaves like a hypergeometric draw (chi-square / df near its
    finite-population value 1 - size/n), as numpy's own sampler does."""
    n, size, blocks = 200000, 20000, 100
    chi = []
    for s in range(40):
        ids = po.prp_sample(5, s, n, size, index_base=0)
        cnt = np.bincount(ids // (n // blocks), minlength=blocks)
        e = size / blocks
        chi.append(((cnt - e) ** 2 / e).sum() / (blocks - 1))
    assert 0.75 < np.mean(chi) < 1.05


def test_bad_arguments():
    from singlecellhaystack_b200.device import perm_sample
    with pytest.raises(ValueError):
        perm_sample(1, 1, 10, 11)
    with pytest.raises(ValueError):
        perm_sample(1, 1, 0, 0)
    with pytest.raises(ValueError):
        po.prp_sample(1, 1, 10, 11)

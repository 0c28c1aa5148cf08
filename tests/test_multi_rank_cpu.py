"""world_size-2 (gloo, CPU) test of the gene-sharded host layer: partitioning, the two exchange steps and
the ownership of the randomization rows.  The compute backend is a stand-in built on the oracle (tests may
use the checker); on GPUs the same code runs with DeviceContext (tests/test_gpu_full_path.py covers the
single-rank case, bench.py --gpus N the NCCL broadcast)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_plan_gene_shards_balances_nnz():
    from singlecellhaystack_b200.distributed import owner_of, plan_gene_shards
    rng = np.random.default_rng(1)
    work = np.exp(rng.uniform(np.log(10), np.log(1e6), 30000))
    for world in (1, 2, 4, 8):
        b = plan_gene_shards(work, world)
        assert b[0] == 0 and b[-1] == work.size and (np.diff(b) >= 0).all() and b.size == world + 1
        loads = np.array([work[b[r]:b[r + 1]].sum() for r in range(world)])
        assert loads.max() / loads.mean() < 1.05          # heavy-tailed rows, still within 5 %
    assert plan_gene_shards(np.ones(3), 8)[-1] == 3       # more ranks than genes: empty slices allowed
    b = plan_gene_shards(np.ones(10), 4)
    assert owner_of(np.arange(10), b).tolist() == sorted(owner_of(np.arange(10), b).tolist())
    assert set(owner_of(np.arange(10), b)) <= set(range(4))


class _OracleBackend:
    """Checker-backed stand-in for DeviceContext (CPU test only)."""

    def __init__(self):
        self.state = None

    def density(self, x, grid, w_q=None, **kw):
        from oracle import haystack_oracle as ho
        self.state = ho.density_stage(x, grid, w_q)
        self._args = (x, grid, w_q)
        return dict(bandwidth=self.state["bandwidth"], Q=self.state["Q"], densities=None, nearest=None)

    def density_broadcast(self, n_cells, n_grid, src=0, group=None):
        import torch.distributed as dist
        box = [self.state if dist.get_rank(group) == src else None]
        dist.broadcast_object_list(box, src=src, group=group)
        self.state = box[0]
        assert self.state["density"].shape == (n_cells, n_grid)

    def dkl_csr(self, p, j, x):
        from oracle import haystack_oracle as ho
        sp = ho.RowSparse(p=np.asarray(p), j=np.asarray(j), x=np.asarray(x), shape=(len(p) - 1, self.state["density"].shape[0]))
        return ho.dkl_observed(sp, self.state["density"], self.state["Q"])

    def dkl_dense(self, e):
        from oracle import haystack_oracle as ho
        return ho.dkl_observed(np.asarray(e), self.state["density"], self.state["Q"])

    def dkl_perm_csr(self, vals, inds, index_base=1):
        from oracle import haystack_oracle as ho
        return ho.dkl_randomized_sparse(vals, [np.asarray(a).T for a in inds], self.state["density"], self.state["Q"])

    def dkl_perm_csr_seeded(self, vals, n_perm, seed, stream_base=0):
        from oracle import haystack_oracle as ho, prp_oracle as po
        n = self.state["density"].shape[0]
        inds = [np.stack([po.prp_sample(seed, stream_base + s * n_perm + r, n, len(v)) for r in range(n_perm)])
                for s, v in enumerate(vals)]
        return ho.dkl_randomized_sparse(vals, inds, self.state["density"], self.state["Q"])

    def dkl_perm_dense_seeded(self, v, n_perm, seed, stream_base=0):
        from oracle import haystack_oracle as ho, prp_oracle as po
        n = self.state["density"].shape[0]
        perms = np.stack([po.prp_sample(seed, stream_base + r, n, n) for r in range(n_perm)])
        return ho.dkl_randomized_dense(np.asarray(v)[None], perms[None], self.state["density"], self.state["Q"])[0]

    def dkl_perm_dense(self, v, perm, index_base=1):
        from oracle import haystack_oracle as ho
        return ho.dkl_randomized_dense(np.asarray(v)[None], np.asarray(perm).T[None], self.state["density"], self.state["Q"])[0]


def _worker(rank, world, port, form, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from singlecellhaystack_b200 import sparse as sp
    from singlecellhaystack_b200 import synth
    from singlecellhaystack_b200.distributed import haystack_continuous_highD_sharded
    from oracle import haystack_oracle as ho
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        toy = np.load(ROOT / "tests" / "golden" / "toy.npz")
        seed = int(toy["seed"])
        expr = toy["expression"].astype(np.float64)
        stats = ho.row_mean_sd(ho.RowSparse.from_dense(expr)) if form == "sparse" else ho.row_mean_sd(expr)
        e = sp.as_CsparseMatrix(expr) if form == "sparse" else expr
        res = haystack_continuous_highD_sharded(
            toy["coord"], e, _OracleBackend(), toy["grid100"], perm_source=synth.perm_source(seed), row_stats=stats,
            cv_folds=(synth.cv_folds(seed, 0, 100), synth.cv_folds(seed, 1, 100)), gene_names=list(toy["gene_names"]))
        out[rank] = (res["results"]["D_KL"].values.copy(), res["results"]["log.p.vals"].values.copy(),
                     res["info"]["randomization"]["KLD_rand"].copy(), res["info"]["shard_bounds"].copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("form", ["sparse", "dense"])
def test_two_rank_gloo_matches_single_process_golden(toy, form):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, form, out), nprocs=2, join=True)
    assert sorted(out.keys()) == [0, 1]
    d0, lp0, kr0, b0 = out[0]
    d1, lp1, kr1, b1 = out[1]
    np.testing.assert_array_equal(d0, d1)                 # every rank returns the same object
    np.testing.assert_array_equal(kr0, kr1)
    assert b0.tolist() == b1.tolist() and 0 < b0[1] < 500  # both ranks got genes
    np.testing.assert_allclose(d0, toy["dkl100"], rtol=1e-12)
    gold_rand = toy["full_KLD_rand"] if form == "sparse" else toy["fulld_KLD_rand"]
    gold_logp = toy["full_log_p"] if form == "sparse" else toy["fulld_log_p"]
    np.testing.assert_allclose(kr0, gold_rand, rtol=1e-12)
    assert np.max(np.abs(lp0 - gold_logp)) < 1e-6


def _worker_device_draw(rank, world, port, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from singlecellhaystack_b200 import sparse as sp
    from singlecellhaystack_b200 import synth
    from singlecellhaystack_b200.distributed import haystack_continuous_highD_sharded
    from oracle import haystack_oracle as ho
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        toy = np.load(ROOT / "tests" / "golden" / "toy.npz")
        seed = int(toy["seed"])
        expr = toy["expression"].astype(np.float64)
        stats = ho.row_mean_sd(ho.RowSparse.from_dense(expr))
        res = haystack_continuous_highD_sharded(
            toy["coord"], sp.as_CsparseMatrix(expr), _OracleBackend(), toy["grid100"], perm_source="device", seed=4711,
            row_stats=stats, n_genes_to_randomize=30, randomization_count=20,
            cv_folds=(synth.cv_folds(seed, 0, 30), synth.cv_folds(seed, 1, 30)))
        out[rank] = (res["info"]["randomization"]["KLD_rand"].copy(), res["results"]["log.p.vals"].values.copy())
    finally:
        dist.destroy_process_group()


def test_device_draws_do_not_depend_on_world_size(toy):
    """perm_source="device": the stream of a (gene, randomization) pair is fixed by its position in the
    selection, so two ranks reproduce the single-process result exactly."""
    import torch.multiprocessing as mp
    from oracle import haystack_oracle as ho
    from singlecellhaystack_b200 import sparse as sp
    from singlecellhaystack_b200 import synth
    from singlecellhaystack_b200.distributed import haystack_continuous_highD_sharded
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_worker_device_draw, args=(2, port, out), nprocs=2, join=True)
    seed = int(toy["seed"])
    expr = toy["expression"]
    single = haystack_continuous_highD_sharded(
        toy["coord"], sp.as_CsparseMatrix(expr), _OracleBackend(), toy["grid100"], perm_source="device", seed=4711,
        row_stats=ho.row_mean_sd(ho.RowSparse.from_dense(expr)), n_genes_to_randomize=30, randomization_count=20,
        cv_folds=(synth.cv_folds(seed, 0, 30), synth.cv_folds(seed, 1, 30)))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][0], single["info"]["randomization"]["KLD_rand"])
    np.testing.assert_allclose(out[0][1], single["results"]["log.p.vals"].values, rtol=0, atol=1e-12)


def test_deal_permutation_units_is_balanced_and_complete():
    from singlecellhaystack_b200.distributed import deal_permutation_units
    rng = np.random.default_rng(1)
    costs = rng.integers(50, 200000, size=100)
    for world in (1, 2, 4, 8):
        deal = deal_permutation_units(costs, 100, world, block=25)
        assert len(deal) == world
        seen = sorted(u for part in deal for u in part)
        assert seen == sorted((s, lo, lo + 25) for s in range(100) for lo in range(0, 100, 25))     # every unit once
        load = [sum(costs[s] * (hi - lo) for s, lo, hi in part) for part in deal]
        assert max(load) <= 1.02 * (sum(load) / world) + 25 * costs.max()                            # LPT bound
        assert deal == deal_permutation_units(costs, 100, world, block=25)                          # deterministic
    # ragged: R not a multiple of the block, fewer units than ranks
    deal = deal_permutation_units([10, 20], 7, 4, block=3)
    assert sorted(u for part in deal for u in part) == [(0, 0, 3), (0, 3, 6), (0, 6, 7), (1, 0, 3), (1, 3, 6), (1, 6, 7)]

#!/usr/bin/env python
"""bench.py -- genes/sec incl. randomizations of the continuous high-D scoring path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c1..c5] [--scaling strong|weak]

One *step* = one pass of the scored path of haystack() over one synthetic data set:
  K1  density stage            (R/haystack_continuous.R:168-186)
  K3  observed D_KL, all genes (R/haystack_continuous.R:199-213, sparse branch)
  K4  randomized D_KL, S genes x R permutations (R/haystack_continuous.R:245-284, sparse branch)
The default workload is BASELINE.json configs[4], the configuration the metric is quoted on
(1M cells x 30k genes, 10 % nnz, 50-D embedding, 100 grid points, 100 x 100 randomizations); it fits one B200
(24 GB of CSR).  With N > 1 the SAME 30 000 genes are cut into N contiguous, nnz-balanced slices
("scaling": "strong", what north_star asks for: `genes_total` is 30 000 at every N); --scaling weak gives every rank
its own 30k genes instead.  The multi-GPU logic is the package's (singlecellhaystack_b200.distributed.DeviceShard):
the density stage is sharded over the CELLS and rebuilt on every rank by three in-place NCCL all-gathers, every rank
scores its gene slice and its LPT-dealt share of the (gene, 25-permutation) units in one batched call, and D_KL is
all-gathered.  `--config c1|c2|c3|c3b|c4|c4b` selects the other BASELINE.json shapes as the workload; by default their
device-resident throughputs ride along under `extra.configs`.

The draws of K4, sample(x = n_cells, size = nnz) (:275), are made on the device from a seed (hs_dkl_perm_csr_seeded*);
the line also carries an end-to-end leg with host-drawn ids through hs_dkl_perm_csr (north_star's contract path) and
one from pageable host memory (what R vectors are) under `extra`.

`value` is measured with every input resident in HBM; `e2e` goes through the host-pointer C ABI with the inputs in
pinned host memory holding R's types (fp64 values, int32 ids).  The `cpu_baseline` object and `--impl reference` time
the fp64 oracle -- a port of the R code, pinned to the reference's printed vignette table
(tests/test_r_parity_cpu.py); R itself exists neither here nor on the GPU box -- on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "genes/sec incl. randomizations (1M cells, 100 grid pts)"
UNIT = "genes/s"
PERM_SEED = 20261019          # seed of the device-side randomization draw


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS), help="BASELINE.json workload shape (default c5 = configs[4])")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: the SAME genes cut into N slices (strong, default) or --genes per GPU (weak)")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the side runs of the other configs")
    # workload overrides (development runs only; unset = the --config shape)
    ap.add_argument("--cells", type=int, default=None)
    ap.add_argument("--genes", type=int, default=None, help="genes of the whole job (strong) / per GPU (weak)")
    ap.add_argument("--density", type=float, default=None)
    ap.add_argument("--dims", type=int, default=None)
    ap.add_argument("--grid-points", type=int, default=100)
    ap.add_argument("--rand-genes", type=int, default=100)
    ap.add_argument("--rand-count", type=int, default=100)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-genes", type=int, default=0, help="0 = all genes if host memory allows")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-f32", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--density-stage", dest="density_mode", default="cells", choices=["cells", "broadcast", "replicate"],
                    help="N > 1: the cells are sharded and the state all-gathered (default); rank 0 computes and "
                         "NCCL-broadcasts; or every rank recomputes everything (all deterministic: identical bits)")
    ap.add_argument("--no-e2e-extra", action="store_true", help="skip the host-draw and pageable end-to-end legs")
    ap.add_argument("--perm-draw", default="device", choices=["device", "host"],
                    help="randomization draw sample(n_cells, nnz) (R/haystack_continuous.R:275): made on the device "
                         "from a seed (hs_dkl_perm_csr_seeded, default) or passed in as ids (hs_dkl_perm_csr)")
    ap.add_argument("--e2e-separate-calls", action="store_true",
                    help="e2e: hs_dkl_csr then hs_dkl_perm_csr_seeded instead of the single hs_dkl_csr_randomized call")
    ap.add_argument("--no-dense", action="store_true", help="skip the side measurement of the dense (tcgen05) kernel")
    ap.add_argument("--dense-genes", type=int, default=1024)
    ap.add_argument("--cpu-genes", type=int, default=24)
    ap.add_argument("--cpu-perms", type=int, default=8)
    ap.add_argument("--cpu-density-cells", type=int, default=50_000)
    ap.add_argument("--ref-genes", type=int, default=0, help="genes per step of --impl reference (0 = 4 per core)")
    a = ap.parse_args()
    return apply_config(a, a.config)


# BASELINE.json configs -> synthetic shapes (cells, genes, embedding dims, mean density)
CONFIGS = {
    "c1": dict(cells=601, genes=500, dims=2, density=0.16,
               label="configs[0]: toy dataset shape (601 cells x 500 genes, 2-D coordinates)"),
    "c2": dict(cells=100_000, genes=16_811, dims=50, density=0.04,
               label="configs[1]: MOCA-like 100k cells x 16 811 genes sparse (4% nnz), 50 PCs"),
    "c3": dict(cells=2_696, genes=12_382, dims=2, density=0.30,
               label="configs[2]: Visium-like 2 696 spots x 12 382 genes (30% nnz), 2-D coordinates"),
    "c3b": dict(cells=2_696, genes=30_000, dims=2, density=0.30,
                label="configs[2]: Visium-like 2 696 spots x 30 000 genes (30% nnz), 2-D coordinates"),
    "c4": dict(cells=1_399, genes=23_341, dims=1, density=0.10,
               label="configs[3]: pseudotime 1 399 cells x 23 341 genes (10% nnz), 1-D coordinate"),
    "c4b": dict(cells=53_173, genes=11_390, dims=2, density=0.015,
                label="configs[3]: Slide-seqV2-like 53 173 beads x 11 390 genes (1.5% nnz), 2-D coordinates"),
    "c5": dict(cells=1_000_000, genes=30_000, dims=50, density=0.10,
               label="configs[4]: synthetic 1M cells x 30k genes sparse (10% nnz), 50-D embedding"),
}


def apply_config(a, name):
    """Fill the unset workload fields of `a` from CONFIGS[name]; remembers whether anything was overridden."""
    cfg = CONFIGS[name]
    a.config = name
    a.overridden = False
    for k in ("cells", "genes", "density", "dims"):
        if getattr(a, k, None) is None:
            setattr(a, k, cfg[k])
        elif getattr(a, k) != cfg[k]:
            a.overridden = True
    return a


def workload_name(a) -> str:
    tail = f", {a.grid_points} grid points, {a.rand_genes}x{a.rand_count} randomizations"
    if not getattr(a, "overridden", True) and (a.grid_points, a.rand_genes, a.rand_count) == (100, 100, 100):
        return CONFIGS[a.config]["label"] + tail
    return (f"DEV synthetic {a.cells} cells x {a.genes} genes sparse ({a.density:.3g} nnz), {a.dims}-D" + tail)


# ------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------
def make_coords(a):
    """Embedding (host, numpy), scale()d as R/haystack_continuous.R:139-145, and the grid (an input)."""
    from singlecellhaystack_b200 import synth
    x, labels = synth.make_embedding(a.cells, a.dims)
    xs, _, _ = synth.scale_coords(x)
    grid = synth.make_grid(xs, a.grid_points)
    return np.asfortranarray(xs), labels, np.asfortranarray(grid)


def make_csr_device(a, labels_np, gene_offset: int, dev, n_genes=None, total_genes=None):
    """Genes [gene_offset, gene_offset + n_genes) of a matrix of `total_genes` genes as device CSR: p int64 [M+1],
    j int32, x fp32 (values are the fp32 rounding of the fp64 table the host path would receive)."""
    import torch
    from singlecellhaystack_b200 import synth
    n_genes = a.genes if n_genes is None else int(n_genes)
    m_total = gene_offset + n_genes
    total = max(m_total, n_genes) if total_genes is None else int(total_genes)
    rho = synth.gene_densities(total, a.density)
    plant = synth.gene_plant(total, int(labels_np.max()) + 1)
    labels = torch.as_tensor(labels_np, device=dev)
    est = int(rho[gene_offset:m_total].sum() * a.cells * 1.02) + 1_000_000
    cols = torch.empty(est, dtype=torch.int32, device=dev)
    vals = torch.empty(est, dtype=torch.float32, device=dev)
    counts = torch.empty(n_genes, dtype=torch.int64, device=dev)
    block = max(1, min(256, (64 << 20) // max(a.cells, 1)))
    pos = 0
    for g0 in range(0, n_genes, block):
        g1 = min(n_genes, g0 + block)
        c, j, v = synth.expression_block(gene_offset + g0, gene_offset + g1, a.cells, rho, plant, labels)
        n = int(j.numel())
        if pos + n > est:
            raise RuntimeError("synthetic nnz estimate too small")
        cols[pos:pos + n] = j
        vals[pos:pos + n] = v
        counts[g0:g1] = c
        pos += n
    p = torch.zeros(n_genes + 1, dtype=torch.int64, device=dev)
    p[1:] = torch.cumsum(counts, 0)
    return p, cols[:pos], vals[:pos]


def row_stats_device(p, vals, n_cells: int):
    """mean, sd (n-1, zeros included), coefficient of variation per gene (R/haystack_continuous.R:57-83,
    :222) -- harness-side input preparation, fp64."""
    import torch
    v = vals.double()
    cs = torch.zeros(v.numel() + 1, dtype=torch.float64, device=v.device)
    cs[1:] = torch.cumsum(v, 0)
    cs2 = torch.zeros_like(cs)
    cs2[1:] = torch.cumsum(v * v, 0)
    s = cs[p[1:]] - cs[p[:-1]]
    ss = cs2[p[1:]] - cs2[p[:-1]]
    mean = s / n_cells
    var = (ss - n_cells * mean * mean) / (n_cells - 1)
    sd = torch.sqrt(torch.clamp(var, min=0))
    return (mean + 1e-300).cpu().numpy(), (sd + 1e-300).cpu().numpy()


def make_perm_inputs_device(a, p, vals, sel_local, sel_global, perm_lo: int, perm_hi: int, dev, want_ids=True):
    """Stand-in for R's sample(x = count.cells, size = nnz) (:275): for every selected gene held by
    this rank, permutations perm_lo..perm_hi-1 as an nnz x R_local column-major int32 block of
    distinct 0-based cell ids (Feistel bijection of [0, N))."""
    import torch
    from singlecellhaystack_b200 import synth
    r_loc = perm_hi - perm_lo
    gp = [0]
    val_parts, ind_parts = [], []
    for gl, gg in zip(sel_local, sel_global):
        lo, hi = int(p[gl]), int(p[gl + 1])
        nnz = hi - lo
        gp.append(gp[-1] + nnz)
        val_parts.append(vals[lo:hi])
        if not want_ids:
            continue
        idx = torch.arange(nnz, device=dev, dtype=torch.int64)[None, :].expand(r_loc, nnz)
        stream = (int(gg) * 100003 + torch.arange(perm_lo, perm_hi, device=dev, dtype=torch.int64))[:, None].expand(r_loc, nnz)
        ids = synth.feistel_perm(synth.BASE_SEED, stream.contiguous(), idx.contiguous(), a.cells)
        ind_parts.append(ids.to(torch.int32).reshape(-1))   # (R_local, nnz) row-major == nnz x R_local col-major
    gene_ptr = torch.tensor(gp, dtype=torch.int64, device=dev)
    val = torch.cat(val_parts) if val_parts else torch.zeros(0, dtype=torch.float32, device=dev)
    ind = torch.cat(ind_parts) if ind_parts else torch.zeros(0, dtype=torch.int32, device=dev)
    return gene_ptr, val.contiguous(), ind.contiguous(), r_loc


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
            self._f = f
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self._f.close()
        sm, mx, reasons, power = [], [], set(), []
        for line in Path(self.path).read_text().splitlines():
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        # samples under load = the upper half of the observed clocks (idle samples precede the region)
        med = float(np.median(sm)) if sm else None
        return {"sm_mhz": med, "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power)) if power else None}


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle = checker and reported baseline; never on the product path)
# ------------------------------------------------------------------------------------------------
def oracle_density_fast(xs, grid):
    """fp64 density matrix / Q for the oracle's per-gene timing, built with scipy's cdist so the
    untimed set-up stays short; the *timed* density sample uses the faithful oracle function."""
    from scipy.spatial.distance import cdist
    from oracle import haystack_oracle as ho
    dist = cdist(xs, grid)
    bw = float(np.median(dist.min(axis=1)))
    dist /= bw
    dens = np.exp(-(dist * dist) / 2)
    q = dens.sum(axis=0) + ho.PSEUDO
    return dens, q / q.sum(), bw


def cpu_sample_times(a, xs, grid, genes_rows, perm_gene, perm_ids1):
    """Time the oracle on a bounded sample.  genes_rows: list of (ind1, val); perm_gene: (val,);
    perm_ids1: (k, nnz) 1-based.  Returns dict of per-unit seconds."""
    from oracle import haystack_oracle as ho
    nsub = min(a.cells, a.cpu_density_cells)
    t0 = time.perf_counter()
    ho.density_stage(xs[:nsub], grid)
    t_density = (time.perf_counter() - t0) * (a.cells / nsub)
    dens, q, _ = oracle_density_fast(xs, grid)
    t0 = time.perf_counter()
    out = [ho.get_D_KL_continuous_highD_SPARSE(ind, val, dens, q, ho.PSEUDO) for ind, val in genes_rows]
    t_gene = (time.perf_counter() - t0) / max(len(genes_rows), 1)
    nnz_sample = float(np.mean([len(v) for _, v in genes_rows])) if genes_rows else 0.0
    t0 = time.perf_counter()
    out_perm = [ho.get_D_KL_continuous_highD_SPARSE(perm_ids1[r], perm_gene, dens, q, ho.PSEUDO)
                for r in range(perm_ids1.shape[0])]
    t_perm = (time.perf_counter() - t0) / max(perm_ids1.shape[0], 1)
    return dict(t_density=t_density, t_gene=t_gene, t_perm=t_perm, nnz_sample=nnz_sample, dkl=np.array(out),
                dkl_perm=np.array(out_perm))


def run_reference(a):
    """--impl reference: the oracle port of the reference's per-gene loop on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from singlecellhaystack_b200 import synth
    from oracle import haystack_oracle as ho
    cores = os.cpu_count() or 1
    xs, labels, grid = make_coords(a)
    n_genes = a.ref_genes or 4 * cores
    rho = synth.gene_densities(a.genes, a.density)
    plant = synth.gene_plant(a.genes, int(labels.max()) + 1)
    pick = np.linspace(0, a.genes - 1, n_genes).astype(np.int64)   # spread over the density spectrum
    rows = []
    for g in pick:
        _, cols, vals = synth.expression_block(int(g), int(g) + 1, a.cells, rho, plant, labels)
        rows.append((cols.astype(np.int64) + 1, vals.astype(np.float64)))
    mean_nnz_all = float(rho.mean() * a.cells)
    mean_nnz_s = float(np.mean([len(v) for _, v in rows]))
    dens, q, _ = oracle_density_fast(xs, grid)
    nsub = min(a.cells, a.cpu_density_cells)
    global _REF_STATE
    _REF_STATE = (rows, dens, q)

    wall = []

    def one_step():
        t00 = t0 = time.perf_counter()
        ho.density_stage(xs[:nsub], grid)
        t_density = (time.perf_counter() - t0) * (a.cells / nsub)   # single-threaded in the reference
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_ref_gene, range(len(rows)), chunksize=max(1, len(rows) // (cores * 2)))
        t_genes = time.perf_counter() - t0
        # per-nnz cost is what scales: normalise the sample to the workload's mean row length
        per_gene = t_genes / len(rows) * (mean_nnz_all / max(mean_nnz_s, 1.0))
        total = t_density + per_gene * (a.genes + a.rand_genes * a.rand_count)
        wall.append(time.perf_counter() - t00)
        return total

    for _ in range(a.warmup):
        one_step()
    wall.clear()
    tot = [one_step() for _ in range(a.steps)]
    t = float(np.mean(tot))
    value = a.genes / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": float(np.mean(wall)) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "full_workload_s_extrapolated": t,
                   "note": "a step is a bounded sample; value = genes of the workload / extrapolated full-workload time"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": (f"fp64 NumPy oracle of R/haystack_continuous.R:586-602 on {len(rows)} genes spread over the "
                                    f"density spectrum (fork pool, {cores} processes) + density stage on {nsub} cells; "
                                    f"extrapolated per nnz to {a.genes} genes + {a.rand_genes}x{a.rand_count} permuted rows")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REF_STATE = None


def _ref_gene(i):
    from oracle import haystack_oracle as ho
    rows, dens, q = _REF_STATE
    ind, val = rows[i]
    return ho.get_D_KL_continuous_highD_SPARSE(ind, val, dens, q, ho.PSEUDO)


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
class Workload:
    """This rank's share of one synthetic job, resident in HBM, plus the host-side bookkeeping of the split."""


def prepare_workload(a, dev, world, rank, scaling):
    """Synthetic inputs of one job.  strong: a.genes genes in all, cut into `world` contiguous nnz-balanced slices
    (distributed.plan_gene_shards); weak: a.genes genes per rank.  The selected genes' permuted rows are dealt to the
    ranks as (gene, 25-permutation) units (distributed.deal_permutation_units); the values of the selected rows are
    exchanged once here (inputs: ~100 rows)."""
    import torch
    import torch.distributed as dist
    from singlecellhaystack_b200 import haystack as hs_host
    from singlecellhaystack_b200 import synth
    from singlecellhaystack_b200.distributed import deal_permutation_units, plan_gene_shards
    w = Workload()
    w.xs, w.labels, w.grid = make_coords(a)
    w.n, w.g = a.cells, a.grid_points
    w.x_dev = torch.as_tensor(np.ascontiguousarray(w.xs.T), device=dev)        # (d, N) == column-major N x d
    w.grid_dev = torch.as_tensor(np.ascontiguousarray(w.grid.T), device=dev)   # (d, G)
    w.m_total = a.genes * (world if scaling == "weak" else 1)
    rho = synth.gene_densities(w.m_total, a.density)
    w.bounds = plan_gene_shards(rho * a.cells, world)
    g0, g1 = int(w.bounds[rank]), int(w.bounds[rank + 1])
    w.g0, w.m_local = g0, g1 - g0
    w.p, w.cols, w.vals = make_csr_device(a, w.labels, g0, dev, n_genes=w.m_local, total_genes=w.m_total)
    w.nnz = int(w.p[-1])
    # ---- gene selection for the randomization (harness-side; stays in R in the drop-in) ----
    mean, sd = row_stats_device(w.p, w.vals, w.n)
    cv_local = sd / mean
    if world > 1:
        cv_all = [None] * world
        dist.all_gather_object(cv_all, cv_local)
        cv = np.concatenate(cv_all)
    else:
        cv = cv_local
    w.sel = hs_host.select_genes_to_randomize(cv, min(a.rand_genes, w.m_total), "heavytails" if w.m_total >= 18 and a.rand_genes >= 18 else "uniform")
    p_host = w.p.cpu().numpy()
    own = {int(gid): w.vals[int(p_host[gid - g0]):int(p_host[gid - g0 + 1])] for gid in w.sel if g0 <= gid < g1}
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, [(gid, v.cpu().numpy()) for gid, v in own.items()])
        row_vals = {gid: torch.as_tensor(v, device=dev) for part in parts for gid, v in part}
    else:
        row_vals = own
    w.sel_nnz = np.array([int(row_vals[int(gid)].numel()) for gid in w.sel], dtype=np.int64)
    w.block = 25 if (world > 1 and a.rand_count % 25 == 0) else a.rand_count
    w.units = deal_permutation_units(w.sel_nnz, a.rand_count, world, w.block)[rank]     # [(s, r_lo, r_hi)]
    gp = [0]
    for s_, _, _ in w.units:
        gp.append(gp[-1] + int(w.sel_nnz[s_]))
    w.gene_ptr = torch.tensor(gp, dtype=torch.int64, device=dev)
    w.pval = (torch.cat([row_vals[int(w.sel[s_])] for s_, _, _ in w.units]).contiguous() if w.units
              else torch.zeros(0, dtype=torch.float32, device=dev))
    # stream of (selected gene s, permutation r) = s * R + r whoever draws it: the randomized matrix is the same for
    # every number of ranks
    w.unit_stream = torch.tensor([s_ * a.rand_count + r_lo for s_, r_lo, _ in w.units] or [0], dtype=torch.int64, device=dev)
    w.sel_rows_local = (w.sel - g0).astype(np.int64) if world == 1 else None
    return w


def run_device_leg(a, w, shard, ctx, dev, world, steps, warmup, with_clocks=False):
    """K1 + K3 + K4 (+ the gather) on HBM-resident inputs, timed with CUDA events on the library's stream."""
    import torch
    import torch.distributed as dist
    n_units = len(w.units)
    dkl = torch.empty(max(w.m_local, 1), dtype=torch.float64, device=dev)
    dkl_rand = torch.empty(max(n_units * w.block, 1), dtype=torch.float64, device=dev)
    ev = {"k1": [], "k3": [], "k4": []}
    state = {}

    def step(record):
        marks = []
        if record:
            marks.append(torch.cuda.Event(enable_timing=True)); marks[-1].record()
        shard.density(w.x_dev, w.grid_dev)
        if record:
            marks.append(torch.cuda.Event(enable_timing=True)); marks[-1].record()
        shard.score(w.p, w.cols, w.vals, dkl)
        if record:
            marks.append(torch.cuda.Event(enable_timing=True)); marks[-1].record()
        shard.randomize(w.gene_ptr, w.pval, w.block, PERM_SEED, w.unit_stream, dkl_rand)
        if record:
            marks.append(torch.cuda.Event(enable_timing=True)); marks[-1].record()
            ev["k1"].append((marks[0], marks[1])); ev["k3"].append((marks[1], marks[2])); ev["k4"].append((marks[2], marks[3]))
        state["dkl_all"] = shard.gather(dkl, w.bounds)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step(False)
    barrier()
    launches0 = ctx.launch_count
    sampler = ClockSampler(dev.index) if with_clocks else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step(True)
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    mean = lambda k: float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev[k]]))
    return dict(ms_step=ms_total / steps, k1_ms=mean("k1"), k3_ms=mean("k3"), k4_ms=mean("k4"), clocks=clocks,
                launches=ctx.launch_count - launches0, dkl=dkl, dkl_rand=dkl_rand, dkl_all=state["dkl_all"])


def hbm_peak():
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        return float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback (B200_PROFILING.md)"


def k3_roofline(w, k3_ms):
    """Algorithmic bytes of one K3 launch: 8 B per stored value (fp32 value + int32 cell id), row
    pointers, results, and the density matrix once (G floats per cell)."""
    k3_bytes = 8.0 * w.nnz + 8.0 * (w.m_local + 1) + 8.0 * w.m_local + 4.0 * w.n * w.g
    peak, peak_src = hbm_peak()
    achieved = k3_bytes / (k3_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": _ncu_traffic(),
            "kernel": "hs_dkl_csr_dev: k_tile_cuts + k_tile_fill + k_tile_mma (tile-bucketed CSR contraction on tcgen05 + fused KL epilogue)",
            "kernel_ms": k3_ms, "algorithmic_bytes": k3_bytes, "peak_source": peak_src}


def run_side_configs(a, ctx, dev):
    """Device-resident throughput of the other BASELINE.json shapes (one GPU), same code path as the headline."""
    import torch
    from singlecellhaystack_b200.distributed import DeviceShard
    out = {}
    for name in ("c1", "c2", "c3", "c3b", "c4", "c4b"):
        if name == a.config:
            continue
        b = argparse.Namespace(**vars(a))
        for k in ("cells", "genes", "density", "dims"):
            setattr(b, k, None)
        apply_config(b, name)
        b.grid_points = min(a.grid_points, max(2, b.cells // 6))
        w = prepare_workload(b, dev, 1, 0, "strong")
        r = run_device_leg(b, w, DeviceShard(ctx), ctx, dev, 1, steps=3, warmup=2)
        roof = k3_roofline(w, r["k3_ms"])
        out[name] = {"workload": workload_name(b), "genes_per_s": b.genes / (r["ms_step"] * 1e-3), "ms_per_step": r["ms_step"],
                     "k1_ms": r["k1_ms"], "k3_ms": r["k3_ms"], "k4_ms": r["k4_ms"], "nnz": w.nnz,
                     "k3_hbm_frac": roof["frac"], "k3_gbs": roof["achieved"]}
        del w, r
        torch.cuda.empty_cache()
    return out


def run_b200(a):
    import torch
    import torch.distributed as dist
    from singlecellhaystack_b200.device import DeviceContext
    from singlecellhaystack_b200.distributed import DeviceShard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t_setup = time.perf_counter()
    w = prepare_workload(a, dev, world, rank, a.scaling)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    ctx = DeviceContext(local)
    work_stream = torch.cuda.Stream(device=dev)          # the library, the collectives and the timing events share it
    torch.cuda.set_stream(work_stream)
    ctx.use_torch_stream(work_stream)
    shard = DeviceShard(ctx, density_mode=a.density_mode)

    r = run_device_leg(a, w, shard, ctx, dev, world, a.steps, a.warmup, with_clocks=(rank == 0))
    ms_step = r["ms_step"]
    value = w.m_total / (ms_step * 1e-3)
    roofline = k3_roofline(w, r["k3_ms"])
    peak = roofline["peak"]
    n_units = len(w.units)

    # ---- side measurements (one GPU only): the dense contraction (K2, tcgen05) and the other config shapes ----
    dense_extra = side = None
    if rank == 0 and world == 1 and not a.no_dense and w.g <= 128 and a.config == "c5" and not a.overridden:
        dense_extra = run_dense_side(a, ctx, dev, peak)
    if rank == 0 and world == 1 and not a.no_extra_configs and not a.overridden:
        side = run_side_configs(a, ctx, dev)
    convert_extra = None
    if rank == 0 and world == 1 and not a.no_dense and a.config == "c5" and not a.overridden:
        convert_extra = run_convert_side(w, ctx, dev, peak)

    # ---- end to end through the host-pointer C ABI (pinned host inputs, H2D inside the timed region) ----
    e2e = None
    e2e_extra = {}
    if not a.no_e2e:
        e2e = run_e2e(a, w, shard, ctx, world, rank, dev)
        if world == 1 and not a.no_e2e_extra:
            e2e_extra["e2e_pageable_host"] = run_e2e(a, w, shard, ctx, world, rank, dev, pinned=False)
            e2e_extra["e2e_host_drawn_ids"] = run_e2e(a, w, shard, ctx, world, rank, dev, perm_draw="host")
            if not a.no_e2e_f32:
                e2e_extra["e2e_float32_host"] = run_e2e(a, w, shard, ctx, world, rank, dev, value_dtype="float32")

    # ---- CPU baseline + parity spot check (rank 0, N = 1 only) ----
    cpu = parity = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu, parity = run_cpu_baseline(a, w, r["dkl"], r["dkl_rand"])

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": a.scaling if world > 1 else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "cells": w.n, "genes_total": w.m_total, "genes_rank0": w.m_local,
                       "nnz_rank0": w.nnz, "grid_points": w.g, "dims": a.dims, "rand_genes": int(len(w.sel)),
                       "rand_count": a.rand_count, "l2": "inputs (CSR %.1f GB on rank 0) larger than L2" % (8.0 * w.nnz / 1e9),
                       "setup_s": round(setup_s, 1), "density_mode": a.density_mode if world > 1 else "single",
                       "rank0_ms": {"density": round(r["k1_ms"], 3), "contraction": round(r["k3_ms"], 3),
                                    "randomization": round(r["k4_ms"], 3), "randomized_rows": int(n_units * w.block)},
                       "parallelism": ("1 process per GPU; the %d genes cut into %d contiguous nnz-balanced slices; density stage "
                                       "sharded over the cells + 3 in-place NCCL all-gathers (mode %s); the S x R randomized rows "
                                       "dealt as (gene, %d-permutation) units, LPT by nnz; D_KL all-gathered"
                                       % (w.m_total, world, a.density_mode, w.block)) if world > 1 else "1 GPU"},
            "clocks": r["clocks"], "gpu_launches": int(r["launches"]), "roofline": roofline,
        }
        if e2e is not None:
            line["e2e"] = e2e
        extra = {}
        if dense_extra is not None:
            extra["k2_dense_tcgen05"] = dense_extra
        if side:
            extra["configs"] = side
        if convert_extra is not None:
            extra["csc_to_csr"] = convert_extra
        extra.update(e2e_extra)
        if extra:
            line["extra"] = extra
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if parity is not None:
            line["parity_spot_check"] = parity
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_convert_side(w, ctx, dev, hbm_peak):
    """as(expression, "RsparseMatrix") on the device (hs_csc_to_csr_dev, R/haystack_continuous.R:113-116) at the full
    workload: the resident CSR matrix is first transposed into dgCMatrix slots by the same routine (rows = cells), then
    converted back and compared bit for bit.  Not part of `value`."""
    import torch
    m, n, nnz = w.m_local, w.n, w.nnz
    free_b, _ = torch.cuda.mem_get_info(dev)
    if free_b < 2.2 * 8 * nnz + (1 << 30):
        return {"skipped": "not enough free device memory for two more copies of the matrix"}
    cp = torch.empty(n + 1, dtype=torch.int64, device=dev)
    ri = torch.empty(nnz, dtype=torch.int32, device=dev)
    xc = torch.empty(nnz, dtype=torch.float32, device=dev)
    t = []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    ev[0].record()
    ctx.csc_to_csr_dev(w.p, w.cols, w.vals, n, m, cp, ri, xc)            # genes x cells CSR == CSC of the transpose
    ev[1].record()
    p2 = torch.empty(m + 1, dtype=torch.int64, device=dev)
    j2 = torch.empty(nnz, dtype=torch.int32, device=dev)
    x2 = torch.empty(nnz, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    ev[1].record()
    ctx.csc_to_csr_dev(cp, ri, xc, m, n, p2, j2, x2)
    ev[2].record()
    torch.cuda.synchronize()
    same = bool(torch.equal(p2, w.p) and torch.equal(j2, w.cols) and torch.equal(x2, w.vals))
    ms = ev[1].elapsed_time(ev[2])
    del cp, ri, xc, p2, j2, x2
    torch.cuda.empty_cache()
    alg = 2.0 * 8.0 * nnz                                                   # read ids + values, write ids + values
    return {"rows": m, "cols": n, "nnz": nnz, "ms": ms, "round_trip_bit_identical": same,
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": alg},
            "note": "dgCMatrix slots -> dgRMatrix slots (ids + fp32 values), host synchronisation of the call included"}


def run_dense_side(a, ctx, dev, hbm_peak):
    """K2 (dense expression, tensor cores) on a device-resident fp32 matrix of --dense-genes genes x all cells;
    not part of `value` (the headline workload is sparse), reported so the dense kernel has a number."""
    import torch
    m, n, g = a.dense_genes, a.cells, a.grid_points
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    e = torch.rand((n, m), device=dev, generator=gen)                  # (cells, genes) == genes x cells column-major
    e = torch.where(e < 0.1, torch.log1p(1 + 20 * e), torch.zeros_like(e)).contiguous()
    out = torch.empty(m, dtype=torch.float64, device=dev)
    for _ in range(2):
        ctx.dkl_dense_dev(e, m, m, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        ctx.dkl_dense_dev(e, m, m, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = 4.0 * n * m / (ms * 1e-3) / 1e9
    del e
    torch.cuda.empty_cache()
    return {"genes": m, "cells": n, "ms": ms, "genes_per_s": m / (ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                         "algorithmic_bytes": 4.0 * n * m},
            "mma_tflops_issued": 3 * 2.0 * n * m * 128 / (ms * 1e-3) / 1e12,
            "note": "fp32 expression, 3 fp16 MMAs per K step (hi/lo split operands), N padded to 128; includes the KL epilogue"}


def _ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    f = ROOT / "profiles" / "k3_traffic.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


def _host_buffer(n, dtype, pinned):
    """R's kind of memory is pageable; the contract leg uses pinned buffers.  Returns (torch view, numpy view)."""
    import torch
    if pinned:
        t = torch.empty(n, dtype=dtype).pin_memory()
    else:
        t = torch.from_numpy(np.empty(n, dtype={torch.int32: np.int32, torch.float32: np.float32, torch.float64: np.float64,
                                                torch.int64: np.int64}[dtype]))
    return t, t.numpy()


def run_e2e(a, w, shard, ctx, world, rank, dev, value_dtype="float64", pinned=True, perm_draw="device"):
    """The same step through the host-pointer C ABI: every input starts in host memory holding R's types (fp64
    values, int32 ids, fp64 coordinates) and every result ends there; host clock around the blocking calls."""
    import psutil
    import torch
    import torch.distributed as dist
    m = w.m_local
    nnz = w.nnz
    seeded = perm_draw == "device"
    n_units = len(w.units)
    r_loc = w.block
    vdt = torch.float64 if value_dtype == "float64" else torch.float32
    n_ind = 0 if seeded else int(w.pval.numel()) * r_loc
    need = (4.0 + (8.0 if value_dtype == "float64" else 4.0)) * nnz + 4.0 * n_ind
    avail = psutil.virtual_memory().available / max(world, 1)
    m_e2e = a.e2e_genes or m
    if a.e2e_genes == 0 and need > 0.6 * avail:
        m_e2e = max(1, int(m * 0.6 * avail / need))
    m_e2e = min(m, m_e2e)
    p_h = w.p[:m_e2e + 1].cpu()
    nnz_e = int(p_h[-1])
    # fp64 host values are narrowed to fp32 by the library's host threads before they cross PCIe (unless the rank has
    # four host threads or fewer to itself, three or more ranks share the node, or HS_NARROW_ON_DEVICE=1; hs_api.cu):
    # count the bytes that are actually copied
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
    threads = min(16, max(1, (os.cpu_count() or 1) // max(local_world, 1)))
    if os.environ.get("HS_HOST_THREADS", "").isdigit() and int(os.environ["HS_HOST_THREADS"]) > 0:
        threads = int(os.environ["HS_HOST_THREADS"])
    nod = os.environ.get("HS_NARROW_ON_DEVICE")
    narrow_dev = (nod not in (None, "", "0")) if nod is not None else (threads <= 4 or local_world >= 3)
    vbytes = 8.0 if (value_dtype == "float64" and narrow_dev) else 4.0
    j_t, j_np = _host_buffer(nnz_e, torch.int32, pinned)
    j_t.copy_(w.cols[:nnz_e])
    x_t, x_np = _host_buffer(nnz_e, vdt, pinned)
    x_t.copy_(w.vals[:nnz_e].to(vdt))
    p_np = p_h.numpy()
    gp_h = w.gene_ptr.cpu().numpy()
    pv_t, pv_np = _host_buffer(int(w.pval.numel()), vdt, pinned)
    pv_t.copy_(w.pval.to(vdt))
    pi_np = None
    if not seeded:
        # stand-in for R's own sample() (north_star's contract path): ids drawn up front, uploaded inside the step
        from singlecellhaystack_b200 import synth
        pi_t, pi_np = _host_buffer(n_ind, torch.int32, pinned)
        pos = 0
        for u, (s_, r_lo, r_hi) in enumerate(w.units):
            nnz_g = int(w.sel_nnz[s_])
            idx = torch.arange(nnz_g, device=dev, dtype=torch.int64)[None, :].expand(r_loc, nnz_g)
            stream = (int(w.sel[s_]) * 100003 + torch.arange(r_lo, r_hi, device=dev, dtype=torch.int64))[:, None].expand(r_loc, nnz_g)
            ids = synth.feistel_perm(synth.BASE_SEED, stream.contiguous(), idx.contiguous(), a.cells)
            pi_t[pos:pos + nnz_g * r_loc].copy_(ids.to(torch.int32).reshape(-1))      # (R, nnz) row-major == nnz x R col-major
            pos += nnz_g * r_loc
            del ids, idx, stream
    sel_rows = w.sel_rows_local
    combined = (seeded and world == 1 and value_dtype == "float64" and sel_rows is not None and n_units
                and int(np.max(sel_rows)) < m_e2e and not a.e2e_separate_calls)
    xs_t, xs_np = _host_buffer(w.n * a.dims, torch.float64, pinned)
    xs_t.copy_(torch.from_numpy(np.ascontiguousarray(w.xs.T)).reshape(-1))          # column-major N x d
    xs_np = xs_np.reshape(a.dims, w.n).T
    grid_np = np.asfortranarray(w.grid)
    x_dev2 = torch.empty((a.dims, w.n), dtype=torch.float64, device=dev) if world > 1 else None
    pv_dev = torch.empty(int(w.pval.numel()), dtype=torch.float32, device=dev) if world > 1 else None
    torch.cuda.synchronize()
    torch.cuda.empty_cache()        # the library allocates its upload buffers with cudaMalloc
    out_rand = np.empty((max(n_units, 1), r_loc), dtype=np.float64, order="F")
    rand_dev = torch.empty(max(n_units * r_loc, 1), dtype=torch.float64, device=dev)

    def step():
        if world == 1:
            ctx.density(xs_np, grid_np)
        else:      # every rank uploads the coordinates and computes its cells
            x_dev2.copy_(xs_t.reshape(a.dims, w.n), non_blocking=True)
            shard.density(x_dev2, w.grid_dev)
        if combined:
            d, _ = ctx.dkl_csr_randomized(p_np, j_np, x_np, sel_rows, r_loc, PERM_SEED, stream_base=0, out_rand=out_rand)
            return d
        d = ctx.dkl_csr(p_np, j_np, x_np)
        if n_units:
            if world > 1:
                pv_dev.copy_(pv_t.to(torch.float32) if vdt != torch.float32 else pv_t, non_blocking=True)
                shard.randomize(w.gene_ptr, pv_dev, r_loc, PERM_SEED, w.unit_stream, rand_dev)
                out_rand[:] = rand_dev[:n_units * r_loc].cpu().numpy().reshape(r_loc, n_units).T
            elif seeded:
                ctx.dkl_perm_csr_seeded_flat(gp_h, pv_np.astype(np.float64, copy=False), r_loc, PERM_SEED, stream_base=0, out=out_rand)
            else:
                ctx.dkl_perm_csr_flat(gp_h, pv_np, pi_np, r_loc, index_base=0, out=out_rand)
        if world > 1:
            d = shard.gather(torch.from_numpy(d).to(dev), w.bounds).cpu().numpy()
        return d

    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    per = dt / a.e2e_steps
    h2d = (8.0 * a.cells * a.dims + 8.0 * a.grid_points * a.dims + 8.0 * (m_e2e + 1) + (4.0 + vbytes) * nnz_e
           + 8.0 * len(gp_h) + vbytes * int(w.pval.numel()) + 4.0 * n_ind)
    d2h = 8.0 * m_e2e + 8.0 * n_units * r_loc + 8.0 * a.grid_points + 8.0
    genes_done = w.m_total if m_e2e == m else m_e2e * world
    calls = ("hs_density + hs_dkl_csr_randomized" if combined else
             ("hs_density" if world == 1 else "hs_density_shard_* + all-gathers") + " + hs_dkl_csr + "
             + ("hs_dkl_perm_csr_seeded_units_dev" if world > 1 else "hs_dkl_perm_csr_seeded" if seeded else "hs_dkl_perm_csr"))
    return {"value": genes_done / per, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": per * 1e3, "steps": a.e2e_steps, "genes_rank0": m_e2e,
            "host_value_bytes": 8 if value_dtype == "float64" else 4, "pcie_value_bytes": int(vbytes),
            "host_memory": "pinned" if pinned else "pageable", "perm_draw": perm_draw,
            "note": (f"{calls}; {'pinned' if pinned else 'pageable'} host buffers ({value_dtype} values, int32 ids), "
                     "timed with the host clock around the blocking calls; per-rank bytes"
                     + ("" if m_e2e == m else f"; host memory limits the e2e pass to the first {m_e2e} genes"))}


def run_cpu_baseline(a, w, dkl, dkl_rand):
    cores_used = 1
    p_np = w.p.cpu().numpy()
    pick = np.linspace(0, w.m_local - 1, min(a.cpu_genes, w.m_local)).astype(np.int64)
    rows = []
    for gidx in pick:
        lo, hi = int(p_np[gidx]), int(p_np[gidx + 1])
        rows.append((w.cols[lo:hi].cpu().numpy().astype(np.int64) + 1, w.vals[lo:hi].double().cpu().numpy()))
    n_units, r_loc = len(w.units), w.block
    k = min(a.cpu_perms, r_loc)
    if n_units:
        gp = w.gene_ptr.cpu().numpy()
        u = int(np.argmax(np.diff(gp) > 0)) if np.any(np.diff(gp) > 0) else 0
        s_, r_lo, _ = w.units[u]
        nnz_s = int(gp[u + 1] - gp[u])
        pv = w.pval[gp[u]:gp[u + 1]].double().cpu().numpy()
        # the ids the device drew for (selected gene s, randomization r), restated on the CPU
        from oracle import prp_oracle as po
        blk = np.stack([po.prp_sample(PERM_SEED, s_ * a.rand_count + r_lo + r, a.cells, nnz_s).astype(np.int64) for r in range(k)])
    else:
        u, pv, blk = 0, rows[0][1], np.zeros((0, len(rows[0][1])), dtype=np.int64)
    tm = cpu_sample_times(a, w.xs, w.grid, rows, pv, blk)
    mean_nnz = float((p_np[-1] - p_np[0]) / max(w.m_local, 1))
    per_nnz = tm["t_gene"] / max(tm["nnz_sample"], 1.0)
    total = tm["t_density"] + per_nnz * mean_nnz * w.m_local + tm["t_perm"] / max(len(pv), 1) * mean_nnz * a.rand_genes * a.rand_count
    got = dkl[pick].cpu().numpy()
    rel = float(np.max(np.abs(got - tm["dkl"]) / np.abs(tm["dkl"])))
    parity = {"genes_checked": int(len(pick)), "max_rel_err_dkl_vs_oracle": rel, "tolerance": 1e-5}
    if n_units and len(tm["dkl_perm"]):
        got_p = dkl_rand[:n_units * r_loc].cpu().numpy().reshape(r_loc, n_units)[:k, u]      # units x R column-major
        parity["randomized_rows_checked"] = int(k)
        parity["max_rel_err_randomized_vs_oracle"] = float(np.max(np.abs(got_p - tm["dkl_perm"]) / np.abs(tm["dkl_perm"])))
    cpu = {"value": w.m_local / total, "unit": UNIT, "cores": cores_used, "kind": "port",
           "sample": (f"fp64 NumPy oracle (port of R/haystack_continuous.R:168-186, :586-602, pinned to the reference's printed "
                      f"vignette table; R itself is not installable here): density stage on {min(a.cells, a.cpu_density_cells)} cells, "
                      f"{len(rows)} genes (mean nnz {tm['nnz_sample']:.0f}) and {blk.shape[0]} permuted rows, single thread like R; "
                      f"extrapolated per nnz to the full workload"),
           "seconds_full_workload_extrapolated": total}
    return cpu, parity


def main():
    a = parse_args()
    # Only the JSON line may reach stdout (NCCL / torchrun print banners there): everything else the
    # process writes to fd 1 is routed to stderr, the JSON goes to the saved descriptor.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


_JSON_OUT = None


def emit(line: dict):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    main()

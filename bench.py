#!/usr/bin/env python
"""bench.py -- genes/sec incl. randomizations of the continuous high-D scoring path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One *step* = one pass of the scored path of haystack() over one synthetic data set:
  K1  density stage            (R/haystack_continuous.R:168-186)
  K3  observed D_KL, all genes (R/haystack_continuous.R:199-213, sparse branch)
  K4  randomized D_KL, S genes x R permutations (R/haystack_continuous.R:245-284, sparse branch)
Workload at N=1 is BASELINE.json configs[4], the configuration the metric is quoted on
(1M cells x 30k genes, 10 % nnz, 50-D embedding, 100 grid points, 100 x 100 randomizations);
it fits one B200 (24 GB of CSR).  With N > 1 every rank holds its own 30k-gene slice of a
30k*N-gene matrix ("scaling": "weak"): rank 0 runs K1 and NCCL-broadcasts the density matrix,
every rank scores its slice and its share of the (gene, permutation-block) units, and D_KL is
all-gathered.

The draws of K4, sample(x = n_cells, size = nnz) (:275), are made on the device from a seed by default
(hs_dkl_perm_csr_seeded; --perm-draw host passes pre-drawn ids through hs_dkl_perm_csr instead).

`value` is measured with every input resident in HBM; `e2e` goes through the host-pointer C ABI
(hs_density + hs_dkl_csr_randomized at N=1, hs_dkl_csr + hs_dkl_perm_csr_seeded per rank at N>1) with the
inputs in pinned host memory holding R's types (fp64 values, int32 ids).  The
`cpu_baseline` object and `--impl reference` time the fp64 oracle (a restatement of the R code;
R itself exists neither here nor on the GPU box) on the box's host cores.
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
    # workload (defaults = BASELINE.json configs[4]); overrides are for development runs only
    ap.add_argument("--cells", type=int, default=1_000_000)
    ap.add_argument("--genes", type=int, default=30_000, help="genes per GPU")
    ap.add_argument("--density", type=float, default=0.10)
    ap.add_argument("--dims", type=int, default=50)
    ap.add_argument("--grid-points", type=int, default=100)
    ap.add_argument("--rand-genes", type=int, default=100)
    ap.add_argument("--rand-count", type=int, default=100)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-genes", type=int, default=0, help="0 = all genes if host memory allows")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-f32", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--density-stage", dest="density_mode", default="broadcast", choices=["broadcast", "replicate"],
                    help="N > 1: rank 0 runs the density stage and NCCL-broadcasts its state (default), or every rank "
                         "recomputes it (deterministic kernels: identical bits)")
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
    return ap.parse_args()


def workload_name(a) -> str:
    if (a.cells, a.genes, a.dims, a.grid_points, a.rand_genes, a.rand_count) == (1_000_000, 30_000, 50, 100, 100, 100) \
            and abs(a.density - 0.10) < 1e-12:
        return "configs[4]: synthetic 1M cells x 30k genes sparse (10% nnz), 50-D embedding, 100 grid points, 100x100 randomizations"
    return (f"DEV synthetic {a.cells} cells x {a.genes} genes sparse ({a.density:.3g} nnz), {a.dims}-D, "
            f"{a.grid_points} grid points, {a.rand_genes}x{a.rand_count} randomizations")


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


def make_csr_device(a, labels_np, gene_offset: int, dev):
    """This rank's gene slice as device CSR: p int64 [M+1], j int32, x fp32 (values are the
    fp32 rounding of the fp64 table the host path would receive)."""
    import torch
    from singlecellhaystack_b200 import synth
    m_total = gene_offset + a.genes
    rho = synth.gene_densities(max(m_total, a.genes), a.density)
    plant = synth.gene_plant(max(m_total, a.genes), int(labels_np.max()) + 1)
    labels = torch.as_tensor(labels_np, device=dev)
    est = int(rho[gene_offset:m_total].sum() * a.cells * 1.02) + 1_000_000
    cols = torch.empty(est, dtype=torch.int32, device=dev)
    vals = torch.empty(est, dtype=torch.float32, device=dev)
    counts = torch.empty(a.genes, dtype=torch.int64, device=dev)
    block = max(1, min(256, (64 << 20) // max(a.cells, 1)))
    pos = 0
    for g0 in range(0, a.genes, block):
        g1 = min(a.genes, g0 + block)
        c, j, v = synth.expression_block(gene_offset + g0, gene_offset + g1, a.cells, rho, plant, labels)
        n = int(j.numel())
        if pos + n > est:
            raise RuntimeError("synthetic nnz estimate too small")
        cols[pos:pos + n] = j
        vals[pos:pos + n] = v
        counts[g0:g1] = c
        pos += n
    p = torch.zeros(a.genes + 1, dtype=torch.int64, device=dev)
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
        "warmup": a.warmup, "ms_per_step": float(np.mean(wall)) * 1e3, "higher_is_better": True, "scaling": "weak",
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
def run_b200(a):
    import torch
    import torch.distributed as dist
    from singlecellhaystack_b200 import haystack as hs_host
    from singlecellhaystack_b200.device import DeviceContext

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
    xs, labels, grid = make_coords(a)
    n, g = a.cells, a.grid_points
    x_dev = torch.as_tensor(np.ascontiguousarray(xs.T), device=dev)        # (d, N) == column-major N x d
    grid_dev = torch.as_tensor(np.ascontiguousarray(grid.T), device=dev)   # (d, G)
    p, cols, vals = make_csr_device(a, labels, rank * a.genes, dev)
    nnz = int(p[-1])
    m_local, m_total = a.genes, a.genes * world

    # ---- gene selection for the randomization (harness-side; stays in R in the drop-in) ----
    mean, sd = row_stats_device(p, vals, n)
    cv_local = sd / mean
    if world > 1:
        cv_all = [None] * world
        dist.all_gather_object(cv_all, cv_local)
        cv = np.concatenate(cv_all)
    else:
        cv = cv_local
    sel = hs_host.select_genes_to_randomize(cv, min(a.rand_genes, m_total), "heavytails")   # global 0-based
    p_host = p.cpu().numpy()
    if world == 1:
        sel_local = sel.astype(np.int64)
        gene_ptr, pval, pind, r_loc = make_perm_inputs_device(a, p_host, vals, sel_local, sel, 0, a.rand_count, dev,
                                                              want_ids=a.perm_draw == "host")
        mine = np.arange(len(sel))
        a._sel_rows = sel_local
    else:
        # units = (selected gene, block of permutations), dealt to ranks by descending cost (LPT) so that the
        # fixed total of S x R permuted rows is balanced; the selected rows' values are exchanged once
        # during set-up (inputs: ~100 rows), the indices are generated where they are used.
        owner = sel // m_local
        own = [(int(gid), vals[int(p_host[gid - rank * m_local]):int(p_host[gid - rank * m_local + 1])].cpu().numpy())
               for gid in sel[owner == rank]]
        rows_all = [None] * world
        dist.all_gather_object(rows_all, own)
        row_vals = {gid: v for part in rows_all for gid, v in part}
        n_blocks = 4 if a.rand_count % 4 == 0 else 1
        r_loc = a.rand_count // n_blocks
        units = sorted(((len(row_vals[int(gid)]), int(gid), b) for gid in sel for b in range(n_blocks)), reverse=True)
        load = np.zeros(world)
        mine_units = []
        for cost, gid, b in units:
            r = int(np.argmin(load))
            load[r] += cost
            if r == rank:
                mine_units.append((gid, b))
        gp = [0]
        val_parts, ind_parts = [], []
        from singlecellhaystack_b200 import synth
        for gid, b in mine_units:
            v = torch.as_tensor(row_vals[gid], device=dev)
            nnz_g = int(v.numel())
            gp.append(gp[-1] + nnz_g)
            val_parts.append(v)
            if a.perm_draw != "host":
                continue
            idx = torch.arange(nnz_g, device=dev, dtype=torch.int64)[None, :].expand(r_loc, nnz_g)
            stream = (gid * 100003 + torch.arange(b * r_loc, (b + 1) * r_loc, device=dev, dtype=torch.int64))[:, None].expand(r_loc, nnz_g)
            ids = synth.feistel_perm(synth.BASE_SEED, stream.contiguous(), idx.contiguous(), a.cells)
            ind_parts.append(ids.to(torch.int32).reshape(-1))
        gene_ptr = torch.tensor(gp, dtype=torch.int64, device=dev)
        pval = torch.cat(val_parts).contiguous() if val_parts else torch.zeros(0, dtype=torch.float32, device=dev)
        pind = torch.cat(ind_parts).contiguous() if ind_parts else torch.zeros(0, dtype=torch.int32, device=dev)
        mine = np.arange(len(mine_units))
    n_sel_local = len(mine)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    ctx = DeviceContext(local)
    work_stream = torch.cuda.Stream(device=dev)          # the library and the timing events share it
    torch.cuda.set_stream(work_stream)
    ctx.use_torch_stream(work_stream)
    dkl = torch.empty(m_local, dtype=torch.float64, device=dev)
    dkl_rand = torch.empty(max(n_sel_local * r_loc, 1), dtype=torch.float64, device=dev)
    dkl_all = torch.empty(m_total, dtype=torch.float64, device=dev) if world > 1 else dkl

    ev_k3 = []

    ev_bc = []
    perm_stream_base = rank << 40          # the ranks draw from disjoint streams
    ev_k4 = []

    def step(record: bool):
        if world == 1 or rank == 0 or a.density_mode == "replicate":
            ctx.density_dev(x_dev, grid_dev)
        if world > 1 and a.density_mode == "broadcast":
            b0 = b1 = None
            if record:
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record()
            ctx.density_broadcast(n, g, src=0)
            if record:
                b1.record()
                ev_bc.append((b0, b1))
        e0 = e1 = None
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        ctx.dkl_csr_dev(p, cols, vals, dkl)
        if record:
            e1.record()
            ev_k3.append((e0, e1))
        if n_sel_local:
            if a.perm_draw == "device":
                ctx.dkl_perm_csr_seeded_dev(gene_ptr, pval, r_loc, PERM_SEED, dkl_rand, stream_base=perm_stream_base)
            else:
                ctx.dkl_perm_csr_dev(gene_ptr, pval, pind, r_loc, dkl_rand, index_base=0)
        if record:
            e2 = torch.cuda.Event(enable_timing=True)
            e2.record()
            ev_k4.append((e1, e2))
        if world > 1:
            dist.all_gather_into_tensor(dkl_all, dkl)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        step(False)
    barrier()
    launches0 = ctx.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        step(True)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count - launches0
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / a.steps
    k3_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev_k3]))
    bc_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev_bc])) if ev_bc else None
    k4_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev_k4]))
    value = m_total / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (K3 CSR contraction), algorithmic bytes per launch ----
    k3_bytes = 8.0 * nnz + 8.0 * (m_local + 1) + 8.0 * m_local + 4.0 * n * g      # algorithmic: G floats per cell
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak = float(json.loads(peaks_file.read_text())["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = k3_bytes / (k3_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": _ncu_traffic(), "kernel": "k_slab_spmm (hs_dkl_csr_dev: slab-staged CSR contraction + KL epilogue)", "kernel_ms": k3_ms,
                "algorithmic_bytes": k3_bytes, "peak_source": peak_src}

    # ---- side measurement: the dense contraction (K2, tcgen05) on the same density state ----
    dense_extra = None
    if rank == 0 and world == 1 and not a.no_dense and g <= 128:
        dense_extra = run_dense_side(a, ctx, dev, peak)

    # ---- end to end through the host-pointer C ABI (pinned host inputs, H2D inside the timed region) ----
    e2e = e2e32 = None
    if not a.no_e2e:
        e2e = run_e2e(a, ctx, xs, grid, p, cols, vals, gene_ptr, pval, pind, r_loc, n_sel_local, world, rank, dev)
        if world == 1 and not a.no_e2e_f32:
            # the same through the float32 entry points (hosts that hold float32 matrices, e.g. AnnData)
            e2e32 = run_e2e(a, ctx, xs, grid, p, cols, vals, gene_ptr, pval, pind, r_loc, n_sel_local, world, rank, dev,
                            value_dtype="float32")

    # ---- CPU baseline + parity spot check (rank 0, N = 1 only) ----
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu, parity = run_cpu_baseline(a, xs, grid, p, cols, vals, dkl, gene_ptr, pval, pind, r_loc, dkl_rand,
                                       n_sel_local)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "cells": n, "genes_per_gpu": m_local, "genes_total": m_total,
                       "nnz_per_gpu": nnz, "grid_points": g, "dims": a.dims, "rand_genes": int(len(sel)),
                       "rand_count": a.rand_count, "l2": "inputs (CSR %.1f GB) larger than L2" % (8.0 * nnz / 1e9),
                       "setup_s": round(setup_s, 1), "density_mode": a.density_mode if world > 1 else "single",
                       "density_broadcast_ms_rank0": bc_ms,
                       "rank0_ms": {"contraction": round(k3_ms, 3), "randomization": round(k4_ms, 3),
                                    "randomized_rows": int(n_sel_local * r_loc)},
                       "parallelism": ("1 process per GPU, gene-sharded (30k genes each); K1 on rank 0 + NCCL broadcast of the "
                                       "density state; the S x R randomized rows are a fixed total, dealt to the ranks as "
                                       "(gene, 25-permutation) units balanced by nnz") if world > 1 else "1 GPU"},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if dense_extra is not None:
            line.setdefault("extra", {})["k2_dense_tcgen05"] = dense_extra
        if e2e32 is not None:
            line.setdefault("extra", {})["e2e_float32_host"] = e2e32
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if parity is not None:
            line["parity_spot_check"] = parity
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


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


def run_e2e(a, ctx, xs, grid, p, cols, vals, gene_ptr, pval, pind, r_loc, n_sel_local, world, rank, dev,
            value_dtype="float64"):
    import psutil
    import torch
    import torch.distributed as dist
    m = a.genes
    nnz = int(p[-1])
    seeded = a.perm_draw == "device"
    n_ind = 0 if seeded else int(pind.numel())
    need = 12.0 * nnz + 12.0 * int(pval.numel()) + 4.0 * n_ind
    avail = psutil.virtual_memory().available / max(world, 1)
    m_e2e = a.e2e_genes or m
    if a.e2e_genes == 0 and need > 0.6 * avail:
        m_e2e = max(1, int(m * 0.6 * avail / need))
    m_e2e = min(m, m_e2e)
    p_h = p[:m_e2e + 1].cpu()
    nnz_e = int(p_h[-1])
    # R hands over doubles (dgRMatrix@x); keep them in pinned memory as the contract asks
    vdt = torch.float64 if value_dtype == "float64" else torch.float32
    # fp64 host values are narrowed to fp32 by the library's host threads before they cross PCIe
    # (HS_NARROW_ON_DEVICE=1 uploads the doubles instead): count the bytes that are actually copied
    vbytes = 8.0 if (value_dtype == "float64" and os.environ.get("HS_NARROW_ON_DEVICE", "0") not in ("", "0")) else 4.0
    j_h = torch.empty(nnz_e, dtype=torch.int32).pin_memory()
    j_h.copy_(cols[:nnz_e])
    x_h = torch.empty(nnz_e, dtype=vdt).pin_memory()
    x_h.copy_(vals[:nnz_e].to(vdt))
    gp_h = gene_ptr.cpu().numpy()
    pv_h = torch.empty(int(pval.numel()), dtype=vdt).pin_memory()
    pv_h.copy_(pval.to(vdt))
    pi_h = torch.empty(n_ind, dtype=torch.int32).pin_memory()
    if not seeded:
        pi_h.copy_(pind)
    pv64 = pv_h.numpy().astype(np.float64) if seeded else None     # the seeded entry point takes doubles
    perm_stream_base = rank << 40
    # one call for the observed scores and the seeded randomization batch (the batch computes while the matrix
    # uploads) when the selected rows are rows of this rank's matrix
    sel_rows = getattr(a, "_sel_rows", None)
    combined = (seeded and world == 1 and value_dtype == "float64" and sel_rows is not None and n_sel_local
                and int(np.max(sel_rows)) < m_e2e and not a.e2e_separate_calls)
    xs_h = torch.from_numpy(np.ascontiguousarray(xs.T)).pin_memory()       # column-major N x d
    grid_h = torch.from_numpy(np.ascontiguousarray(grid.T)).pin_memory()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()        # the library allocates its upload buffers with cudaMalloc
    p_np, j_np, x_np = p_h.numpy(), j_h.numpy(), x_h.numpy()
    out_rand = np.empty((max(n_sel_local, 1), r_loc), dtype=np.float64, order="F")

    def step():
        if world == 1 or rank == 0:
            ctx.density(xs_h.numpy().T, grid_h.numpy().T)
        if world > 1:
            ctx.density_broadcast(a.cells, a.grid_points, src=0)
        if combined:
            d, _ = ctx.dkl_csr_randomized(p_np, j_np, x_np, sel_rows, r_loc, PERM_SEED, stream_base=perm_stream_base,
                                          out_rand=out_rand)
            return d
        d = ctx.dkl_csr(p_np, j_np, x_np)
        if n_sel_local:
            if seeded:
                ctx.dkl_perm_csr_seeded_flat(gp_h, pv64, r_loc, PERM_SEED, stream_base=perm_stream_base, out=out_rand)
            else:
                ctx.dkl_perm_csr_flat(gp_h, pv_h.numpy(), pi_h.numpy(), r_loc, index_base=0, out=out_rand)
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
           + 8.0 * len(gp_h) + vbytes * int(pval.numel()) + 4.0 * n_ind)
    d2h = 8.0 * m_e2e + 8.0 * n_sel_local * r_loc + 8.0 * a.grid_points + 8.0
    return {"value": m_e2e * world / per, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": per * 1e3, "steps": a.e2e_steps, "genes_per_gpu": m_e2e,
            "host_value_bytes": 8 if value_dtype == "float64" else 4, "pcie_value_bytes": int(vbytes),
            "perm_draw": a.perm_draw,
            "note": (f"hs_density + " + ("hs_dkl_csr_randomized" if combined else
                                         f"hs_dkl_csr + {'hs_dkl_perm_csr_seeded' if seeded else 'hs_dkl_perm_csr'}") + " with pinned host "
                     f"buffers ({value_dtype} values, int32 ids), "
                     "timed with the host clock around the blocking calls"
                     + ("" if m_e2e == m else f"; host memory limits the e2e pass to the first {m_e2e} genes"))}


def run_cpu_baseline(a, xs, grid, p, cols, vals, dkl, gene_ptr, pval, pind, r_loc, dkl_rand, n_sel_local):
    cores_used = 1
    p_np = p.cpu().numpy()
    pick = np.linspace(0, a.genes - 1, min(a.cpu_genes, a.genes)).astype(np.int64)
    rows = []
    for gidx in pick:
        lo, hi = int(p_np[gidx]), int(p_np[gidx + 1])
        rows.append((cols[lo:hi].cpu().numpy().astype(np.int64) + 1, vals[lo:hi].double().cpu().numpy()))
    k = min(a.cpu_perms, r_loc)
    if n_sel_local:
        gp = gene_ptr.cpu().numpy()
        s = int(np.argmax(np.diff(gp) > 0)) if np.any(np.diff(gp) > 0) else 0
        nnz_s = int(gp[s + 1] - gp[s])
        pv = pval[gp[s]:gp[s + 1]].double().cpu().numpy()
        if a.perm_draw == "device":    # the ids the device drew for (gene s, randomization r), restated on the CPU
            from oracle import prp_oracle as po
            blk = np.stack([po.prp_sample(PERM_SEED, s * r_loc + r, a.cells, nnz_s).astype(np.int64) for r in range(k)])
        else:
            blk = pind[gp[s] * r_loc: gp[s] * r_loc + nnz_s * r_loc].reshape(r_loc, nnz_s)[:k].cpu().numpy().astype(np.int64) + 1
    else:
        s, pv, blk = 0, rows[0][1], np.zeros((0, len(rows[0][1])), dtype=np.int64)
    tm = cpu_sample_times(a, xs, grid, rows, pv, blk)
    mean_nnz = float((p_np[-1] - p_np[0]) / a.genes)
    per_nnz = tm["t_gene"] / max(tm["nnz_sample"], 1.0)
    total = tm["t_density"] + per_nnz * mean_nnz * a.genes + tm["t_perm"] / max(len(pv), 1) * mean_nnz * a.rand_genes * a.rand_count
    got = dkl[pick].cpu().numpy()
    rel = float(np.max(np.abs(got - tm["dkl"]) / np.abs(tm["dkl"])))
    parity = {"genes_checked": int(len(pick)), "max_rel_err_dkl_vs_oracle": rel, "tolerance": 1e-5}
    if n_sel_local and len(tm["dkl_perm"]):
        got_p = dkl_rand.cpu().numpy().reshape(r_loc, n_sel_local)[:k, s]      # n_sel x R column-major
        parity["randomized_rows_checked"] = int(k)
        parity["max_rel_err_randomized_vs_oracle"] = float(np.max(np.abs(got_p - tm["dkl_perm"]) / np.abs(tm["dkl_perm"])))
    cpu = {"value": a.genes / total, "unit": UNIT, "cores": cores_used, "kind": "port",
           "sample": (f"fp64 NumPy oracle (restatement of R/haystack_continuous.R:168-186, :586-602; R itself is not "
                      f"installable here): density stage on {min(a.cells, a.cpu_density_cells)} cells, {len(rows)} genes "
                      f"(mean nnz {tm['nnz_sample']:.0f}) and {blk.shape[0]} permuted rows, single thread like R; "
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

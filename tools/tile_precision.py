"""Development probe (GPU): error and time of the sparse contraction at C5 scale for several settings of the tile
kernel (HS_TILE_DRAIN, HS_TILE_BIASFIX, HS_TILE_DISABLE), against an fp64 reference of ALL genes computed on the GPU
with torch (cuSPARSE fp64 CSR x dense) -- a checker, not a product path.

    python tools/tile_precision.py [--cells N] [--genes M] "DRAIN=1" "DRAIN=2,BIASFIX=3.5" "DISABLE=1" ...
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from singlecellhaystack_b200.device import DeviceContext  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=1_000_000)
    ap.add_argument("--genes", type=int, default=30_000)
    ap.add_argument("--density", type=float, default=0.10)
    ap.add_argument("configs", nargs="*", default=["DRAIN=1"])
    a0 = ap.parse_args()
    a = argparse.Namespace(cells=a0.cells, genes=a0.genes, density=a0.density, dims=50, grid_points=100)
    dev = torch.device("cuda", 0)
    xs, labels, grid = bench.make_coords(a)
    p, cols, vals = bench.make_csr_device(a, labels, 0, dev)
    x_dev = torch.as_tensor(np.ascontiguousarray(xs.T), device=dev)
    g_dev = torch.as_tensor(np.ascontiguousarray(grid.T), device=dev)
    # fp64 reference
    with DeviceContext(0) as c:
        d = c.density_dev(x_dev, g_dev)
    bw, q = d["bandwidth"], torch.as_tensor(d["Q"], device=dev)
    X = torch.as_tensor(xs, device=dev)
    Gd = torch.as_tensor(grid, device=dev)
    D = torch.cdist(X, Gd) / bw
    D = torch.exp(-D * D / 2)
    ref = torch.empty(a.genes, dtype=torch.float64, device=dev)
    step = 2000
    for g0 in range(0, a.genes, step):
        g1 = min(a.genes, g0 + step)
        lo, hi = int(p[g0]), int(p[g1])
        sp = torch.sparse_csr_tensor((p[g0:g1 + 1] - p[g0]), cols[lo:hi].long(), vals[lo:hi].double(), size=(g1 - g0, a.cells))
        P = torch.sparse.mm(sp, D) + 1e-300
        P = P / P.sum(dim=1, keepdim=True)
        ref[g0:g1] = (P * torch.log(P / q[None, :])).sum(dim=1)
    del D, X
    torch.cuda.empty_cache()
    out = torch.empty(a.genes, dtype=torch.float64, device=dev)
    for cfg in a0.configs:
        for k in list(os.environ):
            if k.startswith("HS_TILE_"):
                del os.environ[k]
        for kv in cfg.split(","):
            k, v = kv.split("=")
            os.environ["HS_TILE_" + k] = v
        with DeviceContext(0) as c:
            st = torch.cuda.Stream(device=dev)
            torch.cuda.set_stream(st)
            c.use_torch_stream(st)
            c.density_dev(x_dev, g_dev)
            for _ in range(2):
                c.dkl_csr_dev(p, cols, vals, out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                c.dkl_csr_dev(p, cols, vals, out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
        rel = ((out - ref).abs() / ref.abs()).cpu().numpy()
        sgn = ((out - ref) / ref.abs()).cpu().numpy()
        print(f"{cfg:28s} {ms:8.2f} ms  rel err max {rel.max():.2e}  p99 {np.quantile(rel, 0.99):.2e}  median {np.median(rel):.2e}"
              f"  mean signed {sgn.mean():+.2e}", flush=True)


if __name__ == "__main__":
    main()

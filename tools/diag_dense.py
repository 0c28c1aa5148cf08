"""Development probe: tensor-core dense path vs SIMT vs oracle."""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import haystack_oracle as ho
from singlecellhaystack_b200 import synth
from singlecellhaystack_b200.device import DeviceContext

def rel(a, b): return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))

toy = np.load(Path(__file__).resolve().parents[1] / "tests/golden/toy.npz")
ctx = DeviceContext(0)
xs, _, _ = ho.scale_coords(toy["coord"])
expr = toy["expression"].astype(np.float64)
for g in (60, 100):
    ctx.density(xs, toy[f"grid{g}"])
    for mode in ("0", "1"):
        os.environ["HS_DISABLE_UMMA"] = mode
        t = time.perf_counter(); d = ctx.dkl_dense(expr); dt = time.perf_counter() - t
        print(f"toy G={g} disable_umma={mode}: rel err {rel(d, toy[f'dkl{g}']):.3e}  nan={np.isnan(d).sum()} time {dt*1e3:.2f} ms", flush=True)
os.environ["HS_DISABLE_UMMA"] = "0"
# perm dense
ref = ho.density_stage(xs, toy["grid100"]); ctx.density(xs, toy["grid100"])
src = synth.perm_source(int(toy["seed"]))
gidx = int(toy["full_genes_to_randomize"][0])
perms = src("dense", gidx, 601, 601, 100)
want = ho.dkl_randomized_dense(expr[[gidx]], perms[None], ref["density"], ref["Q"])[0]
for mode in ("0", "1"):
    os.environ["HS_DISABLE_UMMA"] = mode
    got = ctx.dkl_perm_dense(expr[gidx], np.asfortranarray(perms.T), index_base=1)
    print(f"toy perm dense disable_umma={mode}: rel err {rel(got, want):.3e}", flush=True)
os.environ["HS_DISABLE_UMMA"] = "0"
# synthetic, larger
for (n, d, g, m, rho) in [(20000, 10, 100, 300, 0.08), (100000, 20, 100, 1000, 0.1), (5000, 3, 37, 130, 0.3)]:
    x, labels = synth.make_embedding(n, d); xs2, _, _ = synth.scale_coords(x); grid = synth.make_grid(xs2, g)
    p, j, v = synth.make_csr_numpy(m, n, rho, labels)
    ref = ho.density_stage(xs2, grid)
    sp = ho.RowSparse(p=p, j=j, x=v, shape=(m, n))
    want = ho.dkl_observed(sp, ref["density"], ref["Q"])
    dense = sp.to_dense()
    ctx.density(xs2, grid)
    for mode in ("0", "1"):
        os.environ["HS_DISABLE_UMMA"] = mode
        ctx.dkl_dense(dense[:8])
        t = time.perf_counter(); got = ctx.dkl_dense(dense); dt = time.perf_counter() - t
        print(f"synth N={n} G={g} M={m} disable_umma={mode}: rel err {rel(got, want):.3e} nan={np.isnan(got).sum()} time {dt*1e3:.1f} ms", flush=True)
    # large values -> overflow flag -> SIMT rerun
    os.environ["HS_DISABLE_UMMA"] = "0"
    big = dense[:64] * 1e6
    got = ctx.dkl_dense(big)
    print(f"  overflow rerun: rel err {rel(got, want[:64]):.3e}", flush=True)

"""Development probe: hybrid tensor-core / slab CSR path vs slab-only vs oracle."""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import haystack_oracle as ho
from singlecellhaystack_b200 import synth
from singlecellhaystack_b200.device import DeviceContext
def rel(a, b): return float(np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
ctx = DeviceContext(0)
for (n, d, g, m, rho) in [(20000, 10, 100, 600, 0.15), (100000, 20, 100, 1500, 0.1), (5000, 3, 37, 400, 0.3)]:
    x, labels = synth.make_embedding(n, d); xs2, _, _ = synth.scale_coords(x); grid = synth.make_grid(xs2, g)
    p, j, v = synth.make_csr_numpy(m, n, rho, labels)
    ref = ho.density_stage(xs2, grid)
    sp = ho.RowSparse(p=p, j=j, x=v, shape=(m, n))
    t = time.perf_counter(); want = ho.dkl_observed(sp, ref["density"], ref["Q"]); t_or = time.perf_counter() - t
    ctx.density(xs2, grid)
    dens = np.diff(p) / n
    for mode in ("0", "1"):
        os.environ["HS_DISABLE_UMMA"] = mode
        ctx.dkl_csr(p, j, v)
        t = time.perf_counter(); got = ctx.dkl_csr(p, j, v); dt = time.perf_counter() - t
        print(f"N={n} G={g} M={m} rows>=4%: {(dens>=0.04).sum()} disable_umma={mode}: rel err {rel(got, want):.3e} nan={np.isnan(got).sum()} time {dt*1e3:.1f} ms (oracle {t_or:.1f}s)", flush=True)
    os.environ["HS_DISABLE_UMMA"] = "0"
    # out-of-range id in a dense row, unsorted dense row
    j2 = j.copy(); big = int(np.argmax(np.diff(p))); j2[p[big + 1] - 1] = n + 5
    got = ctx.dkl_csr(p, j2, v); print("  poison:", np.isnan(got[big]), np.isnan(got).sum(), flush=True)
    j3 = j.copy(); v3 = v.copy(); a, b = p[big], p[big + 1] - 1
    j3[a], j3[b] = j3[b], j3[a]; v3[a], v3[b] = v3[b], v3[a]
    got = ctx.dkl_csr(p, j3, v3); print("  unsorted rerun rel err:", rel(got, want), flush=True)

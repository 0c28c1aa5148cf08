"""Development probe: device-resident timing of the dense contraction (tensor-core vs SIMT)."""
import os, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from singlecellhaystack_b200 import synth
from singlecellhaystack_b200.device import DeviceContext
n, d, g = int(sys.argv[1]), 50, 100
m = int(sys.argv[2])
dev = torch.device("cuda", 0)
x, labels = synth.make_embedding(n, d); xs, _, _ = synth.scale_coords(x); grid = synth.make_grid(xs, g)
ctx = DeviceContext(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.use_torch_stream(st)
ctx.density(xs, grid)
gen = torch.Generator(device=dev); gen.manual_seed(1)
E = torch.rand((n, m), device=dev, generator=gen)            # (cells, genes) C-order == genes x cells column-major
E = torch.where(E < 0.1, torch.log1p(1 + 20 * E), torch.zeros_like(E)).contiguous()
out = torch.empty(m, dtype=torch.float64, device=dev)
res = {}
for mode in ("0", "1"):
    os.environ["HS_DISABLE_UMMA"] = mode
    for _ in range(2): ctx.dkl_dense_dev(E, m, m, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): ctx.dkl_dense_dev(E, m, m, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    res[mode] = out.cpu().numpy().copy()
    gb = 4.0 * n * m / 1e9
    print(f"N={n} M={m} disable_umma={mode}: {ms:.3f} ms  E stream {gb/ms*1e3:.0f} GB/s  ({2.0*n*m*g/ms/1e9:.1f} useful TFLOP/s)", flush=True)
print("max rel diff umma vs simt:", float(np.max(np.abs(res["0"] - res["1"]) / np.abs(res["1"]))))

"""Development probe: where does the end-to-end (host-pointer) path spend its time?"""
import sys, time
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from singlecellhaystack_b200.device import DeviceContext

sys.argv = ["bench.py"] + sys.argv[1:]
a = bench.parse_args()
dev = torch.device("cuda", 0)
# raw pinned H2D bandwidth
for mb in (64, 1024, 8192):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"pinned H2D {mb} MiB: {mb / 1024 / dt:.1f} GiB/s", flush=True)
    hp = torch.empty(mb << 20, dtype=torch.uint8)
    t = time.perf_counter(); d.copy_(hp); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"pageable H2D {mb} MiB: {mb / 1024 / dt:.1f} GiB/s", flush=True)
    del h, d, hp
xs, labels, grid = bench.make_coords(a)
p, cols, vals = bench.make_csr_device(a, labels, 0, dev)
nnz = int(p[-1])
print("nnz", nnz, flush=True)
t = time.perf_counter()
j_h = torch.empty(nnz, dtype=torch.int32).pin_memory(); j_h.copy_(cols)
x_h = torch.empty(nnz, dtype=torch.float64).pin_memory(); x_h.copy_(vals.double())
torch.cuda.synchronize()
print("pin+D2H", time.perf_counter() - t, j_h.is_pinned(), x_h.is_pinned(), flush=True)
p_np, j_np, x_np = p.cpu().numpy(), j_h.numpy(), x_h.numpy()
ctx = DeviceContext(0)
ctx.enable_timers(True)
for rep in range(3):
    t = time.perf_counter(); ctx.density(xs, grid); t1 = time.perf_counter() - t
    t = time.perf_counter(); d = ctx.dkl_csr(p_np, j_np, x_np); t2 = time.perf_counter() - t
    print(f"rep {rep}: density {t1*1e3:.1f} ms  dkl_csr {t2*1e3:.1f} ms  ({12*nnz/t2/1e9:.1f} GB/s)  timers {ctx.last_timers()}", flush=True)
# pageable
j_p, x_p = np.array(j_np), np.array(x_np)
t = time.perf_counter(); d2 = ctx.dkl_csr(p_np, j_p, x_p); t2 = time.perf_counter() - t
print(f"pageable: dkl_csr {t2*1e3:.1f} ms ({12*nnz/t2/1e9:.1f} GB/s)", np.array_equal(d, d2), flush=True)
# device-resident for comparison
out = torch.empty(a.genes, dtype=torch.float64, device=dev)
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); ctx.dkl_csr_dev(p, cols, vals, out); ctx.synchronize(); t3 = time.perf_counter() - t
    print(f"dev-resident dkl_csr {t3*1e3:.1f} ms", flush=True)

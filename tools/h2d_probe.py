"""development probe: pinned host -> device copy bandwidth and host memcpy bandwidth of the box"""
import time, torch, numpy as np
n = 4 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    print("H2D pinned GB/s", n / (time.perf_counter() - t) / 1e9)
a = np.empty(1 << 30, dtype=np.uint8); b = np.empty_like(a)
a[:] = 1
t = time.perf_counter(); b[:] = a; print("host memcpy 1 thread GB/s (read+write counted once)", a.size / (time.perf_counter() - t) / 1e9)
import threading
parts = 16
def cp(k):
    s = a.size // parts
    b[k * s:(k + 1) * s] = a[k * s:(k + 1) * s]
t = time.perf_counter(); th = [threading.Thread(target=cp, args=(k,)) for k in range(parts)]
[x.start() for x in th]; [x.join() for x in th]
print("host memcpy 16 threads GB/s", a.size / (time.perf_counter() - t) / 1e9)

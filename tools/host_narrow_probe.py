"""Dev probe: host-side fp64 -> fp32 narrowing rate of this box (numpy, one thread) and core count."""
import os, time
import numpy as np
n = 64 << 20
x = np.random.default_rng(0).random(n)
y = np.empty(n, np.float32)
t0 = time.perf_counter()
for _ in range(3):
    np.copyto(y, x, casting="same_kind")
dt = (time.perf_counter() - t0) / 3
print(f"cpus={os.cpu_count()} affinity={len(os.sched_getaffinity(0))} one-thread narrow: {n / dt / 1e9:.2f} Gval/s")

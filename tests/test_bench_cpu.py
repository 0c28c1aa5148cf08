"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the
driver's keys, and the helpers that prepare the synthetic workload are consistent between the numpy and
torch (device) code paths."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_emits_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--cells", "4000", "--genes", "200",
                          "--dims", "5", "--grid-points", "20", "--steps", "1", "--warmup", "0", "--ref-genes", "8",
                          "--cpu-density-cells", "2000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "genes/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]


def test_device_and_host_generators_agree():
    import torch
    sys.path.insert(0, str(ROOT))
    import bench
    from singlecellhaystack_b200 import synth
    sys.argv = ["bench.py", "--cells", "3000", "--genes", "64", "--dims", "4", "--grid-points", "12"]
    a = bench.parse_args()
    xs, labels, grid = bench.make_coords(a)
    assert xs.shape == (3000, 4) and grid.shape == (12, 4) and xs.flags.f_contiguous
    p, cols, vals = bench.make_csr_device(a, labels, 0, torch.device("cpu"))
    pn, jn, xn = synth.make_csr_numpy(64, 3000, a.density, labels)
    np.testing.assert_array_equal(p.numpy(), pn)
    np.testing.assert_array_equal(cols.numpy(), jn)
    np.testing.assert_allclose(vals.numpy(), xn, rtol=1e-7)
    # a second rank's slice is the continuation of the same matrix
    p2, cols2, _ = bench.make_csr_device(a, labels, 64, torch.device("cpu"))
    pn2, jn2, _ = synth.make_csr_numpy(128, 3000, a.density, labels)
    # densities are rescaled to the requested mean over the genes generated, so compare structure only
    assert p2.numel() == 65 and cols2.numel() == int(p2[-1])
    # permutation stand-in: distinct, in range
    gp, pv, pi, r = bench.make_perm_inputs_device(a, pn, vals, np.array([0, 5]), np.array([0, 5]), 0, 6, torch.device("cpu"))
    n0 = int(gp[1])
    blk = pi[:n0 * r].reshape(r, n0).numpy()
    assert blk.min() >= 0 and blk.max() < 3000 and all(len(np.unique(row)) == n0 for row in blk)

"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, and the
product arm refuses to run (non-zero exit, no CPU fallback) when no CUDA device is visible."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line(built):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "particle-steps/sec" and line["unit"] == "particle-steps/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert line["value"] > 1e5 and line["steps"] == 1 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    allc = line["cpu_replicas_all_cores"]
    assert 1 <= allc["cores"] <= 32 and allc["value"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["extrapolated"] is True and line["scaling"] == "weak"
    # both arms print the identical workload string (the driver's same_config check)
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.config_block(2, 1)


def test_every_baseline_config_is_a_bench_line():
    sys.path.insert(0, ROOT)
    import bench
    assert sorted(bench.CONFIGS) == [2, 3, 4, 5]
    for cfg, c in bench.CONFIGS.items():
        spec, u, y = bench.workload(cfg, 8)
        assert u.shape == (8, c["nu"]) and y.shape == (8, c["ny"]) and np.all(np.isfinite(y))
    assert bench.CONFIGS[2]["alg"] == 3 * 4 * 8 + 16 and bench.CONFIGS[4]["alg"] == 5 * 4 * 8 + 72
    assert bench.CONFIGS[5]["alg"] == 3 * 64 * 4 + 16


def test_product_arm_fails_loudly_without_a_gpu(built):
    import ctypes as C
    import llpf_b200 as L
    n = C.c_int()
    if L.load_library().llpf_device_count(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a CUDA device is visible")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stdout + out.stderr)

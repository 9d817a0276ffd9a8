"""bench.py contract pieces that run without a GPU: the reference arm's JSON line and the synthetic generators."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-budget", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "viterbi_gcups" and d["unit"] == "GCUPS"
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_generators_are_deterministic_and_shaped():
    sys.path.insert(0, ROOT)
    import bench
    a = bench.gen_models(4, 200, 1)
    b = bench.gen_models(4, 200, 1)
    assert all(np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]) for x, y in zip(a, b))
    assert a[0][1].shape == (200, 20) and a[0][2].shape == (201, 7)
    assert np.allclose(np.exp(a[0][1]).sum(1), 1) and np.all(a[0][2][1:200, [2, 6]] <= 0)
    r1 = bench.gen_reads(a, 16, 1000, 2)
    r2 = bench.gen_reads(a, 16, 1000, 2)
    assert r1 == r2 and all(len(r) == 1000 and set(r) <= set(b"ACGT") for r in r1)

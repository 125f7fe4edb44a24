"""bench.py contract on the CPU: the reference arm (oracle port on the host cores, the one place besides `cpu_baseline`
where bench.py may execute oracle/) prints ONE JSON line with the keys the driver reads, and the product arm refuses to run
without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT, env=env)


def test_reference_arm_prints_one_contract_line():
    r = run_bench("--impl", "reference", "--resolution", "64", "--denoise-steps", "2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "stamps/sec" and d["unit"] == "stamps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["warmup"] >= 3 and d["steps"] == 1            # W >= 3 warm-up steps whatever the caller asks for
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["ms_per_stamp"] >= d["ms_per_step"]
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "UNet evaluation" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without a GPU")
def test_product_arm_fails_loudly_without_a_gpu():
    r = run_bench("--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-library-baseline", timeout=300)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]  # no metric line of any kind

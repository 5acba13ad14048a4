"""The bench contract that can be checked without a GPU: the reference arm (`bench.py --impl reference`) prints ONE JSON
line with the keys the driver reads, and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import conftest

ROOT = conftest.ROOT


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    out = _run("--impl", "reference", "--workload", "config4", "--ni", "64", "--nj", "36", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rays_per_s" and d["unit"] == "rays/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["config"]["workload"] == "ks_a0.99_4k_wide" and (d["config"]["ni"], d["config"]["nj"]) == (64, 36)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "rays" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_under_torchrun_only_rank0_prints():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--ni", "64",
                          "--nj", "36", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300,
                         cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    if conftest.has_gpu():
        import pytest
        pytest.skip("only meaningful on a machine without a GPU")
    out = _run("--ni", "64", "--nj", "36", "--steps", "1", "--warmup", "0")
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)

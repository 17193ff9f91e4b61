"""bench.py on a machine without a GPU: the reference arm (`--impl reference`, the oracle port on the host cores) prints
ONE JSON line with the contract's keys; the product arm fails loudly instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

from tests import helpers as H

BENCH = os.path.join(H.REPO, "bench.py")


def _run(args, timeout=600):
    return subprocess.run([sys.executable, BENCH] + args, cwd=H.REPO, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "0", "--cpu-per-worker", "2"])
    assert r.returncode == 0, r.stderr[-800:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line, got %d" % len(lines)
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, "key '%s' missing" % k
    assert d["impl"] == "reference" and d["unit"] == "ticks/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("planning ticks/s") and d["steps"] == 2 and d["n_gpus"] == 1
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "ticks/s" and "sample" in cb
    assert cb["ticks_per_s_cpu_1core"] > 0 and cb["ticks_per_s_cpu_allcores"] == cb["value"]
    assert abs(cb["value"] - d["value"]) <= 1e-9 * d["value"] and d["value"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and e["unit"] == "ticks/s"
    assert abs(e["value"] - d["value"]) <= 1e-9 * d["value"]


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-extra"], timeout=300)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")], "no result line without a GPU"
    assert "NVIDIA" in r.stderr or "CUDA" in r.stderr or "cuda" in r.stderr


def test_reference_arm_under_torchrun_prints_once():
    """the driver launches the reference arm like the product arm; rank 0 alone works and prints, the others exit 0."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", BENCH, "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0", "--cpu-per-worker", "2"], cwd=H.REPO, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0

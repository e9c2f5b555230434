"""bench.py's JSON contract, as far as it can be checked without a GPU: the reference arm
(`--impl reference` = the unmodified reference CPU app from oracle/_ref) runs here on a small
sample and must print ONE line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref is not built (no /root/reference on this box)")
def test_reference_arm_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-scale", "14",
                        "--steps", "2", "--warmup", "1"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "edges/s" and d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    # the traversed edges of a BFS from the maximum-degree vertex of an R-MAT graph: nearly all input edges
    assert 0.5 * d["config"]["input_edges"] < d["config"]["traversed_edges"] <= d["config"]["input_edges"]


def test_gpu_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm must not print a bench line."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--scale", "10",
                        "--no-cpu-baseline", "--sweep", "none"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert p.returncode != 0
    assert not [l for l in p.stdout.splitlines() if l.startswith("{") and '"value"' in l]

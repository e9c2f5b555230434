"""world_size-2/3 gloo runs (CPU): the distributed protocol model and the
bench.py launch contract for the reference arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script, *args, port=29611):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), script, *args]
    return subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)


@pytest.mark.parametrize("nproc", [2, 3])
def test_distributed_protocol_model(nproc):
    p = _torchrun(nproc, os.path.join(ROOT, "tests", "dist_model.py"), port=29611 + nproc)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "OK" in p.stdout


def test_reference_arm_under_torchrun_rank0_only():
    """--impl reference: rank 0 alone runs and prints ONE json line; other ranks exit 0."""
    p = _torchrun(2, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                  "--warmup", "1", "--ref-scale", "12", port=29631)   # 3 queries in one reference process
    assert p.returncode == 0, p.stdout[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "edges/s"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0

"""CPU: the multi-rank MPI shim (oracle/ref/shims/mpi.h) that launches the
reference's host code as N rank processes (GL_MPI_NP=N == mpirun -n N):
collectives, large concurrent exchanges, probe threads stopped by a
zero-length self send, communicator dup / split."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe():
    out = os.path.join(tempfile.mkdtemp(), "mpi_shim_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I" + os.path.join(ROOT, "oracle", "ref", "shims"),
                           os.path.join(ROOT, "tests", "mpi_shim_check.cc"), "-o", out])
    return out


@pytest.mark.parametrize("np_", [1, 2, 3, 5])
def test_mpi_shim_ranks(exe, np_):
    env = dict(os.environ, GL_MPI_NP=str(np_), GL_MPI_TIMEOUT_S="60")
    p = subprocess.run([exe], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
    assert p.returncode == 0, p.stdout
    lines = [l for l in p.stdout.splitlines() if l.startswith("rank ")]
    assert len(lines) == np_
    want_sum = np_ * (np_ + 1) // 2
    for l in lines:
        assert "sum=%d " % want_sum in l and "big_ok=1" in l, l


def test_mpi_shim_reports_dead_rank(exe):
    """A rank that dies makes the others give up instead of hanging (non-zero exit)."""
    env = dict(os.environ, GL_MPI_NP="2", GL_MPI_TIMEOUT_S="20", SHIM_CHECK_DIE="1")
    p = subprocess.run([exe], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert p.returncode != 0

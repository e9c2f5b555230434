"""CPU: the two statements of bench.py's synthetic-input definition agree --
oracle/rmat_gen.h (used by the CPU arm, inside oracle/_ref/ref_driver) and the
product's generator (libgrape-lite_b200/csrc/rmat.h via gl_rmat_edges_host)
produce identical edge lists and weights."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import refdriver
from tests.util import pkg

pytestmark = pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref not built")


@pytest.mark.parametrize("scale,seed,wmode", [(10, 1, 0), (13, 7, 1), (12, 3, 2), (1, 1, 0)])
def test_rmat_definitions_agree(scale, seed, wmode):
    m = 16 << scale
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "e.bin")
        subprocess.check_call([refdriver.EXE, "--rmat", "%d,16,%d,%d" % (scale, seed, wmode), "--dump-edges", f])
        raw = np.fromfile(f, dtype=np.int64)
    src, dst, w = pkg().rmat_edges_host(scale, 16, seed, wmode)
    assert np.array_equal(raw[:m], src) and np.array_equal(raw[m:2 * m], dst)
    if wmode:
        assert np.array_equal(raw[2 * m:].view(np.float64), w.astype(np.float64))


def test_ref_driver_maxdeg_source_and_teps_numerator():
    info = refdriver.run_rmat("bfs", 12, 16, 1, repeat=1)
    src, dst, _ = pkg().rmat_edges_host(12, 16, 1, 0)
    deg = np.bincount(np.concatenate([src, dst]), minlength=1 << 12)
    assert info["source"] == int(np.argmax(deg))
    from oracle import pyoracle
    g = pyoracle.Graph(1 << 12, src, dst, None)
    depth, _ = g.bfs(info["source"])
    reached = depth != np.iinfo(np.int64).max
    assert info["traversed_edges"] == int(np.count_nonzero(reached[src]))


@pytest.mark.parametrize("app", ["bfs", "sssp", "wcc", "pagerank"])
def test_opt_variants_agree_with_plain_apps(app):
    """f4: the reference's --opt CPU apps (the stronger CPU baseline) compute the same results."""
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for opt in (False, True):
            f = os.path.join(d, "o%d.txt" % opt)
            refdriver.run_rmat(app, 11, 16, 5, repeat=1, opt=opt, out=f)
            out[opt] = refdriver.parse_output(open(f).read(), int if app in ("bfs", "wcc") else float)
    assert np.array_equal(out[False][0], out[True][0])
    if app == "wcc":
        from tests import golden_io as G
        assert G.same_partition(out[False][1], out[True][1])
    elif app == "pagerank":
        assert np.allclose(out[False][1], out[True][1], rtol=1e-9, atol=0)
    else:
        assert np.array_equal(out[False][1], out[True][1])

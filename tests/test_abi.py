"""CPU-side checks of the C-ABI boundary: the library loads, exports every
symbol include/grape_b200.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

from tests.util import pkg, have_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "grape_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(gl_[a-z0-9_]+)\s*\(", src))
    # typedef'd function-pointer types are not symbols
    names -= {"gl_allreduce_fn"}
    return sorted(names)


def test_library_exports_every_declared_symbol():
    L = pkg().lib()
    missing = [n for n in header_functions() if not hasattr(L, n)]
    assert not missing, missing


def test_python_mirror_lists_the_same_symbols():
    assert sorted(pkg().capi.SYMBOLS) == header_functions()


def test_abi_version():
    assert pkg().lib().gl_abi_version() == 2


def test_struct_sizes_match_the_header():
    # compile a tiny C program against the header and compare sizeof()
    import subprocess, tempfile
    p = pkg()
    code = r'''
#include <stdio.h>
#include "grape_b200.h"
int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(gl_csr_desc), sizeof(gl_frag_desc),
 sizeof(gl_edges_desc), sizeof(gl_rmat_desc), sizeof(gl_frag_info), sizeof(gl_app_config),
 sizeof(gl_query_stats), sizeof(gl_comm_desc)); return 0;}
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(code)
        exe = os.path.join(d, "s")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mirror = [p.capi.CsrDesc, p.capi.FragDesc, p.capi.EdgesDesc, p.capi.RmatDesc, p.capi.FragInfo,
              p.capi.AppConfig, p.capi.QueryStats, p.capi.CommDesc]
    assert sizes == [ctypes.sizeof(m) for m in mirror]


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU error path")
def test_no_gpu_fails_loudly():
    p = pkg()
    with pytest.raises(p.GrapeError) as e:
        p.device_info()
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(p.GrapeError):
        p.Fragment.rmat(8)


def test_host_rmat_generator_is_deterministic_and_in_range():
    import numpy as np
    p = pkg()
    s1, d1, w1 = p.rmat_edges_host(10, 16, seed=3, weight_mode=1)
    s2, d2, w2 = p.rmat_edges_host(10, 16, seed=3, weight_mode=1)
    assert np.array_equal(s1, s2) and np.array_equal(d1, d2) and np.array_equal(w1, w2)
    assert s1.min() >= 0 and s1.max() < 1024 and d1.max() < 1024
    assert w1.min() >= 1 and w1.max() <= 255 and np.all(w1 == np.round(w1))
    # chunked generation equals whole generation
    s3, d3, _ = p.rmat_edges_host(10, 16, seed=3, weight_mode=0, first=100, count=50)
    assert np.array_equal(s3, s1[100:150]) and np.array_equal(d3, d1[100:150])
    # skewed: the maximum degree is far above the mean
    deg = np.bincount(np.concatenate([s1, d1]), minlength=1024)
    assert deg.max() > 8 * deg.mean()

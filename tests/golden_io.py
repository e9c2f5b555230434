"""Loaders for the committed golden fixtures (tests/golden/*.gz) and the
reference's verifier rules (misc/app_tests.sh:6-40, misc/eps_check.cc,
misc/wcc_check.cc)."""
import gzip
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _open(name):
    return gzip.open(os.path.join(GOLDEN, name + ".gz"), "rt")


def load_p2p31():
    """Returns (oids[int64 sorted], src, dst, w[float64]) of dataset/p2p-31.{v,e}."""
    with _open("p2p-31.v") as f:
        oids = np.array([int(l.split()[0]) for l in f if l.strip()], dtype=np.int64)
    e = np.loadtxt(_open("p2p-31.e"), dtype=np.float64)
    src = e[:, 0].astype(np.int64)
    dst = e[:, 1].astype(np.int64)
    w = e[:, 2].copy()
    return np.sort(oids), src, dst, w


def golden_lines(name):
    with _open(name) as f:
        return f.read()


def golden_pairs(name):
    """-> list of (oid:int, value:str)"""
    out = []
    with _open(name) as f:
        for l in f:
            a, b = l.split()
            out.append((int(a), b))
    return out


def fmt_sci(x):
    """std::scientific << std::setprecision(15) (sssp_context.h:68)."""
    return "%.15e" % x


def render(oids, values):
    """`oid value` lines sorted by oid (misc/app_tests.sh:7 `sort -k1n`)."""
    return "".join("%d %s\n" % (o, v) for o, v in zip(oids, values))


def eps_check(got, want, eps=1e-4):
    """misc/eps_check.cc:24,48-58: fabs(v1-v2) < eps*v1 (relative)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return bool(np.all(np.abs(got - want) <= eps * np.abs(want) + 1e-300))


def same_partition(a, b):
    """misc/wcc_check.cc:36-71: labels must be in bijection."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return False
    f, g = {}, {}
    for x, y in zip(a.tolist(), b.tolist()):
        if f.setdefault(x, y) != y or g.setdefault(y, x) != x:
            return False
    return True

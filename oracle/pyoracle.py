"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_graph_build.restype = C.c_void_p
        L.orc_graph_build.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int]
        L.orc_graph_free.argtypes = [C.c_void_p]
        L.orc_graph_n.restype = C.c_int64
        L.orc_graph_n.argtypes = [C.c_void_p]
        L.orc_graph_entries.restype = C.c_uint64
        L.orc_graph_entries.argtypes = [C.c_void_p]
        for f in ("orc_graph_oids", "orc_graph_rp", "orc_graph_col", "orc_graph_w"):
            getattr(L, f).restype = C.c_void_p
        L.orc_graph_oids.argtypes = [C.c_void_p]
        L.orc_graph_rp.argtypes = [C.c_void_p, C.c_int]
        L.orc_graph_col.argtypes = [C.c_void_p, C.c_int]
        L.orc_graph_w.argtypes = [C.c_void_p, C.c_int]
        L.orc_graph_index_of.restype = C.c_int64
        L.orc_graph_index_of.argtypes = [C.c_void_p, C.c_int64]
        L.orc_graph_max_degree_vertex.restype = C.c_int64
        L.orc_graph_max_degree_vertex.argtypes = [C.c_void_p]
        L.orc_bfs.restype = C.c_int
        L.orc_bfs.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_sssp.restype = C.c_int
        L.orc_sssp.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_wcc.restype = C.c_int
        L.orc_wcc.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_pagerank.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_void_p]
        L.orc_cdlp.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_lcc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Graph:
    """Whole-graph (1-fragment) CPU graph, vertices indexed in ascending-oid order."""

    def __init__(self, n, src, dst, w=None, directed=False, oids=None):
        L = lib()
        src = np.ascontiguousarray(src, dtype=np.int64)
        dst = np.ascontiguousarray(dst, dtype=np.int64)
        w = None if w is None else np.ascontiguousarray(w, dtype=np.float64)
        oids = None if oids is None else np.ascontiguousarray(oids, dtype=np.int64)
        self.h = L.orc_graph_build(n, _p(oids), len(src), _p(src), _p(dst), _p(w),
                                   1 if directed else 0)
        self.n = n
        self.directed = directed
        self.entries = L.orc_graph_entries(self.h)

    def __del__(self):
        if getattr(self, "h", None) and _LIB is not None:
            _LIB.orc_graph_free(self.h)
            self.h = None

    @property
    def oids(self):
        p = lib().orc_graph_oids(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int64)), (self.n,)).copy()

    def csr(self, incoming=False):
        L = lib()
        i = 1 if incoming else 0
        rp = np.ctypeslib.as_array(C.cast(L.orc_graph_rp(self.h, i), C.POINTER(C.c_uint64)),
                                   (self.n + 1,)).copy()
        m = int(rp[-1])
        col = np.ctypeslib.as_array(C.cast(L.orc_graph_col(self.h, i), C.POINTER(C.c_uint32)),
                                    (max(m, 1),))[:m].copy()
        wp = L.orc_graph_w(self.h, i)
        w = None
        if wp:
            w = np.ctypeslib.as_array(C.cast(wp, C.POINTER(C.c_double)), (max(m, 1),))[:m].copy()
        return rp, col, w

    def index_of(self, oid):
        return lib().orc_graph_index_of(self.h, int(oid))

    def max_degree_vertex(self):
        return lib().orc_graph_max_degree_vertex(self.h)

    def bfs(self, source_index):
        out = np.empty(self.n, dtype=np.int64)
        steps = lib().orc_bfs(self.h, int(source_index), _p(out))
        return out, steps

    def sssp(self, source_index):
        out = np.empty(self.n, dtype=np.float64)
        steps = lib().orc_sssp(self.h, int(source_index), _p(out))
        return out, steps

    def wcc(self):
        out = np.empty(self.n, dtype=np.uint32)
        steps = lib().orc_wcc(self.h, _p(out))
        return out, steps

    def pagerank(self, delta=0.85, max_round=10, mode=0):
        out = np.empty(self.n, dtype=np.float64)
        lib().orc_pagerank(self.h, float(delta), int(max_round), int(mode), _p(out))
        return out

    def cdlp(self, max_round=10):
        out = np.empty(self.n, dtype=np.int64)
        lib().orc_cdlp(self.h, int(max_round), _p(out))
        return out

    def lcc(self):
        out = np.empty(self.n, dtype=np.float64)
        tri = np.empty(self.n, dtype=np.int64)
        lib().orc_lcc(self.h, _p(out), _p(tri))
        return out, tri


def num_threads():
    return lib().orc_num_threads()

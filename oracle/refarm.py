"""CPU arm of bench.py (TEST/MEASUREMENT INFRASTRUCTURE, not product code).

Times the reference's CPU implementation of the hot path on the host cores:
oracle/_ref (the UNMODIFIED reference apps compiled against functional shims,
kind="reference") when it was built, else the oracle port (kind="port").
The graph is the same generator/seed as the GPU arm, at a bounded scale.
"""
import importlib
import os
import time

import numpy as np

from oracle import pyoracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def _graph(scale, edgefactor, seed, weighted):
    key = (scale, edgefactor, seed, weighted)
    if key not in _cache:
        pkg = importlib.import_module("libgrape-lite_b200")
        src, dst, w = pkg.rmat_edges_host(scale, edgefactor, seed, 1 if weighted else 0)
        g = pyoracle.Graph(1 << scale, src, dst, None if w is None else w.astype(np.float64))
        _cache.clear()
        _cache[key] = (g, src, dst, w)
    return _cache[key]


def ref_binary():
    p = os.path.join(_HERE, "_ref", "ref_driver")
    return p if os.path.exists(p) else None


def run(app, scale, edgefactor=16, seed=1, reuse=False, repeat=1, keep=1, opt=False):
    exe = ref_binary()
    if exe is not None:
        from oracle import refdriver
        return refdriver.run(exe, app, scale, edgefactor, seed, repeat=repeat, keep=keep, opt=opt)
    g, src, dst, w = _graph(scale, edgefactor, seed, app == "sssp")
    source = g.max_degree_vertex()
    rp, _, _ = g.csr()
    deg = np.diff(rp).astype(np.int64)
    t0 = time.perf_counter()
    if app == "bfs":
        res, _ = g.bfs(source)
        reached = res != np.iinfo(np.int64).max
        edges = int(deg[reached].sum()) // 2
    elif app == "sssp":
        res, _ = g.sssp(source)
        reached = res < 1e300
        edges = int(deg[reached].sum()) // 2
    elif app == "wcc":
        g.wcc()
        edges = edgefactor << scale
    elif app == "pagerank":
        g.pagerank(0.85, 10, 0)
        edges = (edgefactor << scale) * 10
    elif app == "cdlp":
        g.cdlp(10)
        edges = (edgefactor << scale) * 10
    else:
        g.lcc()
        edges = edgefactor << scale
    dt = time.perf_counter() - t0
    cores = 1 if app in ("bfs", "sssp", "wcc") else pyoracle.num_threads()
    return {"value": edges / dt, "unit": "edges/s", "ms": dt * 1e3, "cores": cores, "kind": "port",
            "sample": "%s on R-MAT scale-%d (same generator/seed as the GPU arm), one query, oracle port"
                      % (app.upper(), scale)}

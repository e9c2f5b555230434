"""Runs oracle/_ref/ref_driver — the UNMODIFIED reference CPU apps compiled
against functional MPI/glog shims (oracle/ref/).  TEST / MEASUREMENT ONLY."""
import importlib
import json
import os
import struct
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(_HERE, "_ref", "ref_driver")


def available():
    return os.path.exists(EXE)


def write_graph(path, n, src, dst, w=None, oids=None):
    src = np.ascontiguousarray(src, dtype=np.int64)
    dst = np.ascontiguousarray(dst, dtype=np.int64)
    with open(path, "wb") as f:
        f.write(b"GRB1")
        f.write(struct.pack("<qqii", n, len(src), 0 if w is None else 1, 0 if oids is None else 1))
        if oids is not None:
            f.write(np.ascontiguousarray(oids, dtype=np.int64).tobytes())
        f.write(src.tobytes())
        f.write(dst.tobytes())
        if w is not None:
            f.write(np.ascontiguousarray(w, dtype=np.float64).tobytes())


def run_app(app, graph_path, directed=False, source=0, pr_d=0.85, mr=10, threads=0, repeat=1,
            want_output=True, timeout=3600):
    """-> (info dict, text output or None)"""
    out = None
    cmd = [EXE, "--app", app, "--graph", graph_path, "--directed", "1" if directed else "0",
           "--source", str(int(source)), "--pr_d", repr(float(pr_d)), "--mr", str(int(mr)),
           "--threads", str(int(threads)), "--repeat", str(int(repeat))]
    with tempfile.TemporaryDirectory() as d:
        if want_output:
            out = os.path.join(d, "out.txt")
            cmd += ["--out", out]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError("ref_driver failed (%d): %s" % (p.returncode, p.stderr[-2000:]))
        info = json.loads(p.stdout.strip().splitlines()[-1])
        text = open(out).read() if want_output else None
    return info, text


def parse_output(text, dtype=float):
    """`oid value` lines -> (oids sorted, values in oid order) — the reference's
    verifiers sort by the first column (misc/app_tests.sh:7)."""
    o, v = [], []
    for line in text.splitlines():
        a, b = line.split()
        o.append(int(a))
        if dtype is float:
            v.append(np.finfo(np.float64).max if b == "infinity" else float(b))
        else:
            v.append(int(b))
    o = np.array(o, dtype=np.int64)
    v = np.array(v, dtype=np.float64 if dtype is float else np.int64)
    k = np.argsort(o, kind="stable")
    return o[k], v[k]


def run_rmat(app, scale, edgefactor=16, seed=1, repeat=1, opt=False, threads=0, timeout=3600, out=None):
    """Runs the reference CPU app on bench.py's synthetic input, generated INSIDE ref_driver
    (oracle/rmat_gen.h) -- nothing of the product is loaded.  -> info dict of ref_driver."""
    wmode = 1 if app == "sssp" else 0
    cmd = [EXE, "--app", app, "--rmat", "%d,%d,%d,%d" % (scale, edgefactor, seed, wmode), "--source", "maxdeg",
           "--repeat", str(int(repeat)), "--threads", str(int(threads)), "--opt", "1" if opt else "0"]
    if out:
        cmd += ["--out", out]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError("ref_driver failed (%d): %s" % (p.returncode, p.stderr[-2000:]))
    return json.loads(p.stdout.strip().splitlines()[-1])


def run(exe, app, scale, edgefactor=16, seed=1, repeat=1, keep=1, opt=False):
    """bench.py CPU arm: `repeat` queries of the reference CPU app in ONE process
    (graph generated and loaded once), all host threads; the mean of the last `keep` is reported."""
    info = run_rmat(app, scale, edgefactor, seed, repeat=repeat, opt=opt)
    ms = float(np.mean(info["query_ms"][-keep:]))
    m = int(info["edges"])
    if app in ("bfs", "sssp"):
        edges = int(info["traversed_edges"])     # Graph500 numerator: input edges with a reached source endpoint
    elif app in ("pagerank", "cdlp"):
        edges = m * 10
    else:
        edges = m
    return {"value": edges / (ms * 1e-3), "unit": "edges/s", "ms": ms, "cores": int(info["threads"]),
            "kind": "reference", "all_ms": [float(x) for x in info["query_ms"]], "load_s": info["load_s"],
            "scale": scale, "opt": bool(opt), "source_oid": int(info["source"]), "traversed_edges": edges,
            "sample": "%s on R-MAT scale-%d edgefactor-%d seed-%d (the GPU arm's generator, restated in "
                      "oracle/rmat_gen.h), Query() of the unmodified reference CPU app%s (ParallelEngine, "
                      "%d threads of %d)"
                      % (app.upper(), scale, edgefactor, seed, " (--opt variant)" if opt else "",
                         info["threads"], info["hardware_concurrency"])}

#!/usr/bin/env python
"""Times the reference's OWN GPU code (oracle/_ref/ref_gpu_driver: grape/cuda/** compiled for sm_100a,
see oracle/ref/patch_gpu_reference.py) next to the reference's unchanged app headers on THIS engine
(compat/_build/run_compat_app) on the same synthetic input, per load-balancing mode, and checks the
reference build against the golden files first.  MEASUREMENT INFRASTRUCTURE (run under gpurun).

usage: ref_gpu_compare.py [--scale 22] [--repeat 3] [--out gpurun_out/ref_gpu_compare.json]"""
import argparse
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_driver")
OURS = os.path.join(ROOT, "compat", "_build", "run_compat_app")


def run(exe, args, timeout=3600):
    p = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    if p.returncode != 0:
        return None, p.stderr[-800:]
    return [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")], ""


def golden_check():
    from tests import golden_io as G
    import numpy as np
    d = tempfile.mkdtemp()
    for f in ("p2p-31.e", "p2p-31.v"):
        with gzip.open(os.path.join(G.GOLDEN, f + ".gz"), "rb") as i, open(os.path.join(d, f), "wb") as o:
            shutil.copyfileobj(i, o)
    res = {}
    for app, extra, name in (("bfs", ["--bfs_source", "6"], "p2p-31-BFS"), ("sssp", ["--sssp_source", "6"], "p2p-31-SSSP"),
                             ("pagerank", ["--pr_mr", "10", "--pr_d", "0.85"], "p2p-31-PR"), ("wcc", [], "p2p-31-WCC")):
        out = tempfile.mkdtemp()
        rows, err = run(REF, ["--application", app, "--efile", os.path.join(d, "p2p-31.e"), "--vfile", os.path.join(d, "p2p-31.v"),
                              "--out_prefix", out, "--lb", "cm"] + extra, timeout=600)
        if rows is None:
            res[app] = "FAILED: " + err[-200:]
            continue
        lines = sorted(open(os.path.join(out, "result_frag_0")).read().splitlines(), key=lambda l: int(l.split()[0]))
        if app in ("bfs", "sssp"):
            ok = "\n".join(lines) + "\n" == G.golden_lines(name)
        elif app == "pagerank":
            got = np.array([float(l.split()[1]) for l in lines])
            want = np.array([float(v) for _, v in G.golden_pairs(name)])
            ok = G.eps_check(got, want, 1e-4)
        else:
            got = np.array([int(l.split()[1]) for l in lines])
            want = np.array([int(v) for _, v in G.golden_pairs(name)])
            ok = G.same_partition(got, want)
        res[app] = "golden OK" if ok else "MISMATCH"
        shutil.rmtree(out, ignore_errors=True)
    shutil.rmtree(d, ignore_errors=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--apps", default="bfs,sssp,pagerank,wcc")
    ap.add_argument("--lbs", default="cm,cta,wm,strict,none")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_gpu_compare.json"))
    a = ap.parse_args()
    report = {"scale": a.scale, "repeat": a.repeat, "reference_build_parity": golden_check(), "rows": []}
    print("reference GPU build vs golden files:", report["reference_build_parity"], flush=True)
    for app in a.apps.split(","):
        wmode = 1 if app == "sssp" else 0
        common = ["--application", app, "--rmat", "%d,16,1,%d" % (a.scale, wmode), "--lb", a.lbs, "--repeat", str(a.repeat),
                  "--bfs_source", "maxdeg", "--sssp_source", "maxdeg", "--pr_mr", "10", "--pr_d", "0.85",
                  "--wl_in", "1", "--wl_out_local", "1", "--wl_out_remote", "1"]
        for name, exe in (("reference-gpu", REF), ("b200-compat", OURS)):
            rows, err = run(exe, common)
            if rows is None:
                print(name, app, "FAILED", err[-300:], flush=True)
                report["rows"].append({"impl": name, "app": app, "error": err[-300:]})
                continue
            by_lb = {}
            for r in rows:
                by_lb.setdefault(r["lb"], []).append(r["query_ms"])
            for lb, ms in by_lb.items():
                best = min(ms[1:]) if len(ms) > 1 else ms[0]     # first repetition = warm-up
                report["rows"].append({"impl": name, "app": app, "lb": lb, "query_ms": best, "all_ms": ms,
                                       "load_s": rows[0]["load_s"]})
                print("%-14s %-9s lb=%-7s %10.3f ms  %s" % (name, app, lb, best, ms), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(report, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — TEPS / ms-per-superstep of the PIE hot path on B200.

One "step" = one whole query (PEval + IncEval supersteps) of the app over the
resident fragment(s).  Headline workload = BASELINE.json configs[1] (C2):
BFS on R-MAT scale-24 (edgefactor 16, undirected), 1 x B200; with N GPUs the
graph grows with N (weak scaling: scale 24 + log2 N, one fragment per GPU).

The same line carries, under config.apps, the other BASELINE configs measured
in the same run, each with its own parity check:
  C1  SSSP on the bundled p2p-31 (tests/golden fixture), GPU next to the
      reference's own CPU app (oracle/_ref) on the same input
  C3  PageRank, 10 rounds, d = 0.85, R-MAT scale-24, push IncEval   (N = 1)
  C4  WCC on R-MAT scale 23 + log2 N (= 26 on 8 GPUs), N fragments
  C5  SSSP (near/far) on weighted R-MAT scale 23 + log2 N, N fragments
(at N = 1 the WCC / SSSP entries run the scale-24 graph).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--app bfs] [--scale S]
  python bench.py --impl reference ...     # the CPU arm: the reference's own apps

Prints ONE JSON line (see the driver contract in the task statement).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MAPPING = ("bitmap frontier -> device-ticketed 8192-vertex super tiles walked CTA-cooperatively (the reference's "
           "'cm' mapping) + rows > 1024 entries cut into 1024-entry work items for all SMs; gl_app_config.lb only "
           "selects the kernel of gl_edge_scan_queue / the compat ForEachEdge, not of the built-in apps")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------- workload ---
STATE_BYTES = {"bfs": 8, "sssp": 8, "wcc": 8, "wcc_opt": 8, "pagerank": 24, "cdlp": 16, "lcc": 8}
DTYPE = {"bfs": "u32", "sssp": "f32", "wcc": "u32", "wcc_opt": "u32", "pagerank": "f64", "cdlp": "int64", "lcc": "u32"}


def alg_bytes(app, entries, frontier, touched, weighted):
    """B_alg = m_s*(4+w) + n_f*12 + n_t*(s_r+s_w)."""
    w = 4 if weighted else 0
    return entries * (4 + w) + frontier * 12 + touched * STATE_BYTES[app]


class Group:
    """torch.distributed plumbing of one run (rank / world, barrier, small reductions)."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if args.gpus != self.world and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
        # GL_ONE_DEVICE=1 (functional check of the N > 1 path on a 1-GPU box, never a measurement):
        # every rank's fragment on cuda:0, contexts time-slice the GPU, rendezvous over gloo
        self.one_device = os.environ.get("GL_ONE_DEVICE", "0") == "1"
        if self.one_device:
            self.local = 0
        torch.cuda.set_device(self.local)
        self.dist = None
        self.tdev = "cpu" if self.one_device else "cuda"
        if self.world > 1:
            import torch.distributed as dist
            if self.one_device:
                dist.init_process_group("gloo")
            else:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op="sum"):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.tdev)
        if self.dist:
            self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "max": self.dist.ReduceOp.MAX}[op])
        return [float(x) for x in t]

    def reduce_i64(self, vals, op="sum"):
        t = self.torch.tensor(vals, dtype=self.torch.int64, device=self.tdev)
        if self.dist:
            self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "max": self.dist.ReduceOp.MAX}[op])
        return [int(x) for x in t]

    def gather_i64(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.int64, device=self.tdev)
        if not self.dist:
            return [[int(x) for x in t]]
        allt = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(allt, t)
        return [[int(x) for x in a] for a in allt]

    def free_bytes(self):
        return int(self.torch.cuda.mem_get_info()[0])

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


def global_source(G, frag, n):
    """max-degree vertex of the whole graph, ties -> smallest oid"""
    lid, deg = frag.max_degree_vertex()
    chunk = (n + G.world - 1) // G.world
    rows = G.gather_i64([deg, -(G.rank * chunk + lid)])
    best = max((r[0], r[1]) for r in rows)
    return -best[1]


def reached_entries(pkg, frag, result, app):
    """CSR entries of the reached inner vertices (Graph500 numerator = half of the sum over ranks)."""
    rp = np.zeros(frag.ivnum + 1, dtype=np.uint64)
    pkg.check(pkg.lib().gl_frag_copy_csr(frag.h, 0, pkg.capi._p(rp), None, None))
    deg = np.diff(rp).astype(np.int64)
    if app == "bfs":
        reached = result != np.iinfo(np.int64).max
    elif app == "sssp":
        reached = result < 1e300
    else:
        reached = np.ones(len(result), dtype=bool)
    return int(deg[reached].sum())


def measure_app(G, pkg, frag, comm, kind, cfg, steps, warmup, flush, edges_fn, weighted=False, e2e=True):
    """Creates the app, runs warm-up + `steps` device-timed queries (+ `steps` end-to-end ones) and
    returns (dict, app, last result array).  Times are the MAX over ranks."""
    torch = G.torch
    free0 = G.free_bytes()
    t0 = time.perf_counter()
    app = pkg.App(kind, frag, comm, **cfg)
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t0) * 1e3
    extra_bytes = max(0, free0 - G.free_bytes())
    pinned = pkg.PinnedBuffer(8 * max(frag.ivnum, 1))
    out = pinned.array(pkg.capi.RESULT_DTYPE[app.kind], frag.ivnum)
    st = app.query()
    res = app.result(out).copy()
    edges = edges_fn(res)
    G.barrier()
    # warm-up steps in the cadence of the timed ones (flush, barrier, query) and directly in front of them:
    # a pause between warm-up and timed region (result copy, edge count) lets the NVLink links and the
    # clocks fall back to their idle state, and the first timed multi-GPU query then paid ~190 us for it
    for _ in range(max(warmup, 1)):
        flush.zero_()
        G.barrier()
        st = app.query()
    t_dev, launches, stats = 0.0, 0, []
    for _ in range(steps):
        flush.zero_()                      # L2 flush between timed iterations
        G.barrier()                        # ranks enter the query together (outside the event-timed interval)
        st = app.query()
        t_dev += st.query_ms
        launches += st.kernel_launches
        n = st.n_steps
        stats.append(dict(supersteps=st.supersteps, ms=[st.step_ms[i] for i in range(n)],
                          entries=[int(st.step_entries[i]) for i in range(n)],
                          frontier=[int(st.step_frontier[i]) for i in range(n)],
                          mode=[int(st.step_mode[i]) for i in range(n)],
                          scanned=int(st.entries_scanned), nf=int(st.frontier_vertices),
                          touched=int(st.touched_vertices)))
    G.barrier()
    t_e2e = 0.0
    if e2e:
        for _ in range(steps):
            flush.zero_()
            G.barrier()
            t1 = time.perf_counter()
            app.query()
            app.result(out)                # D2H of the step's result into pinned host memory
            G.barrier()
            t_e2e += time.perf_counter() - t1
    t_dev_ms, t_e2e_ms = G.reduce([t_dev, t_e2e * 1e3], "max")
    ms = t_dev_ms / steps
    last = stats[-1]
    # whole-job algorithmic bytes of one query (sum over ranks)
    scanned, nf, touched = G.reduce_i64([last["scanned"], last["nf"], last["touched"]])
    whole_balg = alg_bytes(kind, scanned, nf, touched, weighted)
    peak, peak_src = load_peaks()
    # dominant superstep of THIS rank (largest mean time over the timed steps)
    nst = len(last["ms"])
    same = [s for s in stats if len(s["ms"]) == nst]
    agg_ms = np.mean([s["ms"] for s in same], axis=0) if nst else np.zeros(0)
    agg_ent = np.mean([s["entries"] for s in same], axis=0) if nst else np.zeros(0)
    agg_fr = np.mean([s["frontier"] for s in same], axis=0) if nst else np.zeros(0)
    dom = int(np.argmax(agg_ms)) if nst else 0
    r = dict(ms_per_query=ms, teps=edges / (ms * 1e-3) if ms > 0 else 0.0, traversed_edges=int(edges),
             supersteps=int(last["supersteps"]),
             ms_per_superstep=[round(float(x), 4) for x in last["ms"]],
             superstep_mode=last["mode"], superstep_entries=last["entries"],
             entries_scanned=int(scanned), alg_bytes=int(whole_balg),
             whole_query_gbs=whole_balg / (ms * 1e-3) / 1e9 / G.world if ms > 0 else 0.0,
             app_setup_ms=round(setup_ms, 2), app_extra_bytes=int(extra_bytes), gpu_launches=int(launches),
             e2e_ms_per_query=t_e2e_ms / steps if e2e else None,
             d2h_bytes_per_step=int(out.nbytes))
    r["frac_whole_query"] = r["whole_query_gbs"] / peak
    if nst:
        touched_dom = last["touched"] * (agg_ent[dom] / max(float(sum(agg_ent)), 1.0))
        b_dom = alg_bytes(kind, float(agg_ent[dom]), float(agg_fr[dom]), touched_dom, weighted)
        r["dominant_superstep"] = dict(index=dom, mode=int(last["mode"][dom]), ms=float(agg_ms[dom]),
                                       alg_bytes=float(b_dom),
                                       gbs=b_dom / (float(agg_ms[dom]) * 1e-3) / 1e9 if agg_ms[dom] > 0 else 0.0)
        r["dominant_superstep"]["frac"] = r["dominant_superstep"]["gbs"] / peak
    r["_peak"] = (peak, peak_src)
    return r, app, res


# ------------------------------------------------------------------ sweeps ---
def sweep_apps(G, pkg, args, flush, frag_bfs, comm_bfs, scale_bfs, res_bfs):
    """config.apps: the other BASELINE configs measured in this run, each with a parity check."""
    apps = {}
    steps, warmup = max(2, min(args.steps, 3)), 1
    world = G.world
    ef = args.edgefactor
    scale = scale_bfs if world == 1 else 23 + int(np.log2(world))
    n = 1 << scale
    m_in = ef << scale
    if world == 1:
        frag_u, comm_u = frag_bfs, comm_bfs
    else:
        frag_u = pkg.Fragment.rmat(scale, ef, seed=args.seed, weight_mode=0, fid=G.rank, fnum=world)
        gdist = importlib.import_module("libgrape-lite_b200.dist")
        comm_u = gdist.make_comm(G.rank, world, frag_u.ivnum, item_bytes=16)
    tag = "R-MAT scale-%d ef-%d, %d fragment(s)" % (scale, ef, world)

    def clean(r):
        r = {k: v for k, v in r.items() if not k.startswith("_")}
        r.pop("superstep_entries", None)
        return r

    # ---- C4: WCC (label propagation, the reference's wcc.h) + wcc_opt; parity: the two algorithms agree
    r1, a1, lab1 = measure_app(G, pkg, frag_u, comm_u, "wcc", {}, steps, warmup, flush, lambda _: m_in, e2e=False)
    a1.close()
    r2, a2, lab2 = measure_app(G, pkg, frag_u, comm_u, "wcc_opt", {}, steps, warmup, flush, lambda _: m_in, e2e=False)
    a2.close()
    same = int(np.array_equal(lab1, lab2))
    same = G.reduce_i64([1 - same])[0] == 0
    r1 = clean(r1)
    r1.update(config="C4 WCC, " + tag, parity="labels(label propagation) == labels(union-find wcc_opt): %s" % same,
              parity_ok=bool(same), wcc_opt=dict(ms_per_query=r2["ms_per_query"], teps=r2["teps"],
                                                 supersteps=r2["supersteps"], app_setup_ms=r2["app_setup_ms"]))
    apps["wcc"] = r1
    # ---- C3: PageRank push, 10 rounds; parity: sum = 1 and push == pull (1e-6)
    if world == 1 or args.sweep == "all":
        cfgp = dict(pr_delta=0.85, max_round=10)
        r3, a3, pr = measure_app(G, pkg, frag_u, comm_u, "pagerank", dict(cfgp, pr_pull=0), steps, warmup, flush,
                                 lambda _: m_in * 10, e2e=False)
        a3.close()
        # the deterministic pull formulation (single fragment: f32 contributions, hub values in shared memory)
        cfg_pull = dict(cfgp, pr_pull=1)
        if world == 1:
            cfg_pull["reserved"] = {5: 1}
        rp_, ap, pull = measure_app(G, pkg, frag_u, comm_u, "pagerank", cfg_pull, 2, 1, flush, lambda _: m_in * 10, e2e=False)
        ap.close()
        tot = G.reduce([float(pr.sum())])[0]
        err = G.reduce([float(np.max(np.abs(pr - pull) / pull))], "max")[0]
        r3 = clean(r3)
        rounds = [x for x in r3["ms_per_superstep"] if x > 0.05]
        prounds = [x for x in rp_["ms_per_superstep"] if x > 0.05]
        r3.update(config="C3 PageRank 10 rounds d=0.85 push IncEval (f64 atomics), " + tag,
                  ms_per_round=float(np.median(rounds)) if rounds else None,
                  pull_variant=dict(what="pull sweep, f32 contributions, f64 sums" + (", hub values in shared memory (k_pr_pull_hub)" if world == 1 else ""),
                                    ms_per_query=rp_["ms_per_query"], ms_per_round=float(np.median(prounds)) if prounds else None,
                                    frac_whole_query=rp_["frac_whole_query"]),
                  parity="sum(rank) = %.12f; max rel |push - pull| = %.2e (bar 1e-6)" % (tot, err),
                  parity_ok=bool(abs(tot - 1.0) < 1e-9 and err < 1e-6))
        apps["pagerank"] = r3
    if world > 1:
        comm_u.close()
        frag_u.close()
    # ---- C5: SSSP near/far on the weighted graph; parity: the f32 run == the f64 run (integer weights: exact)
    frag_w = pkg.Fragment.rmat(scale, ef, seed=args.seed, weight_mode=1, fid=G.rank, fnum=world)
    comm_w = None
    if world > 1:
        gdist = importlib.import_module("libgrape-lite_b200.dist")
        comm_w = gdist.make_comm(G.rank, world, frag_w.ivnum, item_bytes=16)
    src = global_source(G, frag_w, n)

    def sssp_edges(res):
        return G.reduce_i64([reached_entries(pkg, frag_w, res, "sssp")])[0] // 2

    r4, a4, d32 = measure_app(G, pkg, frag_w, comm_w, "sssp", dict(source_oid=int(src)), steps, warmup, flush,
                              sssp_edges, weighted=True, e2e=False)
    a4.close()
    a64 = pkg.App("sssp", frag_w, comm_w, source_oid=int(src), sssp_f64=1)
    a64.query()
    d64 = a64.result()
    a64.close()
    same = G.reduce_i64([0 if np.array_equal(d32, d64) else 1])[0] == 0
    r4 = clean(r4)
    r4.update(config="C5 SSSP near/far (delta-stepping IncEval), integer weights 1..255, " + tag,
              source_oid=int(src), parity="f32 distances == f64 distances bit for bit: %s" % same, parity_ok=bool(same))
    apps["sssp"] = r4
    if comm_w:
        comm_w.close()
    frag_w.close()
    # ---- C1: SSSP on the bundled p2p-31, one fragment, next to the reference's CPU app on the same input
    if G.rank == 0:
        try:
            apps["sssp_p2p31"] = p2p31_sssp(G, pkg, flush)
        except Exception as e:  # the fixture or oracle/_ref may be missing on a stripped box
            apps["sssp_p2p31"] = {"config": "C1 SSSP p2p-31", "unavailable": repr(e)[:300]}
    G.barrier()
    return apps


def p2p31_sssp(G, pkg, flush):
    from tests import golden_io as GO
    oids, src, dst, w = GO.load_p2p31()
    n = len(oids)
    frag = pkg.Fragment.from_edges(n, src, dst, w, oids=oids, directed=False, w_dtype=np.float32)
    app = pkg.App("sssp", frag, source_oid=6)
    pinned = pkg.PinnedBuffer(8 * n)
    out = pinned.array(np.float64, n)
    for _ in range(3):
        app.query()
    dev, e2e = [], []
    for _ in range(10):
        flush.zero_()
        G.torch.cuda.synchronize()
        t1 = time.perf_counter()
        st = app.query()
        app.result(out)
        e2e.append((time.perf_counter() - t1) * 1e3)
        dev.append(st.query_ms)
    res = app.result().copy()
    text = GO.render(oids, ["infinity" if d > 1e300 else GO.fmt_sci(d) for d in res])
    ok = text == GO.golden_lines("p2p-31-SSSP")
    r = {"config": "C1 SSSP on dataset/p2p-31 (62586 vertices / 147892 edges), 1 fragment, source 6",
         "gpu_ms_per_query": float(np.median(dev)), "gpu_e2e_ms_per_query": float(np.median(e2e)),
         "supersteps": int(st.supersteps), "ms_per_superstep": round(float(np.median(dev)) / max(st.supersteps, 1), 4),
         "traversed_edges": 147892, "teps": 147892 / (float(np.median(dev)) * 1e-3),
         "roofline_note": "2.4 MB working set, L2-resident and launch/latency bound: the roofline fraction is not meaningful",
         "parity": "byte-exact against the reference's golden file p2p-31-SSSP: %s" % ok, "parity_ok": bool(ok)}
    app.close()
    frag.close()
    from oracle import refdriver
    if refdriver.available():
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            gp = os.path.join(d, "p2p.bin")
            refdriver.write_graph(gp, n, src, dst, w, oids=oids)
            info, _ = refdriver.run_app("sssp", gp, directed=False, source=6, repeat=6, want_output=False)
        cpu_ms = float(np.median(info["query_ms"][1:]))
        r["cpu_reference"] = {"ms_per_query": cpu_ms, "threads": int(info["threads"]),
                              "what": "the unmodified reference CPU SSSP app (mpirun -n 1 run_app equivalent, "
                                      "oracle/_ref/ref_driver), Query() only"}
        r["gpu_over_cpu"] = cpu_ms / r["gpu_ms_per_query"]
    return r


def run_gpu(args):
    G = Group(args)
    torch = G.torch
    pkg = importlib.import_module("libgrape-lite_b200")
    rank, world = G.rank, G.world
    scale = args.scale if args.scale else 24 + int(np.log2(world))
    weighted = args.app == "sssp"
    n = 1 << scale
    t0 = time.time()
    frag = pkg.Fragment.rmat(scale, args.edgefactor, seed=args.seed, weight_mode=1 if weighted else 0, fid=rank, fnum=world)
    build_s = time.time() - t0
    comm = None
    if world > 1:
        gdist = importlib.import_module("libgrape-lite_b200.dist")
        comm = gdist.make_comm(rank, world, frag.ivnum, item_bytes=16)
    src_oid = global_source(G, frag, n)
    cfg = dict(source_oid=int(src_oid))
    if args.app == "bfs":
        cfg["direction_opt"] = 0 if args.push_only else 1
        cfg["fuse_supersteps"] = 0 if args.no_fuse else 1
        if args.bfs_beta:
            cfg.setdefault("reserved", {})[2] = args.bfs_beta
    if args.no_hub_order:
        cfg.setdefault("reserved", {})[1] = 1
    if args.app in ("pagerank", "cdlp"):
        cfg["max_round"] = 10
    if args.app == "pagerank" and (args.pr_pull or args.pr_f32):
        cfg["pr_pull"] = 1
        if args.pr_f32:
            cfg.setdefault("reserved", {})[5] = 1
        if args.pr_nohub:
            cfg.setdefault("reserved", {})[4] = 1
    kind = "wcc_opt" if (args.app == "wcc" and args.wcc_opt) else args.app
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    iters = 10 if args.app in ("pagerank", "cdlp") else 1

    def edges_fn(res):
        if args.app in ("pagerank", "cdlp"):
            return (args.edgefactor << scale) * iters
        return G.reduce_i64([reached_entries(pkg, frag, res, args.app)])[0] // 2

    sampler = ClockSampler(G.local)
    if rank == 0:
        sampler.start()
    r, app, res = measure_app(G, pkg, frag, comm, kind, cfg, args.steps, args.warmup, flush, edges_fn, weighted=weighted)
    clocks = sampler.stop() if rank == 0 else None
    app.close()
    bfs_parity = None
    if args.app == "bfs" and not args.no_parity:
        # the same query through an independent path of the engine: push only, one superstep per host
        # round (k_frontier_scan + k_hub_scan, per-vertex messages between fragments), the fragment's own
        # vertex order -- the depth arrays of all ranks must be identical (outside every timed region)
        chk = pkg.App("bfs", frag, comm, source_oid=int(src_oid), direction_opt=0, fuse_supersteps=0, reserved={1: 1})
        chk.query()
        ref = chk.result()
        chk.close()
        mismatching_ranks = G.reduce_i64([0 if np.array_equal(ref, res) else 1])[0]
        same = mismatching_ranks == 0
        bfs_parity = {"check": "depths(measured configuration) == depths(push-only, stepwise, fragment order) on every rank",
                      "parity_ok": bool(same)}
    peak, peak_src = r["_peak"]
    ms_per_step = r["ms_per_query"]
    edges = r["traversed_edges"]
    fused_query = args.app == "bfs" and not args.no_fuse
    if fused_query:
        dom_name = "k_bfs_fused" if world == 1 else "k_bfs_fused_multi"
        # ONE cooperative launch per GPU runs the whole query: rank 0's launch and its share of the bytes
        b_alg = r["alg_bytes"] / world
        dom_ms = ms_per_step
    else:
        d = r["dominant_superstep"]
        dom_name = "superstep %d (%s)" % (d["index"], {0: "k_frontier_scan+k_hub_scan", 1: "k_bfs_pull", 2: "dense scan"}[d["mode"]])
        b_alg, dom_ms = d["alg_bytes"], d["ms"]
    achieved = b_alg / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if dom_name in tj and args.app == "bfs" and scale == 24:
            traffic = tj[dom_name]["bytes"]
    except Exception:
        pass

    apps = None
    if args.sweep != "none" and args.app == "bfs":
        apps = sweep_apps(G, pkg, args, flush, frag, comm, scale, res)

    if rank == 0:
        line = {
            "metric": "TEPS (traversed edges/sec), %s" % args.app.upper(),
            "value": r["teps"], "unit": "edges/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE[kind], "data": "synthetic",
            "config": {"workload": "%s on R-MAT scale-%d edgefactor-%d undirected (seed %d), %d fragment(s)%s"
                       % (args.app.upper(), scale, args.edgefactor, args.seed, world,
                          (", push-only" if args.push_only else ", push/pull") if args.app == "bfs" else ""),
                       "mapping": MAPPING,
                       "parity": bfs_parity,
                       "vertices": n, "input_edges": args.edgefactor << scale,
                       "csr_entries_per_gpu": int(frag.oe_num), "traversed_edges": edges,
                       "source_oid": int(src_oid), "supersteps": r["supersteps"],
                       "l2": "L2 flushed (256 MB memset) between timed iterations; CSR (%.2f GB) > L2"
                             % (frag.device_bytes / 1e9),
                       "fragment_build_s": round(build_s, 2),
                       "app_setup_ms": r["app_setup_ms"], "app_extra_bytes": r["app_extra_bytes"],
                       "app_setup_note": "one-off per (fragment, app), outside the timed region: level bitmaps and, for "
                                         "BFS on one fragment, a degree-ordered shadow CSR + hub-neighbour table "
                                         "(amortised over queries like Graph500 kernel-1 preprocessing)",
                       "ms_per_superstep": r["ms_per_superstep"], "superstep_mode": r["superstep_mode"],
                       "superstep_entries": r["superstep_entries"]},
            "e2e": {"value": edges / (r["e2e_ms_per_query"] * 1e-3), "unit": "edges/s", "h2d_bytes_per_step": 8,
                    "d2h_bytes_per_step": r["d2h_bytes_per_step"], "ms_per_step": r["e2e_ms_per_query"],
                    "note": "gl_app_query + gl_app_result (int64 / double per inner vertex) into pinned host memory; "
                            "fragment resident (the reference also times Query() only, run_cuda_app.h:119-124)"},
            "gpu_launches": r["gpu_launches"],
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": dom_name, "alg_bytes_per_launch": b_alg, "ms_per_launch": dom_ms,
                         "whole_query_alg_bytes": r["alg_bytes"], "whole_query_gbs_per_gpu": r["whole_query_gbs"]},
        }
        if apps is not None:
            line["config"]["apps"] = apps
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, sample_scale=min(scale, args.cpu_scale))
        print(json.dumps(line), flush=True)
    if comm:
        comm.close()
    frag.close()
    G.close()


def cpu_baseline(args, sample_scale):
    """Times the reference's CPU path on the host cores on a bounded sample of the workload."""
    from oracle import refarm
    return refarm.run(args.app, sample_scale, args.edgefactor, args.seed, repeat=3, keep=2)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the same path on the same
    configuration (same generator, scale, seed, source rule), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import refarm
    want = args.scale if args.scale else 24 + int(np.log2(world))
    scale = min(want, args.ref_scale)
    t0 = time.time()
    # ONE process generates + loads the graph once and runs warmup + steps queries
    r = refarm.run(args.app, scale, args.edgefactor, args.seed, repeat=args.warmup + args.steps, keep=args.steps,
                   opt=args.ref_opt)
    value = float(r["value"])
    ms = float(r["ms"])
    note = "" if scale == want else (" -- bounded CPU sample of the scale-%d workload (host RAM / load time)" % want)
    line = {"impl": "reference", "metric": "TEPS (traversed edges/sec), %s" % args.app.upper(),
            "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": "%s on R-MAT scale-%d edgefactor-%d undirected (seed %d), 1 fragment(s)%s%s"
                                   % (args.app.upper(), scale, args.edgefactor, args.seed,
                                      ", push/pull" if args.app == "bfs" else "", note),
                       "same_input_as_gpu_arm": bool(scale == want), "source_oid": r.get("source_oid"),
                       "traversed_edges": r.get("traversed_edges"), "load_s": r.get("load_s"),
                       "vertices": 1 << scale, "input_edges": args.edgefactor << scale},
            "cpu_baseline": {"value": value, "unit": "edges/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": r["sample"]},
            "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(time.time() - t0, 1)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--app", default="bfs", choices=["bfs", "sssp", "wcc", "pagerank", "cdlp", "lcc"])
    ap.add_argument("--scale", type=int, default=0, help="R-MAT scale (default 24 + log2(gpus))")
    ap.add_argument("--edgefactor", type=int, default=16)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--push-only", action="store_true")
    ap.add_argument("--sweep", default="default", choices=["default", "all", "none"],
                    help="config.apps: the other BASELINE configs (C1, C3, C4, C5) measured in the same run")
    ap.add_argument("--cpu-scale", type=int, default=22, help="scale of the bounded cpu_baseline sample inside the GPU arm's line")
    ap.add_argument("--ref-scale", type=int, default=24, help="largest scale --impl reference runs (scale 24 = the 1-GPU workload itself)")
    ap.add_argument("--ref-opt", action="store_true", help="--impl reference: the reference's tuned --opt CPU apps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--wcc-opt", action="store_true", help="WCC by union-find (the reference's wcc_opt) instead of label propagation")
    ap.add_argument("--pr-f32", action="store_true", help="PageRank pull gathering f32 contributions (f64 sums)")
    ap.add_argument("--pr-nohub", action="store_true", help="with --pr-f32: without the shared-memory hub table (A/B)")
    ap.add_argument("--no-hub-order", action="store_true", help="disable the hub-first shadow CSR (BFS)")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run BFS parity check (second app, push-only)")
    ap.add_argument("--bfs-beta", type=int, default=0, help="BFS pull->push threshold divisor (0 = library default)")
    ap.add_argument("--pr-pull", action="store_true", help="PageRank: deterministic pull step instead of atomicAdd push")
    ap.add_argument("--no-fuse", action="store_true", help="one launch per superstep (profiling) instead of the fused query kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — TEPS / ms-per-superstep of the PIE hot path on B200.

One "step" = one whole query (PEval + IncEval supersteps) of the app over the
resident fragment(s).  Default workload = BASELINE.json configs[1]:
BFS on R-MAT scale-24 (edgefactor 16, undirected), 1 x B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--app bfs] [--scale S]
  python bench.py --impl reference ...     # the CPU arm (reference apps / port)

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
def alg_bytes(app, entries, frontier, touched, weighted):
    """B_alg = m_s*(4+w) + n_f*12 + n_t*(s_r+s_w)."""
    w = 4 if weighted else 0
    state = {"bfs": 8, "sssp": 8, "wcc": 8, "pagerank": 24, "cdlp": 16, "lcc": 8}[app]
    return entries * (4 + w) + frontier * 12 + touched * state


def traversed_edges(pkg, frag, result, app):
    """Graph500 TEPS numerator: input edges whose endpoints were reached
    (sum of degrees of reached vertices / 2 for an undirected graph)."""
    rp, _, _ = frag.csr(0) if frag.ivnum <= (1 << 22) else (None, None, None)
    if rp is None:
        rp = np.zeros(frag.ivnum + 1, dtype=np.uint64)
        pkg.check(pkg.lib().gl_frag_copy_csr(frag.h, 0, pkg.capi._p(rp), None, None))
    deg = np.diff(rp).astype(np.int64)
    if app == "bfs":
        reached = result != np.iinfo(np.int64).max
    elif app == "sssp":
        reached = result < 1e300
    else:
        reached = np.ones(len(result), dtype=bool)
    return int(deg[reached].sum())  # CSR entries; halved by the caller across ranks


def run_gpu(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("libgrape-lite_b200")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scale = args.scale if args.scale else 24 + int(np.log2(world))
    weighted = args.app == "sssp"
    wmode = 1 if weighted else 0
    n = 1 << scale

    t0 = time.time()
    frag = pkg.Fragment.rmat(scale, args.edgefactor, seed=args.seed, weight_mode=wmode, fid=rank, fnum=world)
    build_s = time.time() - t0

    comm = None
    if world > 1:
        gdist = importlib.import_module("libgrape-lite_b200.dist")
        comm = gdist.make_comm(rank, world, frag.ivnum, item_bytes=16)

    # source = max-degree vertex of the whole graph, ties -> smallest oid
    lid, deg = frag.max_degree_vertex()
    chunk = (n + world - 1) // world
    src_oid = rank * chunk + lid
    if world > 1:
        t = torch.tensor([deg, -src_oid], dtype=torch.int64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        best = max((int(x[0]), int(x[1])) for x in allt)
        src_oid = -best[1]
    cfg = dict(source_oid=int(src_oid), lb=args.lb)
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
    app = pkg.App("wcc_opt" if (args.app == "wcc" and args.wcc_opt) else args.app, frag, comm, **cfg)

    pinned = pkg.PinnedBuffer(8 * max(frag.ivnum, 1))
    out = pinned.array(pkg.capi.RESULT_DTYPE[app.kind], frag.ivnum)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (also gives the TEPS numerator)
    for _ in range(max(args.warmup, 1)):
        st = app.query()
    res = app.result(out)
    entries_reached = traversed_edges(pkg, frag, res, args.app)
    if world > 1:
        t = torch.tensor([entries_reached], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        entries_reached = int(t.item())
    edges = entries_reached // 2
    iters = 10 if args.app in ("pagerank", "cdlp") else 1
    if args.app in ("pagerank", "cdlp"):
        edges = (args.edgefactor << scale) * iters

    # ---- device-timed region: K queries, inputs resident -------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    step_stats = []
    launches = 0
    barrier()
    t_dev = 0.0
    for _ in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations
        torch.cuda.synchronize()
        st = app.query()
        t_dev += st.query_ms
        launches += st.kernel_launches
        step_stats.append((st.supersteps, [st.step_ms[i] for i in range(st.n_steps)],
                           [st.step_entries[i] for i in range(st.n_steps)],
                           [st.step_frontier[i] for i in range(st.n_steps)],
                           [st.step_mode[i] for i in range(st.n_steps)],
                           st.entries_scanned, st.frontier_vertices, st.touched_vertices,
                           [st.step_kernel_ms[i] for i in range(st.n_steps)] if hasattr(st, "step_kernel_ms") else None))
    barrier()
    # ---- end-to-end region: C-ABI call with host buffers -------------------
    t_e2e = 0.0
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        t1 = time.perf_counter()
        app.query()
        app.result(out)                    # D2H of the step's result (pinned)
        barrier()
        t_e2e += time.perf_counter() - t1
    clocks = sampler.stop() if rank == 0 else None

    tt = torch.tensor([t_dev, t_e2e * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev_ms, t_e2e_ms = float(tt[0]), float(tt[1])
    ms_per_step = t_dev_ms / args.steps
    value = edges * args.steps / (t_dev_ms * 1e-3)
    e2e_value = edges * args.steps / (t_e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (live CUDA-event timing) -----------
    peak, peak_src = load_peaks()
    last = step_stats[-1]
    ks = last[8] if last[8] else last[1]
    # dominant superstep = largest kernel time, averaged over the timed steps
    nsteps = len(last[1])
    agg_ms = np.zeros(nsteps)
    agg_ent = np.zeros(nsteps)
    agg_fr = np.zeros(nsteps)
    cnt = 0
    for s in step_stats:
        if len(s[1]) == nsteps:
            agg_ms += np.array(s[8] if s[8] else s[1])
            agg_ent += np.array(s[2], dtype=np.float64)
            agg_fr += np.array(s[3], dtype=np.float64)
            cnt += 1
    agg_ms /= max(cnt, 1)
    agg_ent /= max(cnt, 1)
    agg_fr /= max(cnt, 1)
    dom = int(np.argmax(agg_ms))
    touched_dom = last[7] * (agg_ent[dom] / max(sum(agg_ent), 1.0))
    whole_balg = alg_bytes(args.app, last[5], last[6], last[7], weighted)
    fused_query = args.app == "bfs" and not args.no_fuse
    if fused_query:
        # the whole query is ONE cooperative launch per GPU: the dominant kernel is the step
        # (world > 1: rank 0's launch; its algorithmic bytes are rank 0's share of the query)
        dom_name = "k_bfs_fused" if world == 1 else "k_bfs_fused_multi"
        b_alg = whole_balg
        dom_ms = ms_per_step
    else:
        dom_name = {0: "k_frontier_scan+k_hub_scan", 1: "k_bfs_pull", 2: "dense scan"}[int(last[4][dom])]
        b_alg = alg_bytes(args.app, agg_ent[dom], agg_fr[dom], touched_dom, weighted)
        dom_ms = float(agg_ms[dom])
    achieved = b_alg / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = dom_name if dom_name in tj else None
        if key and args.app == "bfs" and scale == 24:
            traffic = tj[key]["bytes"]
    except Exception:
        pass

    line = None
    if rank == 0:
        line = {
            "metric": "TEPS (traversed edges/sec), %s" % args.app.upper(),
            "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"bfs": "u32", "sssp": "f32", "wcc": "u32", "pagerank": "f64",
                                           "cdlp": "int64", "lcc": "u32"}[args.app],
            "data": "synthetic",
            "config": {"workload": "%s on R-MAT scale-%d edgefactor-%d undirected (seed %d), %d fragment(s), lb=%s%s"
                       % (args.app.upper(), scale, args.edgefactor, args.seed, world, args.lb,
                          ", push-only" if args.push_only else ", push/pull"),
                       "vertices": n, "input_edges": args.edgefactor << scale,
                       "csr_entries_per_gpu": int(frag.oe_num), "traversed_edges": edges,
                       "source_oid": int(src_oid), "supersteps": int(last[0]),
                       "l2": "L2 flushed (256 MB memset) between timed iterations; CSR (%.2f GB) > L2"
                             % (frag.device_bytes / 1e9),
                       "fragment_build_s": round(build_s, 2),
                       "ms_per_superstep": [round(float(x), 4) for x in np.array(last[1])],
                       "superstep_mode": [int(x) for x in last[4]],
                       "superstep_entries": [int(x) for x in last[2]]},
            "e2e": {"value": e2e_value, "unit": "edges/s", "h2d_bytes_per_step": 8,
                    "d2h_bytes_per_step": int(8 * frag.ivnum), "ms_per_step": t_e2e_ms / args.steps,
                    "note": "gl_app_query + gl_app_result to pinned host memory; fragment resident "
                            "(the reference also times Query() only, run_cuda_app.h:119-124)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": dom_name if fused_query else "superstep %d (%s)" % (dom, dom_name),
                         "alg_bytes_per_launch": b_alg, "ms_per_launch": dom_ms,
                         "whole_query_alg_bytes": whole_balg,
                         "whole_query_gbs": whole_balg / (ms_per_step * 1e-3) / 1e9},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, sample_scale=min(scale, args.cpu_scale))
        print(json.dumps(line), flush=True)
    app.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, sample_scale):
    """Times the CPU path on the host cores on a bounded sample of the workload."""
    from oracle import refarm
    return refarm.run(args.app, sample_scale, args.edgefactor, args.seed, repeat=3, keep=2)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the same path."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import refarm
    scale = min(args.scale if args.scale else 24 + int(np.log2(world)), args.cpu_scale)
    t0 = time.time()
    # ONE process loads the graph once and runs warmup + steps queries
    r = refarm.run(args.app, scale, args.edgefactor, args.seed, repeat=args.warmup + args.steps, keep=args.steps)
    value = float(r["value"])
    ms = float(r["ms"])
    line = {"impl": "reference", "metric": "TEPS (traversed edges/sec), %s" % args.app.upper(),
            "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": "%s on R-MAT scale-%d edgefactor-%d undirected (seed %d) — bounded CPU sample of "
                                   "the scale-%d workload" % (args.app.upper(), scale, args.edgefactor, args.seed,
                                                             args.scale if args.scale else 24 + int(np.log2(world)))},
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
    ap.add_argument("--lb", default="cm")
    ap.add_argument("--push-only", action="store_true")
    ap.add_argument("--cpu-scale", type=int, default=22, help="largest scale the CPU arm runs (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--wcc-opt", action="store_true", help="WCC by union-find (the reference's wcc_opt) instead of label propagation")
    ap.add_argument("--pr-f32", action="store_true", help="PageRank pull gathering f32 contributions (f64 sums)")
    ap.add_argument("--no-hub-order", action="store_true", help="disable the hub-first shadow CSR (BFS)")
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

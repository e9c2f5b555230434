"""CPU model of the multi-fragment path (host logic of §8e), run under
torch.distributed/gloo: the same partitioner, gid format, outer-vertex
numbering and round protocol (apply received items -> local superstep -> pack
(lid_at_owner, value) items per owner -> exchange -> one all-reduce as barrier
+ termination vote) as libgrape-lite_b200/csrc/{fragment,comm}.cu, with numpy
standing in for the kernels.  It validates the distributed algorithm; the CUDA
implementation of the same protocol is tested on GPUs (tests/mgpu_worker.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def id_parser(fnum):
    """grape/fragment/id_parser.h:28-41"""
    maxfid = fnum - 1
    if maxfid == 0:
        off = 31
    else:
        off = 32 - int(maxfid).bit_length()
    return off, (1 << off) - 1


class Frag:
    """Edge-cut fragment built with the rules of fragment.cu (segmented
    partitioner, rows sorted inner-first then outer by gid)."""

    def __init__(self, n, src, dst, w, fid, fnum):
        self.fid, self.fnum, self.n = fid, fnum, n
        self.chunk = (n + fnum - 1) // fnum
        self.lo, self.hi = min(n, fid * self.chunk), min(n, (fid + 1) * self.chunk)
        self.ivnum = self.hi - self.lo
        self.off, self.mask = id_parser(fnum)
        a = np.concatenate([src, dst])
        b = np.concatenate([dst, src])
        ww = None if w is None else np.concatenate([w, w])
        keep = (a >= self.lo) & (a < self.hi)
        a, b = a[keep], b[keep]
        ww = None if ww is None else ww[keep]
        inner = (b >= self.lo) & (b < self.hi)
        outer_glob = np.unique(b[~inner])
        self.ovnum = len(outer_glob)
        owner = outer_glob // self.chunk
        self.ovgid = (owner << self.off) | (outer_glob - owner * self.chunk)
        lid = np.where(inner, b - self.lo, self.ivnum + np.searchsorted(outer_glob, b))
        order = np.lexsort((lid, a))
        self.row = (a - self.lo)[order]
        self.col = lid[order]
        self.w = None if ww is None else ww[order]
        self.rp = np.zeros(self.ivnum + 1, dtype=np.int64)
        np.add.at(self.rp, self.row + 1, 1)
        self.rp = np.cumsum(self.rp)
        self.tvnum = self.ivnum + self.ovnum

    def outer_owner_and_lid(self, outer_idx):
        g = self.ovgid[outer_idx]
        return g >> self.off, g & self.mask


def exchange(per_dst):
    """all-to-all of python lists of items + the barrier/termination vote."""
    world = dist.get_world_size()
    gathered = [None] * world
    dist.all_gather_object(gathered, per_dst)
    me = dist.get_rank()
    return [gathered[src][me] for src in range(world) if src != me]


def vote(force_continue, sent):
    t = torch.tensor([1 if force_continue else 0, sent], dtype=torch.int64)
    dist.all_reduce(t)
    return int(t[0]) == 0 and int(t[1]) == 0


def run_min_app(f, init_state, seed_active, relax):
    """Generic min-propagation PIE loop (BFS with unit weights, SSSP, WCC)."""
    state = init_state.copy()
    active = seed_active.copy()
    inbox = []
    rounds = 0
    while True:
        rounds += 1
        for items in inbox:                      # ParallelProcess
            for lid, val in items:
                if val < state[lid]:
                    state[lid] = val
                    active[lid] = True
        remote = np.zeros(f.tvnum, dtype=bool)
        nxt = np.zeros(f.ivnum, dtype=bool)
        for u in np.nonzero(active)[0]:          # edge scan of the active inner vertices
            for e in range(f.rp[u], f.rp[u + 1]):
                v = f.col[e]
                nv = relax(state[u], None if f.w is None else f.w[e])
                if nv < state[v]:
                    state[v] = nv
                    if v < f.ivnum:
                        nxt[v] = True
                    else:
                        remote[v] = True
        per_dst = [[] for _ in range(f.fnum)]
        for v in np.nonzero(remote)[0]:          # pack (lid_at_owner, value) per owner
            o, lid = f.outer_owner_and_lid(v - f.ivnum)
            per_dst[int(o)].append((int(lid), state[v]))
        sent = sum(len(x) for x in per_dst)
        inbox = exchange(per_dst)
        active = nxt
        if vote(bool(nxt.any()), sent):
            break
        if rounds > 10000:
            raise RuntimeError("no convergence")
    return state[:f.ivnum], rounds


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import importlib
    pkg = importlib.import_module("libgrape-lite_b200")
    from oracle import pyoracle
    scale = 9
    n = 1 << scale
    src, dst, w = pkg.rmat_edges_host(scale, 8, seed=31, weight_mode=1)
    f = Frag(n, src, dst, w.astype(np.float64), rank, world)
    g = pyoracle.Graph(n, src, dst, w.astype(np.float64))
    source = g.max_degree_vertex()
    ok = True

    def gather(x):
        out = [None] * world
        dist.all_gather_object(out, x)
        return np.concatenate(out)

    # fragment layout invariants shared with fragment.cu
    assert np.all(np.diff(f.ovgid) > 0)
    assert np.all((f.ovgid >> f.off) != rank)
    orp, ocol, _ = g.csr()
    assert np.array_equal(f.rp, orp[f.lo:f.hi + 1] - orp[f.lo])

    INF = np.finfo(np.float64).max
    # SSSP
    st = np.full(f.tvnum, INF)
    act = np.zeros(f.ivnum, dtype=bool)
    if f.lo <= source < f.hi:
        st[source - f.lo] = 0.0
        act[source - f.lo] = True
    res, _ = run_min_app(f, st, act, lambda d, ww: d + ww)
    got = gather(res)
    ok &= bool(np.array_equal(got, g.sssp(source)[0]))
    # BFS = SSSP with unit weights
    st = np.full(f.tvnum, INF)
    if f.lo <= source < f.hi:
        st[source - f.lo] = 0.0
    res, _ = run_min_app(f, st, act, lambda d, ww: d + 1.0)
    got = gather(res)
    want = g.bfs(source)[0].astype(np.float64)
    want[want > 1e18] = INF
    ok &= bool(np.array_equal(got, want))
    # WCC: labels = gid, result mapped back to the global index
    lab = np.concatenate([(rank << f.off) | np.arange(f.ivnum), f.ovgid]).astype(np.float64)
    res, _ = run_min_app(f, lab, np.ones(f.ivnum, dtype=bool), lambda d, ww: d)
    gid = res.astype(np.int64)
    got = gather((gid >> f.off) * f.chunk + (gid & f.mask))
    ok &= bool(np.array_equal(got, g.wcc()[0].astype(np.int64)))
    flag = torch.tensor([0 if ok else 1])
    dist.all_reduce(flag)
    if rank == 0:
        print("dist_model world=%d %s" % (world, "OK" if int(flag) == 0 else "MISMATCH"))
    dist.destroy_process_group()
    sys.exit(int(flag))


if __name__ == "__main__":
    main()

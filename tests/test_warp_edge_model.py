"""Lane-by-lane model of OpBfsPush::warp_edge (libgrape-lite_b200/csrc/app_bfs.cu): 32 entries of a
hub row per call; the lanes of a run of equal bitmap words OR their bits and the run's last lane
issues one atomicOr.  The model restates the device code's data flow (heads / segment ids /
segmented inclusive OR scan / tails) and is compared with the per-entry semantics of
OpBfsPush::edge on random rows -- sorted and unsorted, with parallel edges and already visited
targets.  CPU only: it pins the algorithm, the GPU tests pin the kernel."""
import numpy as np


def warp_edge_model(vs, valid, vis, nxt, stale_vis):
    """One call; vis / nxt are dicts word -> int.  Returns the number of newly visited vertices."""
    n = 32
    w = [v >> 5 for v in vs]
    bit = [0] * n
    for l in range(n):
        if valid[l]:
            bit[l] = (1 << (vs[l] & 31)) & ~stale_vis.get(w[l], 0) & 0xFFFFFFFF
    if not any(bit):
        return 0
    head = [l == 0 or w[l - 1] != w[l] for l in range(n)]
    heads = sum(1 << l for l in range(n) if head[l])
    seg = [bin(heads & ((2 << l) - 1)).count("1") for l in range(n)]
    incl = list(bit)
    o = 1
    while o < 32:                         # Hillis-Steele, guarded by the segment id
        prev = list(incl)
        for l in range(n):
            if l >= o and seg[l - o] == seg[l]:
                incl[l] = prev[l] | prev[l - o]
        o <<= 1
    excl = [0 if head[l] else incl[l - 1] for l in range(n)]
    tails = (heads >> 1) | 0x80000000
    old = [0] * n
    for l in range(n):                    # same-address atomics of one instruction are serialised
        if (tails >> l) & 1 and incl[l]:
            old[l] = vis.get(w[l], 0)
            vis[w[l]] = old[l] | incl[l]
            fresh = incl[l] & ~old[l]
            if fresh:
                nxt[w[l]] = nxt.get(w[l], 0) | fresh
    new = 0
    for l in range(n):
        t = tails >> l
        mytail = l + ((t & -t).bit_length() - 1)
        if bit[l] & ~old[mytail] & ~excl[l]:
            new += 1
    return new


def per_entry(vs, valid, vis, nxt):
    new = 0
    for v, ok in zip(vs, valid):
        if not ok:
            continue
        w, b = v >> 5, 1 << (v & 31)
        if vis.get(w, 0) & b:
            continue
        vis[w] = vis.get(w, 0) | b
        nxt[w] = nxt.get(w, 0) | b
        new += 1
    return new


def test_warp_edge_model_matches_per_entry_semantics():
    rng = np.random.default_rng(5)
    for trial in range(400):
        span = int(rng.choice([40, 200, 5000]))
        vs = rng.integers(0, span, 32)
        if trial % 2 == 0:
            vs = np.sort(vs)                               # rows of the hub-first shadow graph
        if trial % 5 == 0:
            vs[rng.integers(0, 32, 6)] = vs[rng.integers(0, 32, 6)]   # more parallel edges
        vs = [int(x) for x in vs]
        valid = [bool(x) for x in (rng.random(32) < 0.9)]
        pre = {}
        for v in rng.integers(0, span, 10):                # already visited targets
            pre[int(v) >> 5] = pre.get(int(v) >> 5, 0) | (1 << (int(v) & 31))
        stale = {k: (x if rng.random() < 0.5 else 0) for k, x in pre.items()}   # the plain read may be stale
        vis_a, nxt_a = dict(pre), {}
        vis_b, nxt_b = dict(pre), {}
        na = warp_edge_model(vs, valid, vis_a, nxt_a, stale)
        nb = per_entry(vs, valid, vis_b, nxt_b)
        assert na == nb, (trial, vs)
        assert {k: x for k, x in vis_a.items() if x} == {k: x for k, x in vis_b.items() if x}
        assert {k: x for k, x in nxt_a.items() if x} == {k: x for k, x in nxt_b.items() if x}

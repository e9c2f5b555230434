"""GPU: the C-ABI engine primitives — ForEachOutgoingEdge over a
WorkSourceArray in every load-balancing mode of the reference
(--lb none|cm|wm|cta|strict, parallel_engine.h:51-70) give identical results,
equal to a numpy restatement."""
import numpy as np
import pytest

from oracle import pyoracle
from tests.util import pkg, rmat_graph

pytestmark = pytest.mark.gpu
LBS = ["none", "cm", "wm", "cta", "strict", "cmold"]
INF32 = np.uint32(0xFFFFFFFF)


@pytest.fixture(scope="module")
def graph():
    n, src, dst, w = rmat_graph(13, seed=4, weight_mode=1)
    g = pyoracle.Graph(n, src, dst, w.astype(np.float64))
    frag = pkg().Fragment.rmat(13, 16, seed=4, weight_mode=1)
    rp, col, ww = g.csr()
    yield n, g, frag, rp, col, ww
    frag.close()


@pytest.mark.parametrize("lb", LBS)
def test_bfs_level_by_level(graph, lb):
    """Expanding the frontier queue level by level with GL_OP_BFS_LEVEL
    reproduces the oracle's depths; scanned entries = sum of frontier degrees."""
    P = pkg()
    n, g, frag, rp, col, _ = graph
    want, _ = g.bfs(g.max_degree_vertex())
    level = np.full(n, INF32, dtype=np.uint32)
    src = g.max_degree_vertex()
    level[src] = 0
    d_level = P.DeviceArray(level)
    words = (n + 31) // 32
    d_bm = P.DeviceArray(nbytes=4 * words)
    d_q = P.DeviceArray(nbytes=4 * n)
    frontier = np.array([src], dtype=np.uint32)
    depth = 0
    while len(frontier):
        d_q.upload(frontier)
        d_bm.fill(0)
        scanned = P.edge_scan_queue(frag, d_q, len(frontier), "bfs_level", lb, state=d_level,
                                    out_bitmap=d_bm, depth=depth + 1)
        assert scanned == int((rp[frontier + 1] - rp[frontier]).sum())
        cnt = P.compact_bitmap(d_bm, n, d_q)
        frontier = np.sort(d_q.download(np.uint32, cnt))
        depth += 1
        assert np.array_equal(frontier, np.nonzero(want == depth)[0].astype(np.uint32))
    got = d_level.download(np.uint32, n).astype(np.int64)
    got[got == 0xFFFFFFFF] = np.iinfo(np.int64).max
    assert np.array_equal(got, want)


@pytest.mark.parametrize("lb", LBS)
def test_min_relax_and_add_scatter_one_step(graph, lb):
    P = pkg()
    n, g, frag, rp, col, w = graph
    rng = np.random.default_rng(5)
    q = np.unique(rng.integers(0, n, 3000)).astype(np.uint32)
    hub = g.max_degree_vertex()
    q = np.unique(np.append(q, np.uint32(hub)))                 # include the longest row
    d_q = P.DeviceArray(q)
    rows = np.repeat(q, (rp[q + 1] - rp[q]).astype(np.int64))
    idx = np.concatenate([np.arange(rp[u], rp[u + 1]) for u in q]).astype(np.int64)
    # min-relax f32 with weights: state[v] = min(state[v], state[u] + w)
    st = rng.integers(0, 1000, n).astype(np.float32)
    want = st.copy()
    np.minimum.at(want, col[idx], st[rows] + w[idx].astype(np.float32))
    d_st = P.DeviceArray(st)
    d_bm = P.DeviceArray(nbytes=4 * ((n + 31) // 32))
    d_bm.fill(0)
    P.edge_scan_queue(frag, d_q, len(q), "min_relax_f32", lb, state=d_st, out_bitmap=d_bm, use_weight=1)
    got = d_st.download(np.float32, n)
    # sources may themselves be lowered during the step (Gauss-Seidel style):
    # the result is bounded by the Jacobi step and never above the old state
    assert np.all(got <= want) and np.all(got <= st)
    changed = np.unpackbits(d_bm.download(np.uint8, 4 * ((n + 31) // 32)), bitorder="little")[:n].astype(bool)
    assert np.array_equal(changed, got < st)
    # min-relax u32 without weights (label propagation step)
    lab = rng.permutation(n).astype(np.uint32)
    d_lab = P.DeviceArray(lab)
    P.edge_scan_queue(frag, d_q, len(q), "min_relax_u32", lb, state=d_lab, use_weight=0)
    got = d_lab.download(np.uint32, n)
    jac = lab.copy()
    np.minimum.at(jac, col[idx], lab[rows])
    assert np.all(got <= jac)
    # add-scatter f64: dst[v] += src[u]
    srcv = rng.random(n)
    d_src = P.DeviceArray(srcv)
    d_dst = P.DeviceArray(np.zeros(n))
    sc = P.edge_scan_queue(frag, d_q, len(q), "add_scatter_f64", lb, state=d_src, state2=d_dst)
    want = np.zeros(n)
    np.add.at(want, col[idx], srcv[rows])
    got = d_dst.download(np.float64, n)
    assert sc == len(idx)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_empty_queue_and_empty_bitmap(graph):
    P = pkg()
    n, g, frag, *_ = graph
    d_q = P.DeviceArray(nbytes=16)
    d_level = P.DeviceArray(np.full(n, INF32, dtype=np.uint32))
    assert P.edge_scan_queue(frag, d_q, 0, "bfs_level", "cm", state=d_level, depth=1) == 0
    d_bm = P.DeviceArray(nbytes=4 * ((n + 31) // 32))
    d_bm.fill(0)
    d_out = P.DeviceArray(nbytes=4 * n)
    assert P.compact_bitmap(d_bm, n, d_out) == 0


def test_face2_containers_queue_varray_prepare():
    """gl_queue_* (Queue), gl_varray_* (VertexArray), gl_bitmap_* (DenseVertexSet), gl_frag_prepare."""
    import ctypes as C
    P = pkg()
    L = P.lib()
    n = 100000
    rng = np.random.default_rng(5)
    bits = rng.random(n) < 0.07
    words = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
    idx = np.nonzero(bits)[0]
    np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    bm = C.POINTER(C.c_uint32)()
    P.check(L.gl_bitmap_create(C.byref(bm), C.c_uint64(n)))
    P.check(L.gl_dev_h2d(bm, words.ctypes.data_as(C.c_void_p), C.c_size_t(words.nbytes)))
    cnt = C.c_uint64()
    P.check(L.gl_bitmap_count(None, bm, C.c_uint64(n), C.byref(cnt)))
    assert cnt.value == len(idx)
    q = C.c_void_p()
    P.check(L.gl_queue_create(C.byref(q), C.c_uint32(n)))
    P.check(L.gl_queue_fill_from_bitmap(q, None, bm, C.c_uint32(n)))
    size = C.c_uint32()
    P.check(L.gl_queue_size(q, None, C.byref(size)))
    assert size.value == len(idx)
    data, dcount = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
    P.check(L.gl_queue_data(q, C.byref(data), C.byref(dcount)))
    got = np.empty(size.value, dtype=np.uint32)
    P.check(L.gl_dev_d2h(got.ctypes.data_as(C.c_void_p), data, C.c_size_t(got.nbytes)))
    assert np.array_equal(np.sort(got), idx.astype(np.uint32))
    P.check(L.gl_queue_clear(q, None))
    P.check(L.gl_queue_size(q, None, C.byref(size)))
    assert size.value == 0
    L.gl_queue_destroy(q)
    P.check(L.gl_bitmap_clear(None, bm, C.c_uint64(n)))
    P.check(L.gl_bitmap_count(None, bm, C.c_uint64(n), C.byref(cnt)))
    assert cnt.value == 0
    P.check(L.gl_bitmap_destroy(bm))
    va = C.c_void_p()
    vals = rng.random(1000)
    P.check(L.gl_varray_create(C.byref(va), C.c_uint64(1000), 8, 0))
    P.check(L.gl_varray_h2d(va, vals.ctypes.data_as(C.c_void_p), C.c_uint64(1000), 8))
    back = np.empty(1000)
    P.check(L.gl_varray_d2h(va, back.ctypes.data_as(C.c_void_p), C.c_uint64(1000), 8))
    assert np.array_equal(back, vals)
    P.check(L.gl_varray_destroy(va))
    frag = P.Fragment.rmat(8, 16, seed=3)
    P.check(L.gl_frag_prepare(frag.h, 0, 1, 0))
    frag.offload()
    assert L.gl_frag_prepare(frag.h, 0, 1, 0) != 0
    frag.reload()
    frag.close()

"""GPU: the device-built fragment stores exactly the adjacency the reference's
ImmutableEdgecutFragment would (rows sorted by neighbour lid, inner first,
multi-edges/self loops kept) — compared entry by entry with the oracle CSR."""
import numpy as np
import pytest

from oracle import pyoracle
from tests import golden_io as G
from tests.util import pkg, rmat_graph

pytestmark = pytest.mark.gpu


def _check_single(frag, g, weighted):
    rp, col, w = frag.csr(0)
    orp, ocol, ow = g.csr(False)
    assert np.array_equal(rp, orp)
    assert np.array_equal(col, ocol)
    if weighted:
        # duplicates (same row, same neighbour) may carry different weights in
        # any order: compare per (row, col) group as multisets
        rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp).astype(np.int64))
        a = np.lexsort((w.astype(np.float64), col, rows))
        b = np.lexsort((ow, ocol, rows))
        assert np.array_equal(w.astype(np.float64)[a], ow[b])


@pytest.mark.parametrize("directed", [False, True])
def test_p2p31_layout(directed):
    oids, src, dst, w = G.load_p2p31()
    g = pyoracle.Graph(len(oids), src, dst, w, directed=directed, oids=oids)
    frag = pkg().Fragment.from_edges(len(oids), src, dst, w, directed=directed, oids=oids)
    assert frag.ivnum == 62586 and frag.ovnum == 0
    _check_single(frag, g, True)
    if directed:
        rp, col, _ = frag.csr(1)
        orp, ocol, _ = g.csr(True)
        assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    assert frag.oid2lid(6) == g.index_of(6)
    lid, deg = frag.max_degree_vertex()
    assert lid == g.max_degree_vertex()
    frag.close()


@pytest.mark.parametrize("scale,weight_mode", [(10, 0), (14, 1)])
def test_rmat_device_build_equals_host_edges(scale, weight_mode):
    n, src, dst, w = rmat_graph(scale, seed=5, weight_mode=weight_mode)
    wd = None if w is None else w.astype(np.float64)
    g = pyoracle.Graph(n, src, dst, wd, directed=False)
    f1 = pkg().Fragment.rmat(scale, 16, seed=5, weight_mode=weight_mode)
    _check_single(f1, g, weight_mode != 0)
    f2 = pkg().Fragment.from_edges(n, src, dst, w)
    _check_single(f2, g, weight_mode != 0)
    f1.close()
    f2.close()


def test_from_csr_roundtrip():
    n, src, dst, w = rmat_graph(10, seed=2)
    g = pyoracle.Graph(n, src, dst, None)
    orp, ocol, _ = g.csr(False)
    f = pkg().Fragment.from_csr(n, orp, ocol)
    rp, col, _ = f.csr(0)
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    f.offload()
    with pytest.raises(pkg().GrapeError):
        f.csr(0)
    f.reload()
    rp, col, _ = f.csr(0)
    assert np.array_equal(col, ocol)
    f.close()


@pytest.mark.parametrize("fnum", [2, 3, 8])
def test_edge_cut_partition_layout(fnum):
    """Every fragment's rows, mapped back to global ids, equal the oracle's
    rows; outer lids are in gid order grouped by owner; the reverse adjacency
    of the outer vertices is the transpose of the outer part of oe."""
    scale = 11
    n, src, dst, _ = rmat_graph(scale, seed=9)
    g = pyoracle.Graph(n, src, dst, None)
    orp, ocol, _ = g.csr(False)
    chunk = (n + fnum - 1) // fnum
    for fid in range(fnum):
        f = pkg().Fragment.rmat(scale, 16, seed=9, fid=fid, fnum=fnum)
        lo, hi = min(n, fid * chunk), min(n, (fid + 1) * chunk)
        assert f.ivnum == hi - lo
        rp, col, _ = f.csr(0)
        ovgid = f.ovgid()
        off = f.fid_offset
        assert np.all(np.diff(ovgid.astype(np.int64)) > 0)
        owner = ovgid >> off
        assert np.all(owner != fid)
        glob_outer = owner.astype(np.int64) * chunk + (ovgid & ((1 << off) - 1))
        # local -> global
        lid2g = np.concatenate([np.arange(lo, hi), glob_outer])
        assert np.array_equal(rp - rp[0], orp[lo:hi + 1] - orp[lo])
        seg = ocol[orp[lo]:orp[hi]].astype(np.int64)
        mine = lid2g[col]
        # same multiset per row, and inner neighbours precede outer ones with
        # each part ascending
        rows = np.repeat(np.arange(hi - lo), np.diff(rp).astype(np.int64))
        assert np.array_equal(np.sort(mine + rows * n), np.sort(seg + rows * n))
        key = col.astype(np.int64) + rows * (1 << 33)
        assert np.all(np.diff(key) >= 0)
        # reverse adjacency of outer vertices
        vrp, vcol, _ = f.csr(2)
        is_outer = col >= f.ivnum
        pairs = np.stack([col[is_outer].astype(np.int64) - f.ivnum, rows[is_outer]], 1)
        pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
        assert np.array_equal(np.diff(vrp).astype(np.int64), np.bincount(pairs[:, 0], minlength=f.ovnum))
        assert np.array_equal(vcol.astype(np.int64), pairs[:, 1])
        f.close()


def test_fragment_save_load_roundtrip(tmp_path):
    """gl_frag_save / gl_frag_load (Serialize / Deserialize analogue): every array of the reloaded
    fragment equals the original, for weighted, directed and multi-fragment cases, and apps run on it."""
    P = pkg()
    n, src, dst, w = rmat_graph(11, seed=6, weight_mode=1)
    oids = np.arange(n, dtype=np.int64) * 3 + 5
    cases = [P.Fragment.rmat(11, 16, seed=6, weight_mode=1),
             P.Fragment.from_edges(n, oids[src], oids[dst], w, oids=oids, directed=True),
             P.Fragment.rmat(11, 16, seed=6, fid=1, fnum=3)]
    for k, f in enumerate(cases):
        path = str(tmp_path / ("frag_%d.bin" % k))
        f.save(path)
        g = P.Fragment.load(path)
        assert (g.ivnum, g.ovnum, g.fid, g.fnum, g.oe_num, g.ie_num) == (f.ivnum, f.ovnum, f.fid, f.fnum, f.oe_num, f.ie_num)
        for which in (0, 1, 2):
            a, b = f.csr(which), g.csr(which)
            for x, y in zip(a, b):
                assert (x is None and y is None) or np.array_equal(x, y)
        assert np.array_equal(f.ovgid(), g.ovgid())
        if k < 2:
            src_oid = int(oids[0]) if k == 1 else 0
            ra, rb = [], []
            for frag, out in ((f, ra), (g, rb)):
                app = P.App("sssp", frag, source_oid=src_oid)
                app.query()
                out.append(app.result())
                out.append(app.result_oids())
                app.close()
            assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
        g.close()
        f.close()


def test_device_vertex_map_lookups():
    """gl_vm_*: oid <-> gid for ascending slices (binary search over l2o itself) and for
    hash-partition style slices (sorted side index), unknown oids and invalid gids."""
    P = pkg()
    rng = np.random.default_rng(3)
    all_oids = rng.choice(10**9, size=5000, replace=False).astype(np.int64)
    for ascending in (True, False):
        parts = np.array_split(all_oids, 3)
        if ascending:
            parts = [np.sort(p) for p in parts]
        vm = P.VertexMap(parts)
        gids = vm.oid2gid(all_oids)
        off = 30   # fnum = 3 -> 2 fid bits
        want = np.concatenate([(f << off) | np.arange(len(p)) for f, p in enumerate(parts)]).astype(np.uint32)
        lookup = {int(o): int(g) for p_, f in zip(parts, range(3)) for o, g in zip(p_, ((f << off) | np.arange(len(p_))))}
        assert np.array_equal(gids, np.array([lookup[int(o)] for o in all_oids], dtype=np.uint32))
        assert np.array_equal(vm.gid2oid(want), np.concatenate(parts))
        assert np.all(vm.oid2gid(np.array([-7, 10**12], dtype=np.int64)) == 0xFFFFFFFF)
        assert np.all(vm.gid2oid(np.array([(2 << off) | 4999, 3 << off], dtype=np.uint32)) == -1)
        vm.close()

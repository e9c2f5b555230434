"""GPU parity of the PIE apps, through the C ABI, against the golden vectors
and the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from oracle import pyoracle
from tests import golden_io as G
from tests.util import pkg, rmat_graph

pytestmark = pytest.mark.gpu

INT64_MAX = np.iinfo(np.int64).max


@pytest.fixture(scope="module")
def p2p():
    oids, src, dst, w = G.load_p2p31()
    P = pkg()
    und = P.Fragment.from_edges(len(oids), src, dst, w, directed=False, oids=oids)
    dr = P.Fragment.from_edges(len(oids), src, dst, w, directed=True, oids=oids)
    yield oids, und, dr
    und.close()
    dr.close()


def app_available(kind, frag, **cfg):
    try:
        return pkg().App(kind, frag, **cfg)
    except pkg().GrapeError as e:
        if "not available" in str(e):
            pytest.skip("%s app not built yet" % kind)
        raise


@pytest.mark.parametrize("directed,name", [(False, "p2p-31-BFS"), (True, "p2p-31-BFS-directed")])
@pytest.mark.parametrize("dopt", [0, 1])
@pytest.mark.parametrize("fuse", [0, 1])
def test_bfs_golden(p2p, directed, name, dopt, fuse):
    oids, und, dr = p2p
    frag = dr if directed else und
    app = app_available("bfs", frag, source_oid=6, direction_opt=dopt, fuse_supersteps=fuse)
    app.query()
    depth = app.result()
    assert np.array_equal(app.result_oids(), oids)
    assert G.render(oids, [str(int(d)) for d in depth]) == G.golden_lines(name)
    # a second query on the same app must reset its state
    app.query()
    assert np.array_equal(app.result(), depth)
    app.close()


@pytest.mark.parametrize("scale", [10, 16])
@pytest.mark.parametrize("dopt", [0, 1])
@pytest.mark.parametrize("fuse", [0, 1])
def test_bfs_rmat_vs_oracle(scale, dopt, fuse):
    n, src, dst, _ = rmat_graph(scale, seed=1)
    g = pyoracle.Graph(n, src, dst, None)
    frag = pkg().Fragment.rmat(scale, 16, seed=1)
    for source in (g.max_degree_vertex(), 0, n - 1):
        app = app_available("bfs", frag, source_oid=int(source), direction_opt=dopt, fuse_supersteps=fuse)
        st = app.query()
        want, _ = g.bfs(source)
        assert np.array_equal(app.result(), want)
        assert st.supersteps >= 2 and st.kernel_launches > 0
        app.close()
    frag.close()


@pytest.mark.parametrize("scale", [9, 21])
def test_bfs_result_paths_agree(scale):
    """gl_app_result for BFS: the default path (u8 depths over PCIe in chunks, widened to int64 by
    host threads) and the plain int64 device array (cfg.reserved[6] = 1) return the same array."""
    frag = pkg().Fragment.rmat(scale, 16, seed=3)
    source, _ = frag.max_degree_vertex()
    out = []
    for flag in (0, 1):
        app = app_available("bfs", frag, source_oid=int(source), reserved={6: flag})
        app.query()
        out.append(app.result())
        app.close()
    assert np.array_equal(out[0], out[1])
    assert out[0][source] == 0 and np.any(out[0] == INT64_MAX)
    frag.close()


def test_bfs_source_not_in_graph_and_isolated():
    n, src, dst, _ = rmat_graph(8, seed=4)
    g = pyoracle.Graph(n, src, dst, None)
    frag = pkg().Fragment.rmat(8, 16, seed=4)
    app = app_available("bfs", frag, source_oid=10**9)   # unknown oid: nothing reached
    app.query()
    assert np.all(app.result() == INT64_MAX)
    app.close()
    rp, _, _ = g.csr()
    iso = np.where(np.diff(rp) == 0)[0]
    if len(iso):
        app = app_available("bfs", frag, source_oid=int(iso[0]))
        app.query()
        r = app.result()
        assert r[iso[0]] == 0 and np.sum(r != INT64_MAX) == 1
        app.close()
    frag.close()


@pytest.mark.parametrize("fuse", [0, 1])
@pytest.mark.parametrize("ring", [0, 5, 16])
def test_bfs_high_diameter_spills_level_ring(fuse, ring):
    """A 1000-vertex path with side branches: depth far beyond a tiny level-bitmap
    ring (cfg.reserved[3]) => the finished levels are spilled to the int32 depth
    array and the query resumes; like the reference's depth array (bfs.h:31) the
    depth is unbounded."""
    n = 1000
    src = np.arange(0, n - 1, dtype=np.int64)
    dst = src + 1
    extra_s = np.array([10, 10, 500, 998, 3], dtype=np.int64)
    extra_d = np.array([700, 11, 502, 0, 3], dtype=np.int64)     # chords, a multi-edge and a self loop
    s_all = np.concatenate([src, extra_s])
    d_all = np.concatenate([dst, extra_d])
    g = pyoracle.Graph(n, s_all, d_all, None)
    frag = pkg().Fragment.from_edges(n, s_all, d_all)
    for source in (0, 640):
        app = app_available("bfs", frag, source_oid=source, fuse_supersteps=fuse, reserved={3: ring})
        app.query()
        want, _ = g.bfs(source)
        assert want.max() > 100
        assert np.array_equal(app.result(), want)
        app.query()                      # state (ring + spill array) resets between queries
        assert np.array_equal(app.result(), want)
        app.close()
    frag.close()


# ------------------------------------------------------------------ SSSP ----
def _sssp_render(oids, dist):
    big = np.finfo(np.float64).max
    return G.render(oids, ["infinity" if d == big else G.fmt_sci(d) for d in dist])


@pytest.mark.parametrize("directed,name", [(False, "p2p-31-SSSP"), (True, "p2p-31-SSSP-directed")])
@pytest.mark.parametrize("f64", [0, 1])
def test_sssp_golden(p2p, directed, name, f64):
    oids, und, dr = p2p
    app = app_available("sssp", dr if directed else und, source_oid=6, sssp_f64=f64)
    app.query()
    dist = app.result()
    assert _sssp_render(oids, dist) == G.golden_lines(name)
    app.query()
    assert np.array_equal(app.result(), dist)
    app.close()


@pytest.mark.parametrize("scale,wmode,f64", [(10, 1, 0), (15, 1, 0), (15, 2, 1), (12, 0, 0)])
def test_sssp_rmat_vs_oracle(scale, wmode, f64):
    """Integer weights: f32 sums are exact => bit-exact vs the fp64 oracle.
    Real weights (multiples of 2^-24): fp64 on device => bit-exact as well."""
    n, src, dst, w = rmat_graph(scale, seed=3, weight_mode=wmode)
    g = pyoracle.Graph(n, src, dst, None if w is None else w.astype(np.float64))
    frag = pkg().Fragment.rmat(scale, 16, seed=3, weight_mode=wmode)
    source = g.max_degree_vertex()
    app = app_available("sssp", frag, source_oid=int(source), sssp_f64=f64)
    app.query()
    want, _ = g.sssp(source)
    got = app.result()
    assert np.array_equal(got, want)
    app.close()
    frag.close()


def test_sssp_real_weights_f32_within_tolerance():
    n, src, dst, w = rmat_graph(14, seed=8, weight_mode=2)
    g = pyoracle.Graph(n, src, dst, w.astype(np.float64))
    frag = pkg().Fragment.rmat(14, 16, seed=8, weight_mode=2)
    source = g.max_degree_vertex()
    app = app_available("sssp", frag, source_oid=int(source))
    app.query()
    want, _ = g.sssp(source)
    got = app.result()
    fin = want < 1e300
    assert np.array_equal(got >= 1e300, ~fin)
    assert np.max(np.abs(got[fin] - want[fin]) / np.maximum(want[fin], 1e-30)) < 1e-6  # north-star tolerance
    app.close()
    frag.close()


# ------------------------------------------------------------------- WCC ----
@pytest.mark.parametrize("kind", ["wcc", "wcc_opt"])
def test_wcc_golden(p2p, kind):
    oids, und, dr = p2p
    want = np.array([int(v) for _, v in G.golden_pairs("p2p-31-WCC")])
    for frag in (und, dr):
        app = app_available(kind, frag)
        app.query()
        lab = app.result()
        assert G.same_partition(lab, want)          # misc/wcc_check.cc rule
        # label = min oid of the component (CPU app convention, wcc.h:139-153)
        assert np.all(lab <= oids)
        app.close()


@pytest.mark.parametrize("kind", ["wcc", "wcc_opt"])
@pytest.mark.parametrize("scale", [10, 15])
def test_wcc_rmat_vs_oracle(scale, kind):
    n, src, dst, _ = rmat_graph(scale, seed=6)
    g = pyoracle.Graph(n, src, dst, None)
    frag = pkg().Fragment.rmat(scale, 16, seed=6)
    app = app_available(kind, frag)
    app.query()
    want, _ = g.wcc()
    assert np.array_equal(app.result(), want.astype(np.int64))   # bit-exact min labels
    app.query()                                                   # a second query resets the state
    assert np.array_equal(app.result(), want.astype(np.int64))
    app.close()
    frag.close()


@pytest.mark.parametrize("kind", ["wcc", "wcc_opt"])
@pytest.mark.parametrize("directed", [False, True])
def test_wcc_many_components(kind, directed):
    """Two large components (one of them a star whose centre has the LARGEST id),
    a long path, small cliques, isolated vertices, self loops and a multi-edge:
    the union-find app must not depend on one giant component being present."""
    rng = np.random.RandomState(5)
    parts_s, parts_d = [], []
    a = rng.randint(0, 3000, size=12000)                       # dense random blob on [0, 3000)
    b = rng.randint(0, 3000, size=12000)
    parts_s.append(a); parts_d.append(b)
    parts_s.append(np.full(2999, 6999)); parts_d.append(np.arange(4000, 6999))   # star, centre 6999
    parts_s.append(np.arange(7000, 7999)); parts_d.append(np.arange(7001, 8000))  # path 7000..7999
    for base in range(8000, 8100, 4):                           # 4-cliques
        for i in range(4):
            for j in range(i + 1, 4):
                parts_s.append(np.array([base + j])); parts_d.append(np.array([base + i]))
    parts_s.append(np.array([8200, 8201, 8201])); parts_d.append(np.array([8200, 8202, 8202]))  # loop + multi-edge
    src = np.concatenate(parts_s).astype(np.int64)
    dst = np.concatenate(parts_d).astype(np.int64)
    n = 8300                                                    # 8203.. are isolated
    g = pyoracle.Graph(n, src, dst, None)                       # WCC ignores direction
    frag = pkg().Fragment.from_edges(n, src, dst, directed=directed)
    app = app_available(kind, frag)
    app.query()
    want, _ = g.wcc()
    assert np.array_equal(app.result(), want.astype(np.int64))
    app.close()
    frag.close()


# -------------------------------------------------------------- PageRank ----
def _pr_cfg(pull):
    """0 = push (f64 atomics), 1 = pull, 2 = pull gathering f32 contributions (f64 sums)"""
    return dict(pr_pull=1 if pull else 0, reserved={5: 1} if pull == 2 else {})


@pytest.mark.parametrize("pull", [0, 1, 2])
def test_pagerank_golden(p2p, pull):
    oids, und, _ = p2p
    want = np.array([float(v) for _, v in G.golden_pairs("p2p-31-PR")])
    app = app_available("pagerank", und, pr_delta=0.85, max_round=10, **_pr_cfg(pull))
    app.query()
    got = app.result()
    assert G.eps_check(got, want, 1e-4)              # the reference's own check
    assert np.max(np.abs(got - want) / want) < 1e-6  # north-star tolerance
    app.close()


@pytest.mark.parametrize("pull", [0, 1, 2])
def test_pagerank_rmat_vs_oracle(pull):
    scale = 14
    n, src, dst, _ = rmat_graph(scale, seed=2)
    g = pyoracle.Graph(n, src, dst, None)
    frag = pkg().Fragment.rmat(scale, 16, seed=2)
    app = app_available("pagerank", frag, pr_delta=0.85, max_round=10, **_pr_cfg(pull))
    st = app.query()
    got = app.result()
    for mode in (0, 1):
        want = g.pagerank(0.85, 10, mode)
        assert np.max(np.abs(got - want) / want) < 1e-6
    assert abs(got.sum() - 1.0) < (1e-6 if pull == 2 else 1e-9)
    assert st.supersteps == 12        # PEval + 10 updates + the final message round
    app.close()
    frag.close()


@pytest.mark.parametrize("pull", [0, 1])
def test_pagerank_directed_golden(p2p, pull):
    """p2p-31-PR-directed (misc/app_tests.sh PageRank on the directed load): push along out-edges;
    pr_pull on a directed fragment falls back to push instead of gathering along out-edges."""
    oids, _, dr = p2p
    want = np.array([float(v) for _, v in G.golden_pairs("p2p-31-PR-directed")])
    app = app_available("pagerank", dr, pr_delta=0.85, max_round=10, pr_pull=pull)
    app.query()
    got = app.result()
    assert G.eps_check(got, want, 1e-4)
    assert np.max(np.abs(got - want) / want) < 1e-6
    app.close()


@pytest.mark.parametrize("fuse", [0, 1])
@pytest.mark.parametrize("dopt", [0, 1])
def test_bfs_directed_rmat_vs_oracle(dopt, fuse):
    """Directed graph: push walks oe, the pull levels walk the real ie CSR (bfs.h:225-238)."""
    scale = 15
    n, src, dst, _ = rmat_graph(scale, seed=9)
    g = pyoracle.Graph(n, src, dst, None, directed=True)
    frag = pkg().Fragment.from_edges(n, src, dst, directed=True)
    rp, _, _ = frag.csr(0)
    source = int(np.argmax(np.diff(rp.astype(np.int64))))
    app = app_available("bfs", frag, source_oid=source, direction_opt=dopt, fuse_supersteps=fuse)
    st = app.query()
    want, _ = g.bfs(source)
    assert np.array_equal(app.result(), want)
    if dopt:
        assert any(st.step_mode[i] == 1 for i in range(st.n_steps))     # a pull level really ran
    app.close()
    frag.close()


def test_directed_fragment_without_ie_is_safe():
    """gl_frag_create with directed = 1 and no ie CSR (kOnlyOut): BFS must stay push-only (correct
    levels), WCC must refuse instead of returning components of the out-edge graph."""
    n, src, dst, _ = rmat_graph(12, seed=5)
    g = pyoracle.Graph(n, src, dst, None, directed=True)
    built = pkg().Fragment.from_edges(n, src, dst, directed=True)
    rp, col, _ = built.csr(0)
    built.close()
    frag = pkg().Fragment.from_csr(n, rp, col, directed=True)          # only oe is passed
    source = int(np.argmax(np.diff(rp.astype(np.int64))))
    for fuse in (0, 1):
        app = app_available("bfs", frag, source_oid=source, direction_opt=1, fuse_supersteps=fuse)
        st = app.query()
        assert np.array_equal(app.result(), g.bfs(source)[0])
        assert all(st.step_mode[i] == 0 for i in range(st.n_steps))
        app.close()
    for kind in ("wcc", "wcc_opt"):
        with pytest.raises(pkg().GrapeError):
            pkg().App(kind, frag)
    frag.close()


def test_sssp_rejects_negative_weights():
    n, src, dst, w = rmat_graph(8, seed=2, weight_mode=1)
    w = w.copy()
    w[3] = -1.0
    frag = pkg().Fragment.from_edges(n, src, dst, w)
    with pytest.raises(pkg().GrapeError):
        pkg().App("sssp", frag, source_oid=0)
    frag.close()


# ------------------------------------------------------------------ CDLP ----
def test_cdlp_golden(p2p):
    oids, und, _ = p2p
    app = app_available("cdlp", und, max_round=10)
    app.query()
    lab = app.result()
    assert G.render(oids, [str(int(x)) for x in lab]) == G.golden_lines("p2p-31-CDLP")
    app.close()


@pytest.mark.parametrize("scale,rounds", [(10, 10), (14, 5), (8, 0), (8, 1)])
def test_cdlp_rmat_vs_oracle(scale, rounds):
    n, src, dst, _ = rmat_graph(scale, seed=11)
    g = pyoracle.Graph(n, src, dst, None)
    frag = pkg().Fragment.rmat(scale, 16, seed=11)
    app = app_available("cdlp", frag, max_round=rounds)
    app.query()
    assert np.array_equal(app.result(), g.cdlp(rounds))
    app.close()
    frag.close()


# ------------------------------------------------------------------- LCC ----
def test_lcc_golden(p2p):
    oids, und, _ = p2p
    app = app_available("lcc", und)
    app.query()
    lcc = app.result()
    assert G.render(oids, [G.fmt_sci(x) for x in lcc]) == G.golden_lines("p2p-31-LCC")
    app.close()


@pytest.mark.parametrize("scale", [8, 12])
def test_lcc_rmat_vs_oracle(scale):
    """R-MAT keeps duplicate edges and self loops, so this pins the
    multiplicity rules of lcc.h:96-190."""
    n, src, dst, _ = rmat_graph(scale, seed=13)
    g = pyoracle.Graph(n, src, dst, None)
    frag = pkg().Fragment.rmat(scale, 16, seed=13)
    app = app_available("lcc", frag)
    app.query()
    want, _ = g.lcc()
    assert np.array_equal(app.result(), want)    # same integers, same fp64 division
    app.close()
    frag.close()

"""Pins the CPU oracle against the reference's golden vectors
(dataset/p2p-31-*, parameters of misc/app_tests.sh:54-109)."""
import numpy as np
import pytest

from oracle import pyoracle
from tests import golden_io as G

INT64_MAX = np.iinfo(np.int64).max


@pytest.fixture(scope="module")
def p2p():
    oids, src, dst, w = G.load_p2p31()
    und = pyoracle.Graph(len(oids), src, dst, w, directed=False, oids=oids)
    dr = pyoracle.Graph(len(oids), src, dst, w, directed=True, oids=oids)
    return oids, und, dr


def test_graph_shape(p2p):
    oids, und, dr = p2p
    assert len(oids) == 62586
    assert und.entries == 295784      # M = 2E
    assert dr.entries == 147892


@pytest.mark.parametrize("directed,name", [(False, "p2p-31-BFS"), (True, "p2p-31-BFS-directed")])
def test_bfs(p2p, directed, name):
    oids, und, dr = p2p
    g = dr if directed else und
    depth, _ = g.bfs(g.index_of(6))
    assert G.render(oids, [str(int(d)) for d in depth]) == G.golden_lines(name)


@pytest.mark.parametrize("directed,name", [(False, "p2p-31-SSSP"), (True, "p2p-31-SSSP-directed")])
def test_sssp(p2p, directed, name):
    oids, und, dr = p2p
    g = dr if directed else und
    dist, _ = g.sssp(g.index_of(6))
    vals = ["infinity" if d == np.finfo(np.float64).max else G.fmt_sci(d) for d in dist]
    assert G.render(oids, vals) == G.golden_lines(name)


def test_pagerank_undirected(p2p):
    oids, und, _ = p2p
    want = np.array([float(v) for _, v in G.golden_pairs("p2p-31-PR")])
    for mode in (0, 1):
        got = und.pagerank(0.85, 10, mode)
        assert G.eps_check(got, want, 1e-4)
        assert np.max(np.abs(got - want) / want) < 1e-9, mode


def test_pagerank_directed(p2p):
    oids, _, dr = p2p
    want = np.array([float(v) for _, v in G.golden_pairs("p2p-31-PR-directed")])
    got = dr.pagerank(0.85, 10, 2)
    assert G.eps_check(got, want, 1e-4)
    assert np.max(np.abs(got - want) / want) < 1e-9


def test_cdlp(p2p):
    oids, und, _ = p2p
    lab = und.cdlp(10)
    assert G.render(oids, [str(int(x)) for x in lab]) == G.golden_lines("p2p-31-CDLP")


def test_lcc(p2p):
    oids, und, _ = p2p
    lcc, _ = und.lcc()
    assert G.render(oids, [G.fmt_sci(x) for x in lcc]) == G.golden_lines("p2p-31-LCC")


def test_wcc(p2p):
    oids, und, dr = p2p
    want = np.array([int(v) for _, v in G.golden_pairs("p2p-31-WCC")])
    for g in (und, dr):
        lab, _ = g.wcc()
        assert G.same_partition(lab, want)
        # label is the minimum index of its component
        assert np.all(lab <= np.arange(len(lab)))
        assert np.all(lab[lab] == lab)

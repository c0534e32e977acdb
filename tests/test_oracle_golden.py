"""Pins the CPU oracle against every known-answer row the reference's own tests hold for the
hot path. Table rows are the reference's:
  internal/core/algorithm/distance_test.go   (L2 10-64, mismatch 66-72, cosine 74-160, IP 162-209,
                                              factory 252-285, BatchDistance 287-310, normalise 312-366,
                                              magnitude 368-409, dot 411-451)
  internal/core/algorithm/hnsw_test.go       (47-75, 77-122, 124-160, 162-219, 221-240, 242-278, 280-309)
  internal/core/algorithm/hnsw_graph_state_test.go (92-107)
  internal/server/grpc/vector_ops_test.go    (118-208)
"""
import math

import numpy as np
import pytest

import oracle
from oracle import METRIC_COSINE, METRIC_IP, METRIC_L2, OracleError, OracleHNSW

SQ = math.sqrt


# ---- distance_test.go ----------------------------------------------------------------------

@pytest.mark.parametrize("a,b,want,delta", [
    ([1, 2, 3], [1, 2, 3], 0.0, 1e-6),
    ([1, 0], [0, 1], np.float32(SQ(2)), 1e-6),
    ([1, 1], [4, 5], 5.0, 1e-6),
    ([-1, -2], [1, 2], np.float32(SQ(20)), 1e-6),
    ([0, 0, 0], [0, 0, 0], 0.0, 1e-6),
])
def test_l2_table(a, b, want, delta):
    assert abs(float(oracle.distance(METRIC_L2, a, b)) - float(want)) <= delta


@pytest.mark.parametrize("metric", [METRIC_L2, METRIC_COSINE, METRIC_IP])
def test_mismatched_dimensions_inf(metric):
    assert oracle.distance(metric, [1, 2], [1, 2, 3]) == np.inf


@pytest.mark.parametrize("a,b,want,delta", [
    ([1, 2, 3], [1, 2, 3], 0.0, 1e-6),
    ([1, 0], [0, 1], 1.0, 1e-6),
    ([1, 0], [-1, 0], 2.0, 1e-6),
    ([1, 2], [2, 4], 0.0, 1e-6),
    ([1, 0], [0.5, np.float32(SQ(3) / 2)], 0.5, 1e-5),
])
def test_cosine_table(a, b, want, delta):
    assert abs(float(oracle.distance(METRIC_COSINE, a, b)) - want) <= delta


@pytest.mark.parametrize("a,b", [([0, 0], [0, 0]), ([1, 2], [0, 0])])
def test_cosine_zero_vectors_exactly_one(a, b):
    assert oracle.distance(METRIC_COSINE, a, b) == np.float32(1.0)


@pytest.mark.parametrize("a,b,want", [
    ([1, 2, 3], [1, 1, 1], -6.0),
    ([1, 2], [-1, -1], 3.0),
    ([1, 0], [0, 1], 0.0),
    ([2, 3], [2, 3], -13.0),
])
def test_ip_table(a, b, want):
    assert abs(float(oracle.distance(METRIC_IP, a, b)) - want) <= 1e-6


@pytest.mark.parametrize("metric,ok", [(1, True), (2, True), (3, True), (0, False), (999, False)])
def test_new_distance_calculator(metric, ok):
    if ok:
        oracle.distance(metric, [1.0], [1.0])
    else:
        with pytest.raises(OracleError) as e:
            oracle.distance(metric, [1.0], [1.0])
        assert e.value.code == 3007


def test_batch_distance():
    got = oracle.batch_distance(METRIC_L2, [0, 0], [[1, 0], [0, 1], [1, 1], [2, 2]])
    want = np.array([1, 1, SQ(2), SQ(8)], np.float32)
    assert np.all(np.abs(got - want) <= 1e-6)


@pytest.mark.parametrize("v,want", [
    ([1, 0], [1, 0]),
    ([3, 4], [0.6, 0.8]),
    ([0, 0], [0, 0]),
    ([-1, 1], [-1 / np.float32(SQ(2)), 1 / np.float32(SQ(2))]),
])
def test_normalize(v, want):
    got = oracle.normalize(v)
    assert np.all(np.abs(got - np.array(want, np.float32)) <= 1e-6)
    if any(v):
        assert abs(float(oracle.magnitude(got)) - 1.0) <= 1e-6


@pytest.mark.parametrize("v,want", [([1, 0], 1.0), ([3, 4], 5.0), ([0, 0, 0], 0.0), ([-1, -1], SQ(2))])
def test_magnitude(v, want):
    assert abs(float(oracle.magnitude(v)) - want) <= 1e-6


@pytest.mark.parametrize("a,b,want", [
    ([1, 0], [0, 1], 0.0), ([1, 2, 3], [1, 2, 3], 14.0), ([1, -2, 3], [4, 5, 6], 12.0), ([1, 2], [1, 2, 3], 0.0),
])
def test_dot(a, b, want):
    assert float(oracle.dot(a, b)) == want


def test_sequential_fp32_accumulation_is_what_is_restated():
    # The restatement must round like a scalar float32 loop (no FMA, no reassociation):
    rng = np.random.default_rng(7)
    a = rng.standard_normal(768).astype(np.float32)
    b = rng.standard_normal(768).astype(np.float32)
    s = np.float32(0)
    for x, y in zip(a, b):
        d = np.float32(x - y)
        s = np.float32(s + np.float32(d * d))
    assert oracle.distance(METRIC_L2, a, b) == np.float32(np.sqrt(np.float64(s)))
    dot = na = nb = np.float32(0)
    for x, y in zip(a, b):
        dot = np.float32(dot + np.float32(x * y))
        na = np.float32(na + np.float32(x * x))
        nb = np.float32(nb + np.float32(y * y))
    assert oracle.distance(METRIC_IP, a, b) == -dot
    na, nb = np.float32(np.sqrt(np.float64(na))), np.float32(np.sqrt(np.float64(nb)))
    cs = np.float32(dot / np.float32(na * nb))
    assert oracle.distance(METRIC_COSINE, a, b) == np.float32(np.float32(1) - cs)


# ---- hnsw_test.go --------------------------------------------------------------------------

def _hnsw(metric=METRIC_L2, **kw):
    p = dict(M=16, ef_construction=200, ef_search=50, max_layers=16, seed=42)
    p.update(kw)
    return OracleHNSW(metric=metric, **p)


def test_new_hnsw_invalid_metric():
    with pytest.raises(OracleError):
        OracleHNSW(metric=0)


def test_empty_index():
    h = _hnsw()
    assert h.size() == 0 and h.layers() == 0
    ids, ds = h.search([1.0, 2.0], 5)
    assert len(ids) == 0


def test_single_vector():
    h = _hnsw()
    h.insert(1, [1.0, 2.0, 3.0])
    assert h.size() == 1
    ids, ds = h.search([1.1, 2.1, 3.1], 1)
    assert list(ids) == [1]


def test_multiple_vectors_sorted():
    h = _hnsw()
    h.build(np.array([[1, 0], [0, 1], [1, 1], [2, 2]], np.float32))
    assert h.size() == 4
    ids, ds = h.search([0.0, 0.0], 2)
    assert len(ids) == 2 and ds[0] <= ds[1]
    assert set(ids) == {1, 2} and np.all(ds == np.float32(1.0))


def test_delete():
    h = _hnsw()
    h.build(np.array([[1, 0], [0, 1], [1, 1]], np.float32))
    h.delete(2)
    assert h.size() == 2
    ids, _ = h.search([0.0, 1.0], 3)
    assert 2 not in ids
    with pytest.raises(OracleError) as e:
        h.delete(12345)
    assert e.value.code == 3004
    h.delete(2)  # already deleted: succeeds
    assert h.size() == 2


def test_duplicate_insert():
    h = _hnsw()
    h.insert(1, [1.0, 2.0])
    with pytest.raises(OracleError):
        h.insert(1, [1.0, 2.0])


@pytest.mark.parametrize("metric", [METRIC_L2, METRIC_COSINE, METRIC_IP])
def test_different_metrics(metric):
    h = _hnsw(metric)
    h.build(np.array([[1, 0], [0, 1]], np.float32))
    ids, _ = h.search([1.0, 0.0], 1)
    assert len(ids) == 1 and ids[0] == 1


def test_set_ef_search():
    h = _hnsw(M=32, ef_construction=400, ef_search=100, max_layers=10, seed=12345)
    h.set_ef_search(200)
    assert h.params["ef_search"] == 200


def test_topk_larger_than_ef_returns_only_ef():
    # hnsw.go:319-347 — result count <= min(TopK, ef, reachable)
    rng = np.random.default_rng(0)
    h = _hnsw()
    h.build(rng.standard_normal((200, 8)).astype(np.float32))
    ids, _ = h.search(rng.standard_normal(8).astype(np.float32), 50, ef_search=7)
    assert len(ids) == 7


def test_grpc_search_case():
    # vector_ops_test.go:118-208 — query (1,0.1,0) over the three unit axes, k=2, distance >= 0
    h = _hnsw()
    h.build(np.eye(3, dtype=np.float32))
    ids, ds = h.search([1.0, 0.1, 0.0], 2)
    assert len(ids) == 2 and ids[0] == 1 and np.all(ds >= 0)


# ---- hnsw_graph_state_test.go ---------------------------------------------------------------

def test_export_import_graph_state_roundtrip():
    p = dict(M=16, ef_construction=200, ef_search=50, max_layers=16, seed=12345)
    h = OracleHNSW(metric=METRIC_L2, **p)
    vecs = np.arange(1, 16, dtype=np.float32).reshape(5, 3)
    h.build(vecs)
    st = h.export_graph_state()
    assert st.size == 5 and len(st.ids) == 5 and st.max_layer > -1 and st.entrypoint != 0
    assert np.array_equal(st.vectors, vecs) and not st.deleted.any()
    h2 = OracleHNSW(metric=METRIC_L2, **p)
    h2.import_graph_state(st)
    assert h2.size() == h.size() and h2.layers() == h.layers()
    st2 = h2.export_graph_state()
    assert np.array_equal(st.edges, st2.edges) and np.array_equal(st.edge_counts, st2.edge_counts)
    a = h.search([1.1, 2.1, 3.1], 3)
    b = h2.search([1.1, 2.1, 3.1], 3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


# ---- structural quirks spelled out in SURVEY.md Appendix A -----------------------------------

def test_levels_follow_one_over_ln2():
    h = _hnsw(seed=1)
    rng = np.random.default_rng(1)
    h.build(rng.standard_normal((4000, 4)).astype(np.float32))
    st = h.export_graph_state()
    lv = st.list_counts - 1
    # P(level >= l) = 2^-l  (mL = 1/ln 2, hnsw.go:460)
    assert abs((lv >= 1).mean() - 0.5) < 0.04
    assert abs((lv >= 2).mean() - 0.25) < 0.03
    assert lv.max() <= 15


def test_degree_caps_and_one_way_edges_allowed():
    h = _hnsw(M=4, ef_construction=32, seed=3)
    rng = np.random.default_rng(3)
    h.build(rng.standard_normal((500, 6)).astype(np.float32))
    st = h.export_graph_state()
    li = 0
    for lc in st.list_counts:
        for l in range(lc):
            assert st.edge_counts[li] <= (8 if l == 0 else 4)
            li += 1


def test_deleted_nodes_are_walls_not_bridges():
    # chain 1-2-3 on a line; deleting 2 must make 3 unreachable from entry 1 (hnsw.go:527-530)
    h = _hnsw(M=1, ef_construction=1)
    h.insert(1, [0.0], level=0)
    h.insert(2, [1.0], level=0)
    h.insert(3, [2.0], level=0)
    h.delete(2)
    ids, _ = h.search([2.0], 3, ef_search=10)
    st = h.export_graph_state()
    # entrypoint is node 1; whatever remains reachable never includes the deleted node
    assert 2 not in ids and st.entrypoint == 1


def test_flat_search_stable_ties_lower_id_first():
    db = np.array([[1, 0], [0, 1], [1, 1], [2, 2], [0, 1]], np.float32)
    ids, ds, counts = oracle.flat_search(METRIC_L2, db, [[0, 0]], 4)
    assert list(ids[0]) == [1, 2, 5, 3] and counts[0] == 4
    ids, ds, counts = oracle.flat_search(METRIC_L2, db, [[0, 0]], 8)
    assert counts[0] == 5 and list(ids[0][5:]) == [0, 0, 0] and np.all(np.isinf(ds[0][5:]))


def test_hnsw_recall_sane_small():
    rng = np.random.default_rng(5)
    db = rng.standard_normal((3000, 16)).astype(np.float32)
    q = rng.standard_normal((50, 16)).astype(np.float32)
    h = _hnsw(seed=42)
    h.build(db)
    gt, _, _ = oracle.flat_search(METRIC_L2, db, q, 10)
    ids, ds, counts, (evals, hops) = h.search_batch(q, 10, ef_search=100, nthreads=2)
    rec = np.mean([len(set(ids[i]) & set(gt[i])) / 10 for i in range(len(q))])
    assert rec > 0.9 and evals > 0 and hops > 0
    one = h.search(q[0], 10, ef_search=100)
    assert np.array_equal(one[0], ids[0]) and np.array_equal(one[1], ds[0])

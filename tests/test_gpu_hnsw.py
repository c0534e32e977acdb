"""GPU parity: batched HNSW search over the oracle-built graph (the graph the reference's
algorithm builds) vs the oracle's own Search on the same graph, same queries, same ef.
Bar (north_star): recall@10 within 0.5 % absolute of the reference at equal efSearch; returned
distances must be the exact reference arithmetic for the returned ids."""
import numpy as np
import pytest

import oracle
from scintirete_b200 import DistanceMetric, GPUHNSWIndex, GraphState, HNSWParams, ScintireteError, SearchParams, Vector
from util import gaussian, recall, to_graph_state

pytestmark = pytest.mark.gpu


def _pair(metric, n, d, M=16, efc=200, seed=42, max_layers=16):
    db = gaussian(n, d, 1234)
    h = oracle.OracleHNSW(M=M, ef_construction=efc, ef_search=50, max_layers=max_layers, seed=seed, metric=int(metric))
    h.build(db)
    g = GPUHNSWIndex(HNSWParams(m=M, ef_construction=efc, ef_search=50, max_layers=max_layers, seed=seed), metric, d)
    g.import_graph_state(to_graph_state(h.export_graph_state(), M))
    return db, h, g


@pytest.mark.parametrize("metric", [DistanceMetric.L2, DistanceMetric.COSINE, DistanceMetric.INNER_PRODUCT])
def test_recall_parity_and_exact_distances(metric):
    n, d, nq, k, ef = 8000, 64, 200, 10, 100
    db, h, g = _pair(metric, n, d)
    q = gaussian(nq, d, 4321)
    gt, _, _ = oracle.flat_search(int(metric), db, q, k, nthreads=8)
    o_ids, o_dist, o_cnt, _ = h.search_batch(q, k, ef, nthreads=8)
    ids, dist, cnt = g.search_batch(q, SearchParams(top_k=k, ef_search=ef))
    r_gpu, r_ref = recall(ids, gt), recall(o_ids, gt)
    assert abs(r_gpu - r_ref) <= 0.005, (r_gpu, r_ref)
    assert np.array_equal(cnt, o_cnt)
    # exact-set agreement with the oracle (expected ~1: only float near-ties can reorder the walk)
    agree = recall(ids, o_ids)
    assert agree >= 0.99, agree
    # every returned distance is the reference's exact arithmetic for that id, list sorted ascending
    for i in range(0, nq, 17):
        want = oracle.batch_distance(int(metric), q[i], db[ids[i].astype(np.int64) - 1])
        assert np.array_equal(dist[i], want)
        assert np.all(np.diff(dist[i]) >= 0)


def test_c1_shape_small_and_default_ef():
    # C1 shape (128-d L2, M=16, efC=200, efS=100) at a size the oracle builds in seconds
    n, d, nq, k = 6000, 128, 100, 10
    db, h, g = _pair(DistanceMetric.L2, n, d)
    q = gaussian(nq, d, 4321)
    h.set_ef_search(100)
    g.set_ef_search(100)
    o_ids, _, _, _ = h.search_batch(q, k, None, nthreads=8)
    ids, _, _ = g.search_batch(q, SearchParams(top_k=k))
    gt, _, _ = oracle.flat_search(1, db, q, k, nthreads=8)
    assert abs(recall(ids, gt) - recall(o_ids, gt)) <= 0.005


def test_topk_larger_than_ef_and_k1():
    db, h, g = _pair(DistanceMetric.L2, 2000, 16)
    q = gaussian(20, 16, 7)
    ids, dist, cnt = g.search_batch(q, SearchParams(top_k=50, ef_search=7))
    assert np.all(cnt == 7) and np.all(ids[:, 7:] == 0) and np.all(np.isinf(dist[:, 7:]))
    o_ids, _, o_cnt, _ = h.search_batch(q, 50, 7)
    assert np.array_equal(cnt, o_cnt) and recall(ids[:, :7], o_ids[:, :7]) >= 0.97
    ids1, _, _ = g.search_batch(q, SearchParams(top_k=1, ef_search=1))
    o1, _, _, _ = h.search_batch(q, 1, 1)
    assert np.mean(ids1 == o1) >= 0.95


def test_deleted_nodes_are_walls():
    db, h, g = _pair(DistanceMetric.L2, 3000, 24)
    dead = np.arange(2, 3000, 5).astype(np.uint64)
    ep = h.entrypoint()
    dead = dead[dead != ep]
    for i in dead:
        h.delete(int(i))
    for i in dead:
        g.delete(str(int(i)))
    assert g.size() == h.size()
    q = gaussian(100, 24, 8)
    o_ids, _, o_cnt, _ = h.search_batch(q, 10, 64, nthreads=4)
    ids, dist, cnt = g.search_batch(q, SearchParams(top_k=10, ef_search=64))
    assert not np.isin(ids, dead).any()
    assert np.array_equal(cnt, o_cnt) and recall(ids, o_ids) >= 0.99


def test_empty_and_single_and_toy_graphs():
    g = GPUHNSWIndex(HNSWParams(), DistanceMetric.L2, 3)
    ids, dist, cnt = g.search_batch(np.zeros((2, 3), np.float32), SearchParams(top_k=5))  # hnsw_test.go:47-75
    assert np.all(cnt == 0)
    h = oracle.OracleHNSW(metric=1)
    h.insert(1, [1.0, 2.0, 3.0])
    g.import_graph_state(to_graph_state(h.export_graph_state(), 16))
    res = g.search([1.1, 2.1, 3.1], SearchParams(top_k=1))                                  # hnsw_test.go:77-122
    assert len(res) == 1 and res[0].vector.id == 1
    h = oracle.OracleHNSW(metric=1)
    h.build(np.array([[1, 0], [0, 1], [1, 1], [2, 2]], np.float32))
    g2 = GPUHNSWIndex(HNSWParams(), DistanceMetric.L2, 2)
    g2.import_graph_state(to_graph_state(h.export_graph_state(), 16))
    res = g2.search([0.0, 0.0], SearchParams(top_k=2))                                      # hnsw_test.go:124-160
    o = h.search([0.0, 0.0], 2)
    assert [r.vector.id for r in res] == list(o[0]) and [np.float32(r.distance) for r in res] == list(o[1])
    g2.build([])                                                                            # Build clears first (hnsw.go:148-156)
    assert g2.size() == 0 and len(g2.search([0.0, 0.0], SearchParams(top_k=2))) == 0
    # Build on the device: the same toy vectors, linked by scn_hnsw_insert with the oracle's level draws
    st = h.export_graph_state()
    g2.build([Vector(i + 1, v) for i, v in enumerate(st.vectors)], levels=st.list_counts - 1)
    res = g2.search([0.0, 0.0], SearchParams(top_k=2))
    assert [r.vector.id for r in res] == list(o[0]) and [np.float32(r.distance) for r in res] == list(o[1])


def test_visited_overflow_path_gives_same_answer():
    # tiny M with a large ef forces far more visits per expansion budget than the shared-memory
    # table is sized for -> exercises the global-memory overflow pass
    db, h, g = _pair(DistanceMetric.L2, 5000, 8, M=48, efc=64)
    q = gaussian(50, 8, 3)
    o_ids, _, o_cnt, _ = h.search_batch(q, 10, 16, nthreads=4)
    ids, _, cnt = g.search_batch(q, SearchParams(top_k=10, ef_search=16))
    assert np.array_equal(cnt, o_cnt) and recall(ids, o_ids) >= 0.99


@pytest.mark.parametrize("metric,d", [(DistanceMetric.L2, 128), (DistanceMetric.COSINE, 96), (DistanceMetric.INNER_PRODUCT, 40)])
def test_walk_is_identical_to_the_reference_walk(metric, d):
    # Stronger than the recall bar: traversal distances are accumulated in the reference's order
    # and admissions follow its stable (distance, admission) order, so on data without exact
    # float ties the GPU returns the very same ids, in the same order, with the same distance bits
    # as hnsw.go:292-350 restated by the oracle, after the same number of expansions.
    n, nq, k, ef = 6000, 300, 10, 64
    db, h, g = _pair(metric, n, d)
    q = gaussian(nq, d, 99)
    o_ids, o_dist, o_cnt, o_stats = h.search_batch(q, k, ef, nthreads=8)
    g.store.set_option("profile", 1)
    ids, dist, cnt = g.search_batch(q, SearchParams(top_k=k, ef_search=ef))
    counters = g.store.last_counters()
    g.store.set_option("profile", 0)
    assert np.array_equal(cnt, o_cnt)
    assert np.array_equal(ids, o_ids)
    assert np.array_equal(dist, o_dist)
    assert counters[1] == o_stats[1], (counters, o_stats)   # expansions, all layers


@pytest.mark.parametrize("d", [128, 200, 768])
@pytest.mark.parametrize("gather,global_tables", [(0, 0), (0, 1), (1, 0), (1, 1), (3, 0), (3, 1)])
def test_walk_is_identical_in_every_gather_and_table_mode(d, gather, global_tables):
    # the kernel's switches (rows through registers / 512-byte / 256-byte shared-memory stages;
    # visited tables in shared or global memory) change how a walk is executed, never the walk:
    # ids, distance bits and expansion counts stay the oracle's. d = 200 and 768 take the ragged
    # last stage and the two-buffer path (rows longer than 512 bytes).
    n, nq, k, ef = 3000, 100, 10, 48
    db, h, g = _pair(DistanceMetric.L2, n, d, efc=60)
    q = gaussian(nq, d, 7)
    o_ids, o_dist, o_cnt, o_stats = h.search_batch(q, k, ef, nthreads=8)
    g.store.set_option("hnsw_gather", gather)
    g.store.set_option("hnsw_global", global_tables)
    g.store.set_option("profile", 1)
    ids, dist, cnt = g.search_batch(q, SearchParams(top_k=k, ef_search=ef))
    counters = g.store.last_counters()
    assert np.array_equal(cnt, o_cnt) and np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist)
    assert counters[1] == o_stats[1], (counters, o_stats)


def test_repeated_neighbours_are_dropped_at_upload_like_visited_rows():
    # a list that names a neighbour twice walks exactly like the list without the repeat
    # (hnsw.go:523-525 skips the second occurrence as visited)
    n, d = 400, 16
    db, h, g = _pair(DistanceMetric.L2, n, d, M=8, efc=40)
    st = h.export_graph_state()
    edges, counts = [], []
    pos = 0
    for c in st.edge_counts:
        lst = list(st.edges[pos:pos + c])
        pos += c
        if 0 < c < 8:            # room for one repeat below the per-layer cap
            lst.append(lst[0])
        edges.extend(lst)
        counts.append(len(lst))
    g2 = GPUHNSWIndex(HNSWParams(m=8, ef_search=32), DistanceMetric.L2, d)
    g2.import_graph_state(GraphState(st.ids, st.list_counts, np.asarray(counts, np.uint32), np.asarray(edges, np.uint64),
                                     st.entrypoint, st.max_layer, st.size, st.deleted, st.vectors, m=8))
    q = gaussian(40, d, 5)
    a = g.search_batch(q, SearchParams(top_k=5, ef_search=32))
    b = g2.search_batch(q, SearchParams(top_k=5, ef_search=32))
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_deleting_the_entry_point_picks_a_new_one_like_the_reference():
    # hnsw.go:280-283 + findNewEntrypoint (617-634): the live node with the highest getNodeLayer
    # takes over and its layer becomes maxLayer (ties: insertion order, as in the oracle; the Go map
    # order is random). Walks after the hand-over stay identical — three hand-overs in a row.
    n, d = 4000, 32
    db, h, g = _pair(DistanceMetric.L2, n, d, efc=100)
    q = gaussian(120, d, 17)
    for _ in range(3):
        ep = h.entrypoint()
        h.delete(int(ep))
        g.delete(str(int(ep)))
        assert h.entrypoint() != ep and g.store.stats().entry_id == h.entrypoint()
        assert g.get_layers() == h.max_layer() + 1
        o_ids, o_dist, o_cnt, _ = h.search_batch(q, 10, 64, nthreads=4)
        ids, dist, cnt = g.search_batch(q, SearchParams(top_k=10, ef_search=64))
        assert np.array_equal(cnt, o_cnt) and np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist)
        assert not np.isin(ids, [ep]).any()


def test_restored_snapshot_keeps_a_deleted_entry_point_verbatim():
    # ImportGraphState takes entrypoint / maxLayer verbatim (hnsw.go:791-793): a snapshot whose entry
    # point is flagged deleted answers every search with nothing (searchLayer drops the deleted entry
    # point, hnsw.go:492-502), in the reference, in the oracle and here — no new entry point is picked.
    n, d = 1500, 24
    db, h, g = _pair(DistanceMetric.L2, n, d, efc=60)
    st = h.export_graph_state()
    st.deleted[int(st.entrypoint) - 1] = 1          # ids are 1..n in insertion order
    st.deleted[7] = 1
    h2 = oracle.OracleHNSW(M=16, ef_construction=60, ef_search=32, max_layers=16, seed=42, metric=1)
    h2.import_graph_state(st)
    g2 = GPUHNSWIndex(HNSWParams(m=16, ef_search=32), DistanceMetric.L2, d)
    g2.import_graph_state(to_graph_state(st, 16))
    assert g2.store.stats().entry_id == st.entrypoint and g2.get_layers() == st.max_layer + 1
    q = gaussian(20, d, 3)
    o_ids, o_dist, o_cnt, _ = h2.search_batch(q, 5, 32)
    ids, dist, cnt = g2.search_batch(q, SearchParams(top_k=5, ef_search=32))
    assert not o_cnt.any() and np.array_equal(cnt, o_cnt) and np.array_equal(ids, o_ids)
    # the raw C-ABI get hides soft-deleted rows like HNSW.Get (hnsw.go:364-366)
    with pytest.raises(ScintireteError) as e:
        g2.store.get([8])
    assert e.value.code == 3004
    assert np.array_equal(g2.store.get([9])[0], db[8])


@pytest.mark.parametrize("n,d,hi,M,efc,metric,ef", [(3000, 8, 3, 8, 40, 1, 32), (3000, 16, 2, 16, 64, 1, 64), (4000, 12, 4, 16, 100, 2, 64),
                                                    (3000, 24, 3, 16, 200, 3, 128), (500, 4, 1, 4, 16, 1, 8)])
def test_search_follows_the_reference_through_exact_distance_ties(n, d, hi, M, efc, metric, ef):
    # small integer coordinates: most distances tie exactly (last case: all vectors identical). The fast walk
    # flags every query whose W dropped an entry that ties with W[ef-1]; those are redone by the exact walk
    # (ghost candidates, sequential admission — hnsw.go:516-518, 536-542), so ids, distances and counts are
    # the reference's even here
    db = np.random.default_rng(5).integers(0, hi, (n, d)).astype(np.float32) + (0.0 if metric == 1 else 1.0)
    q = np.random.default_rng(6).integers(0, hi, (300, d)).astype(np.float32) + (0.0 if metric == 1 else 1.0)
    h = oracle.OracleHNSW(M=M, ef_construction=efc, ef_search=ef, max_layers=16, seed=42, metric=metric)
    h.build(db)
    g = GPUHNSWIndex(HNSWParams(m=M, ef_search=ef), DistanceMetric(metric), d)
    g.import_graph_state(to_graph_state(h.export_graph_state(), M))
    g.store.set_option("hnsw_exact_ties", 1)
    for k in (10, ef + 5):
        ids, dist, cnt = g.search_batch(q, SearchParams(top_k=k, ef_search=ef))
        o_ids, o_dist, o_cnt, _ = h.search_batch(q, k, ef, nthreads=4)
        assert np.array_equal(cnt, o_cnt) and np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist)

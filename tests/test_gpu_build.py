"""GPU-assisted HNSW construction (scn_hnsw_insert): the graph must be edge for edge
the one the reference's serial insertVector builds (hnsw.go:190-257, 560-614; restated by the oracle)
for the same level draws — adjacency lists in stored order, entry point, maxLayer — and searches over
the device copy of that graph must be the oracle's."""
import os

import numpy as np
import pytest

import oracle
from scintirete_b200 import DeviceStore, DistanceMetric, GPUHNSWIndex, HNSWParams, ScintireteError, SearchParams, Vector
from util import gaussian, to_graph_state

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle(db, metric, M, efc, ids=None):
    h = oracle.OracleHNSW(M=M, ef_construction=efc, ef_search=64, max_layers=16, seed=42, metric=int(metric))
    h.build(db, ids)
    return h


def _same_graph(a, b):
    """a: oracle.GraphState, b: scintirete_b200.GraphState"""
    assert np.array_equal(a.ids, b.node_ids)
    assert np.array_equal(a.list_counts, b.list_counts)
    assert np.array_equal(a.edge_counts, b.edge_counts), "degree of some list differs"
    assert np.array_equal(a.edges, b.edges), "an adjacency list differs"
    assert a.entrypoint == b.entry_point and a.max_layer == b.max_layer


@pytest.mark.parametrize("metric", [DistanceMetric.L2, DistanceMetric.COSINE, DistanceMetric.INNER_PRODUCT])
@pytest.mark.parametrize("n,d,M,efc,window", [(1500, 24, 8, 40, 1), (1500, 24, 8, 40, 0), (12000, 64, 16, 100, 0),
                                              (3000, 200, 16, 64, 0), (2500, 40, 32, 80, 0)])
def test_built_graph_is_the_reference_graph(metric, n, d, M, efc, window):
    # window = 1: no speculation (every insert is searched on the up-to-date graph): isolates the device
    # walk + host link phase; window = 0 (adaptive): adds the speculative window and its validation
    db = gaussian(n, d, 7)
    h = _oracle(db, metric, M, efc)
    st = h.export_graph_state()
    s = DeviceStore(d, metric)
    s.append(db)
    s.set_option("build_window", window)
    stats = s.hnsw_insert(st.list_counts - 1, M, efc)
    assert stats["inserted"] == n and stats["rounds"] <= n
    g = s.graph_export(M)
    _same_graph(st, g)
    if window == 0:
        assert stats["rounds"] < n / 2, stats      # speculation must actually commit several inserts per round
    # the device copy of the graph is the same graph: walks equal the oracle's walks
    q = gaussian(64, d, 8)
    ids, dist, cnt = s.search_hnsw(q, 10, 64)
    o_ids, o_dist, o_cnt, _ = h.search_batch(q, 10, 64, nthreads=4)
    assert np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist) and np.array_equal(cnt, o_cnt)
    s.close()


def test_incremental_insert_continues_an_uploaded_graph():
    # restore (ImportGraphState) then Insert: the next inserts link into the uploaded graph exactly as the
    # reference's would (edge distances of the uploaded graph are recomputed on the device for the prunes)
    n0, n1, d, M, efc = 4000, 2500, 32, 16, 80
    db = gaussian(n0 + n1, d, 11)
    h_all = _oracle(db, DistanceMetric.L2, M, efc)
    st_all = h_all.export_graph_state()
    h0 = oracle.OracleHNSW(M=M, ef_construction=efc, ef_search=64, max_layers=16, seed=42, metric=1)
    for i in range(n0):   # the same first n0 inserts with the same level draws
        h0.insert(i + 1, db[i], level=int(st_all.list_counts[i]) - 1)
    st0 = h0.export_graph_state()
    g = GPUHNSWIndex(HNSWParams(m=M, ef_construction=efc, ef_search=64), DistanceMetric.L2, d)
    g.import_graph_state(to_graph_state(st0, M))
    for i in range(n0, n0 + 5):   # one at a time (HNSW.Insert) ...
        g.insert(Vector(i + 1, db[i]), level=int(st_all.list_counts[i]) - 1)
    g.store.append(db[n0 + 5:])   # ... and the rest as one batch
    g.store.hnsw_insert(st_all.list_counts[n0 + 5:] - 1, M, efc)
    _same_graph(st_all, g.store.graph_export(M))
    q = gaussian(50, d, 12)
    ids, dist, cnt = g.search_batch(q, SearchParams(top_k=10, ef_search=64))
    o_ids, o_dist, o_cnt, _ = h_all.search_batch(q, 10, 64, nthreads=4)
    assert np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist) and np.array_equal(cnt, o_cnt)


def test_insert_after_soft_deletes_treats_deleted_nodes_as_walls():
    # searchLayer skips deleted neighbours (hnsw.go:527-530) and pruneConnections drops them (596-601)
    n0, n1, d, M, efc = 3000, 1500, 32, 8, 60
    db = gaussian(n0 + n1, d, 13)
    levels = np.minimum(np.floor(-np.log(np.random.default_rng(3).random(n0 + n1)) / np.log(2.0)), 15).astype(np.int32)
    h = oracle.OracleHNSW(M=M, ef_construction=efc, ef_search=64, max_layers=16, seed=42, metric=2)
    for i in range(n0):
        h.insert(i + 1, db[i], level=int(levels[i]))
    dead = np.random.default_rng(4).choice(np.arange(1, n0 + 1), 300, replace=False)
    dead = dead[dead != h.entrypoint()]
    for v in dead:
        h.delete(int(v))
    s = DeviceStore(d, DistanceMetric.COSINE)
    s.append(db[:n0])
    s.hnsw_insert(levels[:n0], M, efc)
    s.mark_deleted(dead.astype(np.uint64))
    for i in range(n0, n0 + n1):
        h.insert(i + 1, db[i], level=int(levels[i]))
    s.append(db[n0:])
    s.hnsw_insert(levels[n0:], M, efc)
    _same_graph(h.export_graph_state(), s.graph_export(M))
    s.close()


def test_build_argument_checks():
    s = DeviceStore(8, DistanceMetric.L2)
    s.append(gaussian(10, 8, 1))
    with pytest.raises(ScintireteError) as e:
        s.hnsw_insert(np.zeros(11, np.int32), 16, 200)       # more inserts than rows outside the graph
    assert e.value.code == 3007
    with pytest.raises(ScintireteError) as e:
        s.hnsw_insert(np.zeros(10, np.int32), 16, 16)        # efConstruction < 2M
    assert e.value.code == 3007
    s.hnsw_insert(np.zeros(10, np.int32), 4, 16)
    assert s.stats().has_graph == 1 and s.stats().entry_id == 1
    s.close()


def _cached(n, d, metric):
    p = os.path.join(ROOT, "bench_cache", f"hnsw_n{n}_d{d}_m{metric}_M16_efc200_s42_db1234.npz")
    return p if os.path.exists(p) else None


@pytest.mark.parametrize("n,d,metric", [(100_000, 128, 1), (100_000, 768, 2)])
def test_full_size_build_equals_the_cached_reference_graph(n, d, metric):
    # BASELINE.json configs[0] (100k x 128 L2, M=16, efC=200) and the embedding-sized 100k x 768 cosine graph:
    # the graphs the oracle built serially (minutes to 46 minutes on one core; cached by bench.py) must come
    # out of the GPU-assisted build edge for edge
    path = _cached(n, d, metric)
    if path is None:
        pytest.skip("bench_cache/ holds no reference-built graph for this configuration")
    import bench

    z = np.load(path)
    db = bench.gen_rows_numpy(0, n, d, bench.SEED_DB)
    s = DeviceStore(d, DistanceMetric(metric))
    s.append(db)
    stats = s.hnsw_insert(z["list_counts"] - 1, 16, 200)
    g = s.graph_export(16)
    assert np.array_equal(g.node_ids, z["ids"]) and np.array_equal(g.list_counts, z["list_counts"])
    assert np.array_equal(g.edge_counts, z["edge_counts"])
    assert np.array_equal(g.edges, z["edges"].astype(np.uint64))
    assert g.entry_point == int(z["entrypoint"]) and g.max_layer == int(z["max_layer"])
    print(f"\nGPU-assisted build {n}x{d}: {stats['seconds']:.1f}s, {stats['rounds']} rounds, "
          f"{stats['inserted'] / max(stats['rounds'], 1):.1f} commits/round, {stats['conflicts']} conflicts")
    s.close()


@pytest.mark.parametrize("n,d,hi,M,efc,metric", [(3000, 8, 3, 8, 40, 1), (3000, 16, 2, 16, 64, 1), (4000, 12, 4, 16, 100, 2),
                                                 (3000, 24, 3, 16, 200, 3), (600, 4, 1, 4, 16, 1)])
@pytest.mark.parametrize("window", [1, 0])
def test_exact_distance_ties_follow_the_reference(n, d, hi, M, efc, metric, window):
    # small integer coordinates: most distances tie exactly (the last case: all vectors identical). The
    # walk must reproduce the reference's tie rules: stable (distance, admission) order, sequential
    # admission inside one adjacency list (strict <, hnsw.go:536-542), and candidates that were pushed out
    # of `candidates` but still equal W[ef-1] stay expandable in `dynamic` (hnsw.go:516-518)
    db = np.random.default_rng(5).integers(0, hi, (n, d)).astype(np.float32) + (0.0 if metric == 1 else 1.0)
    h = _oracle(db, DistanceMetric(metric), M, efc)
    st = h.export_graph_state()
    s = DeviceStore(d, DistanceMetric(metric))
    s.append(db)
    s.set_option("build_window", window)
    s.hnsw_insert(st.list_counts - 1, M, efc)
    _same_graph(st, s.graph_export(M))
    s.close()

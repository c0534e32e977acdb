"""GPU tests of the API-edge pieces: micro-batching of concurrent single-query
Search calls, and Collection.Compact mirrored on the device store."""
import threading

import numpy as np
import pytest

import oracle
from scintirete_b200 import (Batcher, DeviceStore, DistanceMetric, GPUFlatIndex, GPUHNSWIndex, HNSWParams, ScintireteError,
                             SearchParams, Vector)
from util import gaussian, to_graph_state

pytestmark = pytest.mark.gpu


def _run_threads(n_threads, fn):
    errs = []

    def wrap(t):
        try:
            fn(t)
        except Exception as e:  # pragma: no cover - surfaced below
            errs.append(e)

    ts = [threading.Thread(target=wrap, args=(t,)) for t in range(n_threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs[:3]


def test_concurrent_single_query_flat_calls_are_coalesced_and_exact():
    # vector_operations_test.go:345-362 drives Search from concurrent goroutines, one query each
    n, d, nq, k = 20000, 96, 256, 10
    db, q = gaussian(n, d, 1), gaussian(nq, d, 2)
    s = DeviceStore(d, DistanceMetric.COSINE)
    s.append(db)
    want_ids, want_dist, want_cnt = s.search_flat(q, k)
    o = oracle.flat_search(2, db, q, k, nthreads=8)
    assert np.array_equal(want_ids, o[0]) and np.array_equal(want_dist, o[1])
    b = Batcher(s, Batcher.FLAT, max_batch=64, window_us=2000)
    got_ids = np.zeros_like(want_ids)
    got_dist = np.zeros_like(want_dist)
    n_threads = 32

    def work(t):
        for i in range(t, nq, n_threads):
            ids, dist, cnt = b.search(q[i], k)
            got_ids[i], got_dist[i] = ids, dist
            assert cnt == want_cnt[i]

    _run_threads(n_threads, work)
    assert np.array_equal(got_ids, want_ids) and np.array_equal(got_dist, want_dist)
    st = b.stats()
    assert st["calls"] == nq and st["batches"] < nq / 2 and st["max_batch"] > 1, st
    b.close()
    s.close()


def test_hnsw_index_search_through_the_batcher_matches_the_reference_walk():
    n, d, nq, k, ef = 5000, 48, 128, 10, 64
    db = gaussian(n, d, 1234)
    h = oracle.OracleHNSW(M=16, ef_construction=100, ef_search=ef, max_layers=16, seed=42, metric=1)
    h.build(db)
    g = GPUHNSWIndex(HNSWParams(m=16, ef_construction=100, ef_search=ef), DistanceMetric.L2, d)
    g.import_graph_state(to_graph_state(h.export_graph_state(), 16))
    g.enable_micro_batching(max_batch=256, window_us=1000)
    q = gaussian(nq, d, 5)
    o_ids, o_dist, o_cnt, _ = h.search_batch(q, k, ef, nthreads=8)
    out = [None] * nq

    def work(t):
        for i in range(t, nq, 16):
            # mixed (k, ef) requests in flight at once: grouped into separate launches
            kk = k if i % 3 else 5
            out[i] = (kk, g.search(q[i], SearchParams(top_k=kk, ef_search=ef if i % 5 else 32)))

    _run_threads(16, work)
    for i in range(nq):
        kk, res = out[i]
        if i % 5:
            assert [r.vector.id for r in res] == list(o_ids[i][:kk])
            assert [np.float32(r.distance) for r in res] == list(o_dist[i][:kk])
        else:
            ref = h.search(q[i], kk, 32)
            assert [r.vector.id for r in res] == list(ref[0])
    st = g._batcher.stats()
    assert st["calls"] == nq and st["launches"] >= st["batches"]


def test_batcher_argument_errors_reach_the_caller():
    s = DeviceStore(8, DistanceMetric.L2)
    s.append(gaussian(100, 8, 1))
    with pytest.raises(ScintireteError):
        Batcher(s, 7)
    b = Batcher(s, Batcher.HNSW)
    with pytest.raises(ScintireteError):
        b.search(np.zeros(8, np.float32), 10, 0)       # ef = 0
    with pytest.raises(ScintireteError):
        b.search(np.zeros(9, np.float32), 10, 16)      # dimension mismatch (3005)
    ids, dist, cnt = b.search(np.zeros(8, np.float32), 3, 16)   # no graph yet: empty result, hnsw.go:296-298
    assert cnt == 0 and np.all(ids == 0)
    b.close()
    s.close()


@pytest.mark.parametrize("explicit_ids", [False, True])
def test_compact_drops_deleted_rows_and_keeps_results_exact(explicit_ids):
    n, d, k = 9000, 40, 10
    db, q = gaussian(n, d, 3), gaussian(64, d, 4)
    ids = (np.arange(n, dtype=np.uint64) * 7 + 11) if explicit_ids else (np.arange(n, dtype=np.uint64) + 1)
    s = DeviceStore(d, DistanceMetric.L2)
    s.append(db, ids if explicit_ids else None)
    dead_rows = np.arange(0, n, 3)
    s.mark_deleted(ids[dead_rows])
    before = s.search_flat(q, k)
    removed = s.compact()
    st = s.stats()
    assert removed == dead_rows.size and st.rows == st.live_rows == n - dead_rows.size and not st.has_graph
    after = s.search_flat(q, k)
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[1], after[1])
    keep = np.setdiff1d(np.arange(n), dead_rows)
    o = oracle.flat_search(1, db[keep], q, k, nthreads=8)
    assert np.array_equal(after[0], ids[keep][o[0].astype(np.int64) - 1]) and np.array_equal(after[1], o[1])
    assert np.array_equal(s.get(ids[keep[:5]]), db[keep[:5]])
    with pytest.raises(ScintireteError):
        s.get([int(ids[0])])                              # dropped for good
    # the store keeps working: append after compaction, tensor path (large enough store) still exact
    extra = gaussian(500, d, 9)
    new_ids = np.arange(500, dtype=np.uint64) + 10_000_000
    s.append(extra, new_ids)
    s.set_option("flat_path", 2)
    got = s.search_flat(q, k)
    all_db, all_ids = np.concatenate([db[keep], extra]), np.concatenate([ids[keep], new_ids])
    o = oracle.flat_search(1, all_db, q, k, nthreads=8)
    assert np.array_equal(got[0], all_ids[o[0].astype(np.int64) - 1]) and np.array_equal(got[1], o[1])
    assert s.compact() == 0
    s.close()


def test_flat_index_compact_mirrors_collection_compact():
    idx = GPUFlatIndex(4, DistanceMetric.L2)
    idx.build([Vector(i + 1, [float(i), 0, 0, 0], {"n": i}) for i in range(10)])
    idx.delete("3")
    idx.delete("4")
    assert idx.size() == 8
    assert idx.compact() == 2 and idx.size() == 8
    res = idx.search([2.2, 0, 0, 0], SearchParams(top_k=3))
    assert [r.vector.id for r in res] == [2, 5, 1]


def test_concurrent_uncoalesced_searches_with_different_shapes():
    # The reference allows any number of concurrent Search calls under RLock (hnsw.go:293). Calls of
    # different shapes need different amounts of dynamic shared memory; the per-kernel opt-in limit
    # must not be lowered by one thread between another thread's set and its launch.
    n, d = 12000, 64
    db = gaussian(n, d, 1)
    s = DeviceStore(d, DistanceMetric.L2)
    s.append(db)
    shapes = [(1, 1), (3, 10), (40, 24), (7, 100), (130, 5), (2, 300)]
    want = {}
    for nq, k in shapes:
        q = gaussian(nq, d, 100 + nq)
        want[(nq, k)] = (q, oracle.flat_search(1, db, q, k, nthreads=4))

    def work(t):
        for rep in range(12):
            nq, k = shapes[(t + rep) % len(shapes)]
            q, o = want[(nq, k)]
            ids, dist, _ = s.search_flat(q, k)
            assert np.array_equal(ids, o[0]) and np.array_equal(dist, o[1])

    _run_threads(24, work)
    s.close()

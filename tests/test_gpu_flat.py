"""GPU parity: exact flat scan / rerank / batched distance through the C ABI vs the CPU oracle.
Bar: bit-exact ids AND bit-exact fp32 distances (the kernels reproduce the reference's sequential
fp32 arithmetic), which is stricter than north_star's "1e-5 relative, ties within epsilon"."""
import numpy as np
import pytest

import oracle
from scintirete_b200 import (DeviceStore, DistanceMetric, GPUFlatIndex, ScintireteError, SearchParams, Vector,
                             batch_distance, new_distance_calculator)
from util import gaussian

pytestmark = pytest.mark.gpu
METRICS = [DistanceMetric.L2, DistanceMetric.COSINE, DistanceMetric.INNER_PRODUCT]


def _flat(metric, db, q, k, **opts):
    s = DeviceStore(db.shape[1], metric)
    for name, v in opts.items():
        s.set_option(name, v)
    s.append(db)
    out = s.search_flat(q, k)
    s.close()
    return out


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("n,d,nq,k", [(5000, 128, 9, 10), (20011, 768, 3, 10), (777, 33, 17, 5), (4096, 64, 1, 1),
                                      (3000, 100, 20, 100)])
def test_exact_scan_bit_exact(metric, n, d, nq, k):
    db, q = gaussian(n, d, 1234), gaussian(nq, d, 4321)
    ids, dist, cnt = _flat(metric, db, q, k, flat_path=1)
    o_ids, o_dist, o_cnt = oracle.flat_search(int(metric), db, q, k, nthreads=8)
    assert np.array_equal(ids, o_ids)
    assert np.array_equal(dist, o_dist)
    assert np.array_equal(cnt, o_cnt)


@pytest.mark.parametrize("metric", METRICS)
def test_golden_distance_rows(metric):
    # the reference's own known-answer rows (distance_test.go) through scn_distance_batch
    rows = {
        DistanceMetric.L2: [([1, 2, 3], [1, 2, 3]), ([1, 0, 0], [0, 1, 0]), ([1, 1, 0], [4, 5, 0]), ([-1, -2, 0], [1, 2, 0])],
        DistanceMetric.COSINE: [([1, 2, 3], [1, 2, 3]), ([1, 0, 0], [0, 1, 0]), ([1, 0, 0], [-1, 0, 0]), ([1, 2, 0], [2, 4, 0]),
                                ([0, 0, 0], [0, 0, 0]), ([1, 2, 0], [0, 0, 0])],
        DistanceMetric.INNER_PRODUCT: [([1, 2, 3], [1, 1, 1]), ([1, 2, 0], [-1, -1, 0]), ([1, 0, 0], [0, 1, 0]), ([2, 3, 0], [2, 3, 0])],
    }[metric]
    calc = new_distance_calculator(metric)
    for a, b in rows:
        assert calc.distance(a, b) == oracle.distance(int(metric), a, b)
    got = batch_distance(new_distance_calculator(1), [0, 0], [[1, 0], [0, 1], [1, 1], [2, 2]])
    assert np.array_equal(got, oracle.batch_distance(1, [0, 0], [[1, 0], [0, 1], [1, 1], [2, 2]]))


@pytest.mark.parametrize("metric", METRICS)
def test_distance_batch_random_bit_exact(metric):
    a, b = gaussian(7, 257, 1), gaussian(11, 257, 2)
    got = new_distance_calculator(metric).pairwise(a, b)
    want = np.stack([oracle.batch_distance(int(metric), a[i], b) for i in range(len(a))])
    assert np.array_equal(got, want)


def test_ties_resolve_to_lower_row_and_padding():
    db = np.array([[1, 0], [0, 1], [1, 1], [2, 2], [0, 1]], np.float32)
    ids, dist, cnt = _flat(DistanceMetric.L2, db, np.zeros((1, 2), np.float32), 8)
    o = oracle.flat_search(1, db, np.zeros((1, 2), np.float32), 8)
    assert list(ids[0]) == [1, 2, 5, 3, 4, 0, 0, 0] and cnt[0] == 5
    assert np.array_equal(ids, o[0]) and np.array_equal(dist, o[1])


def test_all_identical_vectors_like_reference_benchmark():
    # hnsw_test.go:473-479 uses one deterministic vector for every row: everything ties
    v = (np.arange(128, dtype=np.float32) / 128)
    db = np.tile(v, (1000, 1))
    ids, dist, _ = _flat(DistanceMetric.L2, db, v[None, :], 10)
    assert list(ids[0]) == list(range(1, 11)) and np.all(dist == 0)


@pytest.mark.parametrize("metric", METRICS)
def test_deleted_rows_and_custom_ids(metric):
    n, d = 3000, 48
    db, q = gaussian(n, d, 5), gaussian(6, d, 6)
    ids_ext = (np.arange(n, dtype=np.uint64) * 7 + 100)
    s = DeviceStore(d, metric)
    s.append(db[:1000], ids_ext[:1000])
    s.append(db[1000:], ids_ext[1000:])          # second append exercises store growth
    dele = np.zeros(n, np.uint8)
    dele[::3] = 1
    s.mark_deleted(ids_ext[dele.astype(bool)])
    s.mark_deleted(ids_ext[:3])                   # already deleted / repeated: no-op
    ids, dist, cnt = s.search_flat(q, 10)
    dele[:3] = 1
    o = oracle.flat_search(int(metric), db, q, 10, ids=ids_ext, deleted=dele)
    assert np.array_equal(ids, o[0]) and np.array_equal(dist, o[1])
    assert s.stats().live_rows == n - int(dele.sum())
    with pytest.raises(ScintireteError) as e:
        s.mark_deleted([5])                       # unknown id
    assert e.value.code == 3004
    with pytest.raises(ScintireteError) as e:
        s.append(db[:1], [ids_ext[0]])            # duplicate id (hnsw.go:192-194)
    assert e.value.code == 3007
    s.close()


def test_zero_norm_cosine_exactly_one():
    db = gaussian(100, 16, 9)
    db[10] = 0
    q = np.zeros((1, 16), np.float32)
    ids, dist, _ = _flat(DistanceMetric.COSINE, db, q, 5)
    assert np.all(dist == np.float32(1.0)) and list(ids[0]) == [1, 2, 3, 4, 5]
    ids, dist, _ = _flat(DistanceMetric.COSINE, db, gaussian(1, 16, 10), 100)
    o = oracle.flat_search(2, db, gaussian(1, 16, 10), 100)
    assert np.array_equal(ids, o[0]) and np.array_equal(dist, o[1])


def test_empty_store_and_errors():
    s = DeviceStore(8, DistanceMetric.L2)
    ids, dist, cnt = s.search_flat(np.zeros((2, 8), np.float32), 3)
    assert np.all(ids == 0) and np.all(np.isinf(dist)) and np.all(cnt == 0)
    with pytest.raises(ScintireteError) as e:
        s.search_flat(np.zeros((2, 9), np.float32), 3)
    assert e.value.code == 3005
    with pytest.raises(ScintireteError) as e:
        s.search_flat(np.zeros((2, 8), np.float32), 0)
    assert e.value.code == 3007
    s.close()
    with pytest.raises(ScintireteError):
        DeviceStore(8, 0)


@pytest.mark.parametrize("metric", METRICS)
def test_rerank_matches_oracle_on_given_candidates(metric):
    n, d, nq, nc, k = 4000, 96, 12, 50, 10
    db, q = gaussian(n, d, 11), gaussian(nq, d, 12)
    rng = np.random.default_rng(3)
    cand = rng.integers(1, n + 1, size=(nq, nc)).astype(np.uint64)
    cand[:, 0] = 0                 # empty slot
    cand[:, 1] = cand[:, 2]        # duplicate candidate
    s = DeviceStore(d, metric)
    s.append(db)
    ids, dist, cnt = s.rerank(q, cand, k)
    s.close()
    for i in range(nq):
        rows = np.unique(cand[i][cand[i] > 0]) - 1
        dd = oracle.batch_distance(int(metric), q[i], db[rows])
        order = np.lexsort((rows, dd))[:k]
        assert np.array_equal(ids[i], rows[order] + 1)
        assert np.array_equal(dist[i], dd[order])


def test_vector_index_interface_like_reference_tests():
    # hnsw_test.go:124-160 / 162-219 shape, on the flat index
    idx = GPUFlatIndex(2, DistanceMetric.L2)
    idx.build([Vector(1, [1.0, 0.0]), Vector(2, [0.0, 1.0]), Vector(3, [1.0, 1.0]), Vector(4, [2.0, 2.0])])
    assert idx.size() == 4
    res = idx.search([0.0, 0.0], SearchParams(top_k=2))
    assert len(res) == 2 and res[0].distance <= res[1].distance
    idx.delete("2")
    assert idx.size() == 3
    with pytest.raises(ScintireteError):
        idx.get("2")
    assert all(r.vector.id != 2 for r in idx.search([0.0, 1.0], SearchParams(top_k=3)))
    with pytest.raises(ScintireteError):
        idx.delete("nonexistent")
    idx.delete("2")  # already deleted: succeeds
    with pytest.raises(ScintireteError):
        idx.insert(Vector(1, [1.0, 2.0]))  # duplicate insert
    assert np.array_equal(idx.get("4").elements, np.array([2.0, 2.0], np.float32))


def test_get_of_many_ids_gathers_on_the_device():
    # result decoration (include_vector) for whole batches: rows come back through one device gather
    # per block of ids; ragged dim (row pitch != dim), repeated and out-of-order ids, unknown id -> 3004
    n, d = 5000, 33
    db = gaussian(n, d, 11)
    s = DeviceStore(d, DistanceMetric.L2)
    s.append(db)
    ids = np.random.default_rng(5).integers(1, n + 1, 3000).astype(np.uint64)
    assert np.array_equal(s.get(ids), db[ids.astype(np.int64) - 1])
    with pytest.raises(ScintireteError) as e:
        s.get(np.array([1, 2, 3, 4, 5, n + 7], np.uint64))
    assert e.value.code == 3004
    s.close()


# ---- distance.go:152-192 helpers -----------------------------------------------

@pytest.mark.parametrize("v,want", [([1, 0], [1, 0]), ([3, 4], [0.6, 0.8]), ([0, 0], [0, 0]),
                                    ([-1, 1], [-1 / np.sqrt(np.float32(2)), 1 / np.sqrt(np.float32(2))])])
def test_normalize_vector_golden_rows(v, want):          # distance_test.go:312-366
    from scintirete_b200 import normalize_vector, vector_magnitude
    got = normalize_vector(v)
    assert np.all(np.abs(got - np.array(want, np.float32)) <= 1e-6)
    assert np.array_equal(got, oracle.normalize(v))
    if any(v):
        assert abs(float(vector_magnitude(got)) - 1.0) <= 1e-6


@pytest.mark.parametrize("v,want", [([1, 0], 1.0), ([3, 4], 5.0), ([0, 0, 0], 0.0), ([-1, -1], float(np.sqrt(2.0)))])
def test_vector_magnitude_golden_rows(v, want):          # distance_test.go:368-409
    from scintirete_b200 import vector_magnitude
    assert abs(float(vector_magnitude(v)) - want) <= 1e-6
    assert vector_magnitude(v) == oracle.magnitude(v)


@pytest.mark.parametrize("a,b,want", [([1, 0], [0, 1], 0.0), ([1, 2, 3], [1, 2, 3], 14.0), ([1, -2, 3], [4, 5, 6], 12.0),
                                      ([1, 2], [1, 2, 3], 0.0)])
def test_dot_product_golden_rows(a, b, want):            # distance_test.go:411-451 (length mismatch -> 0)
    from scintirete_b200 import dot_product
    assert float(dot_product(a, b)) == want


def test_vector_helpers_batched_bit_exact():
    # batches of random vectors, ragged dim: every result carries the bits of the scalar Go loop;
    # the store's precomputed ||x|| (cosine's normB) is the same VectorMagnitude
    from scintirete_b200 import dot_product, normalize_vector, vector_magnitude
    n, d = 700, 257
    a, b = gaussian(n, d, 31), gaussian(n, d, 32)
    a[5] = 0                                              # zero vector: returned unchanged, magnitude 0
    mag, nrm, dot = vector_magnitude(a), normalize_vector(a), dot_product(a, b)
    for i in range(n):
        assert mag[i] == oracle.magnitude(a[i])
        assert dot[i] == oracle.dot(a[i], b[i])
        assert np.array_equal(nrm[i], oracle.normalize(a[i]))
    assert np.array_equal(nrm[5], a[5]) and mag[5] == 0


def test_pinned_and_pageable_host_buffers_give_the_same_answer():
    # scn_host_alloc buffers are DMA-ed directly; a pageable buffer (what a Go []float32 is) goes
    # through the library's pinned staging chunks (several chunks at this size). Same bits either way.
    import ctypes as C

    from scintirete_b200 import PinnedBuffer, _native
    from scintirete_b200.index import _check

    n, d, nq, k = 20000, 768, 4000, 10          # 12.3 MB of queries: three staging chunks
    db, q = gaussian(n, d, 21), gaussian(nq, d, 22)
    s = DeviceStore(d, DistanceMetric.COSINE)
    s.append(db)
    ids_p, dist_p, cnt_p = s.search_flat(q, k)   # pageable numpy buffers
    pq, pi, pd, pc = PinnedBuffer((nq, d), np.float32), PinnedBuffer((nq, k), np.uint64), PinnedBuffer((nq, k), np.float32), \
        PinnedBuffer((nq,), np.uint32)
    pq.array[:] = q
    _check(_native.lib().scn_search_flat(s.handle, C.c_void_p(pq.ptr), nq, k, C.c_void_p(pi.ptr), C.c_void_p(pd.ptr),
                                         C.c_void_p(pc.ptr)))
    assert np.array_equal(pi.array, ids_p) and np.array_equal(pd.array, dist_p) and np.array_equal(pc.array, cnt_p)
    o = oracle.flat_search(2, db, q[:64], k, nthreads=8)
    assert np.array_equal(ids_p[:64], o[0]) and np.array_equal(dist_p[:64], o[1])
    for b in (pq, pi, pd, pc):
        b.close()
    s.close()

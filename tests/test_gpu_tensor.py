"""GPU parity of the tensor-core flat path (tcgen05 filter + certified exact rerank): the result
must be IDENTICAL (ids and fp32 distance bits) to the CPU oracle, because every returned distance
is recomputed in the reference's arithmetic and the certificate (or the exact re-scan behind it)
guarantees that no row was missed."""
import ctypes as C

import numpy as np
import pytest

import oracle
from scintirete_b200 import DeviceStore, DistanceMetric, _native
from util import gaussian

pytestmark = pytest.mark.gpu
METRICS = [DistanceMetric.L2, DistanceMetric.COSINE, DistanceMetric.INNER_PRODUCT]


def bf16_round(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def debug_scores(store, q):
    q = np.ascontiguousarray(q, np.float32)
    n = int(store.stats().rows)
    out = np.empty((q.shape[0], n), np.float32)
    rc = _native.lib().scn_debug_tensor_scores(store.handle, q.ctypes.data_as(C.c_void_p), q.shape[0],
                                               out.ctypes.data_as(C.c_void_p))
    assert rc == 0, _native.last_error()
    return out


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("n,d,nq", [(3000, 768, 130), (2500, 128, 70), (1000, 200, 5), (1500, 1536, 140), (700, 800, 9)])
def test_filter_scores_match_bf16_emulation(metric, n, d, nq):
    db, q = gaussian(n, d, 1), gaussian(nq, d, 2)
    s = DeviceStore(d, metric)
    s.append(db)
    got = debug_scores(s, q)
    s.close()
    x = db.astype(np.float32)
    if metric == DistanceMetric.COSINE:
        x = x * (np.float32(1) / np.sqrt((x.astype(np.float64) ** 2).sum(1)).astype(np.float32))[:, None]
    xb, qb = bf16_round(x).astype(np.float64), bf16_round(q).astype(np.float64)
    dot = qb @ xb.T
    want = (xb ** 2).sum(1)[None, :] - 2 * dot if metric == DistanceMetric.L2 else -dot
    err = np.abs(got - want).max()
    scale = np.abs(want).max()
    assert err <= 2e-5 * scale + 2e-3, (err, scale)


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("n,d,nq,k", [(20000, 768, 300, 10), (6000, 128, 130, 10), (9999, 200, 77, 1), (5000, 64, 129, 24),
                                      (9000, 1536, 260, 10), (5000, 1000, 33, 24)])
def test_tensor_path_bit_exact_vs_oracle(metric, n, d, nq, k):
    db, q = gaussian(n, d, 1234), gaussian(nq, d, 4321)
    s = DeviceStore(d, metric)
    s.set_option("flat_path", 2)
    s.append(db)
    ids, dist, cnt = s.search_flat(q, k)
    tensor_q, widened, rescanned = s.last_counters()[:3]
    s.close()
    o_ids, o_dist, o_cnt = oracle.flat_search(int(metric), db, q, k, nthreads=8)
    assert np.array_equal(ids, o_ids)
    assert np.array_equal(dist, o_dist)
    assert np.array_equal(cnt, o_cnt)
    assert tensor_q == nq
    assert widened <= 0.02 * nq + 1, f"{widened} of {nq} queries failed the first certificate on Gaussian data"
    assert rescanned <= widened


@pytest.mark.parametrize("metric", METRICS)
def test_tensor_path_with_deleted_rows_and_ids(metric):
    n, d, nq = 8000, 96, 140
    db, q = gaussian(n, d, 5), gaussian(nq, d, 6)
    ids_ext = np.arange(n, dtype=np.uint64) * 3 + 11
    dele = np.zeros(n, np.uint8)
    dele[::4] = 1
    s = DeviceStore(d, metric)
    s.set_option("flat_path", 2)
    s.append(db, ids_ext)
    s.mark_deleted(ids_ext[dele.astype(bool)])
    ids, dist, cnt = s.search_flat(q, 10)
    s.close()
    o = oracle.flat_search(int(metric), db, q, 10, ids=ids_ext, deleted=dele, nthreads=8)
    assert np.array_equal(ids, o[0]) and np.array_equal(dist, o[1])


def test_ties_force_exact_rescan_and_stay_exact():
    # every row identical (the reference's own benchmark data, hnsw_test.go:473-479): the filter
    # cannot separate anything, the certificate must refuse, the exact scan must answer
    v = (np.arange(128, dtype=np.float32) / 128)
    db = np.tile(v, (5000, 1))
    q = np.tile(v, (130, 1)) + gaussian(130, 128, 3) * np.float32(1e-3)
    for metric in METRICS:
        s = DeviceStore(128, metric)
        s.set_option("flat_path", 2)
        s.append(db)
        ids, dist, cnt = s.search_flat(q, 10)
        c = s.last_counters()
        s.close()
        o = oracle.flat_search(int(metric), db, q, 10, nthreads=8)
        assert np.array_equal(ids, o[0]) and np.array_equal(dist, o[1])
        assert c[1] == 130 and c[2] == 130   # no certificate can pass: every query is re-scanned exactly


def test_near_duplicates_cluster():
    # tight clusters: many rows within bf16 noise of each other -> some queries need the re-scan,
    # all must still be exact
    rng = np.random.default_rng(9)
    centers = rng.standard_normal((50, 256)).astype(np.float32)
    db = (centers[rng.integers(0, 50, 6000)] + rng.standard_normal((6000, 256)).astype(np.float32) * np.float32(1e-3))
    q = centers[rng.integers(0, 50, 150)] + rng.standard_normal((150, 256)).astype(np.float32) * np.float32(1e-3)
    for metric in METRICS:
        s = DeviceStore(256, metric)
        s.set_option("flat_path", 2)
        s.append(db)
        ids, dist, _ = s.search_flat(q, 10)
        s.close()
        o = oracle.flat_search(int(metric), db, q, 10, nthreads=8)
        assert np.array_equal(ids, o[0]) and np.array_equal(dist, o[1])


def test_auto_dispatch():
    # large stores: the tensor filter serves every batch size (it streams the bf16 mirror, half the
    # bytes of the fp32 scan), dims beyond the TMEM budget through the smem-streamed variant; small
    # stores use the exact scan
    s = DeviceStore(128, DistanceMetric.L2)
    s.append(gaussian(8192, 128, 1))
    s.search_flat(gaussian(1, 128, 2), 10)
    assert s.last_counters()[0] == 1
    s.search_flat(gaussian(64, 128, 2), 10)
    assert s.last_counters()[0] == 64
    s.set_option("tensor_min_batch", 16)
    s.search_flat(gaussian(4, 128, 2), 10)
    assert s.last_counters()[0] == 0
    s.close()
    s = DeviceStore(128, DistanceMetric.L2)
    s.append(gaussian(1000, 128, 1))
    s.search_flat(gaussian(64, 128, 2), 10)
    assert s.last_counters()[0] == 0
    s.close()
    s = DeviceStore(1536, DistanceMetric.INNER_PRODUCT)
    db, q = gaussian(5000, 1536, 1), gaussian(20, 1536, 2)
    s.append(db)
    ids, dist, _ = s.search_flat(q, 10)
    assert s.last_counters()[0] == 20
    o = oracle.flat_search(3, db, q, 10, nthreads=8)
    assert np.array_equal(ids, o[0]) and np.array_equal(dist, o[1])
    s.close()


@pytest.mark.parametrize("metric", METRICS)
def test_small_batch_on_many_rows_selects_candidates_by_radix(metric):
    # few queries over many rows: the query block is spread over ~148 row chunks, so the merge sees
    # thousands of candidates per query and takes the radix-select path (k'' + 1 of them matter)
    n, d, nq, k = 200_000, 64, 3, 10
    db, q = gaussian(n, d, 21), gaussian(nq, d, 22)
    s = DeviceStore(d, metric)
    s.set_option("flat_path", 2)
    s.append(db)
    ids, dist, _ = s.search_flat(q, k)
    c = s.last_counters()
    s.close()
    o = oracle.flat_search(int(metric), db, q, k, nthreads=8)
    assert np.array_equal(ids, o[0]) and np.array_equal(dist, o[1])
    assert c[2] == 0                      # certified without the exact re-scan


def test_small_batch_with_mass_ties_falls_back_to_the_full_sort_and_stays_exact():
    # 150 000 identical rows among 200 000: far more keys tie at the selection boundary than the
    # select buffer holds -> full sort, certificate refuses, exact re-scan answers
    n, d, nq, k = 200_000, 64, 2, 10
    db = gaussian(n, d, 23)
    db[10_000:160_000] = db[9_999]
    q = np.stack([db[9_999] + np.float32(1e-3), gaussian(1, d, 24)[0]])
    s = DeviceStore(d, DistanceMetric.L2)
    s.set_option("flat_path", 2)
    s.append(db)
    ids, dist, _ = s.search_flat(q, k)
    s.close()
    o = oracle.flat_search(1, db, q, k, nthreads=8)
    assert np.array_equal(ids, o[0]) and np.array_equal(dist, o[1])


@pytest.mark.parametrize("metric", [DistanceMetric.L2, DistanceMetric.COSINE, DistanceMetric.INNER_PRODUCT])
@pytest.mark.parametrize("n,d,nq,k", [(30001, 768, 300, 10), (20000, 600, 256, 24), (50000, 512, 1000, 10), (9000, 470, 513, 5)])
def test_cta_pair_filter_equals_oracle(metric, n, d, nq, k):
    # the tcgen05 cta_group::2 kernel (kpad 512 / 640 / 768, batches >= 256): an odd number of query blocks
    # (the last pair is half empty), rows that do not fill the last 128-row tile, deleted rows, both k' sizes;
    # ids and distance bits equal the oracle's and the single-CTA kernel's
    db, q = gaussian(n, d, 31), gaussian(nq, d, 32)
    db[n - 1] = db[3]
    q[5] = db[3]
    s = DeviceStore(d, metric)
    s.append(db)
    dead = np.array([4, n // 3, n - 5], np.uint64)
    s.mark_deleted(dead)
    deleted = np.zeros(n, np.uint8)
    deleted[dead.astype(np.int64) - 1] = 1
    s.set_option("flat_path", 2)
    s.set_option("profile", 1)
    s.set_option("tensor_pair", 1)
    ids, dist, cnt = s.search_flat(q, k)
    assert "tensor_filter" in s.last_timings()
    s.set_option("tensor_pair", 0)
    ids1, dist1, cnt1 = s.search_flat(q, k)
    assert np.array_equal(ids, ids1) and np.array_equal(dist, dist1) and np.array_equal(cnt, cnt1)
    o_ids, o_dist, o_cnt = oracle.flat_search(int(metric), db, q, k, deleted=deleted, nthreads=8)
    assert np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist) and np.array_equal(cnt, o_cnt)
    s.close()


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("n,d,nq,k", [(20000, 768, 5, 10), (6000, 128, 130, 10), (9999, 200, 1, 1), (5000, 64, 129, 24),
                                      (9000, 1536, 40, 10), (5000, 1000, 33, 10), (4200, 33, 17, 10), (150000, 96, 3, 10)])
def test_fused_tail_equals_three_kernel_tail_and_oracle(metric, n, d, nq, k):
    # finish_queries_kernel (candidate merge + exact rerank + certificate in one launch, option tensor_fused) against
    # merge_candidates -> rerank -> certify and against the oracle: ids, distance bits, counts, and the same
    # certificate decisions (counters). Covers one query, ragged dims, k'' = 64 (rows > 768 elements), k' = 32,
    # deleted rows, a duplicated row, and long candidate lists (few queries over many row chunks: radix select).
    db, q = gaussian(n, d, 77), gaussian(nq, d, 78)
    db[n - 1] = db[7]
    q[0] = db[7]
    s = DeviceStore(d, metric)
    s.append(db)
    dead = np.array([3, n // 2, n - 2], np.uint64)
    s.mark_deleted(dead)
    deleted = np.zeros(n, np.uint8)
    deleted[dead.astype(np.int64) - 1] = 1
    s.set_option("flat_path", 2)
    s.set_option("profile", 1)
    got = {}
    for fused in (0, 1):
        s.set_option("tensor_fused", fused)
        s.last_timings()
        ids, dist, cnt = s.search_flat(q, k)
        names = s.last_timings()
        assert ("finish_queries" in names) == bool(fused) and ("merge_candidates" in names) != bool(fused), names
        got[fused] = (ids, dist, cnt, s.last_counters()[:3])
    s.close()
    o_ids, o_dist, o_cnt = oracle.flat_search(int(metric), db, q, k, deleted=deleted, nthreads=8)
    for fused in (0, 1):
        ids, dist, cnt, _ = got[fused]
        assert np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist) and np.array_equal(cnt, o_cnt), fused
    assert got[0][3][0] == got[1][3][0] == nq


def test_fused_tail_with_ties_everywhere():
    # all rows identical: every filter score ties, no certificate can hold, every query ends in the exact scan —
    # through the fused tail exactly as through the three kernels
    n, d, nq, k = 5000, 128, 9, 10
    db = np.tile(gaussian(1, d, 5), (n, 1))
    q = gaussian(nq, d, 6)
    for fused in (0, 1):
        s = DeviceStore(d, DistanceMetric.L2)
        s.append(db)
        s.set_option("flat_path", 2)
        s.set_option("tensor_fused", fused)
        ids, dist, cnt = s.search_flat(q, k)
        rescanned = s.last_counters()[2]
        s.close()
        o = oracle.flat_search(1, db, q, k, nthreads=8)
        assert np.array_equal(ids, o[0]) and np.array_equal(dist, o[1]) and np.array_equal(cnt, o[2])
        assert rescanned == nq


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("n,d,nq,k", [(150000, 96, 3, 10), (60000, 512, 256, 10), (100000, 768, 300, 10), (40000, 200, 130, 24)])
def test_shared_bounds_between_lists_keep_results_and_certificates(metric, n, d, nq, k):
    # few queries over many row chunks: >= 16 candidate lists per query exchange bounds while they are built
    # (shared_bound, option tensor_share) — both filter kernels. Results stay the oracle's bit for bit, and the
    # tighter thresholds do not cost certificates (a bound never drops below the query's 32nd best score).
    db, q = gaussian(n, d, 91), gaussian(nq, d, 92)
    s = DeviceStore(d, metric)
    s.append(db)
    s.set_option("flat_path", 2)
    o_ids, o_dist, o_cnt = oracle.flat_search(int(metric), db, q, k, nthreads=8)
    for share in (2, 0):   # 2 = whenever a query has >= 16 lists
        s.set_option("tensor_share", share)
        ids, dist, cnt = s.search_flat(q, k)
        tensor_q, widened, rescanned = s.last_counters()[:3]
        assert np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist) and np.array_equal(cnt, o_cnt), share
        assert tensor_q == nq and widened <= 0.02 * nq + 1 and rescanned <= widened, (share, widened, rescanned)
    s.close()


@pytest.mark.parametrize("metric", [DistanceMetric.L2, DistanceMetric.COSINE])
@pytest.mark.parametrize("n,d,nq,k", [(30001, 768, 300, 10), (20000, 600, 256, 24), (50000, 512, 700, 10), (9000, 470, 513, 24)])
def test_cta_pair_filter_one_and_two_epilogue_warps_per_quarter(metric, n, d, nq, k):
    # tensor_pair_ew: every TMEM lane quarter of the pair kernel is gated by one warp (128 columns of a tile, one list per
    # chunk) or by two (64 columns each, two lists per chunk); one and two accumulators, both k' sizes
    db, q = gaussian(n, d, 41), gaussian(nq, d, 42)
    s = DeviceStore(d, metric)
    s.append(db)
    s.set_option("flat_path", 2)
    o_ids, o_dist, o_cnt = oracle.flat_search(int(metric), db, q, k, nthreads=8)
    for ew in (1, 2):
        s.set_option("tensor_pair_ew", ew)
        ids, dist, cnt = s.search_flat(q, k)
        assert np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist) and np.array_equal(cnt, o_cnt), ew
        assert s.last_counters()[0] == nq
    s.close()

"""N > 1 path on CPU: two gloo ranks run the row-sharded exact-search protocol end to end
(shard plan -> shard-local top-k as 64-bit merge keys -> all_gather -> k-way merge) and must
reproduce the single-shard oracle bit for bit. Shard-local search and the merge are stood in for
by the oracle / a numpy model of scn_merge_topk_dev's contract; what is under test is the
protocol (row bases, key layout, tie order across shards, padding), i.e. the host logic that
bench.py --gpus N and the Go shim drive."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def f32_ord(d):
    """numpy twin of scn::f32_ord (csrc/common.cuh)."""
    d = np.asarray(d, np.float32) + np.float32(0)
    u = d.view(np.uint32)
    o = np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)
    return np.where(np.isnan(d), np.uint32(0xFFFFFFFF), o)


def make_keys(dist_, rows, valid):
    k = (f32_ord(dist_).astype(np.uint64) << np.uint64(32)) | rows.astype(np.uint64)
    return np.where(valid, k, np.uint64(0xFFFFFFFFFFFFFFFF))


def merge_model(keys, ids, k):
    """Contract of scn_merge_topk_dev: ascending key order over [G][nq][k], first k, pad with 0/+inf."""
    g, nq, _ = keys.shape
    out_ids = np.zeros((nq, k), np.uint64)
    out_d = np.full((nq, k), np.inf, np.float32)
    for q in range(nq):
        kk = keys[:, q, :].ravel()
        ii = ids[:, q, :].ravel()
        order = np.argsort(kk, kind="stable")[:k]
        for j, o in enumerate(order):
            if kk[o] == np.uint64(0xFFFFFFFFFFFFFFFF):
                break
            out_ids[q, j] = ii[o]
            o32 = np.uint32(kk[o] >> np.uint64(32))
            bits = (o32 & np.uint32(0x7FFFFFFF)) if (o32 & np.uint32(0x80000000)) else ~o32
            out_d[q, j] = np.array([bits], np.uint32).view(np.float32)[0]
    return out_ids, out_d


def _worker(rank, world, port, n, d, nq, k, metric, ret):
    sys.path.insert(0, ROOT)
    import oracle
    from scintirete_b200.sharding import gather_shape, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1234)
    db = rng.standard_normal((n, d)).astype(np.float32)
    db[n // 2 + 3] = db[5]          # an exact tie across the shard boundary
    db[n // 2 + 4] = db[5]
    q = np.random.default_rng(4321).standard_normal((nq, d)).astype(np.float32)
    q[0] = db[5]
    lo, hi = shard_range(n, world, rank)
    ids_all = np.arange(n, dtype=np.uint64) * 5 + 7
    l_ids, l_dist, l_cnt = oracle.flat_search(metric, db[lo:hi], q, k, ids=ids_all[lo:hi])
    # local rows from ids, then global rows = row_base + local row
    l_rows = ((l_ids.astype(np.int64) - 7) // 5)
    valid = np.arange(k)[None, :] < l_cnt[:, None]
    keys = make_keys(l_dist, np.where(valid, l_rows, 0), valid)
    tk = torch.from_numpy(keys.view(np.int64))
    ti = torch.from_numpy(l_ids.view(np.int64))
    shape = gather_shape(world, nq, k)
    # gloo wants the concatenated form [G*nq][k]; it is the same memory as [G][nq][k]
    all_k = torch.zeros((shape[0] * shape[1], shape[2]), dtype=torch.int64)
    all_i = torch.zeros((shape[0] * shape[1], shape[2]), dtype=torch.int64)
    dist.all_gather_into_tensor(all_k, tk)
    dist.all_gather_into_tensor(all_i, ti)
    all_k, all_i = all_k.view(shape), all_i.view(shape)
    m_ids, m_dist = merge_model(all_k.numpy().view(np.uint64), all_i.numpy().view(np.uint64), k)
    o_ids, o_dist, _ = oracle.flat_search(metric, db, q, k, ids=ids_all)
    ok = bool(np.array_equal(m_ids, o_ids) and np.array_equal(m_dist, o_dist))
    ret[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("metric", [1, 2, 3])
def test_two_rank_sharded_search_matches_single_shard(metric):
    world = 2
    port = 29500 + os.getpid() % 2000 + metric
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, 1001, 24, 9, 10, metric, ret), nprocs=world, join=True)
        assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shard_plan_covers_rows_exactly_once():
    sys.path.insert(0, ROOT)
    from scintirete_b200.sharding import shard_range

    for n in (0, 1, 7, 1000, 1001, 10_000_000):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)

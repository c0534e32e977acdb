"""Row-sharded exact search on one GPU: G stores stand in for G ranks; shard-local keys are
concatenated as the all-gather would leave them and merged by scn_merge_topk_dev. The result
must be bit-identical to the single-shard oracle (and to a single store)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import oracle
from scintirete_b200 import DeviceStore, DistanceMetric, _native
from scintirete_b200.index import _check
from scintirete_b200.sharding import gather_shape, shard_range
from util import gaussian

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("metric", [DistanceMetric.L2, DistanceMetric.COSINE, DistanceMetric.INNER_PRODUCT])
@pytest.mark.parametrize("world,n,d,nq,k,path", [(2, 6001, 64, 40, 10, 1), (4, 20000, 128, 150, 10, 2), (8, 9000, 768, 130, 5, 0),
                                                 (3, 10, 16, 5, 10, 1)])
def test_sharded_equals_single(metric, world, n, d, nq, k, path):
    lib = _native.lib()
    db, q = gaussian(n, d, 1234), gaussian(nq, d, 4321)
    db[n // 2] = db[1]  # a tie across shards must resolve to the lower global row
    q[0] = db[1]
    ids_ext = np.arange(n, dtype=np.uint64) * 3 + 5
    dev = torch.device("cuda", 0)
    qd = torch.from_numpy(q).to(dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    shape = gather_shape(world, nq, k)
    all_keys = torch.zeros(shape, dtype=torch.int64, device=dev)
    all_ids = torch.zeros(shape, dtype=torch.int64, device=dev)
    stores = []
    for r in range(world):
        lo, hi = shard_range(n, world, r)
        s = DeviceStore(d, metric)
        s.set_option("flat_path", path)
        if hi > lo:
            s.append(db[lo:hi], ids_ext[lo:hi])
        stores.append(s)
        _check(lib.scn_search_flat_shard_dev(s.handle, _p(qd), nq, k, lo, _p(all_keys[r]), _p(all_ids[r]), stream))
    out_ids = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    out_dist = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    out_cnt = torch.zeros((nq,), dtype=torch.int32, device=dev)
    _check(lib.scn_merge_topk_dev(0, _p(all_keys), _p(all_ids), world, nq, k, _p(out_ids), _p(out_dist), _p(out_cnt), stream))
    torch.cuda.synchronize()
    o_ids, o_dist, o_cnt = oracle.flat_search(int(metric), db, q, k, ids=ids_ext, nthreads=8)
    assert np.array_equal(out_ids.cpu().numpy().view(np.uint64), o_ids)
    assert np.array_equal(out_dist.cpu().numpy(), o_dist)
    assert np.array_equal(out_cnt.cpu().numpy().view(np.uint32), o_cnt)
    for s in stores:
        s.close()


@pytest.mark.parametrize("metric", [DistanceMetric.L2, DistanceMetric.COSINE])
@pytest.mark.parametrize("q_is_slice", [False, True])
@pytest.mark.parametrize("world,n,d,nq,k,path", [(2, 6001, 64, 40, 10, 1), (4, 30000, 128, 300, 10, 2), (8, 9000, 96, 130, 10, 0),
                                                 (3, 5000, 33, 2, 5, 0)])
def test_fused_peer_exchange_equals_single(metric, q_is_slice, world, n, d, nq, k, path):
    # both exchanges fused over peer memory (query gather + top-k lists to the owner of each query slice,
    # P2P stores + flags instead of all-gathers): G ranks of ONE process, one GPU each, as the reference's
    # single server process would drive its GPUs. Three rounds exercise the parity double buffering. Rank r
    # ends up with the results of query slice r; together they are the oracle's answer.
    # The ranks' kernels wait on one another, so every rank needs its own GPU (two launches on one GPU are
    # not guaranteed to run at the same time); on a smaller box the same kernels and protocol are covered
    # by the shard-set tests below, which synchronise between the steps when shards share a device.
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (one per rank)")
    from scintirete_b200.sharding import ShardExchange, query_slice

    db = gaussian(n, d, 1234)
    db[n // 2] = db[1]
    ids_ext = np.arange(n, dtype=np.uint64) * 3 + 5
    stores, exs, streams, outs = [], [], [], []
    for r in range(world):
        dev = torch.device("cuda", r)
        lo, hi = shard_range(n, world, r)
        s = DeviceStore(d, metric, device=r)
        s.set_option("flat_path", path)
        s.append(db[lo:hi], ids_ext[lo:hi])
        stores.append(s)
        exs.append(ShardExchange(r, r, world, 512, k, d))
        streams.append(torch.cuda.Stream(device=dev))
        outs.append((torch.zeros((nq, k), dtype=torch.int64, device=dev), torch.zeros((nq, k), dtype=torch.float32, device=dev),
                     torch.zeros((nq,), dtype=torch.int32, device=dev)))
    for e in exs:
        e.connect_local(exs)
    for rnd in range(3):
        q = gaussian(nq, d, 4321 + rnd)
        q[0] = db[1]
        srcs = []
        for r in range(world):
            qlo, qcnt = exs[r].slice(nq)
            assert (qlo, qlo + qcnt) == query_slice(nq, world, r)
            srcs.append(torch.from_numpy(np.ascontiguousarray(q[qlo:qlo + qcnt]) if q_is_slice else q).to(torch.device("cuda", r)))
        torch.cuda.synchronize()
        for r in range(world):
            lo, _ = shard_range(n, world, r)
            oi, od, oc = outs[r]
            exs[r].search(stores[r], srcs[r].data_ptr(), nq, lo, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), streams[r].cuda_stream,
                          q_is_slice=q_is_slice)
        for r in range(world):
            exs[r].status(streams[r].cuda_stream)
        o_ids, o_dist, o_cnt = oracle.flat_search(int(metric), db, q, k, ids=ids_ext, nthreads=8)
        for r in range(world):
            qlo, qcnt = exs[r].slice(nq)
            oi, od, oc = outs[r]
            assert np.array_equal(oi.cpu().numpy().view(np.uint64)[:qcnt], o_ids[qlo:qlo + qcnt])
            assert np.array_equal(od.cpu().numpy()[:qcnt], o_dist[qlo:qlo + qcnt])
            assert np.array_equal(oc.cpu().numpy().view(np.uint32)[:qcnt], o_cnt[qlo:qlo + qcnt])
    for e in exs:
        e.close()
    for s in stores:
        s.close()


def _devices(world):
    """`world` device ordinals: distinct GPUs when the box has them, else all shards on GPU 0 (the
    workers of a shard set own different streams, so the protocol is the same)."""
    n = torch.cuda.device_count()
    return [r % n for r in range(world)] if n >= world else [0] * world


@pytest.mark.parametrize("metric", [DistanceMetric.L2, DistanceMetric.COSINE, DistanceMetric.INNER_PRODUCT])
@pytest.mark.parametrize("world,n,d,nq,k,path", [(2, 9001, 64, 77, 10, 0), (4, 20000, 128, 301, 10, 2), (3, 700, 40, 5, 3, 0),
                                                 (8, 40000, 768, 130, 10, 0)])
def test_shard_set_host_call_equals_oracle(metric, world, n, d, nq, k, path):
    # scn_shards_*: one blocking host-buffer call drives all shards (one worker thread per device)
    from scintirete_b200.sharding import ShardedStore

    db, q = gaussian(n, d, 1234), gaussian(nq, d, 4321)
    db[n // 2] = db[1]
    q[0] = db[1]
    sh = ShardedStore(_devices(world), d, metric, n)
    sh.set_option("flat_path", path)
    for lo in range(0, n, 4097):           # appended in uneven pieces that straddle shard boundaries
        sh.append(db[lo:lo + 4097])
    assert int(sh.stats().rows) == n
    for rnd in range(2):
        ids, dist, cnt = sh.search_flat(q, k)
        o_ids, o_dist, o_cnt = oracle.flat_search(int(metric), db, q, k, nthreads=8)
        assert np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist) and np.array_equal(cnt, o_cnt)
    # soft delete through the shard set, a larger batch (the exchanges are rebuilt), one query
    dead = np.array([2, n // 2 + 1, n], np.uint64)
    sh.mark_deleted(dead)
    deleted = np.zeros(n, np.uint8)
    deleted[dead.astype(np.int64) - 1] = 1
    q2 = gaussian(1500, d, 99)
    for qq in (q2, q2[:1]):
        ids, dist, cnt = sh.search_flat(qq, k)
        o_ids, o_dist, o_cnt = oracle.flat_search(int(metric), db, qq, k, deleted=deleted, nthreads=8)
        assert np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist) and np.array_equal(cnt, o_cnt)
    sh.close()


def test_shard_set_rows_beyond_capacity_and_explicit_ids():
    from scintirete_b200.sharding import ShardedStore

    n, d, k = 3000, 32, 7
    db, q = gaussian(n, d, 5), gaussian(33, d, 6)
    ids_ext = np.arange(n, dtype=np.uint64) * 7 + 11
    sh = ShardedStore(_devices(2), d, DistanceMetric.L2, 2000)    # 1000 rows more than declared: last shard grows
    sh.append(db, ids_ext)
    ids, dist, cnt = sh.search_flat(q, k)
    o_ids, o_dist, o_cnt = oracle.flat_search(1, db, q, k, ids=ids_ext, nthreads=4)
    assert np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist) and np.array_equal(cnt, o_cnt)
    sh.close()


def _ipc_rank(rank, world, port, n, d, nq, k, metric, ret):
    # one process per GPU: the shape bench.py --gpus N runs (CUDA IPC handles + NVLink P2P stores)
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scintirete_b200.sharding import ShardExchange

    torch.cuda.set_device(rank)
    db, q = gaussian(n, d, 1234), gaussian(nq, d, 4321)
    db[n // 2] = db[1]
    q[0] = db[1]
    lo, hi = shard_range(n, world, rank)
    store = DeviceStore(d, DistanceMetric(metric), device=rank)
    store.append(db[lo:hi], np.arange(lo, hi, dtype=np.uint64) + 1)
    ex = ShardExchange(rank, rank, world, 4096, k, d)
    handles = [None] * world
    dist.all_gather_object(handles, ex.local_handle())
    ex.connect(handles)
    qlo, qcnt = ex.slice(nq)
    ids = np.zeros((max(qcnt, 1), k), np.uint64)
    dd = np.zeros((max(qcnt, 1), k), np.float32)
    cnt = np.zeros(max(qcnt, 1), np.uint32)
    ok = True
    for rnd in range(3):
        qs = np.ascontiguousarray(q[qlo:qlo + qcnt])
        ex.search_host(store, qs.ctypes.data, nq, lo, ids.ctypes.data, dd.ctypes.data, cnt.ctypes.data)
        o_ids, o_dist, o_cnt = oracle.flat_search(metric, db, q[qlo:qlo + qcnt], k, nthreads=4)
        ok = ok and np.array_equal(ids[:qcnt], o_ids) and np.array_equal(dd[:qcnt], o_dist) and np.array_equal(cnt[:qcnt], o_cnt)
        dist.barrier()
    ret[rank] = bool(ok)
    ex.close()
    store.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("metric", [1, 2])
def test_two_process_ipc_exchange_equals_oracle(metric):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    world = 2
    port = 29600 + os.getpid() % 2000 + metric
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_ipc_rank, args=(world, port, 50000, 128, 257, 10, metric, ret), nprocs=world, join=True)
        assert all(ret.get(r) for r in range(world)), dict(ret)

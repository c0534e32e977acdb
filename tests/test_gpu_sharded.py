"""Row-sharded exact search on one GPU: G stores stand in for G ranks; shard-local keys are
concatenated as the all-gather would leave them and merged by scn_merge_topk_dev. The result
must be bit-identical to the single-shard oracle (and to a single store)."""
import ctypes as C

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
@pytest.mark.parametrize("world,n,d,nq,k,path", [(2, 6001, 64, 40, 10, 1), (4, 30000, 128, 300, 10, 2), (8, 9000, 96, 130, 10, 0)])
def test_fused_peer_exchange_equals_single(metric, world, n, d, nq, k, path):
    # the shard exchange fused into the search epilogue (P2P stores + flags instead of all-gathers):
    # G ranks of ONE process, each on its own stream, as the reference's single server process would
    # drive its GPUs. Two rounds exercise the parity double buffering.
    from scintirete_b200.sharding import ShardExchange

    db = gaussian(n, d, 1234)
    db[n // 2] = db[1]
    ids_ext = np.arange(n, dtype=np.uint64) * 3 + 5
    dev = torch.device("cuda", 0)
    stores, exs, streams, outs = [], [], [], []
    for r in range(world):
        lo, hi = shard_range(n, world, r)
        s = DeviceStore(d, metric)
        s.set_option("flat_path", path)
        s.append(db[lo:hi], ids_ext[lo:hi])
        stores.append(s)
        exs.append(ShardExchange(0, r, world, 512, k))
        streams.append(torch.cuda.Stream(device=dev))
        outs.append((torch.zeros((nq, k), dtype=torch.int64, device=dev), torch.zeros((nq, k), dtype=torch.float32, device=dev),
                     torch.zeros((nq,), dtype=torch.int32, device=dev)))
    for e in exs:
        e.connect_local(exs)
    for rnd in range(3):
        q = gaussian(nq, d, 4321 + rnd)
        q[0] = db[1]
        qd = torch.from_numpy(q).to(dev)
        torch.cuda.synchronize()
        for r in range(world):
            lo, _ = shard_range(n, world, r)
            oi, od, oc = outs[r]
            exs[r].search(stores[r], qd.data_ptr(), nq, lo, oi.data_ptr(), od.data_ptr(), oc.data_ptr(), streams[r].cuda_stream)
        for r in range(world):
            exs[r].status(streams[r].cuda_stream)
        o_ids, o_dist, o_cnt = oracle.flat_search(int(metric), db, q, k, ids=ids_ext, nthreads=8)
        for r in range(world):   # every rank ends up with the full, identical answer
            oi, od, oc = outs[r]
            assert np.array_equal(oi.cpu().numpy().view(np.uint64), o_ids)
            assert np.array_equal(od.cpu().numpy(), o_dist)
            assert np.array_equal(oc.cpu().numpy().view(np.uint32), o_cnt)
    for e in exs:
        e.close()
    for s in stores:
        s.close()

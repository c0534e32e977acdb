"""Row-sharding plan for the exact scan across the GPUs of one box.

Host-side bookkeeping only: which contiguous block of rows a rank owns and how per-shard results
are laid out for the all-gather. The search and the merge themselves are CUDA
(`scn_search_flat_shard_dev`, `scn_merge_topk_dev`)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) held by `rank`: contiguous blocks of ceil(n/world) rows (the last may be short
    or empty). Contiguity is what makes `row_base + local row` a global row, so that ascending
    64-bit merge keys reproduce the single-GPU (distance, row) order exactly."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = (n_rows + world - 1) // world
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def gather_shape(world: int, nq: int, k: int) -> Tuple[int, int, int]:
    """Layout of the all-gathered key / id tensors expected by scn_merge_topk_dev: [G][nq][k]."""
    return (world, nq, k)


def query_slice(nq: int, world: int, rank: int) -> Tuple[int, int]:
    """Queries [lo, hi) of a batch of nq that `rank` answers in the fused exchange
    (scn_exchange_slice): contiguous slices of ceil(nq / world) queries."""
    return shard_range(nq, world, rank)


class ShardExchange:
    """One rank's end of the fused shard exchange (`scn_exchange_*`). Per call every rank contributes
    its slice of the query batch (gathered over NVLink peer memory), scans its rows for the whole
    batch, stores each top-k list into the buffer of the rank that owns the query, and merges its own
    slice from local memory — no NCCL collective, 1/world of the PCIe traffic per rank. Every rank must
    issue the call for the same batch."""

    HANDLE_BYTES = 64

    def __init__(self, device: int, rank: int, world: int, max_nq: int, k: int, dim: int = 0):
        import ctypes as C

        from . import _native
        from .index import _check

        self._C, self._lib, self._check = C, _native.lib(), _check
        h = C.c_void_p()
        _check(self._lib.scn_exchange_create(device, rank, world, max_nq, k, dim, C.byref(h)))
        self._h, self.rank, self.world, self.k, self.max_nq, self.dim = h, rank, world, k, max_nq, dim

    @property
    def handle(self):
        return self._h

    def slice(self, nq: int, rank: int = None) -> Tuple[int, int]:
        """(first query, number of queries) of the slice `rank` (default: this rank) answers."""
        C = self._C
        lo, cnt = C.c_uint64(0), C.c_uint64(0)
        self._check(self._lib.scn_exchange_slice(self._h, nq, self.rank if rank is None else rank, C.byref(lo), C.byref(cnt)))
        return int(lo.value), int(cnt.value)

    def local_handle(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's buffer (to be all-gathered between processes)."""
        buf = self._C.create_string_buffer(self.HANDLE_BYTES)
        self._check(self._lib.scn_exchange_local_handle(self._h, buf))
        return buf.raw

    def connect(self, handles) -> None:
        """handles[r] = rank r's local_handle() (another process)."""
        blob = b"".join(bytes(h) for h in handles)
        assert len(blob) == self.world * self.HANDLE_BYTES
        self._check(self._lib.scn_exchange_connect(self._h, blob))

    def connect_local(self, peers) -> None:
        """peers[r] = rank r's ShardExchange in this process (the reference server's shape)."""
        arr = (self._C.c_void_p * self.world)(*[p.handle for p in peers])
        self._check(self._lib.scn_exchange_connect_local(self._h, arr))

    def search(self, store, q_ptr: int, nq: int, row_base: int, out_ids_ptr: int, out_dist_ptr: int, out_cnt_ptr: int,
               stream: int, q_is_slice: bool = False) -> None:
        """Device buffers, asynchronous on `stream`. q_ptr: the whole batch, or (q_is_slice) this rank's
        slice only; the outputs receive the results of this rank's slice."""
        C = self._C
        self._check(self._lib.scn_search_flat_exchange_dev(store.handle, self._h, C.c_void_p(q_ptr), 1 if q_is_slice else 0,
                                                            nq, self.k, row_base, C.c_void_p(out_ids_ptr),
                                                            C.c_void_p(out_dist_ptr), C.c_void_p(out_cnt_ptr),
                                                            C.c_void_p(stream)))

    def search_host(self, store, q_slice_ptr: int, nq: int, row_base: int, out_ids_ptr: int, out_dist_ptr: int,
                    out_cnt_ptr: int) -> None:
        """Host buffers, blocking (scn_search_flat_exchange): this rank's query slice in, its results out."""
        C = self._C
        self._check(self._lib.scn_search_flat_exchange(store.handle, self._h, C.c_void_p(q_slice_ptr), nq, self.k, row_base,
                                                        C.c_void_p(out_ids_ptr), C.c_void_p(out_dist_ptr),
                                                        C.c_void_p(out_cnt_ptr)))

    def status(self, stream: int) -> None:
        self._check(self._lib.scn_exchange_status(self._h, self._C.c_void_p(stream)))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.scn_exchange_destroy(h)

    __del__ = close


class ShardedStore:
    """`scn_shards`: one collection row-sharded over several GPUs of this process, searched with ONE
    blocking host-buffer call (the multi-device form of core.VectorIndex.Search, interfaces.go:87-111)."""

    def __init__(self, devices, dim: int, metric, capacity_rows: int):
        import ctypes as C

        import numpy as np

        from . import _native
        from .index import _check

        self._C, self._np, self._lib, self._check = C, np, _native.lib(), _check
        self._native = _native
        devs = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        _check(self._lib.scn_shards_create(devs, len(devices), dim, int(metric), capacity_rows, C.byref(h)))
        self._h, self.dim, self.world = h, dim, len(devices)

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.scn_shards_destroy(h)

    __del__ = close

    def append(self, vectors, ids=None) -> None:
        np = self._np
        v = np.ascontiguousarray(vectors, dtype=np.float32)
        ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        self._check(self._lib.scn_shards_append(self._h, v.ctypes.data_as(self._C.c_void_p),
                                                None if ids_a is None else ids_a.ctypes.data_as(self._C.c_void_p), v.shape[0]))

    def mark_deleted(self, ids) -> None:
        a = self._np.ascontiguousarray(ids, dtype=self._np.uint64).ravel()
        self._check(self._lib.scn_shards_mark_deleted(self._h, a.ctypes.data_as(self._C.c_void_p), a.size))

    def set_option(self, name: str, value: int) -> None:
        self._check(self._lib.scn_shards_set_option(self._h, name.encode(), value))

    def stats(self):
        st = self._native.Stats()
        self._check(self._lib.scn_shards_stats(self._h, self._C.byref(st)))
        return st

    def search_flat(self, queries, k: int):
        np, C = self._np, self._C
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq = q.shape[0]
        ids = np.zeros((nq, max(k, 1)), np.uint64)
        dist = np.full((nq, max(k, 1)), np.inf, np.float32)
        cnt = np.zeros(nq, np.uint32)
        self._check(self._lib.scn_shards_search_flat(self._h, q.ctypes.data_as(C.c_void_p), nq, k, ids.ctypes.data_as(C.c_void_p),
                                                     dist.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p)))
        return ids, dist, cnt

    def search_flat_ptr(self, q_ptr: int, nq: int, k: int, ids_ptr: int, dist_ptr: int, cnt_ptr: int) -> None:
        """Raw host pointers (e.g. pinned buffers from scn_host_alloc)."""
        C = self._C
        self._check(self._lib.scn_shards_search_flat(self._h, C.c_void_p(q_ptr), nq, k, C.c_void_p(ids_ptr), C.c_void_p(dist_ptr),
                                                     C.c_void_p(cnt_ptr)))

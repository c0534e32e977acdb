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


class ShardExchange:
    """One rank's end of the fused shard exchange (`scn_exchange_*`): the shard-local search stores
    its top-k lists straight into every rank's buffer over NVLink peer memory and each rank merges
    from local memory — no all-gather. Every rank must issue `search` for the same batch."""

    HANDLE_BYTES = 64

    def __init__(self, device: int, rank: int, world: int, max_nq: int, k: int):
        import ctypes as C

        from . import _native
        from .index import _check

        self._C, self._lib, self._check = C, _native.lib(), _check
        h = C.c_void_p()
        _check(self._lib.scn_exchange_create(device, rank, world, max_nq, k, C.byref(h)))
        self._h, self.rank, self.world, self.k, self.max_nq = h, rank, world, k, max_nq

    @property
    def handle(self):
        return self._h

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
               stream: int) -> None:
        C = self._C
        self._check(self._lib.scn_search_flat_exchange_dev(store.handle, self._h, C.c_void_p(q_ptr), nq, self.k, row_base,
                                                            C.c_void_p(out_ids_ptr), C.c_void_p(out_dist_ptr),
                                                            C.c_void_p(out_cnt_ptr), C.c_void_p(stream)))

    def status(self, stream: int) -> None:
        self._check(self._lib.scn_exchange_status(self._h, self._C.c_void_p(stream)))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.scn_exchange_destroy(h)

    __del__ = close

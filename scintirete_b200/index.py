"""Host-side mirror of the reference's index/distance interfaces for the search hot path.

Same names, argument meaning and error behaviour as internal/core/interfaces.go
(VectorIndex 87-111, HNSWIndex 114-134, DistanceCalculator 154-163, IndexFactory 187-196) and
internal/core/algorithm/distance.go, on top of the C ABI in include/scn_gpu.h. This is what the Go
`GPUIndex` in go/ does through cgo; Go is not installed in this image, so this Python twin is the
tested binding. All arithmetic happens on the GPU: there is no CPU fallback anywhere in here.

Graph construction (HNSW.Insert/Build: searchLayer with efConstruction, selectNeighbors,
pruneConnections — hnsw.go:190-257, 560-614) is GPU-assisted with the reference's serial semantics
(``scn_hnsw_insert``: speculative window searches on the device, in-order commits); a graph built
elsewhere is handed over with ``import_graph_state`` exactly as persistence does on restore
(database.go:398-493 -> hnsw.go:749-804).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .types import (DistanceMetric, ErrorCode, GraphState, HNSWParams, ScintireteError, SearchParams, SearchResult,
                    Vector)


def _check(rc: int) -> None:
    if rc != 0:
        raise ScintireteError(rc, _native.last_error())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- DistanceCalculator (distance.go) --------------------------------------------------------

class DistanceCalculator:
    """core.DistanceCalculator backed by scn_distance_batch (exact reference arithmetic on GPU)."""

    def __init__(self, metric: DistanceMetric, device: int = 0):
        self._metric = DistanceMetric(metric)
        self._device = device

    def distance(self, a, b) -> np.float32:
        a, b = _f32(a).ravel(), _f32(b).ravel()
        if a.size != b.size:  # distance.go:22-24: +Inf, not an error
            return np.float32(np.inf)
        return self.pairwise(a[None, :], b[None, :])[0, 0]

    def pairwise(self, queries, targets) -> np.ndarray:
        q, x = _f32(queries), _f32(targets)
        if q.shape[1] != x.shape[1]:
            return np.full((q.shape[0], x.shape[0]), np.inf, np.float32)
        out = np.empty((q.shape[0], x.shape[0]), np.float32)
        _check(_native.lib().scn_distance_batch(self._device, int(self._metric), _ptr(q), q.shape[0], _ptr(x), x.shape[0],
                                                q.shape[1], _ptr(out)))
        return out

    def distance_type(self) -> DistanceMetric:
        return self._metric

    def is_similarity(self) -> bool:  # distance.go:40-42, 90-92, 124-126
        return False


def new_distance_calculator(metric, device: int = 0) -> DistanceCalculator:
    """algorithm.NewDistanceCalculator (distance.go:129-140)."""
    try:
        m = DistanceMetric(metric)
    except ValueError:
        m = DistanceMetric.UNSPECIFIED
    if m == DistanceMetric.UNSPECIFIED:
        raise ScintireteError(ErrorCode.INVALID_PARAMETERS, "unsupported distance metric")
    return DistanceCalculator(m, device)


def batch_distance(calc: DistanceCalculator, query, targets) -> np.ndarray:
    """algorithm.BatchDistance (distance.go:144-150)."""
    t = _f32(targets)
    if t.shape[0] == 0:
        return np.empty(0, np.float32)
    return calc.pairwise(_f32(query)[None, :], t)[0]


# The vector helpers of distance.go:152-192 on the GPU, bit-identical to the Go functions. One
# vector ([dim]) or a batch ([n][dim]) per call.
_VEC_MAGNITUDE, _VEC_NORMALIZE, _VEC_DOT = 1, 2, 3


def _vector_ops(op: int, a, b=None, device: int = 0) -> np.ndarray:
    a = _f32(a)
    single = a.ndim == 1
    a2 = a[None, :] if single else a
    n, dim = a2.shape
    b2 = None
    if b is not None:
        b2 = _f32(b)
        b2 = b2[None, :] if b2.ndim == 1 else b2
    out = np.empty((n, dim) if op == _VEC_NORMALIZE else (n,), np.float32)
    _check(_native.lib().scn_vector_ops(device, op, _ptr(a2), _ptr(b2), n, dim, _ptr(out)))
    return out[0] if single else out


def normalize_vector(vector, device: int = 0) -> np.ndarray:
    """algorithm.NormalizeVector (distance.go:154-172): a zero vector is returned unchanged."""
    return _vector_ops(_VEC_NORMALIZE, vector, device=device)


def vector_magnitude(vector, device: int = 0):
    """algorithm.VectorMagnitude (distance.go:175-181)."""
    return _vector_ops(_VEC_MAGNITUDE, vector, device=device)


def dot_product(a, b, device: int = 0):
    """algorithm.DotProduct (distance.go:184-192): vectors of different lengths give 0."""
    a, b = _f32(a), _f32(b)
    if a.shape != b.shape:
        return np.float32(0.0)
    return _vector_ops(_VEC_DOT, a, b, device=device)


# ---- device store ------------------------------------------------------------------------------

class DeviceStore:
    """Owns one scn_store (device-memory mirror of the node store, hnsw.go:17-26,115)."""

    def __init__(self, dim: int, metric: DistanceMetric, device: int = 0):
        h = C.c_void_p()
        _check(_native.lib().scn_store_create(device, dim, int(metric), C.byref(h)))
        self._h = h
        self.dim, self.metric, self.device = dim, DistanceMetric(metric), device

    @classmethod
    def from_rdb(cls, path: str, database: str, collection: str, device: int = 0) -> Tuple["DeviceStore", "_native.RdbInfo"]:
        """Restore one collection of a reference-written RDB snapshot straight into device memory
        (RDBManager.Load -> RestoreFromSnapshot -> ImportGraphState; rdb.go:179-237, database.go:398-493)."""
        h = C.c_void_p()
        info = _native.RdbInfo()
        _check(_native.lib().scn_store_load_rdb(str(path).encode(), database.encode(), collection.encode(), device,
                                                C.byref(h), C.byref(info)))
        self = cls.__new__(cls)
        self._h = h
        self.dim, self.metric, self.device = int(info.dim), DistanceMetric(info.metric), device
        return self, info

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _native.lib().scn_store_destroy(h)

    __del__ = close

    @property
    def handle(self):
        return self._h

    def reserve(self, rows: int):
        _check(_native.lib().scn_store_reserve(self._h, rows))

    def clear(self):
        _check(_native.lib().scn_store_clear(self._h))

    def append(self, vectors, ids=None):
        v = _f32(vectors)
        if v.ndim == 1:
            v = v[None, :]
        if v.shape[1] != self.dim:
            raise ScintireteError(ErrorCode.DIMENSION_MISMATCH, f"vector has dimension {v.shape[1]}, expected {self.dim}")
        ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        _check(_native.lib().scn_store_append(self._h, _ptr(v), _ptr(ids_a), v.shape[0]))

    def append_device(self, data_ptr: int, n: int, ids=None):
        ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        _check(_native.lib().scn_store_append_dev(self._h, C.c_void_p(data_ptr), _ptr(ids_a), n))

    def mark_deleted(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.uint64).ravel()
        _check(_native.lib().scn_store_mark_deleted(self._h, _ptr(a), a.size))

    def restore_deleted(self, ids):
        """Soft-delete flags of a restored snapshot: entry point / maxLayer stay verbatim (hnsw.go:791-793)."""
        a = np.ascontiguousarray(ids, dtype=np.uint64).ravel()
        _check(_native.lib().scn_store_restore_deleted(self._h, _ptr(a), a.size))

    def compact(self) -> int:
        """Collection.Compact (collection.go:283-313), device half: drops the soft-deleted rows and the
        graph; returns the number of rows removed."""
        n = C.c_uint64(0)
        _check(_native.lib().scn_store_compact(self._h, C.byref(n)))
        return int(n.value)

    def stats(self) -> _native.Stats:
        st = _native.Stats()
        _check(_native.lib().scn_store_stats(self._h, C.byref(st)))
        return st

    def get(self, ids) -> np.ndarray:
        a = np.ascontiguousarray(ids, dtype=np.uint64).ravel()
        out = np.empty((a.size, self.dim), np.float32)
        _check(_native.lib().scn_store_get(self._h, _ptr(a), a.size, _ptr(out)))
        return out

    def graph_upload(self, st: GraphState):
        ids = np.ascontiguousarray(st.node_ids, dtype=np.uint64)
        lc = np.ascontiguousarray(st.list_counts, dtype=np.int32)
        ec = np.ascontiguousarray(st.edge_counts, dtype=np.uint32)
        ed = np.ascontiguousarray(st.edges, dtype=np.uint64)
        _check(_native.lib().scn_graph_upload(self._h, st.m, st.max_layer, st.entry_point, ids.size, _ptr(ids), _ptr(lc),
                                              _ptr(ec), _ptr(ed)))

    def hnsw_insert(self, levels, m: int, ef_construction: int) -> Dict[str, Any]:
        """GPU-assisted HNSW.Insert / Build (scn_hnsw_insert): the next len(levels) rows of the store that
        are not in the graph yet are inserted with the reference's serial semantics (hnsw.go:190-257);
        levels[i] = the selectLayer() draw of the i-th new node. Returns the build statistics."""
        lv = np.ascontiguousarray(levels, dtype=np.int32).ravel()
        st = _native.BuildStats()
        _check(_native.lib().scn_hnsw_insert(self._h, lv.size, lv.ctypes.data_as(_native.i32p), m, ef_construction, C.byref(st)))
        return {f: (list(getattr(st, f)) if f == "conflict_kind" else getattr(st, f)) for f, _ in st._fields_}

    def graph_export(self, m: int = 16) -> GraphState:
        """The store's graph as flattened core.HNSWGraphState (ExportGraphState, hnsw.go:703-746)."""
        nn, nl, ne = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(_native.lib().scn_graph_export_sizes(self._h, C.byref(nn), C.byref(nl), C.byref(ne)))
        ids = np.zeros(nn.value, np.uint64)
        lc = np.zeros(nn.value, np.int32)
        ec = np.zeros(nl.value, np.uint32)
        ed = np.zeros(ne.value, np.uint64)
        entry, ml = C.c_uint64(0), C.c_int32(-1)
        _check(_native.lib().scn_graph_export(self._h, _ptr(ids), _ptr(lc), _ptr(ec), _ptr(ed), C.byref(entry), C.byref(ml)))
        return GraphState(ids, lc, ec, ed, int(entry.value), int(ml.value), int(self.stats().live_rows), m=m)

    def set_option(self, name: str, value: int):
        _check(_native.lib().scn_set_option(self._h, name.encode(), value))

    def last_timings(self) -> Dict[str, Tuple[float, int]]:
        """{kernel: (total ms, launches)} since the previous call (needs option profile=1)."""
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        cnt = (C.c_uint32 * 32)()
        n = _native.lib().scn_last_timings(self._h, names, ms, cnt, 32)
        return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def last_counters(self) -> List[int]:
        c = (C.c_uint64 * 4)()
        n = _native.lib().scn_last_counters(self._h, c, 4)
        return [int(c[i]) for i in range(n)]

    # -- search, host buffers (the call a Go Collection.Search would make) --
    def _prep(self, queries, k):
        q = _f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            raise ScintireteError(ErrorCode.DIMENSION_MISMATCH,
                                  f"query has dimension {q.shape[1]}, expected {self.dim}")
        nq = q.shape[0]
        return q, np.zeros((nq, k), np.uint64), np.full((nq, k), np.inf, np.float32), np.zeros(nq, np.uint32)

    def search_flat(self, queries, k: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        q, ids, dist, cnt = self._prep(queries, max(k, 1))
        _check(_native.lib().scn_search_flat(self._h, _ptr(q), q.shape[0], k, _ptr(ids), _ptr(dist), _ptr(cnt)))
        return ids, dist, cnt

    def search_hnsw(self, queries, k: int, ef: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        q, ids, dist, cnt = self._prep(queries, max(k, 1))
        _check(_native.lib().scn_search_hnsw(self._h, _ptr(q), q.shape[0], k, ef, _ptr(ids), _ptr(dist), _ptr(cnt)))
        return ids, dist, cnt

    def rerank(self, queries, cand_ids, k: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        q, ids, dist, cnt = self._prep(queries, max(k, 1))
        c = np.ascontiguousarray(cand_ids, dtype=np.uint64)
        if c.ndim == 1:
            c = c[None, :]
        _check(_native.lib().scn_rerank(self._h, _ptr(q), q.shape[0], _ptr(c), c.shape[1], k, _ptr(ids), _ptr(dist),
                                        _ptr(cnt)))
        return ids, dist, cnt


class Batcher:
    """scn_batcher: coalesces concurrent single-query searches (the reference's one-query-per-call
    API, collection.go:193-204, called from many goroutines) into batched launches."""

    FLAT, HNSW = 0, 1

    def __init__(self, store: DeviceStore, kind: int, max_batch: int = 1024, window_us: int = 100):
        h = C.c_void_p()
        _check(_native.lib().scn_batcher_create(store.handle, kind, max_batch, window_us, C.byref(h)))
        self._h, self._store, self.kind = h, store, kind

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _check(_native.lib().scn_batcher_destroy(h))

    def search(self, query, k: int, ef: int = 0) -> Tuple[np.ndarray, np.ndarray, int]:
        """One query, blocking; safe to call from many threads at once (ctypes drops the GIL)."""
        q = _f32(query).ravel()
        if q.size != self._store.dim:
            raise ScintireteError(ErrorCode.DIMENSION_MISMATCH, f"query has dimension {q.size}, expected {self._store.dim}")
        ids = np.zeros(max(k, 1), np.uint64)
        dist = np.full(max(k, 1), np.inf, np.float32)
        cnt = C.c_uint32(0)
        _check(_native.lib().scn_batcher_search(self._h, _ptr(q), k, ef, _ptr(ids), _ptr(dist), C.byref(cnt)))
        return ids, dist, int(cnt.value)

    def stats(self) -> Dict[str, int]:
        c = (C.c_uint64 * 4)()
        _check(_native.lib().scn_batcher_stats(self._h, c, 4))
        return {"calls": int(c[0]), "batches": int(c[1]), "launches": int(c[2]), "max_batch": int(c[3])}


class PinnedBuffer:
    """A page-locked host buffer from scn_host_alloc, exposed as a numpy array (query / result
    buffers of the host-buffer entry points are DMA-ed directly from pinned memory)."""

    def __init__(self, shape, dtype):
        self._dtype = np.dtype(dtype)
        self._nbytes = int(np.prod(shape)) * self._dtype.itemsize
        p = C.c_void_p()
        _check(_native.lib().scn_host_alloc(self._nbytes, C.byref(p)))
        self._p = p
        buf = (C.c_char * max(self._nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self._dtype, count=int(np.prod(shape))).reshape(shape)

    @property
    def ptr(self) -> int:
        return self._p.value

    def close(self):
        p, self._p = getattr(self, "_p", None), None
        if p:
            self.array = None
            _native.lib().scn_host_free(p)

    __del__ = close


# ---- VectorIndex / HNSWIndex ---------------------------------------------------------------------

class _GPUIndexBase:
    """Shared VectorIndex plumbing (interfaces.go:87-111)."""

    def __init__(self, dim: int, metric, device: int = 0):
        try:
            m = DistanceMetric(metric)
        except ValueError:
            m = DistanceMetric.UNSPECIFIED
        if m == DistanceMetric.UNSPECIFIED:  # NewHNSW -> NewDistanceCalculator error (hnsw.go:129-132)
            raise ScintireteError(ErrorCode.INVALID_PARAMETERS, "unsupported distance metric")
        self.metric = m
        self.dim = dim
        self.store = DeviceStore(dim, m, device)
        self._metadata: Dict[int, Dict[str, Any]] = {}
        self._deleted: set = set()
        self._batcher: Optional[Batcher] = None

    def enable_micro_batching(self, max_batch: int = 1024, window_us: int = 100) -> None:
        """Route single-query search() calls through a Batcher so that concurrent callers share
        one batched launch."""
        kind = Batcher.HNSW if isinstance(self, GPUHNSWIndex) else Batcher.FLAT
        self._batcher = Batcher(self.store, kind, max_batch, window_us)

    def _search_one(self, query, k: int, ef: int):
        ids, dist, n = self._batcher.search(query, k, ef)
        return ids[None, :], dist[None, :], np.array([n], np.uint32)

    # -- mutation --
    def insert(self, vector: Vector) -> None:
        e = _f32(vector.elements)
        if vector.id == 0 or self._has(vector.id):
            raise ScintireteError(ErrorCode.INSERT_FAILED, f"failed to insert vector {vector.id}")  # hnsw.go:181-183
        try:
            self.store.append(e[None, :], [vector.id])
        except ScintireteError as err:
            if err.code == ErrorCode.INVALID_PARAMETERS:   # e.g. the id of a soft-deleted node: still "already exists"
                raise ScintireteError(ErrorCode.INSERT_FAILED, f"failed to insert vector {vector.id}") from err
            raise
        if vector.metadata:
            self._metadata[vector.id] = vector.metadata

    def _has(self, id_: int) -> bool:
        try:
            self.store.get([id_])
            return True
        except ScintireteError:
            return False

    def delete(self, id_: str) -> None:
        try:
            vid = int(str(id_).strip())  # hnsw.go:265-268 Sscanf("%d")
        except ValueError:
            raise ScintireteError(ErrorCode.INVALID_PARAMETERS, f"invalid ID format: {id_}")
        self.store.mark_deleted([vid])
        self._deleted.add(vid)

    # -- read --
    def get(self, id_: str) -> Vector:
        try:
            vid = int(str(id_).strip())
        except ValueError:
            raise ScintireteError(ErrorCode.INVALID_PARAMETERS, f"invalid ID format: {id_}")
        if vid in self._deleted:
            raise ScintireteError(ErrorCode.VECTOR_NOT_FOUND, f"vector {id_} not found")  # hnsw.go:364-366
        return Vector(vid, self.store.get([vid])[0], self._metadata.get(vid))

    def size(self) -> int:
        return int(self.store.stats().live_rows)

    def memory_usage(self) -> int:
        return int(self.store.stats().device_bytes)

    def _results(self, ids, dist, cnt, include_vector: bool) -> List[List[SearchResult]]:
        out = []
        for qi in range(ids.shape[0]):
            n = int(cnt[qi])
            vecs = self.store.get(ids[qi, :n]) if (include_vector and n) else None
            out.append([SearchResult(Vector(int(ids[qi, j]), None if vecs is None else vecs[j],
                                            self._metadata.get(int(ids[qi, j]))), float(dist[qi, j])) for j in range(n)])
        return out


class GPUFlatIndex(_GPUIndexBase):
    """Exact-scan VectorIndex ("flat-gpu"): BatchDistance + stable sort, on the GPU."""

    def build(self, vectors: Sequence[Vector]) -> None:  # VectorIndex.Build: clear, then insert all
        self.store.clear()
        self._metadata.clear()
        self._deleted.clear()
        if len(vectors):
            self.store.append(np.stack([_f32(v.elements) for v in vectors]), [v.id for v in vectors])
            for v in vectors:
                if v.metadata:
                    self._metadata[v.id] = v.metadata

    def search_batch(self, queries, params: SearchParams):
        if params.top_k <= 0:
            raise ScintireteError(ErrorCode.INVALID_PARAMETERS, "top_k must be positive")
        return self.store.search_flat(queries, params.top_k)

    def search(self, query, params: SearchParams, include_vector: bool = False) -> List[SearchResult]:
        if self._batcher is not None and params.top_k > 0:
            return self._results(*self._search_one(query, params.top_k, 0), include_vector)[0]
        return self._results(*self.search_batch(_f32(query)[None, :], params), include_vector)[0]

    def compact(self) -> int:
        """Collection.Compact for the flat index: the deleted rows are dropped on the device."""
        n = self.store.compact()
        for vid in self._deleted:
            self._metadata.pop(vid, None)
        self._deleted.clear()
        return n

    def get_statistics(self):
        st = self.store.stats()
        return {"nodes": int(st.live_rows), "memory_usage": int(st.device_bytes)}


class GPUHNSWIndex(_GPUIndexBase):
    """core.HNSWIndex ("hnsw-gpu") whose Search runs on the GPU over a graph built by the
    reference algorithm and handed over via import_graph_state."""

    def __init__(self, params: HNSWParams, metric, dim: int, device: int = 0):
        super().__init__(dim, metric, device)
        self.params = params
        self._graph: Optional[GraphState] = None

    def _select_layer(self) -> int:
        """selectLayer (hnsw.go:458-469): floor(-ln(U) * 1/ln 2), capped at MaxLayers - 1. The draws come
        from this index's own seeded generator (the Go shim uses the CPU index's math/rand stream)."""
        if not hasattr(self, "_rng"):
            self._rng = np.random.default_rng(self.params.seed)
        u = max(float(self._rng.random()), 2.0 ** -53)
        return min(int(np.floor(-np.log(u) * (1.0 / np.log(2.0)))), self.params.max_layers - 1)

    def build(self, vectors: Sequence[Vector], levels=None) -> Dict[str, Any]:
        """HNSW.Build (hnsw.go:148-174): clear, then insert every vector in slice order — the searches run
        on the GPU, the graph is the one the reference's serial insertVector builds for the same level
        draws (`levels`, default: this index's own selectLayer stream)."""
        self.store.clear()
        self._metadata.clear()
        self._deleted.clear()
        self._graph = None
        if not len(vectors):
            return {}
        self.store.append(np.stack([_f32(v.elements) for v in vectors]), [v.id for v in vectors])
        for v in vectors:
            if v.metadata:
                self._metadata[v.id] = v.metadata
        if levels is None:
            levels = [self._select_layer() for _ in vectors]
        return self.store.hnsw_insert(levels, self.params.m, self.params.ef_construction)

    def insert(self, vector: Vector, level: Optional[int] = None) -> None:
        """HNSW.Insert (hnsw.go:177-187): store the vector and link it into the graph on the GPU."""
        super().insert(vector)
        self.store.hnsw_insert([self._select_layer() if level is None else level], self.params.m, self.params.ef_construction)
        self._graph = None

    def get_parameters(self) -> HNSWParams:
        return self.params

    def set_ef_search(self, ef_search: int) -> None:  # hnsw.go:449-453
        self.params.ef_search = ef_search

    def get_layers(self) -> int:  # hnsw.go:394-401
        ml = self.store.stats().max_layer
        return 0 if ml < 0 else ml + 1

    def import_graph_state(self, state: GraphState) -> None:
        """hnsw.go:749-804: replace all nodes, then take entrypoint / maxLayer / size verbatim."""
        if state.vectors is None:
            raise ScintireteError(ErrorCode.INVALID_PARAMETERS, "graph state carries no vectors")
        self.store.clear()
        self._deleted.clear()
        self.store.append(state.vectors, state.node_ids)
        self.store.graph_upload(state)
        if state.deleted is not None and np.any(state.deleted):
            dead = np.asarray(state.node_ids)[np.asarray(state.deleted).astype(bool)]
            self.store.restore_deleted(dead)
            self._deleted.update(int(x) for x in dead)
        self._graph = state

    def export_graph_state(self) -> Optional[GraphState]:
        if self._graph is None and self.store.stats().has_graph:
            self._graph = self.store.graph_export(self.params.m)
        return self._graph

    def get_graph_statistics(self):  # hnsw.go:404-443
        st = self.store.stats()
        live = int(st.live_rows)
        return {"layers": int(st.max_layer) + 1, "nodes": live, "connections": int(st.graph_edges),
                "avg_degree": (st.graph_edges / live) if live else 0.0, "memory_usage": int(st.device_bytes)}

    get_statistics = get_graph_statistics

    def _ef(self, params: SearchParams) -> int:  # hnsw.go:300-303
        if params.ef_search is not None and params.ef_search > 0:
            return params.ef_search
        return self.params.ef_search

    def search_batch(self, queries, params: SearchParams):
        if params.top_k <= 0:
            raise ScintireteError(ErrorCode.INVALID_PARAMETERS, "top_k must be positive")
        return self.store.search_hnsw(queries, params.top_k, self._ef(params))

    def search(self, query, params: SearchParams, include_vector: bool = False) -> List[SearchResult]:
        if self._batcher is not None and params.top_k > 0:
            return self._results(*self._search_one(query, params.top_k, self._ef(params)), include_vector)[0]
        return self._results(*self.search_batch(_f32(query)[None, :], params), include_vector)[0]

    def search_exact(self, queries, params: SearchParams):
        """Flat ground truth over the same rows (SURVEY.md §8b `SearchExact`)."""
        return self.store.search_flat(queries, params.top_k)


class IndexFactory:
    """core.IndexFactory (interfaces.go:187-196) for the two GPU index types."""

    def __init__(self, device: int = 0):
        self.device = device

    def create_index(self, config: Dict[str, Any]):
        kind = config.get("type", "hnsw-gpu")
        metric, dim = config["metric"], config["dim"]
        if kind == "flat-gpu":
            return GPUFlatIndex(dim, metric, self.device)
        if kind == "hnsw-gpu":
            p = config.get("parameters", {})
            hp = HNSWParams(**{k: p[k] for k in ("m", "ef_construction", "ef_search", "max_layers", "seed") if k in p})
            return GPUHNSWIndex(hp, metric, dim, self.device)
        raise ScintireteError(ErrorCode.INVALID_PARAMETERS, f"unsupported index type {kind}")

    def supported_metrics(self) -> List[DistanceMetric]:
        return [DistanceMetric.L2, DistanceMetric.COSINE, DistanceMetric.INNER_PRODUCT]

    def default_parameters(self) -> Dict[str, Any]:
        return {"m": 16, "ef_construction": 200, "ef_search": 50, "max_layers": 16}

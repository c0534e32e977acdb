"""CPU oracle bindings — TEST INFRASTRUCTURE ONLY.

ctypes wrapper over ``oracle/liboracle.so`` (built from ``scn_oracle.cpp``, a line-faithful
restatement of the reference's ``internal/core/algorithm/distance.go`` and ``hnsw.go``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
package. The product (``scintirete_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

METRIC_L2, METRIC_COSINE, METRIC_IP = 1, 2, 3


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "scn_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None

_f32p = C.POINTER(C.c_float)
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.scn_oracle_distance.argtypes = [C.c_int32, _f32p, C.c_uint64, _f32p, C.c_uint64, _f32p]
        L.scn_oracle_distance.restype = C.c_int32
        L.scn_oracle_batch_distance.argtypes = [C.c_int32, _f32p, C.c_uint64, _f32p, C.c_uint64, _f32p]
        L.scn_oracle_batch_distance.restype = C.c_int32
        L.scn_oracle_normalize.argtypes = [_f32p, C.c_uint64, _f32p]
        L.scn_oracle_normalize.restype = C.c_int32
        L.scn_oracle_magnitude.argtypes = [_f32p, C.c_uint64]
        L.scn_oracle_magnitude.restype = C.c_float
        L.scn_oracle_dot.argtypes = [_f32p, C.c_uint64, _f32p, C.c_uint64]
        L.scn_oracle_dot.restype = C.c_float
        L.scn_oracle_flat_search.argtypes = [C.c_int32, _f32p, C.c_uint64, C.c_uint64, _u64p, _u8p, _f32p,
                                             C.c_uint64, C.c_uint32, _u64p, _f32p, _u32p, C.c_int32]
        L.scn_oracle_flat_search.restype = C.c_int32
        L.scn_oracle_hnsw_new.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32]
        L.scn_oracle_hnsw_new.restype = C.c_void_p
        L.scn_oracle_hnsw_free.argtypes = [C.c_void_p]
        L.scn_oracle_hnsw_free.restype = None
        L.scn_oracle_hnsw_insert.argtypes = [C.c_void_p, C.c_uint64, _f32p, C.c_uint64]
        L.scn_oracle_hnsw_insert.restype = C.c_int32
        L.scn_oracle_hnsw_insert_level.argtypes = [C.c_void_p, C.c_uint64, _f32p, C.c_uint64, C.c_int32]
        L.scn_oracle_hnsw_insert_level.restype = C.c_int32
        L.scn_oracle_hnsw_build.argtypes = [C.c_void_p, _u64p, _f32p, C.c_uint64, C.c_uint64]
        L.scn_oracle_hnsw_build.restype = C.c_int32
        L.scn_oracle_hnsw_delete.argtypes = [C.c_void_p, C.c_uint64]
        L.scn_oracle_hnsw_delete.restype = C.c_int32
        L.scn_oracle_hnsw_set_ef_search.argtypes = [C.c_void_p, C.c_int32]
        L.scn_oracle_hnsw_set_ef_search.restype = None
        for name in ("size", "layers", "max_layer"):
            f = getattr(L, "scn_oracle_hnsw_" + name)
            f.argtypes = [C.c_void_p]
            f.restype = C.c_int32
        for name in ("entrypoint", "node_count"):
            f = getattr(L, "scn_oracle_hnsw_" + name)
            f.argtypes = [C.c_void_p]
            f.restype = C.c_uint64
        L.scn_oracle_hnsw_search.argtypes = [C.c_void_p, _f32p, C.c_uint64, C.c_int32, C.c_int32, _u64p, _f32p, _u64p]
        L.scn_oracle_hnsw_search.restype = C.c_int32
        L.scn_oracle_hnsw_search_batch.argtypes = [C.c_void_p, _f32p, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32,
                                                   _u64p, _f32p, _u32p, C.c_int32, _u64p]
        L.scn_oracle_hnsw_search_batch.restype = C.c_int32
        L.scn_oracle_hnsw_search_layer.argtypes = [C.c_void_p, _f32p, C.c_uint64, _u64p, C.c_uint32, C.c_int32,
                                                   C.c_int32, _u64p]
        L.scn_oracle_hnsw_search_layer.restype = C.c_int32
        L.scn_oracle_hnsw_export_sizes.argtypes = [C.c_void_p, _u64p, _u64p, _u64p]
        L.scn_oracle_hnsw_export_sizes.restype = None
        L.scn_oracle_hnsw_export.argtypes = [C.c_void_p, _u64p, _u8p, _i32p, _u32p, _u64p, _f32p]
        L.scn_oracle_hnsw_export.restype = None
        L.scn_oracle_hnsw_import.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, _u64p, _u8p, _i32p, _u32p, _u64p,
                                             _f32p, C.c_uint64, C.c_int32, C.c_int32]
        L.scn_oracle_hnsw_import.restype = C.c_int32
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class OracleError(Exception):
    def __init__(self, code, msg=""):
        super().__init__(f"oracle error {code} {msg}")
        self.code = code


# ---- distance.go ---------------------------------------------------------------------------

def distance(metric: int, a, b) -> np.float32:
    a, b = _f32(a), _f32(b)
    out = C.c_float()
    rc = lib().scn_oracle_distance(metric, _p(a, _f32p), a.size, _p(b, _f32p), b.size, C.byref(out))
    if rc:
        raise OracleError(rc, "unsupported distance metric")
    return np.float32(out.value)


def batch_distance(metric: int, query, targets) -> np.ndarray:
    q, t = _f32(query), _f32(targets)
    out = np.empty(t.shape[0], np.float32)
    rc = lib().scn_oracle_batch_distance(metric, _p(q, _f32p), q.size, _p(t, _f32p), t.shape[0], _p(out, _f32p))
    if rc:
        raise OracleError(rc)
    return out


def normalize(v) -> np.ndarray:
    v = _f32(v)
    out = np.empty_like(v)
    lib().scn_oracle_normalize(_p(v, _f32p), v.size, _p(out, _f32p))
    return out


def magnitude(v) -> np.float32:
    v = _f32(v)
    return np.float32(lib().scn_oracle_magnitude(_p(v, _f32p), v.size))


def dot(a, b) -> np.float32:
    a, b = _f32(a), _f32(b)
    return np.float32(lib().scn_oracle_dot(_p(a, _f32p), a.size, _p(b, _f32p), b.size))


# ---- flat scan -----------------------------------------------------------------------------

def flat_search(metric: int, db, queries, k: int, ids=None, deleted=None, nthreads: int = 1):
    """Exact scan oracle: (ids[nq,k] u64, dist[nq,k] f32, counts[nq] u32)."""
    db, q = _f32(db), _f32(queries)
    if q.ndim == 1:
        q = q[None, :]
    n, dim = db.shape if db.ndim == 2 else (0, q.shape[1])
    nq = q.shape[0]
    out_ids = np.zeros((nq, k), np.uint64)
    out_d = np.full((nq, k), np.inf, np.float32)
    counts = np.zeros(nq, np.uint32)
    ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
    del_a = None if deleted is None else np.ascontiguousarray(deleted, dtype=np.uint8)
    rc = lib().scn_oracle_flat_search(metric, _p(db, _f32p), n, dim, _p(ids_a, _u64p), _p(del_a, _u8p),
                                      _p(q, _f32p), nq, k, _p(out_ids, _u64p), _p(out_d, _f32p),
                                      _p(counts, _u32p), nthreads)
    if rc:
        raise OracleError(rc)
    return out_ids, out_d, counts


# ---- HNSW ----------------------------------------------------------------------------------

class GraphState:
    """Flattened core.HNSWGraphState (interfaces.go:137-151); node order = insertion order."""

    def __init__(self, ids, deleted, list_counts, edge_counts, edges, vectors, entrypoint, max_layer, size):
        self.ids, self.deleted, self.list_counts = ids, deleted, list_counts
        self.edge_counts, self.edges, self.vectors = edge_counts, edges, vectors
        self.entrypoint, self.max_layer, self.size = int(entrypoint), int(max_layer), int(size)


class OracleHNSW:
    """Restatement of algorithm.HNSW (hnsw.go:107-145)."""

    def __init__(self, M=16, ef_construction=200, ef_search=50, max_layers=16, seed=42, metric=METRIC_L2):
        self._h = lib().scn_oracle_hnsw_new(M, ef_construction, ef_search, max_layers, seed, metric)
        if not self._h:
            raise OracleError(3007, "unsupported distance metric")
        self.params = dict(M=M, ef_construction=ef_construction, ef_search=ef_search, max_layers=max_layers, seed=seed)
        self.metric = metric
        self.dim = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.scn_oracle_hnsw_free(h)

    def insert(self, id_: int, vec, level: int | None = None):
        v = _f32(vec)
        self.dim = v.size
        if level is None:
            rc = lib().scn_oracle_hnsw_insert(self._h, id_, _p(v, _f32p), v.size)
        else:
            rc = lib().scn_oracle_hnsw_insert_level(self._h, id_, _p(v, _f32p), v.size, level)
        if rc:
            raise OracleError(rc, f"failed to insert vector {id_}")

    def build(self, vectors, ids=None):
        v = _f32(vectors)
        self.dim = v.shape[1]
        ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        rc = lib().scn_oracle_hnsw_build(self._h, _p(ids_a, _u64p), _p(v, _f32p), v.shape[0], v.shape[1])
        if rc:
            raise OracleError(rc)

    def delete(self, id_: int):
        rc = lib().scn_oracle_hnsw_delete(self._h, id_)
        if rc:
            raise OracleError(rc, "vector not found")

    def set_ef_search(self, ef: int):
        lib().scn_oracle_hnsw_set_ef_search(self._h, ef)
        self.params["ef_search"] = ef

    def size(self):
        return lib().scn_oracle_hnsw_size(self._h)

    def layers(self):
        return lib().scn_oracle_hnsw_layers(self._h)

    def entrypoint(self):
        return lib().scn_oracle_hnsw_entrypoint(self._h)

    def max_layer(self):
        return lib().scn_oracle_hnsw_max_layer(self._h)

    def node_count(self):
        return lib().scn_oracle_hnsw_node_count(self._h)

    def search(self, query, top_k: int, ef_search: int | None = None, with_stats=False):
        q = _f32(query)
        ids = np.zeros(max(top_k, 1), np.uint64)
        ds = np.zeros(max(top_k, 1), np.float32)
        st = np.zeros(2, np.uint64)
        n = lib().scn_oracle_hnsw_search(self._h, _p(q, _f32p), q.size, top_k, ef_search or 0, _p(ids, _u64p),
                                         _p(ds, _f32p), _p(st, _u64p))
        if with_stats:
            return ids[:n], ds[:n], (int(st[0]), int(st[1]))
        return ids[:n], ds[:n]

    def search_batch(self, queries, top_k: int, ef_search: int | None = None, nthreads: int = 1):
        q = _f32(queries)
        nq, dim = q.shape
        ids = np.zeros((nq, top_k), np.uint64)
        ds = np.zeros((nq, top_k), np.float32)
        counts = np.zeros(nq, np.uint32)
        st = np.zeros(2, np.uint64)
        lib().scn_oracle_hnsw_search_batch(self._h, _p(q, _f32p), nq, dim, top_k, ef_search or 0, _p(ids, _u64p),
                                           _p(ds, _f32p), _p(counts, _u32p), nthreads, _p(st, _u64p))
        return ids, ds, counts, (int(st[0]), int(st[1]))

    def search_layer(self, query, entry_ids, ef: int, layer: int):
        q = _f32(query)
        e = np.ascontiguousarray(entry_ids, dtype=np.uint64)
        out = np.zeros(max(ef, len(e), 1), np.uint64)
        n = lib().scn_oracle_hnsw_search_layer(self._h, _p(q, _f32p), q.size, _p(e, _u64p), e.size, ef, layer,
                                               _p(out, _u64p))
        return out[:n]

    def export_graph_state(self, with_vectors=True) -> GraphState:
        nn, nl, ne = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().scn_oracle_hnsw_export_sizes(self._h, C.byref(nn), C.byref(nl), C.byref(ne))
        ids = np.zeros(nn.value, np.uint64)
        deleted = np.zeros(nn.value, np.uint8)
        lc = np.zeros(nn.value, np.int32)
        ec = np.zeros(nl.value, np.uint32)
        edges = np.zeros(ne.value, np.uint64)
        vec = np.zeros((nn.value, self.dim or 0), np.float32) if with_vectors and self.dim else None
        lib().scn_oracle_hnsw_export(self._h, _p(ids, _u64p), _p(deleted, _u8p), _p(lc, _i32p), _p(ec, _u32p),
                                     _p(edges, _u64p), _p(vec, _f32p))
        return GraphState(ids, deleted, lc, ec, edges, vec, self.entrypoint(), self.max_layer(), self.size())

    def import_graph_state(self, st: GraphState):
        self.dim = st.vectors.shape[1]
        v = _f32(st.vectors)
        lib().scn_oracle_hnsw_import(self._h, len(st.ids), self.dim, _p(st.ids, _u64p), _p(st.deleted, _u8p),
                                     _p(st.list_counts, _i32p), _p(st.edge_counts, _u32p), _p(st.edges, _u64p),
                                     _p(v, _f32p), st.entrypoint, st.max_layer, st.size)

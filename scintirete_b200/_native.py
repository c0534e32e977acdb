"""ctypes binding of libscn_gpu.so — the same C ABI the Go cgo shim binds (include/scn_gpu.h).

There is no CPU fallback: if the library is missing or no B200 is present, calls fail loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscn_gpu.so")

f32p = C.POINTER(C.c_float)
u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)


class Stats(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("live_rows", C.c_uint64), ("capacity_rows", C.c_uint64),
                ("device_bytes", C.c_uint64), ("dim", C.c_uint32), ("metric", C.c_int32), ("device", C.c_int32),
                ("has_graph", C.c_int32), ("max_layer", C.c_int32), ("m", C.c_int32), ("entry_id", C.c_uint64),
                ("graph_edges", C.c_uint64)]


class BuildStats(C.Structure):
    _fields_ = [("inserted", C.c_uint64), ("rounds", C.c_uint64), ("searches", C.c_uint64), ("conflicts", C.c_uint64),
                ("table_overflows", C.c_uint64), ("distance_evals", C.c_uint64), ("expansions", C.c_uint64),
                ("conflict_kind", C.c_uint64 * 6), ("seconds", C.c_double), ("device_seconds", C.c_double),
                ("commit_seconds", C.c_double)]


class RdbInfo(C.Structure):
    _fields_ = [("metric", C.c_int32), ("dim", C.c_uint32), ("m", C.c_int32), ("ef_construction", C.c_int32),
                ("ef_search", C.c_int32), ("max_layers", C.c_int32), ("seed", C.c_int64), ("nodes", C.c_uint64),
                ("deleted", C.c_uint64), ("entry_id", C.c_uint64), ("max_layer", C.c_int32), ("graph_size", C.c_int32),
                ("vector_count", C.c_int64), ("deleted_count", C.c_int64), ("has_graph", C.c_int32)]


# every symbol include/scn_gpu.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "scn_last_error": (C.c_char_p, []),
    "scn_launch_count": (C.c_uint64, []),
    "scn_host_alloc": (C.c_int32, [C.c_uint64, C.POINTER(C.c_void_p)]),
    "scn_host_free": (C.c_int32, [C.c_void_p]),
    "scn_store_create": (C.c_int32, [C.c_int32, C.c_uint32, C.c_int32, C.POINTER(C.c_void_p)]),
    "scn_store_destroy": (C.c_int32, [C.c_void_p]),
    "scn_store_reserve": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "scn_store_clear": (C.c_int32, [C.c_void_p]),
    "scn_store_append": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "scn_store_append_dev": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "scn_store_mark_deleted": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "scn_store_restore_deleted": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "scn_store_compact": (C.c_int32, [C.c_void_p, u64p]),
    "scn_store_stats": (C.c_int32, [C.c_void_p, C.POINTER(Stats)]),
    "scn_store_get": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "scn_graph_upload": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "scn_hnsw_insert": (C.c_int32, [C.c_void_p, C.c_uint64, i32p, C.c_int32, C.c_int32, C.POINTER(BuildStats)]),
    "scn_graph_export_sizes": (C.c_int32, [C.c_void_p, u64p, u64p, u64p]),
    "scn_graph_export": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, u64p, i32p]),
    "scn_search_flat": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scn_search_hnsw": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "scn_rerank": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                C.c_void_p, C.c_void_p]),
    "scn_distance_batch": (C.c_int32, [C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32,
                                        C.c_void_p]),
    "scn_vector_ops": (C.c_int32, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]),
    "scn_search_flat_dev": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "scn_search_hnsw_dev": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "scn_search_flat_shard_dev": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p,
                                               C.c_void_p, C.c_void_p]),
    "scn_merge_topk_dev": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "scn_store_load_rdb": (C.c_int32, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(C.c_void_p),
                                        C.POINTER(RdbInfo)]),
    "scn_exchange_create": (C.c_int32, [C.c_int32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32,
                                         C.POINTER(C.c_void_p)]),
    "scn_exchange_slice": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint32, u64p, u64p]),
    "scn_search_flat_exchange": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64,
                                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "scn_shards_create": (C.c_int32, [i32p, C.c_int32, C.c_uint32, C.c_int32, C.c_uint64, C.POINTER(C.c_void_p)]),
    "scn_shards_destroy": (C.c_int32, [C.c_void_p]),
    "scn_shards_count": (C.c_int32, [C.c_void_p]),
    "scn_shards_store": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "scn_shards_append": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "scn_shards_mark_deleted": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "scn_shards_stats": (C.c_int32, [C.c_void_p, C.POINTER(Stats)]),
    "scn_shards_set_option": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_int64]),
    "scn_shards_search_flat": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "scn_exchange_local_handle": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "scn_exchange_connect": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "scn_exchange_connect_local": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "scn_exchange_destroy": (C.c_int32, [C.c_void_p]),
    "scn_search_flat_exchange_dev": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_uint64, C.c_uint32,
                                                  C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scn_exchange_status": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "scn_batcher_create": (C.c_int32, [C.c_void_p, C.c_int32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "scn_batcher_destroy": (C.c_int32, [C.c_void_p]),
    "scn_batcher_search": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scn_batcher_stats": (C.c_int32, [C.c_void_p, u64p, C.c_int32]),
    "scn_set_option": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_int64]),
    "scn_last_timings": (C.c_int32, [C.c_void_p, C.POINTER(C.c_char_p), f32p, u32p, C.c_int32]),
    "scn_last_counters": (C.c_int32, [C.c_void_p, u64p, C.c_int32]),
    "scn_debug_tensor_scores": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m scintirete_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI and the header diverge
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return (lib().scn_last_error() or b"").decode("utf-8", "replace")

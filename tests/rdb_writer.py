"""Test infrastructure: writes RDB snapshot files in the reference's on-disk format.

The reference serialises snapshots with FlatBuffers (schemas/flatbuffers/rdb.fbs) through the Go
builder, call by call as in internal/persistence/rdb/rdb.go:239-533. Neither Go nor the flatbuffers
package is available here, so this module restates the FlatBuffers builder algorithm (back-to-front
construction, vtables, uoffset/soffset encoding) and replays rdb.go's create* functions on top of
it — same field ids, same child-before-parent order. It is used only by tests/ to feed
scn_store_load_rdb; nothing in the product imports it.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, List, Optional, Sequence

import numpy as np


class Builder:
    """Minimal FlatBuffers builder (grows downwards, like flatbuffers.Builder)."""

    def __init__(self, initial: int = 1024):
        self.bytes = bytearray(initial)
        self.head = initial
        self.minalign = 1
        self.vtable: Optional[List[int]] = None
        self.object_end = 0

    def offset(self) -> int:
        return len(self.bytes) - self.head

    def _grow(self):
        old = len(self.bytes)
        new = bytearray(old * 2)
        new[old:] = self.bytes
        self.bytes = new
        self.head += old

    def pad(self, n: int):
        for _ in range(n):
            self.head -= 1
            self.bytes[self.head] = 0

    def prep(self, size: int, additional: int):
        self.minalign = max(self.minalign, size)
        align = (~(len(self.bytes) - self.head + additional) + 1) & (size - 1)
        while self.head < align + size + additional:
            self._grow()
        self.pad(align)

    def _place(self, fmt: str, v):
        n = struct.calcsize(fmt)
        self.head -= n
        struct.pack_into(fmt, self.bytes, self.head, v)

    def prepend(self, fmt: str, v):
        self.prep(struct.calcsize(fmt), 0)
        self._place(fmt, v)

    def prepend_uoffset(self, off: int):
        self.prep(4, 0)
        assert off <= self.offset()
        self._place("<I", self.offset() - off + 4)

    def start_vector(self, elem_size: int, n: int, alignment: int):
        self.prep(4, elem_size * n)
        self.prep(alignment, elem_size * n)

    def end_vector(self, n: int) -> int:
        self._place("<I", n)
        return self.offset()

    def create_string(self, s) -> int:
        raw = s if isinstance(s, (bytes, bytearray)) else str(s).encode()
        self.prep(4, len(raw) + 1)
        self._place("<B", 0)
        self.head -= len(raw)
        self.bytes[self.head:self.head + len(raw)] = raw
        return self.end_vector(len(raw))

    def create_float_vector(self, a) -> int:
        raw = np.ascontiguousarray(a, dtype="<f4").tobytes()
        n = len(raw) // 4
        self.start_vector(4, n, 4)
        self.head -= len(raw)
        self.bytes[self.head:self.head + len(raw)] = raw
        return self.end_vector(n)

    def create_offset_vector(self, offs: Sequence[int]) -> int:
        self.start_vector(4, len(offs), 4)
        for o in reversed(offs):
            self.prepend_uoffset(o)
        return self.end_vector(len(offs))

    def start_object(self, nfields: int):
        self.vtable = [0] * nfields
        self.object_end = self.offset()

    def slot(self, i: int):
        self.vtable[i] = self.offset()

    def add_scalar(self, i: int, fmt: str, v, default):
        if v != default:
            self.prepend(fmt, v)
            self.slot(i)

    def add_offset(self, i: int, off: int):
        if off:
            self.prepend_uoffset(off)
            self.slot(i)

    def end_object(self) -> int:
        self.prepend("<i", 0)
        obj = self.offset()
        vt = list(self.vtable)
        while vt and vt[-1] == 0:
            vt.pop()
        for f in reversed(vt):
            self.prepend("<H", obj - f if f else 0)
        self.prepend("<H", obj - self.object_end)
        self.prepend("<H", (len(vt) + 2) * 2)
        struct.pack_into("<i", self.bytes, len(self.bytes) - obj, self.offset() - obj)
        self.vtable = None
        return obj

    def finish(self, root: int) -> bytes:
        self.prep(self.minalign, 4)
        self.prepend_uoffset(root)
        return bytes(self.bytes[self.head:])


# ---- rdb.go's create* functions, replayed ---------------------------------------------------------

def _layer_connections(b: Builder, layer: int, ids: Sequence) -> int:      # rdb.go:513-533
    strs = [b.create_string(str(i)) for i in ids]
    vec = b.create_offset_vector(strs)
    b.start_object(2)
    b.add_scalar(0, "<i", int(layer), 0)
    b.add_offset(1, vec)
    return b.end_object()


def _hnsw_node(b: Builder, node_id, elements, metadata, deleted: bool, lists: List[Sequence], id_text=None,
               raw_layers=None) -> int:                                                # rdb.go:466-510
    el = b.create_float_vector(elements)
    # ConvertHNSWGraphState (rdb.go:982-1025): only non-empty lists are written, layer = list index
    layer_entries = raw_layers if raw_layers is not None else [(l, c) for l, c in enumerate(lists) if len(c)]
    lcs = [_layer_connections(b, l, c) for l, c in layer_entries]
    lvec = b.create_offset_vector(lcs)
    ids = b.create_string(id_text if id_text is not None else str(int(node_id)))
    meta = b.create_string(json.dumps(metadata if metadata is not None else None))
    b.start_object(6)
    b.add_offset(0, ids)
    b.add_offset(1, el)
    b.add_offset(2, meta)
    b.add_scalar(3, "<B", 1 if deleted else 0, 0)
    b.add_offset(4, lvec)
    b.add_scalar(5, "<i", len(lists) - 1, 0)
    return b.end_object()


def write_rdb(path, databases: Dict[str, Dict[str, dict]]):
    """databases = {db: {collection: {"metric", "m", "ef_construction", "ef_search", "max_layers",
    "seed", "ids", "vectors", "deleted", "lists" (per node: list of per-layer id lists), "entry",
    "max_layer", optional "no_graph", "id_text" {row: str}, "raw_layers" {row: [(layer, ids)]}}}}"""
    b = Builder()
    db_offs = []
    for db_name, colls in databases.items():
        coll_offs = []
        for c_name, c in colls.items():
            graph = 0
            if not c.get("no_graph"):
                n = len(c["ids"])
                nodes = []
                for i in range(n):
                    nodes.append(_hnsw_node(b, c["ids"][i], c["vectors"][i], c.get("metadata", {}).get(i),
                                            bool(c["deleted"][i]) if c.get("deleted") is not None else False, c["lists"][i],
                                            c.get("id_text", {}).get(i), c.get("raw_layers", {}).get(i)))
                nvec = b.create_offset_vector(nodes)
                ep = b.create_string(c.get("entry_text", str(int(c["entry"]))))
                b.start_object(4)                                      # rdb.go:434-463
                b.add_offset(0, nvec)
                b.add_offset(1, ep)
                b.add_scalar(2, "<i", int(c["max_layer"]), 0)
                b.add_scalar(3, "<i", int(c.get("size", n)), 0)
                graph = b.end_object()
            b.start_object(5)                                          # HNSWParams, rdb.go:422-431
            b.add_scalar(0, "<i", int(c.get("m", 16)), 0)
            b.add_scalar(1, "<i", int(c.get("ef_construction", 200)), 0)
            b.add_scalar(2, "<i", int(c.get("ef_search", 50)), 0)
            b.add_scalar(3, "<i", int(c.get("max_layers", 16)), 0)
            b.add_scalar(4, "<q", int(c.get("seed", 42)), 0)
            hp = b.end_object()
            cname = b.create_string(c_name)
            b.start_object(3)                                          # CollectionConfig, rdb.go:402-419
            b.add_offset(0, cname)
            b.add_scalar(1, "<b", int(c["metric"]), 0)
            b.add_offset(2, hp)
            cfg = b.end_object()
            vectors = b.create_offset_vector([])                        # legacy field, written empty
            name = b.create_string(c_name)
            b.start_object(8)                                          # CollectionSnapshot, rdb.go:320-371
            b.add_offset(0, name)
            b.add_offset(1, cfg)
            b.add_offset(2, vectors)
            b.add_offset(3, graph)
            b.add_scalar(4, "<q", int(len(c["ids"])), 0)
            b.add_scalar(5, "<q", int(np.sum(c["deleted"])) if c.get("deleted") is not None else 0, 0)
            b.add_scalar(6, "<q", 1_700_000_000, 0)
            b.add_scalar(7, "<q", 1_700_000_100, 0)
            coll_offs.append(b.end_object())
        cvec = b.create_offset_vector(coll_offs)
        dname = b.create_string(db_name)
        b.start_object(3)                                              # DatabaseSnapshot, rdb.go:289-317
        b.add_offset(0, dname)
        b.add_offset(1, cvec)
        b.add_scalar(2, "<q", 1_700_000_000, 0)
        db_offs.append(b.end_object())
    dvec = b.create_offset_vector(db_offs)
    version = b.create_string("1.0")
    meta = b.create_string(json.dumps({"created_by": "scintirete"}))
    b.start_object(4)                                                  # RDBSnapshot, rdb.go:239-286
    b.add_offset(0, version)
    b.add_scalar(1, "<q", 1_700_000_200, 0)
    b.add_offset(2, dvec)
    b.add_offset(3, meta)
    root = b.end_object()
    data = b.finish(root)
    with open(path, "wb") as f:
        f.write(data)
    return data


def collection_from_oracle(h, db: np.ndarray, metric: int, m: int = 16) -> dict:
    """CollectionSnapshot content for an oracle-built index (ExportGraphState -> ConvertHNSWGraphState)."""
    st = h.export_graph_state(with_vectors=False)
    lists, e, l = [], 0, 0
    for i in range(len(st.ids)):
        node = []
        for _ in range(int(st.list_counts[i])):
            c = int(st.edge_counts[l])
            node.append([int(x) for x in st.edges[e:e + c]])
            e += c
            l += 1
        lists.append(node)
    rows = st.ids.astype(np.int64) - 1
    return {"metric": metric, "m": m, "ids": [int(x) for x in st.ids], "vectors": db[rows], "deleted": st.deleted.astype(bool),
            "lists": lists, "entry": int(st.entrypoint), "max_layer": int(st.max_layer), "size": int(st.size)}

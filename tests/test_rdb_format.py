"""CPU checks of the RDB test writer itself: the bytes it produces must parse back, field by field,
with an independent pure-Python FlatBuffers walk (so that the GPU-side loader is tested against a
file format that is pinned twice, not against its own mirror image)."""
import struct

import numpy as np

from rdb_writer import write_rdb


def _tbl(buf, pos):
    vt = pos - struct.unpack_from("<i", buf, pos)[0]
    vlen = struct.unpack_from("<H", buf, vt)[0]
    return pos, vt, vlen


def _field(buf, t, i):
    pos, vt, vlen = t
    if 4 + 2 * i + 2 > vlen:
        return 0
    off = struct.unpack_from("<H", buf, vt + 4 + 2 * i)[0]
    return pos + off if off else 0


def _ind(buf, t, i):
    f = _field(buf, t, i)
    return f + struct.unpack_from("<I", buf, f)[0] if f else 0


def _str(buf, p):
    n = struct.unpack_from("<I", buf, p)[0]
    return bytes(buf[p + 4:p + 4 + n]).decode()


def _vec(buf, p):
    n = struct.unpack_from("<I", buf, p)[0]
    return [p + 4 + 4 * i + struct.unpack_from("<I", buf, p + 4 + 4 * i)[0] for i in range(n)]


def test_writer_round_trips_through_an_independent_walk(tmp_path):
    vec = np.arange(12, dtype=np.float32).reshape(3, 4) * 0.5
    coll = {"metric": 2, "m": 8, "ef_construction": 77, "ef_search": 33, "max_layers": 9, "seed": -5,
            "ids": [7, 8, 20], "vectors": vec, "deleted": np.array([False, True, False]),
            "lists": [[[8, 20], [20]], [[7]], [[7, 8], [7], []]], "entry": 20, "max_layer": 2}
    data = write_rdb(tmp_path / "a.rdb", {"db0": {"c0": coll}, "other": {}})
    root = _tbl(data, struct.unpack_from("<I", data, 0)[0])
    assert _str(data, _ind(data, root, 0)) == "1.0"
    dbs = _vec(data, _ind(data, root, 2))
    assert [_str(data, _ind(data, _tbl(data, d), 0)) for d in dbs] == ["db0", "other"]
    c = _tbl(data, _vec(data, _ind(data, _tbl(data, dbs[0]), 1))[0])
    assert _str(data, _ind(data, c, 0)) == "c0"
    cfg = _tbl(data, _ind(data, c, 1))
    assert struct.unpack_from("<b", data, _field(data, cfg, 1))[0] == 2
    hp = _tbl(data, _ind(data, cfg, 2))
    assert [struct.unpack_from("<i", data, _field(data, hp, i))[0] for i in range(4)] == [8, 77, 33, 9]
    assert struct.unpack_from("<q", data, _field(data, hp, 4))[0] == -5
    g = _tbl(data, _ind(data, c, 3))
    assert _str(data, _ind(data, g, 1)) == "20" and struct.unpack_from("<i", data, _field(data, g, 2))[0] == 2
    nodes = [_tbl(data, p) for p in _vec(data, _ind(data, g, 0))]
    assert [_str(data, _ind(data, n, 0)) for n in nodes] == ["7", "8", "20"]
    for i, n in enumerate(nodes):
        p = _ind(data, n, 1)
        cnt = struct.unpack_from("<I", data, p)[0]
        assert p % 4 == 0 and cnt == 4
        assert np.array_equal(np.frombuffer(data, "<f4", cnt, p + 4), vec[i])
    assert _field(data, nodes[0], 3) == 0 and struct.unpack_from("<B", data, _field(data, nodes[1], 3))[0] == 1
    # node 20: lists [[7, 8], [7], []] -> two LayerConnections (the empty list is not written), max_layer 2
    lcs = [_tbl(data, p) for p in _vec(data, _ind(data, nodes[2], 4))]
    assert len(lcs) == 2 and struct.unpack_from("<i", data, _field(data, nodes[2], 5))[0] == 2
    assert _field(data, lcs[0], 0) == 0                      # layer 0 is the default: field absent
    assert [_str(data, p) for p in _vec(data, _ind(data, lcs[0], 1))] == ["7", "8"]
    assert struct.unpack_from("<i", data, _field(data, lcs[1], 0))[0] == 1

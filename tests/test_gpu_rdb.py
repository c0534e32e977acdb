"""GPU tests of scn_store_load_rdb: a snapshot in the reference's RDB
format goes straight to device memory and answers searches exactly like the index it was taken from
(hnsw_restore_integration_test.go:70-142 asserts the same for the reference's own restore)."""
import numpy as np
import pytest

import oracle
from rdb_writer import collection_from_oracle, write_rdb
from scintirete_b200 import DeviceStore, DistanceMetric, ScintireteError
from util import gaussian

pytestmark = pytest.mark.gpu


def _built(metric, n, d, efc=100):
    db = gaussian(n, d, 1234)
    h = oracle.OracleHNSW(M=16, ef_construction=efc, ef_search=50, max_layers=16, seed=42, metric=int(metric))
    h.build(db)
    return db, h


@pytest.mark.parametrize("metric", [DistanceMetric.L2, DistanceMetric.COSINE, DistanceMetric.INNER_PRODUCT])
def test_restored_store_searches_exactly_like_the_source_index(tmp_path, metric):
    n, d, nq, k, ef = 4000, 48, 120, 10, 64
    db, h = _built(metric, n, d)
    for i in (5, 77, 1234):                                   # soft-deleted nodes travel through the snapshot
        if i != h.entrypoint():
            h.delete(i)
    coll = collection_from_oracle(h, db, int(metric))
    write_rdb(tmp_path / "dump.rdb", {"default": {"other": collection_from_oracle(*_built(metric, 50, d)[::-1], int(metric)),
                                                  "docs": coll}})
    s, info = DeviceStore.from_rdb(tmp_path / "dump.rdb", "default", "docs")
    assert (info.dim, info.metric, info.m, info.nodes, info.deleted) == (d, int(metric), 16, n, int(np.sum(coll["deleted"])))
    assert info.entry_id == h.entrypoint() and info.max_layer == h.max_layer() and info.has_graph == 1
    st = s.stats()
    assert st.rows == n and st.live_rows == h.size() and st.has_graph and st.entry_id == h.entrypoint()
    q = gaussian(nq, d, 7)
    o_ids, o_dist, o_cnt, _ = h.search_batch(q, k, ef, nthreads=8)
    ids, dist, cnt = s.search_hnsw(q, k, ef)
    assert np.array_equal(ids, o_ids) and np.array_equal(dist, o_dist) and np.array_equal(cnt, o_cnt)
    live = np.ones(n, bool)
    live[np.array(coll["ids"])[np.array(coll["deleted"])] - 1] = False
    f = oracle.flat_search(int(metric), db[live], q, k, nthreads=8)
    ids, dist, _ = s.search_flat(q, k)
    assert np.array_equal(ids, (np.nonzero(live)[0] + 1).astype(np.uint64)[f[0].astype(np.int64) - 1]) and np.array_equal(dist, f[1])
    assert np.array_equal(s.get([1, 2, n]), db[[0, 1, n - 1]])
    s.close()


def test_conversion_rules_and_errors(tmp_path):
    d = 8
    vec = gaussian(4, d, 3)
    base = {"metric": 1, "m": 4, "ids": [1, 2, 3, 9], "vectors": vec, "deleted": np.zeros(4, bool),
            "lists": [[[2, 3, 9]], [[1], [3]], [[1, 2], [2]], [[1]]], "entry": 2, "max_layer": 1}
    # a list above the node's max_layer is dropped and an unparsable neighbour id is skipped (rdb.go:1050-1061)
    quirky = dict(base, raw_layers={0: [(0, [2, "x3", 3, 9]), (5, [9])]})
    write_rdb(tmp_path / "q.rdb", {"db": {"c": quirky}})
    s, info = DeviceStore.from_rdb(tmp_path / "q.rdb", "db", "c")
    assert s.stats().graph_edges == sum(len(l) for node in base["lists"] for l in node)
    ids, dist, cnt = s.search_hnsw(vec[3], 4, 8)
    assert cnt[0] == 4 and ids[0, 0] == 9
    s.close()
    with pytest.raises(ScintireteError) as e:
        DeviceStore.from_rdb(tmp_path / "q.rdb", "nope", "c")
    assert e.value.code == 3000
    with pytest.raises(ScintireteError) as e:
        DeviceStore.from_rdb(tmp_path / "q.rdb", "db", "nope")
    assert e.value.code == 3002
    with pytest.raises(ScintireteError) as e:
        DeviceStore.from_rdb(tmp_path / "missing.rdb", "db", "c")
    assert e.value.code == 4001
    write_rdb(tmp_path / "g.rdb", {"db": {"c": dict(base, no_graph=True)}})
    with pytest.raises(ScintireteError) as e:                 # database.go:461-464
        DeviceStore.from_rdb(tmp_path / "g.rdb", "db", "c")
    assert e.value.code == 4001 and "graph state missing" in str(e.value)
    write_rdb(tmp_path / "i.rdb", {"db": {"c": dict(base, id_text={1: "12a"})}})
    with pytest.raises(ScintireteError) as e:                 # rdb.go:1038-1041
        DeviceStore.from_rdb(tmp_path / "i.rdb", "db", "c")
    assert e.value.code == 4002
    data = write_rdb(tmp_path / "t.rdb", {"db": {"c": base}})
    for cut in (3, 40, len(data) // 2, len(data) - 9):       # truncated files never crash the loader
        (tmp_path / "cut.rdb").write_bytes(data[:cut])
        with pytest.raises(ScintireteError) as e:
            DeviceStore.from_rdb(tmp_path / "cut.rdb", "db", "c")
        assert e.value.code in (4001, 4002, 3000, 3002, 3005)
    ragged = dict(base, vectors=[vec[0], vec[1][:5], vec[2], vec[3]])
    write_rdb(tmp_path / "r.rdb", {"db": {"c": ragged}})
    with pytest.raises(ScintireteError) as e:
        DeviceStore.from_rdb(tmp_path / "r.rdb", "db", "c")
    assert e.value.code == 3005

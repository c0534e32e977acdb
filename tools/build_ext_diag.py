"""GPU: upload the cached oracle graph, insert the `extra` following rows with scn_hnsw_insert, compare with the
oracle's continuation (tools/make_ext_graph.py).   python tools/build_ext_diag.py ROWS DIM METRIC EXTRA [WINDOW]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, oracle
from scintirete_b200 import DeviceStore, DistanceMetric, GraphState

n, d, metric, extra = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
window = int(sys.argv[5]) if len(sys.argv) > 5 else 0
z0 = np.load(bench.graph_cache_path(n, d, metric))
z1 = np.load(bench.graph_cache_path(n, d, metric).replace(".npz", f"_ext{extra}.npz"))
db = bench.gen_rows_numpy(0, n + extra, d, bench.SEED_DB)
s = DeviceStore(d, DistanceMetric(metric))
s.append(db[:n])
s.graph_upload(GraphState(z0["ids"], z0["list_counts"], z0["edge_counts"], z0["edges"].astype(np.uint64), int(z0["entrypoint"]),
                          int(z0["max_layer"]), n, m=16))
s.append(db[n:])
s.set_option("build_window", window)
stats = s.hnsw_insert(z1["levels"], 16, 200)
print(json.dumps(stats))
g = s.graph_export(16)
eo, eg = z1["edges"].astype(np.uint64), g.edges
ec_o, ec_g, lc = z1["edge_counts"], g.edge_counts, z1["list_counts"]
same = np.array_equal(ec_o, ec_g) and np.array_equal(eo, eg) and g.entry_point == int(z1["entrypoint"])
print("window", window, "identical:", same)
if not same:
    li = oo = og = 0
    diffs = []
    for r in range(n + extra):
        for l in range(int(lc[r])):
            a, b = eo[oo:oo + ec_o[li]], eg[og:og + ec_g[li]]
            if len(a) != len(b) or not np.array_equal(a, b):
                diffs.append((r, l, a.copy(), b.copy()))
            oo += ec_o[li]; og += ec_g[li]; li += 1
    print("lists that differ:", len(diffs), "| new nodes among them:", sorted({r + 1 for r, _, _, _ in diffs if r >= n})[:20])
    new = [x for x in diffs if x[0] >= n]
    for r, l, a, b in (new[:3] if new else diffs[:3]):
        print(f"node id {r + 1} layer {l}\n oracle: {a.tolist()}\n gpu   : {b.tolist()}")
        for v in sorted(set(a.tolist()) ^ set(b.tolist())):
            dv = oracle.distance(metric, db[r], db[int(v) - 1])
            print(f"   id {v}: d = {float(dv)!r} bits {np.float32(dv).view(np.uint32):#x} in={'oracle' if v in a else 'gpu'}")

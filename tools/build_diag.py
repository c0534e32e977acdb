"""GPU-assisted build vs the cached oracle graph: statistics, and where the graphs first differ (if they do).
    python tools/build_diag.py ROWS DIM METRIC [WINDOW]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, oracle
from scintirete_b200 import DeviceStore, DistanceMetric

n, d, metric = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
window = int(sys.argv[4]) if len(sys.argv) > 4 else 0
z = np.load(bench.graph_cache_path(n, d, metric))
db = bench.gen_rows_numpy(0, n, d, bench.SEED_DB)
s = DeviceStore(d, DistanceMetric(metric))
s.append(db)
s.set_option("build_window", window)
stats = s.hnsw_insert(z["list_counts"] - 1, 16, 200)
print(json.dumps(stats))
g = s.graph_export(16)
same = (np.array_equal(g.edge_counts, z["edge_counts"]) and np.array_equal(g.edges, z["edges"].astype(np.uint64))
        and g.entry_point == int(z["entrypoint"]) and g.max_layer == int(z["max_layer"]))
print("identical:", same, "| %.1f s, %.2f commits/round, %.2f ms/round (device %.1f s, commit %.1f s)" % (
    stats["seconds"], stats["inserted"] / max(stats["rounds"], 1), 1e3 * stats["seconds"] / max(stats["rounds"], 1),
    stats["device_seconds"], stats["commit_seconds"]))
if not same:
    lc = z["list_counts"]
    ec_o, ec_g = z["edge_counts"], g.edge_counts
    eo, eg = z["edges"].astype(np.uint64), g.edges
    li = 0
    oo = og = 0
    ndiff, first = 0, None
    for r in range(n):
        for l in range(int(lc[r])):
            a = eo[oo:oo + ec_o[li]]
            b = eg[og:og + ec_g[li]]
            if len(a) != len(b) or not np.array_equal(a, b):
                ndiff += 1
                if first is None:
                    first = (r, l, a.copy(), b.copy())
            oo += ec_o[li]
            og += ec_g[li]
            li += 1
    print("lists that differ:", ndiff)
    r, l, a, b = first
    print(f"first: node id {r + 1} layer {l}\n oracle: {a.tolist()}\n gpu   : {b.tolist()}")
    both = sorted(set(a.tolist()) ^ set(b.tolist()))
    for v in both:
        dv = oracle.distance(metric, db[r], db[int(v) - 1])
        print(f"   id {v}: distance to node = {float(dv)!r} bits {np.float32(dv).view(np.uint32):#x}  in={'oracle' if v in a else 'gpu'}")
    da = [float(oracle.distance(metric, db[r], db[int(v) - 1])) for v in a]
    print(" oracle list distances:", ["%.9g" % x for x in da])

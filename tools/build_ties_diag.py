"""Heavy exact-tie data (small integer coordinates): oracle build vs GPU-assisted build."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from scintirete_b200 import DeviceStore, DistanceMetric

for (n, d, hi, M, efc, metric) in [(3000, 8, 3, 8, 40, 1), (3000, 16, 2, 16, 64, 1), (4000, 12, 4, 16, 100, 2), (3000, 24, 3, 16, 200, 3)]:
    db = np.random.default_rng(5).integers(0, hi, (n, d)).astype(np.float32)
    if metric != 1:
        db += 1.0
    h = oracle.OracleHNSW(M=M, ef_construction=efc, ef_search=64, max_layers=16, seed=42, metric=metric)
    h.build(db)
    st = h.export_graph_state()
    for window in (1, 0):
        s = DeviceStore(d, DistanceMetric(metric))
        s.append(db)
        s.set_option("build_window", window)
        stats = s.hnsw_insert(st.list_counts - 1, M, efc)
        g = s.graph_export(M)
        same = np.array_equal(st.edge_counts, g.edge_counts) and np.array_equal(st.edges, g.edges) and st.entrypoint == g.entry_point
        print(f"n={n} d={d} values<{hi} M={M} efc={efc} metric={metric} window={window}: identical={same} rounds={stats['rounds']}")
        if not same:
            li = oo = og = 0
            nd, first = 0, None
            for r in range(n):
                for l in range(int(st.list_counts[r])):
                    a, b = st.edges[oo:oo + st.edge_counts[li]], g.edges[og:og + g.edge_counts[li]]
                    if len(a) != len(b) or not np.array_equal(a, b):
                        nd += 1
                        if first is None:
                            first = (r, l, a.copy(), b.copy())
                    oo += st.edge_counts[li]; og += g.edge_counts[li]; li += 1
            r, l, a, b = first
            print(f"  {nd} lists differ; first: node {r + 1} layer {l}\n   oracle {a.tolist()}\n   gpu    {b.tolist()}")
            print("   d(oracle):", [float(oracle.distance(metric, db[r], db[int(v) - 1])) for v in a])
            print("   d(gpu)   :", [float(oracle.distance(metric, db[r], db[int(v) - 1])) for v in b])
        s.close()

"""Heavy exact-tie data: oracle HNSW.Search vs the GPU search kernel on the same (oracle-built) graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from scintirete_b200 import DistanceMetric, GPUHNSWIndex, GraphState, HNSWParams, SearchParams

for (n, d, hi, M, efc, metric, ef) in [(3000, 8, 3, 8, 40, 1, 32), (3000, 16, 2, 16, 64, 1, 64), (4000, 12, 4, 16, 100, 2, 64), (3000, 24, 3, 16, 200, 3, 128)]:
    db = np.random.default_rng(5).integers(0, hi, (n, d)).astype(np.float32) + (0.0 if metric == 1 else 1.0)
    q = np.random.default_rng(6).integers(0, hi, (300, d)).astype(np.float32) + (0.0 if metric == 1 else 1.0)
    h = oracle.OracleHNSW(M=M, ef_construction=efc, ef_search=ef, max_layers=16, seed=42, metric=metric)
    h.build(db)
    st = h.export_graph_state()
    g = GPUHNSWIndex(HNSWParams(m=M, ef_search=ef), DistanceMetric(metric), d)
    g.import_graph_state(GraphState(st.ids, st.list_counts, st.edge_counts, st.edges, st.entrypoint, st.max_layer, st.size, st.deleted, st.vectors, m=M))
    o_ids, o_dist, o_cnt, _ = h.search_batch(q, 10, ef, nthreads=4)
    for exact in (0, 1):   # option hnsw_exact_ties: walks that end on a tie at the edge of W are redone by the exact walk kernel
        g.store.set_option("hnsw_exact_ties", exact)
        g.store.set_option("profile", 1)
        ids, dist, cnt = g.search_batch(q, SearchParams(top_k=10, ef_search=ef))
        redone = g.store.last_counters()[2]
        same_q = [bool(np.array_equal(ids[i], o_ids[i])) for i in range(len(q))]
        print(f"n={n} d={d} values<{hi} metric={metric} ef={ef} hnsw_exact_ties={exact}: ids identical for {sum(same_q)}/{len(q)} queries; "
              f"distances identical: {bool(np.array_equal(dist, o_dist))}; counts identical: {bool(np.array_equal(cnt, o_cnt))}; walks redone: {redone}")

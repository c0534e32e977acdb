"""Debug aid: compares the GPU HNSW results with the oracle's, query by query."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import oracle
from scintirete_b200 import DistanceMetric, GPUHNSWIndex, HNSWParams, SearchParams
from util import gaussian, to_graph_state

for metric, d in [(DistanceMetric.L2, 128), (DistanceMetric.COSINE, 96), (DistanceMetric.INNER_PRODUCT, 40)]:
    n, nq, k, ef = 6000, 300, 10, 64
    db = gaussian(n, d, 1234)
    h = oracle.OracleHNSW(M=16, ef_construction=200, ef_search=50, max_layers=16, seed=42, metric=int(metric))
    h.build(db)
    g = GPUHNSWIndex(HNSWParams(m=16, ef_construction=200, ef_search=50, max_layers=16, seed=42), metric, d)
    g.import_graph_state(to_graph_state(h.export_graph_state(), 16))
    q = gaussian(nq, d, 99)
    o_ids, o_dist, o_cnt, o_stats = h.search_batch(q, k, ef, nthreads=8)
    g.store.set_option("profile", 1)
    ids, dist, cnt = g.search_batch(q, SearchParams(top_k=k, ef_search=ef))
    c = g.store.last_counters()
    bad = [i for i in range(nq) if not np.array_equal(ids[i], o_ids[i])]
    print(metric, "mismatching queries", len(bad), "of", nq, "stats gpu", c[:2], "oracle", o_stats)
    for i in bad[:3]:
        print(" q", i, "\n  gpu", ids[i], dist[i], "\n  ora", o_ids[i], o_dist[i])

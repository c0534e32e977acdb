"""Extends a cached oracle graph by `extra` more inserts on the CPU oracle and caches the result (debug fixture
for the GPU-assisted build at scale). CPU only.   python tools/make_ext_graph.py ROWS DIM METRIC EXTRA"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, oracle

n, d, metric, extra = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
db = bench.gen_rows_numpy(0, n + extra, d, bench.SEED_DB)
h, _ = bench.hnsw_graph_cached(db[:n], metric, 128)
levels = np.minimum(np.floor(-np.log(np.random.default_rng(99).random(extra)) / np.log(2.0)), 15).astype(np.int32)
t = time.perf_counter()
for i in range(extra):
    h.insert(n + i + 1, db[n + i], level=int(levels[i]))
print(f"{extra} oracle inserts into the {n}x{d} graph: {time.perf_counter() - t:.0f}s")
st = h.export_graph_state(with_vectors=False)
out = bench.graph_cache_path(n, d, metric).replace(".npz", f"_ext{extra}.npz")
np.savez(out, ids=st.ids, list_counts=st.list_counts, edge_counts=st.edge_counts, edges=st.edges.astype(np.uint32),
         entrypoint=st.entrypoint, max_layer=st.max_layer, levels=levels)
print("saved", out)

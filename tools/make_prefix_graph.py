"""Builds the oracle (reference-algorithm) HNSW graph of the first n rows of a bench.py dataset and caches
it under bench_cache/ (same file layout as bench.hnsw_graph_cached). CPU only."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

n, d, metric = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
db = bench.gen_rows_numpy(0, n, d, bench.SEED_DB)
t = time.perf_counter()
h, dt = bench.hnsw_graph_cached(db, metric, 128)
print(f"{n}x{d} metric {metric}: {'built in %.0fs' % dt if dt else 'cached'}; total {time.perf_counter()-t:.0f}s")

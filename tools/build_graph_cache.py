"""Pre-builds the cached HNSW graphs bench.py uses (serial reference algorithm on the CPU)."""
import sys, time
sys.path.insert(0, "/root/repo")
import bench
for wl in sys.argv[1:]:
    rows, dim, metric, nq, k, kind = bench.WORKLOADS[wl]
    t = time.time()
    db = bench.gen_rows_numpy(0, rows, dim, bench.SEED_DB)
    h, dt = bench.hnsw_graph_cached(db, metric, 128)
    print(wl, "rows", rows, "build_s", dt, "total_s", time.time() - t, "layers", h.layers(), flush=True)

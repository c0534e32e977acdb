"""The graph nobody had: 1 000 000 x 768 cosine, M=16, efConstruction=200, built on one B200 with the reference's
serial semantics (scn_hnsw_insert), then searched (ef = 128, 10 000 queries) and checked against the CPU oracle:
the oracle imports the device-built graph, walks a query sample (ids must be identical), and performs the next
`extra` inserts on one core while the GPU does the same — the two graphs must stay identical edge for edge.
    python tools/build_1m.py [ROWS] [DIM] [METRIC] [EXTRA]   ->  gpurun_out/r02_build_<rows>_<dim>.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, oracle
from scintirete_b200 import DeviceStore, DistanceMetric

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
metric = int(sys.argv[3]) if len(sys.argv) > 3 else 2
extra = int(sys.argv[4]) if len(sys.argv) > 4 else 300
nq, k, ef = 10_000, 10, 128
out = {"workload": f"{rows}x{dim} {bench.METRIC_NAME[metric]} HNSW M=16 efC=200: GPU-assisted build, then search ef={ef}"}
t0 = time.perf_counter()
db = bench.gen_rows_numpy(0, rows + extra, dim, bench.SEED_DB)
q = bench.gen_rows_numpy(0, nq, dim, bench.SEED_Q)
levels = np.minimum(np.floor(-np.log(np.random.default_rng(42).random(rows + extra)) / np.log(2.0)), 15).astype(np.int32)
out["generate_seconds"] = time.perf_counter() - t0
s = DeviceStore(dim, DistanceMetric(metric))
for lo in range(0, rows, 100_000):
    s.append(db[lo:min(rows, lo + 100_000)])
stages, done, t_build = [], 0, 0.0
for hi in [10_000, 100_000, 250_000, 500_000, 750_000, rows]:
    hi = min(hi, rows)
    if hi <= done:
        continue
    st = s.hnsw_insert(levels[done:hi], 16, 200)
    t_build += st["seconds"]
    stages.append({"nodes": hi, "seconds": st["seconds"], "inserts_per_s": (hi - done) / st["seconds"], "rounds": st["rounds"],
                   "commits_per_round": (hi - done) / max(st["rounds"], 1), "ms_per_round": 1e3 * st["seconds"] / max(st["rounds"], 1)})
    print(json.dumps(stages[-1]), flush=True)
    done = hi
out["build"] = {"seconds": t_build, "inserts_per_s": rows / t_build, "stages": stages}
stats = s.stats()
out["graph"] = {"edges": int(stats.graph_edges), "max_layer": int(stats.max_layer), "entry_id": int(stats.entry_id)}
# ---- search on the device-built graph ----
s.set_option("profile", 1)
for _ in range(3):
    ids, dist, cnt = s.search_hnsw(q, k, ef)
s.last_timings()
t0 = time.perf_counter()
for _ in range(5):
    ids, dist, cnt = s.search_hnsw(q, k, ef)
e2e = 5 * nq / (time.perf_counter() - t0)
tm = s.last_timings()
counters = s.last_counters()
kern_ms = tm["hnsw_search"][0] / tm["hnsw_search"][1]
gt, _, _ = s.search_flat(q, k)
recall = float(np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)]))
out["search"] = {"ef": ef, "queries": nq, "kernel_ms": kern_ms, "queries_per_s_kernel": nq / (kern_ms * 1e-3), "queries_per_s_e2e_pageable": e2e,
                 "recall_at_10": recall, "evals_per_query": counters[0] / nq, "expansions_per_query": counters[1] / nq,
                 "algorithmic_GBps": (counters[0] * dim * 4 + counters[1] * 128) / (kern_ms * 1e-3) / 1e9}
print(json.dumps(out["search"]), flush=True)
# ---- the oracle on the same graph ----
g = s.graph_export(16)
h = oracle.OracleHNSW(M=16, ef_construction=200, ef_search=ef, max_layers=16, seed=42, metric=metric)
t0 = time.perf_counter()
h.import_graph_state(oracle.GraphState(g.node_ids, np.zeros(rows, np.uint8), g.list_counts, g.edge_counts, g.edges, db[:rows],
                                       g.entry_point, g.max_layer, rows))
ns = 200
t1 = time.perf_counter()
o_ids, o_dist, o_cnt, _ = h.search_batch(q[:ns], k, ef, nthreads=os.cpu_count() or 1)
t_cpu = time.perf_counter() - t1
out["oracle_search"] = {"queries": ns, "identical": bool(np.array_equal(ids[:ns], o_ids) and np.array_equal(dist[:ns], o_dist)),
                        "cpu_queries_per_s": ns / t_cpu, "cores": os.cpu_count(), "import_seconds": t1 - t0}
t1 = time.perf_counter()
for i in range(extra):
    h.insert(rows + i + 1, db[rows + i], level=int(levels[rows + i]))
cpu_s = time.perf_counter() - t1
s.append(db[rows:])
st2 = s.hnsw_insert(levels[rows:], 16, 200)
o = h.export_graph_state(with_vectors=False)
g2 = s.graph_export(16)
out["next_inserts"] = {"inserts": extra, "cpu_inserts_per_s": extra / cpu_s, "gpu_inserts_per_s": extra / st2["seconds"],
                       "identical": bool(np.array_equal(g2.edge_counts, o.edge_counts) and np.array_equal(g2.edges, o.edges)
                                         and g2.entry_point == o.entrypoint and g2.max_layer == o.max_layer),
                       "cpu_full_build_estimate_hours": rows / (extra / cpu_s) / 3600.0}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/r02_build_{rows}_{dim}.json", "w"), indent=1)

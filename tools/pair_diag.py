"""CTA-pair filter kernel vs the TMEM-stationary one: same results (and the oracle's on a sample), timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from scintirete_b200 import DeviceStore, DistanceMetric

n, d, nq, metric = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
rng = np.random.default_rng(3)
db = rng.standard_normal((n, d), dtype=np.float32)
q = rng.standard_normal((nq, d), dtype=np.float32)
s = DeviceStore(d, DistanceMetric(metric))
s.append(db)
res = {}
for pair in (0, 1):
    s.set_option("tensor_pair", pair)
    s.set_option("profile", 1)
    s.last_timings()
    ids, dist, cnt = s.search_flat(q, 10)
    for _ in range(3):
        s.search_flat(q, 10)
    t = s.last_timings()
    res[pair] = (ids, dist, cnt)
    print(f"tensor_pair={pair}: tensor_filter {t['tensor_filter'][0] / t['tensor_filter'][1]:.3f} ms/launch, counters {s.last_counters()}", flush=True)
print("pair == stationary:", np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]))
m = min(nq, 16)
o = oracle.flat_search(metric, db, q[:m], 10, nthreads=8)
print("pair == oracle (sample):", np.array_equal(res[1][0][:m], o[0]) and np.array_equal(res[1][1][:m], o[1]))

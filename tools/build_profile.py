"""A short GPU-assisted build for profiling (random level draws, no comparison): python tools/build_profile.py ROWS DIM METRIC"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from scintirete_b200 import DeviceStore, DistanceMetric

n, d, metric = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
db = bench.gen_rows_numpy(0, n, d, bench.SEED_DB)
levels = np.minimum(np.floor(-np.log(np.random.default_rng(42).random(n)) / np.log(2.0)), 15).astype(np.int32)
s = DeviceStore(d, DistanceMetric(metric))
s.append(db)
print(json.dumps(s.hnsw_insert(levels, 16, 200)))

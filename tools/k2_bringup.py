import sys, numpy as np, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_tensor import bf16_round, debug_scores
from scintirete_b200 import DeviceStore, DistanceMetric
from util import gaussian
for (n, d, nq) in [(640, 64, 128), (3000, 768, 130), (2500, 128, 70)]:
    db, q = gaussian(n, d, 1), gaussian(nq, d, 2)
    s = DeviceStore(d, DistanceMetric.INNER_PRODUCT); s.append(db)
    got = debug_scores(s, q); s.close()
    xb, qb = bf16_round(db).astype(np.float64), bf16_round(q).astype(np.float64)
    want = -(qb @ xb.T)
    err = np.abs(got - want)
    print(n, d, nq, "max err", err.max(), "scale", np.abs(want).max(), "bad frac", (err > 1e-2).mean(), flush=True)
    if err.max() > 1e-2:
        bad = np.argwhere(err > 1e-2)
        print(" first bad (q,col):", bad[:8].tolist(), "rows bad:", np.unique(bad[:,0])[:16], "cols bad:", np.unique(bad[:,1])[:32])
        print(" got", got[0,:8], "\n want", want[0,:8])

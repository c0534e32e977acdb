"""Experiment: DRAM traffic and time of tensor_filter vs. the number of row chunks per query block
(C2 shape). Run under `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum -k regex:tensor_filter`."""
import sys
sys.path.insert(0, "/root/repo")
import ctypes as C
import torch
from scintirete_b200 import DeviceStore, DistanceMetric, _native
from scintirete_b200.index import _check

rows, dim, nq, k = 1_000_000, 768, 10_000, 10
dev = torch.device("cuda", 0)
store = DeviceStore(dim, DistanceMetric.COSINE)
store.reserve(rows)
g = torch.Generator(device=dev); g.manual_seed(1234)
for r in range(0, rows, 65536):
    n = min(65536, rows - r)
    blk = torch.randn((n, dim), generator=g, device=dev)
    store.append_device(blk.data_ptr(), n)
q = torch.randn((nq, dim), generator=g, device=dev)
oi = torch.zeros((nq, k), dtype=torch.int64, device=dev); od = torch.zeros((nq, k), device=dev); oc = torch.zeros(nq, dtype=torch.int32, device=dev)
lib = _native.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for chunks in [int(x) for x in sys.argv[1:]] or [0]:
    store.set_option("tensor_chunks", chunks)
    store.set_option("profile", 0)
    for _ in range(2):
        _check(lib.scn_search_flat_dev(store.handle, C.c_void_p(q.data_ptr()), nq, k, C.c_void_p(oi.data_ptr()), C.c_void_p(od.data_ptr()), C.c_void_p(oc.data_ptr()), st))
    torch.cuda.synchronize()
    store.set_option("profile", 1); store.last_timings()
    for _ in range(3):
        _check(lib.scn_search_flat_dev(store.handle, C.c_void_p(q.data_ptr()), nq, k, C.c_void_p(oi.data_ptr()), C.c_void_p(od.data_ptr()), C.c_void_p(oc.data_ptr()), st))
    torch.cuda.synchronize()
    t = store.last_timings()
    print("chunks", chunks, {k_: round(v[0] / v[1], 3) for k_, v in t.items() if v[0] / v[1] > 0.05}, store.last_counters()[:3], flush=True)

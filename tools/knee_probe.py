"""One configuration of the exact flat search on 1 M rows, run a few times (for ncu captures of the filter kernel).

    python tools/knee_probe.py DIM NQ [option=value ...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from scintirete_b200 import DeviceStore, DistanceMetric, _native
from scintirete_b200.index import _check

dim, nq = int(sys.argv[1]), int(sys.argv[2])
rows = 1_000_000
lib = _native.lib()
dev = torch.device("cuda", 0)
store = DeviceStore(dim, DistanceMetric(2))
store.reserve(rows)
g = torch.Generator(device=dev)
g.manual_seed(1234)
for r in range(0, rows, 65536):
    n = min(65536, rows - r)
    blk = torch.randn((n, dim), generator=g, device=dev, dtype=torch.float32)
    store.append_device(blk.data_ptr(), n)
for kv in sys.argv[3:]:
    store.set_option(kv.split("=")[0], int(kv.split("=")[1]))
g.manual_seed(4321)
q = torch.randn((nq, dim), generator=g, device=dev, dtype=torch.float32)
ids = torch.zeros((nq, 10), dtype=torch.int64, device=dev)
dist = torch.zeros((nq, 10), dtype=torch.float32, device=dev)
cnt = torch.zeros((nq,), dtype=torch.int32, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(4):
    _check(lib.scn_search_flat_dev(store.handle, C.c_void_p(q.data_ptr()), nq, 10, C.c_void_p(ids.data_ptr()), C.c_void_p(dist.data_ptr()),
                                   C.c_void_p(cnt.data_ptr()), st))
torch.cuda.synchronize()
print("ok", store.last_counters())

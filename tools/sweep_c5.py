"""C5 (BASELINE.json configs[4]): query-batch sweep 1..65536 at dim 128 / 768 / 1536 on one B200,
showing the HBM-bound -> tensor-bound crossover of the exact flat search.

    python tools/sweep_c5.py [--rows 1000000] [--dims 128,768,1536] [--metric 3] [--out gpurun_out/c5.json]

For every (dim, batch) the device-resident search time is measured with CUDA events (3 warm-ups, mean over
30 / 5 / 3 back-to-back calls), the per-kernel split from the library's own events in a second pass, and converted into
  * effective HBM GB/s  = one pass over the rows actually streamed (fp32 rows for the exact scan,
    bf16 mirror for the tensor filter) / time
  * TFLOP/s             = 2*nq*N*D / time
so the knee between the two rooflines is visible. Database >> L2 at every point (>= 256 MB)."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from scintirete_b200 import DeviceStore, DistanceMetric, _native
from scintirete_b200.index import _check


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dims", default="128,768,1536")
    ap.add_argument("--metric", type=int, default=3)
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512,1024,2048,4096,8192,16384,32768,65536")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/c5.json")
    args = ap.parse_args()
    lib = _native.lib()
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {
        "hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
    results = []
    for dim in [int(x) for x in args.dims.split(",")]:
        store = DeviceStore(dim, DistanceMetric(args.metric))
        store.reserve(args.rows)
        g = torch.Generator(device=dev)
        g.manual_seed(1234)
        for r in range(0, args.rows, 65536):
            n = min(65536, args.rows - r)
            blk = torch.randn((n, dim), generator=g, device=dev, dtype=torch.float32)
            store.append_device(blk.data_ptr(), n)
        del blk
        batches = [int(x) for x in args.batches.split(",")]
        qmax = max(batches)
        g.manual_seed(4321)
        q_all = torch.randn((qmax, dim), generator=g, device=dev, dtype=torch.float32)
        out_ids = torch.zeros((qmax, args.k), dtype=torch.int64, device=dev)
        out_dist = torch.zeros((qmax, args.k), dtype=torch.float32, device=dev)
        out_cnt = torch.zeros((qmax,), dtype=torch.int32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def run(nq):
            _check(lib.scn_search_flat_dev(store.handle, C.c_void_p(q_all.data_ptr()), nq, args.k, C.c_void_p(out_ids.data_ptr()),
                                           C.c_void_p(out_dist.data_ptr()), C.c_void_p(out_cnt.data_ptr()), stream))

        for nq in batches:
            store.set_option("profile", 0)
            for _ in range(3):
                run(nq)
            torch.cuda.synchronize()
            # the call time WITHOUT the library's per-kernel events (they cost about 2 us per kernel and switch the chained
            # launches off: rounds 1 and 2a timed with them on and read about 0.06 ms high for small batches)
            reps = 30 if nq <= 1024 else (5 if nq <= 8192 else 3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                run(nq)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            store.set_option("profile", 1)   # second pass: the per-kernel split
            store.last_timings()
            preps = 3
            for _ in range(preps):
                run(nq)
            torch.cuda.synchronize()
            tim = {k_: v[0] / preps for k_, v in store.last_timings().items()}
            cnt = store.last_counters()
            tensor = cnt[0] > 0
            if tensor:
                stream_bytes = args.rows * store.stats().dim * 0 + args.rows * (((dim + 63) // 64) * 64) * 2
            else:
                stream_bytes = ((nq + 7) // 8) * args.rows * dim * 4
            rec = {"dim": dim, "nq": nq, "ms": ms, "qps": nq / (ms * 1e-3), "path": "tensor" if tensor else "exact",
                   "gbps_streamed": stream_bytes / (ms * 1e-3) / 1e9, "hbm_frac": stream_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                   "tflops": 2.0 * nq * args.rows * dim / (ms * 1e-3) / 1e12,
                   "tensor_frac": 2.0 * nq * args.rows * dim / (ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
                   "kernels_ms": tim, "counters": cnt[:3]}
            results.append(rec)
            print(json.dumps(rec), flush=True)
        store.close()
        del q_all
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"rows": args.rows, "metric": args.metric, "k": args.k, "peaks": peaks, "results": results}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

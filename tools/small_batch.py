"""Small and knee-sized batches of the exact flat search on one B200: time per call with the fused tail
(finish_queries_kernel) on and off, device-resident and through the blocking host-buffer call, the per-kernel
split, and result equality between the variants and against the fp32 exact scan.

    python tools/small_batch.py [--rows 1000000] [--dims 768,128] [--batches 1,16,128,256,512,1024,2048,4096] [--out gpurun_out/small.json]

Timing: CUDA events around `reps` back-to-back device-resident calls WITHOUT the library's per-kernel events
(they cost about 2 us per kernel); the kernel split comes from a second, profiled pass. Host-buffer latency: wall
clock around scn_search_flat from pinned memory, one call at a time."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from scintirete_b200 import DeviceStore, DistanceMetric, _native
from scintirete_b200.index import _check


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dims", default="768,128")
    ap.add_argument("--metric", type=int, default=2)
    ap.add_argument("--batches", default="1,16,128,256,512,1024,2048,4096")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--variants", default="tensor_fused=0;tensor_fused=1")
    ap.add_argument("--out", default="gpurun_out/small.json")
    args = ap.parse_args()
    lib = _native.lib()
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    variants = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in v.split(",") if kv) for v in args.variants.split(";")]
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
        q_host = q_all.cpu().numpy()
        out_ids = torch.zeros((qmax, args.k), dtype=torch.int64, device=dev)
        out_dist = torch.zeros((qmax, args.k), dtype=torch.float32, device=dev)
        out_cnt = torch.zeros((qmax,), dtype=torch.int32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        mirror_bytes = args.rows * (((dim + 63) // 64) * 64) * 2

        def run(nq):
            _check(lib.scn_search_flat_dev(store.handle, C.c_void_p(q_all.data_ptr()), nq, args.k, C.c_void_p(out_ids.data_ptr()),
                                           C.c_void_p(out_dist.data_ptr()), C.c_void_p(out_cnt.data_ptr()), stream))

        # reference results of the first 16 queries from the fp32 exact scan
        store.set_option("flat_path", 1)
        run(min(16, qmax))
        torch.cuda.synchronize()
        ref_ids, ref_dist = out_ids[:16].cpu().numpy().copy(), out_dist[:16].cpu().numpy().copy()
        store.set_option("flat_path", 0)

        for nq in batches:
            base = None
            for var in variants:
                for name, val in var.items():
                    store.set_option(name, val)
                store.set_option("profile", 0)
                for _ in range(3):
                    run(nq)
                torch.cuda.synchronize()
                reps = 50 if nq <= 1024 else 10
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    run(nq)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                ids, dist = out_ids[:nq].cpu().numpy().copy(), out_dist[:nq].cpu().numpy().copy()
                m = min(nq, 16)
                same_exact = bool(np.array_equal(ids[:m], ref_ids[:m]) and np.array_equal(dist[:m].view(np.uint32), ref_dist[:m].view(np.uint32)))
                if base is None:
                    base = (ids, dist)
                same_base = bool(np.array_equal(ids, base[0]) and np.array_equal(dist.view(np.uint32), base[1].view(np.uint32)))
                # blocking host-buffer call, one at a time (a caller's latency)
                t = []
                for _ in range(20 if nq <= 1024 else 5):
                    t0 = time.perf_counter()
                    store.search_flat(q_host[:nq], args.k)
                    t.append(time.perf_counter() - t0)
                host_ms = float(np.median(t[2:]) * 1e3)
                store.set_option("profile", 1)
                store.last_timings()
                for _ in range(5):
                    run(nq)
                torch.cuda.synchronize()
                tim = {k_: round(v[0] / 5, 4) for k_, v in store.last_timings().items()}
                cnt = store.last_counters()
                store.set_option("profile", 0)
                rec = {"dim": dim, "nq": nq, "variant": var, "ms": ms, "host_call_ms": host_ms, "qps": nq / (ms * 1e-3),
                       "hbm_frac": mirror_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                       "tensor_frac": 2.0 * nq * args.rows * dim / (ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
                       "identical_to_exact_scan_first16": same_exact, "identical_to_first_variant": same_base, "kernels_ms": tim, "counters": cnt[:3]}
                results.append(rec)
                print(json.dumps(rec), flush=True)
        store.close()
        del q_all
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"rows": args.rows, "metric": args.metric, "k": args.k, "peaks": peaks, "results": results}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

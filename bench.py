#!/usr/bin/env python
"""bench.py — queries/sec of the search hot path on B200 (see BASELINE.json / SURVEY.md §8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|hnsw|...]
    torchrun --nproc-per-node N bench.py --gpus N ...        (N > 1: one rank per GPU, NCCL)

A "step" is one pass of the hot path over one batch of synthetic queries. Default workload (C2):
1M x 768 fp32 Gaussian rows, cosine, exact k=10 search of a 10 000-query batch on one B200. With
N > 1 the same database is row-sharded over the ranks (strong scaling); every rank scans its shard
for the whole batch, the per-shard top-k lists are all-gathered over NCCL and merged by
scn_merge_topk_dev.

Printed JSON line: `value` = queries/s with the queries already resident in HBM (CUDA events on
the launching stream, max over ranks); `e2e` = queries/s through the blocking host-buffer C-ABI call
(pinned host queries in, ids/distances out, copies inside the timed region); `roofline` for the
dominant kernel from live CUDA-event timings; `cpu_baseline` = the CPU oracle (restatement of the
reference's Go code; "port") timed on this box's host cores on a bounded query sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

WORKLOADS = {
    # name: (rows, dim, metric, nq, k, kind)
    "c2": (1_000_000, 768, 2, 10_000, 10, "flat"),     # BASELINE.json configs[1] — the headline
    "c4": (10_000_000, 768, 3, 10_000, 10, "flat"),    # configs[3]: 10M x 768 IP, row-sharded
    "c2-small": (100_000, 768, 2, 2_000, 10, "flat"),
    "c1": (100_000, 128, 1, 1_000, 10, "hnsw"),        # configs[0]
    "c3": (1_000_000, 128, 1, 10_000, 10, "hnsw"),     # configs[2]
    "c1-768": (100_000, 768, 2, 10_000, 10, "hnsw"),   # embedding-sized rows (cosine): the two-stage row gather
}
METRIC_NAME = {1: "L2", 2: "cosine", 3: "inner_product"}
SEED_DB, SEED_Q = 1234, 4321


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self._nv = self._h = None
        try:  # NVML is initialised here, outside the timed region
            import pynvml as nv

            nv.nvmlInit()
            self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def _sample(self):
        nv, h = self._nv, self._h
        self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for bit, name in ((nv.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"),
                          (nv.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                          (nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                          (nv.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap")):
            if r & bit:
                self.reasons.add(name)

    def run(self):
        if self._nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self._sample()
            except Exception as e:
                self.reasons.add(f"nvml_error:{type(e).__name__}")
                return
            time.sleep(0.004)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def gen_rows_numpy(row0: int, n: int, dim: int, seed: int) -> np.ndarray:
    """Host generator (reference arm / CPU baseline): block-seeded so any row range is reproducible."""
    out = np.empty((n, dim), np.float32)
    blk = 65536
    r = row0
    while r < row0 + n:
        b = r // blk
        lo, hi = max(r, b * blk), min(row0 + n, (b + 1) * blk)
        block = np.random.default_rng([seed, b]).standard_normal((blk, dim), dtype=np.float32)
        out[lo - row0:hi - row0] = block[lo - b * blk:hi - b * blk]
        r = hi
    return out


def hnsw_graph_cached(db: np.ndarray, metric: int, ef_search: int, M: int = 16, efc: int = 200, seed: int = 42):
    """The graph the reference's algorithm builds (serial CPU construction, hnsw.go:148-257, restated
    by the oracle). Built once per (data, parameters) and cached under bench_cache/ because the
    serial build takes minutes (100k x 128) to an hour (1M x 128) on one core; the cache travels to
    the GPU box with the repo snapshot. Returns (oracle index, build seconds or None if cached)."""
    import oracle

    n, dim = db.shape
    tag = f"hnsw_n{n}_d{dim}_m{metric}_M{M}_efc{efc}_s{seed}_db{SEED_DB}"
    path = os.path.join(ROOT, "bench_cache", tag + ".npz")
    h = oracle.OracleHNSW(M=M, ef_construction=efc, ef_search=ef_search, max_layers=16, seed=seed, metric=metric)
    if os.path.exists(path):
        z = np.load(path)
        st = oracle.GraphState(z["ids"], z["deleted"], z["list_counts"], z["edge_counts"], z["edges"].astype(np.uint64),
                               db, int(z["entrypoint"]), int(z["max_layer"]), int(z["size"]))
        h.import_graph_state(st)
        return h, None
    t0 = time.perf_counter()
    h.build(db)
    dt = time.perf_counter() - t0
    st = h.export_graph_state(with_vectors=False)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez(path, ids=st.ids, deleted=st.deleted, list_counts=st.list_counts, edge_counts=st.edge_counts,
             edges=st.edges.astype(np.uint32), entrypoint=st.entrypoint, max_layer=st.max_layer, size=st.size)
    return h, dt


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port; Go cannot run here) on host cores
# ------------------------------------------------------------------------------------------------

def cpu_flat_qps(db: np.ndarray, queries: np.ndarray, metric: int, k: int, threads: int, rounds: int = 1):
    import oracle

    t0 = time.perf_counter()
    for _ in range(rounds):
        oracle.flat_search(metric, db, queries, k, nthreads=threads)
    dt = time.perf_counter() - t0
    return rounds * len(queries) / dt, dt


def run_reference(args, wl):
    rows, dim, metric, nq, k, kind = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t_gen = time.perf_counter()
    if kind == "flat":
        db = gen_rows_numpy(0, rows, dim, SEED_DB)
        q = gen_rows_numpy(0, nq, dim, SEED_Q)
        t_gen = time.perf_counter() - t_gen
        # bounded sample per step: one query per host thread (goroutine-per-request under RLock)
        per_step = min(nq, threads)
        for _ in range(args.warmup):
            cpu_flat_qps(db, q[:per_step], metric, k, threads)
        t0 = time.perf_counter()
        for s in range(args.steps):
            lo = (s * per_step) % max(1, nq - per_step + 1)
            cpu_flat_qps(db, q[lo:lo + per_step], metric, k, threads)
        dt = time.perf_counter() - t0
        qps = args.steps * per_step / dt
        sample = f"{per_step} queries/step x {args.steps} steps over the full {rows}x{dim} database, one query per thread"
    else:
        import oracle

        db = gen_rows_numpy(0, rows, dim, SEED_DB)
        q = gen_rows_numpy(0, nq, dim, SEED_Q)
        h, _ = hnsw_graph_cached(db, metric, args.ef)
        t_gen = time.perf_counter() - t_gen
        per_step = min(nq, 10000)
        for _ in range(args.warmup):
            h.search_batch(q[:per_step], k, args.ef, nthreads=threads)
        t0 = time.perf_counter()
        for s in range(args.steps):
            h.search_batch(q[:per_step], k, args.ef, nthreads=threads)
        dt = time.perf_counter() - t0
        qps = args.steps * per_step / dt
        sample = f"{per_step} queries/step x {args.steps} steps, HNSW M=16 efC=200 ef={args.ef}"
    line = {
        "impl": "reference", "metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic gaussian (numpy, seeded)",
        "config": {"workload": f"{args.workload}: {rows}x{dim} {METRIC_NAME[metric]} {kind} k={k}", "rows": rows,
                   "dim": dim, "metric": METRIC_NAME[metric], "nq_per_step": per_step, "k": k},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "reference algorithm restated in C++ (Go toolchain unavailable); setup %.1fs" % t_gen},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    from scintirete_b200 import DeviceStore, DistanceMetric, _native
    from scintirete_b200.index import _check

    rows, dim, metric, nq, k, kind = wl
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _native.lib()

    store = None
    if kind == "flat":
        # ---- row shard of the database, generated on the device ----------------------------------
        per = (rows + world - 1) // world
        row0, row1 = rank * per, min(rows, (rank + 1) * per)
        n_local = row1 - row0
        store = DeviceStore(dim, DistanceMetric(metric), device=local)
        store.reserve(n_local)
        blk = 65536
        r = row0
        while r < row1:
            b = r // blk
            lo, hi = max(r, b * blk), min(row1, (b + 1) * blk)
            g = torch.Generator(device=dev)
            g.manual_seed(SEED_DB * 1_000_003 + b)
            block = torch.randn((blk, dim), generator=g, device=dev, dtype=torch.float32)
            chunk = block[lo - b * blk:hi - b * blk].contiguous()
            store.append_device(chunk.data_ptr(), hi - lo)
            r = hi
        g = torch.Generator(device=dev)
        g.manual_seed(SEED_Q)
        q_dev = torch.randn((nq, dim), generator=g, device=dev, dtype=torch.float32)
        q_host = q_dev.cpu().pin_memory()
    else:
        # ---- HNSW: replicas only. Every rank holds the whole store + graph and answers its slice
        # of the query batch; no data-path collective. Host-generated data (numpy, seeded) so the
        # cached graph (bench_cache/) matches the vectors bit for bit on any machine. The graph is
        # the one the reference's algorithm builds (hnsw.go:148-257, restated by the oracle) and
        # is constructed / loaded outside every timed region.
        from scintirete_b200 import GraphState

        row0, n_local = 0, rows
        db_host = gen_rows_numpy(0, rows, dim, SEED_DB)
        q_all = gen_rows_numpy(0, nq, dim, SEED_Q)
        qlo, qhi = (nq * rank) // world, (nq * (rank + 1)) // world
        nq_total, nq = nq, qhi - qlo
        h, build_s = hnsw_graph_cached(db_host, metric, args.ef)
        store = DeviceStore(dim, DistanceMetric(metric), device=local)
        store.append(db_host)
        st = h.export_graph_state(with_vectors=False)
        store.graph_upload(GraphState(st.ids, st.list_counts, st.edge_counts, st.edges, st.entrypoint, st.max_layer,
                                      st.size, m=16))
        q_host = torch.from_numpy(q_all[qlo:qhi].copy()).pin_memory()
        q_dev = q_host.to(dev)
    torch.cuda.synchronize()

    out_ids = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    out_dist = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    out_cnt = torch.zeros((nq,), dtype=torch.int32, device=dev)
    keys = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    if world > 1:
        all_keys = torch.zeros((world, nq, k), dtype=torch.int64, device=dev)
        all_ids = torch.zeros((world, nq, k), dtype=torch.int64, device=dev)
    h_ids = torch.zeros((nq, k), dtype=torch.int64).pin_memory()
    h_dist = torch.zeros((nq, k), dtype=torch.float32).pin_memory()
    h_cnt = torch.zeros((nq,), dtype=torch.int32).pin_memory()

    # Row-sharded exchange: fused into the search epilogue over NVLink peer memory (P2P stores +
    # flags, scn_search_flat_exchange_dev), or the NCCL formulation (2 all-gathers + merge).
    merge_mode, exchange = "none", None
    if world > 1 and kind == "flat":
        merge_mode = args.merge
        if merge_mode == "p2p":
            from scintirete_b200.sharding import ShardExchange

            handle = None
            try:
                exchange = ShardExchange(local, rank, world, nq, k)
                handle = exchange.local_handle()
            except Exception as e:
                print(f"[rank {rank}] peer-memory exchange unavailable ({e})", file=sys.stderr)
            handles = [None] * world
            dist.all_gather_object(handles, handle)          # every rank takes part, whatever happened above
            ok = torch.ones(1, device=dev)
            try:
                if any(h is None for h in handles):
                    raise RuntimeError("a rank could not create its exchange buffer")
                exchange.connect(handles)
            except Exception as e:  # e.g. CUDA IPC not permitted in this container
                print(f"[rank {rank}] peer-memory exchange unavailable ({e}); using NCCL all-gather", file=sys.stderr)
                ok = torch.zeros(1, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)        # all ranks must agree on the protocol
            if ok.item() == 0:
                exchange, merge_mode = None, "nccl"

    def p(t):
        return C.c_void_p(t.data_ptr())

    def step_device(qd):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if kind == "hnsw":
            _check(lib.scn_search_hnsw_dev(store.handle, p(qd), nq, k, args.ef, p(out_ids), p(out_dist), p(out_cnt), stream))
        elif world == 1 and kind == "flat":
            _check(lib.scn_search_flat_dev(store.handle, p(qd), nq, k, p(out_ids), p(out_dist), p(out_cnt), stream))
        elif exchange is not None:
            exchange.search(store, qd.data_ptr(), nq, row0, out_ids.data_ptr(), out_dist.data_ptr(), out_cnt.data_ptr(),
                            torch.cuda.current_stream().cuda_stream)
        else:
            _check(lib.scn_search_flat_shard_dev(store.handle, p(qd), nq, k, row0, p(keys), p(out_ids), stream))
            dist.all_gather_into_tensor(all_keys, keys)
            dist.all_gather_into_tensor(all_ids, out_ids)
            _check(lib.scn_merge_topk_dev(local, p(all_keys), p(all_ids), world, nq, k, p(out_ids), p(out_dist),
                                          p(out_cnt), stream))

    def step_e2e():
        # the call a user of the C ABI makes: host buffers in, host buffers out
        if world == 1 and kind == "flat":
            _check(lib.scn_search_flat(store.handle, p(q_host), nq, k, p(h_ids), p(h_dist), p(h_cnt)))
        elif kind == "hnsw":
            _check(lib.scn_search_hnsw(store.handle, p(q_host), nq, k, args.ef, p(h_ids), p(h_dist), p(h_cnt)))
        else:
            q_dev.copy_(q_host, non_blocking=True)
            step_device(q_dev)
            h_ids.copy_(out_ids, non_blocking=True)
            h_dist.copy_(out_dist, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for kv in args.opt:
        name, val = kv.split("=")
        store.set_option(name, int(val))

    # ---- timed region 1: device-resident queries -------------------------------------------------
    store.set_option("profile", 0)
    for _ in range(args.warmup):
        step_device(q_dev)
    store.set_option("profile", 1)
    store.last_timings()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.scn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device(q_dev)
    e1.record()
    barrier()
    launches = lib.scn_launch_count() - launches0
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    timings = store.last_timings()
    counters = store.last_counters()
    store.set_option("profile", 0)

    # ---- timed region 2: end to end through the host-buffer call ---------------------------------
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())

    if rank == 0:
        pk = peaks()
        ms_step = ms_total / args.steps
        nq_job = nq_total if kind == "hnsw" else nq   # replicas split the batch; shards share it
        value = nq_job / (ms_step * 1e-3)
        # dominant kernel + its roofline
        roof = None
        if timings:
            top = max(timings.items(), key=lambda kv: kv[1][0])
            name, (tot_ms, cnt) = top
            avg_ms = tot_ms / max(cnt, 1)
            share = tot_ms / max(sum(v[0] for v in timings.values()), 1e-9)
            if name == "tensor_filter":
                flops = 2.0 * nq * n_local * dim
                ach = flops / (avg_ms * 1e-3) / 1e12
                # DRAM bytes per launch from the committed ncu --set full capture of this exact workload
                traffic = 2.387e9 if (rows, dim, nq, world) == (1_000_000, 768, 10_000, 1) else None
                roof = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                        "frac": ach / pk["bf16_sustained"], "traffic": traffic,
                        "traffic_source": "profiles/r01_ncu_tensor_filter_metrics.json (dram__bytes_read+write, bytes/launch)" if traffic else None,
                        "algorithmic_flop": flops, "peak_source": pk["source"] + " (sustained bf16)",
                        "launch_ms": avg_ms, "share_of_step": share}
            elif name == "flat_exact_scan":
                passes = (nq + 7) // 8
                byts = float(passes) * n_local * (dim * 4 + (4 if metric == 2 else 0))
                ach = byts / (avg_ms * 1e-3) / 1e9
                roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                        "frac": ach / pk["hbm"], "traffic": None, "peak_source": pk["source"], "launch_ms": avg_ms,
                        "share_of_step": share, "note": f"{passes} passes of 8 queries over fp32 rows per launch"}
            elif name.startswith("hnsw_search") and counters:
                evals, hops = counters[0], counters[1]
                byts = evals * dim * 4.0 + hops * 32 * 4.0
                ach = byts / (avg_ms * 1e-3) / 1e9
                # DRAM bytes per launch from the committed ncu --set full capture of this exact workload: BELOW the
                # algorithmic bytes, because the walks of a batch share hub rows and L2 serves them
                traffic = 8.701e9 if (rows, dim, nq, world, args.ef) == (1_000_000, 128, 10_000, 1, 128) else None
                roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                        "frac": ach / pk["hbm"], "traffic": traffic,
                        "traffic_source": "profiles/r01_ncu_hnsw_search_metrics.json (dram__bytes_read+write, bytes/launch)" if traffic else None,
                        "algorithmic_bytes": byts, "peak_source": pk["source"], "launch_ms": avg_ms,
                        "share_of_step": share, "note": f"{evals / nq:.0f} distance evals, {hops / nq:.0f} expansions per query (counted on device); "
                                                        "algorithmic bytes = evals*dim*4 + expansions*2M*4"}
        # CPU baseline on a bounded sample (rank 0, N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            if kind == "flat":
                db_host = np.concatenate([store.get(np.arange(i, min(i + 65536, n_local), dtype=np.uint64) + 1)
                                          for i in range(0, n_local, 65536)])
                ns = min(nq, threads)
                rounds = 16 if rows * dim >= 5e8 else 32   # about 10 s of host work at C2 (0.6 s per round)
                qps, dt = cpu_flat_qps(db_host, q_host.numpy()[:ns], metric, k, threads, rounds)
                cpu = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                       "sample": f"{ns * rounds} queries ({rounds} rounds x {ns}, one per thread) over the full database, {dt:.1f}s wall"}
                del db_host
            else:
                # the whole batch, repeated until about 10 s of host work have been timed
                ns, reps = nq, 0
                t0 = time.perf_counter()
                while reps < 64 and time.perf_counter() - t0 < 10.0:
                    h.search_batch(q_host.numpy()[:ns], k, args.ef, nthreads=threads)
                    reps += 1
                dt = time.perf_counter() - t0
                cpu = {"value": ns * reps / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                       "sample": f"{reps} x {ns} queries, ef={args.ef}, same graph, {dt:.1f}s wall; graph "
                                 + (f"built in {build_s:.0f}s (1 thread)" if build_s else "loaded from bench_cache/")}
        line = {
            "metric": "queries/sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16 filter + f32 exact rerank" if (timings and "tensor_filter" in timings) else "f32",
            "data": "synthetic gaussian (torch.randn on device, seeded)",
            "config": {"workload": f"{args.workload}: {rows}x{dim} {METRIC_NAME[metric]} {kind} k={k} nq={nq}"
                                   + (f" ef={args.ef}" if kind == "hnsw" else ""),
                       "rows": rows, "dim": dim, "metric": METRIC_NAME[metric], "nq": nq, "k": k,
                       "sharding": f"rows/{world}" if kind == "flat" else f"replicas x{world}, query batch split",
                       "shard_merge": {"p2p": "fused into the search epilogue over NVLink peer memory (P2P stores + flags)",
                                       "nccl": "NCCL all_gather of keys and ids + merge kernel", "none": None}[merge_mode],
                       "l2_policy": "database (>= 3 GB fp32 + bf16 mirror per pass) is far larger than the 126 MB L2; no flush needed"
                       if rows * dim * 4 > 4e8 else "working set fits L2: flush not applied (small workload, not the headline)"},
            "e2e": {"value": nq_job * args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4,
                    "d2h_bytes_per_step": nq * k * 12 + nq * 4, "timer": "host wall clock around the blocking C-ABI call"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "kernels_ms_per_step": {n_: v[0] / args.steps for n_, v in timings.items()},
            "counters": counters,
        }
        emit(line)
    if exchange is not None:
        exchange.status(torch.cuda.current_stream().cuda_stream)   # a missed peer arrival would have invalidated the run
        barrier()
        exchange.close()
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def claim_stdout():
    """stdout carries the ONE JSON line and nothing else: file descriptor 1 is pointed at stderr for
    everything that writes to it behind Python's back (NCCL prints its version banner there), and
    the line itself goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int)
    ap.add_argument("--dim", type=int)
    ap.add_argument("--nq", type=int)
    ap.add_argument("--ef", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (scn_set_option), repeatable")
    ap.add_argument("--merge", default="p2p", choices=["p2p", "nccl"], help="row-shard exchange for --gpus > 1")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.rows:
        wl[0] = args.rows
    if args.dim:
        wl[1] = args.dim
    if args.nq:
        wl[3] = args.nq
    if args.workload == "c1" and args.ef == 128:
        args.ef = 100
    if args.impl == "reference":
        run_reference(args, tuple(wl))
    else:
        run_ours(args, tuple(wl))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — queries/sec of the search hot path on B200 (see BASELINE.json / SURVEY.md §8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|c3|...]
    torchrun --nproc-per-node N bench.py --gpus N ...        (N > 1: one rank per GPU)

A "step" is one pass of the hot path over one batch of synthetic queries. Default workload (C2):
1M x 768 fp32 Gaussian rows, cosine, exact k=10 search of a 10 000-query batch. With N > 1 the same
database is row-sharded over the ranks (strong scaling); a batch is cut into N query slices, every
rank scans its rows for the whole batch and answers its slice (fused exchange over NVLink peer
memory, csrc/exchange.cu; `--merge nccl` = the NCCL all-gather formulation for comparison).

Printed JSON line: `value` = queries/s with the queries already resident in HBM (CUDA events on
the launching stream, max over ranks); `e2e` = queries/s through the blocking host-buffer C-ABI call
(pinned host buffers from scn_host_alloc; copies inside the timed region; `pageable` = the same call
from ordinary pageable memory); `roofline` for the dominant kernel from live CUDA-event timings;
`cpu_baseline` = the CPU oracle (restatement of the reference's Go code; "port") timed on this box's
host cores on a bounded query sample; `verified` = a query sample of the results compared with the
oracle over the same rows (ids and distance bits). `secondary` carries the other BASELINE.json
configurations measured in the same run: C3 (HNSW, N = 1; recall and an ef sweep), C4 (N >= 2).
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


def traffic_for(kernel: str, rows: int, dim: int, nq: int, world: int, ef=None):
    """DRAM bytes per launch of `kernel` on exactly this configuration, from the committed
    `ncu --set full` captures (profiles/traffic.json: one entry per capture, with its source file).
    None when this configuration was never captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    key = f"{kernel}|rows={rows}|dim={dim}|nq={nq}|world={world}" + (f"|ef={ef}" if ef is not None else "")
    e = json.load(open(p)).get(key)
    return (e["bytes"], e["source"]) if e else (None, None)


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
    """Host generator (reference arm / HNSW workloads): block-seeded so any row range is reproducible."""
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


def graph_cache_path(n, dim, metric, M=16, efc=200, seed=42):
    return os.path.join(ROOT, "bench_cache", f"hnsw_n{n}_d{dim}_m{metric}_M{M}_efc{efc}_s{seed}_db{SEED_DB}.npz")


def hnsw_graph_cached(db: np.ndarray, metric: int, ef_search: int, M: int = 16, efc: int = 200, seed: int = 42):
    """The graph the reference's algorithm builds (serial CPU construction, hnsw.go:148-257, restated
    by the oracle). Built once per (data, parameters) and cached under bench_cache/ because the
    serial build takes minutes (100k x 128) to hours (1M x 128) on one core; the cache travels to
    the GPU box with the repo snapshot. Returns (oracle index, build seconds or None if cached)."""
    import oracle

    n, dim = db.shape
    path = graph_cache_path(n, dim, metric, M, efc, seed)
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

class Ctx:
    """Process-wide pieces shared by the workloads of one run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        from scintirete_b200 import _native

        self.args, self.torch, self.dist = args, torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torchrun --nproc-per-node {args.gpus}")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.lib = _native.lib()
        self.peaks = peaks()
        self.threads = os.cpu_count() or 1

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


def time_device(ctx: Ctx, step, steps: int, warmup: int, store):
    """W warm-ups, then K steps bracketed by barrier + synchronize, CUDA events on the launching
    stream, max over ranks. Per-kernel event timings are collected during the timed steps."""
    torch = ctx.torch
    store.set_option("profile", 0)
    for _ in range(warmup):
        step()
    store.set_option("profile", 1)
    store.last_timings()
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    sampler.start()
    launches0 = ctx.lib.scn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    ctx.barrier()
    launches = ctx.lib.scn_launch_count() - launches0
    clocks = sampler.stop()
    ms_total = ctx.max_over_ranks(e0.elapsed_time(e1))
    timings = store.last_timings()
    counters = store.last_counters()
    store.set_option("profile", 0)
    return ms_total, timings, counters, int(launches), clocks


def time_host(ctx: Ctx, step, steps: int, warmup: int) -> float:
    """Wall clock around blocking host-buffer calls, max over ranks (seconds for `steps` calls)."""
    for _ in range(max(1, warmup)):
        step()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    ctx.torch.cuda.synchronize()
    return ctx.max_over_ranks(time.perf_counter() - t0)


def pick_peak(pk, clocks, kind):
    """Tensor roofline denominator chosen by the clock record of the timed region: the burst figure
    when the SMs held (nearly) their maximum clock and no power cap was reported, else the sustained one."""
    if kind == "hbm":
        return pk["hbm"], pk["source"]
    capped = "sw_power_cap" in (clocks.get("reasons") or [])
    held = clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"]
    if held and not capped:
        return pk["bf16_burst"], pk["source"] + " (burst bf16: clocks held, no power cap)"
    return pk["bf16_sustained"], pk["source"] + " (sustained bf16: power-capped or clocks below max)"


def roofline_of(ctx: Ctx, timings, counters, clocks, *, rows, n_local, dim, metric, nq, world, ef=None):
    if not timings:
        return None
    pk = ctx.peaks
    name, (tot_ms, cnt) = max(timings.items(), key=lambda kv: kv[1][0])
    avg_ms = tot_ms / max(cnt, 1)
    share = tot_ms / max(sum(v[0] for v in timings.values()), 1e-9)
    if name == "tensor_filter":
        flops = 2.0 * nq * n_local * dim
        ach = flops / (avg_ms * 1e-3) / 1e12
        peak, src = pick_peak(pk, clocks, "tensor")
        traffic, tsrc = traffic_for(name, rows, dim, nq, world)
        return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "frac_sustained": ach / pk["bf16_sustained"], "frac_burst": ach / pk["bf16_burst"], "traffic": traffic,
                "traffic_source": tsrc, "algorithmic_flop": flops, "peak_source": src, "launch_ms": avg_ms, "share_of_step": share}
    if name == "flat_exact_scan":
        passes = (nq + 7) // 8
        byts = float(passes) * n_local * (dim * 4 + (4 if metric == 2 else 0))
        ach = byts / (avg_ms * 1e-3) / 1e9
        traffic, tsrc = traffic_for(name, rows, dim, nq, world)
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                "traffic": traffic, "traffic_source": tsrc, "peak_source": pk["source"], "launch_ms": avg_ms,
                "share_of_step": share, "note": f"{passes} passes of 8 queries over fp32 rows per launch"}
    if name.startswith("hnsw_search") and counters:
        evals, hops = counters[0], counters[1]
        byts = evals * dim * 4.0 + hops * 32 * 4.0
        ach = byts / (avg_ms * 1e-3) / 1e9
        # DRAM bytes per launch (ncu --set full of this exact workload): the walks of a batch share hub rows
        # that L2 serves, and the visited tables add traffic the algorithm does not count — report both
        traffic, tsrc = traffic_for("hnsw_search", rows, dim, nq, world, ef)
        r = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
             "traffic": traffic, "traffic_source": tsrc, "algorithmic_bytes": byts, "peak_source": pk["source"],
             "launch_ms": avg_ms, "share_of_step": share,
             "note": f"{evals / max(nq, 1):.0f} distance evals, {hops / max(nq, 1):.0f} expansions per query (counted on device); "
                     "algorithmic bytes = evals*dim*4 + expansions*2M*4; the walk is latency-bound (DESIGN.md 4.2)"}
        if traffic:
            r["achieved_dram"] = traffic / (avg_ms * 1e-3) / 1e9
            r["frac_dram"] = r["achieved_dram"] / pk["hbm"]
        return r
    return None


def oracle_flat_sharded(ctx: Ctx, store, n_local: int, row0: int, qs: np.ndarray, metric: int, k: int):
    """Exact top-k of the sample queries over the WHOLE database by the CPU oracle: every rank scans
    the rows its own GPU holds (read back block by block), the per-shard lists are merged by
    (distance, global row) — the flat oracle's order — on every rank. Returns (ids, dist)."""
    import oracle

    threads = max(1, ctx.threads // ctx.world)
    cand_i, cand_d = [], []
    blk = 131072
    for lo in range(0, n_local, blk):
        hi = min(n_local, lo + blk)
        gids = np.arange(row0 + lo + 1, row0 + hi + 1, dtype=np.uint64)
        rows = store.get(gids)
        i_, d_, c_ = oracle.flat_search(metric, rows, qs, k, ids=gids, nthreads=threads)
        cand_i.append(i_)
        cand_d.append(d_)
    if cand_i:
        li, ld = np.concatenate(cand_i, axis=1), np.concatenate(cand_d, axis=1)
    else:
        li, ld = np.zeros((len(qs), 0), np.uint64), np.zeros((len(qs), 0), np.float32)
    parts = ctx.gather_objects((li, ld))
    ai = np.concatenate([p[0] for p in parts], axis=1)
    ad = np.concatenate([p[1] for p in parts], axis=1)
    out_i = np.zeros((len(qs), k), np.uint64)
    out_d = np.full((len(qs), k), np.inf, np.float32)
    for q in range(len(qs)):
        valid = ai[q] != 0
        vi, vd = ai[q][valid], ad[q][valid]
        order = np.lexsort((vi, vd))[:k]            # distance ascending, then id (= global row + 1) ascending
        out_i[q, :len(order)] = vi[order]
        out_d[q, :len(order)] = vd[order]
    return out_i, out_d


def bench_flat(ctx: Ctx, wl, name: str, steps: int, warmup: int, cpu_baseline: bool):
    torch, lib, args = ctx.torch, ctx.lib, ctx.args
    from scintirete_b200 import DeviceStore, DistanceMetric, PinnedBuffer
    from scintirete_b200.index import _check

    rows, dim, metric, nq, k, _ = wl
    world, rank, local, dev = ctx.world, ctx.rank, ctx.local, ctx.dev
    # ---- row shard of the database, generated on the device (global rows row0..row1) ---------------
    per = (rows + world - 1) // world
    row0, row1 = min(rows, rank * per), min(rows, (rank + 1) * per)
    n_local = row1 - row0
    store = DeviceStore(dim, DistanceMetric(metric), device=local)
    store.set_option("auto_id_base", row0)          # ids are global row + 1 on every shard
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
    q_dev = torch.randn((nq, dim), generator=g, device=dev, dtype=torch.float32)   # same batch on every rank
    q_np = q_dev.cpu().numpy()
    for kv in args.opt:
        oname, val = kv.split("=")
        store.set_option(oname, int(val))
    torch.cuda.synchronize()

    def p(t):
        return C.c_void_p(t.data_ptr())

    # ---- query slice this rank answers; pinned host buffers for the C-ABI calls -----------------------
    merge_mode, exchange = "none", None
    qlo, qcnt = 0, nq
    if world > 1:
        merge_mode = args.merge
        from scintirete_b200.sharding import ShardExchange, query_slice

        qlo, qhi = query_slice(nq, world, rank)
        qcnt = qhi - qlo
        if merge_mode == "p2p":
            handle = None
            try:
                exchange = ShardExchange(local, rank, world, nq, k, dim)
                handle = exchange.local_handle()
            except Exception as e:
                print(f"[rank {rank}] peer-memory exchange unavailable ({e})", file=sys.stderr)
            handles = ctx.gather_objects(handle)        # every rank takes part, whatever happened above
            ok = 1.0
            try:
                if any(h is None for h in handles):
                    raise RuntimeError("a rank could not create its exchange buffer")
                exchange.connect(handles)
            except Exception as e:  # e.g. CUDA IPC not permitted in this container
                print(f"[rank {rank}] peer-memory exchange unavailable ({e}); using NCCL all-gather", file=sys.stderr)
                ok = 0.0
            if -ctx.max_over_ranks(-ok) == 0.0:           # all ranks must agree on the protocol
                exchange, merge_mode = None, "nccl"
    n_out = max(qcnt, 1) if (world > 1 and merge_mode == "p2p") else nq
    out_ids = torch.zeros((n_out, k), dtype=torch.int64, device=dev)
    out_dist = torch.zeros((n_out, k), dtype=torch.float32, device=dev)
    out_cnt = torch.zeros((n_out,), dtype=torch.int32, device=dev)
    if world > 1 and merge_mode == "nccl":
        keys = torch.zeros((nq, k), dtype=torch.int64, device=dev)
        all_keys = torch.zeros((world, nq, k), dtype=torch.int64, device=dev)
        all_ids = torch.zeros((world, nq, k), dtype=torch.int64, device=dev)
    n_host_q = max(qcnt, 1) if exchange is not None else nq
    hq = PinnedBuffer((n_host_q, dim), np.float32)
    h_ids, h_dist, h_cnt = PinnedBuffer((n_out, k), np.uint64), PinnedBuffer((n_out, k), np.float32), PinnedBuffer((n_out,), np.uint32)
    if exchange is not None:
        hq.array[:qcnt] = q_np[qlo:qlo + qcnt]
    else:
        hq.array[:] = q_np
    q_pageable = np.ascontiguousarray(hq.array.copy())
    pg_ids, pg_dist, pg_cnt = np.zeros((n_out, k), np.uint64), np.zeros((n_out, k), np.float32), np.zeros((n_out,), np.uint32)

    def step_device():
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if world == 1:
            _check(lib.scn_search_flat_dev(store.handle, p(q_dev), nq, k, p(out_ids), p(out_dist), p(out_cnt), stream))
        elif exchange is not None:
            exchange.search(store, q_dev.data_ptr(), nq, row0, out_ids.data_ptr(), out_dist.data_ptr(), out_cnt.data_ptr(),
                            torch.cuda.current_stream().cuda_stream)
        else:
            _check(lib.scn_search_flat_shard_dev(store.handle, p(q_dev), nq, k, row0, p(keys), p(out_ids), stream))
            ctx.dist.all_gather_into_tensor(all_keys, keys)
            ctx.dist.all_gather_into_tensor(all_ids, out_ids)
            _check(lib.scn_merge_topk_dev(local, p(all_keys), p(all_ids), world, nq, k, p(out_ids), p(out_dist),
                                          p(out_cnt), stream))

    def step_e2e(q_ptr=None, o=None):
        # the call a user of the C ABI makes: host buffers in, host buffers out, blocking
        q_ptr = hq.ptr if q_ptr is None else q_ptr
        oi, od, oc = (h_ids.ptr, h_dist.ptr, h_cnt.ptr) if o is None else o
        if world == 1:
            _check(lib.scn_search_flat(store.handle, C.c_void_p(q_ptr), nq, k, C.c_void_p(oi), C.c_void_p(od), C.c_void_p(oc)))
        elif exchange is not None:
            exchange.search_host(store, q_ptr, nq, row0, oi, od, oc)
        else:   # NCCL comparison path: torch copies around the device-buffer entry points
            q_dev.copy_(torch.from_numpy(hq.array), non_blocking=True)
            step_device()
            h_ids.array[:] = out_ids.cpu().numpy().view(np.uint64)
            h_dist.array[:] = out_dist.cpu().numpy()

    ms_total, timings, counters, launches, clocks = time_device(ctx, step_device, steps, warmup, store)
    e2e_s = time_host(ctx, step_e2e, steps, max(1, warmup // 2))
    pageable_s = None
    if merge_mode != "nccl":
        pageable_s = time_host(ctx, lambda: step_e2e(q_pageable.ctypes.data, (pg_ids.ctypes.data, pg_dist.ctypes.data, pg_cnt.ctypes.data)),
                               steps, 1)

    # ---- verification: a sample of this run's results against the CPU oracle over the same rows ----------
    step_e2e()
    n_s = min(16 if rows * dim < 4e9 else 8, qcnt if world > 1 else nq, nq)
    n_s = int(-ctx.max_over_ranks(-float(n_s)))      # the same sample size on every rank (smallest slice)
    sample_q = np.ascontiguousarray(q_np[:max(n_s, 1)])   # the first queries: rank 0's slice
    t0 = time.perf_counter()
    o_ids, o_dist = oracle_flat_sharded(ctx, store, n_local, row0, sample_q, metric, k)
    verify_s = time.perf_counter() - t0
    verified = None
    if rank == 0:
        same = bool(np.array_equal(h_ids.array[:n_s], o_ids[:n_s]) and np.array_equal(h_dist.array[:n_s], o_dist[:n_s]))
        same_pg = None if pageable_s is None else bool(np.array_equal(pg_ids[:n_s], o_ids[:n_s]) and np.array_equal(pg_dist[:n_s], o_dist[:n_s]))
        verified = {"queries": n_s, "identical": same, "identical_pageable_call": same_pg,
                    "against": f"CPU oracle flat scan over all {rows} rows (each rank scans the rows read back from its own GPU; "
                               "lists merged by (distance, row)); ids and fp32 distance bits compared",
                    "seconds": round(verify_s, 1)}

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) --------------------------------------------
    cpu = None
    if world == 1 and cpu_baseline and rank == 0:
        threads = ctx.threads
        db_host = np.concatenate([store.get(np.arange(i, min(i + 65536, n_local), dtype=np.uint64) + 1)
                                  for i in range(0, n_local, 65536)])
        ns = min(nq, threads)
        rounds = 16 if rows * dim >= 5e8 else 32   # about 10 s of host work at C2 (0.6 s per round)
        qps, dt = cpu_flat_qps(db_host, q_np[:ns], metric, k, threads, rounds)
        cpu = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
               "sample": f"{ns * rounds} queries ({rounds} rounds x {ns}, one per thread) over the full database, {dt:.1f}s wall"}
        del db_host

    block = None
    if rank == 0:
        ms_step = ms_total / steps
        roof = roofline_of(ctx, timings, counters, clocks, rows=rows, n_local=n_local, dim=dim, metric=metric, nq=nq, world=world)
        h2d = (qcnt if exchange is not None else nq) * dim * 4
        d2h = (qcnt if exchange is not None else nq) * (k * 12 + 4)
        e2e = {"value": nq * steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "timer": "host wall clock around the blocking host-buffer C-ABI call ("
                        + ("scn_search_flat" if world == 1 else "scn_search_flat_exchange: this rank's query slice in, its results out"
                           if exchange is not None else "torch copies + scn_search_flat_shard_dev + NCCL")
                        + "), pinned buffers (scn_host_alloc); bytes are per rank"}
        if pageable_s is not None:
            e2e["pageable"] = {"value": nq * steps / pageable_s, "unit": "queries/s",
                               "note": "the same call from pageable host memory (a Go slice): staged through the library's pinned chunks"}
        block = {
            "workload": f"{name}: {rows}x{dim} {METRIC_NAME[metric]} flat k={k} nq={nq}", "value": nq / (ms_step * 1e-3),
            "unit": "queries/s", "ms_per_step": ms_step, "steps": steps, "warmup": warmup,
            "dtype": "bf16 filter + f32 exact rerank" if (timings and "tensor_filter" in timings) else "f32",
            "config": {"workload": f"{name}: {rows}x{dim} {METRIC_NAME[metric]} flat k={k} nq={nq}",
                       "rows": rows, "dim": dim, "metric": METRIC_NAME[metric], "nq": nq, "k": k, "sharding": f"rows/{world}",
                       "shard_merge": {"p2p": "fused over NVLink peer memory: query slices gathered and top-k lists sent to the owner of "
                                              "each query slice as P2P stores + flags; each rank merges its slice",
                                       "nccl": "NCCL all_gather of keys and ids + merge kernel", "none": None}[merge_mode],
                       "l2_policy": "database (>= 3 GB fp32 + bf16 mirror per pass) is far larger than the 126 MB L2; no flush needed"
                       if n_local * dim * 4 > 4e8 else "working set fits L2: flush not applied (small workload, not the headline)"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "verified": verified,
            "kernels_ms_per_step": {n_: v[0] / steps for n_, v in timings.items()}, "counters": counters,
        }
    if exchange is not None:
        ctx.barrier()
        exchange.close()
    for b in (hq, h_ids, h_dist, h_cnt):
        b.close()
    store.close()
    del q_dev, out_ids, out_dist, out_cnt
    torch.cuda.empty_cache()
    return block


def recall_at_k(ids: np.ndarray, gt: np.ndarray) -> float:
    k = gt.shape[1]
    return float(np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / k for i in range(len(gt))]))


def bench_hnsw(ctx: Ctx, wl, name: str, ef: int, steps: int, warmup: int, cpu_baseline: bool, ef_sweep):
    """HNSW: replicas only. Every rank holds the whole store + graph and answers its slice of the
    query batch; no data-path collective. Host-generated data (numpy, seeded) so the cached graph
    (bench_cache/) matches the vectors bit for bit on any machine. The graph is the one the
    reference's algorithm builds (hnsw.go:148-257, restated by the oracle), constructed / loaded
    outside every timed region."""
    torch, lib = ctx.torch, ctx.lib
    from scintirete_b200 import DeviceStore, DistanceMetric, GraphState, PinnedBuffer
    from scintirete_b200.index import _check

    rows, dim, metric, nq_total, k, _ = wl
    world, rank, local, dev = ctx.world, ctx.rank, ctx.local, ctx.dev
    db_host = gen_rows_numpy(0, rows, dim, SEED_DB)
    q_all = gen_rows_numpy(0, nq_total, dim, SEED_Q)
    qlo, qhi = (nq_total * rank) // world, (nq_total * (rank + 1)) // world
    nq = qhi - qlo
    h, build_s = hnsw_graph_cached(db_host, metric, ef)
    store = DeviceStore(dim, DistanceMetric(metric), device=local)
    store.append(db_host)
    st = h.export_graph_state(with_vectors=False)
    store.graph_upload(GraphState(st.ids, st.list_counts, st.edge_counts, st.edges, st.entrypoint, st.max_layer, st.size, m=16))
    for kv in ctx.args.opt:
        oname, val = kv.split("=")
        store.set_option(oname, int(val))
    hq = PinnedBuffer((max(nq, 1), dim), np.float32)
    hq.array[:nq] = q_all[qlo:qhi]
    q_dev = torch.from_numpy(hq.array).to(dev)
    out_ids = torch.zeros((max(nq, 1), k), dtype=torch.int64, device=dev)
    out_dist = torch.zeros((max(nq, 1), k), dtype=torch.float32, device=dev)
    out_cnt = torch.zeros((max(nq, 1),), dtype=torch.int32, device=dev)
    h_ids, h_dist, h_cnt = PinnedBuffer((max(nq, 1), k), np.uint64), PinnedBuffer((max(nq, 1), k), np.float32), PinnedBuffer((max(nq, 1),), np.uint32)
    torch.cuda.synchronize()

    def p(t):
        return C.c_void_p(t.data_ptr())

    cur_ef = [ef]

    def step_device():
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _check(lib.scn_search_hnsw_dev(store.handle, p(q_dev), nq, k, cur_ef[0], p(out_ids), p(out_dist), p(out_cnt), stream))

    def step_e2e():
        _check(lib.scn_search_hnsw(store.handle, C.c_void_p(hq.ptr), nq, k, cur_ef[0], C.c_void_p(h_ids.ptr), C.c_void_p(h_dist.ptr),
                                   C.c_void_p(h_cnt.ptr)))

    ms_total, timings, counters, launches, clocks = time_device(ctx, step_device, steps, warmup, store)
    e2e_s = time_host(ctx, step_e2e, steps, max(1, warmup // 2))
    block = None
    if rank == 0:
        threads = ctx.threads
        step_e2e()
        gpu_ids = h_ids.array.copy()
        # exact ground truth for recall@k: the GPU flat scan over the same store (itself oracle-checked in tests/)
        gt_ids, _, _ = store.search_flat(hq.array[:nq], k)
        # the oracle's walk of the same graph on a sample: ids must be identical (the kernels reproduce
        # the reference's walk), so recall is identical by construction — and measured both ways
        n_s = min(nq, 500)
        o_ids, o_dist, o_cnt, _ = h.search_batch(hq.array[:n_s], k, ef, nthreads=threads)
        same = bool(np.array_equal(gpu_ids[:n_s], o_ids) and np.array_equal(h_dist.array[:n_s], o_dist))
        verified = {"queries": n_s, "identical": same,
                    "against": "CPU oracle HNSW.Search over the same graph (ids and fp32 distance bits)",
                    "recall_at_k_gpu": recall_at_k(gpu_ids[:n_s], gt_ids[:n_s]), "recall_at_k_oracle": recall_at_k(o_ids, gt_ids[:n_s]),
                    "recall_at_k_gpu_all_queries": recall_at_k(gpu_ids[:nq], gt_ids)}
        cpu = None
        if world == 1 and cpu_baseline:
            reps, t0 = 0, time.perf_counter()
            while reps < 64 and time.perf_counter() - t0 < 10.0:   # the whole batch, repeated for about 10 s of host work
                h.search_batch(hq.array[:nq], k, ef, nthreads=threads)
                reps += 1
            dt = time.perf_counter() - t0
            cpu = {"value": nq * reps / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                   "sample": f"{reps} x {nq} queries, ef={ef}, same graph, {dt:.1f}s wall; graph "
                             + (f"built in {build_s:.0f}s (1 thread)" if build_s else "loaded from bench_cache/")}
        ms_step = ms_total / steps
        roof = roofline_of(ctx, timings, counters, clocks, rows=rows, n_local=rows, dim=dim, metric=metric, nq=nq, world=world, ef=ef)
        sweep = []
        for e in ef_sweep:
            # queries/s at recall: one point per ef on the same graph (3 timed steps each; recall of the GPU on
            # all queries, of the oracle on the sample; ef < k returns fewer than k results, as in the reference)
            cur_ef[0] = e
            for _ in range(2):
                step_device()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                step_device()
            e1.record()
            torch.cuda.synchronize()
            step_e2e()
            oi, _, _, _ = h.search_batch(hq.array[:n_s], k, e, nthreads=threads)
            sweep.append({"ef": e, "value": nq * 3 / (e0.elapsed_time(e1) * 1e-3), "unit": "queries/s",
                          "recall_at_k_gpu": recall_at_k(h_ids.array[:nq], gt_ids), "recall_at_k_oracle_sample": recall_at_k(oi, gt_ids[:n_s]),
                          "identical_to_oracle_sample": bool(np.array_equal(h_ids.array[:n_s], oi))})
        cur_ef[0] = ef
        # the same batch with exact tie handling on: walks that end on a distance tie at the edge of W are redone
        store.set_option("hnsw_exact_ties", 1)
        store.set_option("profile", 1)
        for _ in range(2):
            step_device()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step_device()
        e1.record()
        torch.cuda.synchronize()
        redone = store.last_counters()[2]
        step_e2e()
        oi, _, _, _ = h.search_batch(hq.array[:n_s], k, ef, nthreads=threads)
        exact_ties = {"value": nq * 5 / (e0.elapsed_time(e1) * 1e-3), "unit": "queries/s", "walks_redone_per_batch": int(redone),
                      "identical_to_oracle_sample": bool(np.array_equal(h_ids.array[:n_s], oi)),
                      "note": "option hnsw_exact_ties=1: results equal the reference's even through exact float ties; one re-walk is a "
                              "single-warp latency chain"}
        store.set_option("hnsw_exact_ties", 0)
        store.set_option("profile", 0)
        block = {
            "workload": f"{name}: {rows}x{dim} {METRIC_NAME[metric]} hnsw k={k} nq={nq_total} ef={ef}",
            "value": nq_total / (ms_step * 1e-3), "unit": "queries/s", "ms_per_step": ms_step, "steps": steps, "warmup": warmup,
            "dtype": "f32",
            "config": {"workload": f"{name}: {rows}x{dim} {METRIC_NAME[metric]} hnsw k={k} nq={nq_total} ef={ef}", "rows": rows,
                       "dim": dim, "metric": METRIC_NAME[metric], "nq": nq_total, "k": k, "ef": ef, "M": 16, "efConstruction": 200,
                       "sharding": f"replicas x{world}, query batch split",
                       "l2_policy": "1M x 128 fp32 rows + adjacency (0.64 GB) exceed the 126 MB L2; walks are random gathers"
                       if rows * dim * 4 > 4e8 else "working set fits L2 (small workload, not the headline)",
                       "operating_point": "the reference's plain closest-M neighbour selection (hnsw.go:560-583, no diversity "
                                          "heuristic) on i.i.d. Gaussian data caps recall well below 1; GPU and oracle walks are "
                                          "identical, so recall is the reference's own at every ef — see ef_sweep"},
            "e2e": {"value": nq_total * steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4,
                    "d2h_bytes_per_step": nq * (k * 12 + 4),
                    "timer": "host wall clock around the blocking host-buffer C-ABI call (scn_search_hnsw), pinned buffers"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "verified": verified,
            "ef_sweep": sweep, "exact_ties": exact_ties, "kernels_ms_per_step": {n_: v[0] / steps for n_, v in timings.items()}, "counters": counters,
        }
    ctx.barrier()
    for b in (hq, h_ids, h_dist, h_cnt):
        b.close()
    store.close()
    del q_dev, out_ids, out_dist, out_cnt, h, db_host
    torch.cuda.empty_cache()
    return block


def bench_build(ctx: Ctx, rows: int, dim: int, metric: int, extra: int = 2000):
    """GPU-assisted HNSW construction (scn_hnsw_insert) on the data of a cached reference graph:
    the whole graph is rebuilt on the GPU from the same level draws and compared edge for edge with the graph the
    reference's serial algorithm built (oracle; bench_cache/). CPU baseline on a bounded sample: `extra` further
    inserts into the finished graph by the oracle on one core (insertVector is serial), the same inserts on the GPU."""
    import oracle
    from scintirete_b200 import DeviceStore, DistanceMetric

    if ctx.rank != 0:
        return None
    z = np.load(graph_cache_path(rows, dim, metric))
    db = gen_rows_numpy(0, rows + extra, dim, SEED_DB)
    store = DeviceStore(dim, DistanceMetric(metric), device=ctx.local)
    store.append(db[:rows])
    launches0 = ctx.lib.scn_launch_count()
    st = store.hnsw_insert(z["list_counts"] - 1, 16, 200)
    launches = ctx.lib.scn_launch_count() - launches0
    g = store.graph_export(16)
    same = bool(np.array_equal(g.node_ids, z["ids"]) and np.array_equal(g.list_counts, z["list_counts"])
                and np.array_equal(g.edge_counts, z["edge_counts"]) and np.array_equal(g.edges, z["edges"].astype(np.uint64))
                and g.entry_point == int(z["entrypoint"]) and g.max_layer == int(z["max_layer"]))
    # the same `extra` inserts into the finished graph: oracle (one core) vs GPU
    levels = np.minimum(np.floor(-np.log(np.random.default_rng(99).random(extra)) / np.log(2.0)), 15).astype(np.int32)
    h = oracle.OracleHNSW(M=16, ef_construction=200, ef_search=128, max_layers=16, seed=42, metric=metric)
    h.import_graph_state(oracle.GraphState(z["ids"], z["deleted"], z["list_counts"], z["edge_counts"], z["edges"].astype(np.uint64),
                                           db[:rows], int(z["entrypoint"]), int(z["max_layer"]), int(z["size"])))
    t0 = time.perf_counter()
    for i in range(extra):
        h.insert(rows + i + 1, db[rows + i], level=int(levels[i]))
    cpu_s = time.perf_counter() - t0
    store.append(db[rows:])
    st2 = store.hnsw_insert(levels, 16, 200)
    o = h.export_graph_state(with_vectors=False)
    g2 = store.graph_export(16)
    same2 = bool(np.array_equal(g2.edge_counts, o.edge_counts) and np.array_equal(g2.edges, o.edges)
                 and g2.entry_point == o.entrypoint and g2.max_layer == o.max_layer)
    store.close()
    return {
        "workload": f"build: {rows}x{dim} {METRIC_NAME[metric]} HNSW construction M=16 efC=200 (reference-serial semantics, GPU-assisted)",
        "metric": "inserts/sec", "value": rows / st["seconds"], "unit": "inserts/s", "seconds": st["seconds"],
        "rounds": st["rounds"], "commits_per_round": rows / max(st["rounds"], 1), "speculative_searches": st["searches"],
        "rounds_ended_by": dict(zip(["added_while_not_full", "added_would_be_admitted", "admitted_neighbour_left", "entry_or_maxlayer_moved",
                                     "log_incomplete", "beyond_pair_matrix"], st["conflict_kind"])),
        "device_seconds": st["device_seconds"], "commit_seconds": st["commit_seconds"], "gpu_launches": int(launches),
        "verified": {"identical": same, "against": "the cached graph the oracle (reference algorithm, serial) built from the same data and "
                                                   "level draws: every adjacency list in stored order, entry point, maxLayer"},
        "at_full_size": {"inserts": extra, "gpu_inserts_per_s": extra / st2["seconds"], "cpu_inserts_per_s": extra / cpu_s,
                         "identical": same2, "note": f"the same {extra} further inserts into the finished {rows}-node graph"},
        "cpu_baseline": {"value": extra / cpu_s, "unit": "inserts/s", "cores": 1, "kind": "port",
                         "sample": f"{extra} inserts into the finished {rows}-node graph, {cpu_s:.1f}s (insertVector is serial: one core)"},
    }


def run_ours(args, wl, name):
    ctx = Ctx(args)
    rows, dim, metric, nq, k, kind = wl
    ef_sweep = [int(x) for x in args.ef_sweep.split(",") if x] if args.ef_sweep else []
    if kind == "flat":
        main = bench_flat(ctx, wl, name, args.steps, args.warmup, not args.no_cpu_baseline)
    else:
        main = bench_hnsw(ctx, wl, name, args.ef, args.steps, args.warmup, not args.no_cpu_baseline, ef_sweep)
    secondary = []
    if not args.no_secondary and name == "c2" and not (args.rows or args.dim or args.nq):
        # the other BASELINE.json configurations, measured in the same run so that they are on the driver's record
        if ctx.world == 1 and os.path.exists(graph_cache_path(*WORKLOADS["c3"][:3])):
            secondary.append(bench_hnsw(ctx, WORKLOADS["c3"], "c3", 128, args.steps, args.warmup, not args.no_cpu_baseline,
                                        [16, 32, 64, 256, 512]))
        if ctx.world == 1 and os.path.exists(graph_cache_path(*WORKLOADS["c1"][:3])) and not args.no_cpu_baseline:
            secondary.append(bench_build(ctx, *WORKLOADS["c1"][:3]))
        if ctx.world == 1 and ctx.rank == 0 and not os.path.exists(graph_cache_path(*WORKLOADS["c3"][:3])):
            secondary.append({"workload": "c3", "skipped": "bench_cache/ holds no reference-built 1M x 128 graph on this box "
                                                           "(the serial reference construction takes hours; see DESIGN.md)"})
        if ctx.world >= 2:
            secondary.append(bench_flat(ctx, WORKLOADS["c4"], "c4", max(2, min(args.steps, 5)), 3, False))
    if ctx.rank == 0:
        line = {
            "metric": "queries/sec", "value": main["value"], "unit": "queries/s", "n_gpus": ctx.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": main["dtype"],
            "data": "synthetic gaussian (torch.randn on device, seeded; the oracle checks read the same rows back from the GPU)"
            if kind == "flat" else "synthetic gaussian (numpy, seeded)",
            "config": main["config"], "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "clocks": main["clocks"],
            "roofline": main["roofline"], "cpu_baseline": main["cpu_baseline"], "verified": main["verified"],
            "kernels_ms_per_step": main["kernels_ms_per_step"], "counters": main["counters"],
        }
        if "ef_sweep" in main:
            line["ef_sweep"] = main["ef_sweep"]
        if secondary:
            line["secondary"] = [s for s in secondary if s]
        emit(line)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


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
    ap.add_argument("--metric", type=int, choices=[1, 2, 3])
    ap.add_argument("--ef", type=int, default=128)
    ap.add_argument("--ef-sweep", default="", help="HNSW workloads: comma-separated ef values measured after the main point")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="default c2 run: skip the C3 (N = 1) / C4 (N >= 2) blocks")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (scn_set_option), repeatable")
    ap.add_argument("--merge", default="p2p", choices=["p2p", "nccl"], help="row-shard exchange for --gpus > 1")
    args = ap.parse_args()
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    wl = list(WORKLOADS[args.workload])
    if args.rows:
        wl[0] = args.rows
    if args.dim:
        wl[1] = args.dim
    if args.nq:
        wl[3] = args.nq
    if args.metric:
        wl[2] = args.metric
    if args.workload == "c1" and args.ef == 128:
        args.ef = 100
    if args.impl == "reference":
        run_reference(args, tuple(wl))
    else:
        run_ours(args, tuple(wl), args.workload)


if __name__ == "__main__":
    main()

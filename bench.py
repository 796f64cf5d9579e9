#!/usr/bin/env python
"""bench.py — QPS@recall@10 of the IVF-FLAT batched search on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|small] [--shard rows|replicas|lists]

A step = one pass of the hot path over one batch: coarse quantiser -> list-major scan of the probed lists ->
per-query top-k, for the whole 10k-query batch.  `value` is measured with the index and the queries resident in
HBM; `e2e` goes through the C ABI with HOST (pinned) query buffers and host result buffers.

N = 1: BASELINE.json configs[1] (IVF-FLAT 1M x 768, nlist 1024, nprobe 32, 10k queries, top-10).
N > 1 (torchrun, one process per GPU): ONE global IVF-FLAT index of N x 12.5M rows (100M x 768 at N = 8, nlist =
8192 N = 65,536 at N = 8: configs[3]'s build), rows sharded in contiguous blocks.  Build = data-parallel k-means inside the
library (hb_sharded_ivf_build: one NCCL all-reduce per Lloyd round); a step = hb_sharded_search on every rank: coarse
routing (replicated centroids), scan of this rank's part of every probed list, exchange of the local top-k over NVLink peer
windows fused with the merge (one kernel; `--opt comm_p2p=0`: ncclAllGather + merge kernel).  Weak scaling: rows per GPU are
fixed, the database grows with N.  The r01 layout (every GPU a replica of configs[1], no collective) is kept as the
secondary key `replicas`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# per-GPU shard of the N > 1 workload: N x 12.5M rows = 100M at N = 8; nlist = 8192 per GPU (65,536 at N = 8, configs[3])
SHARD = dict(n_per=12_500_000, d=768, nlist_per=8192, nprobe=32, nq=10_000, k=10, centres_per=16384, noise=0.1, iters=10,
             truth_queries=1024)
SHARD_SMALL = dict(n_per=500_000, d=768, nlist_per=512, nprobe=16, nq=4_000, k=10, centres_per=1024, noise=0.1, iters=4,
                   truth_queries=512)
WORKLOADS = {
    # BASELINE.json configs[1]: IVF-FLAT 1M x 768 fp32 cosine, nlist=1024, nprobe=32, 10k-query batch, top-10
    "c2": dict(n=1_000_000, d=768, nlist=1024, nprobe=32, nq=10_000, k=10, centres=2048, noise=0.1, iters=10),
    "small": dict(n=100_000, d=768, nlist=256, nprobe=16, nq=2_000, k=10, centres=512, noise=0.1, iters=5),
}
METRIC = "QPS@recall@10 (IVF-FLAT 1Mx768 fp32 cosine, nlist=1024, nprobe=32, 10k-query batch, top-10)"
METRIC_SHARDED = "QPS@recall@10 (IVF-FLAT Nx12.5Mx768 fp32 cosine row-sharded over N GPUs, 100Mx768 at N=8, nprobe=32, 10k-query batch, top-10)"


def workload_name(w, key):
    return (f"ivf-flat {w['n']}x{w['d']} fp32 cosine nlist={w['nlist']} nprobe={w['nprobe']} nq={w['nq']} k={w['k']} "
            f"(BASELINE configs[1])" if key == "c2" else f"ivf-flat reduced ({key})")


# ---------------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons via NVML (the counters nvidia-smi prints), sampled every 5 ms."""

    def __init__(self, index: int):
        self.index, self.sm, self.mx, self.reasons, self.power = index, [], None, set(), []
        self._stop = threading.Event()
        self.t = None

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.nv = nv
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def stop(self) -> dict:
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self._stop.set()
        self.t.join(timeout=1)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------------------
# synthetic data (clustered, cf. test/data_generator.clj:74-79)
# ---------------------------------------------------------------------------------------------------------
def gen_gpu(w, device, query_seed=43):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(42)
    rows = torch.empty((w["n"], w["d"]), dtype=torch.float32, device=device)
    step = 131072
    if w.get("structureless"):  # i.i.d. Gaussian rows and queries: no cluster structure for IVF (or the pruning) to exploit
        for i in range(0, w["n"], step):
            m = min(step, w["n"] - i)
            rows[i:i + m] = torch.randn((m, w["d"]), generator=g, device=device)
        g.manual_seed(query_seed)
        return rows, torch.randn((w["nq"], w["d"]), generator=g, device=device).contiguous()
    centres = torch.randn((w["centres"], w["d"]), generator=g, device=device)
    for i in range(0, w["n"], step):
        m = min(step, w["n"] - i)
        idx = torch.randint(0, w["centres"], (m,), generator=g, device=device)
        rows[i:i + m] = centres[idx] + w["noise"] * torch.randn((m, w["d"]), generator=g, device=device)
    g.manual_seed(query_seed)
    idx = torch.randint(0, w["centres"], (w["nq"],), generator=g, device=device)
    queries = centres[idx] + w["noise"] * torch.randn((w["nq"], w["d"]), generator=g, device=device)
    return rows, queries.contiguous()


def gen_cpu(w):
    r = np.random.default_rng(42)
    centres = r.standard_normal((w["centres"], w["d"]), dtype=np.float32)
    rows = np.empty((w["n"], w["d"]), dtype=np.float32)
    step = 65536
    for i in range(0, w["n"], step):
        m = min(step, w["n"] - i)
        rows[i:i + m] = centres[r.integers(0, w["centres"], m)] + np.float32(w["noise"]) * r.standard_normal((m, w["d"]), dtype=np.float32)
    r = np.random.default_rng(43)
    queries = centres[r.integers(0, w["centres"], w["nq"])] + np.float32(w["noise"]) * r.standard_normal((w["nq"], w["d"]), dtype=np.float32)
    return rows, np.ascontiguousarray(queries, dtype=np.float32)


def cpu_partitions(rows, nlist, iters=2):
    """Setup for the CPU arm only: a plain BLAS Lloyd (seeded random rows) that yields an IVF partitioning of the
    same shape.  The reference's own k-means++ is O(N*k^2*D) (ivf_flat.clj:43-49) — hours at 1M x 1024 on CPU."""
    import torch

    r = np.random.default_rng(7)
    x = torch.from_numpy(rows)
    xn = x / x.norm(dim=1, keepdim=True)
    cents = x[torch.from_numpy(r.choice(rows.shape[0], nlist, replace=False))].clone()
    asg = None
    for it in range(iters + 1):
        cn = cents / cents.norm(dim=1, keepdim=True).clamp_min(1e-30)
        asg = torch.empty(rows.shape[0], dtype=torch.int64)
        for i in range(0, rows.shape[0], 65536):
            asg[i:i + 65536] = (xn[i:i + 65536] @ cn.T).argmax(dim=1)
        if it == iters:
            break
        sums = torch.zeros((nlist, rows.shape[1]), dtype=torch.float32).index_add_(0, asg, x)
        cnt = torch.bincount(asg, minlength=nlist).clamp_min(1).unsqueeze(1)
        cents = sums / cnt
    return cents.double().numpy(), asg.numpy().astype(np.int32)


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU path restated (oracle/) on the host cores, bounded sample per step
# ---------------------------------------------------------------------------------------------------------
def time_cpu_search(rows, cents, asg, queries, k, nprobe, budget_s, steps=1, warmup=0):
    from oracle import oracle as orc

    cores = orc.ncores()
    probe = min(len(queries), 2 * cores)
    t0 = time.perf_counter()
    orc.ivf_search(rows, cents, asg, queries[:probe], k, nprobe, nthreads=cores)
    per_query_core_s = (time.perf_counter() - t0) * cores / probe  # includes the one-off list/norm setup: conservative
    per_step = budget_s / max(steps + warmup, 1)
    sample = int(max(cores, min(len(queries), per_step / max(per_query_core_s, 1e-9) * cores)))
    sample -= sample % cores or 0
    sample = max(sample, cores)
    qs = queries[:sample]
    # setup (lists, norms) outside the timed loop: the reference precomputes them at build time (ivf_flat.clj:161-179)
    off, lrows = orc.build_lists(asg, cents.shape[0])
    norms = orc.row_norms(rows)
    L = orc.lib()
    ids = np.empty((sample, k), dtype=np.int64)
    dist = np.empty((sample, k), dtype=np.float64)
    args = (orc._ptr(rows), rows.shape[0], rows.shape[1], orc._ptr(cents), cents.shape[0], orc._ptr(off), orc._ptr(lrows),
            orc._ptr(norms), orc._ptr(qs), sample, k, nprobe, orc.COSINE, orc._ptr(ids), orc._ptr(dist), None, cores)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        L.orc_ivf_search(*args)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return {"qps": sample / (sum(times) / len(times)), "cores": cores, "sample": sample, "ids": ids, "dist": dist,
            "ms_per_step": 1e3 * sum(times) / len(times)}


def run_reference(args, w, key):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows, queries = gen_cpu(w)
    cents, asg = cpu_partitions(rows, w["nlist"])
    r = time_cpu_search(rows, cents, asg, queries, w["k"], w["nprobe"], budget_s=150.0, steps=args.steps, warmup=args.warmup)
    sample_desc = (f"{r['sample']} of {w['nq']} queries per step, {r['cores']} threads, one query per task "
                   f"(parallel-search-futures); index partitions from a BLAS Lloyd (setup, untimed)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["qps"], "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(w, key), "note": "C restatement of the reference's Clojure CPU path "
                   "(oracle/; no JVM in this image)"},
        "cpu_baseline": {"value": r["qps"], "unit": "queries/s", "cores": r["cores"], "kind": "port", "sample": sample_desc},
        "e2e": {"value": r["qps"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
# dram__bytes_read.sum + dram__bytes_write.sum of one tc_pass_kernel<2,EMIT> list-scan launch, from `ncu --set full` captures of
# this command: (workload, digits, probe pruning on) -> bytes.  profiles/r01c_tc_pass_ncu_raw.csv (5.42 GB + 0.03 GB, no pruning),
# profiles/r01v_tc_pass_ncu_raw.csv (final build, probe pruning and M = 64 units on: 1.760 GB + 0.023 GB)
TRAFFIC = {("c2", 2, False): 5.45e9, ("c2", 2, True): 1.783e9}


def setup(args):
    """One process per GPU: device, library, process group (NCCL) and the library's own communicator."""
    import torch
    import torch.distributed as dist

    from hnsw_clj_b200 import _lib as hb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    hb.check(hb.lib().hb_init(local))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    hb.set_option("fast_digits", args.digits)
    for kv in args.opt:
        name, _, val = kv.partition("=")
        hb.set_option(name, int(val))
    return rank, world, local, device


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def run_ours(args, w, key, ctx, replicas_only=False):
    import torch
    import torch.distributed as dist

    from hnsw_clj_b200 import _lib as hb
    from hnsw_clj_b200 import ivf_flat
    from hnsw_clj_b200.flat import FlatIndex, recall_at_k
    from hnsw_clj_b200.sharded import ShardedIVFFlat

    rank, world, local, device = ctx
    peaks = load_peaks()

    fast = args.mode == "fast"
    k, nprobe, nq = w["k"], w["nprobe"], w["nq"]
    # N > 1: "replicas" = every GPU holds the whole index (3 GB of 180 GB) and answers its OWN batch of nq queries: no
    # data-path collective, weak scaling, value = N * nq / t.  "lists" = ONE batch against the lists of the index
    # sharded l mod N, all-gather + merge kernel: strong scaling (the layout for an index larger than one GPU).
    replicas = world > 1 and (args.shard == "replicas" or replicas_only)
    rows, queries = gen_gpu(w, device, query_seed=43 + (rank if replicas else 0))
    torch.cuda.synchronize()

    # ---- build (untimed setup; reported) -------------------------------------------------------------
    t0 = time.perf_counter()
    if rank == 0:
        gix = ivf_flat.build_index(rows, num_partitions=w["nlist"], max_iterations=w["iters"])
        cents, asg = gix.export()
    build_s = time.perf_counter() - t0
    if world > 1:
        if rank != 0:
            cents = np.empty((w["nlist"], w["d"]), np.float64)
            asg = np.empty(w["n"], np.int32)
        tc, ta = torch.from_numpy(cents).to(device), torch.from_numpy(asg).to(device)
        dist.broadcast(tc, 0)
        dist.broadcast(ta, 0)
        cents, asg = tc.cpu().numpy(), ta.cpu().numpy()
    if world > 1 and not replicas:
        if rank == 0:
            gix.close()
        shard = ShardedIVFFlat(rows, cents, asg, rank, world)
        search = lambda q: shard.search(q, k, nprobe)  # noqa: E731
    else:
        if rank != 0:
            gix = ivf_flat.import_index(rows, cents, asg)  # the same index as rank 0's, without re-clustering
        out_ids = torch.empty((nq, k), dtype=torch.int64, device=device)
        out_dist = torch.empty((nq, k), dtype=torch.float64, device=device)
        search = lambda q: gix.search_raw(q, k, nprobe, out_ids=out_ids, out_dist=out_dist)  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: device-resident inputs -------------------------------------------------------------
    hb.set_mode(hb.MODE_FAST if fast else hb.MODE_EXACT)
    for _ in range(args.warmup):
        search(queries)
    barrier()
    hb.set_option("profile", 1)
    hb.launch_count(reset=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.ncu_region:
        hb.set_option("cuda_profiler", 1)
    e0.record()
    for _ in range(args.steps):
        ids, dists = search(queries)
    e1.record()
    barrier()
    if args.ncu_region:
        hb.set_option("cuda_profiler", 0)
    ms = e0.elapsed_time(e1) / args.steps
    launches = hb.launch_count()
    clk = clocks.stop() if rank == 0 else {}
    scan_ms = hb.get_stat("scan_ms")
    scan_n = max(hb.get_stat("scan_count"), 1.0)
    stats = {n: hb.get_stat(n) / args.steps for n in ("scan_ms", "coarse_ms", "select_ms", "plan_ms", "tc_ms", "tc_sample_ms",
                                                         "pack_ms", "rescore_ms")}
    tc_ms, tc_n = hb.get_stat("tc_ms"), max(hb.get_stat("tc_count"), 1.0)
    fast_served, fast_fell = hb.get_stat("fast_queries"), hb.get_stat("fast_fallbacks")
    # (query, probed list) pairs the scan dropped by the angle bound, and the rows they held, per step
    pruned_pairs = hb.get_stat("fast_pruned_pairs") / args.steps
    pruned_rows = hb.get_stat("fast_pruned_rows") / args.steps
    probe_pairs = hb.get_stat("fast_probe_pairs") / args.steps
    hb_stats = {n: hb.get_stat(n) for n in ("tc_units", "tc_items", "tc_tiles", "tc_half_units", "tc_narrow_units", "tc_narrow_items",
                                            "tc_narrow_slots")}
    hb.set_option("profile", 0)
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    jobs = world if replicas else 1  # query batches answered per step over all ranks
    value = jobs * nq / (ms * 1e-3)

    # ---- e2e: host (pinned) queries in, host results out, through the C ABI -----------------------------
    hq = torch.empty((nq, w["d"]), dtype=torch.float32, pin_memory=True)
    hq.copy_(queries)
    h_ids = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
    h_dist = torch.empty((nq, k), dtype=torch.float64, pin_memory=True)
    dq = torch.empty_like(queries)

    def e2e_step():
        if world > 1 and not replicas:
            dq.copy_(hq, non_blocking=True)  # every rank needs the whole batch
            i, d_ = search(dq)
            if rank == 0:
                h_ids.copy_(i, non_blocking=True)
                h_dist.copy_(d_, non_blocking=True)
            torch.cuda.synchronize()
        else:
            hb.check(hb.lib().hb_search(gix._h, hq.data_ptr(), hb.F32, nq, k, nprobe, h_ids.data_ptr(), h_dist.data_ptr()))

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {"value": jobs * nq / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(hq.numel() * 4) * (world if world > 1 else 1),
           "d2h_bytes_per_step": int(h_ids.numel() * 8 + h_dist.numel() * 8) * jobs}

    if replicas_only or rank != 0:
        gix.close() if (world == 1 or replicas) else shard.close()
        if world > 1:
            dist.barrier()
        return {"value": value, "unit": "queries/s", "ms_per_step": ms, "e2e": e2e, "gpu_launches": int(launches),
                "build_s": build_s, "workload": workload_name(w, key),
                "sharding": f"replicas: the whole index on each of {world} GPUs, one {nq}-query batch per GPU per step, no collective"}

    # ---- recall@10 against the exact flat search of the same rows (device, exact path) ----------------
    hb.set_mode(hb.MODE_EXACT)
    ids_np = ids.cpu().numpy() if hb._is_torch(ids) else ids
    dists_np = dists.cpu().numpy() if hb._is_torch(dists) else dists
    e2e_same = bool((h_ids.numpy() == ids_np).all())
    with FlatIndex(rows) as fx:
        exact_ids, _ = fx.search_raw(queries, k)
    recall = recall_at_k(ids_np, exact_ids)
    fast_vs_exact = None
    single = world == 1 or replicas  # this rank ran the whole single-GPU path
    if fast and single:
        # the same search in EXACT mode (fp64 for every pair): FAST must return the same ids and distance bits
        x_ids = torch.empty((nq, k), dtype=torch.int64, device=device)
        x_dist = torch.empty((nq, k), dtype=torch.float64, device=device)
        gix.search_raw(queries, k, nprobe, out_ids=x_ids, out_dist=x_dist)
        fast_vs_exact = {"queries": nq, "ids_equal": bool((x_ids.cpu().numpy() == ids_np).all()),
                         "dist_bits_equal": bool((x_dist.cpu().numpy().view(np.int64) == dists_np.view(np.int64)).all()),
                         "served_by_candidate_pass": int(fast_served - fast_fell), "exact_fallbacks": int(fast_fell),
                         "per_steps": args.steps}

    # ---- roofline of the dominant kernel (the list scan) ---------------------------------------------------
    rows_np = rows.cpu().numpy()
    q_np = queries.cpu().numpy()
    items = None
    if single:
        probes = gix.probes(queries, nprobe)
        lens = np.bincount(asg, minlength=w["nlist"])
        pairs = float(lens[probes.reshape(-1)].sum())
        per_list_q = np.bincount(probes.reshape(-1)[probes.reshape(-1) >= 0], minlength=w["nlist"])
        items = float((np.ceil(per_list_q / 128.0) * np.ceil(lens / 128.0)).sum())
    else:
        pairs = float("nan")
    flops = 2.0 * pairs * w["d"]
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1.59 PFLOP/s"
    if fast:
        # tc_pass_kernel: the candidate pass over every probed (query, list) pair.  Algorithmic work = one 768-dim dot
        # per pair; the kernel executes it as digits^2 int8 GEMM products on zero-padded 128 x 128 tiles.
        t_tc = (tc_ms / tc_n) * 1e-3
        nprod = 4 if args.digits == 2 else 6
        dpad = -(-w["d"] // 128) * 128
        # What the launch covered, counted by the library while profiling: units (<= 128 query slots of one list), items
        # (unit x row tile = one 128 x 128 x dpad block of digit products) and the distinct row tiles they read.  With probe
        # pruning the kernel only sees the (query, list) pairs that survive: its roofline is quoted on the rows it actually
        # scores and the bytes it has to read; the step's algorithmic work (every probed row, as the reference scans them)
        # stands beside it.
        algorithmic_pairs = pairs
        pairs = pairs - pruned_rows
        flops = 2.0 * pairs * w["d"]
        tc_units, tc_items, tc_tiles, tc_half = (hb_stats[n] / args.steps for n in ("tc_units", "tc_items", "tc_tiles", "tc_half_units"))
        tc_narrow, tc_narrow_items = hb_stats["tc_narrow_units"] / args.steps, hb_stats["tc_narrow_items"] / args.steps
        if tc_items > 0:
            items = tc_items
        int8_ops = 2.0 * items * 128 * 128 * dpad * nprod if items else None
        # digit images read once: the row tiles some unit scans + the units' query images (128 slots each, 64 for the
        # units that run with M = 64, the 8-slot groups in use for the narrow units of tc_narrow_kernel)
        slots = 128.0 * (tc_units - tc_half) + 64.0 * (tc_half - tc_narrow) + hb_stats["tc_narrow_slots"] / args.steps
        img_bytes = (tc_tiles * 128.0 * dpad * args.digits + slots * dpad * args.digits) if tc_items > 0 else (
            float(w["n"]) * dpad * args.digits + nq * nprobe * dpad * args.digits)
        hbm_peak = peaks.get("hbm_gbs", 7700.0)
        t_flops, t_bytes = flops / (bf16_peak * 1e12), img_bytes / (hbm_peak * 1e9)
        hbm_bound = t_bytes > t_flops  # the binding roofline is the larger of the two lower bounds on the launch time
        tensor = {"achieved_tflops": flops / t_tc / 1e12 if single else None, "peak_tflops": bf16_peak,
                  "frac": (flops / t_tc / 1e12 / bf16_peak) if single else None}
        hbm = {"achieved_gbs": img_bytes / t_tc / 1e9 if single else None, "peak_gbs": hbm_peak,
               "frac": (img_bytes / t_tc / 1e9 / hbm_peak) if single else None}
        roofline = {
            "kernel": (f"tc_narrow_kernel<{args.digits}> (tcgen05.mma kind::i8, rows on M, <= 32 queries on N: {tc_narrow_items:.0f} of "
                       f"{items or 0:.0f} items) + tc_pass_kernel<{args.digits},EMIT> (the rest): IVF list scan candidate pass"),
            "bound": "hbm" if hbm_bound else "tensor",
            "achieved": (hbm["achieved_gbs"] if hbm_bound else tensor["achieved_tflops"]),
            "peak": hbm_peak if hbm_bound else bf16_peak, "unit": "GB/s" if hbm_bound else "TFLOP/s",
            "frac": hbm["frac"] if hbm_bound else tensor["frac"],
            "peak_source": ("MEASURED_PEAKS.json " + ("hbm_gbs" if hbm_bound else "bf16_tflops (burst)")) if peaks else "fallback",
            "why_this_bound": f"lower bounds on the launch: unique digit-image bytes / HBM peak = {t_bytes * 1e3:.3f} ms, "
                              f"algorithmic flops / bf16 peak = {t_flops * 1e3:.3f} ms",
            # dram__bytes_read.sum + dram__bytes_write.sum of the candidate-pass launch: measured by an ncu pass of this very
            # command (measure_traffic, below); the constant is the last committed capture, used when ncu is unavailable
            "traffic": TRAFFIC.get((key, args.digits, pruned_rows > 0)),
            "launch_ms": t_tc * 1e3, "launches_per_step": tc_n / args.steps,
            "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": img_bytes,
            "tensor": tensor, "hbm": hbm,
            "units_per_launch": tc_units, "m64_units_per_launch": tc_half, "narrow_units_per_launch": tc_narrow,
            "narrow_items_per_launch": tc_narrow_items, "items_per_launch": items,
            "row_tiles_read_per_launch": tc_tiles,
            "scored_pairs_per_launch": pairs, "algorithmic_pairs_per_step": algorithmic_pairs,
            "probe_pruning": {"probe_pairs": probe_pairs, "pruned_probe_pairs": pruned_pairs, "pruned_rows": pruned_rows,
                              "note": "exact: a probed list is dropped for a query when cos(angle(q, centroid) - list radius) "
                                      "cannot reach the query's candidate threshold (results stay bit-identical); "
                                      "--opt fast_prune=0 scans every probed list"},
            "int8_pipe": {"executed_tops": int8_ops / t_tc / 1e12 if int8_ops else None,
                          "nominal_peak_tops": 2.0 * bf16_peak, "frac": (int8_ops / t_tc / 1e12 / (2.0 * bf16_peak)) if int8_ops else None,
                          "note": "executed = digit products on padded tiles; peak = 2 x measured bf16 (int8 runs at twice the "
                                  "bf16 rate on tcgen05)"},
            "step_breakdown_ms": stats,
        }
    else:
        unique_bytes = float(w["n"]) * w["d"] * 4 + nq * w["d"] * 4 + pairs * 8  # slab once + queries + distance scratch out
        t_scan = (scan_ms / scan_n) * 1e-3
        try:
            fp64_peak = hb.get_stat("fp64_peak_tflops")
        except Exception:
            fp64_peak = None
        roofline = {
            "kernel": "pairscan_kernel<float,float,FMA> (IVF list-major scan, exact fp64)",
            "bound": "tensor", "achieved": flops / t_scan / 1e12 if single else None, "peak": bf16_peak, "unit": "TFLOP/s",
            "frac": (flops / t_scan / 1e12 / bf16_peak) if single else None, "peak_source": peak_src,
            "traffic": None,
            "launch_ms": t_scan * 1e3, "launches_per_step": scan_n / args.steps,
            "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": unique_bytes,
            "hbm_gbs_at_unique_bytes": unique_bytes / t_scan / 1e9 if single else None,
            "fp64_pipe": {"achieved_tflops": flops / t_scan / 1e12 if single else None, "peak_tflops_measured": fp64_peak,
                          "frac": (flops / t_scan / 1e12 / fp64_peak) if (single and fp64_peak) else None,
                          "note": "EXACT mode: one sequential fp64 FMA chain per (query,row) pair, bound by the fp64 pipe"},
            "step_breakdown_ms": stats,
        }

    # ---- the same index without probe pruning, and a second data distribution on which pruning finds nothing ----------
    unpruned = hard = None
    if fast and world == 1 and not args.no_extras:
        hb.set_mode(hb.MODE_FAST)  # (the recall / fast_vs_exact legs above left the process in EXACT mode)
        unpruned = measure_variant(args, hb, lambda q: gix.search_raw(q, k, nprobe, out_ids=out_ids, out_dist=out_dist), queries, nq,
                                   opts={"fast_prune": 0})
        unpruned["ids_equal_pruned_run"] = bool((out_ids.cpu().numpy() == ids_np).all())
        flops_all = 2.0 * float(np.bincount(asg, minlength=w["nlist"])[probes.reshape(-1)].sum()) * w["d"]
        unpruned["roofline"] = {"bound": "tensor", "achieved": flops_all / (unpruned["tc_ms"] * 1e-3) / 1e12, "peak": bf16_peak,
                                "unit": "TFLOP/s", "frac": flops_all / (unpruned["tc_ms"] * 1e-3) / 1e12 / bf16_peak,
                                "note": "every probed list scanned: the candidate pass is bound by the tensor pipe / its operand traffic"}
        hard = measure_hard_distribution(args, hb, w, device)
        hb.set_mode(hb.MODE_EXACT)
    traffic = None
    if fast and world == 1 and not args.no_traffic:
        traffic = measure_traffic(args)
        if traffic.get("bytes") is not None:
            roofline["traffic"] = traffic["bytes"]
        roofline["traffic_source"] = traffic

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded sample + parity on that sample ---------
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu:
        r = time_cpu_search(rows_np, cents, asg, q_np, k, nprobe, budget_s=20.0)
        cpu = {"value": r["qps"], "unit": "queries/s", "cores": r["cores"], "kind": "port",
               "sample": f"first {r['sample']} of {nq} queries, one pass, {r['cores']} threads (one query per task)"}
        s = r["sample"]
        parity = {"sample_queries": s, "ids_equal": bool((ids_np[:s] == r["ids"]).all()),
                  "dist_bits_equal": bool((dists_np[:s].view(np.int64) == r["dist"].view(np.int64)).all())}

    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if (world > 1 and not replicas) else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(w, key), "recall_at_10": recall,
                   "mode": (f"fast: tcgen05 int8 x{args.digits}-digit candidate pass + fp64 re-score + proof, exact fallback "
                            "(ids and distance bits identical to exact mode)") if fast else "exact (fp64 for every pair)",
                   "l2": "inputs_larger_than_l2 (index slab 3.07 GB vs 126 MB L2)",
                   "sharding": ("single GPU" if world == 1 else
                                f"replicas: the whole index on each of {world} GPUs, one {nq}-query batch per GPU per step, no collective"
                                if replicas else "lists of one global index, l mod N; one batch; all-gather + merge"),
                   "build_s": build_s, "e2e_results_equal_device_results": e2e_same},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "fast_vs_exact": fast_vs_exact, "unpruned": unpruned, "hard_distribution": hard,
    }
    gix.close() if (world == 1 or replicas) else shard.close()
    if world > 1:
        dist.barrier()
    return line


# ---------------------------------------------------------------------------------------------------------
# secondary measurements of the N = 1 line
# ---------------------------------------------------------------------------------------------------------
def measure_variant(args, hb, search, queries, nq, opts):
    """The timed region again (device-resident inputs, CUDA events) with library knobs changed; knobs restored afterwards."""
    import torch

    for name, val in opts.items():
        hb.set_option(name, val)
    try:
        for _ in range(3):
            search(queries)
        torch.cuda.synchronize()
        hb.set_option("profile", 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            search(queries)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out = {"value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "opts": opts,
               "tc_ms": hb.get_stat("tc_ms") / args.steps, "exact_fallbacks_per_step": hb.get_stat("fast_fallbacks") / args.steps,
               "pruned_probe_pairs_per_step": hb.get_stat("fast_pruned_pairs") / args.steps,
               "probe_pairs_per_step": hb.get_stat("fast_probe_pairs") / args.steps}
        hb.set_option("profile", 0)
        return out
    finally:
        for name in opts:
            hb.set_option(name, 1)


def measure_hard_distribution(args, hb, w, device):
    """configs[1]'s shape on data where IVF is not trivially right: i.i.d. Gaussian rows and queries (no clusters at all), so
    recall@10 at nprobe = 32 of 1024 is far below 1, the exact probe pruning can drop next to nothing and the candidate pass
    has no score gap to exploit (queries whose proof fails are recomputed by the exact kernels: `exact_fallbacks_per_step`).
    Same index build, same search call, results checked against EXACT mode."""
    import torch

    from hnsw_clj_b200 import ivf_flat
    from hnsw_clj_b200.flat import FlatIndex, recall_at_k

    wh = dict(w, structureless=True)
    rows, queries = gen_gpu(wh, device)
    k, nprobe, nq = wh["k"], wh["nprobe"], wh["nq"]
    t0 = time.perf_counter()
    ix = ivf_flat.build_index(rows, num_partitions=wh["nlist"], max_iterations=wh["iters"])
    build_s = time.perf_counter() - t0
    ids = torch.empty((nq, k), dtype=torch.int64, device=device)
    dist = torch.empty((nq, k), dtype=torch.float64, device=device)
    out = measure_variant(args, hb, lambda q: ix.search_raw(q, k, nprobe, out_ids=ids, out_dist=dist), queries, nq, opts={})
    f_ids, f_d = ids.cpu().numpy().copy(), dist.cpu().numpy().copy()
    hb.set_mode(hb.MODE_EXACT)
    s = min(nq, 2048)
    x_ids, x_d = ix.search_raw(queries[:s].contiguous(), k, nprobe)
    with FlatIndex(rows) as fx:
        exact_ids, _ = fx.search_raw(queries, k)
    hb.set_mode(hb.MODE_FAST)
    ix.close()
    out.update({"workload": f"ivf-flat {wh['n']}x{wh['d']} fp32 cosine nlist={wh['nlist']} nprobe={nprobe} nq={nq} k={k}, "
                            f"i.i.d. Gaussian rows and queries (structureless)",
                "recall_at_10": recall_at_k(f_ids, exact_ids), "build_s": build_s,
                "fast_equals_exact_mode": {"queries": s, "ids_equal": bool((f_ids[:s] == x_ids).all()),
                                           "dist_bits_equal": bool((f_d[:s].view(np.int64) == x_d.view(np.int64)).all())}})
    del rows
    torch.cuda.empty_cache()
    return out


def measure_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from an ncu pass over a child run of
    this very command (one profiled step, the kernel replayed by ncu; nothing from that run is reported as a timing)."""
    import shutil

    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return {"bytes": None, "why": "ncu not found"}
    kernel = "tc_narrow_kernel" if not any(o.startswith("tc_narrow=0") for o in args.opt) else "tc_pass_kernel"
    cmd = [ncu, "--profile-from-start", "off", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum",
           "--clock-control", "none", "-k", f"regex:{kernel}", "--csv", sys.executable, os.path.abspath(__file__), "--steps", "1",
           "--warmup", "3", "--no-cpu", "--no-extras", "--no-traffic", "--ncu-region", "--workload", args.workload,
           "--digits", str(args.digits)] + [x for o in args.opt for x in ("--opt", o)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    except Exception as e:  # noqa: BLE001
        return {"bytes": None, "why": repr(e)[:200]}
    best = None  # the largest launch of the kernel inside the profiled step = the list scan's main pass
    rows = {}
    for line in r.stdout.splitlines():
        parts = [p.strip('"') for p in line.split('","')]
        if len(parts) < 5 or not parts[0].lstrip('"').isdigit():
            continue
        lid, metric, unit, val = parts[0].lstrip('"'), parts[-3], parts[-2], parts[-1].rstrip('"').replace(",", "")
        try:
            v = float(val)
        except ValueError:
            continue
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(unit, 1.0)
        rows.setdefault(lid, {})[metric] = v * scale
    for lid, m in rows.items():
        if "dram__bytes_read.sum" in m and "dram__bytes_write.sum" in m:
            tot = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
            if best is None or tot > best["bytes"]:
                best = {"bytes": tot, "read": m["dram__bytes_read.sum"], "write": m["dram__bytes_write.sum"],
                        "kernel": kernel, "launches_profiled": len(rows),
                        "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over a child run of this command (1 step)"}
    return best or {"bytes": None, "why": ("no kernel matched; ncu said: " + (r.stderr or r.stdout)[-300:])}


# ---------------------------------------------------------------------------------------------------------
# N > 1: one global index, rows sharded, data plane inside the library
# ---------------------------------------------------------------------------------------------------------
def gen_shard(S, device, rank, world):
    """This rank's rows of the global clustered data set (centres and queries identical on every rank; SURVEY §8d C4:
    generated on the device per shard)."""
    import torch

    d, n, nq = S["d"], S["n_per"], S["nq"]
    ncent = S["centres_per"] * world
    g = torch.Generator(device=device)
    g.manual_seed(42)
    centres = torch.randn((ncent, d), generator=g, device=device)
    g.manual_seed(43)
    idx = torch.randint(0, ncent, (nq,), generator=g, device=device)
    queries = (centres[idx] + S["noise"] * torch.randn((nq, d), generator=g, device=device)).contiguous()
    g.manual_seed(1000 + rank)
    rows = torch.empty((n, d), dtype=torch.float32, device=device)
    step = 131072
    for i in range(0, n, step):
        m = min(step, n - i)
        idx = torch.randint(0, ncent, (m,), generator=g, device=device)
        rows[i:i + m] = centres[idx] + S["noise"] * torch.randn((m, d), generator=g, device=device)
    del centres
    return rows, queries


def draw_seed_rows(n_total, nlist, seed=42):
    """nlist distinct global rows in draw order (seeding of the sharded k-means; see config.seeding)."""
    r = np.random.default_rng(seed)
    out, seen = [], set()
    while len(out) < nlist:
        for v in r.integers(0, n_total, size=2 * nlist).tolist():
            if v not in seen:
                seen.add(v)
                out.append(v)
                if len(out) == nlist:
                    break
    return np.asarray(out, dtype=np.int64)


def run_sharded(args, S, ctx):
    import torch
    import torch.distributed as dist

    from hnsw_clj_b200 import _lib as hb
    from hnsw_clj_b200 import sharded
    from hnsw_clj_b200.flat import recall_at_k
    from oracle import oracle as orc

    rank, world, local, device = ctx
    peaks = load_peaks()
    sharded.comm_init(rank, world)  # the 128-byte id travels over torch.distributed; everything else is the library's
    d, n_per, nq, k, nprobe = S["d"], S["n_per"], S["nq"], S["k"], S["nprobe"]
    nlist, n_total, first_row = S["nlist_per"] * world, S["n_per"] * world, rank * S["n_per"]
    fast = args.mode == "fast"
    t0 = time.perf_counter()
    rows, queries = gen_shard(S, device, rank, world)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return float(sharded.comm_allreduce([x], "max")[0])

    hb.set_mode(hb.MODE_FAST if fast else hb.MODE_EXACT)
    # ---- ground truth for recall@10: exact flat search of a query sample over all shards (sharded flat index, freed again) ----
    T = min(S["truth_queries"], nq)
    t0 = time.perf_counter()
    with sharded.RowShardedFlat(rows, first_row) as fx:
        truth_ids, _ = fx.search_raw(queries[:T], k)
    truth_s = time.perf_counter() - t0
    truth_ids = truth_ids.copy()

    # ---- build (untimed setup; reported): data-parallel k-means + local slabs, inside the library ----------------------
    # seeds: the reference's k-means++ (ivf_flat.clj:32-60, bit-exact device implementation: hb_kmeanspp_init) over the first
    # 8 * nlist rows of rank 0's shard (the data are i.i.d. over the ranks, rank 0 holds global rows 0..n_per), broadcast
    # through the library's communicator.  k-means++ over all N * 12.5M rows would walk every row once per seed.
    hb.set_option("profile", 1)
    barrier()
    t0 = time.perf_counter()
    sample = min(n_per, max(4 * nlist, min(8 * nlist, 262144)))  # seeding cost grows with sample x nlist
    seeds = np.zeros(nlist, dtype=np.int64)
    if args.seeding == "random":
        seeds = draw_seed_rows(n_total, nlist)
    elif rank == 0:
        from hnsw_clj_b200 import ivf_flat

        seeds = np.ascontiguousarray(ivf_flat.kmeanspp_init(rows[:sample], nlist), dtype=np.int64)
    if args.seeding != "random":
        sharded.comm_broadcast(seeds, 0)
    seed_s = time.perf_counter() - t0
    ix = sharded.RowShardedIVFFlat(rows, first_row, nlist, seeds, max_iterations=S["iters"])
    barrier()
    build_s = time.perf_counter() - t0
    build = {"build_s": build_s, "lloyd_rounds": S["iters"], "assign_ms": hb.get_stat("assign_ms"), "update_ms": hb.get_stat("update_ms"),
             "allreduce_ms": hb.get_stat("allreduce_ms"), "allreduce_bytes_per_round": nlist * d * 8 + nlist * 8,
             "seeding_s": seed_s, "kpp_rows_scored": hb.get_stat("kpp_rows_scored"), "kpp_chunks_walked": hb.get_stat("kpp_chunks_walked"),
             "seeding": ("nlist distinct random global rows (numpy default_rng(42))" if args.seeding == "random" else
                         f"k-means++ (ivf_flat.clj:32-60: java.util.Random(42), exact) over the first {sample} rows of rank 0's shard "
                         "(an i.i.d. sample of the data), seeds broadcast; Lloyd rounds over all rows of all ranks")}
    hb.set_option("profile", 0)

    out_ids = torch.empty((nq, k), dtype=torch.int64, device=device)
    out_dist = torch.empty((nq, k), dtype=torch.float64, device=device)
    search = lambda q: ix.search_raw(q, k, nprobe, out_ids=out_ids, out_dist=out_dist)  # noqa: E731

    # ---- timed region: device-resident inputs ------------------------------------------------------------------------
    for _ in range(args.warmup):
        search(queries)
    barrier()
    hb.set_option("profile", 1)
    hb.launch_count(reset=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        search(queries)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    launches = hb.launch_count()
    clk = clocks.stop() if rank == 0 else {}
    names = ("coarse_ms", "plan_ms", "pack_ms", "tc_sample_ms", "tc_ms", "select_ms", "rescore_ms", "scan_ms", "exchange_ms", "merge_ms")
    stats = {n: hb.get_stat(n) / args.steps for n in names}
    tc_ms, tc_n = hb.get_stat("tc_ms"), max(hb.get_stat("tc_count"), 1.0)
    fast_served, fast_fell = hb.get_stat("fast_queries"), hb.get_stat("fast_fallbacks")
    pruned_pairs = hb.get_stat("fast_pruned_pairs") / args.steps
    pruned_rows = hb.get_stat("fast_pruned_rows") / args.steps
    probe_pairs = hb.get_stat("fast_probe_pairs") / args.steps
    tc = {n: hb.get_stat(n) / args.steps for n in ("tc_units", "tc_items", "tc_tiles", "tc_half_units", "tc_narrow_units", "tc_narrow_items",
                                                   "tc_narrow_slots")}
    hb.set_option("profile", 0)
    # the slowest rank's share of the step spent in the exchange + merge (what the collective costs)
    comm_ms = max_over_ranks(stats["exchange_ms"] + stats["merge_ms"])
    value = nq / (ms * 1e-3)
    ids_np, dist_np = out_ids.cpu().numpy(), out_dist.cpu().numpy()

    # ---- e2e: pinned host queries in, host results out, one C-ABI call per rank ---------------------------------------
    hq = torch.empty((nq, d), dtype=torch.float32, pin_memory=True)
    hq.copy_(queries)
    h_ids = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
    h_dist = torch.empty((nq, k), dtype=torch.float64, pin_memory=True)
    for _ in range(max(1, args.warmup // 2)):
        ix.search_raw(hq, k, nprobe, out_ids=h_ids, out_dist=h_dist)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ix.search_raw(hq, k, nprobe, out_ids=h_ids, out_dist=h_dist)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    e2e = {"value": nq / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(hq.numel() * 4) * world, "d2h_bytes_per_step": int(h_ids.numel() * 16) * world,
           "note": "every rank copies the whole query batch in and the merged results out"}
    e2e_same = bool((h_ids.numpy() == ids_np).all())

    # ---- parity: (1) the same search in EXACT mode (fp64 for every pair) on a query sample — collective;
    #              (2) the returned distances recomputed by the oracle's pairwise arithmetic from the rows each rank holds ----
    P = min(256, nq)
    fast_vs_exact = None
    if fast:
        hb.check(hb.lib().hb_index_set_mode(ix._h, hb.MODE_EXACT))
        x_ids, x_dist = ix.search_raw(queries[:P], k, nprobe)
        hb.check(hb.lib().hb_index_set_mode(ix._h, -1))
        fast_vs_exact = {"queries": P, "ids_equal": bool((x_ids == ids_np[:P]).all()),
                         "dist_bits_equal": bool((x_dist.view(np.int64) == dist_np[:P].view(np.int64)).all()),
                         "exact_fallbacks_per_step": fast_fell / max(args.steps + 0, 1), "served_per_step": fast_served / args.steps}
    O = min(64, nq)
    mine = (ids_np[:O] >= first_row) & (ids_np[:O] < first_row + n_per)
    qi, ji = np.nonzero(mine)
    loc = torch.from_numpy(ids_np[:O][mine] - first_row).to(device)
    vec = rows[loc].cpu().numpy()
    q_np = queries[:O].cpu().numpy()
    same = sum(1 for t in range(len(qi))
               if np.float64(orc.cosine_distance_direct(q_np[qi[t]], vec[t])).view(np.int64) == dist_np[qi[t], ji[t]].view(np.int64))
    tot = sharded.comm_allreduce([same, len(qi)], "sum")
    oracle_parity = {"queries": O, "results_checked": int(tot[1]), "dist_bits_equal_oracle": int(tot[0]),
                     "what": "distance bits of the returned (query, row) pairs vs the oracle's pairwise restatement, each rank "
                             "checking the rows it holds"}
    recall = recall_at_k(ids_np[:T], truth_ids)
    info = sharded.comm_info()
    ix.close()
    del rows
    torch.cuda.empty_cache()
    barrier()
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel on this rank (tc_pass_kernel, the candidate pass over the probed lists) --------
    dpad = -(-d // 128) * 128
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = None
    if fast and tc["tc_items"] > 0:
        t_tc = (tc_ms / tc_n) * 1e-3
        slots = 128.0 * (tc["tc_units"] - tc["tc_half_units"]) + 64.0 * (tc["tc_half_units"] - tc["tc_narrow_units"]) + tc["tc_narrow_slots"]
        img_bytes = tc["tc_tiles"] * 128.0 * dpad * args.digits + slots * dpad * args.digits
        roofline = {"kernel": f"tc_narrow_kernel<{args.digits}> ({tc['tc_narrow_items']:.0f} of {tc['tc_items']:.0f} items) + tc_pass_kernel<{args.digits},EMIT> (rank 0's shard)", "bound": "hbm",
                    "achieved": img_bytes / t_tc / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": img_bytes / t_tc / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s", "traffic": None,
                    "launch_ms": t_tc * 1e3, "launches_per_step": tc_n / args.steps, "algorithmic_bytes_per_launch": img_bytes,
                    "units_per_launch": tc["tc_units"], "items_per_launch": tc["tc_items"], "row_tiles_read_per_launch": tc["tc_tiles"],
                    "probe_pruning": {"probe_pairs": probe_pairs, "pruned_probe_pairs": pruned_pairs, "pruned_rows": pruned_rows},
                    "step_breakdown_ms": stats}
    exchange_bytes = nq * k * 16
    line = {
        "metric": METRIC_SHARDED, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"ivf-flat {n_total}x{d} fp32 cosine ({world} x {n_per} rows, contiguous row blocks) nlist={nlist} "
                               f"nprobe={nprobe} nq={nq} k={k}" + (" (100M x 768: BASELINE metric, configs[3]'s index)" if n_total == 100_000_000 else ""),
                   "recall_at_10": recall, "recall_queries": T,
                   "mode": (f"fast: tcgen05 int8 x{args.digits}-digit candidate pass + fp64 re-score + proof, exact fallback") if fast else "exact",
                   "l2": f"inputs_larger_than_l2 (slab {n_per * d * 4 / 1e9:.1f} GB per GPU vs 126 MB L2)",
                   "sharding": ("rows: one global index, rank g holds rows [g*n_per, (g+1)*n_per) of every list; centroids replicated; "
                                "per search one collective step: " +
                                ("peer-window exchange fused with the merge (exchange_merge_kernel<P2P>: stores to cudaIpc-mapped peer "
                                 "memory over NVLink, release/acquire flags, merge)" if info["p2p"] else
                                 "ncclAllGather of the packed local top-k + merge kernel")),
                   "collective": {"kind": "p2p-window" if info["p2p"] else "nccl-allgather", "bytes_per_rank_per_step": exchange_bytes,
                                  "exchange_ms": stats["exchange_ms"], "merge_ms": stats["merge_ms"],
                                  "allgather_ms": None if info["p2p"] else stats["exchange_ms"],
                                  "slowest_rank_exchange_plus_merge_ms": comm_ms,
                                  "note": "p2p-window: the exchange happens inside the merge kernel (merge_ms covers push + wait + merge, "
                                          "exchange_ms is 0); wait time includes the skew between ranks"},
                   "build": build, "gen_s": gen_s, "truth_s": truth_s, "e2e_results_equal_device_results": e2e_same},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": None,
        "parity": oracle_parity, "fast_vs_exact": fast_vs_exact,
    }
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: skip the unpruned / hard-distribution secondary measurements")
    ap.add_argument("--no-traffic", action="store_true", help="N = 1: skip the ncu child run that measures roofline.traffic")
    ap.add_argument("--shard", default="rows", choices=["rows", "replicas", "lists"],
                    help="N > 1: rows (ONE global index of N x 12.5M rows, contiguous row blocks, data plane inside the library: "
                         "hb_sharded_ivf_build / hb_sharded_search; weak scaling), replicas (whole configs[1] index per GPU, one "
                         "batch per GPU, no collective) or lists (configs[1], lists sharded l mod N, host-orchestrated all-gather + merge)")
    ap.add_argument("--shard-workload", default="c100m", choices=["c100m", "small"],
                    help="--shard rows: c100m = 12.5M rows per GPU (100M x 768 at N = 8), small = 0.5M rows per GPU (smoke runs)")
    ap.add_argument("--no-replicas", action="store_true", help="--shard rows: skip the secondary replicas measurement")
    ap.add_argument("--seeding", default="kpp-sample", choices=["kpp-sample", "random"],
                    help="--shard rows: seeds of the sharded k-means (k-means++ over a sample held by rank 0, or random rows)")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"],
                    help="fast: tensor-core candidate pass + fp64 re-score + proof (same results); exact: fp64 for every pair")
    ap.add_argument("--opt", action="append", default=[], help="library knob name=value (hb_set_option), repeatable")
    ap.add_argument("--ncu-region", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps (for `ncu --profile-from-start off`)")
    ap.add_argument("--digits", type=int, default=2, choices=[2, 3], help="int8 digits per element in fast mode")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, args.workload)
        return
    ctx = setup(args)
    rank, world = ctx[0], ctx[1]
    if world > 1 and args.shard == "rows":
        line = run_sharded(args, SHARD if args.shard_workload == "c100m" else SHARD_SMALL, ctx)
        if not args.no_replicas:
            rep = run_ours(args, w, args.workload, ctx, replicas_only=True)
            if rank == 0:
                line["replicas"] = rep
    else:
        line = run_ours(args, w, args.workload, ctx)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

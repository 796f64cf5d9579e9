"""Measurements of the BASELINE.json configs that are not the bench.py headline (configs[1]):

    python tools/bench_configs.py c1            # 31,173 x 768 fp32 cosine, 1000 queries, flat top-10 (configs[0])
    python tools/bench_configs.py c3 [--n N]    # flat N x 768 bf16 inner product, 4096 queries, top-100 (configs[2], one GPU's rows)
    python tools/bench_configs.py c4 [--n N]    # one Lloyd round (assign + update) on one GPU's shard of configs[3]
    torchrun ... tools/bench_configs.py c4full  # configs[3] as named: 100 M rows over the GPUs of the box, 10 Lloyd rounds
    python tools/bench_configs.py c5 [--n N]    # batched HNSW search, M=16 ef=128, 16k queries (configs[4]; --n 1000000 --graph bulk = as named)

Every command prints ONE JSON line (device-resident inputs, CUDA events on the launching stream, inputs larger than L2
or stated otherwise) and checks the device results against the exact mode / the CPU oracle on a bounded sample.
The oracle is only the checker here, never the thing measured -- except in the explicitly named cpu_baseline field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def timed(fn, reps=3, warm=1):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def same_bits(a, b):
    return bool((np.asarray(a).view(np.int64) == np.asarray(b).view(np.int64)).all())


def unit_rows(n, d, seed, device):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.randn((n, d), generator=g, device=device)
    return x / x.norm(dim=1, keepdim=True)


def run_c1(args):
    import torch

    from hnsw_clj_b200 import _lib as hb
    from hnsw_clj_b200.flat import FlatIndex
    from oracle import oracle as orc

    dev = torch.device("cuda", 0)
    n, d, nq, k = 31173, 768, 1000, 10
    rows = unit_rows(n, d, 42, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(43)
    queries = (rows[torch.arange(nq, device=dev) * 31] + 0.1 / d ** 0.5 * torch.randn((nq, d), generator=g, device=dev)).contiguous()
    out_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    out_d = torch.empty((nq, k), dtype=torch.float64, device=dev)
    res = {}
    with FlatIndex(rows) as ix:
        for mode, code in (("exact", hb.MODE_EXACT), ("fast", hb.MODE_FAST)):
            hb.set_mode(code)
            hb.set_option("profile", 1)
            ms = timed(lambda: ix.search_raw(queries, k, out_ids=out_ids, out_dist=out_d), reps=5, warm=2)
            res[mode] = {"ms": ms, "qps": nq / ms * 1e3, "ids": out_ids.cpu().numpy().copy(), "dist": out_d.cpu().numpy().copy(),
                         "fallbacks": hb.get_stat("fast_fallbacks") if mode == "fast" else None}
            hb.set_option("profile", 0)
        hb.set_mode(hb.MODE_EXACT)
    rows_np, q_np = rows.cpu().numpy(), queries.cpu().numpy()
    s = 128
    t0 = time.perf_counter()
    want_ids, want_d = orc.exact_knn(rows_np, q_np[:s], k)
    cpu_s = time.perf_counter() - t0
    flops = 2.0 * nq * n * d
    pk = peaks()
    line = {
        "config": "BASELINE configs[0]: 31,173x768 fp32 cosine, 1000 queries, exact flat top-10 (unit-norm Gaussian rows, queries = rows + noise)",
        "metric": "queries/s", "value_fast": res["fast"]["qps"], "value_exact": res["exact"]["qps"],
        "ms_fast": res["fast"]["ms"], "ms_exact": res["exact"]["ms"],
        "fast_equals_exact": bool((res["fast"]["ids"] == res["exact"]["ids"]).all()) and same_bits(res["fast"]["dist"], res["exact"]["dist"]),
        "fast_fallbacks_per_7_calls": res["fast"]["fallbacks"],
        "parity_vs_oracle": {"queries": s, "ids_equal": bool((res["exact"]["ids"][:s] == want_ids).all()),
                             "dist_bits_equal": same_bits(res["exact"]["dist"][:s], want_d)},
        "cpu_baseline": {"value": s / cpu_s, "unit": "queries/s", "cores": orc.ncores(), "kind": "port",
                         "sample": f"first {s} queries, all host threads"},
        "roofline": {"bound": "tensor", "algorithmic_flops": flops, "achieved_tflops_fast": flops / res["fast"]["ms"] / 1e9,
                     "achieved_tflops_exact_fp64": flops / res["exact"]["ms"] / 1e9, "peak_bf16_tflops": pk.get("bf16_tflops"),
                     "note": "index (96 MB) fits L2: not an HBM measurement"},
    }
    print(json.dumps(line), flush=True)


def run_c3(args):
    """One GPU: the whole matrix.  Under torchrun (WORLD_SIZE = G): rows sharded [g*N/G, (g+1)*N/G), queries replicated,
    local top-100 -> NCCL all-gather -> merge kernel (hnsw_clj_b200/sharded.py: ShardedFlat); strong scaling."""
    import torch

    from hnsw_clj_b200 import _lib as hb
    from hnsw_clj_b200.sharded import ShardedFlat, row_range

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    hb.check(hb.lib().hb_init(local))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n, d, nq, k = args.n or 10_000_000, 768, 4096, 100
    lo, hi = row_range(n, rank, world)
    g = torch.Generator(device=dev)
    g.manual_seed(42 + rank)
    rows = torch.empty((hi - lo, d), dtype=torch.bfloat16, device=dev)
    for i in range(0, hi - lo, 1 << 20):
        m = min(1 << 20, hi - lo - i)
        rows[i:i + m] = torch.randn((m, d), generator=g, device=dev).to(torch.bfloat16)
    g.manual_seed(43)
    queries = torch.randn((nq, d), generator=g, device=dev).to(torch.bfloat16).float().contiguous()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sh = ShardedFlat(rows, lo, rank, world, "ip")
    del rows
    torch.cuda.synchronize()
    create_s = time.perf_counter() - t0
    hb.set_mode(hb.MODE_FAST)
    t0 = time.perf_counter()
    sh.search(queries, k)  # first call quantises the rows (digit images)
    torch.cuda.synchronize()
    first_s = time.perf_counter() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    hb.set_option("profile", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        out_ids, out_d = sh.search(queries, k)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.reps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    stats = {nm: hb.get_stat(nm) / args.reps for nm in ("tc_ms", "tc_sample_ms", "select_ms", "pack_ms", "rescore_ms")}
    served, fell = hb.get_stat("fast_queries"), hb.get_stat("fast_fallbacks")
    hb.set_option("profile", 0)
    fast_ids, fast_d = out_ids.cpu().numpy().copy(), out_d.cpu().numpy().copy()
    # the exact mode (fp64 for every pair) on a sample of the queries, through the same sharded path
    hb.set_mode(hb.MODE_EXACT)
    s = 64
    qs = queries[:s].contiguous()
    e_ids, e_d = sh.search(qs, k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sh.search(qs, k)
    torch.cuda.synchronize()
    ms_exact = (time.perf_counter() - t0) * 1e3
    e_ids, e_d = e_ids.cpu().numpy(), e_d.cpu().numpy()
    info = sh.index.info()
    sh.close()
    if rank == 0:
        flops = 2.0 * nq * n * d
        pk = peaks()
        bf16 = pk.get("bf16_tflops", 1590.0)
        line = {
            "config": f"BASELINE configs[2]: flat {n}x{d} bf16 inner product, {nq} queries, top-{k} (Gaussian rows), rows sharded over {world} GPU(s)",
            "metric": "queries/s", "value": nq / ms * 1e3, "ms_per_batch": ms, "n_gpus": world, "scaling": "strong",
            "mode": "fast (int8 x2-digit tcgen05 candidate pass, 128 re-scored, proof)" + ("; NCCL all-gather + merge kernel" if world > 1 else ""),
            "exact_mode_qps_on_sample": s / ms_exact * 1e3,
            "parity": {"sample_queries": s, "ids_equal": bool((fast_ids[:s] == e_ids).all()), "dist_bits_equal": same_bits(fast_d[:s], e_d)},
            "fast_queries_rank0": served, "fast_fallbacks_rank0": fell, "index_create_s": create_s, "first_call_s": first_s,
            "device_bytes_rank0": info["device_bytes"], "step_breakdown_ms_rank0": stats,
            "roofline": {"bound": "tensor", "algorithmic_flops": flops, "achieved_tflops": flops / ms / 1e9,
                         "peak_bf16_tflops_per_gpu": bf16, "frac": flops / ms / 1e9 / (bf16 * world),
                         "algorithmic_bytes": float(n) * d * 2, "hbm_gbs_at_unique_bytes": n * d * 2 / ms / 1e6},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c4(args):
    import torch

    from hnsw_clj_b200 import _lib as hb

    dev = torch.device("cuda", 0)
    n, d, nlist = args.n or 12_500_000, 768, args.nlist
    L = hb.lib()
    hb.check(L.hb_init(0))
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    ncent = max(nlist // 2, 8)
    centres = torch.randn((ncent, d), generator=g, device=dev)
    rows = torch.empty((n, d), dtype=torch.float32, device=dev)
    for i in range(0, n, 1 << 19):
        m = min(1 << 19, n - i)
        idx = torch.randint(0, ncent, (m,), generator=g, device=dev)
        rows[i:i + m] = centres[idx] + 0.1 * torch.randn((m, d), generator=g, device=dev)
    del centres
    seeds = torch.randperm(n, generator=g, device=dev)[:nlist]
    cents = rows[seeds].double().contiguous()  # seeds = data rows, as k-means++ would hand over (ivf_flat.clj:37-41)
    asg = torch.empty(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def assign(r, out):
        hb.check(L.hb_kmeans_assign(r.data_ptr(), r.shape[0], d, hb.F32, 0, cents.data_ptr(), nlist, out.data_ptr()))

    def update():
        hb.check(L.hb_kmeans_update(rows.data_ptr(), n, d, hb.F32, asg.data_ptr(), nlist, cents.data_ptr(), None, None))

    hb.set_mode(hb.MODE_FAST)
    hb.set_option("profile", 1)
    ms_assign = timed(lambda: assign(rows, asg), reps=args.reps, warm=1)
    served, fell = hb.get_stat("fast_queries"), hb.get_stat("fast_fallbacks")
    hb.set_option("profile", 0)
    # exact mode on a sample of the rows
    hb.set_mode(hb.MODE_EXACT)
    s = min(n, 16384)
    sample = rows[:s].contiguous()
    e_asg = torch.empty(s, dtype=torch.int32, device=dev)
    ms_exact = timed(lambda: assign(sample, e_asg), reps=1, warm=0)
    equal = bool((e_asg == asg[:s]).all().item())
    c_before = cents.clone()
    ms_update = timed(update, reps=1, warm=1)
    moved = float((cents - c_before).abs().max().item())
    pairs = float(n) * nlist
    flops = 2.0 * pairs * d
    pk = peaks()
    bf16 = pk.get("bf16_tflops", 1590.0)
    line = {
        "config": f"BASELINE configs[3], one GPU's shard: k-means on {n}x{d} fp32, nlist={nlist}: one Lloyd round = assign + update "
                  "(seeds given; configs[3] is 100M rows over 8 GPUs = 12.5M per GPU)",
        "assign_ms": ms_assign, "update_ms": ms_update, "rows_per_s_assign": n / ms_assign * 1e3,
        "assign_mode": "fast (k=1 candidate pass over the centroids on tcgen05 + fp64 re-score + proof)",
        "fast_rows": served, "fast_fallbacks": fell,
        "exact_assign_ms_on_sample": ms_exact, "exact_rows_per_s": s / ms_exact * 1e3,
        "parity": {"sample_rows": s, "assignments_equal_exact_mode": equal},
        "centroids_moved_max_abs": moved,
        "roofline": {"bound": "tensor", "algorithmic_flops": flops, "achieved_tflops": flops / ms_assign / 1e9, "peak_bf16_tflops": bf16,
                     "frac": flops / ms_assign / 1e9 / bf16,
                     "update_hbm_gbs": float(n) * d * 4 / ms_update / 1e6, "hbm_peak_gbs": pk.get("hbm_gbs")},
    }
    print(json.dumps(line), flush=True)


def run_c4_full(args):
    """BASELINE configs[3] as named: k-means on N x 768 fp32 (default 100 M), nlist 65,536, `--iters` Lloyd rounds + the final
    assignment, rows sharded over the GPUs of the box (torchrun).  Seeds are given (rows of rank 0's shard): the
    reference's k-means++ is a sequential O(N) prefix walk per seed and cannot be run at this nlist."""
    import torch
    import torch.distributed as dist

    from hnsw_clj_b200 import _lib as hb
    from hnsw_clj_b200.sharded import lloyd_round_device, row_range

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n, d, nlist, iters = args.n or 100_000_000, 768, args.nlist, args.iters
    lo, hi = row_range(n, rank, world)
    nl = hi - lo
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    ncent = max(nlist // 2, 8)
    centres = torch.randn((ncent, d), generator=g, device=dev)  # same centres on every rank
    g.manual_seed(1000 + rank)
    rows = torch.empty((nl, d), dtype=torch.float32, device=dev)
    for i in range(0, nl, 1 << 19):
        m = min(1 << 19, nl - i)
        idx = torch.randint(0, ncent, (m,), generator=g, device=dev)
        rows[i:i + m] = centres[idx] + 0.1 * torch.randn((m, d), generator=g, device=dev)
    del centres
    cents = torch.empty((nlist, d), dtype=torch.float64, device=dev)
    if rank == 0:
        cents.copy_(rows[torch.randperm(nl, generator=g, device=dev)[:nlist]].double())
    if world > 1:
        dist.broadcast(cents, 0)
    asg = torch.empty(nl, dtype=torch.int32, device=dev)
    hb.set_mode(hb.MODE_FAST)
    hb.set_option("profile", 1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    per_round = []
    for _ in range(iters):
        per_round.append(lloyd_round_device(rows, cents, asg, world=world))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hb.check(hb.lib().hb_kmeans_assign(rows.data_ptr(), nl, d, hb.F32, hb.COSINE, cents.data_ptr(), nlist, asg.data_ptr()))
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_s = time.perf_counter() - t0
    served, fell = hb.get_stat("fast_queries"), hb.get_stat("fast_fallbacks")
    hb.set_option("profile", 0)
    tt = torch.tensor([total_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_s = float(tt.item())
    # exact mode on a sample of this rank's rows against the final centroids
    hb.set_mode(hb.MODE_EXACT)
    s = min(nl, 8192)
    sample = rows[:s].contiguous()
    e_asg = torch.empty(s, dtype=torch.int32, device=dev)
    hb.check(hb.lib().hb_kmeans_assign(sample.data_ptr(), s, d, hb.F32, hb.COSINE, cents.data_ptr(), nlist, e_asg.data_ptr()))
    equal = torch.tensor([int((e_asg == asg[:s]).all().item())], device=dev)
    csum = cents.sum(dtype=torch.float64).reshape(1).clone()
    if world > 1:
        dist.all_reduce(equal, op=dist.ReduceOp.MIN)
        gathered = [torch.empty_like(csum) for _ in range(world)]
        dist.all_gather(gathered, csum)
        same_c = all(bool((x == gathered[0]).all().item()) for x in gathered)
    else:
        same_c = True
    if rank == 0:
        a_ms = [r[0] for r in per_round]
        u_ms = [r[1] for r in per_round]
        r_ms = [r[2] for r in per_round]
        flops = 2.0 * n * nlist * d
        pk = peaks()
        bf16 = pk.get("bf16_tflops", 1590.0)
        mean_a = sum(a_ms) / max(len(a_ms), 1)
        line = {
            "config": f"BASELINE configs[3]: IVF k-means build {n}x{d} fp32, nlist={nlist}, {iters} Lloyd rounds + final assignment, "
                      f"rows sharded over {world} GPU(s) ({nl} rows = {nl * d * 4 / 1e9:.1f} GB per GPU); seeds given",
            "metric": "seconds for the Lloyd rounds + final assignment", "value": total_s, "n_gpus": world, "higher_is_better": False,
            "assign_ms_per_round_rank0": a_ms, "update_ms_per_round_rank0": u_ms, "allreduce_divide_ms_per_round_rank0": r_ms,
            "final_assign_ms_rank0": e0.elapsed_time(e1),
            "assign_mode": "fast (k=1 candidate pass over the centroids on tcgen05 + fp64 re-score + proof)",
            "fast_rows_rank0": served, "fast_fallbacks_rank0": fell,
            "parity": {"sample_rows_per_rank": s, "final_assignments_equal_exact_mode_on_every_rank": bool(equal.item()),
                       "centroids_identical_on_every_rank": same_c},
            "roofline": {"bound": "tensor", "algorithmic_flops_per_assign_pass": flops,
                         "achieved_tflops_assign": flops / mean_a / 1e9 if a_ms else None, "peak_bf16_tflops_per_gpu": bf16,
                         "frac": (flops / mean_a / 1e9 / (bf16 * world)) if a_ms else None,
                         "allreduce_bytes_per_round": nlist * d * 8 + nlist * 8},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c5(args):
    import torch

    from hnsw_clj_b200 import _lib as hb
    from oracle import oracle as orc

    dev = torch.device("cuda", 0)
    n, d, nq, k, ef = args.n or 20000, 768, 16384, 10, 128
    g = torch.Generator(device=dev)
    if args.data == "gaussian":
        # structureless: i.i.d. Gaussian directions.  Every graph index struggles here (the reference's own insert-built graph
        # reaches the same recall as the bulk graph, profiles/r01c_config4_hnsw_20k.json.log)
        rows = unit_rows(n, d, 42, dev)
        data_desc = "unit-norm i.i.d. Gaussian (structureless)"
    else:
        # unit-norm rows with cluster structure (cf. :clustered, test/data_generator.clj:74-79): n / 500 centres, per-dimension
        # noise args.noise, then normalised — embeddings-like data on which a navigable graph is meaningful
        g.manual_seed(42)
        ncent = max(8, n // 500)
        centres = torch.randn((ncent, d), generator=g, device=dev)
        rows = torch.empty((n, d), dtype=torch.float32, device=dev)
        for i in range(0, n, 1 << 18):
            m = min(1 << 18, n - i)
            idx = torch.randint(0, ncent, (m,), generator=g, device=dev)
            rows[i:i + m] = centres[idx] + args.noise * torch.randn((m, d), generator=g, device=dev)
        rows = rows / rows.norm(dim=1, keepdim=True)
        data_desc = f"unit-norm clustered ({ncent} Gaussian centres, noise {args.noise} per dimension)"
    g.manual_seed(43)
    queries = (rows[torch.randint(0, n, (nq,), generator=g, device=dev)] +
               0.3 / d ** 0.5 * torch.randn((nq, d), generator=g, device=dev)).contiguous()
    rows_np = rows.cpu().numpy()
    from hnsw_clj_b200.ultra_fast import HnswIndex, bulk_knn_graph

    t0 = time.perf_counter()
    if args.graph == "oracle":
        # the reference's incremental insert-single, restated on the host (feasible up to a few 10k nodes)
        graph = orc.Hnsw(rows_np, M=16, ef_construction=200, level_seed=42)
        levels, entry = graph.levels(), graph.entry
        adjacency = [graph.export_level(l) for l in range(graph.max_level + 1)]
        how = "graph built on the host by the oracle (insert-single, efConstruction=200)"
    else:
        # bulk k-NN graph built on the device by the flat search (FAST mode): configs[4]'s 1M nodes in seconds
        hb.set_mode(hb.MODE_FAST)
        levels, entry, adjacency = bulk_knn_graph(rows, M=16, level_seed=42)
        hb.set_mode(hb.MODE_EXACT)
        graph = orc.Hnsw.from_graph(rows_np, levels, entry, adjacency)
        how = "bulk k-NN graph (32 nearest per node at level 0, 16 above, reference level distribution) built on the device by the flat search"
    build_s = time.perf_counter() - t0
    ix = HnswIndex(rows, levels, entry, adjacency, distance_fn="cosine")
    out_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    out_d = torch.empty((nq, k), dtype=torch.float64, device=dev)
    hb.set_option("profile", 1)
    ms = timed(lambda: ix.search_raw(queries, k, ef, out_ids=out_ids, out_dist=out_d), reps=args.reps, warm=1)
    scored = hb.get_stat("hnsw_scored") / (args.reps + 1)
    hb.set_option("profile", 0)
    ids = out_ids.cpu().numpy()
    s = 256
    q_np = queries[:s].cpu().numpy()
    t0 = time.perf_counter()
    want_ids, want_d = graph.search(q_np, k, ef)
    cpu_s = time.perf_counter() - t0
    from hnsw_clj_b200.flat import FlatIndex, recall_at_k

    with FlatIndex(rows) as fx:
        ex_ids, _ = fx.search_raw(queries, k)
    if hasattr(ex_ids, "cpu"):
        ex_ids = ex_ids.cpu().numpy()
    pk = peaks()
    bytes_scored = scored * (d * 4 + 4)
    line = {
        "config": f"BASELINE configs[4]{'' if n >= 1000000 else ' on a reduced graph'}: HNSW M=16 efSearch={ef}, {n}x{d} fp32 {data_desc}, {nq} concurrent queries, top-{k} "
                  f"({how}, {build_s:.0f} s)",
        "metric": "queries/s", "value": nq / ms * 1e3, "ms_per_batch": ms, "recall_at_10": recall_at_k(ids, ex_ids),
        "pairs_scored_per_batch": scored,
        "parity": {"sample_queries": s, "ids_equal_oracle_traversal": bool((ids[:s] == want_ids).all()),
                   "dist_bits_equal": same_bits(out_d.cpu().numpy()[:s], want_d)},
        "cpu_baseline": {"value": s / cpu_s, "unit": "queries/s", "cores": 1, "kind": "port", "sample": f"first {s} queries, one thread"},
        "roofline": {"bound": "hbm", "achieved": bytes_scored / ms / 1e6, "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                     "frac": bytes_scored / ms / 1e6 / pk.get("hbm_gbs", 6500.0), "algorithmic_bytes_per_scored_candidate": d * 4 + 4,
                     "note": (f"the {n * d * 4 / 1e6:.0f} MB of vectors fit L2 at this graph size: the gather is L2-bound here, not HBM-bound"
                              if n * d * 4 < 120e6 else f"random {d * 4}-byte row gathers over {n * d * 4 / 1e9:.2f} GB of vectors")},
        "graph": {"max_level": int(levels.max()), "edges_level0": int(adjacency[0][0][-1])},
    }
    print(json.dumps(line), flush=True)
    if args.sweep:
        name, _, vals = args.sweep.partition("=")
        for v in vals.split(","):
            hb.set_option(name, int(v))
            hb.set_option("profile", 1)
            ms_v = timed(lambda: ix.search_raw(queries, k, ef, out_ids=out_ids, out_dist=out_d), reps=args.reps, warm=1)
            same = bool((out_ids.cpu().numpy() == ids).all())
            print(json.dumps({"sweep": name, "value": int(v), "ms_per_batch": ms_v, "queries_per_s": nq / ms_v * 1e3,
                              "gather_gbs": bytes_scored / ms_v / 1e6, "ids_unchanged": same,
                              "queue_overflows_per_batch": hb.get_stat("hnsw_overflows") / (args.reps + 1)}), flush=True)
            hb.set_option("profile", 0)


def run_lsh(args):
    """Hybrid LSH (src/hnsw/ann/hash/hybrid_lsh.clj; reference README: 0.5-1 s build, 2-5 ms per query, ~45 % recall) on
    configs[0]'s shape: 31,173 x 768 unit-norm rows, 1000 noisy queries, search-knn's default mode (6 tables, radius 2)."""
    import torch

    from hnsw_clj_b200 import hybrid_lsh
    from hnsw_clj_b200.flat import FlatIndex
    from oracle import oracle as orc

    dev = torch.device("cuda", 0)
    n, d, nq, k = args.n or 31173, 768, 1000, 10
    rows = unit_rows(n, d, 42, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(43)
    queries = (rows[torch.arange(nq, device=dev) * (n // nq)] + 0.1 / d ** 0.5 * torch.randn((nq, d), generator=g, device=dev)).contiguous()
    rows_np, q_np = rows.cpu().numpy(), queries.cpu().numpy()
    t0 = time.perf_counter()
    ix = hybrid_lsh.build_index(rows_np)
    build_s = time.perf_counter() - t0
    res = {}
    for mode in ("turbo", "balanced", "precise"):
        p, r = hybrid_lsh._MODES[mode]
        hybrid_lsh.search_hybrid_multiprobe_raw(ix, q_np, k, p, r)
        t0 = time.perf_counter()
        for _ in range(args.reps):
            ids, dist = hybrid_lsh.search_hybrid_multiprobe_raw(ix, q_np, k, p, r)
        ms = (time.perf_counter() - t0) * 1e3 / args.reps
        res[mode] = (ms, ids, dist)
    with FlatIndex(rows) as fx:
        exact_ids, _ = fx.search_raw(queries, k)
    s = 64
    t0 = time.perf_counter()
    want_ids, want_d = orc.lsh_search(rows_np, orc.lsh_matrices(d), ix.buckets, q_np[:s], k, 6, 2, True, 2)
    cpu_s = time.perf_counter() - t0
    info = hybrid_lsh.index_info(ix)
    ix.close()
    line = {
        "config": f"Hybrid LSH {n}x768 fp32 cosine, 8 tables x 12 bits, {nq} queries per batch, top-10 (host buffers, host-side bucket tables)",
        "build_s": build_s, "index_info": info,
        "modes": {m: {"ms_per_batch": v[0], "queries_per_s": nq / v[0] * 1e3,
                      "recall_at_10": orc.recall(v[1], exact_ids)} for m, v in res.items()},
        "parity_vs_oracle": {"queries": s, "mode": "balanced", "ids_equal": bool((res["balanced"][1][:s] == want_ids).all()),
                             "dist_bits_equal": same_bits(res["balanced"][2][:s], want_d)},
        "cpu_baseline": {"value": s / cpu_s, "unit": "queries/s", "cores": 1, "kind": "port",
                         "sample": f"first {s} queries, one thread, balanced mode"},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c1", "c3", "c4", "c4full", "c5", "lsh"])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--nlist", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--graph", choices=["oracle", "bulk"], default="bulk")
    ap.add_argument("--data", choices=["gaussian", "clustered"], default="clustered", help="c5: row distribution")
    ap.add_argument("--noise", type=float, default=0.5, help="c5 --data clustered: per-dimension noise around the centres")
    ap.add_argument("--opt", action="append", default=[], help="library knob name=value (hb_set_option), repeatable")
    ap.add_argument("--sweep", default="", help="c5: name=v1,v2,... re-times the search for each value of a library knob")
    args = ap.parse_args()
    from hnsw_clj_b200 import _lib as hb

    hb.check(hb.lib().hb_init(int(os.environ.get("LOCAL_RANK", "0"))))
    for kv in args.opt:
        name, _, val = kv.partition("=")
        hb.set_option(name, int(val))
    {"c1": run_c1, "c3": run_c3, "c4": run_c4, "c4full": run_c4_full, "c5": run_c5, "lsh": run_lsh}[args.config](args)


if __name__ == "__main__":
    main()

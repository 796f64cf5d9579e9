"""Small-batch latency: `search-knn` with 1 / 8 / 64 queries per call (the reference's own calling pattern: one query per
call, `src/hnsw/ann/partition/ivf_flat.clj:300-317`, `src/hnsw/bench.clj:72-84`) through the C ABI with HOST buffers.
Prints one JSON line per (index, mode, batch).  Every result is compared with the same rows of a large-batch exact call."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from hnsw_clj_b200 import _lib as hb
    from hnsw_clj_b200 import ivf_flat
    from hnsw_clj_b200.flat import FlatIndex

    ap = argparse.ArgumentParser()
    ap.add_argument("--n-ivf", type=int, default=1000000)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--batches", default="1,8,16,32,64")
    ap.add_argument("--indexes", default="flat31k,flat1m,flat1mc,ivf")
    ap.add_argument("--modes", default="exact,fast")
    ap.add_argument("--grid", default="", help="name=v1,v2;name2=... : repeat every sweep for each combination of library knobs")
    ap.add_argument("--ncu-region", action="store_true", help="cudaProfilerStart/Stop around ONE call per (index, mode, batch)")
    args = ap.parse_args()
    hb.check(hb.lib().hb_init(0))
    for o in args.opt:
        name, v = o.split("=")
        hb.set_option(name, int(v))
    dev = torch.device("cuda", 0)
    d, k = 768, 10
    g = torch.Generator(device=dev)
    g.manual_seed(42)

    def lat(ix, q_host, nq, nprobe):
        ids = np.empty((nq, k), dtype=np.int64)
        dist = np.empty((nq, k), dtype=np.float64)
        q = np.ascontiguousarray(q_host[:nq])

        def call():
            hb.check(hb.lib().hb_search(ix._h, q.ctypes.data, hb.F32, nq, k, nprobe, ids.ctypes.data, dist.ctypes.data))

        for _ in range(5):
            call()
        if args.ncu_region:
            hb.set_option("cuda_profiler", 1)
            call()
            hb.set_option("cuda_profiler", 0)
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        return ts[len(ts) // 2] * 1e6, ts[0] * 1e6, ids, dist

    def sweep(name, ix, q_host, nprobe, unique_bytes):
        hb.set_mode(hb.MODE_EXACT)
        nbig = q_host.shape[0]
        want_ids = np.empty((nbig, k), dtype=np.int64)
        want_d = np.empty((nbig, k), dtype=np.float64)
        hb.check(hb.lib().hb_search(ix._h, q_host.ctypes.data, hb.F32, nbig, k, nprobe, want_ids.ctypes.data, want_d.ctypes.data))
        import itertools
        axes = [(a.split("=")[0], [int(v) for v in a.split("=")[1].split(",")]) for a in args.grid.split(";") if a]
        combos = list(itertools.product(*[[(nm, v) for v in vals] for nm, vals in axes])) or [()]
        for combo, (mode, code) in itertools.product(combos, (("exact", hb.MODE_EXACT), ("fast", hb.MODE_FAST))):
            if mode not in args.modes.split(","):
                continue
            for nm, v in combo:
                hb.set_option(nm, v)
            hb.set_mode(code)
            for nq in [int(x) for x in args.batches.split(',')]:
                med, best, ids, dist = lat(ix, q_host, nq, nprobe)
                calls = 10  # a second, profiled loop: CUDA events around the scan and select kernels of each call
                qd = torch.from_numpy(np.ascontiguousarray(q_host[:nq])).to(dev)
                hb.set_option("profile", 1)
                for _ in range(calls):
                    ix.search_raw(qd, k, nprobe) if nprobe else ix.search_raw(qd, k)
                scan_ms, sel_ms = hb.get_stat("scan_ms"), hb.get_stat("select_ms")
                served, fell = hb.get_stat("fast_queries"), hb.get_stat("fast_fallbacks")
                hb.set_option("profile", 0)
                same = bool((ids == want_ids[:nq]).all() and (dist.view(np.int64) == want_d[:nq].view(np.int64)).all())
                print(json.dumps({"index": name, "mode": mode, "knobs": dict(combo), "queries_per_call": nq, "median_us": med, "best_us": best,
                                  "qps": nq / med * 1e6, "equals_large_batch_exact": same,
                                  "fast_served_per_call": served / calls, "exact_fallbacks_per_call": fell / calls,
                                  "hbm_floor_us": unique_bytes(nq) / 6562.6e3,
                                                                    "scan_us_per_call": scan_ms / calls * 1e3, "select_us_per_call": sel_ms / calls * 1e3,
                                  "scan_hbm_gbs": unique_bytes(nq) / (scan_ms / calls * 1e-3) / 1e9 if scan_ms > 0 else None,
                                  "scan_hbm_frac": unique_bytes(nq) / (scan_ms / calls * 1e-3) / 1e9 / 6562.6 if scan_ms > 0 else None}),
                      flush=True)
        hb.set_mode(hb.MODE_EXACT)

    # configs[0]: flat 31,173 x 768
    n = 31173
    rows = torch.randn((n, d), generator=g, device=dev)
    rows = rows / rows.norm(dim=1, keepdim=True)
    q = (rows[torch.arange(64, device=dev) * 31] + 0.1 / d ** 0.5 * torch.randn((64, d), generator=g, device=dev)).contiguous()
    if "flat31k" in args.indexes:
        with FlatIndex(rows) as fx:
            sweep("flat 31173x768 fp32 cosine top-10", fx, q.cpu().numpy(), 0, lambda nq: n * d * 4.0)
    del rows
    # a flat index far larger than the L2 (3.07 GB): the small-batch scan against the HBM roofline
    n = args.n_ivf
    if "flat1m" in args.indexes:
        rows = torch.randn((n, d), generator=g, device=dev)
        q = torch.randn((64, d), generator=g, device=dev)
        with FlatIndex(rows) as fx:
            sweep(f"flat {n}x768 fp32 cosine top-10", fx, q.cpu().numpy(), 0, lambda nq: n * d * 4.0)
        del rows
    if "flat1mc" in args.indexes:
        # the same size with cluster structure (configs[1]'s rows as a flat index): what embeddings look like
        c = torch.randn((2048, d), generator=g, device=dev)
        rows = (c[torch.randint(0, 2048, (n,), generator=g, device=dev)] + 0.1 * torch.randn((n, d), generator=g, device=dev)).contiguous()
        q = (c[torch.randint(0, 2048, (64,), generator=g, device=dev)] + 0.1 * torch.randn((64, d), generator=g, device=dev)).contiguous()
        with FlatIndex(rows) as fx:
            sweep(f"flat {n}x768 fp32 cosine top-10, clustered rows", fx, q.cpu().numpy(), 0, lambda nq: n * d * 4.0)
        del rows, c
    if "ivf" not in args.indexes:
        return
    # configs[1]: IVF-FLAT, nlist 1024, nprobe 32
    n = args.n_ivf
    c = torch.randn((2048, d), generator=g, device=dev)
    rows = (c[torch.randint(0, 2048, (n,), generator=g, device=dev)] + 0.1 * torch.randn((n, d), generator=g, device=dev)).contiguous()
    q = (c[torch.randint(0, 2048, (64,), generator=g, device=dev)] + 0.1 * torch.randn((64, d), generator=g, device=dev)).contiguous()
    hb.set_mode(hb.MODE_FAST)
    ix = ivf_flat.build_index(rows, num_partitions=1024, max_iterations=10)
    hb.set_mode(hb.MODE_EXACT)
    try:
        sweep(f"ivf-flat {n}x768 fp32 cosine nlist=1024 nprobe=32 top-10", ix, q.cpu().numpy(), 32,
              lambda nq: min(nq * 32, 1024) * (n / 1024.0) * d * 4.0 + 1024 * d * 8.0)
    finally:
        ix.close()


if __name__ == "__main__":
    main()

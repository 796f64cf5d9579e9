"""What the per-rank slice of the sharded coarse routing costs, measured on one GPU: a FAST flat scan of `--rows` fp64 rows
(a centroid slice) x `--nq` queries, k = nprobe.  Prints the stage times the library records (profile = 1) for the levelled
path and for the sample + main path (fast_level_min raised above the slice's tile count).

    python tools/probe_slice.py [--rows 8192] [--nq 10000] [--k 32] [--reps 10]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from hnsw_clj_b200 import _lib as hb  # noqa: E402
from hnsw_clj_b200.flat import FlatIndex  # noqa: E402

STAGES = ("tc_ms", "tc_sample_ms", "select_ms", "pack_ms", "rescore_ms", "plan_ms")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default=None, help="run one variant only (e.g. 'levelled x16'), for a launch list under ncu")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    rows = torch.randn((args.rows, args.d), generator=g, device=dev, dtype=torch.float64)
    rows /= rows.norm(dim=1, keepdim=True)
    pick = torch.randint(0, args.rows, (args.nq,), generator=g, device=dev)
    queries = (rows[pick] + 0.02 * torch.randn((args.nq, args.d), generator=g, device=dev, dtype=torch.float64)).float().contiguous()
    ix = FlatIndex(rows, "cosine")
    out = {}
    for name, level_min, ratio, dense0 in (("levelled x16", 33, 16, 8), ("levelled x8", 33, 8, 8), ("levelled x16 dense 4", 33, 16, 4),
                                           ("levelled x16 dense 16", 33, 16, 16), ("levelled x16 emit-first", 33, 16, 0),
                                           ("sample+main", 1 << 20, 16, 8)):
        if args.only and name != args.only:
            continue
        hb.set_mode(hb.MODE_FAST)
        hb.set_option("fast_level_min", level_min)
        hb.set_option("fast_level_ratio", ratio)
        hb.set_option("fast_level_dense", dense0)
        ix.search_raw(queries, args.k)
        torch.cuda.synchronize()
        hb.set_option("profile", 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = hb.launch_count(reset=True)
        e0.record()
        for _ in range(args.reps):
            ids, dist = ix.search_raw(queries, args.k)
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"ms": e0.elapsed_time(e1) / args.reps, "launches": hb.launch_count() / args.reps,
                     **{s: hb.get_stat(s) / args.reps for s in STAGES},
                     "fallbacks": hb.get_stat("fast_fallbacks"), "served": hb.get_stat("fast_queries")}
        hb.set_option("profile", 0)
        del n0
    hb.set_option("fast_level_min", 33)
    hb.set_option("fast_level_ratio", 16)
    hb.set_option("fast_level_dense", 8)
    hb.set_mode(hb.MODE_EXACT)
    print(json.dumps({"rows": args.rows, "nq": args.nq, "k": args.k, **out}))


if __name__ == "__main__":
    main()

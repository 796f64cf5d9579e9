"""Times the exact k-means++ seeding (hb_kpp.cu) at the shape of the 8-GPU sharded build: `--rows` sample rows x 768 fp32 of
clustered data with `--centres` clusters, `--nlist` seeds.  Prints seconds and the pruning statistics.

    python tools/probe_kpp.py [--rows 262144] [--nlist 65536] [--centres 131072]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from hnsw_clj_b200 import _lib as hb  # noqa: E402
from hnsw_clj_b200 import ivf_flat  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=262144)
    ap.add_argument("--nlist", type=int, default=65536)
    ap.add_argument("--centres", type=int, default=131072)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--noise", type=float, default=0.1)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    centres = torch.randn((args.centres, args.d), generator=g, device=dev)
    idx = torch.randint(0, args.centres, (args.rows,), generator=g, device=dev)
    rows = (centres[idx] + args.noise * torch.randn((args.rows, args.d), generator=g, device=dev)).contiguous()
    del centres
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    seeds = ivf_flat.kmeanspp_init(rows, args.nlist)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"rows": args.rows, "nlist": args.nlist, "centres": args.centres, "seconds": dt,
                      "kpp_rows_scored": hb.get_stat("kpp_rows_scored"), "kpp_chunks_walked": hb.get_stat("kpp_chunks_walked"),
                      "kpp_steps": hb.get_stat("kpp_steps"), "first_seeds": [int(x) for x in list(seeds[:4])]}))


if __name__ == "__main__":
    main()

"""Debug: a reduced configs[1] (clustered rows, IVF-FLAT) in FAST mode with the library's fast_debug prints."""
import sys
import time

import torch

sys.path.insert(0, "/root/repo")
from hnsw_clj_b200 import _lib, ivf_flat

_lib.check(_lib.lib().hb_init(0))
n, d, nlist, nprobe, nq, k = 200000, 768, 512, 32, 4000, 10
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
cent = torch.randn((1024, d), generator=g, device=dev)
rows = (cent[torch.randint(0, 1024, (n,), generator=g, device=dev)] + 0.1 * torch.randn((n, d), generator=g, device=dev)).contiguous()
q = (cent[torch.randint(0, 1024, (nq,), generator=g, device=dev)] + 0.1 * torch.randn((nq, d), generator=g, device=dev)).contiguous()
_lib.set_mode(_lib.MODE_FAST)
ix = ivf_flat.build_index(rows, num_partitions=nlist, max_iterations=3)
oi = torch.empty((nq, k), dtype=torch.int64, device=dev)
od = torch.empty((nq, k), dtype=torch.float64, device=dev)
for so in (0, 1):
    _lib.set_option("fast_set_only", so)
    ix.search_raw(q, k, nprobe, out_ids=oi, out_dist=od)
    _lib.set_option("fast_debug", 1)
    _lib.set_option("profile", 1)
    print(f"--- fast_set_only={so}", file=sys.stderr, flush=True)
    ix.search_raw(q, k, nprobe, out_ids=oi, out_dist=od)
    torch.cuda.synchronize()
    print({s: _lib.get_stat(s) for s in ("coarse_ms", "rescore_ms", "fast_queries", "fast_fallbacks")}, file=sys.stderr, flush=True)
    _lib.set_option("fast_debug", 0)
    _lib.set_option("profile", 0)

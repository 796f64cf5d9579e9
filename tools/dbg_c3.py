import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from hnsw_clj_b200 import _lib as hb
from hnsw_clj_b200.flat import FlatIndex
hb.check(hb.lib().hb_init(0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 256
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(42)
rows = torch.randn((n, 768), generator=g, device=dev).to(torch.bfloat16)
q = torch.randn((nq, 768), generator=g, device=dev).to(torch.bfloat16).float().contiguous()
ix = FlatIndex(rows, 'ip')
hb.set_option('fast_debug', 1)
hb.set_mode(hb.MODE_FAST)
ix.search_raw(q, k)

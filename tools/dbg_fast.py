import numpy as np, sys
sys.path.insert(0,'/root/repo')
from hnsw_clj_b200 import _lib, ivf_flat
_lib.check(_lib.lib().hb_init(0))
def clustered(n, d, seed, centres=16, noise=0.1):
    r = np.random.default_rng(seed)
    c = r.standard_normal((centres, d))
    return (c[r.integers(0, centres, n)] + noise * r.standard_normal((n, d))).astype(np.float32)
n,d,nlist,nprobe,nq,k=4000,128,300,40,129,5
rows = clustered(n, d, n + nlist, centres=max(8, nlist // 2))
queries = clustered(nq, d, n + nlist + 1, centres=max(8, nlist // 2))
ix = ivf_flat.build_index(rows, num_partitions=nlist, max_iterations=2)
_lib.set_option("fast_debug",1)
_lib.set_mode(_lib.MODE_FAST)
fids, fdist = ix.search_raw(queries, k, nprobe)

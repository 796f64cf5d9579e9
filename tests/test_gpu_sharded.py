"""Row-sharded search on the device (single process, both shards on cuda:0, merged by the device merge
kernel) equals the unsharded search and the oracle."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_two_shards_on_one_gpu_equal_global():
    from hnsw_clj_b200 import _lib, ivf_flat
    from hnsw_clj_b200.sharded import ShardedIVFFlat

    r = np.random.default_rng(3)
    c = r.standard_normal((30, 64))
    rows = (c[r.integers(0, 30, 5000)] + 0.1 * r.standard_normal((5000, 64))).astype(np.float32)
    q = (c[r.integers(0, 30, 77)] + 0.1 * r.standard_normal((77, 64))).astype(np.float32)
    cents, asg = orc.kmeans(rows, 16, iters=2, seed=42)
    k, nprobe, world = 10, 5, 2
    shards = [ShardedIVFFlat(rows, cents, asg, rk, world) for rk in range(world)]
    parts = [s.local_search(q, k, nprobe) for s in shards]
    all_ids = np.stack([p[0] for p in parts])
    all_d = np.stack([p[1] for p in parts])
    out_i = np.empty((77, k), np.int64)
    out_d = np.empty((77, k), np.float64)
    _lib.check(_lib.lib().hb_topk_merge(all_d.ctypes.data, all_ids.ctypes.data, world, 77, k, out_i.ctypes.data,
                                        out_d.ctypes.data))
    want_i, want_d = orc.ivf_search(rows, cents, asg, q, k, nprobe)
    assert out_i.tolist() == want_i.tolist()
    assert (out_d.view(np.int64) == want_d.view(np.int64)).all()
    with ivf_flat.import_index(rows, cents, asg) as ix:
        gi, gd = ix.search_raw(q, k, nprobe)
    assert gi.tolist() == out_i.tolist()
    for s in shards:
        s.close()


def test_row_sharded_flat_equals_global():
    """BASELINE configs[2]'s layout on one GPU: two row blocks, local top-k, device merge == unsharded search == oracle;
    duplicated rows across the block boundary keep global row order."""
    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.flat import FlatIndex
    from hnsw_clj_b200.sharded import ShardedFlat, row_range

    r = np.random.default_rng(5)
    rows = r.standard_normal((3001, 48)).astype(np.float32)
    rows[1500] = rows[1499]  # a tie across the shard boundary
    rows[2000] = rows[10]
    q = np.concatenate([r.standard_normal((40, 48)).astype(np.float32), rows[[1499, 10]]])
    k, world = 20, 2
    shards = []
    for rk in range(world):
        lo, hi = row_range(rows.shape[0], rk, world)
        shards.append(ShardedFlat(rows[lo:hi], lo, rk, world, "ip"))
    parts = [s.local_search(q, k) for s in shards]
    all_ids = np.stack([p[0] for p in parts])
    all_d = np.stack([p[1] for p in parts])
    out_i = np.empty((len(q), k), np.int64)
    out_d = np.empty((len(q), k), np.float64)
    _lib.check(_lib.lib().hb_topk_merge(all_d.ctypes.data, all_ids.ctypes.data, world, len(q), k, out_i.ctypes.data,
                                        out_d.ctypes.data))
    want_i, want_d = orc.exact_knn(rows, q, k, orc.IP)
    assert out_i.tolist() == want_i.tolist()
    assert (out_d.view(np.int64) == want_d.view(np.int64)).all()
    with FlatIndex(rows, "ip") as ix:
        gi, gd = ix.search_raw(q, k)
    assert gi.tolist() == out_i.tolist()
    for s in shards:
        s.close()

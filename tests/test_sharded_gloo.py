"""N > 1 host logic on CPU: world_size-2 gloo run of the row-sharded search (list ownership, local top-k with
global ids, all-gather, merge rule) and of the sharded Lloyd update reduction.  The local scan is played by the
oracle here (no GPU in this container); on the GPU box the same code path runs the CUDA kernels
(tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    r = np.random.default_rng(9)
    c = r.standard_normal((10, 16))
    rows = (c[r.integers(0, 10, 900)] + 0.1 * r.standard_normal((900, 16))).astype(np.float32)
    q = (c[r.integers(0, 10, 21)] + 0.1 * r.standard_normal((21, 16))).astype(np.float32)
    return rows, q


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hnsw_clj_b200 import sharded

    rows, q = _data()
    cents, asg = orc.kmeans(rows, 8, iters=3, seed=42)
    k, nprobe = 10, 3
    # this rank's shard: the rows of the lists it owns; lists owned elsewhere are empty here
    mine = np.nonzero(sharded.list_owner(8, world)[asg] == rank)[0]
    l_ids, l_dist = orc.ivf_search(rows[mine], cents, asg[mine], q, k, nprobe)
    g_ids = np.where(l_ids >= 0, mine[np.maximum(l_ids, 0)], -1)
    ids, d = sharded.all_gather_merge(g_ids, l_dist, world)
    want_ids, want_d = orc.ivf_search(rows, cents, asg, q, k, nprobe)
    ok = ids.tolist() == want_ids.tolist() and d.tolist() == want_d.tolist()
    # sharded Lloyd update: all-reduced partial sums == single-process update (to fp64 rounding)
    half = np.array_split(np.arange(rows.shape[0]), world)[rank]
    sums = np.zeros((8, 16))
    cnt = np.zeros(8, dtype=np.int64)
    np.add.at(sums, asg[half], rows[half].astype(np.float64))
    np.add.at(cnt, asg[half], 1)
    ts, tc = torch.from_numpy(sums), torch.from_numpy(cnt)
    dist.all_reduce(ts)
    dist.all_reduce(tc)
    new = cents.copy()
    nz = tc.numpy() > 0
    new[nz] = ts.numpy()[nz] / tc.numpy()[nz, None]
    ok2 = np.allclose(new, orc.update_centroids(rows, asg, cents), rtol=1e-13, atol=0)
    # row-sharded flat search (configs[2]): contiguous row blocks, local exact top-k, all-gather, merge
    lo, hi = sharded.row_range(rows.shape[0], rank, world)
    f_ids, f_dist = orc.exact_knn(rows[lo:hi], q, k)
    f_ids = np.where(f_ids >= 0, f_ids + lo, -1)
    m_ids, m_d = sharded.all_gather_merge(f_ids, f_dist, world)
    w_ids, w_d = orc.exact_knn(rows, q, k)
    ok3 = m_ids.tolist() == w_ids.tolist() and m_d.tolist() == w_d.tolist()
    ret[rank] = (ok, ok2 and ok3)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_sharded_search_and_update():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: (True, True), 1: (True, True)}


def test_merge_host_rule():
    from hnsw_clj_b200.sharded import merge_host

    d = np.array([[[0.1, 0.3]], [[0.1, 0.2]]])  # [world=2, nq=1, k=2]
    i = np.array([[[5, 6]], [[7, 8]]])
    ids, dd = merge_host(i, d)
    assert ids.tolist() == [[5, 7]] and dd.tolist() == [[0.1, 0.1]]  # tie: lower rank first

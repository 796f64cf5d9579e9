"""The library's multi-GPU data plane (include/hnswb200.h "multi-GPU"; reference model: partitioned_hnsw.clj:149-196).

One-GPU part (always runs on the GPU box): a one-rank communicator drives hb_sharded_search through BOTH exchange
paths — pack + ncclAllGather + merge kernel, and the fused peer-window kernel looped back onto its own window — and the
sharded k-means / IVF build, whose all-reduce is then the identity, must equal the single-GPU build bit for bit.

Two-GPU part (skipped on a one-GPU box; run with `gpurun --gpus 2`): two processes, one per GPU, contiguous row blocks;
the sharded flat and IVF-FLAT searches must equal the oracle on the global rows (ids and fp64 distance bits), through the
peer-window kernel and through NCCL.
"""
import os
import socket
import tempfile
import threading

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _data(n=6000, d=64, nq=90, seed=17):
    r = np.random.default_rng(seed)
    c = r.standard_normal((24, d))
    rows = (c[r.integers(0, 24, n)] + 0.1 * r.standard_normal((n, d))).astype(np.float32)
    q = (c[r.integers(0, 24, nq)] + 0.1 * r.standard_normal((nq, d))).astype(np.float32)
    return rows, q


@pytest.fixture(scope="module")
def comm1():
    from hnsw_clj_b200 import _lib, sharded

    _lib.check(_lib.lib().hb_init(0))
    info = sharded.comm_init(rank=0, world=1, id_bytes=sharded.comm_unique_id())
    assert info["nranks"] == 1 and info["rank"] == 0
    yield sharded
    _lib.set_option("comm_p2p", 1)
    sharded.comm_shutdown()


@pytest.mark.parametrize("p2p", [0, 2])
def test_one_rank_sharded_flat_applies_id_base(comm1, p2p):
    from hnsw_clj_b200 import _lib

    rows, q = _data()
    _lib.set_option("comm_p2p", p2p)
    want_i, want_d = orc.exact_knn(rows, q, 10)
    with comm1.RowShardedFlat(rows, first_row=1000) as sh:
        for _ in range(3):  # both parities of the window, and a re-used epoch counter
            ids, d = sh.search_raw(q, 10)
            assert ids.tolist() == (want_i + 1000).tolist()
            assert (d.view(np.int64) == want_d.view(np.int64)).all()
        ids, d = sh.search_raw(q[:1], 10)  # a small batch after a large one (the window is large enough already)
        assert ids.tolist() == (want_i[:1] + 1000).tolist()
        # k > n_local pads with -1 / inf and the pad keeps id -1 (no id_base added)
        ids, d = sh.search_raw(q[:5], 10)
    with comm1.RowShardedFlat(rows[:4], first_row=7) as sh:
        ids, d = sh.search_raw(q[:3], 6)
        assert (ids[:, 4:] == -1).all() and np.isinf(d[:, 4:]).all() and (ids[:, :4] >= 7).all()
    assert comm1.comm_info()["p2p"] == (p2p == 2) or p2p == 0


def test_one_rank_sharded_build_equals_single_gpu_build(comm1):
    from hnsw_clj_b200 import ivf_flat

    rows, q = _data()
    seeds = ivf_flat.kmeanspp_init(rows, 16)
    cents1, asg1 = ivf_flat.partition_vectors_kmeans(rows, 16, max_iterations=3, seed_rows=seeds)
    cents2, asg2 = comm1.sharded_kmeans(rows, 0, 16, seeds, max_iterations=3)
    assert (cents1.view(np.int64) == cents2.view(np.int64)).all() and (asg1 == asg2).all()
    oc, oa = orc.kmeans(rows, 16, iters=3, seed=42)
    assert (oc.view(np.int64) == cents2.view(np.int64)).all() and (oa == asg2).all()
    want_i, want_d = orc.ivf_search(rows, oc, oa, q, 10, 4)
    with comm1.RowShardedIVFFlat(rows, 0, 16, seeds, max_iterations=3) as sh:
        ids, d = sh.search_raw(q, 10, 4)
        assert ids.tolist() == want_i.tolist() and (d.view(np.int64) == want_d.view(np.int64)).all()
    with pytest.raises(ValueError):
        comm1.sharded_kmeans(rows, 0, 16, np.full(16, 10 ** 9), max_iterations=1)  # seed row outside the global rows


def test_per_index_mode_and_combined_small_calls(comm1):
    """hb_index_set_mode pins the mode per index; concurrent one-query calls on host buffers are answered in combined
    batches (parallel-search-futures, helper/parallel_search.clj:15-49) with the bits a lone call returns."""
    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.flat import FlatIndex

    rows, q = _data(nq=200)
    want_i, want_d = orc.exact_knn(rows, q, 10)
    with FlatIndex(rows) as fx:
        _lib.check(_lib.lib().hb_index_set_mode(fx._h, _lib.MODE_FAST))
        _lib.set_option("profile", 1)
        ids, d = fx.search_raw(q, 10)
        assert _lib.get_stat("fast_queries") == len(q)  # FAST although the process default is EXACT
        _lib.check(_lib.lib().hb_index_set_mode(fx._h, -1))
        assert ids.tolist() == want_i.tolist() and (d.view(np.int64) == want_d.view(np.int64)).all()
        out = {}

        def worker(t):
            for j in range(t, len(q), 32):
                i1, d1 = fx.search_raw(q[j], 10)
                out[j] = (i1[0].tolist(), d1[0].view(np.int64).tolist())

        b0, r0 = _lib.get_stat("micro_batches"), _lib.get_stat("micro_batch_requests")
        ts = [threading.Thread(target=worker, args=(t,)) for t in range(32)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        _lib.set_option("profile", 0)
        nb, nr = _lib.get_stat("micro_batches") - b0, _lib.get_stat("micro_batch_requests") - r0
        assert nr > nb >= 1  # at least one combined batch held several callers
        for j in range(len(q)):
            assert out[j] == (want_i[j].tolist(), want_d[j].view(np.int64).tolist())
        with pytest.raises(ValueError):
            fx.search_raw(q[0][:5], 10)


# ---- two GPUs, two processes ---------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, id_path, ret):
    import torch

    from hnsw_clj_b200 import _lib, sharded

    torch.cuda.set_device(rank)
    _lib.check(_lib.lib().hb_init(rank))
    if rank == 0:
        with open(id_path + ".tmp", "wb") as f:
            f.write(sharded.comm_unique_id())
        os.replace(id_path + ".tmp", id_path)
    else:
        import time

        while not os.path.exists(id_path):
            time.sleep(0.01)
    sharded.comm_init(rank, world, open(id_path, "rb").read())
    rows, q = _data(n=9001, nq=300)
    rows[4500] = rows[4499]  # an exact tie across the shard boundary
    q[0] = rows[4499]
    lo, hi = sharded.row_range(rows.shape[0], rank, world)
    res = {}
    want_i, want_d = orc.exact_knn(rows, q, 10)
    oc, oa = orc.kmeans(rows, 16, iters=2, seed=42)
    wi, wd = orc.ivf_search(rows, oc, oa, q, 10, 4)
    for p2p in (1, 0):
        _lib.set_option("comm_p2p", p2p)
        with sharded.RowShardedFlat(rows[lo:hi], lo) as sh:
            for rep in range(3):
                ids, d = sh.search_raw(q, 10)
            res[f"flat_p2p{p2p}"] = ids.tolist() == want_i.tolist() and bool((d.view(np.int64) == want_d.view(np.int64)).all())
            # device buffers in and out, FAST mode
            _lib.check(_lib.lib().hb_index_set_mode(sh._h, _lib.MODE_FAST))
            tq = torch.from_numpy(q).cuda()
            ti = torch.empty((len(q), 10), dtype=torch.int64, device="cuda")
            td = torch.empty((len(q), 10), dtype=torch.float64, device="cuda")
            sh.search_raw(tq, 10, out_ids=ti, out_dist=td)
            torch.cuda.synchronize()
            res[f"flat_fast_dev_p2p{p2p}"] = ti.cpu().numpy().tolist() == want_i.tolist() and bool(
                (td.cpu().numpy().view(np.int64) == want_d.view(np.int64)).all())
        # rows of the oracle's partitions split by row block: every rank scans its part of every probed list
        with sharded.import_row_shard(rows[lo:hi], lo, oc, oa[lo:hi]) as sh:
            ids, d = sh.search_raw(q, 10, 4)
            res[f"ivf_p2p{p2p}"] = ids.tolist() == wi.tolist() and bool((d.view(np.int64) == wd.view(np.int64)).all())
        res[f"info_p2p{p2p}"] = sharded.comm_info()["p2p"]
    # coarse routing split over the ranks (needs >= 256 centroids per rank): rank r ranks its half of the 512 centroids exactly,
    # the top-nprobe lists are exchanged and merged; results must still be the oracle's, with and without the split
    rows2, q2 = _data(n=20000, d=64, nq=300, seed=29)
    lo2, hi2 = sharded.row_range(rows2.shape[0], rank, world)
    oc2, oa2 = orc.kmeans(rows2, 512, iters=1, seed=42)
    wi2, wd2, wp2 = orc.ivf_search(rows2, oc2, oa2, q2, 10, 8, return_probes=True)
    for split in (1, 0):
        for p2p in (1, 0):
            _lib.set_option("comm_p2p", p2p)
            with sharded.import_row_shard(rows2[lo2:hi2], lo2, oc2, oa2[lo2:hi2]) as sh:
                _lib.check(_lib.lib().hb_index_set_coarse_sharded(sh._h, split))
                _lib.check(_lib.lib().hb_index_set_mode(sh._h, _lib.MODE_FAST))
                for rep in range(2):
                    ids, d = sh.search_raw(q2, 10, 8)
                res[f"coarse_split{split}_p2p{p2p}"] = ids.tolist() == wi2.tolist() and bool((d.view(np.int64) == wd2.view(np.int64)).all())
    _lib.set_option("comm_p2p", 1)
    # data-parallel k-means: same seeds as the single-GPU build; assignments equal, centroids to the last ulps
    seeds = orc.kmeanspp_init(rows, 16, seed=42)
    cents, asg = sharded.sharded_kmeans(rows[lo:hi], lo, 16, seeds, max_iterations=2)
    res["kmeans_assign"] = bool((asg == oa[lo:hi]).all())
    res["kmeans_cents"] = bool(np.allclose(cents, oc, rtol=1e-12, atol=0))
    with sharded.RowShardedIVFFlat(rows[lo:hi], lo, 16, seeds, max_iterations=2) as sh:
        ids, d = sh.search_raw(q, 10, 4)
        res["sharded_build_search_ids"] = ids.tolist() == wi.tolist()
    ret[rank] = res
    sharded.comm_shutdown()


def test_two_gpus_row_sharded_equals_oracle():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = 2
    with tempfile.TemporaryDirectory() as tmp, mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_rank_main, args=(world, os.path.join(tmp, "comm.id"), ret), nprocs=world, join=True)
        got = dict(ret)
    assert set(got) == {0, 1}
    for rank, res in got.items():
        bad = [k for k, v in res.items() if v is not True and not k.startswith("info_")]
        assert not bad, (rank, res)
        assert res["info_p2p0"] in (True, False)


def test_second_device_from_other_threads():
    """hb_init(1) then calls from threads that never called cudaSetDevice (the current device is per thread)."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_device1_main, args=(ret,), nprocs=1, join=True)
        assert dict(ret) == {0: True}


def _device1_main(_, ret):
    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.flat import FlatIndex

    _lib.check(_lib.lib().hb_init(1))
    rows, q = _data(n=3000, nq=64)
    want_i, _ = orc.exact_knn(rows, q, 5)
    ok = []
    with FlatIndex(rows) as fx:
        def worker(t):
            i1, _d = fx.search_raw(q[t * 8:(t + 1) * 8 + 1], 5)  # 9 queries: not combined, straight to the device
            i2, _d = fx.search_raw(q[t], 5)
            ok.append(i1.tolist() == want_i[t * 8:(t + 1) * 8 + 1].tolist() and i2.tolist() == want_i[t:t + 1].tolist())

        ts = [threading.Thread(target=worker, args=(t,)) for t in range(7)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    ret[0] = len(ok) == 7 and all(ok)

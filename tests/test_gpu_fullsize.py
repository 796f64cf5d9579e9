"""Parity at BASELINE.json's FULL sizes (configs[0], [2], [4]; configs[1] is checked inside every bench.py run, configs[3] by
tests/test_gpu_multi.py at test size and bench.py --gpus N at full size).

The oracle cannot finish these sizes in seconds, so each case combines (a) the oracle on a bounded sample of the queries,
(b) the device's own EXACT mode (the reference's fp64 sum for every pair) against the FAST mode the benchmarks run, and
(c) size-independent properties: results ascending, ids unique and in range, every returned distance re-derived by the
oracle's pairwise arithmetic from the row it names.  Marked slow: about a minute on one B200 (set HB_FULLSIZE=0 to skip)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = [pytest.mark.gpu, pytest.mark.slow,
              pytest.mark.skipif(os.environ.get("HB_FULLSIZE", "1") == "0", reason="HB_FULLSIZE=0")]


def _bits(a):
    return np.asarray(a).view(np.int64)


@pytest.fixture(scope="module")
def dev():
    import torch

    from hnsw_clj_b200 import _lib

    _lib.check(_lib.lib().hb_init(0))
    yield torch.device("cuda", 0)
    _lib.set_mode(_lib.MODE_EXACT)
    _lib.check(_lib.lib().hb_shutdown())  # give the workspace back before the next module
    torch.cuda.empty_cache()


def _unit_rows(n, d, seed, device):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.randn((n, d), generator=g, device=device)
    return x / x.norm(dim=1, keepdim=True)


def test_config0_flat_31173x768_top10(dev):
    """configs[0]: Karoli-Bible shape, 31,173 x 768 fp32 cosine, 1000 queries, exact flat top-10."""
    import torch

    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.flat import FlatIndex

    n, d, nq, k = 31173, 768, 1000, 10
    rows = _unit_rows(n, d, 42, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(43)
    q = (rows[torch.arange(nq, device=dev) * 31] + 0.1 / d ** 0.5 * torch.randn((nq, d), generator=g, device=dev)).contiguous()
    with FlatIndex(rows) as ix:
        e_ids, e_d = ix.search_raw(q, k)
        _lib.set_mode(_lib.MODE_FAST)
        f_ids, f_d = ix.search_raw(q, k)
        _lib.set_mode(_lib.MODE_EXACT)
        one_ids, one_d = ix.search_raw(q[7:8], k)  # the reference's calling pattern: one query per call
    assert (f_ids == e_ids).all() and (_bits(f_d) == _bits(e_d)).all()
    assert (one_ids == e_ids[7:8]).all() and (_bits(one_d) == _bits(e_d[7:8])).all()
    s = 96
    want_i, want_d = orc.exact_knn(rows.cpu().numpy(), q[:s].cpu().numpy(), k)
    assert e_ids[:s].tolist() == want_i.tolist() and (_bits(e_d[:s]) == _bits(want_d)).all()
    assert (np.diff(e_d, axis=1) >= 0).all() and (e_ids[:, 0] == np.arange(nq) * 31).all()  # a query's own row comes first


def test_config2_flat_10Mx768_bf16_ip_top100(dev):
    """configs[2] on one GPU: 10M x 768 bf16 inner product, 4096 queries, top-100 (FAST mode, as benchmarked)."""
    import torch

    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.flat import FlatIndex

    n, d, nq, k = 10_000_000, 768, 4096, 100
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    rows = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    for i in range(0, n, 1 << 20):
        m = min(1 << 20, n - i)
        rows[i:i + m] = torch.randn((m, d), generator=g, device=dev).to(torch.bfloat16)
    g.manual_seed(43)
    q = torch.randn((nq, d), generator=g, device=dev).to(torch.bfloat16).float().contiguous()
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    dist = torch.empty((nq, k), dtype=torch.float64, device=dev)
    with FlatIndex(rows, "ip") as ix:
        _lib.set_mode(_lib.MODE_FAST)
        _lib.set_option("profile", 1)
        ix.search_raw(q, k, out_ids=ids, out_dist=dist)
        served, fell = _lib.get_stat("fast_queries"), _lib.get_stat("fast_fallbacks")
        _lib.set_option("profile", 0)
        _lib.set_mode(_lib.MODE_EXACT)
        s = 24
        e_ids, e_d = ix.search_raw(q[:s].contiguous(), k)  # fp64 for every one of the 10M rows
    ids_np, d_np = ids.cpu().numpy(), dist.cpu().numpy()
    assert served == nq and fell <= nq // 100  # the candidate pass answered (nearly) every query itself
    assert (ids_np[:s] == e_ids).all() and (_bits(d_np[:s]) == _bits(e_d)).all()
    assert (np.diff(d_np, axis=1) >= 0).all() and ids_np.min() >= 0 and ids_np.max() < n
    assert all(len(set(r)) == k for r in ids_np[::97].tolist())
    # every returned distance is -dot(q, row) in the reference's sequential fp64 arithmetic on the bf16 values
    qs = [0, 1, 4095]
    got_rows = rows[torch.from_numpy(ids_np[qs].reshape(-1)).to(dev)].float().cpu().numpy().reshape(len(qs), k, d)
    q_np = q.cpu().numpy()
    for a, qi in enumerate(qs):
        for j in range(0, k, 9):
            want = -orc.dot(q_np[qi].astype(np.float64), got_rows[a, j].astype(np.float64))
            assert np.float64(want).view(np.int64) == d_np[qi, j].view(np.int64)
    del rows
    torch.cuda.empty_cache()


def test_config4_hnsw_1M_ef128(dev):
    """configs[4]: M = 16, efSearch = 128, 1M x 768 unit-norm rows, 16,384 concurrent queries; the batched traversal must
    return what the oracle's restatement of search-knn (ultra_fast.clj:346-374) returns on the SAME graph."""
    import torch

    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.ultra_fast import HnswIndex, bulk_knn_graph

    n, d, nq, k, ef = 1_000_000, 768, 16384, 10, 128
    rows = _unit_rows(n, d, 42, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(43)
    q = (rows[torch.randint(0, n, (nq,), generator=g, device=dev)] + 0.3 / d ** 0.5 * torch.randn((nq, d), generator=g, device=dev)).contiguous()
    _lib.set_mode(_lib.MODE_FAST)
    levels, entry, adjacency = bulk_knn_graph(rows, M=16, level_seed=42)
    _lib.set_mode(_lib.MODE_EXACT)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    dist = torch.empty((nq, k), dtype=torch.float64, device=dev)
    with HnswIndex(rows, levels, entry, adjacency, distance_fn="cosine") as ix:
        ix.search_raw(q, k, ef, out_ids=ids, out_dist=dist)
    ids_np, d_np = ids.cpu().numpy(), dist.cpu().numpy()
    s = 48
    graph = orc.Hnsw.from_graph(rows.cpu().numpy(), levels, entry, adjacency)
    want_i, want_d = graph.search(q[:s].cpu().numpy(), k, ef)
    assert ids_np[:s].tolist() == want_i.tolist() and (_bits(d_np[:s]) == _bits(want_d)).all()
    assert ids_np.min() >= 0 and ids_np.max() < n and all(len(set(r)) == k for r in ids_np[::331].tolist())
    del rows
    torch.cuda.empty_cache()

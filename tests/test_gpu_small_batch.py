"""Small batches (1..8 queries per call — the reference's own calling pattern, `search-knn` with one query:
src/hnsw/ann/partition/ivf_flat.clj:300-317, src/hnsw/bench.clj:72-84) take the HBM-bound thread-per-row scan
(`smallscan_kernel`) and a split selection instead of the 128 x 64 fp64 tiles.  Same bits as the oracle and as the
large-batch path, in EXACT and FAST mode."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hnsw_clj_b200 as pkg
    from hnsw_clj_b200 import _lib

    _lib.check(_lib.lib().hb_init(0))
    return pkg


def rng_rows(n, d, seed, clustered=0):
    r = np.random.default_rng(seed)
    if clustered:
        c = r.standard_normal((clustered, d))
        x = c[r.integers(0, clustered, n)] + 0.1 * r.standard_normal((n, d))
    else:
        x = r.standard_normal((n, d))
    return x.astype(np.float32)


def same_bits(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool((a.view(np.int64) == b.view(np.int64)).all())


@pytest.mark.parametrize("n,d,k", [(5000, 768, 10), (300, 100, 7), (129, 4, 1), (9000, 36, 100)])
@pytest.mark.parametrize("nq", [1, 2, 4, 5, 8])
@pytest.mark.parametrize("metric", ["cosine", "euclidean", "ip"])
def test_flat_small_batch_parity(hb, n, d, k, nq, metric):
    from hnsw_clj_b200.flat import FlatIndex

    rows, q = rng_rows(n, d, 10 + nq), rng_rows(nq, d, 11)
    code = {"cosine": orc.COSINE, "euclidean": orc.L2, "ip": orc.IP}[metric]
    want_ids, want_d = orc.exact_knn(rows, q, k, code)
    with FlatIndex(rows, distance_fn=metric) as ix:
        ids, dist = ix.search_raw(q, k)
    assert ids.tolist() == want_ids.tolist()
    assert same_bits(dist, want_d)


def test_flat_small_batch_dtypes_and_unaligned_rows(hb):
    import torch

    from hnsw_clj_b200.flat import FlatIndex

    # bf16 rows, fp64 queries (the reference's double[]), inner product
    rows = torch.from_numpy(rng_rows(4100, 64, 5)).to(torch.bfloat16)
    q = torch.from_numpy(rng_rows(3, 64, 6)).to(torch.bfloat16).float()
    want_ids, want_d = orc.exact_knn(rows.float().numpy(), q.numpy(), 50, orc.IP)
    with FlatIndex(rows.cuda(), distance_fn="ip") as ix:
        ids, dist = ix.search_raw(q.cuda(), 50)
        ids2, dist2 = ix.search_raw(q.numpy().astype(np.float64), 50)
    assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)
    assert ids2.tolist() == want_ids.tolist() and same_bits(dist2, want_d)
    # fp64 rows whose values are not fp32-representable: separately rounded product (mul, then add)
    r64 = np.random.default_rng(7).standard_normal((3000, 40))
    q64 = np.random.default_rng(8).standard_normal((2, 40))
    with FlatIndex(r64) as ix:
        ids, dist = ix.search_raw(q64, 10)
    for qi, qv in enumerate(q64):
        dd = np.array([orc.cosine_distance(qv, r) for r in r64])
        order = np.argsort(dd, kind="stable")[:10]
        assert ids[qi].tolist() == order.tolist() and same_bits(dist[qi], dd[order])
    # a row length that is not a multiple of 16 bytes: the vectorised scan does not apply, the tiled kernel answers
    rows = rng_rows(2500, 7, 9)
    q = rng_rows(1, 7, 10)
    want_ids, want_d = orc.exact_knn(rows, q, 10)
    with FlatIndex(rows) as ix:
        ids, dist = ix.search_raw(q, 10)
    assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)


def test_flat_small_batch_ties_and_k_larger_than_n(hb):
    from hnsw_clj_b200.flat import FlatIndex

    dup = np.repeat(rng_rows(700, 16, 2), 5, axis=0)  # 3500 rows: the split selection sees ties across sub-ranges
    q = dup[[0, 1700]]
    want_ids, want_d = orc.exact_knn(dup, q, 12)
    with FlatIndex(dup) as ix:
        ids, dist = ix.search_raw(q, 12)
    assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)
    assert ids[0, :5].tolist() == [0, 1, 2, 3, 4]
    rows = rng_rows(6, 8, 3)
    with FlatIndex(rows) as ix:
        ids, dist = ix.search_raw(rows[:1], 9)
    assert ids[0, 6:].tolist() == [-1, -1, -1] and np.isinf(dist[0, 6:]).all()


@pytest.mark.parametrize("nq", [1, 3, 8])
@pytest.mark.parametrize("nprobe", [1, 8, 16])
@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_ivf_small_batch_parity(hb, nq, nprobe, mode):
    from hnsw_clj_b200 import _lib, ivf_flat

    rows = rng_rows(20000, 64, 40, clustered=30)
    q = rng_rows(nq, 64, 41 + nq, clustered=30)
    cents, asg = orc.kmeans(rows, 16, iters=2, seed=42)
    want_ids, want_d, want_p = orc.ivf_search(rows, cents, asg, q, 10, nprobe, return_probes=True)
    with ivf_flat.import_index(rows, cents, asg) as ix:
        _lib.set_mode(_lib.MODE_FAST if mode == "fast" else _lib.MODE_EXACT)
        try:
            ids, dist = ix.search_raw(q, 10, nprobe)  # 8 x ~1250 rows per query: split selection + merge
        finally:
            _lib.set_mode(_lib.MODE_EXACT)
        probes = ix.probes(q, nprobe)
    assert probes.tolist() == want_p.tolist()
    assert ids.tolist() == want_ids.tolist()
    assert same_bits(dist, want_d)


def test_small_batch_equals_large_batch(hb):
    """The first rows of a 200-query call and the same queries asked 1, 3 and 8 at a time: identical ids and bits."""
    from hnsw_clj_b200 import _lib, ivf_flat
    from hnsw_clj_b200.flat import FlatIndex

    rows = rng_rows(30000, 128, 50, clustered=64)
    q = rng_rows(200, 128, 51, clustered=64)
    with FlatIndex(rows) as fx, ivf_flat.build_index(rows, num_partitions=32, max_iterations=2) as ix:
        big_f = fx.search_raw(q, 10)
        big_i = ix.search_raw(q, 10, 8)
        for mode in (_lib.MODE_EXACT, _lib.MODE_FAST):
            _lib.set_mode(mode)
            try:
                for nq in (1, 3, 8):
                    f = fx.search_raw(q[:nq], 10)
                    i = ix.search_raw(q[:nq], 10, 8)
                    assert f[0].tolist() == big_f[0][:nq].tolist() and same_bits(f[1], big_f[1][:nq])
                    assert i[0].tolist() == big_i[0][:nq].tolist() and same_bits(i[1], big_i[1][:nq])
            finally:
                _lib.set_mode(_lib.MODE_EXACT)


def test_knobs_never_change_results(hb):
    """hb_set_option knobs pick kernels, never results: the bulk-copy ring in every shape (segment bytes, stages, warps),
    the register-buffered scan (rowstream = 0), M = 64 / M = 128 tensor-core units, probe pruning on / off."""
    from hnsw_clj_b200 import _lib, ivf_flat
    from hnsw_clj_b200.flat import FlatIndex

    rows = rng_rows(12000, 200, 60, clustered=40)  # 800-byte rows: segments of 256 / 384 / 512 bytes all end in a partial one
    q = rng_rows(150, 200, 61, clustered=40)
    defaults = {"rowstream": 1, "stream_seg": 256, "stream_stages": 2, "stream_warps": 0, "tc_half_m": 1, "fast_prune": 1}
    with FlatIndex(rows) as fx, ivf_flat.build_index(rows, num_partitions=256, max_iterations=2) as ix:
        want_f = fx.search_raw(q[:5], 10)
        want_i = ix.search_raw(q[:5], 10, 8)
        want_big = ix.search_raw(q, 10, 8)
        try:
            for knobs in ({"stream_seg": 384}, {"stream_seg": 512, "stream_stages": 3}, {"stream_stages": 4, "stream_warps": 3},
                          {"stream_warps": 1}, {"rowstream": 0}):
                for name, v in {**defaults, **knobs}.items():
                    _lib.set_option(name, v)
                for nq in (1, 5):
                    f = fx.search_raw(q[:nq], 10)
                    i = ix.search_raw(q[:nq], 10, 8)
                    assert f[0].tolist() == want_f[0][:nq].tolist() and same_bits(f[1], want_f[1][:nq]), knobs
                    assert i[0].tolist() == want_i[0][:nq].tolist() and same_bits(i[1], want_i[1][:nq]), knobs
            _lib.set_mode(_lib.MODE_FAST)
            for knobs in ({"tc_half_m": 0}, {"tc_half_m": 1, "fast_prune": 0}, {"tc_half_m": 0, "fast_prune": 0}, {}):
                for name, v in {**defaults, **knobs}.items():
                    _lib.set_option(name, v)
                got = ix.search_raw(q, 10, 8)  # 150 queries over 256 lists: almost every unit holds <= 64 selections
                assert got[0].tolist() == want_big[0].tolist() and same_bits(got[1], want_big[1]), knobs
        finally:
            _lib.set_mode(_lib.MODE_EXACT)
            for name, v in defaults.items():
                _lib.set_option(name, v)


def test_more_lists_than_queries_and_full_units(hb):
    """Unit counts stay on the device (the host sizes by a bound): plans with far more lists than queries, and a plan whose
    units are all full (every query probes the same few lists)."""
    from hnsw_clj_b200 import _lib, ivf_flat

    rows = rng_rows(9000, 64, 70, clustered=3)
    q = rng_rows(700, 64, 71, clustered=3)
    for nlist, nprobe in ((600, 5), (4, 3)):
        with ivf_flat.build_index(rows, num_partitions=nlist, max_iterations=2) as ix:
            want = ix.search_raw(q, 10, nprobe)
            _lib.set_mode(_lib.MODE_FAST)
            try:
                got = ix.search_raw(q, 10, nprobe)
                got9 = ix.search_raw(q[:9], 10, nprobe)
            finally:
                _lib.set_mode(_lib.MODE_EXACT)
        assert got[0].tolist() == want[0].tolist() and same_bits(got[1], want[1]), (nlist, nprobe)
        assert got9[0].tolist() == want[0][:9].tolist() and same_bits(got9[1], want[1][:9]), (nlist, nprobe)


@pytest.mark.parametrize("metric", ["cosine", "ip"])
def test_flat_fast_9_to_33_queries_levelled_narrow_scan(metric):
    """9..32 queries over a long flat list run the levelled candidate scan on tc_narrow_kernel (one narrow unit, thresholds
    raised between the levels); 33 queries take the 128 x 128 kernel.  Same ids and fp64 distance bits as the exact mode and
    the oracle, on clustered and on structureless rows (where a failed proof falls back to the exact kernels)."""
    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.flat import FlatIndex

    _lib.check(_lib.lib().hb_init(0))
    r = np.random.default_rng(77)
    c = r.standard_normal((50, 64))
    clustered = (c[r.integers(0, 50, 70000)] + 0.2 * r.standard_normal((70000, 64))).astype(np.float32)
    gauss = r.standard_normal((70000, 64)).astype(np.float32)
    code = orc.COSINE if metric == "cosine" else orc.IP
    for rows in (clustered, gauss):
        q = (rows[r.integers(0, len(rows), 33)] + 0.1 * r.standard_normal((33, 64))).astype(np.float32)
        want_i, want_d = orc.exact_knn(rows, q, 10, code)
        with FlatIndex(rows, metric) as fx:
            _lib.set_mode(_lib.MODE_FAST)
            try:
                for nq in (9, 20, 32, 33):
                    ids, d = fx.search_raw(q[:nq], 10)
                    assert ids.tolist() == want_i[:nq].tolist(), (metric, nq)
                    assert (d.view(np.int64) == want_d[:nq].view(np.int64)).all(), (metric, nq)
                ids, d = fx.search_raw(q[:20], 100)  # k = 100: kk = 128 candidates per query
                wi, wd = orc.exact_knn(rows, q[:20], 100, code)
                assert ids.tolist() == wi.tolist() and (d.view(np.int64) == wd.view(np.int64)).all()
            finally:
                _lib.set_mode(_lib.MODE_EXACT)


def test_flat_fast_2_to_8_queries_over_a_long_list():
    """In FAST mode 2..8 queries over a list of >= 1 GB take the levelled narrow scan over the digit images (half the bytes of
    the fp32 rows); one query stays on the exact row stream.  Same bits either way."""
    import torch

    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.flat import FlatIndex

    _lib.check(_lib.lib().hb_init(0))
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    n, d = 400_000, 768  # 1.23 GB of fp32 rows
    c = torch.randn((800, d), generator=g, device=dev)
    rows = (c[torch.randint(0, 800, (n,), generator=g, device=dev)] + 0.1 * torch.randn((n, d), generator=g, device=dev)).contiguous()
    q = (c[torch.randint(0, 800, (8,), generator=g, device=dev)] + 0.1 * torch.randn((8, d), generator=g, device=dev)).contiguous()
    with FlatIndex(rows) as fx:
        want_i, want_d = fx.search_raw(q, 10)  # exact mode
        _lib.set_mode(_lib.MODE_FAST)
        _lib.set_option("profile", 1)
        try:
            for nq in (1, 2, 5, 8):
                ids, dist = fx.search_raw(q[:nq].contiguous(), 10)
                assert (ids == want_i[:nq]).all() and (dist.view(np.int64) == want_d[:nq].view(np.int64)).all(), nq
            assert _lib.get_stat("fast_queries") == 2 + 5 + 8  # the one-query call stayed on the exact path
        finally:
            _lib.set_option("profile", 0)
            _lib.set_mode(_lib.MODE_EXACT)
    s = 3
    oi, od = orc.exact_knn(rows.cpu().numpy(), q[:s].cpu().numpy(), 10)
    assert want_i[:s].tolist() == oi.tolist() and (want_d[:s].view(np.int64) == od.view(np.int64)).all()
    del rows
    torch.cuda.empty_cache()

"""GPU parity of HB_MODE_FAST (tcgen05 int8-digit candidate pass + fp64 re-score + proof, hb_fast.cuh).

1. the tensor-core scores are pinned BIT-EXACTLY against a numpy restatement of the quantisation and the
   integer digit products (any descriptor / swizzle / pipeline mistake shows up here);
2. FAST searches return the same ids and the same fp64 distance bits as the exact device path and the CPU
   oracle, whatever fraction of queries the proof accepts (the rest take the exact path)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hnsw_clj_b200 as pkg
    from hnsw_clj_b200 import _lib

    _lib.check(_lib.lib().hb_init(0))
    yield pkg
    _lib.set_mode(_lib.MODE_EXACT)
    _lib.set_option("fast_digits", 2)


def same_bits(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool((a.view(np.int64) == b.view(np.int64)).all())


def clustered(n, d, seed, centres=16, noise=0.1):
    r = np.random.default_rng(seed)
    c = r.standard_normal((centres, d))
    return (c[r.integers(0, centres, n)] + noise * r.standard_normal((n, d))).astype(np.float32)


# ---- numpy restatement of hb_fastprep.cu's quantisation and hb_tc.cu's digit products ---------------------
def quantise(x, ns):
    x = x.astype(np.float64)
    amax = np.abs(x).max(axis=1)
    qmax = 32000.0 if ns == 2 else 8000000.0
    u = np.where(amax > 0, amax / qmax, 1.0)
    m = np.rint(x * (1.0 / u)[:, None]).astype(np.int64)
    digs = []
    rest = m
    for _ in range(ns - 1):
        lo = ((rest & 255) ^ 128) - 128  # signed low byte
        digs.append(lo)
        rest = (rest - lo) >> 8
    digs.append(rest)
    return u, digs[::-1]  # most significant first


def emulate_scores(rows, queries, ns, cosine):
    ur, dr = quantise(rows, ns)
    uq, dq = quantise(queries, ns)
    c = [np.zeros((len(queries), len(rows)), dtype=np.int64) for _ in range(3)]
    for a in range(ns):
        for b in range(ns):
            if a + b <= 2:
                c[a + b] += dq[a] @ dr[b].T
    # epilogue of hb_tc.cu: lo = c1 + (c2 >> 8) in int32; s = fmaf(f(c0), 256, f(lo)); v = s * (rs * 256)
    lo = (c[1] + (c[2] >> 8)).astype(np.int32)
    s = (c[0].astype(np.float32).astype(np.float64) * 256.0 + lo.astype(np.float32).astype(np.float64)).astype(np.float32)
    rn = orc.row_norms(rows) if cosine else np.ones(len(rows))
    rs = (ur * (1.0 / rn) * (1.0 if ns == 2 else 65536.0)).astype(np.float32)
    v = (s.astype(np.float64) * (rs.astype(np.float64) * 256.0)[None, :]).astype(np.float32)
    qn = orc.row_norms(queries) if cosine else np.ones(len(queries))
    return v, uq * (1.0 / qn)


def fast_scores(hb, ix, queries, n):
    from hnsw_clj_b200 import _lib

    nq = len(queries)
    nu, nt = -(-nq // 128), -(-n // 128)
    out = np.empty((nu, nt, 128, 128), dtype=np.float32)
    scale = np.empty(nq, dtype=np.float64)
    eps = np.empty(nq, dtype=np.float64)
    _lib.check(_lib.lib().hb_fast_scores(ix._h, _lib.ptr(queries), _lib.dtype_code(queries), nq, _lib.ptr(out), _lib.ptr(scale),
                                         _lib.ptr(eps)))
    dense = out.transpose(0, 2, 1, 3).reshape(nu * 128, nt * 128)[:nq, :n]
    return dense, scale, eps


@pytest.mark.parametrize("ns", [2, 3])
@pytest.mark.parametrize("n,d,nq,metric", [(300, 768, 130, "cosine"), (128, 128, 1, "cosine"), (1000, 100, 257, "ip"),
                                           (129, 40, 5, "cosine"), (2048, 768, 128, "ip")])
def test_tensor_core_scores_bit_exact(hb, ns, n, d, nq, metric):
    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.flat import FlatIndex

    _lib.set_option("fast_digits", ns)
    rows = clustered(n, d, 11 * n + d)
    queries = clustered(nq, d, 7 * n + d + 1)
    with FlatIndex(rows, metric) as ix:
        got, scale, eps = fast_scores(hb, ix, queries, n)
    want, want_scale = emulate_scores(rows, queries, ns, metric == "cosine")
    assert (got.view(np.int32) == want.view(np.int32)).all(), f"{(got != want).sum()} of {got.size} scores differ"
    assert np.allclose(scale, want_scale, rtol=1e-15)
    # the a-priori bound really bounds the error of the candidate scores
    exact = queries.astype(np.float64) @ rows.astype(np.float64).T
    if metric == "cosine":
        exact = exact / orc.row_norms(queries)[:, None] / orc.row_norms(rows)[None, :]
    err = np.abs(got.astype(np.float64) * scale[:, None] - exact)
    assert (err <= eps[:, None]).all()
    if metric == "cosine":
        assert eps.max() < (2e-3 if ns == 2 else 1e-4)


@pytest.mark.parametrize("ns", [2, 3])
@pytest.mark.parametrize("n,d,nq,k,metric", [(3000, 96, 70, 10, "cosine"), (2000, 768, 200, 10, "cosine"), (129, 5, 3, 1, "cosine"),
                                             (5000, 128, 300, 48, "ip"), (40, 16, 9, 10, "cosine"), (700, 64, 130, 10, "ip"),
                                             (20000, 128, 200, 100, "ip"), (9000, 96, 130, 64, "cosine"), (300, 64, 40, 112, "ip")])
def test_flat_fast_equals_exact(hb, ns, n, d, nq, k, metric):
    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.flat import FlatIndex

    rows = clustered(n, d, n + d, centres=8)
    queries = clustered(nq, d, n + d + 1, centres=8)
    with FlatIndex(rows, metric) as ix:
        _lib.set_mode(_lib.MODE_EXACT)
        eids, edist = ix.search_raw(queries, k)
        _lib.set_option("fast_digits", ns)
        _lib.set_option("profile", 1)
        _lib.set_mode(_lib.MODE_FAST)
        try:
            fids, fdist = ix.search_raw(queries, k)
        finally:
            _lib.set_mode(_lib.MODE_EXACT)
        served, fell = _lib.get_stat("fast_queries"), _lib.get_stat("fast_fallbacks")
        _lib.set_option("profile", 0)
    assert fids.tolist() == eids.tolist()
    assert same_bits(fdist, edist)
    assert served == (nq if nq > 8 else 0)  # batches of <= 8 queries take the HBM-bound exact scan (smallscan_kernel)
    if metric == "cosine" and n >= 1000:
        assert fell <= 0.5 * nq, f"{fell} of {nq} queries fell back to the exact path"
    if metric == "cosine":
        oids, odist = orc.exact_knn(rows, queries, k)
        assert fids.tolist() == oids.tolist() and same_bits(fdist, odist)


@pytest.mark.parametrize("ns", [2, 3])
@pytest.mark.parametrize("n,d,nlist,nprobe,nq,k", [(6000, 64, 32, 8, 300, 10), (20000, 768, 512, 16, 700, 10),
                                                   (3000, 32, 24, 4, 50, 10), (4000, 128, 300, 40, 129, 5)])
def test_ivf_fast_equals_exact_and_oracle(hb, ns, n, d, nlist, nprobe, nq, k):
    from hnsw_clj_b200 import _lib, ivf_flat

    rows = clustered(n, d, n + nlist, centres=max(8, nlist // 2))
    queries = clustered(nq, d, n + nlist + 1, centres=max(8, nlist // 2))
    ix = ivf_flat.build_index(rows, num_partitions=nlist, max_iterations=2)
    try:
        _lib.set_mode(_lib.MODE_EXACT)
        eids, edist = ix.search_raw(queries, k, nprobe)
        _lib.set_option("fast_digits", ns)
        _lib.set_option("profile", 1)
        _lib.set_mode(_lib.MODE_FAST)
        try:
            fids, fdist = ix.search_raw(queries, k, nprobe)
        finally:
            _lib.set_mode(_lib.MODE_EXACT)
        served, fell = _lib.get_stat("fast_queries"), _lib.get_stat("fast_fallbacks")
        _lib.set_option("profile", 0)
        cents, asg = ix.export()
    finally:
        ix.close()
    assert fids.tolist() == eids.tolist()
    assert same_bits(fdist, edist)
    assert served == nq
    print(f"ivf fast ns={ns} n={n} nlist={nlist}: {int(fell)} of {nq} queries fell back")
    oids, odist = orc.ivf_search(rows, cents, asg, queries[:64], k, nprobe)
    assert fids[:64].tolist() == oids.tolist() and same_bits(fdist[:64], odist)


def test_flat_fast_bf16_top100_inner_product(hb):
    """BASELINE configs[2] in small: bf16 rows, inner product, top-100 -- candidate pass with 128 re-scored candidates."""
    import torch

    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.flat import FlatIndex

    r = np.random.default_rng(77)
    rows = torch.from_numpy(r.standard_normal((30000, 256)).astype(np.float32)).to(torch.bfloat16)
    q = torch.from_numpy(r.standard_normal((150, 256)).astype(np.float32)).to(torch.bfloat16).float()
    want_ids, want_d = orc.exact_knn(rows.float().numpy(), q.numpy(), 100, orc.IP)
    with FlatIndex(rows.cuda(), distance_fn="ip") as ix:
        _lib.set_option("profile", 1)
        _lib.set_mode(_lib.MODE_FAST)
        try:
            ids, dist = ix.search_raw(q.cuda(), 100)
        finally:
            _lib.set_mode(_lib.MODE_EXACT)
        served, fell = _lib.get_stat("fast_queries"), _lib.get_stat("fast_fallbacks")
        _lib.set_option("profile", 0)
    if hasattr(ids, "cpu"):
        ids, dist = ids.cpu().numpy(), dist.cpu().numpy()
    assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)
    assert served == 150
    assert fell <= 75, f"{fell} of 150 queries fell back to the exact path"


@pytest.mark.parametrize("n,d,nlist", [(20000, 64, 300), (5000, 768, 256), (40000, 32, 1000)])
def test_kmeans_assign_fast_equals_exact_and_oracle(hb, n, d, nlist):
    """assign-to-nearest-centroid through the tensor-core candidate pass (k = 1 over the centroids): same assignments as the
    fp64 kernel and the oracle, ties to the lowest centroid index."""
    from hnsw_clj_b200 import _lib, ivf_flat

    rows = clustered(n, d, 3 * n + d, centres=nlist // 2)
    r = np.random.default_rng(n)
    cents = rows[r.choice(n, nlist, replace=False)].astype(np.float64) + 1e-3 * r.standard_normal((nlist, d))
    cents[7] = cents[3]  # duplicate centroid: every row nearest to it must go to the lower index
    want = orc.assign(rows, cents)
    exact = ivf_flat.assign_to_nearest_centroid(rows, cents)
    _lib.set_option("profile", 1)
    _lib.set_mode(_lib.MODE_FAST)
    try:
        fast = ivf_flat.assign_to_nearest_centroid(rows, cents)
    finally:
        _lib.set_mode(_lib.MODE_EXACT)
    served, fell = _lib.get_stat("fast_queries"), _lib.get_stat("fast_fallbacks")
    _lib.set_option("profile", 0)
    assert exact.tolist() == want.tolist()
    assert fast.tolist() == want.tolist()
    assert served == n
    assert (fast != 7).all()
    print(f"k-means assign fast n={n} nlist={nlist}: {int(fell)} rows fell back")


def test_ivf_build_in_fast_mode_is_bit_identical(hb):
    """The whole build (k-means++ seeds, Lloyd rounds, final assignment) with FAST-mode assignment passes gives the centroids
    and assignments of the exact build, bit for bit."""
    from hnsw_clj_b200 import _lib, ivf_flat

    rows = clustered(30000, 96, 123, centres=200)
    a = ivf_flat.build_index(rows, num_partitions=300, max_iterations=3)
    ca, aa = a.export()
    a.close()
    _lib.set_mode(_lib.MODE_FAST)
    try:
        b = ivf_flat.build_index(rows, num_partitions=300, max_iterations=3)
    finally:
        _lib.set_mode(_lib.MODE_EXACT)
    cb, ab = b.export()
    b.close()
    assert aa.tolist() == ab.tolist()
    assert same_bits(ca, cb)


def test_fast_duplicates_fall_back_and_stay_exact(hb):
    """Exact duplicates make the k-th / (k+1)-th gap zero: the proof must refuse and the exact path answer."""
    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.flat import FlatIndex

    base = clustered(30, 64, 5)
    rows = np.concatenate([base] * 100)  # every row 100 times: the tie spans the candidate cut, order is by row index
    queries = base[:20] + np.float32(0.01)
    with FlatIndex(rows, "cosine") as ix:
        eids, edist = ix.search_raw(queries, 10)
        _lib.set_option("profile", 1)
        _lib.set_mode(_lib.MODE_FAST)
        try:
            fids, fdist = ix.search_raw(queries, 10)
        finally:
            _lib.set_mode(_lib.MODE_EXACT)
        fell = _lib.get_stat("fast_fallbacks")
        _lib.set_option("profile", 0)
    assert fids.tolist() == eids.tolist() and same_bits(fdist, edist)
    assert fell == 20
    oids, _ = orc.exact_knn(rows, queries, 10)
    assert fids.tolist() == oids.tolist()


def test_ivf_fast_cross_list_ties_follow_probe_rank(hb):
    """The FAST coarse stage proves the probed SET and leaves the probe ORDER approximate (set_only); the reference breaks
    a distance tie between rows of different lists by probe rank (ivf_flat.clj:281-294), so such queries must be handed
    to the exact path.  Rows and their duplicates are planted in different lists (hb_ivf_import takes any assignment)."""
    from hnsw_clj_b200 import _lib, ivf_flat

    n, d, nlist, nprobe, k = 12000, 64, 256, 8, 3  # nlist >= 256: the coarse stage runs the candidate pass
    base = clustered(n, d, 77, centres=150)
    cents, asg = orc.kmeans(base, nlist, iters=3, seed=42)
    dmat = np.stack([[orc.cosine_distance(base[i].astype(np.float64), c) for c in cents] for i in range(200)])
    second = np.argsort(dmat, axis=1, kind="stable")[:, 1].astype(np.int32)
    rows = np.concatenate([base, base[:200]])  # row n + i duplicates row i ...
    asg2 = np.concatenate([asg, asg[:200]]).astype(np.int32)  # ... and takes its place in the nearest list,
    asg2[:200] = second  # while row i itself moves to its second-nearest list: probe rank and row order now disagree
    queries = (base[:200] + np.float32(1e-3)).astype(np.float32)
    want_ids, want_d = orc.ivf_search(rows, cents, asg2, queries, k, nprobe)
    assert (want_ids[:, :2] % n == np.arange(200)[:, None]).all() and (want_d[:, 0] == want_d[:, 1]).all()
    assert (want_ids[:, 0] >= n).sum() > 150  # the pair leads every result, the copy in the nearer list first
    ix = ivf_flat.import_index(rows, cents, asg2)
    try:
        eids, edist = ix.search_raw(queries, k, nprobe)
        _lib.set_option("profile", 1)
        _lib.set_mode(_lib.MODE_FAST)
        try:
            fids, fdist = ix.search_raw(queries, k, nprobe)
        finally:
            _lib.set_mode(_lib.MODE_EXACT)
        fell = _lib.get_stat("fast_fallbacks")
        _lib.set_option("profile", 0)
    finally:
        ix.close()
    assert eids.tolist() == want_ids.tolist() and same_bits(edist, want_d)
    assert fids.tolist() == want_ids.tolist() and same_bits(fdist, want_d)
    assert fell == 200


@pytest.mark.parametrize("data", ["clustered", "gaussian", "boundary"])
def test_ivf_fast_probe_pruning_is_exact(hb, data):
    """Probed lists whose rows cannot reach a query's candidate threshold (cos(angle(q, centroid) - list radius) below it)
    are dropped before the scan.  Clustered rows: nearly every non-nearest list goes; structureless rows: the true
    neighbours sit in many lists and none of those may go; queries half-way between two clusters need both lists."""
    from hnsw_clj_b200 import _lib, ivf_flat

    n, d, nlist, nprobe, nq, k = 20000, 96, 256, 16, 500, 10
    r = np.random.default_rng(17)
    if data == "gaussian":
        rows = r.standard_normal((n, d)).astype(np.float32)
        queries = r.standard_normal((nq, d)).astype(np.float32)
    else:
        c = r.standard_normal((300, d))
        rows = (c[r.integers(0, 300, n)] + 0.1 * r.standard_normal((n, d))).astype(np.float32)
        if data == "clustered":
            queries = (c[r.integers(0, 300, nq)] + 0.1 * r.standard_normal((nq, d))).astype(np.float32)
        else:
            a, b = r.integers(0, 300, nq), r.integers(0, 300, nq)
            queries = (0.5 * (c[a] + c[b]) + 0.05 * r.standard_normal((nq, d))).astype(np.float32)
    ix = ivf_flat.build_index(rows, num_partitions=nlist, max_iterations=3)
    try:
        eids, edist = ix.search_raw(queries, k, nprobe)
        _lib.set_mode(_lib.MODE_FAST)
        try:
            _lib.set_option("fast_prune", 0)
            uids, udist = ix.search_raw(queries, k, nprobe)
            _lib.set_option("fast_prune", 1)
            _lib.set_option("profile", 1)
            fids, fdist = ix.search_raw(queries, k, nprobe)
            pruned, total = _lib.get_stat("fast_pruned_pairs"), _lib.get_stat("fast_probe_pairs")
            fell = _lib.get_stat("fast_fallbacks")
            _lib.set_option("profile", 0)
        finally:
            _lib.set_option("fast_prune", 1)
            _lib.set_mode(_lib.MODE_EXACT)
        cents, asg = ix.export()
    finally:
        ix.close()
    assert total == nq * nprobe
    assert fids.tolist() == eids.tolist() and same_bits(fdist, edist)
    assert uids.tolist() == eids.tolist() and same_bits(udist, edist)
    oids, odist = orc.ivf_search(rows, cents, asg, queries[:64], k, nprobe)
    assert fids[:64].tolist() == oids.tolist() and same_bits(fdist[:64], odist)
    print(f"probe pruning on {data} rows: {int(pruned)} of {int(total)} (query, list) pairs dropped, {int(fell)} fallbacks")
    if data == "clustered":
        assert pruned > 0.5 * total

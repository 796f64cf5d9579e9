"""GPU parity of the batched HNSW search (hb_search on an uploaded graph) against the oracle's traversal of the
SAME graph (src/hnsw/ultra_fast.clj:151-212, 346-374): ids and fp64 distance bits, ties included.  The graph is
built on the host by the oracle (graph mutation is out of scope; SURVEY §8c: level RNG seeded, neighbour sets in
insertion order) and uploaded level by level."""
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


def rows_of(n, d, seed, unit=True, dup=0):
    r = np.random.default_rng(seed)
    x = r.standard_normal((n, d))
    if unit:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    x = x.astype(np.float32)
    if dup:  # exact duplicates: equal distances exercise the PriorityQueue tie behaviour
        src = r.integers(0, n, dup)
        dst = r.integers(0, n, dup)
        x[dst] = x[src]
    return x


def upload(g, rows, metric="cosine"):
    from hnsw_clj_b200.ultra_fast import HnswIndex

    adjacency = [g.export_level(l) for l in range(g.max_level + 1)]
    return HnswIndex(rows, g.levels(), g.entry, adjacency, distance_fn=metric)


def same_bits(a, b):
    return a.shape == b.shape and bool((np.asarray(a).view(np.int64) == np.asarray(b).view(np.int64)).all())


@pytest.fixture(scope="module")
def graph_case():
    rows = rows_of(3000, 64, 7, dup=40)
    g = orc.Hnsw(rows, M=8, ef_construction=60, level_seed=42)
    return rows, g


@pytest.mark.parametrize("k,ef", [(10, 0), (10, 128), (1, 1), (60, 0), (100, 300)])
def test_hnsw_search_matches_oracle_traversal(hb, graph_case, k, ef):
    rows, g = graph_case
    q = np.concatenate([rows_of(150, 64, 8), rows[:50]])
    want_ids, want_d = g.search(q, k, ef)
    with upload(g, rows) as ix:
        ids, dist = ix.search_raw(q, k, ef)
    assert ids.tolist() == want_ids.tolist()
    assert same_bits(dist, want_d)


def test_hnsw_search_768_dims_fp64_queries_and_recall(hb):
    rows = rows_of(1200, 768, 11)
    g = orc.Hnsw(rows, M=16, ef_construction=80, level_seed=1)
    q = rows_of(64, 768, 12)
    want_ids, want_d = g.search(q, 10, 128)
    with upload(g, rows) as ix:
        ids, dist = ix.search_raw(q, 10, 128)
        ids64, dist64 = ix.search_raw(q.astype(np.float64), 10, 128)
    assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)
    assert ids64.tolist() == want_ids.tolist() and same_bits(dist64, want_d)
    ex, _ = orc.exact_knn(rows, q, 10)
    assert orc.recall(ids, ex) > 0.9


def test_hnsw_search_euclidean(hb):
    rows = rows_of(1500, 48, 21, unit=False)
    g = orc.Hnsw(rows, metric=orc.L2, M=8, ef_construction=50, level_seed=3)
    q = rows_of(90, 48, 22, unit=False)
    want_ids, want_d = g.search(q, 10, 0)
    with upload(g, rows, "euclidean") as ix:
        ids, dist = ix.search_raw(q, 10, 0)
    assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)


def test_hnsw_candidate_queue_overflow_path(hb, graph_case):
    """A 16-slot shared-memory candidate queue overflows for most queries: they are re-run with the queue in
    global memory and must give the same results."""
    from hnsw_clj_b200 import _lib

    rows, g = graph_case
    q = rows_of(70, 64, 9)
    want_ids, want_d = g.search(q, 10, 100)
    with upload(g, rows) as ix:
        _lib.set_option("profile", 1)
        _lib.set_option("hnsw_cand_cap", 16)
        try:
            ids, dist = ix.search_raw(q, 10, 100)
            over = _lib.get_stat("hnsw_overflows")
            scored = _lib.get_stat("hnsw_scored")
        finally:
            _lib.set_option("hnsw_cand_cap", 0)
            _lib.set_option("profile", 0)
    assert over > 0 and scored > 0
    assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)


def test_hnsw_edge_cases(hb):
    # a single node; k > n pads with -1 / +inf (test/hnsw/core_test.clj:90-96)
    rows = rows_of(1, 16, 1)
    g = orc.Hnsw(rows, M=4, ef_construction=10, level_seed=5)
    q = rows_of(3, 16, 2)
    want_ids, want_d = g.search(q, 5, 0)
    with upload(g, rows) as ix:
        ids, dist = ix.search_raw(q, 5, 0)
        assert ix.search_raw(q[:0], 5, 0)[0].shape == (0, 5)
    assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)
    assert (ids[:, 1:] == -1).all() and np.isinf(dist[:, 1:]).all()
    # a few nodes, odd dimension (unaligned rows take the scalar load path)
    rows = rows_of(37, 5, 3)
    g = orc.Hnsw(rows, M=4, ef_construction=10, level_seed=6)
    q = rows_of(9, 5, 4)
    want_ids, want_d = g.search(q, 50, 0)
    with upload(g, rows) as ix:
        ids, dist = ix.search_raw(q, 50, 0)
    assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)


def test_search_knn_mirror_returns_maps(hb, graph_case):
    from hnsw_clj_b200 import ultra_fast

    rows, g = graph_case
    want_ids, want_d = g.search(rows[:1], 5, 0)
    with upload(g, rows) as ix:
        res = ultra_fast.search_knn(ix, rows[0], 5)
    assert [r["id"] for r in res] == want_ids[0].tolist()
    assert [r["distance"] for r in res] == want_d[0].tolist()


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_bulk_knn_graph_and_traversal(hb, mode):
    """The device-built bulk graph (ultra_fast.bulk_knn_graph): level-l neighbour lists are the m nearest members of the
    level in the oracle's (distance, row) order, and the batched search over it equals the oracle traversal of the same
    graph (orc.Hnsw.from_graph)."""
    from hnsw_clj_b200 import _lib
    from hnsw_clj_b200.ultra_fast import HnswIndex, bulk_knn_graph

    rows = rows_of(2500, 64, 11, dup=20)
    _lib.set_mode(_lib.MODE_FAST if mode == "fast" else _lib.MODE_EXACT)
    try:
        levels, entry, adjacency = bulk_knn_graph(rows, M=8, level_seed=3, batch=1000)
    finally:
        _lib.set_mode(_lib.MODE_EXACT)
    assert levels.max() == len(adjacency) - 1 and levels[entry] == levels.max()
    assert entry == int(np.flatnonzero(levels == levels.max())[0])
    for l, (off, ids) in enumerate(adjacency[:3]):
        members = np.flatnonzero(levels >= l)
        m = 16 if l == 0 else 8
        kk = min(m + 1, len(members))
        want_ids, _ = orc.exact_knn(rows[members], rows[members], kk)
        for j in (0, 1, len(members) // 2, len(members) - 1):
            w = [int(members[x]) for x in want_ids[j] if x != j][: kk - 1]
            if len(w) < kk - 1:
                w = [int(members[x]) for x in want_ids[j][: kk - 1]]
            node = int(members[j])
            assert ids[off[node]:off[node + 1]].tolist() == w, (l, j)
        absent = np.flatnonzero(levels < l)
        if len(absent):
            assert off[absent[0]] == off[absent[0] + 1]
    g = orc.Hnsw.from_graph(rows, levels, entry, adjacency, M=8)
    q = np.concatenate([rows_of(120, 64, 12), rows[:40]])
    want_ids, want_d = g.search(q, 10, 96)
    with HnswIndex(rows, levels, entry, adjacency) as ix:
        ids, dist = ix.search_raw(q, 10, 96)
    assert ids.tolist() == want_ids.tolist() and same_bits(dist, want_d)
    from hnsw_clj_b200.flat import recall_at_k

    ex_ids, _ = orc.exact_knn(rows, q, 10)
    assert recall_at_k(ids, ex_ids) > 0.5

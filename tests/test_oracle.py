"""Pins the CPU oracle against every known-answer the reference's own tests hold for the path
(test/hnsw/core_test.clj:9-31) and against published java.util.Random(42) values, then checks the
oracle's internal consistency (literal vs incremental k-means++, numpy twin of the arithmetic)."""
import math

import numpy as np
import pytest

from oracle import oracle as orc


# ---- reference KATs: test/hnsw/core_test.clj:9-20 (test-vector-distance) ---------------------
def test_euclid_kats():
    assert orc.euclidean_distance([1, 2, 3], [1, 2, 3]) == 0.0
    assert orc.euclidean_distance([0, 0], [3, 4]) == 5.0
    assert abs(orc.euclidean_distance([1, 2, 3], [4, 5, 6]) - 5.196152422706632) < 1e-5
    assert orc.euclidean_distance([1, 2, 3], [4, 5, 6]) == 5.196152422706632


# ---- reference KATs: test/hnsw/core_test.clj:22-31 (test-cosine-distance) --------------------
def test_cosine_kats():
    assert orc.cosine_distance([1, 2, 3], [1, 2, 3]) < 0.001
    assert abs(orc.cosine_distance([1, 0], [-1, 0]) - 2.0) < 1e-3
    assert abs(orc.cosine_distance([1, 2, 3], [4, 5, 6]) - 0.0253) < 0.01
    # SURVEY Appendix C
    assert orc.cosine_distance([1, 2, 3], [4, 5, 6]) == 0.025368153802923787
    assert orc.cosine_distance([1, 0], [-1, 0]) == 2.0
    # zero-norm guard, src/hnsw/ultra_fast.clj:92-95
    assert orc.cosine_distance([0, 0, 0], [1, 2, 3]) == 1.0
    assert orc.cosine_distance_direct([0, 0, 0], [1, 2, 3]) == 1.0


def test_pairwise_matches_python_left_fold():
    rng = np.random.default_rng(0)
    a = rng.standard_normal(768).astype(np.float32).astype(np.float64)
    b = rng.standard_normal(768).astype(np.float32).astype(np.float64)
    dot = n1 = n2 = 0.0
    for x, y in zip(a.tolist(), b.tolist()):  # python floats: IEEE fp64, no FMA
        dot = dot + x * y
        n1 = n1 + x * x
        n2 = n2 + y * y
    assert orc.cosine_distance(a, b) == 1.0 - dot / (math.sqrt(n1) * math.sqrt(n2))
    assert orc.dot(a, b) == dot
    assert orc.norm(a) == math.sqrt(n1)
    s = 0.0
    for x, y in zip(a.tolist(), b.tolist()):
        s = s + (x - y) * (x - y)
    assert orc.euclidean_distance(a, b) == math.sqrt(s)


# ---- java.util.Random(42): published values ------------------------------
def test_java_random_kats():
    r = orc.JavaRandom(42)
    assert [r.next_int() for _ in range(3)] == [-1170105035, 234785527, -1360544799]
    r = orc.JavaRandom(42)
    assert [r.next_int(10) for _ in range(5)] == [0, 3, 8, 4, 0]
    r = orc.JavaRandom(42)
    assert [r.next_int(100) for _ in range(5)] == [30, 63, 48, 84, 70]  # SURVEY Appendix C (restated twice, agree)
    r = orc.JavaRandom(42)
    assert r.next_double() == 0.7275636800328681
    assert r.next_double() == 0.6832234717598454
    assert orc.JavaRandom(42).next_int(31173) == 9197
    assert orc.JavaRandom(42).next_int(1000000) == 431130
    r = orc.JavaRandom(42)
    g = r.next_gaussian()
    assert abs(g - 1.1419053154730547) < 1e-15


def test_exact_knn_matches_numpy_twin():
    rng = np.random.default_rng(1)
    rows = rng.standard_normal((300, 48)).astype(np.float32)
    q = rng.standard_normal((7, 48)).astype(np.float32)
    ids, dist = orc.exact_knn(rows, q, 10)
    R, Q = rows.astype(np.float64), q.astype(np.float64)
    for qi in range(7):
        d = np.empty(300)
        for i in range(300):
            dot = nv = nq = 0.0
            for x, y in zip(R[i].tolist(), Q[qi].tolist()):
                dot = dot + x * y
                nv = nv + x * x
                nq = nq + y * y
            d[i] = 1.0 - dot / (math.sqrt(nv) * math.sqrt(nq))
        order = np.argsort(d, kind="stable")[:10]
        assert ids[qi].tolist() == order.tolist()
        assert dist[qi].tolist() == d[order].tolist()


def test_exact_knn_edge_cases():
    rows = np.eye(4, dtype=np.float32)
    q = np.eye(4, dtype=np.float32)[:2]
    ids, dist = orc.exact_knn(rows, q, 6)  # k > n: test/hnsw/core_test.clj:90-96
    assert ids[0, 0] == 0 and (ids[:, 4:] == -1).all() and np.isinf(dist[:, 4:]).all()
    # ties keep row order (stable sort): rows 1..3 are all at distance 1.0 from q0
    assert ids[0].tolist()[:4] == [0, 1, 2, 3]
    ids, dist = orc.exact_knn(np.zeros((0, 4), np.float32), q, 3)  # empty index -> nothing
    assert (ids == -1).all()


def test_kmeanspp_incremental_equals_literal():
    rows = orc.gen_dataset(400, 16, orc.CLUSTERED, num_clusters=8, noise=0.1, seed=7).astype(np.float32)
    a = orc.kmeanspp_init(rows, 12, seed=42)
    b = orc.kmeanspp_init(rows, 12, seed=42, literal=True)
    assert a.tolist() == b.tolist()
    assert a[0] == orc.JavaRandom(42).next_int(400)
    a2 = orc.kmeanspp_init(rows, 12, metric=orc.L2, seed=42)
    b2 = orc.kmeanspp_init(rows, 12, metric=orc.L2, seed=42, literal=True)
    assert a2.tolist() == b2.tolist()


def test_kmeans_and_ivf_search_properties():
    rows = orc.gen_dataset(1500, 24, orc.CLUSTERED, num_clusters=12, noise=0.15, seed=3).astype(np.float32)
    q = orc.gen_dataset(20, 24, orc.CLUSTERED, num_clusters=12, noise=0.15, seed=4).astype(np.float32)
    cents, asg = orc.kmeans(rows, 16, iters=5, seed=42)
    assert asg.min() >= 0 and asg.max() < 16
    # final assignment is consistent with the centroids
    assert orc.assign(rows, cents).tolist() == asg.tolist()
    # probing every list == exact search (ids and distance bits)
    ids_all, d_all = orc.ivf_search(rows, cents, asg, q, 10, 16)
    ids_ex, d_ex = orc.exact_knn(rows, q, 10)
    assert ids_all.tolist() == ids_ex.tolist()
    assert d_all.tolist() == d_ex.tolist()
    ids4, d4, probes = orc.ivf_search(rows, cents, asg, q, 10, 4, return_probes=True)
    assert probes.shape == (20, 4)
    assert (np.diff(d4, axis=1) >= 0).all()  # ascending, test/hnsw/integration_test.clj:134-136
    assert 0.5 < orc.recall(ids4, ids_ex) <= 1.0


def test_hnsw_oracle_recall_and_shape():
    rows = orc.gen_dataset(600, 16, orc.UNIT, seed=5).astype(np.float32)
    g = orc.Hnsw(rows, M=8, ef_construction=60, level_seed=42)
    q = rows[:25]
    ids, dist = g.search(q, 10)
    # self-query (test/hnsw/core_test.clj:46): the graph is approximate, most queries find themselves
    assert (ids[:, 0] == np.arange(25)).mean() >= 0.8
    assert (dist[:, 0] < 0.01).mean() >= 0.8
    ex, _ = orc.exact_knn(rows, q, 10)
    assert orc.recall(ids, ex) > 0.8
    off, nb = g.export_level(0)
    deg = np.diff(off)
    assert deg.max() <= 16  # max-M = 2M at layer 0, src/hnsw/ultra_fast.clj:131,252
    # gather_score of (query, neighbour) pairs equals the pairwise function
    pq = np.array([0, 1, 2], np.int32)
    pr = np.array([5, 6, 7], np.int32)
    s = orc.gather_score(rows, q, pq, pr)
    for t in range(3):
        assert s[t] == orc.cosine_distance(q[pq[t]].astype(np.float64), rows[pr[t]].astype(np.float64))


def test_hnsw_oracle_import_round_trip():
    """from_graph (used to run the oracle traversal over graphs built elsewhere) reproduces the searches of the graph
    it was exported from."""
    rows = np.random.default_rng(5).standard_normal((600, 32)).astype(np.float32)
    g = orc.Hnsw(rows, M=6, ef_construction=40, level_seed=7)
    adjacency = [g.export_level(l) for l in range(g.max_level + 1)]
    g2 = orc.Hnsw.from_graph(rows, g.levels(), g.entry, adjacency, M=6)
    assert g2.entry == g.entry and g2.max_level == g.max_level
    for l in range(g.max_level + 1):
        o1, i1 = g.export_level(l)
        o2, i2 = g2.export_level(l)
        assert o1.tolist() == o2.tolist() and i1.tolist() == i2.tolist()
    q = rows[:40] + 0.01
    a_ids, a_d = g.search(q, 5, 30)
    b_ids, b_d = g2.search(q, 5, 30)
    assert a_ids.tolist() == b_ids.tolist() and (a_d.view(np.int64) == b_d.view(np.int64)).all()


def test_lightning_seeding_matches_a_literal_python_walk():
    """orc.lightning_seeds against a direct Python restatement of src/hnsw/ann/partition/lightning.clj:86-109
    (min over ALL chosen centroids, weights d_i, first i with cumsum + d_i >= r)."""
    r = np.random.default_rng(3)
    rows = r.standard_normal((120, 12)).astype(np.float32)
    got = orc.lightning_seeds(rows, 7, seed=42)
    state = (42 ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def nxt(bits):
        nonlocal state
        state = (state * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        v = state >> (48 - bits)
        return v - (1 << bits) if bits == 32 and v >= (1 << 31) else v

    def next_int(bound):
        rr = nxt(31)
        m = bound - 1
        if bound & m == 0:
            return (bound * rr) >> 31
        u = rr
        while True:
            rr = u % bound
            if u - rr + m < (1 << 31):
                return rr
            u = nxt(31)

    def next_double():
        return ((nxt(26) << 27) + nxt(27)) * (1.0 / (1 << 53))

    chosen = [next_int(len(rows))]
    for _ in range(6):
        dist = [min([1.7976931348623157e308] + [orc.cosine_distance(rows[i].astype(np.float64), rows[c].astype(np.float64))
                                                 for c in chosen]) for i in range(len(rows))]
        s = 0.0
        for v in dist:
            s = s + v
        rr = next_double() * s
        cum, i = 0.0, 0
        while not (cum + dist[i] >= rr):
            cum = cum + dist[i]
            i += 1
        chosen.append(i)
    assert got.tolist() == chosen
    cents, asg = orc.lightning_build(rows, 7)
    seeds = rows[got].astype(np.float64)
    want_a = [int(np.argmin([orc.cosine_distance(rows[i].astype(np.float64), s_) for s_ in seeds])) for i in range(len(rows))]
    assert asg.tolist() == want_a
    for c in range(7):
        members = rows[asg == c].astype(np.float64)
        acc = np.zeros(12)
        for v in members:
            acc = acc + v
        assert (cents[c] == (acc / len(members) if len(members) else acc)).all()


def test_simd_lane_order_is_within_the_fp32_bound():
    """a4 (src/hnsw/simd.clj:73-115): the float[] cosine restated with 4 / 8 / 16 lanes, left-to-right or pairwise-tree lane
    sums, and the double[] cosine-distance-direct all agree within the north star's 1e-5 relative for fp32; the KATs of
    test/hnsw/core_test.clj:22-31 hold for it."""
    r = np.random.default_rng(11)
    for _ in range(20):
        a = r.standard_normal(768).astype(np.float32)
        b = (a + 0.5 * r.standard_normal(768)).astype(np.float32)
        ref = orc.cosine_distance_direct(a.astype(np.float64), b.astype(np.float64))
        for lanes in (4, 8, 16):
            for tree in (False, True):
                assert abs(orc.simd_cosine(a, b, lanes, tree) - ref) <= 1e-5 * abs(ref)
    assert abs(orc.simd_cosine([1, 2, 3], [4, 5, 6], 8) - 0.025368153802923787) < 1e-7   # scalar tail only (d < lanes)
    assert orc.simd_cosine([1, 0], [-1, 0], 4) == 2.0 and orc.simd_cosine([0, 0], [1, 2], 4) == 1.0
    assert orc.simd_euclidean([0, 0], [3, 4], 8) == 5.0
    # d a multiple of the lane count: no tail; the chunk sums are fp32
    a = np.full(16, 0.1, np.float32)
    assert orc.simd_dot(a, a, 8) == float(np.float32(8 * 0).item()) + 2 * float(sum_f32([np.float32(0.1) * np.float32(0.1)] * 8))


def sum_f32(xs):
    s = np.float32(0.0)
    for x in xs:
        s = np.float32(s + x)
    return s


def test_pcaf_matrix_is_the_java_gaussian_stream_and_matches_the_library():
    """create-random-projection (src/hnsw/ann/dimreduct/pcaf.clj:33-46): (float)(1/sqrt(t)) * (float) Random(42).nextGaussian();
    hb_pcaf_matrix is host arithmetic (no device needed) and must equal the oracle bit for bit."""
    m = orc.pcaf_matrix(768, 100)
    g = orc.JavaRandom(42)
    scale = np.float32(1.0 / np.sqrt(100.0))
    first = [np.float32(float(scale) * float(np.float32(g.next_gaussian()))) for _ in range(5)]
    assert m.reshape(-1)[:5].tolist() == [float(x) for x in first]
    assert abs(float(m[0, 0]) - 0.11419053) < 1e-7  # 1.1419053154730547 / 10
    from hnsw_clj_b200 import pcaf

    lib_m = pcaf.create_random_projection(768, 100)
    assert (lib_m.view(np.int32) == m.view(np.int32)).all()
    low = orc.pcaf_project(m, np.ones((1, 768), np.float32))
    assert low.shape == (1, 100) and np.isfinite(low).all()

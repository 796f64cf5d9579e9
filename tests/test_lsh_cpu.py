"""Hybrid LSH (src/hnsw/ann/hash/hybrid_lsh.clj), the parts that need no GPU: java.util.Random.nextGaussian and
StrictMath.log as restated in the oracle and in the library's host-side generator (hb_lsh_matrices), pinned by the
published outputs of `new Random(42).nextGaussian()`; the oracle's own LSH search against a plain numpy reading."""
import math

import numpy as np

from oracle import oracle as orc


def test_next_gaussian_known_answers():
    r = orc.JavaRandom(42)
    # new Random(42).nextGaussian() x 3 on any JVM (StrictMath makes it platform-independent)
    assert [r.next_gaussian() for _ in range(3)] == [1.1419053154730547, 0.9194079489827879, -0.9498666368908959]
    r = orc.JavaRandom(0)
    assert r.next_gaussian() == 0.8025330637390305  # new Random(0).nextGaussian()


def test_strict_log_is_fdlibm_not_libm():
    assert orc.strict_log(1.0) == 0.0 and orc.strict_log(2.0) == 0.6931471805599453
    assert orc.strict_log(0.0) == -math.inf and math.isnan(orc.strict_log(-1.0)) and orc.strict_log(math.inf) == math.inf
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.random(20000), rng.random(2000) * 1e-300, 1.0 + rng.random(2000) * 1e-7, rng.random(2000) * 1e300])
    worst = max(abs(orc.strict_log(float(x)) - math.log(float(x))) / max(abs(math.log(float(x))), 1e-300) for x in xs)
    assert worst < 3e-16  # within an ulp of libm everywhere, though not bit-identical (which is the point)


def test_library_generator_matches_the_oracle():
    """hb_lsh_matrices is host arithmetic (no device): the same stream as the oracle's, bit for bit."""
    from hnsw_clj_b200 import hybrid_lsh

    for d in (5, 96):
        got = hybrid_lsh.projection_matrices(d)
        want = orc.lsh_matrices(d)
        assert got.shape == (8, 64, d)
        assert (got.view(np.int64) == want.view(np.int64)).all()
    assert hybrid_lsh.projection_matrices(3)[0, 0, 0] == 1.1419053154730547


def test_oracle_lsh_search_against_numpy_reading():
    rng = np.random.default_rng(5)
    c = rng.standard_normal((6, 24))
    rows = (c[rng.integers(0, 6, 4000)] + 0.3 * rng.standard_normal((4000, 24))).astype(np.float32)
    q = (c[rng.integers(0, 6, 9)] + 0.3 * rng.standard_normal((9, 24))).astype(np.float32)
    M = orc.lsh_matrices(24)
    b = orc.lsh_hash(rows, M)
    qb = orc.lsh_hash(q, M)
    # the hash itself: sign of the fp64 projections, first 12 rows of each table
    R = rows.astype(np.float64)
    for t in (0, 7):
        bits = (R @ M[t, :12].T) >= 0
        assert (b[:, t] == (bits * (1 << np.arange(12))).sum(axis=1)).mean() > 0.999  # (matmul order differs in the last ulp)
    norms = orc.row_norms(rows)
    for multiprobe, probes, radius, mult in ((True, 6, 2, 2), (False, 2, 0, 3), (False, 3, 0, 2), (True, 8, 4, 2)):
        ids, dist = orc.lsh_search(rows, M, b, q, 10, probes, radius, multiprobe, mult)
        for qi in range(q.shape[0]):
            cand = []
            for t in range(min(probes, 8)):
                buckets = [(qb[qi, t], 10 * (2 if multiprobe else mult))]
                if multiprobe:
                    buckets += [((qb[qi, t] ^ (1 << bit)) & 4095, 10) for bit in range(min(radius, 12))]
                for bk, limit in buckets:
                    members = np.nonzero(b[:, t] == bk)[0]
                    dd = [orc.dot(rows[m], q[qi]) for m in members]
                    qn = orc.norm(q[qi])
                    hits = [(1.0 - dd[i] / (qn * norms[m]), int(m)) for i, m in enumerate(members)]
                    if len(hits) > limit:
                        hits = sorted(hits, key=lambda h: h[0])[:limit]  # Python's sort is stable
                    cand += hits
            seen, uniq = set(), []
            for h in cand:
                if h[1] not in seen:
                    seen.add(h[1])
                    uniq.append(h)
            uniq = sorted(uniq, key=lambda h: h[0])[:10]
            assert [h[1] for h in uniq] == [i for i in ids[qi].tolist() if i >= 0]
            assert [h[0] for h in uniq] == dist[qi][: len(uniq)].tolist()

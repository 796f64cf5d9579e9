"""kmeans-plus-plus-init at scale (hb_kpp.cu) is the reference's algorithm bit for bit (src/hnsw/ann/partition/ivf_flat.clj:32-60):
the chunked ordered sum returns the sequential loop's bits on adversarial weights, and the seeds equal the oracle's and the
one-thread walk's on clustered / structureless / degenerate data, for cosine, euclidean and Lightning's d_i-weighted variant."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc
from tests.test_kpp_sum_model import cases, sequential

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    from hnsw_clj_b200 import _lib

    _lib.check(_lib.lib().hb_init(0))
    yield _lib
    _lib.set_option("kpp_scale", 1)


def _sum_pick(hb, xs, u):
    w = np.ascontiguousarray(xs, dtype=np.float64)
    tot, pick = C.c_double(), C.c_int64()
    hb.check(hb.lib().hb_kpp_sum_pick(hb.ptr(w), w.size, float(u), C.byref(tot), C.byref(pick)))
    return tot.value, pick.value


def test_ordered_sum_and_pick_equal_the_sequential_loop(hb):
    r = np.random.default_rng(1)
    for name, xs in cases():
        run = sequential(xs)
        for u in [0.0, 0.5, 0.999999999, float(r.random()), float(r.random())]:
            tot, pick = _sum_pick(hb, xs, u)
            assert tot.hex() == run[-1].hex(), name
            target = u * run[-1]
            want = next((i for i, c in enumerate(run) if c >= target), len(xs) - 1)
            assert pick == want, (name, u)


def test_ordered_sum_large(hb):
    g = np.random.default_rng(2)
    xs = (g.random(3_000_000) * 0.4) ** 2
    run = np.add.accumulate(xs)  # numpy accumulates left to right in fp64
    for u in (0.123, 0.77):
        tot, pick = _sum_pick(hb, xs, u)
        assert tot.hex() == float(run[-1]).hex()
        assert pick == int(np.argmax(run >= u * run[-1]))


def _datasets():
    r = np.random.default_rng(5)
    c = r.standard_normal((40, 96))
    clustered = (c[r.integers(0, 40, 9000)] + 0.1 * r.standard_normal((9000, 96))).astype(np.float32)
    yield "clustered", clustered, 64
    yield "structureless", r.standard_normal((4000, 64)).astype(np.float32), 48
    dup = clustered[:3000].copy()
    dup[100:200] = dup[0]          # exact duplicates: d_i = 0 once one of them is a seed
    dup[500] = 0.0                 # a zero row: distance 1.0 to everything by the guard
    dup[501] = 0.0
    yield "duplicates and zero rows", dup, 40
    yield "more seeds than clusters", clustered[:1500], 300


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_seeds_equal_oracle_and_the_sequential_walk(hb, metric):
    from hnsw_clj_b200 import ivf_flat

    code = orc.COSINE if metric == "cosine" else orc.L2
    for name, rows, nlist in _datasets():
        want = orc.kmeanspp_init(rows, nlist, metric=code, seed=42)
        hb.set_option("kpp_scale", 1)
        hb.set_option("profile", 1)
        got = ivf_flat.kmeanspp_init(rows, nlist, distance_fn=metric)
        scored, steps = hb.get_stat("kpp_rows_scored"), hb.get_stat("kpp_steps")
        hb.set_option("profile", 0)
        assert got.tolist() == want.tolist(), (name, metric)
        hb.set_option("kpp_scale", 0)
        assert ivf_flat.kmeanspp_init(rows, nlist, distance_fn=metric).tolist() == want.tolist(), (name, metric, "walk")
        hb.set_option("kpp_scale", 1)
        assert steps == nlist - 1 and scored <= (nlist - 1) * len(rows)
        if name == "clustered":
            assert scored < 0.7 * (nlist - 1) * len(rows)  # the triangle inequality skips rows once their cluster holds a seed


def test_lightning_linear_weights(hb):
    """build-lightning-index seeds with d_i, not d_i^2 (lightning.clj:100-106)."""
    from hnsw_clj_b200 import lightning

    r = np.random.default_rng(9)
    c = r.standard_normal((30, 64))
    rows = (c[r.integers(0, 30, 5000)] + 0.15 * r.standard_normal((5000, 64))).astype(np.float32)
    want_c, want_a = orc.lightning_build(rows, 24)
    for scale in (1, 0):
        hb.set_option("kpp_scale", scale)
        with lightning.build_index(rows, num_partitions=24, smart_partition=True) as ix:
            cents, asg = ix.export()
        assert (asg == want_a).all() and (cents.view(np.int64) == want_c.view(np.int64)).all(), scale
    hb.set_option("kpp_scale", 1)

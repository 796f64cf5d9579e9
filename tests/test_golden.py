"""The committed golden vectors (tests/golden/golden.json, made by make_golden.py) against the oracle on CPU
and against the CUDA path on the GPU."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.golden import make_golden as mg

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def unhex(xs, shape=None):
    a = np.array([float.fromhex(x) for x in xs], dtype=np.float64)
    return a.reshape(shape) if shape else a


def test_oracle_reproduces_golden():
    assert mg.build() == GOLD


def test_reference_kats_in_golden():
    # test/hnsw/core_test.clj:16-31
    assert abs(float.fromhex(GOLD["pairwise"]["euclid_123_456"]) - 5.196152422706632) < 1e-5
    assert abs(float.fromhex(GOLD["pairwise"]["cos_123_456"]) - 0.0253) < 0.01


@pytest.mark.gpu
def test_cuda_flat_matches_golden():
    from hnsw_clj_b200.flat import FlatIndex

    c = mg.CASES["flat_unit"]
    rows, q = mg.dataset(c["kind"], c["n"], c["d"], c["seed"]), mg.dataset(c["kind"], c["nq"], c["d"], c["qseed"])
    for metric in ("cosine", "euclidean", "ip"):
        with FlatIndex(rows, distance_fn=metric) as ix:
            ids, dist = ix.search_raw(q, c["k"])
        g = GOLD[f"flat_unit/{metric}"]
        assert ids.tolist() == g["ids"]
        assert (dist.view(np.int64) == unhex(g["dist"], dist.shape).view(np.int64)).all()


@pytest.mark.gpu
def test_cuda_ivf_matches_golden():
    from hnsw_clj_b200 import ivf_flat

    c = mg.CASES["ivf_clustered"]
    rows = mg.dataset(c["kind"], c["n"], c["d"], c["seed"], c["clusters"])
    q = mg.dataset(c["kind"], c["nq"], c["d"], c["qseed"], c["clusters"])
    g = GOLD["ivf_clustered"]
    assert ivf_flat.kmeanspp_init(rows, c["nlist"], seed=42).tolist() == g["seed_rows"]
    with ivf_flat.build_index(rows, num_partitions=c["nlist"], max_iterations=c["iters"]) as ix:
        cents, asg = ix.export()
        ids, dist = ix.search_raw(q, c["k"], c["nprobe"])
        probes = ix.probes(q, c["nprobe"])
    assert asg.tolist() == g["assign"]
    assert [float(x).hex() for x in cents[0]] == g["centroid0"]
    assert ids.tolist() == g["ids"] and probes.tolist() == g["probes"]
    assert (dist.view(np.int64) == unhex(g["dist"], dist.shape).view(np.int64)).all()

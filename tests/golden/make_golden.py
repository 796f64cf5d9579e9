"""Generates tests/golden/*.json from the CPU oracle (python tests/golden/make_golden.py).

The reference (pure Clojure) cannot run in the build image (no JVM), and its own tests pin only three pairwise
KATs on this path (test/hnsw/core_test.clj:9-31).  These fixtures therefore pin the ORACLE's outputs for top-k
ids / IVF partitions (parity "unpinned by the reference": see oracle/hnsw_oracle.c header and DESIGN.md), so that
(a) any later edit of the oracle that changes results is caught on CPU, and (b) the CUDA path is compared with
committed numbers, not only with a same-run oracle.  Inputs are regenerated from the seeds below with the restated
java.util.Random generator of test/data_generator.clj:28-87; fp64 values are stored as hex for bit-exactness.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402


def hexes(a):
    return [float(x).hex() for x in np.asarray(a, np.float64).reshape(-1)]


def dataset(kind, n, d, seed, clusters=10):
    return orc.gen_dataset(n, d, kind, num_clusters=clusters, noise=0.1, seed=seed).astype(np.float32)


CASES = {
    "flat_unit": dict(kind=orc.UNIT, n=500, d=32, nq=12, k=10, seed=42, qseed=43),
    "ivf_clustered": dict(kind=orc.CLUSTERED, n=1200, d=24, nq=16, k=10, seed=42, qseed=43, clusters=12, nlist=16, iters=10,
                          nprobe=4),
}


def build():
    out = {}
    c = CASES["flat_unit"]
    rows, q = dataset(c["kind"], c["n"], c["d"], c["seed"]), dataset(c["kind"], c["nq"], c["d"], c["qseed"])
    for metric, code in (("cosine", orc.COSINE), ("euclidean", orc.L2), ("ip", orc.IP)):
        ids, dist = orc.exact_knn(rows, q, c["k"], code)
        out[f"flat_unit/{metric}"] = {"ids": ids.tolist(), "dist": hexes(dist)}
    out["flat_unit/row0"] = hexes(rows[0])  # guards the data generator itself
    c = CASES["ivf_clustered"]
    rows = dataset(c["kind"], c["n"], c["d"], c["seed"], c["clusters"])
    q = dataset(c["kind"], c["nq"], c["d"], c["qseed"], c["clusters"])
    seeds = orc.kmeanspp_init(rows, c["nlist"], seed=42)
    cents, asg = orc.kmeans(rows, c["nlist"], iters=c["iters"], seed=42)
    ids, dist, probes = orc.ivf_search(rows, cents, asg, q, c["k"], c["nprobe"], return_probes=True)
    ex, _ = orc.exact_knn(rows, q, c["k"])
    out["ivf_clustered"] = {"seed_rows": seeds.tolist(), "assign": asg.tolist(), "centroid0": hexes(cents[0]),
                            "centroid_sum": float(cents.sum()).hex(), "ids": ids.tolist(), "dist": hexes(dist),
                            "probes": probes.tolist(), "recall": orc.recall(ids, ex)}
    out["pairwise"] = {"cos_123_456": orc.cosine_distance([1, 2, 3], [4, 5, 6]).hex(),
                       "euclid_123_456": orc.euclidean_distance([1, 2, 3], [4, 5, 6]).hex()}
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(build(), f)
    print("wrote golden.json")

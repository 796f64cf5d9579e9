"""Mirror of hnsw.simd (src/hnsw/simd.clj): the float[] Vector-API variants of the distance functions, on the device.

`lanes` is FloatVector/SPECIES_PREFERRED .length of the machine the reference runs on (simd.clj:7-10: 4 with 128-bit
vectors, 8 with AVX2, 16 with AVX-512); the chunking is part of the arithmetic (fp32 lane products summed in fp32 per chunk,
chunk sums accumulated in fp64 — :18-43, :73-115).  The JDK does not specify the lane order of reduceLanes(ADD): results
agree with any JVM within 1e-5 relative (in practice ~1e-9) and with the CPU restatement used by the tests bit for bit.  double[] inputs take the
-direct / -doubles functions (:129-160, :176-185), i.e. the sequential fp64 kernels of simd_optimized."""
from __future__ import annotations

import numpy as np

from . import _lib as hb
from . import simd_optimized

SPECIES_LENGTH = 8  # AVX2; hnsw.simd prints its own at load time (simd.clj:12-14)


def _f32(x):
    a = np.ascontiguousarray(np.atleast_2d(np.asarray(x)), dtype=np.float32)
    if a.ndim != 2:
        raise hb.HbInvalid(hb.ERR_INVALID, "expected a vector or a [n, d] matrix")
    return a


def pairwise(a, b, metric=hb.COSINE, lanes: int = SPECIES_LENGTH) -> np.ndarray:
    """out[i, j] = metric(a_i, b_j) in the Vector-API arithmetic (hb_pairwise_f32lanes)."""
    A, B = _f32(a), _f32(b)
    if A.shape[1] != B.shape[1]:
        raise hb.HbInvalid(hb.ERR_INVALID, "vectors must have the same dimension")
    out = np.empty((A.shape[0], B.shape[0]), dtype=np.float64)
    hb.check(hb.lib().hb_pairwise_f32lanes(hb.ptr(A), A.shape[0], hb.ptr(B), B.shape[0], A.shape[1], metric, int(lanes), hb.ptr(out)))
    return out


def dot_product_simd_optimized(a, b, lanes: int = SPECIES_LENGTH) -> float:
    """simd.clj:18-43."""
    return float(pairwise(a, b, hb.IP, lanes)[0, 0])


def euclidean_distance_simd_optimized(a, b, lanes: int = SPECIES_LENGTH) -> float:
    """simd.clj:45-71."""
    return float(pairwise(a, b, hb.L2, lanes)[0, 0])


def cosine_distance_simd_optimized(a, b, lanes: int = SPECIES_LENGTH) -> float:
    """simd.clj:73-115."""
    return float(pairwise(a, b, hb.COSINE, lanes)[0, 0])


def batch_cosine_distances_simd(query, vectors, lanes: int = SPECIES_LENGTH) -> np.ndarray:
    """simd.clj:119-125: one query against a list of vectors (pmap above 8 vectors in the reference; one launch here)."""
    return pairwise(query, vectors, hb.COSINE, lanes)[0]


# (def cosine-distance cosine-distance-simd-optimized) etc., simd.clj:164-171
cosine_distance = cosine_distance_simd = cosine_distance_simd_optimized
euclidean_distance = euclidean_distance_simd = euclidean_distance_simd_optimized
dot_product = dot_product_simd = dot_product_simd_optimized

# double[] wrappers (simd.clj:129-160, :174-185): the sequential fp64 functions
cosine_distance_direct = cosine_distance_simd_doubles = simd_optimized.cosine_distance
euclidean_distance_direct = euclidean_distance_simd_doubles = simd_optimized.euclidean_distance

"""ctypes binding of libhnswb200.so — the same C ABI (include/hnswb200.h) the Clojure shim binds through
java.lang.foreign.  There is no CPU fallback: if the library is missing or no sm_100
device is visible every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhnswb200.so")

F32, BF16, F64 = 0, 1, 2
COSINE, L2, IP = 0, 1, 2
INDEX_FLAT, INDEX_IVF_FLAT, INDEX_HNSW = 0, 1, 2
MODE_EXACT, MODE_FAST = 0, 1

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_OOM, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5


class HbError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"libhnswb200 status {status}: {msg}")
        self.status = status


class HbInvalid(HbError, ValueError):
    """IllegalArgumentException of the reference (src/hnsw/api/simple.clj:13-14)."""


class HbInfo(C.Structure):
    _fields_ = [("type", C.c_int32), ("dtype", C.c_int32), ("metric", C.c_int32), ("dim", C.c_int32),
                ("n", C.c_int64), ("nlist", C.c_int32), ("max_level", C.c_int32), ("device_bytes", C.c_int64)]


_p, _i64, _i32, _int = C.c_void_p, C.c_int64, C.c_int32, C.c_int
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/hnswb200.h declares
SIGNATURES = {
    "hb_init": (_int, [_int]),
    "hb_shutdown": (_int, []),
    "hb_last_error": (C.c_char_p, []),
    "hb_version": (_int, []),
    "hb_set_stream": (_int, [_p]),
    "hb_set_mode": (_int, [_int]),
    "hb_set_option": (_int, [C.c_char_p, _i64]),
    "hb_get_stat": (_int, [C.c_char_p, C.POINTER(C.c_double)]),
    "hb_launch_count": (_i64, [_int]),
    "hb_row_norms": (_int, [_p, _i64, _i32, _int, _p]),
    "hb_pairwise": (_int, [_p, _i64, _int, _p, _i64, _int, _i32, _int, _p]),
    "hb_flat_create": (_int, [_p, _i64, _i32, _int, _int, _pp]),
    "hb_ivf_build": (_int, [_p, _i64, _i32, _int, _int, _i32, _i32, _i64, _pp]),
    "hb_lightning_build": (_int, [_p, _i64, _i32, _int, _int, _i32, _i64, _pp]),
    "hb_ivf_import": (_int, [_p, _i64, _i32, _int, _int, _p, _i32, _p, _pp]),
    "hb_ivf_export": (_int, [_p, _p, _p]),
    "hb_search": (_int, [_p, _p, _int, _i64, _i32, _i32, _p, _p]),
    "hb_ivf_probes": (_int, [_p, _p, _int, _i64, _i32, _p]),
    "hb_lsh_matrices": (_int, [_i32, _i32, _i32, _i64, _p]),
    "hb_kmeanspp_init": (_int, [_p, _i64, _i32, _int, _int, _i32, _i64, _p]),
    "hb_kmeans_assign": (_int, [_p, _i64, _i32, _int, _int, _p, _i32, _p]),
    "hb_kmeans_update": (_int, [_p, _i64, _i32, _int, _p, _i32, _p, _p, _p]),
    "hb_kmeans": (_int, [_p, _i64, _i32, _int, _int, _i32, _i32, _i64, _p, _p, _p]),
    "hb_hnsw_create": (_int, [_p, _i64, _i32, _int, _int, _p, _i32, _i32, _pp, _pp, _pp]),
    "hb_gather_score": (_int, [_p, _p, _int, _i64, _p, _p, _i64, _p]),
    "hb_fast_scores": (_int, [_p, _p, _int, _i64, _p, _p, _p]),
    "hb_topk_merge": (_int, [_p, _p, _i32, _i64, _i32, _p, _p]),
    "hb_comm_unique_id": (_int, [_p]),
    "hb_comm_init": (_int, [_p, _i32, _i32]),
    "hb_comm_info": (_int, [C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "hb_comm_shutdown": (_int, []),
    "hb_comm_broadcast": (_int, [_p, _i64, _i32]),
    "hb_comm_allreduce_f64": (_int, [_p, _i64, _i32]),
    "hb_index_set_id_base": (_int, [_p, _i64]),
    "hb_index_set_mode": (_int, [_p, _int]),
    "hb_index_set_coarse_sharded": (_int, [_p, _int]),
    "hb_sharded_search": (_int, [_p, _p, _int, _i64, _i32, _i32, _p, _p]),
    "hb_sharded_kmeans": (_int, [_p, _i64, _i32, _int, _int, _i32, _i32, _p, _i64, _p, _p]),
    "hb_sharded_ivf_build": (_int, [_p, _i64, _i32, _int, _int, _i32, _i32, _p, _i64, _pp]),
    "hb_pairwise_f32lanes": (_int, [_p, _i64, _p, _i64, _i32, _int, _i32, _p]),
    "hb_pcaf_matrix": (_int, [_i32, _i32, _i64, _p]),
    "hb_pcaf_project": (_int, [_p, _i32, _i32, _p, _i64, _i32, _p]),
    "hb_pcaf_search": (_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _p, _p]),
    "hb_kpp_sum_pick": (_int, [_p, _i64, C.c_double, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "hb_index_save": (_int, [_p, C.c_char_p]),
    "hb_index_load": (_int, [C.c_char_p, _pp]),
    "hb_index_info": (_int, [_p, C.POINTER(HbInfo)]),
    "hb_index_free": (_int, [_p]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HbError(ERR_NO_DEVICE, f"{LIB_PATH} is missing: build it with `python -m hnsw_clj_b200.build` "
                                         "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != OK:
        msg = lib().hb_last_error().decode("utf-8", "replace")
        raise (HbInvalid if status == ERR_INVALID else HbError)(status, msg)


# ---- buffers: numpy arrays (host) or torch tensors (host or CUDA) pass straight through ----------
def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def dtype_code(x) -> int:
    if _is_torch(x):
        import torch

        return {torch.float32: F32, torch.bfloat16: BF16, torch.float64: F64}[x.dtype]
    return {np.dtype(np.float32): F32, np.dtype(np.float64): F64}[x.dtype]


def as_matrix(x, allow=(F32, F64, BF16)):
    """Contiguous 2-D float buffer: torch tensors are kept (zero copy), everything else becomes numpy.
    float64 input stays float64 (the reference's double[]); other numpy dtypes become float32."""
    if _is_torch(x):
        x = x.contiguous()
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if dtype_code(x) not in allow:
            raise HbInvalid(ERR_INVALID, f"unsupported tensor dtype {x.dtype}")
        return x
    a = np.asarray(x)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64 if a.dtype.kind == "f" and a.dtype.itemsize > 4 else np.float32)
    if a.ndim == 1:
        a = a[None, :]
    if a.ndim != 2:
        raise HbInvalid(ERR_INVALID, "expected a [n, d] matrix")
    return np.ascontiguousarray(a)


def ptr(x) -> int | None:
    if x is None:
        return None
    if _is_torch(x):
        return x.data_ptr()
    return x.ctypes.data


def set_option(name: str, value: int) -> None:
    check(lib().hb_set_option(name.encode(), int(value)))


def get_stat(name: str) -> float:
    v = C.c_double()
    check(lib().hb_get_stat(name.encode(), C.byref(v)))
    return v.value


def set_mode(mode: int) -> None:
    """MODE_EXACT (default) or MODE_FAST (tensor-core candidate pass + fp64 re-score + proof; same results)."""
    check(lib().hb_set_mode(int(mode)))


def launch_count(reset: bool = False) -> int:
    return int(lib().hb_launch_count(1 if reset else 0))

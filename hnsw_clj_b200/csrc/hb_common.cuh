// hb_common.cuh — shared device/host helpers for libhnswb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#include "../../include/hnswb200.h"

namespace hb {

// ---- errors --------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &m) : std::runtime_error(m), status(s) {}
};

#define HB_CUDA(expr)                                                                               \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw hb::Error(_e == cudaErrorMemoryAllocation ? HB_ERR_OOM : HB_ERR_CUDA,             \
                            std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
    } while (0)

#define HB_REQUIRE(cond, msg)                                                                       \
    do {                                                                                            \
        if (!(cond)) throw hb::Error(HB_ERR_INVALID, std::string(msg));                             \
    } while (0)

// launch bookkeeping (bench.py reports `gpu_launches`)
extern int64_t g_launches;
extern cudaStream_t g_stream;
#define HB_LAUNCH_CHECK()                                                                           \
    do {                                                                                            \
        ++hb::g_launches;                                                                           \
        HB_CUDA(cudaGetLastError());                                                                \
    } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- arithmetic that must not be contracted or reassociated --------------------------------
// The reference accumulates acc = acc + a*b in fp64 with a separately rounded product
// (SURVEY Appendix A.1).  When both factors are fp32-representable the product is exact in fp64
// (24+24 <= 53 bits), so a fused multiply-add rounds identically: kExactProd selects DFMA.
template <bool kExactProd>
__device__ __forceinline__ double mac_seq(double a, double b, double acc) {
    if constexpr (kExactProd) return __fma_rn(a, b, acc);
    else return __dadd_rn(acc, __dmul_rn(a, b));
}
// euclidean-distance-ultra, src/hnsw/ultra_fast.clj:43-51: d = a-b; acc + d*d
__device__ __forceinline__ double l2_seq(double a, double b, double acc) {
    double t = __dsub_rn(a, b);
    return __dadd_rn(acc, __dmul_rn(t, t));
}

// ---- element loads ------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ double to_f64(T v);
template <>
__device__ __forceinline__ double to_f64<float>(float v) { return (double)v; }
template <>
__device__ __forceinline__ double to_f64<double>(double v) { return v; }
template <>
__device__ __forceinline__ double to_f64<__nv_bfloat16>(__nv_bfloat16 v) { return (double)__bfloat162float(v); }

template <typename T>
struct is_f32_repr { static constexpr bool value = true; };
template <>
struct is_f32_repr<double> { static constexpr bool value = false; };

// ---- sortable keys ------------------------------------------------------------------------
// Total order on fp64 distances matching Double/compare for the values that occur: ascending,
// NaN (canonicalised) after +inf.
__device__ __forceinline__ uint64_t dist_key(double d) {
    if (d != d) return 0xFFFFFFFFFFFFFFFFull;
    uint64_t b = (uint64_t)__double_as_longlong(d);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_dist(uint64_t k) {
    if (k == 0xFFFFFFFFFFFFFFFFull) return __longlong_as_double(0x7FF8000000000000ll);
    uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace hb

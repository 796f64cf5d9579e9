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
extern int g_num_sms;
#define HB_LAUNCH_CHECK()                                                                           \
    do {                                                                                            \
        ++hb::g_launches;                                                                           \
        HB_CUDA(cudaGetLastError());                                                                \
    } while (0)

// ---- arithmetic that must not be contracted or reassociated --------------------------------
// The reference accumulates acc = acc + a*b in fp64 with a separately rounded product
// (SURVEY Appendix A.1).  When both factors are fp32-representable the product is exact in fp64
// (24+24 <= 53 bits), so a fused multiply-add rounds identically: ARITH_FMA selects DFMA.
enum Arith { ARITH_FMA = 0, ARITH_MULADD = 1, ARITH_L2 = 2 };

template <int ARITH>
__device__ __forceinline__ double mac_seq(double a, double b, double acc) {
    if constexpr (ARITH == ARITH_FMA) return __fma_rn(a, b, acc);
    else if constexpr (ARITH == ARITH_MULADD) return __dadd_rn(acc, __dmul_rn(a, b));
    else {  // euclidean-distance-ultra, src/hnsw/ultra_fast.clj:43-51: t = a-b; acc + t*t
        double t = __dsub_rn(a, b);
        return __dadd_rn(acc, __dmul_rn(t, t));
    }
}

// epilogues turning the accumulated sum into the reference's distance value
enum Epi {
    EPI_COS = 0,        // 1 - dot/(qn*vn)                 (ivf_flat.clj:226, bench.clj:83)
    EPI_COS_GUARD = 1,  // same, 1.0 unless both norms > 0 (ultra_fast.clj:92-95)
    EPI_L2 = 2,         // sqrt(sum)                       (ultra_fast.clj:51)
    EPI_NEGDOT = 3,     // -dot (inner-product ranking, extension)
    EPI_DOT = 4         // dot
};
__device__ __forceinline__ double apply_epi(int epi, double acc, double qn, double vn) {
    switch (epi) {
        case EPI_COS: return __dsub_rn(1.0, __ddiv_rn(acc, __dmul_rn(qn, vn)));
        case EPI_COS_GUARD:
            return (qn > 0.0 && vn > 0.0) ? __dsub_rn(1.0, __ddiv_rn(acc, __dmul_rn(qn, vn))) : 1.0;
        case EPI_L2: return __dsqrt_rn(acc);
        case EPI_NEGDOT: return -acc;
        default: return acc;
    }
}

// ---- element loads ------------------------------------------------------------------------
__device__ __forceinline__ double to_f64(float v) { return (double)v; }
__device__ __forceinline__ double to_f64(double v) { return v; }
__device__ __forceinline__ double to_f64(__nv_bfloat16 v) { return (double)__bfloat162float(v); }

template <typename T>
struct is_f32_repr { static constexpr bool value = true; };
template <>
struct is_f32_repr<double> { static constexpr bool value = false; };

static inline size_t dtype_size(int dtype) { return dtype == HB_F32 ? 4 : dtype == HB_BF16 ? 2 : 8; }

// ---- sortable keys ------------------------------------------------------------------------
// Total order on fp64 distances matching Double/compare for the values that occur: ascending,
// NaN (canonicalised) after +inf.
__host__ __device__ __forceinline__ uint64_t dist_key(double d) {
    if (d != d) return 0xFFFFFFFFFFFFFFFEull;
    union { double f; uint64_t u; } c;
    c.f = d;
    uint64_t b = c.u;
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double key_dist(uint64_t k) {
    union { double f; uint64_t u; } c;
    if (k >= 0xFFFFFFFFFFFFFFFEull) {
        c.u = 0x7FF8000000000000ull;
        return c.f;
    }
    c.u = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return c.f;
}
constexpr uint64_t kKeyEmpty = 0xFFFFFFFFFFFFFFFFull;  // sorts after everything, incl. NaN

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- growable device buffer ---------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void *get(size_t bytes) {
        if (bytes > cap) {
            if (p) {
                cudaStreamSynchronize(g_stream);
                cudaFree(p);
                p = nullptr;
                cap = 0;
            }
            size_t want = bytes + bytes / 8 + 256;
            cudaError_t e = cudaMalloc(&p, want);
            if (e != cudaSuccess) {
                cudaGetLastError();
                want = bytes;
                e = cudaMalloc(&p, want);
            }
            if (e != cudaSuccess) {
                p = nullptr;
                cudaGetLastError();
                throw Error(HB_ERR_OOM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed");
            }
            cap = want;
        }
        return p;
    }
    template <typename T>
    T *as(size_t count) { return (T *)get(count * sizeof(T)); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    ~DevBuf() {}  // freed explicitly (hb_shutdown / index free): no CUDA calls at static destruction
};

bool is_device_ptr(const void *p);

// A read-only input that may live on the host (copied into `stage`) or on the device (used as is).
const void *stage_in(const void *p, size_t bytes, DevBuf &stage);
// An output: device pointer to write to; finish_out copies back if the caller's buffer is host memory.
struct OutStage {
    void *user = nullptr;
    void *dev = nullptr;
    size_t bytes = 0;
    bool host = false;
};
OutStage stage_out(void *user, size_t bytes, DevBuf &stage);
void finish_out(const OutStage &o);

}  // namespace hb

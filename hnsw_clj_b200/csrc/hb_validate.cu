// hb_validate.cu — range checks of caller-supplied index arrays before they drive device addressing.
// The reference raises on such input (NullPointerException for an unknown id, src/hnsw/ultra_fast.clj:189-192;
// IllegalArgumentException, src/hnsw/api/simple.clj:13-14); here a bad index must become HB_ERR_INVALID, never an
// out-of-bounds read (a sticky CUDA fault poisons the whole process).
#include "hb_validate.cuh"

namespace hb {
namespace {

template <typename T>
__global__ void check_range_kernel(const T *__restrict__ p, int64_t n, int64_t lo, int64_t hi, int32_t *__restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t v = (int64_t)p[i];
    if (v < lo || v >= hi) atomicOr(flag, 1);
}

// off[0] == 0, non-decreasing, off[count-1] == total (total < 0: any)
__global__ void check_offsets_kernel(const int64_t *__restrict__ off, int64_t count, int64_t total, int32_t *__restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    bool bad = false;
    if (i == 0) bad = off[0] != 0;
    else bad = off[i] < off[i - 1];
    if (i == count - 1 && total >= 0) bad = bad || off[i] != total;
    if (bad) atomicOr(flag, 1);
}

// every id inside a CSR: in [0, n) ; per node the neighbour count must not exceed max_deg (0: any)
__global__ void check_csr_kernel(const int64_t *__restrict__ off, const int32_t *__restrict__ ids, int64_t n, int64_t max_deg,
                                 int32_t *__restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t b = off[i], e = off[i + 1];
    if (max_deg > 0 && e - b > max_deg) atomicOr(flag, 1);
    for (int64_t t = b; t < e; ++t) {
        const int32_t v = ids[t];
        if (v < 0 || v >= n) {
            atomicOr(flag, 1);
            break;
        }
    }
}

inline int blocks_for(int64_t n) { return (int)ceil_div(n > 0 ? n : 1, 256); }

}  // namespace

void launch_check_range_i32(const int32_t *p, int64_t n, int64_t lo, int64_t hi, int32_t *flag) {
    if (n <= 0) return;
    check_range_kernel<int32_t><<<blocks_for(n), 256, 0, g_stream>>>(p, n, lo, hi, flag);
    HB_LAUNCH_CHECK();
}
void launch_check_range_i64(const int64_t *p, int64_t n, int64_t lo, int64_t hi, int32_t *flag) {
    if (n <= 0) return;
    check_range_kernel<int64_t><<<blocks_for(n), 256, 0, g_stream>>>(p, n, lo, hi, flag);
    HB_LAUNCH_CHECK();
}
void launch_check_offsets(const int64_t *off, int64_t count, int64_t total, int32_t *flag) {
    if (count <= 0) return;
    check_offsets_kernel<<<blocks_for(count), 256, 0, g_stream>>>(off, count, total, flag);
    HB_LAUNCH_CHECK();
}
void launch_check_csr(const int64_t *off, const int32_t *ids, int64_t n, int64_t max_deg, int32_t *flag) {
    if (n <= 0) return;
    check_csr_kernel<<<blocks_for(n), 256, 0, g_stream>>>(off, ids, n, max_deg, flag);
    HB_LAUNCH_CHECK();
}

}  // namespace hb

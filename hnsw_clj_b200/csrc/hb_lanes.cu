// hb_lanes.cu — the float[] Vector-API variants of the reference's distance functions (SURVEY §8 a4,
// src/hnsw/simd.clj:18-115) and the two phases of PCAF built on them (src/hnsw/ann/dimreduct/pcaf.clj:48-81,195-253).
//
// Arithmetic: per SPECIES-LENGTH chunk the reference multiplies fp32 lanes and reduces them with
// reduceLanes(VectorOperators/ADD), then adds the float chunk sum into a double accumulator
// ((+ sum (double (.reduceLanes ...))), simd.clj:33,56,96-98); the tail (len mod lanes) is scalar in double.  The JDK
// leaves the lane order of the ADD reduction unspecified; the kernels take the order of its scalar fallback and of
// HotSpot's ordered AddReductionVF: fp32 adds left to right from 0.0f.  __fmul_rn / __fadd_rn keep the compiler from
// contracting the pair into an FMA.  With the same `lanes` the results equal the oracle's restatement bit for bit; against a
// JVM with another lane order or width they agree within the north star's 1e-5 relative bound for fp32
// (tests/test_gpu_pcaf.py).
//
// One thread owns one (a, b) pair and walks the dimension in order: the chunk order IS the arithmetic.  The low-dimensional
// scan of PCAF is 100 floats per row; neither phase is bound by anything but HBM latency at these sizes.
#include "hb_lanes.cuh"

namespace hb {
namespace {

template <int METRIC>  // HB_COSINE / HB_L2 / HB_IP
__device__ __forceinline__ double lanes_pair(const float *__restrict__ a, const float *__restrict__ b, int d, int lanes) {
    const int ub = d - d % lanes;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    for (int i = 0; i < ub; i += lanes) {
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
        for (int l = 0; l < lanes; ++l) {
            const float x = a[i + l], y = __ldg(b + i + l);
            if (METRIC == HB_L2) {
                const float df = __fsub_rn(x, y);
                s0 = __fadd_rn(s0, __fmul_rn(df, df));
            } else {
                s0 = __fadd_rn(s0, __fmul_rn(x, y));
                if (METRIC == HB_COSINE) {
                    s1 = __fadd_rn(s1, __fmul_rn(x, x));
                    s2 = __fadd_rn(s2, __fmul_rn(y, y));
                }
            }
        }
        acc0 = __dadd_rn(acc0, (double)s0);
        if (METRIC == HB_COSINE) {
            acc1 = __dadd_rn(acc1, (double)s1);
            acc2 = __dadd_rn(acc2, (double)s2);
        }
    }
    for (int j = ub; j < d; ++j) {
        const double x = (double)a[j], y = (double)__ldg(b + j);
        if (METRIC == HB_L2) {
            const double df = __dsub_rn(x, y);
            acc0 = __dadd_rn(acc0, __dmul_rn(df, df));
        } else {
            acc0 = __dadd_rn(acc0, __dmul_rn(x, y));
            if (METRIC == HB_COSINE) {
                acc1 = __dadd_rn(acc1, __dmul_rn(x, x));
                acc2 = __dadd_rn(acc2, __dmul_rn(y, y));
            }
        }
    }
    if (METRIC == HB_L2) return __dsqrt_rn(acc0);
    if (METRIC == HB_IP) return acc0;
    const double mag = __dmul_rn(__dsqrt_rn(acc1), __dsqrt_rn(acc2));  // (zero? magnitude) -> 1.0, simd.clj:112-115
    return mag == 0.0 ? 1.0 : __dsub_rn(1.0, __ddiv_rn(acc0, mag));
}

// out[i * out_stride + j] = metric(a_i, b_j) for i in [i0, i0 + ni): CTA = 128 rows of b x up to 8 rows of a (staged in
// shared memory, read by broadcast)
constexpr int LQ = 8;
template <int METRIC, typename TOut>
__global__ void __launch_bounds__(128) lanes_pairwise_kernel(const float *__restrict__ a, int64_t na, const float *__restrict__ b,
                                                             int64_t nb, int d, int lanes, TOut *__restrict__ out,
                                                             int64_t out_stride) {
    extern __shared__ float s_a[];  // [LQ][d]
    const int64_t i0 = (int64_t)blockIdx.y * LQ;
    const int ni = (int)min((int64_t)LQ, na - i0);
    for (int t = threadIdx.x; t < ni * d; t += blockDim.x) s_a[t] = a[i0 * d + t];
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    const float *bj = b + j * d;
    for (int i = 0; i < ni; ++i) out[(i0 + i) * out_stride + j] = (TOut)lanes_pair<METRIC>(s_a + i * d, bj, d, lanes);
}

// out[q * c + s] = cosine(query q, row cand[q * c + s]) (+inf where the candidate slot is unused)
__global__ void __launch_bounds__(128) lanes_gather_kernel(const float *__restrict__ queries, const float *__restrict__ rows, int d,
                                                           int lanes, const int64_t *__restrict__ cand, int64_t nq, int c,
                                                           double *__restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nq * c) return;
    const int64_t r = cand[p];
    out[p] = r < 0 ? INFINITY : lanes_pair<HB_COSINE>(queries + (p / c) * d, rows + r * d, d, lanes);
}

}  // namespace

template <typename TOut>
static void pairwise_launch(const float *a, int64_t na, const float *b, int64_t nb, int d, int metric, int lanes, TOut *out,
                            int64_t out_stride) {
    if (na == 0 || nb == 0) return;
    HB_REQUIRE(lanes >= 1 && lanes <= 64, "lanes must be 1..64");
    HB_REQUIRE((size_t)LQ * d * 4 <= 96 * 1024, "dimension too large for the lanes kernels");
    const dim3 grid((unsigned)ceil_div(nb, 128), (unsigned)ceil_div(na, LQ));
    HB_REQUIRE(grid.y <= 65535, "too many `a` rows for one launch");
    const size_t smem = (size_t)LQ * d * 4;
#define HB_LP(M)                                                                                              \
    do {                                                                                                      \
        auto kern = lanes_pairwise_kernel<M, TOut>;                                                           \
        HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
        kern<<<grid, 128, smem, g_stream>>>(a, na, b, nb, d, lanes, out, out_stride);                         \
    } while (0)
    if (metric == HB_COSINE) HB_LP(HB_COSINE);
    else if (metric == HB_L2) HB_LP(HB_L2);
    else if (metric == HB_IP) HB_LP(HB_IP);
    else throw Error(HB_ERR_INVALID, "unknown metric");
#undef HB_LP
    HB_LAUNCH_CHECK();
}

void launch_lanes_pairwise(const float *a, int64_t na, const float *b, int64_t nb, int d, int metric, int lanes, double *out,
                           int64_t out_stride) {
    // rows of `a` in slices of 65535 * 8 (grid.y)
    const int64_t step = (int64_t)65535 * LQ;
    for (int64_t i0 = 0; i0 < na; i0 += step)
        pairwise_launch<double>(a + i0 * d, std::min(step, na - i0), b, nb, d, metric, lanes, out + i0 * out_stride, out_stride);
}
void launch_lanes_project(const float *matrix, int64_t target_dim, const float *rows, int64_t n, int d, int lanes, float *out) {
    // out[r][t] = (float) dot(matrix_t, row_r): `a` = the matrix rows (staged), `b` = the data rows; transposed store
    // through out_stride is not available, so rows play `a` here: out[r * target + t]
    const int64_t step = (int64_t)65535 * LQ;
    for (int64_t r0 = 0; r0 < n; r0 += step)
        pairwise_launch<float>(rows + r0 * d, std::min(step, n - r0), matrix, target_dim, d, HB_IP, lanes, out + r0 * target_dim,
                               target_dim);
}
void launch_lanes_gather(const float *queries, const float *rows, int d, int lanes, const int64_t *cand, int64_t nq, int c,
                         double *out) {
    if (nq * c == 0) return;
    HB_REQUIRE(lanes >= 1 && lanes <= 64, "lanes must be 1..64");
    lanes_gather_kernel<<<(unsigned)ceil_div(nq * c, 128), 128, 0, g_stream>>>(queries, rows, d, lanes, cand, nq, c, out);
    HB_LAUNCH_CHECK();
}

}  // namespace hb

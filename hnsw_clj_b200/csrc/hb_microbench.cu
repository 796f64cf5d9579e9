// hb_microbench.cu — measures the peak of the fp64 FMA pipe (the pipe that bounds the exact kernels), so the
// roofline fractions reported for them are "of measured", like the HBM / bf16 figures in MEASURED_PEAKS.json.
#include "hb_common.cuh"

namespace hb {
namespace {
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = __fma_rn(x0, a, b);
            x1 = __fma_rn(x1, a, b);
            x2 = __fma_rn(x2, a, b);
            x3 = __fma_rn(x3, a, b);
            x4 = __fma_rn(x4, a, b);
            x5 = __fma_rn(x5, a, b);
            x6 = __fma_rn(x6, a, b);
            x7 = __fma_rn(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678) out[0] = s;  // keep the chain alive
}
}  // namespace

double fp64_peak_tflops() {
    double *d;
    HB_CUDA(cudaMalloc(&d, 8));
    const int iters = 4096, grid = g_num_sms * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a, g_stream);
        dfma_peak_kernel<<<grid, 256, 0, g_stream>>>(d, iters, 0.999999, 1e-9);
        ++g_launches;
        cudaEventRecord(b, g_stream);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = 2.0 * 64.0 * iters * 256.0 * grid;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    HB_CUDA(cudaGetLastError());
    return best;
}
}  // namespace hb

// ---- tensor-pipe issue rate per operand kind (tcgen05.mma M = 128, N = 128, operands in shared memory) ----------------
// Measures how many MMA instructions per second one SM sustains for kind::i8 / f16 (bf16) / f8f6f4 (e4m3) / tf32, so
// the FAST-mode design can state which pipe it is bound by with a measured peak instead of a nominal one.
namespace hb {
namespace {
__device__ __forceinline__ uint32_t mb_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int KIND>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long *out_clocks) {
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const uint32_t base = (mb_smem_u32(smem_dyn) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 2 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_dyn + (base - mb_smem_u32(smem_dyn)))[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb_smem_u32(&s_bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mb_smem_u32(&s_tmem)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        const uint64_t ad = (uint64_t)((base & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        const uint64_t bd = (uint64_t)(((base + 16384) & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
                            ((uint64_t)2 << 61);
        // idesc: c_format @4 (F32 = 1, S32 = 2), a/b format @7/@10, N>>3 @17, M>>4 @24
        uint32_t idesc = ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (KIND == 0) idesc |= (2u << 4) | (1u << 7) | (1u << 10);       // i8: S32, signed int8
        else if (KIND == 1) idesc |= (1u << 4) | (1u << 7) | (1u << 10);  // f16 kind: F32 accum, BF16 operands
        else if (KIND == 2) idesc |= (1u << 4);                           // f8f6f4: F32 accum, E4M3 operands
        else idesc |= (1u << 4) | (2u << 7) | (2u << 10);                 // tf32
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (KIND == 0) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
            else if (KIND == 1) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
            else if (KIND == 2) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
            else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb_smem_u32(&s_bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(mb_smem_u32(&s_bar)), "r"(0) : "memory");
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) out_clocks[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128) : "memory");
}
}  // namespace

// clocks per M128 x N128 MMA instruction (K = 32 bytes of operand per row) for kind 0 = i8, 1 = bf16, 2 = e4m3, 3 = tf32
double mma_clocks_per_instr(int kind) {
    long long *d;
    HB_CUDA(cudaMalloc(&d, 8));
    const int iters = 4096, smem = 2 * 16384 + 1024;
    auto run = [&](auto kernel) {
        HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int rep = 0; rep < 2; ++rep) {
            kernel<<<g_num_sms, 128, smem, g_stream>>>(iters, d);
            ++g_launches;
        }
    };
    if (kind == 0) run(mma_rate_kernel<0>);
    else if (kind == 1) run(mma_rate_kernel<1>);
    else if (kind == 2) run(mma_rate_kernel<2>);
    else run(mma_rate_kernel<3>);
    long long h = 0;
    HB_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, g_stream));
    HB_CUDA(cudaStreamSynchronize(g_stream));
    cudaFree(d);
    HB_CUDA(cudaGetLastError());
    return (double)h / iters;
}
}  // namespace hb

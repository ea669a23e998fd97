// Measured ceiling of the tensor pipe this library runs on: tcgen05.mma kind::tf32, 128 x 256 x 8, issued back to back
// from shared memory by one thread per CTA, one CTA per SM (operands are whatever sits in shared memory; the pipe
// rate does not depend on the values).  bench.py calls it live, at the clock the bench itself runs at, and reports the
// roofline fraction against this number as well as against the bf16 figure of MEASURED_PEAKS.json.
// A measurement utility, not on the hot path: it synchronises.
#include "common.cuh"

namespace {
__device__ __forceinline__ uint32_t pr_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool pr_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t pr_desc(uint32_t saddr) {   // K-major, 128B swizzle, SBO = 1 KB
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(128) tf32_pipe_rate_kernel(int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 0.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pr_smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pr_smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        constexpr int N = 256;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint64_t ad = pr_desc(pr_smem_u32(smem)), bd = pr_desc(pr_smem_u32(smem) + 16 * 1024);
        if (pr_elect_one()) {
            for (int i = 0; i < iters; ++i) {
                const uint32_t acc = tmem + (uint32_t)((i & 1) * N);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(acc), "l"(ad + 2 * kk), "l"(bd + 2 * kk), "r"(idesc), "r"(1u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pr_smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(pr_smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}
}  // namespace

extern "C" int ramnet_tf32_pipe_rate(ramnet_handle *h, double *tflops_out, double *ms_out) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && tflops_out, "tf32_pipe_rate: bad argument");
    const int smem = 64 * 1024, iters = 20000;     // 80000 MMAs of 128x256x8 per SM: ~5 ms at 128 clk each
    RAMNET_CUDA(cudaFuncSetAttribute(tf32_pipe_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    RAMNET_CUDA(cudaEventCreate(&e0));
    RAMNET_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {            // first repetition is the warm-up
        RAMNET_CUDA(cudaEventRecord(e0, 0));
        tf32_pipe_rate_kernel<<<h->sm_count, 128, smem, 0>>>(iters);
        RAMNET_LAUNCH_CHECK(h);
        RAMNET_CUDA(cudaEventRecord(e1, 0));
        RAMNET_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        RAMNET_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = (double)h->sm_count * iters * 4.0 * 2.0 * 128.0 * 256.0 * 8.0;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return RAMNET_OK;
}

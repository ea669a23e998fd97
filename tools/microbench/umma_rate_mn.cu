// Microbenchmark: tcgen05.mma kind::tf32 rate with MN-major operands (the weight-gradient kernels' layout:
// SWIZZLE_128B_BASE32B, K-groups of 4 rows, SBO = 512 B) against K-major operands, as a function of N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate_mn umma_rate_mn.cu && ./umma_rate_mn
// mode bit0: A MN-major, bit1: B MN-major, bit2: B blocks are the same box shifted by one pixel (LBO = 128 B)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {       // K-major, SWIZZLE_128B
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo) {   // MN-major, SWIZZLE_128B_BASE32B
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}

__global__ void __launch_bounds__(128) rate_kernel(int N, int iters, int mode, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        if (mode & 1) idesc |= 1u << 15;
        if (mode & 2) idesc |= 1u << 16;
        const uint32_t a_base = smem_u32(smem), b_base = a_base + 64 * 1024;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint64_t ad = (mode & 1) ? desc_mn(a_base, 4096) : desc_k(a_base);
                const uint64_t bd = (mode & 2) ? desc_mn(b_base, (mode & 4) ? 128 : 4096) : desc_k(b_base);
                const uint32_t a_step = (mode & 1) ? 64 : 2, b_step = (mode & 2) ? 64 : 2;   // per K = 8 step, in 16-byte units
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_tf32(tmem, ad + a_step * kk, bd + b_step * kk, idesc, 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            t1 = clock64();
        }
        __syncwarp();
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        const long long t2 = clock64();
        if (elect_one() && blockIdx.x == 0) { out[0] = t1 - t0; }
        if (threadIdx.x == 32 && blockIdx.x == 0) out[1] = t2;
        if (threadIdx.x == 32 && blockIdx.x == 0) out[2] = t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

int main() {
    long long *out;
    cudaMallocManaged(&out, 64);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    printf("%6s %5s %14s %10s   (mode bit0: A MN-major, bit1: B MN-major, bit2: B blocks shifted by 128 B)\n", "N", "mode",
           "cyc/MMA", "N/2");
    const int modes[] = {0, 1, 2, 3, 7};
    const int Ns[] = {32, 64, 96, 128, 160, 256};
    for (int mode : modes)
        for (int N : Ns) {
            rate_kernel<<<148, 128, smem>>>(N, iters, mode, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s (N=%d mode=%d)\n", cudaGetErrorString(e), N, mode); return 1; }
            printf("%6d %5d %14.1f %10d\n", N, mode, (double)(out[1] - out[2]) / (4.0 * iters), N / 2);
        }
    return 0;
}

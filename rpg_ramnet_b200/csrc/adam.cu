// Fused Adam over one flat fp32 buffer.
//
// Replaces torch.optim.Adam.step as built by base/base_trainer.py:36-37 and called at
// trainer/lstm_trainer.py:453 (one foreach/for-loop chain per tensor) with a single
// launch over all parameters.  HBM-bound: reads p, g, m, v and writes p, m, v = 28 B / param;
// 128-bit vector accesses, grid = SMs x 8 resident CTAs, grid-stride.
// Update order follows torch.optim.Adam (non-amsgrad, non-decoupled weight decay):
//   g += wd*p ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include "common.cuh"

struct AdamK {
    float b1, b2, omb1, omb2, eps, wd, step_size, inv_sqrt_bc2;  // omb = 1 - beta, rounded once from double
};

__device__ __forceinline__ void adam1(float &p, float g, float &m, float &v, const AdamK &k) {
    const float b1 = k.b1, b2 = k.b2, eps = k.eps, wd = k.wd, step_size = k.step_size, inv_sqrt_bc2 = k.inv_sqrt_bc2;
    if (wd != 0.f) g = fmaf(wd, p, g);
    m = b1 * m + k.omb1 * g;
    v = b2 * v + k.omb2 * g * g;
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ p, const float *__restrict__ g,
                                                   float *__restrict__ m, float *__restrict__ v, int64_t n, AdamK k) {
    const int64_t n4 = n >> 2;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n4; i += nthr) {
        float4 pp = reinterpret_cast<float4 *>(p)[i], mm = reinterpret_cast<float4 *>(m)[i];
        float4 vv = reinterpret_cast<float4 *>(v)[i];
        const float4 gg = reinterpret_cast<const float4 *>(g)[i];
        adam1(pp.x, gg.x, mm.x, vv.x, k);
        adam1(pp.y, gg.y, mm.y, vv.y, k);
        adam1(pp.z, gg.z, mm.z, vv.z, k);
        adam1(pp.w, gg.w, mm.w, vv.w, k);
        reinterpret_cast<float4 *>(p)[i] = pp;
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nthr) adam1(p[i], g[i], m[i], v[i], k);
}

// CUDA-graph-capturable variant: the step number lives in device memory (incremented by the kernel launch itself),
// so a captured training step can be replayed without baking the bias corrections into the graph.
__global__ void adam_bump_step_kernel(int *step) { *step += 1; }

__global__ void __launch_bounds__(256) adam_dev_step_kernel(float *__restrict__ p, const float *__restrict__ g,
                                                            float *__restrict__ m, float *__restrict__ v, int64_t n,
                                                            double lr, double beta1, double beta2, float eps, float wd,
                                                            const int *__restrict__ step) {
    const int t = *step;
    AdamK k;
    k.b1 = (float)beta1; k.b2 = (float)beta2; k.omb1 = (float)(1.0 - beta1); k.omb2 = (float)(1.0 - beta2);
    k.eps = eps; k.wd = wd;
    k.step_size = (float)(lr / (1.0 - pow(beta1, (double)t)));
    k.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow(beta2, (double)t)));
    const int64_t n4 = n >> 2;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n4; i += nthr) {
        float4 pp = reinterpret_cast<float4 *>(p)[i], mm = reinterpret_cast<float4 *>(m)[i];
        float4 vv = reinterpret_cast<float4 *>(v)[i];
        const float4 gg = reinterpret_cast<const float4 *>(g)[i];
        adam1(pp.x, gg.x, mm.x, vv.x, k);
        adam1(pp.y, gg.y, mm.y, vv.y, k);
        adam1(pp.z, gg.z, mm.z, vv.z, k);
        adam1(pp.w, gg.w, mm.w, vv.w, k);
        reinterpret_cast<float4 *>(p)[i] = pp;
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nthr) adam1(p[i], g[i], m[i], v[i], k);
}

extern "C" int ramnet_adam_step_dev(ramnet_handle *h, float *p, const float *g, float *m, float *v, int64_t n, double lr,
                                    double beta1, double beta2, double eps, double weight_decay, int *step_counter,
                                    int increment, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && p && g && m && v && step_counter && n > 0, "adam_step_dev: bad argument");
    RAMNET_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0,
                     "adam_step_dev: buffers must be 16-byte aligned");
    if (increment) {   // the first slice of a bucketed step increments; the other slices of the same step only read
        adam_bump_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_counter);
        RAMNET_LAUNCH_CHECK(h);
    }
    const int blocks = (int)imin64(((n >> 2) + 255) / 256 + 1, (int64_t)h->sm_count * 8);
    adam_dev_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, (float)eps,
                                                                   (float)weight_decay, step_counter);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_adam_step(ramnet_handle *h, float *p, const float *g, float *m, float *v, int64_t n, double lr,
                                double beta1, double beta2, double eps, double weight_decay, int step, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && p && g && m && v, "adam_step: NULL argument");
    RAMNET_CHECK_ARG(n > 0 && step >= 1, "adam_step: n=%lld step=%d", (long long)n, step);
    RAMNET_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0,
                     "adam_step: buffers must be 16-byte aligned");
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    AdamK k;
    k.b1 = (float)beta1; k.b2 = (float)beta2; k.omb1 = (float)(1.0 - beta1); k.omb2 = (float)(1.0 - beta2);
    k.eps = (float)eps; k.wd = (float)weight_decay;
    k.step_size = (float)(lr / bc1);
    k.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const int blocks = (int)imin64(((n >> 2) + 255) / 256 + 1, (int64_t)h->sm_count * 8);
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, k);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

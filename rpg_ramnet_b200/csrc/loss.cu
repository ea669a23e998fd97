// scale_invariant_loss forward statistics, value and gradient.
//
// Replaces RAM_Net/model/loss.py:6-9: `log_diff[~is_nan]` boolean-mask gathers (dynamic shape,
// host sync) + two means.  Here: one streaming reduction (sum d, sum d^2, count of non-NaN d) in
// float64, and one streaming pass writing d loss / d pred.  Both are HBM-bound at 8 B / pixel
// (stats) and 12 B / pixel (grad); n and mean(d) are read from device memory so the host never
// synchronises, and the three statistics can be all-reduced between the two calls for the
// exact global-batch loss under data parallelism.
#include "common.cuh"

__global__ void __launch_bounds__(256) si_stats_kernel(const float *__restrict__ pred, const float *__restrict__ target,
                                                       int64_t n, double *__restrict__ stats, int log_space) {
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        // loss.py:7 (difference of the network's normalised log depths) or loss.py:13 (scale_invariant_log_loss: the
        // logarithm is taken here; log of a non-positive value is NaN / -inf exactly as torch.log)
        const float d = log_space ? logf(pred[i]) - logf(target[i]) : pred[i] - target[i];
        if (d == d) {  // ~isnan(log_diff), loss.py:8
            s1 += (double)d;
            s2 += (double)d * (double)d;
            cnt += 1.0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ double sh[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = s1; sh[1][warp] = s2; sh[2][warp] = cnt; }
    __syncthreads();
    if (warp == 0) {
        s1 = lane < 8 ? sh[0][lane] : 0.0;
        s2 = lane < 8 ? sh[1][lane] : 0.0;
        cnt = lane < 8 ? sh[2][lane] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (lane == 0) {
            atomicAdd(stats + 0, s1);
            atomicAdd(stats + 1, s2);
            atomicAdd(stats + 2, cnt);
        }
    }
}

__global__ void si_value_kernel(const double *__restrict__ stats, float weight, float n_lambda, float *out) {
    const double n = stats[2];
    const double mean = stats[0] / n, mean2 = stats[1] / n;  // n == 0 -> NaN, as torch's mean of an empty tensor
    *out = (float)((double)weight * (mean2 - (double)n_lambda * mean * mean));
}

__global__ void __launch_bounds__(256) si_grad_kernel(const float *__restrict__ pred, const float *__restrict__ target,
                                                      int64_t n, const double *__restrict__ stats, float weight,
                                                      float n_lambda, float scale, const float *__restrict__ scale_dev,
                                                      int log_space, float *__restrict__ grad) {
    const double cnt = stats[2];
    const float mean = (float)(stats[0] / cnt);
    // scale_dev: the upstream gradient d total / d loss_term as a device scalar (autograd's grad_output), so that no
    // separate elementwise multiply over the map and no host read-back is needed
    const double up = (double)scale * (scale_dev ? (double)*scale_dev : 1.0);
    const float k = (float)(2.0 * (double)weight * up / cnt);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float p = pred[i];
        const float d = log_space ? logf(p) - logf(target[i]) : p - target[i];
        const float g = k * (d - n_lambda * mean);
        grad[i] = (d == d) ? (log_space ? g / p : g) : 0.f;
    }
}

extern "C" int ramnet_si_loss_stats(ramnet_handle *h, const float *pred, const float *target, int64_t n, double *stats,
                                    int flags, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && pred && target && stats && n > 0, "si_loss_stats: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    RAMNET_CUDA(cudaMemsetAsync(stats, 0, 3 * sizeof(double), s));
    const int blocks = (int)imin64((n + 255) / 256, (int64_t)h->sm_count * 8);
    si_stats_kernel<<<blocks, 256, 0, s>>>(pred, target, n, stats, (flags & RAMNET_LOSS_LOG_SPACE) ? 1 : 0);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_si_loss_value(ramnet_handle *h, const double *stats, float weight, float n_lambda, float *loss_out,
                                    void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && stats && loss_out, "si_loss_value: bad argument");
    si_value_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(stats, weight, n_lambda, loss_out);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_si_loss_grad(ramnet_handle *h, const float *pred, const float *target, int64_t n,
                                   const double *stats, float weight, float n_lambda, float scale,
                                   const float *scale_dev, int flags, float *grad, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && pred && target && stats && grad && n > 0, "si_loss_grad: bad argument");
    const int blocks = (int)imin64((n + 255) / 256, (int64_t)h->sm_count * 8);
    si_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, target, n, stats, weight, n_lambda, scale, scale_dev,
                                                             (flags & RAMNET_LOSS_LOG_SPACE) ? 1 : 0, grad);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

// ---------------------------------------------------------------------------------------------
// MultiScaleGradient loss (RAM_Net/model/loss.py:22-70):
//   diff = pred - target; for scale s in 0..S-1: p_s = AvgPool2d(2^s)(diff); g = kornia.spatial_gradient(p_s)
//   (normalised 3x3 Sobel /8, replicate padding, two components); loss += sum|g[~nan]| / count(~nan) * B * 2; loss /= S.
// NaN semantics as the reference's conv / pooling: a pooled value is NaN if any pixel of its window is, a gradient
// pixel (both components) is NaN if any of its 3x3 pooled neighbours is (0 * NaN = NaN).
// kornia 0.4.0 is a third-party dependency that is not vendored (requirements.txt:32): the Sobel semantics are restated
// from its published behaviour (loss.py:54's comment gives the output shape) — parity for this loss is pinned on the
// oracle restatement only.
// One thread per pooled pixel; everything is recomputed from pred/target (the whole loss touches 8 B per pixel per scale).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float msg_pooled(const float *__restrict__ pred, const float *__restrict__ target, int n, int H,
                                            int W, int k, int py, int px) {
    float acc = 0.f;
    const float *p = pred + ((int64_t)n * H + (int64_t)py * k) * W + (int64_t)px * k;
    const float *t = target + ((int64_t)n * H + (int64_t)py * k) * W + (int64_t)px * k;
    for (int y = 0; y < k; ++y)
        for (int x = 0; x < k; ++x) acc += p[(int64_t)y * W + x] - t[(int64_t)y * W + x];   // NaN propagates
    return acc / (float)(k * k);
}

// gx, gy of pooled pixel (qy, qx); returns false when NaN
__device__ __forceinline__ bool msg_sobel(const float *pred, const float *target, int n, int H, int W, int k, int hp, int wp,
                                          int qy, int qx, float &gx, float &gy) {
    float v[3][3];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx)
            v[dy + 1][dx + 1] = msg_pooled(pred, target, n, H, W, k, min(max(qy + dy, 0), hp - 1), min(max(qx + dx, 0), wp - 1));
    gx = ((v[0][2] - v[0][0]) + 2.f * (v[1][2] - v[1][0]) + (v[2][2] - v[2][0])) * 0.125f + 0.f * v[1][1] + 0.f * v[0][1] + 0.f * v[2][1];
    gy = ((v[2][0] - v[0][0]) + 2.f * (v[2][1] - v[0][1]) + (v[2][2] - v[0][2])) * 0.125f + 0.f * v[1][1] + 0.f * v[1][0] + 0.f * v[1][2];
    return gx == gx && gy == gy;
}

__global__ void __launch_bounds__(256) msg_stats_kernel(const float *__restrict__ pred, const float *__restrict__ target,
                                                        int N, int H, int W, int k, double *__restrict__ stats) {
    const int hp = H / k, wp = W / k;
    const int64_t total = (int64_t)N * hp * wp;
    double s = 0.0, cnt = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int qx = (int)(i % wp), qy = (int)((i / wp) % hp), n = (int)(i / ((int64_t)wp * hp));
        float gx, gy;
        if (msg_sobel(pred, target, n, H, W, k, hp, wp, qy, qx, gx, gy)) {
            s += (double)fabsf(gx) + (double)fabsf(gy);
            cnt += 2.0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats + 0, s);
        atomicAdd(stats + 1, cnt);
    }
}

__global__ void msg_value_kernel(const double *__restrict__ stats, int N, int scales, float *out) {
    double loss = 0.0;
    for (int s = 0; s < scales; ++s) loss += stats[2 * s] / stats[2 * s + 1] * (double)N * 2.0;
    *out = (float)(loss / (double)scales);
}

__global__ void __launch_bounds__(256) msg_grad_kernel(const float *__restrict__ pred, const float *__restrict__ target,
                                                       int N, int H, int W, int k, const double *__restrict__ stats,
                                                       int scales, int n_batch, float gscale,
                                                       const float *__restrict__ scale_dev, float *__restrict__ grad) {
    const int hp = H / k, wp = W / k;
    const int64_t total = (int64_t)N * hp * wp;
    // d loss / d g = sign(g) * B * 2 / (count * scales); Sobel /8; avg-pool adjoint 1/k^2
    // (B = n_batch: the GLOBAL batch size when the statistics were all-reduced over data-parallel ranks)
    const float up = gscale * (scale_dev ? *scale_dev : 1.f);
    const float coef = up * (float)((double)n_batch * 2.0 / (stats[1] * (double)scales)) * 0.125f / (float)(k * k);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int qx = (int)(i % wp), qy = (int)((i / wp) % hp), n = (int)(i / ((int64_t)wp * hp));
        float gx, gy;
        if (!msg_sobel(pred, target, n, H, W, k, hp, wp, qy, qx, gx, gy)) continue;
        const float sx = (gx > 0.f) - (gx < 0.f), sy = (gy > 0.f) - (gy < 0.f);
        const float kx[3][3] = {{-1, 0, 1}, {-2, 0, 2}, {-1, 0, 1}};
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const float c = coef * (sx * kx[dy + 1][dx + 1] + sy * kx[dx + 1][dy + 1]);
                if (c == 0.f) continue;
                const int py = min(max(qy + dy, 0), hp - 1), px = min(max(qx + dx, 0), wp - 1);   // replicate-pad adjoint
                float *gp = grad + ((int64_t)n * H + (int64_t)py * k) * W + (int64_t)px * k;
                for (int y = 0; y < k; ++y)
                    for (int x = 0; x < k; ++x) atomicAdd(gp + (int64_t)y * W + x, c);
            }
    }
}

extern "C" int ramnet_msg_loss_stats(ramnet_handle *h, const float *pred, const float *target, int N, int H, int W,
                                     int start_scale, int scales, double *stats, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && pred && target && stats && N > 0 && H > 0 && W > 0 && scales > 0 && scales <= 8 && start_scale >= 1,
                     "msg_loss_stats: bad argument");
    RAMNET_CHECK_ARG(H % (start_scale << (scales - 1)) == 0 && W % (start_scale << (scales - 1)) == 0,
                     "msg_loss_stats: H, W must be divisible by start_scale * 2^(scales-1)");
    cudaStream_t s = (cudaStream_t)stream;
    RAMNET_CUDA(cudaMemsetAsync(stats, 0, 2 * scales * sizeof(double), s));
    for (int sc = 0; sc < scales; ++sc) {
        const int k = start_scale << sc;
        const int64_t total = (int64_t)N * (H / k) * (W / k);
        msg_stats_kernel<<<(int)imin64((total + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, s>>>(pred, target, N, H, W, k,
                                                                                                    stats + 2 * sc);
        RAMNET_LAUNCH_CHECK(h);
    }
    return RAMNET_OK;
}

extern "C" int ramnet_msg_loss_value(ramnet_handle *h, const double *stats, int N, int scales, float *loss_out, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && stats && loss_out && scales > 0, "msg_loss_value: bad argument");
    msg_value_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(stats, N, scales, loss_out);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_msg_loss_grad(ramnet_handle *h, const float *pred, const float *target, int N, int H, int W,
                                    int start_scale, int scales, const double *stats, int n_batch, float scale,
                                    const float *scale_dev, float *grad, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && pred && target && stats && grad && N > 0 && scales > 0 && scales <= 8 && start_scale >= 1,
                     "msg_loss_grad: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    RAMNET_CUDA(cudaMemsetAsync(grad, 0, (size_t)N * H * W * sizeof(float), s));
    for (int sc = 0; sc < scales; ++sc) {
        const int k = start_scale << sc;
        const int64_t total = (int64_t)N * (H / k) * (W / k);
        msg_grad_kernel<<<(int)imin64((total + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, s>>>(
            pred, target, N, H, W, k, stats + 2 * sc, scales, n_batch > 0 ? n_batch : N, scale, scale_dev, grad);
        RAMNET_LAUNCH_CHECK(h);
    }
    return RAMNET_OK;
}

// preview=True branch of MultiScaleGradient.forward (loss.py:46-47): kornia `sobel` = gradient magnitude
// sqrt(gx^2 + gy^2 + eps), eps = 1e-6, of the pooled difference at one scale.  Logging only (TensorBoard previews,
// lstm_trainer.py:162-165); the bicubic resize that follows stays a host-side torch call.
__global__ void __launch_bounds__(256) msg_sobel_mag_kernel(const float *__restrict__ pred, const float *__restrict__ target,
                                                            int N, int H, int W, int k, float *__restrict__ out) {
    const int hp = H / k, wp = W / k;
    const int64_t total = (int64_t)N * hp * wp;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int qx = (int)(i % wp), qy = (int)((i / wp) % hp), n = (int)(i / ((int64_t)wp * hp));
        float gx, gy;
        msg_sobel(pred, target, n, H, W, k, hp, wp, qy, qx, gx, gy);
        out[i] = sqrtf(gx * gx + gy * gy + 1e-6f);   // NaN where any pooled neighbour is NaN, as the reference
    }
}

extern "C" int ramnet_msg_sobel_preview(ramnet_handle *h, const float *pred, const float *target, int N, int H, int W,
                                        int pool, float *out, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && pred && target && out && N > 0 && H > 0 && W > 0 && pool >= 1 && H % pool == 0 && W % pool == 0,
                     "msg_sobel_preview: bad argument");
    const int64_t total = (int64_t)N * (H / pool) * (W / pool);
    msg_sobel_mag_kernel<<<(int)imin64((total + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, (cudaStream_t)stream>>>(
        pred, target, N, H, W, pool, out);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

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
                                                       int64_t n, double *__restrict__ stats) {
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float d = pred[i] - target[i];
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
                                                      float n_lambda, float scale, float *__restrict__ grad) {
    const double cnt = stats[2];
    const float mean = (float)(stats[0] / cnt);
    const float k = (float)(2.0 * (double)weight * (double)scale / cnt);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float d = pred[i] - target[i];
        grad[i] = (d == d) ? k * (d - n_lambda * mean) : 0.f;
    }
}

extern "C" int ramnet_si_loss_stats(ramnet_handle *h, const float *pred, const float *target, int64_t n, double *stats,
                                    void *stream) {
    RAMNET_CHECK_ARG(h && pred && target && stats && n > 0, "si_loss_stats: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    RAMNET_CUDA(cudaMemsetAsync(stats, 0, 3 * sizeof(double), s));
    const int blocks = (int)imin64((n + 255) / 256, (int64_t)h->sm_count * 8);
    si_stats_kernel<<<blocks, 256, 0, s>>>(pred, target, n, stats);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_si_loss_value(ramnet_handle *h, const double *stats, float weight, float n_lambda, float *loss_out,
                                    void *stream) {
    RAMNET_CHECK_ARG(h && stats && loss_out, "si_loss_value: bad argument");
    si_value_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(stats, weight, n_lambda, loss_out);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_si_loss_grad(ramnet_handle *h, const float *pred, const float *target, int64_t n,
                                   const double *stats, float weight, float n_lambda, float scale, float *grad,
                                   void *stream) {
    RAMNET_CHECK_ARG(h && pred && target && stats && grad && n > 0, "si_loss_grad: bad argument");
    const int blocks = (int)imin64((n + 255) / 256, (int64_t)h->sm_count * 8);
    si_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, target, n, stats, weight, n_lambda, scale, grad);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

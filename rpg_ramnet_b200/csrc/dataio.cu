// Device side of the loader -> network wire format and of the trainer's per-step metrics
// (SURVEY.md §8f ranks 2 and 3).  All of it is streaming HBM work: one or two passes over a
// [bins, H, W] voxel grid or a [N, 1, H, W] depth map, statistics reduced in float64 on the
// device so the host never has to pull a full tensor back.
//
//   ramnet_voxel_normalize   RAM_Net/data_loader/event_dataset.py:144-151 (twins:
//                            dataset_asynchronous.py:300-308, utils/event_tensor_utils.py:52-66):
//                            mean / stddev of the NON-ZERO voxels -> (x - mean) / stddev on them
//   ramnet_depth_to_label    RAM_Net/data_loader/dataset.py:296-305: metric depth -> normalised
//                            log depth in [0, 1], NaN (no ground truth) preserved
//   ramnet_depth_metrics     RAM_Net/model/metric.py:8-57 as called by
//                            trainer/lstm_trainer.py:100-106,291-294: masked error sums per sample
#include "common.cuh"

namespace {

__device__ __forceinline__ void block_reduce_add(double *vals, int nvals, double *out) {
    // vals: per-thread partial sums (registers, nvals <= 8); atomically adds the block totals to out[0..nvals)
    __shared__ double sh[8][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < nvals; ++k) {
        double v = vals[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[k][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
        for (int k = 0; k < nvals; ++k) {
            double v = lane < 8 ? sh[k][lane] : 0.0;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(out + k, v);
        }
    }
}

// stats[0] = sum, stats[1] = sum of squares, stats[2] = count over the non-zero voxels
// blockIdx.y = sample of a batch of grids (n voxels each, statistics per sample)
__global__ void __launch_bounds__(256) voxel_stats_kernel(const float *__restrict__ grid, int64_t n, double *__restrict__ stats) {
    grid += (int64_t)blockIdx.y * n;
    stats += blockIdx.y * 3;
    double v[3] = {0.0, 0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = grid[i];
        if (x != 0.f) {          // np.nonzero: NaN counts as non-zero, exactly as the reference
            v[0] += (double)x;
            v[1] += (double)x * (double)x;
            v[2] += 1.0;
        }
    }
    block_reduce_add(v, 3, stats);
}

__global__ void __launch_bounds__(256) voxel_normalize_kernel(float *__restrict__ grid, int64_t n, const double *__restrict__ stats) {
    grid += (int64_t)blockIdx.y * n;
    stats += blockIdx.y * 3;
    const double cnt = stats[2];
    if (!(cnt > 0.0)) return;                               // event_dataset.py:147 `if mask[0].size > 0`
    const double mean = stats[0] / cnt;
    double var = stats[1] / cnt - mean * mean;              // population variance (np.std, ddof = 0)
    if (var < 0.0) var = 0.0;
    const double sd = sqrt(var);
    if (!(sd > 0.0)) return;                                // :149 `if stddev > 0`
    const float m = (float)mean, inv = (float)(1.0 / sd);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = grid[i];
        if (x != 0.f) grid[i] = (x - m) * inv;
    }
}

// label = clip(1 + log(clip(d, 0, clip) / clip) / reg, 0, 1); NaN stays NaN (np.clip / np.log propagate it)
__global__ void __launch_bounds__(256) depth_label_kernel(const float *__restrict__ depth, float *__restrict__ label, int64_t n,
                                                          float clip_distance, float reg_factor) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float d = depth[i];
        float y = d;
        if (d == d) {
            const float c = fminf(fmaxf(d, 0.f), clip_distance) / clip_distance;
            y = 1.0f + logf(c) / reg_factor;                // log(0) = -inf -> clipped to 0 below
            y = fminf(fmaxf(y, 0.f), 1.0f);
        }
        label[i] = y;
    }
}

// Per sample s (blockIdx.y), over its HW pixels, out[s*8 + k]:
//   0: count of non-NaN (target - pred)      1: sum |d| / (target + eps)      2: sum d^2 / (target^2 + eps)
//   3: sum d^2                               4: sum |d|
//   5: count of non-NaN target               6: sum (pred - target)^2 over non-NaN target        7: unused
__global__ void __launch_bounds__(256) depth_metrics_kernel(const float *__restrict__ pred, const float *__restrict__ target,
                                                            int64_t hw, float eps, double *__restrict__ out) {
    const float *p = pred + (int64_t)blockIdx.y * hw, *t = target + (int64_t)blockIdx.y * hw;
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
        const float ti = t[i], pi = p[i];
        const float d = fabsf(ti - pi);
        if (d == d) {
            v[0] += 1.0;
            v[1] += (double)(d / (ti + eps));
            v[2] += (double)(d * d / (ti * ti + eps));
            v[3] += (double)d * (double)d;
            v[4] += (double)d;
        }
        if (ti == ti) {
            const double e = (double)pi - (double)ti;
            v[5] += 1.0;
            v[6] += e * e;
        }
    }
    block_reduce_add(v, 7, out + (int64_t)blockIdx.y * 8);
}
}  // namespace

extern "C" int ramnet_voxel_normalize(ramnet_handle *h, float *grid, int64_t n, int batch, double *stats, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && grid && stats && n > 0 && batch > 0 && batch <= 65535, "voxel_normalize: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    RAMNET_CUDA(cudaMemsetAsync(stats, 0, (size_t)batch * 3 * sizeof(double), s));
    int bx = (int)imin64((n + 255) / 256, (int64_t)h->sm_count * 8 / batch + 1);
    const dim3 blocks((unsigned)(bx < 1 ? 1 : bx), (unsigned)batch);
    voxel_stats_kernel<<<blocks, 256, 0, s>>>(grid, n, stats);
    RAMNET_LAUNCH_CHECK(h);
    voxel_normalize_kernel<<<blocks, 256, 0, s>>>(grid, n, stats);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_depth_to_label(ramnet_handle *h, const float *depth, float *label, int64_t n, float clip_distance,
                                     float reg_factor, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && depth && label && n > 0 && clip_distance > 0.f && reg_factor != 0.f, "depth_to_label: bad argument");
    const int blocks = (int)imin64((n + 255) / 256, (int64_t)h->sm_count * 8);
    depth_label_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(depth, label, n, clip_distance, reg_factor);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_depth_metrics(ramnet_handle *h, const float *pred, const float *target, int N, int64_t hw, float eps,
                                    double *out, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && pred && target && out && N > 0 && N <= 65535 && hw > 0, "depth_metrics: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    RAMNET_CUDA(cudaMemsetAsync(out, 0, (size_t)N * 8 * sizeof(double), s));
    int bx = (int)imin64((hw + 255) / 256, (int64_t)h->sm_count * 8 / N + 1);
    if (bx < 1) bx = 1;
    depth_metrics_kernel<<<dim3((unsigned)bx, (unsigned)N), 256, 0, s>>>(pred, target, hw, eps, out);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// test.py output stage (RAM_Net/test.py:259-360,365-379).  Per depth map the inference driver pulls
// the fp32 map to the host and derives from it: an 8-bit grey PNG (`cv2.imwrite(img * 255.0)`, :271), a colour-mapped
// PNG (`make_colormap`, :31-38,285), and with --calculate_scale the metric-space scale factor (:365-379).  Here one
// reduction + one streaming kernel produce those payloads on the device, so 1 + 3 bytes per pixel cross PCIe instead of
// 4 (plus 4 more for every colour map the host recomputed), and the scale needs two doubles:
//   grey[p]  = saturate_u8(rint(255 v))                                   (OpenCV's float -> 8U conversion; NaN -> 0)
//   x        = (max(v) - v) / max(max(v) - v)  with nan_to_num as make_colormap applies it (:32-35): a map that contains
//              a NaN has max = NaN, hence x = 1 everywhere; max - min == 0 gives 0
//   bgr[p]   = saturate_u8(rint(255 lut[min(int(256 x), 255)][2 - c]))    (matplotlib Colormap.__call__ LUT indexing,
//              colours reversed to BGR at :37; the LUT = the reference's color mapper sampled by the caller)
//   scale    = sum(p t) / sum(p p) with p, t = clip * exp(reg (. - 1)) over ALL pixels (:370-376; NaN propagates)
// ---------------------------------------------------------------------------------------------------------------
namespace {
// mm[0] = max key, mm[1] = ~(min key) (order-preserving uint32 keys of the floats), mm[2] = NaN flag; 4 slots per map
__device__ __forceinline__ unsigned f2key(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

__global__ void __launch_bounds__(256) depth_minmax_kernel(const float *__restrict__ d, int64_t hw, unsigned *__restrict__ mm) {
    d += (int64_t)blockIdx.y * hw;
    mm += blockIdx.y * 4;
    unsigned kmax = 0u, kmin = 0xffffffffu, nan = 0u;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = d[i];
        if (v != v) { nan = 1u; continue; }
        const unsigned k = f2key(v);
        kmax = max(kmax, k); kmin = min(kmin, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        nan |= __shfl_xor_sync(0xffffffffu, nan, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(mm + 0, kmax);
        atomicMax(mm + 1, ~kmin);          // min kept as the max of inverted keys: both slots start at 0
        if (nan) atomicOr(mm + 2, 1u);
    }
}

__device__ __forceinline__ unsigned char sat_u8(float v) {     // cv::saturate_cast<uchar>(float): cvRound, clamp; NaN -> 0
    if (!(v == v)) return 0;
    const float r = rintf(v);
    return (unsigned char)(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));
}

__global__ void __launch_bounds__(256) depth_output_kernel(const float *__restrict__ d, const float *__restrict__ target,
                                                           int64_t hw, const unsigned *__restrict__ mm,
                                                           const float *__restrict__ lut, unsigned char *__restrict__ grey,
                                                           unsigned char *__restrict__ bgr, double *__restrict__ scale_sums,
                                                           float reg_factor, float clip_distance) {
    d += (int64_t)blockIdx.y * hw;
    if (target) target += (int64_t)blockIdx.y * hw;
    if (grey) grey += (int64_t)blockIdx.y * hw;
    if (bgr) bgr += (int64_t)blockIdx.y * hw * 3;
    mm += blockIdx.y * 4;
    const bool has_nan = mm[2] != 0u;
    const float vmax = key2f(mm[0]), vmin = key2f(~mm[1]);
    const float range = vmax - vmin;                 // = amax(max - img) for a NaN-free map
    double s_pt = 0.0, s_pp = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = d[i];
        if (grey) grey[i] = sat_u8(v * 255.0f);
        if (bgr) {
            float x;
            if (has_nan) x = 1.0f;                   // amax = NaN -> all NaN -> nan_to_num(nan=1) -> / 1
            else {
                x = (vmax - v) / range;              // range == 0 -> 0/0 = NaN -> nan_to_num -> 0
                if (!(x == x)) x = 0.0f;
            }
            int idx = (int)(x * 256.0f);
            idx = idx < 0 ? 0 : (idx > 255 ? 255 : idx);
            bgr[3 * i + 0] = sat_u8(lut[3 * idx + 2] * 255.0f);
            bgr[3 * i + 1] = sat_u8(lut[3 * idx + 1] * 255.0f);
            bgr[3 * i + 2] = sat_u8(lut[3 * idx + 0] * 255.0f);
        }
        if (scale_sums) {
            const float p = clip_distance * expf(reg_factor * (v - 1.0f));
            const float t = clip_distance * expf(reg_factor * (target[i] - 1.0f));
            s_pt += (double)(p * t);
            s_pp += (double)(p * p);
        }
    }
    if (scale_sums) {
        double v2[2] = {s_pt, s_pp};
        block_reduce_add(v2, 2, scale_sums + blockIdx.y * 2);
    }
}
}  // namespace

extern "C" int ramnet_depth_output(ramnet_handle *h, const float *depth, const float *target, int N, int64_t hw,
                                   const float *lut_rgb256, unsigned char *grey, unsigned char *bgr, double *scale_sums,
                                   float reg_factor, float clip_distance, unsigned *scratch, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && depth && scratch && N > 0 && N <= 65535 && hw > 0, "depth_output: bad argument");
    RAMNET_CHECK_ARG(!bgr || lut_rgb256, "depth_output: the colour map needs the 256 x 3 LUT");
    RAMNET_CHECK_ARG(!scale_sums || target, "depth_output: the scale factor needs the target map");
    cudaStream_t s = (cudaStream_t)stream;
    // scratch: N x {max key, min key, NaN flag, pad}
    RAMNET_CUDA(cudaMemsetAsync(scratch, 0, (size_t)N * 4 * sizeof(unsigned), s));
    if (scale_sums) RAMNET_CUDA(cudaMemsetAsync(scale_sums, 0, (size_t)N * 2 * sizeof(double), s));
    int bx = (int)imin64((hw + 255) / 256, (int64_t)h->sm_count * 8 / N + 1);
    if (bx < 1) bx = 1;
    const dim3 blocks((unsigned)bx, (unsigned)N);
    if (bgr) {
        depth_minmax_kernel<<<blocks, 256, 0, s>>>(depth, hw, scratch);
        RAMNET_LAUNCH_CHECK(h);
    }
    depth_output_kernel<<<blocks, 256, 0, s>>>(depth, target, hw, scratch, lut_rgb256, grey, bgr, scale_sums, reg_factor,
                                               clip_distance);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

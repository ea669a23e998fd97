// events_to_voxel_grid: bilinear-in-time event voting.
//
// Replaces RAM_Net/utils/event_tensor_utils.py:71-117 (numpy, np.add.at) and its
// torch twin :120-187 (two index_add_ + boolean-mask compactions).
//
// Roofline: HBM / L2-atomic bound.  Algorithmic bytes = 32 B per event (one float64 row
// [t,x,y,p]) + 4*bins*H*W B for the grid written once.  Each event issues two fire-and-forget
// fp32 reductions (RED.E.ADD.F32) that resolve in L2.
//
// Layout / mapping: one thread per event, each reading its 32-byte row as two 16-byte
// vector loads; a warp therefore reads 1024 contiguous bytes (fully coalesced, streaming
// .nc, no L1 allocation).  t0 / t_last are two broadcast loads per thread served by L1/L2, so
// no host synchronisation or extra reduction pass is needed (timestamps are sorted by
// contract, exactly as the reference assumes at :86-88).  Arithmetic follows the reference
// operation by operation in float64 (B200 has full-rate-enough FP64 for 6 flops/event);
// index arithmetic is int64 and bit-exact; the vote value is rounded to fp32 once, where
// numpy casts it on accumulation.
#include <cooperative_groups.h>

#include "common.cuh"
namespace cg = cooperative_groups;

struct Vote {
    int64_t il, ir;  // flat voxel index or -1
    float vl, vr;
    bool oob;
    int64_t pix;     // x + y * width
    int ti;          // left time bin
};

__device__ __forceinline__ Vote event_vote(const double *__restrict__ ev, int64_t i, double t0, double dT, int bins,
                                           int width, int height) {
    double2 a, b;  // [t, x], [y, p]
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(a.x), "=d"(a.y) : "l"(ev + 4 * i));
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(b.x), "=d"(b.y) : "l"(ev + 4 * i + 2));
    // (num_bins - 1) * (t - first_stamp) / deltaT, left-to-right in float64 (:95)
    const double ts = __ddiv_rn(__dmul_rn((double)(bins - 1), __dsub_rn(a.x, t0)), dT);
    const int64_t x = (int64_t)a.y;   // astype(int): truncation toward zero (:96-97)
    const int64_t y = (int64_t)b.x;
    const double p = (b.y == 0.0) ? -1.0 : b.y;  // :100
    const int64_t ti = (int64_t)ts;              // :102
    const double dt = __dsub_rn(ts, (double)ti);
    Vote v;
    v.vl = (float)__dmul_rn(p, __dsub_rn(1.0, dt));  // :104, cast at accumulate
    v.vr = (float)__dmul_rn(p, dt);                  // :105
    const int64_t plane = (int64_t)width * height;
    const bool inb = (x >= 0) & (x < width) & (y >= 0) & (y < height) & (ti >= 0);
    v.oob = !inb;
    const int64_t base = x + y * width;
    v.pix = base;
    v.ti = (int)(ti < 0 ? -1 : (ti > 0x7fffffff ? 0x7fffffff : ti));
    v.il = (inb && ti < bins) ? base + ti * plane : -1;            // :107-109
    v.ir = (inb && ti + 1 < bins) ? base + (ti + 1) * plane : -1;  // :111-113
    return v;
}

__device__ __forceinline__ void red_add_f32(float *addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__global__ void __launch_bounds__(256) voxel_grid_kernel(const double *__restrict__ ev, int64_t n, int bins, int width,
                                                         int height, float *__restrict__ grid,
                                                         int32_t *__restrict__ oob_count) {
    const double t0 = __ldg(ev);
    double dT = __dsub_rn(__ldg(ev + 4 * (n - 1)), t0);
    if (dT == 0.0) dT = 1.0;  // :92-93
    int oob = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Vote v = event_vote(ev, i, t0, dT, bins, width, height);
        if (v.il >= 0) red_add_f32(grid + v.il, v.vl);
        if (v.ir >= 0) red_add_f32(grid + v.ir, v.vr);
        oob += v.oob;
    }
    if (oob_count != nullptr && oob) atomicAdd(oob_count, oob);
}

// Same kernel with the zero fill of the grid folded in (small event counts: the separate 2.6 MB memset node and the
// kernel boundary behind it were ~3 of the 8.3 us the call took up to 1e5 events).  Cooperative launch: every block is
// resident, so the grid-wide barrier between "zero" and "vote" cannot deadlock; the runtime refuses the launch otherwise
// and the caller falls back to memset + voxel_grid_kernel.
__global__ void __launch_bounds__(256) voxel_grid_fused_kernel(const double *__restrict__ ev, int64_t n, int bins, int width,
                                                               int height, float *__restrict__ grid, int64_t total,
                                                               int32_t *__restrict__ oob_count) {
    cg::grid_group gg = cg::this_grid();
    const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, gsize = (int64_t)gridDim.x * blockDim.x;
    float4 *g4 = reinterpret_cast<float4 *>(grid);       // 16-byte aligned (checked by the launcher)
    const int64_t total4 = total >> 2;
    for (int64_t i = gtid; i < total4; i += gsize) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = (total4 << 2) + gtid; i < total; i += gsize) grid[i] = 0.f;
    if (oob_count != nullptr && gtid == 0) *oob_count = 0;
    gg.sync();
    const double t0 = __ldg(ev);
    double dT = __dsub_rn(__ldg(ev + 4 * (n - 1)), t0);
    if (dT == 0.0) dT = 1.0;  // :92-93
    int oob = 0;
    for (int64_t i = gtid; i < n; i += gsize) {
        const Vote v = event_vote(ev, i, t0, dT, bins, width, height);
        if (v.il >= 0) red_add_f32(grid + v.il, v.vl);
        if (v.ir >= 0) red_add_f32(grid + v.ir, v.vr);
        oob += v.oob;
    }
    if (oob_count != nullptr && oob) atomicAdd(oob_count, oob);
}

__global__ void __launch_bounds__(256) voxel_votes_kernel(const double *__restrict__ ev, int64_t n, int bins, int width,
                                                          int height, int64_t *__restrict__ il, float *__restrict__ vl,
                                                          int64_t *__restrict__ ir, float *__restrict__ vr) {
    const double t0 = __ldg(ev);
    double dT = __dsub_rn(__ldg(ev + 4 * (n - 1)), t0);
    if (dT == 0.0) dT = 1.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Vote v = event_vote(ev, i, t0, dT, bins, width, height);
        il[i] = v.il; vl[i] = v.vl; ir[i] = v.ir; vr[i] = v.vr;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Packed accumulation for large event counts.  At 1e7 events the kernel above is bound by the ~2e7 fp32 reductions
// the L2 has to resolve (147 us against a 49 us HBM floor, profiles/r01_voxel_bench.json), not by the event stream.
// Both votes of an event go to the SAME pixel, time bins ti and ti + 1, so with the bins of a pixel adjacent in memory
// (accumulator [H*W][8] floats, one 32-byte sector per pixel) they are ONE vector reduction
// (red.global.add.v4.f32 -> REDG.E.ADD.F32x4: slots ti%4, ti%4+1 carry the votes, the others +0.0) unless the pair
// straddles the two quads (ti % 4 == 3), which takes the two scalar reductions.  A second pass transposes the
// accumulator into the reference's [bins, H, W] layout and can fold in the statistics of the non-zero voxels that the
// loaders' normalisation needs (event_dataset.py:144-151), so that "scatter -> normalise" never re-reads the grid for them.
// Per-voxel accumulation order differs from the kernel above only in which votes are summed first: same tolerance
// (<= 1 ulp per accumulation, fp32 atomics reorder sums either way), indices bit-exact.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(256) voxel_grid_packed_kernel(const double *__restrict__ ev, int64_t n, int bins, int width,
                                                                int height, float *__restrict__ acc,
                                                                int32_t *__restrict__ oob_count) {
    const double t0 = __ldg(ev);
    double dT = __dsub_rn(__ldg(ev + 4 * (n - 1)), t0);
    if (dT == 0.0) dT = 1.0;
    int oob = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Vote v = event_vote(ev, i, t0, dT, bins, width, height);
        oob += v.oob;
        if (v.il < 0 && v.ir < 0) continue;
        float *cell = acc + v.pix * 8;
        const int j = v.ti & 3;
        if (v.il >= 0 && v.ir >= 0 && j != 3) {
            float *quad = cell + (v.ti & ~3);
            if (j == 0) red_add_v4(quad, v.vl, v.vr, 0.f, 0.f);
            else if (j == 1) red_add_v4(quad, 0.f, v.vl, v.vr, 0.f);
            else red_add_v4(quad, 0.f, 0.f, v.vl, v.vr);
        } else {
            if (v.il >= 0) red_add_f32(cell + v.ti, v.vl);
            if (v.ir >= 0) red_add_f32(cell + v.ti + 1, v.vr);
        }
    }
    if (oob_count != nullptr && oob) atomicAdd(oob_count, oob);
}

// acc [H*W][8] -> grid [bins][H*W]; stats (nullable) += (sum, sum of squares, count) of the non-zero voxels
__global__ void __launch_bounds__(256) voxel_unpack_kernel(const float4 *__restrict__ acc, float *__restrict__ grid, int bins,
                                                           int64_t hw, double *__restrict__ stats) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < hw; p += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = acc[2 * p], b = acc[2 * p + 1];
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < bins) {
                grid[(int64_t)k * hw + p] = v[k];
                if (v[k] != 0.f) { s0 += (double)v[k]; s1 += (double)v[k] * (double)v[k]; s2 += 1.0; }
            }
    }
    if (stats != nullptr) {
        __shared__ double sh[3][8];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) { sh[0][warp] = s0; sh[1][warp] = s1; sh[2][warp] = s2; }
        __syncthreads();
        if (threadIdx.x < 3) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
            atomicAdd(stats + threadIdx.x, t);
        }
    }
}

// stats of the non-zero voxels of a finished [bins*H*W] grid (path without the packed accumulator)
__global__ void __launch_bounds__(256) voxel_nz_stats_kernel(const float *__restrict__ grid, int64_t n, double *__restrict__ stats) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = grid[i];
        if (x != 0.f) { s0 += (double)x; s1 += (double)x * (double)x; s2 += 1.0; }
    }
    __shared__ double sh[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) { sh[0][warp] = s0; sh[1][warp] = s1; sh[2][warp] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
        atomicAdd(stats + threadIdx.x, t);
    }
}

// (x - mean) / stddev on the non-zero voxels, nothing when there are none or the stddev is 0 (event_dataset.py:144-151)
__global__ void __launch_bounds__(256) voxel_apply_norm_kernel(float *__restrict__ grid, int64_t n, const double *__restrict__ stats) {
    const double cnt = stats[2];
    if (!(cnt > 0.0)) return;
    const double mean = stats[0] / cnt;
    double var = stats[1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const double sd = sqrt(var);
    if (!(sd > 0.0)) return;
    const float m = (float)mean, inv = (float)(1.0 / sd);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = grid[i];
        if (x != 0.f) grid[i] = (x - m) * inv;
    }
}

static int check_args(ramnet_handle *h, const double *events, int64_t n, int bins, int width, int height) {
    RAMNET_CHECK_ARG(h != nullptr, "voxel_grid: handle is NULL");
    RAMNET_CHECK_ARG(n >= 0, "voxel_grid: n < 0");
    RAMNET_CHECK_ARG(n == 0 || events != nullptr, "voxel_grid: events is NULL");
    RAMNET_CHECK_ARG(bins > 0 && width > 0 && height > 0, "voxel_grid: num_bins/width/height must be > 0 (got %d,%d,%d)",
                     bins, width, height);  // the reference's asserts, event_tensor_utils.py:80-83
    RAMNET_CHECK_ARG(((uintptr_t)events & 15) == 0, "voxel_grid: events must be 16-byte aligned");
    return RAMNET_OK;
}

// Zero fill + votes in one cooperative launch; returns false when that path does not apply (the caller then runs
// cudaMemsetAsync + voxel_grid_kernel).  RAMNET_VOXEL_FUSED=0 disables it.
static bool launch_voxel_fused(ramnet_handle *h, const double *events, int64_t n, int bins, int width, int height, float *grid,
                               int32_t *oob_count, cudaStream_t s) {
    static const bool enabled = [] { const char *e = getenv("RAMNET_VOXEL_FUSED"); return !(e && e[0] == '0'); }();
    static const int64_t fused_max = [] { const char *e = getenv("RAMNET_VOXEL_FUSED_MAX"); return e ? atoll(e) : 400000ll; }();
    if (!enabled || n <= 0 || n > fused_max || (((uintptr_t)grid) & 15)) return false;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;      // a refused launch must not invalidate a capture
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
        (void)cudaGetLastError();
        return false;
    }
    static int per_sm = -1, coop = -1;
    if (coop < 0) {
        int dev = h->device;
        if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) coop = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, voxel_grid_fused_kernel, 256, 0) != cudaSuccess) per_sm = 0;
    }
    if (!coop || per_sm <= 0) return false;
    int64_t total = (int64_t)bins * width * height;
    int64_t want = (n + 255) / 256;
    if (want < (int64_t)h->sm_count * 2) want = (int64_t)h->sm_count * 2;       // enough stores in flight for the zero fill
    const int blocks = (int)imin64(want, (int64_t)h->sm_count * imin64(per_sm, 4));
    void *args[] = {(void *)&events, (void *)&n, (void *)&bins, (void *)&width, (void *)&height, (void *)&grid, (void *)&total,
                    (void *)&oob_count};
    if (cudaLaunchCooperativeKernel((const void *)voxel_grid_fused_kernel, dim3(blocks), dim3(256), args, 0, s) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    h->launches++;
    return true;
}

extern "C" int ramnet_voxel_grid(ramnet_handle *h, const double *events, int64_t n, int bins, int width, int height,
                                 float *grid, int32_t *oob_count, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    int rc = check_args(h, events, n, bins, width, height);
    if (rc) return rc;
    RAMNET_CHECK_ARG(grid != nullptr, "voxel_grid: grid is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    if (launch_voxel_fused(h, events, n, bins, width, height, grid, oob_count, s)) return RAMNET_OK;
    RAMNET_CUDA(cudaMemsetAsync(grid, 0, sizeof(float) * (size_t)bins * width * height, s));
    if (oob_count) RAMNET_CUDA(cudaMemsetAsync(oob_count, 0, sizeof(int32_t), s));
    if (n == 0) return RAMNET_OK;
    // grid: a multiple of the SM count, 8 resident 256-thread CTAs per SM at most
    const int64_t want = (n + 255) / 256;
    const int blocks = (int)imin64(want, (int64_t)h->sm_count * 8);
    voxel_grid_kernel<<<blocks, 256, 0, s>>>(events, n, bins, width, height, grid, oob_count);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_voxel_votes(ramnet_handle *h, const double *events, int64_t n, int bins, int width, int height,
                                  int64_t *idx_left, float *val_left, int64_t *idx_right, float *val_right,
                                  void *stream) {
    RAMNET_DEVICE_GUARD(h);
    int rc = check_args(h, events, n, bins, width, height);
    if (rc) return rc;
    if (n == 0) return RAMNET_OK;
    RAMNET_CHECK_ARG(idx_left && val_left && idx_right && val_right, "voxel_votes: NULL output");
    const int blocks = (int)imin64((n + 255) / 256, (int64_t)h->sm_count * 8);
    voxel_votes_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(events, n, bins, width, height, idx_left, val_left,
                                                                idx_right, val_right);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" size_t ramnet_voxel_grid_workspace_bytes(int bins, int width, int height) {
    return (bins > 0 && bins <= 8 && width > 0 && height > 0) ? (size_t)width * height * 8 * sizeof(float) : 0;
}

// events -> grid (+ optional statistics of the non-zero voxels, + optional normalisation) in one call.
// workspace (nullable, ramnet_voxel_grid_workspace_bytes): enables the packed accumulator for n >= packed_min events.
extern "C" int ramnet_voxel_grid_ex(ramnet_handle *h, const double *events, int64_t n, int bins, int width, int height,
                                    float *grid, int32_t *oob_count, void *workspace, size_t workspace_bytes,
                                    double *stats, int flags, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    int rc = check_args(h, events, n, bins, width, height);
    if (rc) return rc;
    RAMNET_CHECK_ARG(grid != nullptr, "voxel_grid: grid is NULL");
    RAMNET_CHECK_ARG(!(flags & RAMNET_VOXEL_NORMALIZE) || stats, "voxel_grid: normalisation needs the stats buffer (3 doubles)");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t hw = (int64_t)width * height, total = hw * bins;
    if (oob_count) RAMNET_CUDA(cudaMemsetAsync(oob_count, 0, sizeof(int32_t), s));
    if (stats) RAMNET_CUDA(cudaMemsetAsync(stats, 0, 3 * sizeof(double), s));
    static const int64_t packed_min = [] { const char *e = getenv("RAMNET_VOXEL_PACKED_MIN"); return e ? atoll(e) : 1500000ll; }();
    const size_t need = ramnet_voxel_grid_workspace_bytes(bins, width, height);
    const bool packed = workspace != nullptr && need > 0 && workspace_bytes >= need && n >= packed_min &&
                        (((uintptr_t)workspace) & 15) == 0;
    const int pw_blocks = (int)imin64((total + 255) / 256, (int64_t)h->sm_count * 8);
    if (packed) {
        RAMNET_CUDA(cudaMemsetAsync(workspace, 0, need, s));
        voxel_grid_packed_kernel<<<(int)imin64((n + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, s>>>(
            events, n, bins, width, height, (float *)workspace, oob_count);
        RAMNET_LAUNCH_CHECK(h);
        voxel_unpack_kernel<<<(int)imin64((hw + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, s>>>(
            (const float4 *)workspace, grid, bins, hw, stats);
        RAMNET_LAUNCH_CHECK(h);
    } else if (launch_voxel_fused(h, events, n, bins, width, height, grid, oob_count, s)) {
        if (stats) {
            voxel_nz_stats_kernel<<<pw_blocks, 256, 0, s>>>(grid, total, stats);
            RAMNET_LAUNCH_CHECK(h);
        }
    } else {
        RAMNET_CUDA(cudaMemsetAsync(grid, 0, sizeof(float) * (size_t)total, s));
        if (n > 0) {
            voxel_grid_kernel<<<(int)imin64((n + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, s>>>(events, n, bins, width,
                                                                                                    height, grid, oob_count);
            RAMNET_LAUNCH_CHECK(h);
        }
        if (stats && n > 0) {
            voxel_nz_stats_kernel<<<pw_blocks, 256, 0, s>>>(grid, total, stats);
            RAMNET_LAUNCH_CHECK(h);
        }
    }
    if ((flags & RAMNET_VOXEL_NORMALIZE) && n > 0) {
        voxel_apply_norm_kernel<<<pw_blocks, 256, 0, s>>>(grid, total, stats);
        RAMNET_LAUNCH_CHECK(h);
    }
    return RAMNET_OK;
}

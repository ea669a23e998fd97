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
#include "common.cuh"

struct Vote {
    int64_t il, ir;  // flat voxel index or -1
    float vl, vr;
    bool oob;
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

static int check_args(ramnet_handle *h, const double *events, int64_t n, int bins, int width, int height) {
    RAMNET_CHECK_ARG(h != nullptr, "voxel_grid: handle is NULL");
    RAMNET_CHECK_ARG(n >= 0, "voxel_grid: n < 0");
    RAMNET_CHECK_ARG(n == 0 || events != nullptr, "voxel_grid: events is NULL");
    RAMNET_CHECK_ARG(bins > 0 && width > 0 && height > 0, "voxel_grid: num_bins/width/height must be > 0 (got %d,%d,%d)",
                     bins, width, height);  // the reference's asserts, event_tensor_utils.py:80-83
    RAMNET_CHECK_ARG(((uintptr_t)events & 15) == 0, "voxel_grid: events must be 16-byte aligned");
    return RAMNET_OK;
}

extern "C" int ramnet_voxel_grid(ramnet_handle *h, const double *events, int64_t n, int bins, int width, int height,
                                 float *grid, int32_t *oob_count, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    int rc = check_args(h, events, n, bins, width, height);
    if (rc) return rc;
    RAMNET_CHECK_ARG(grid != nullptr, "voxel_grid: grid is NULL");
    cudaStream_t s = (cudaStream_t)stream;
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

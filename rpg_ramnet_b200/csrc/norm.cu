// Live normalisation layers (SURVEY.md §8 a-3 / a-6 / a-7 / a-8 with config key norm = 'BN' | 'IN'):
// nn.BatchNorm2d in train mode (batch statistics + running-statistics update, submodules.py:21-22,29-30,189-190),
// nn.InstanceNorm2d(track_running_stats=True) in train mode (:23-24) and the ResidualBlock's plain
// nn.InstanceNorm2d (:192-194, per-instance statistics in train AND eval mode), each followed by the layer's
// activation (+ the residual add of ResidualBlock.forward :213-214).  Eval-mode norms with running statistics never get
// here: they are folded into the packed conv weights (engine._fold_norm).
//
// All tensors NHWC fp32 [N, H*W, C].  Statistics are accumulated in float64 (ATen's CPU kernels use double accumulators
// for float input; E[z^2] - E[z]^2 in double has no cancellation problem at these magnitudes).  Every kernel is a
// flat, fully coalesced float4 stream whose grid stride is a multiple of C/4, so a thread owns the same four channels
// for its whole life: per-channel constants sit in registers and the per-channel sums fold through shared memory into
// one float64 atomic per channel per block.  HBM-bound: forward = read z twice + write y (12 B / element), backward =
// read (dy, y, z) twice + write dz (28 B / element).
#include "common.cuh"

namespace {

template <int V> struct Vec;
template <> struct Vec<4> {
    typedef float4 T;
    static __device__ __forceinline__ void get(const float4 &v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
    static __device__ __forceinline__ float4 put(const float (&o)[4]) { return make_float4(o[0], o[1], o[2], o[3]); }
};
template <> struct Vec<1> {
    typedef float T;
    static __device__ __forceinline__ void get(const float &v, float (&o)[1]) { o[0] = v; }
    static __device__ __forceinline__ float put(const float (&o)[1]) { return o[0]; }
};

__device__ __forceinline__ float act_fwd(float v, int flags) {
    if (flags & RAMNET_NORM_RELU) return fmaxf(v, 0.f);
    if (flags & RAMNET_NORM_SIGMOID) return sigmoidf_(v);
    return v;
}
// derivative of the activation expressed through its output
__device__ __forceinline__ float act_bwd(float dy, float y, int flags) {
    if (flags & RAMNET_NORM_RELU) return y > 0.f ? dy : 0.f;
    if (flags & RAMNET_NORM_SIGMOID) return dy * y * (1.f - y);
    return dy;
}

// The per-(group, channel) sums live in kReplicas copies ([G][kReplicas][C][2] doubles): block b adds into copy
// b % kReplicas and the finalize kernels add the copies up.  One copy made ~1000 blocks queue their float64 atomics on
// the same 2 C addresses (ncu: the BatchNorm statistics kernel took 49 us for 33 MB, the InstanceNorm one -- 4 x the
// addresses -- 29 us).
constexpr int kReplicas = 16;


// Folds the block's per-thread float64 partial sums (NS per channel lane) and adds them to out[(c * NS) + k].
template <int V, int NS>
__device__ __forceinline__ void fold_to_global(double (&acc)[V][NS], int CV, double *__restrict__ out) {
    __shared__ double part[256 * V * NS];
#pragma unroll
    for (int e = 0; e < V; ++e)
#pragma unroll
        for (int k = 0; k < NS; ++k) part[(threadIdx.x * V + e) * NS + k] = acc[e][k];
    __syncthreads();
    if ((int)threadIdx.x < CV) {
#pragma unroll
        for (int e = 0; e < V; ++e)
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                double s = 0.0;
                for (int j = threadIdx.x; j < 256; j += CV) s += part[(j * V + e) * NS + k];
                atomicAdd(out + ((int64_t)threadIdx.x * V + e) * NS + k, s);
            }
    }
}

// sums[g][c] = (sum z, sum z^2) over the rows of group g (grid.y = groups)
template <int V>
__global__ void __launch_bounds__(256) norm_stats_kernel(const float *__restrict__ z, int64_t vec_per_group, int C,
                                                         double *__restrict__ sums) {
    typedef typename Vec<V>::T T;
    const int CV = C / V;
    const T *zg = reinterpret_cast<const T *>(z) + (int64_t)blockIdx.y * vec_per_group;
    double acc[V][2];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e][0] = acc[e][1] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    // U independent loads in flight per thread; every element goes into the float64 sums (fp32 partial sums of z^2 over
    // the run were measured to move a gradient sample of the reference golden by 6e-3, at no gain in time)
    constexpr int U = 8;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < vec_per_group; i += stride * U) {
        T ld[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < vec_per_group) ld[u] = zg[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < vec_per_group) {
                float v[V];
                Vec<V>::get(ld[u], v);
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const double d = (double)v[e];
                    acc[e][0] += d;
                    acc[e][1] = fma(d, d, acc[e][1]);
                }
            }
    }
    fold_to_global<V, 2>(acc, CV, sums + ((int64_t)blockIdx.y * kReplicas + blockIdx.x % kReplicas) * C * 2);
}

// (sum, sum of squares) -> (mean, 1/sqrt(biased var + eps)); running statistics as nn.BatchNorm2d / F.instance_norm
// update them: running = (1 - momentum) * running + momentum * stat, with the UNBIASED variance, averaged over the
// instances for InstanceNorm (ATen runs it as a batch norm over [1, N*C, H, W] and averages the N updated copies).
__device__ __forceinline__ void replica_sum(const double *__restrict__ sums, int g, int c, int C, double &s0, double &s1) {
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
    for (int r = 0; r < kReplicas; r += 2) {      // 2 * kReplicas independent loads
        const double *p = sums + (((int64_t)g * kReplicas + r) * C + c) * 2;
        a0 += p[0];
        a1 += p[1];
        b0 += p[(int64_t)C * 2];
        b1 += p[(int64_t)C * 2 + 1];
    }
    s0 = a0 + b0;
    s1 = a1 + b1;
}

__global__ void norm_finalize_kernel(const double *__restrict__ sums, int G, int C, double count, double eps,
                                     double momentum, float *__restrict__ running_mean, float *__restrict__ running_var,
                                     float *__restrict__ stats) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (group, channel)
    if (idx >= G * C) return;
    const int g = idx / C, c = idx - g * C;
    const double unbias = count / (count > 1.0 ? count - 1.0 : 1.0);
    double s, ss;
    replica_sum(sums, g, c, C, s, ss);
    const double mean = s / count;
    double var = ss / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[(int64_t)idx * 2] = (float)mean;
    stats[(int64_t)idx * 2 + 1] = (float)(1.0 / sqrt(var + eps));
    if (g == 0 && (running_mean || running_var)) {
        double mean_acc = mean, var_acc = var * unbias;
        for (int gg = 1; gg < G; ++gg) {        // InstanceNorm with running statistics: average over the instances
            replica_sum(sums, gg, c, C, s, ss);
            const double m = s / count;
            double v = ss / count - m * m;
            if (v < 0.0) v = 0.0;
            mean_acc += m;
            var_acc += v * unbias;
        }
        if (running_mean) running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean_acc / G);
        if (running_var) running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * var_acc / G);
    }
}

// RAMNET_NORM_RUNNING: the statistics are the running ones (an eval-mode norm that gradients flow through)
__global__ void norm_running_stats_kernel(const float *__restrict__ running_mean, const float *__restrict__ running_var,
                                          int C, double eps, float *__restrict__ stats) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    stats[2 * c] = running_mean[c];
    stats[2 * c + 1] = (float)(1.0 / sqrt((double)running_var[c] + eps));
}

// y = act((z - mean) * invstd * gamma + beta (+ res))
template <int V>
__global__ void __launch_bounds__(256) norm_apply_kernel(const float *__restrict__ z, const float *__restrict__ res,
                                                         const float *__restrict__ gamma, const float *__restrict__ beta,
                                                         const float *__restrict__ stats,
                                                         int64_t vec_per_group, int C, int flags, float *__restrict__ y) {
    typedef typename Vec<V>::T T;
    const int CV = C / V;
    const int q = threadIdx.x % CV;
    const float *st = stats + (int64_t)blockIdx.y * C * 2;
    float mean[V], sc[V], sh[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const int c = q * V + e;
        mean[e] = st[2 * c];
        sc[e] = st[2 * c + 1] * (gamma ? gamma[c] : 1.f);
        sh[e] = beta ? beta[c] : 0.f;
    }
    const int64_t base = (int64_t)blockIdx.y * vec_per_group;
    const T *zg = reinterpret_cast<const T *>(z) + base;
    const T *rg = res ? reinterpret_cast<const T *>(res) + base : nullptr;
    T *yg = reinterpret_cast<T *>(y) + base;
    const int64_t stride = (int64_t)gridDim.x * 256;
#pragma unroll 4
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < vec_per_group; i += stride) {
        float v[V], r[V], o[V];
        Vec<V>::get(zg[i], v);
        if (rg) Vec<V>::get(rg[i], r);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            float t = fmaf(v[e] - mean[e], sc[e], sh[e]);
            if (rg) t += r[e];
            t = act_fwd(t, flags);
            o[e] = (flags & RAMNET_NORM_ROUND_TF32) ? round_tf32(t) : t;
        }
        yg[i] = Vec<V>::put(o);
    }
}

// backward pass 1: g = dy * act'(y); sums[g][c] = (sum g, sum g * xhat); dres = g (the residual branch's gradient)
template <int V>
__global__ void __launch_bounds__(256) norm_bwd_reduce_kernel(const float *__restrict__ dy, const float *__restrict__ y,
                                                              const float *__restrict__ z, const float *__restrict__ stats,
                                                              int64_t vec_per_group, int C, int flags,
                                                              double *__restrict__ sums, float *__restrict__ dres) {
    typedef typename Vec<V>::T T;
    const int CV = C / V;
    const int q = threadIdx.x % CV;
    const float *st = stats + (int64_t)blockIdx.y * C * 2;
    float mean[V], inv[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { mean[e] = st[2 * (q * V + e)]; inv[e] = st[2 * (q * V + e) + 1]; }
    const int64_t base = (int64_t)blockIdx.y * vec_per_group;
    const T *dyg = reinterpret_cast<const T *>(dy) + base, *zg = reinterpret_cast<const T *>(z) + base;
    const T *yg = y ? reinterpret_cast<const T *>(y) + base : nullptr;
    T *drg = dres ? reinterpret_cast<T *>(dres) + base : nullptr;
    double acc[V][2];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e][0] = acc[e][1] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    constexpr int U = 4;          // U x 3 independent loads in flight per thread
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < vec_per_group; i += stride * U) {
        T ld_d[U], ld_z[U], ld_y[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < vec_per_group) {
                ld_d[u] = dyg[i + u * stride];
                ld_z[u] = zg[i + u * stride];
                if (yg) ld_y[u] = yg[i + u * stride];
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < vec_per_group) {
                float d[V], o[V], v[V], g[V];
                Vec<V>::get(ld_d[u], d);
                Vec<V>::get(ld_z[u], v);
                if (yg) Vec<V>::get(ld_y[u], o);
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    g[e] = yg ? act_bwd(d[e], o[e], flags) : d[e];
                    acc[e][0] += (double)g[e];
                    acc[e][1] = fma((double)g[e], (double)((v[e] - mean[e]) * inv[e]), acc[e][1]);
                }
                if (drg) drg[i + u * stride] = Vec<V>::put(g);
            }
    }
    fold_to_global<V, 2>(acc, CV, sums + ((int64_t)blockIdx.y * kReplicas + blockIdx.x % kReplicas) * C * 2);
}

// (sum g, sum g xhat) -> per-(group, channel) means the data gradient subtracts; dgamma / dbeta accumulate (+=)
__global__ void norm_bwd_finalize_kernel(const double *__restrict__ sums, int G, int C, double count, int running,
                                         float *__restrict__ coef, float *__restrict__ dgamma, float *__restrict__ dbeta) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (group, channel)
    if (idx >= G * C) return;
    const int g = idx / C, c = idx - g * C;
    double a, b;
    replica_sum(sums, g, c, C, a, b);
    coef[(int64_t)idx * 2] = running ? 0.f : (float)(a / count);
    coef[(int64_t)idx * 2 + 1] = running ? 0.f : (float)(b / count);
    if (g == 0 && (dgamma || dbeta)) {
        double s1 = a, s2 = b;
        for (int gg = 1; gg < G; ++gg) {
            replica_sum(sums, gg, c, C, a, b);
            s1 += a;
            s2 += b;
        }
        if (dbeta) dbeta[c] += (float)s1;
        if (dgamma) dgamma[c] += (float)s2;
    }
}

// backward pass 2: dz = gamma * invstd * (g - mean(g) - xhat * mean(g xhat))
template <int V>
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const float *__restrict__ dy, const float *__restrict__ y,
                                                             const float *__restrict__ z, const float *__restrict__ stats,
                                                             const float *__restrict__ coef,
                                                             const float *__restrict__ gamma, int64_t vec_per_group, int C,
                                                             int flags, float *__restrict__ dz) {
    typedef typename Vec<V>::T T;
    const int CV = C / V;
    const int q = threadIdx.x % CV;
    const float *st = stats + (int64_t)blockIdx.y * C * 2;
    const float *cf = coef + (int64_t)blockIdx.y * C * 2;
    float mean[V], inv[V], sc[V], a[V], b[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const int c = q * V + e;
        mean[e] = st[2 * c];
        inv[e] = st[2 * c + 1];
        sc[e] = inv[e] * (gamma ? gamma[c] : 1.f);
        a[e] = cf[2 * c];
        b[e] = cf[2 * c + 1];
    }
    const int64_t base = (int64_t)blockIdx.y * vec_per_group;
    const T *dyg = reinterpret_cast<const T *>(dy) + base, *zg = reinterpret_cast<const T *>(z) + base;
    const T *yg = y ? reinterpret_cast<const T *>(y) + base : nullptr;
    T *dzg = reinterpret_cast<T *>(dz) + base;
    const int64_t stride = (int64_t)gridDim.x * 256;
#pragma unroll 4
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < vec_per_group; i += stride) {
        float d[V], o[V], v[V], r[V];
        Vec<V>::get(dyg[i], d);
        Vec<V>::get(zg[i], v);
        if (yg) Vec<V>::get(yg[i], o);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const float g = yg ? act_bwd(d[e], o[e], flags) : d[e];
            const float t = sc[e] * (g - a[e] - (v[e] - mean[e]) * inv[e] * b[e]);
            r[e] = (flags & RAMNET_NORM_ROUND_TF32) ? round_tf32(t) : t;
        }
        dzg[i] = Vec<V>::put(r);
    }
}

struct NormGeom {
    int V, G;
    int64_t vec_per_group;
    double count;
    dim3 grid;
};

// Vector width and grid: 256 % (C / V) == 0 keeps a thread on the same channels under a grid stride of 256 * blocks.
int plan_norm(const ramnet_handle *h, const void *p0, int N, int64_t HW, int C, int flags, NormGeom *g) {
    const bool inst = (flags & RAMNET_NORM_INSTANCE) != 0;
    if (C % 4 == 0 && 256 % (C / 4) == 0 && (((uintptr_t)p0) & 15) == 0 && ((HW * C) % 4 == 0)) g->V = 4;
    else if (C <= 256 && 256 % C == 0) g->V = 1;
    else return ramnet_set_error(RAMNET_EUNSUPPORTED, "norm: C=%d (supported: C/4 or C a divisor of 256)", C);
    g->G = inst ? N : 1;
    const int64_t rows = inst ? HW : HW * N;
    g->count = (double)rows;
    g->vec_per_group = rows * (C / g->V);
    int64_t blocks = (g->vec_per_group + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = (int64_t)h->sm_count * 4 / g->G > 0 ? (int64_t)h->sm_count * 4 / g->G : 1;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    g->grid = dim3((unsigned)blocks, (unsigned)g->G, 1);
    return 0;
}
}  // namespace

extern "C" size_t ramnet_norm_scratch_bytes(int N, int C) { return (size_t)N * kReplicas * C * 2 * sizeof(double); }

extern "C" int ramnet_norm_fwd(ramnet_handle *h, const float *z, const float *res, const float *gamma,
                               const float *beta, float *running_mean, float *running_var, double momentum, double eps,
                               int N, int64_t HW, int C, int flags, double *sums, float *stats, float *y,
                               void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && z && stats && y && N > 0 && HW > 0 && C > 0, "norm_fwd: bad argument");
    const bool running = (flags & RAMNET_NORM_RUNNING) != 0;
    RAMNET_CHECK_ARG(!running || (running_mean && running_var), "norm_fwd: RAMNET_NORM_RUNNING needs running statistics");
    RAMNET_CHECK_ARG(running || sums, "norm_fwd: the statistics scratch is missing");
    if (running) flags &= ~RAMNET_NORM_INSTANCE;
    NormGeom g;
    if (int rc = plan_norm(h, z, N, HW, C, flags, &g)) return rc;
    if (res && g.V == 4 && (((uintptr_t)res) & 15)) g.V = 1, g.vec_per_group *= 4;
    if (g.V == 1 && (C > 256 || 256 % C)) return ramnet_set_error(RAMNET_EUNSUPPORTED, "norm_fwd: unaligned operand with C=%d", C);
    cudaStream_t s = (cudaStream_t)stream;
    const int cb = (C + 127) / 128, gcb = (g.G * C + 127) / 128;
    if (running) {
        norm_running_stats_kernel<<<cb, 128, 0, s>>>(running_mean, running_var, C, eps, stats);
        RAMNET_LAUNCH_CHECK(h);
    } else {
        RAMNET_CUDA(cudaMemsetAsync(sums, 0, (size_t)g.G * kReplicas * C * 2 * sizeof(double), s));
        if (g.V == 4) norm_stats_kernel<4><<<g.grid, 256, 0, s>>>(z, g.vec_per_group, C, sums);
        else norm_stats_kernel<1><<<g.grid, 256, 0, s>>>(z, g.vec_per_group, C, sums);
        RAMNET_LAUNCH_CHECK(h);
        norm_finalize_kernel<<<gcb, 128, 0, s>>>(sums, g.G, C, g.count, eps, momentum, running_mean, running_var, stats);
        RAMNET_LAUNCH_CHECK(h);
    }
    if (g.V == 4) norm_apply_kernel<4><<<g.grid, 256, 0, s>>>(z, res, gamma, beta, stats, g.vec_per_group, C, flags, y);
    else norm_apply_kernel<1><<<g.grid, 256, 0, s>>>(z, res, gamma, beta, stats, g.vec_per_group, C, flags, y);
    RAMNET_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int ramnet_norm_bwd(ramnet_handle *h, const float *dy, const float *y, const float *z, const float *stats,
                               const float *gamma, int N, int64_t HW, int C, int flags, double *sums, float *coef,
                               float *dz, float *dres, float *dgamma, float *dbeta, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && dy && z && stats && sums && coef && dz && N > 0 && HW > 0 && C > 0, "norm_bwd: bad argument");
    RAMNET_CHECK_ARG(y || !(flags & (RAMNET_NORM_RELU | RAMNET_NORM_SIGMOID)), "norm_bwd: the activation's adjoint needs y");
    const bool running = (flags & RAMNET_NORM_RUNNING) != 0;
    if (running) flags &= ~RAMNET_NORM_INSTANCE;
    NormGeom g;
    if (int rc = plan_norm(h, z, N, HW, C, flags, &g)) return rc;
    if (g.V == 4 && ((((uintptr_t)dy) | ((uintptr_t)dz) | ((uintptr_t)y) | ((uintptr_t)dres)) & 15)) {
        if (C > 256 || 256 % C) return ramnet_set_error(RAMNET_EUNSUPPORTED, "norm_bwd: unaligned operand with C=%d", C);
        g.V = 1, g.vec_per_group *= 4;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int gcb = (g.G * C + 127) / 128;
    RAMNET_CUDA(cudaMemsetAsync(sums, 0, (size_t)g.G * kReplicas * C * 2 * sizeof(double), s));
    if (g.V == 4) norm_bwd_reduce_kernel<4><<<g.grid, 256, 0, s>>>(dy, y, z, stats, g.vec_per_group, C, flags, sums, dres);
    else norm_bwd_reduce_kernel<1><<<g.grid, 256, 0, s>>>(dy, y, z, stats, g.vec_per_group, C, flags, sums, dres);
    RAMNET_LAUNCH_CHECK(h);
    norm_bwd_finalize_kernel<<<gcb, 128, 0, s>>>(sums, g.G, C, g.count, running ? 1 : 0, coef, dgamma, dbeta);
    RAMNET_LAUNCH_CHECK(h);
    if (g.V == 4) norm_bwd_apply_kernel<4><<<g.grid, 256, 0, s>>>(dy, y, z, stats, coef, gamma, g.vec_per_group, C, flags, dz);
    else norm_bwd_apply_kernel<1><<<g.grid, 256, 0, s>>>(dy, y, z, stats, coef, gamma, g.vec_per_group, C, flags, dz);
    RAMNET_LAUNCH_CHECK(h);
    return 0;
}

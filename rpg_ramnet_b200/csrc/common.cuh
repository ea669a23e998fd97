// Shared device/host helpers for libramnet_sm100a.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ramnet_b200.h"

struct ramnet_handle {
    int device;
    int sm_count;
    int64_t launches;
    void *encode_tiled;  // PFN_cuTensorMapEncodeTiled, resolved at create
};

// ---- error plumbing (api.cu) ------------------------------------------------
int ramnet_set_error(int code, const char *fmt, ...);
#define RAMNET_CHECK_ARG(cond, ...)                                      \
    do {                                                                  \
        if (!(cond)) return ramnet_set_error(RAMNET_EINVAL, __VA_ARGS__); \
    } while (0)
#define RAMNET_CUDA(expr)                                                                    \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return ramnet_set_error(RAMNET_ECUDA, "%s failed: %s (%s:%d)", #expr,            \
                                    cudaGetErrorString(e__), __FILE__, __LINE__);            \
    } while (0)
#define RAMNET_LAUNCH_CHECK(h)                                                               \
    do {                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess)                                                              \
            return ramnet_set_error(RAMNET_ECUDA, "kernel launch failed: %s (%s:%d)",        \
                                    cudaGetErrorString(e__), __FILE__, __LINE__);            \
        (h)->launches++;                                                                     \
    } while (0)

// ---- device selection -------------------------------------------------------------
// The reference selects its GPU only through tensor placement (config['gpu'] -> .to(self.gpu), no set_device), so the
// caller's current device can differ from the handle's.  Every entry point that takes a handle makes h->device current
// for the duration of the call and restores the previous device on return; cudaGetDevice / cudaSetDevice on the
// already-current device are host-side no-ops (~100 ns).
struct ramnet_device_guard {
    int prev = -1;
    bool switched = false;
    explicit ramnet_device_guard(int device) {
        if (device < 0 || cudaGetDevice(&prev) != cudaSuccess) return;
        if (prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    explicit ramnet_device_guard(const ramnet_handle *h) : ramnet_device_guard(h ? h->device : -1) {}
    ~ramnet_device_guard() {
        if (switched) cudaSetDevice(prev);
    }
};
#define RAMNET_DEVICE_GUARD(h) ramnet_device_guard ramnet_guard__(h)

// ---- programmatic dependent launch ----------------------------------------------
// Kernels of one pass run back to back on one stream.  A kernel launched through ramnet_launch(pdl = true) may start
// while its predecessor is still draining: its CTAs are placed as SMs free up and run their prologue (barrier init,
// TMEM allocation, tensor-map prefetch) before pdl_wait(), which returns once the predecessor grid has completed and
// its writes are visible.  Every global read or write of such a kernel sits after pdl_wait().  A kernel that never
// calls pdl_launch_dependents() triggers implicitly when it exits (no overlap, no hazard).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool ramnet_pdl_enabled();   // api.cu: RAMNET_PDL=1 turns it on (off by default: measured no gain inside CUDA graphs)

template <typename... KArgs, typename... Args>
inline cudaError_t ramnet_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                                 Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && ramnet_pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- device math --------------------------------------------------------------
// Accurate (not --use_fast_math) transcendental forms: parity with torch.sigmoid / tanh
// to ~1 ulp matters more here than the handful of SFU cycles.
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// FAST = the TF32 tensor-core kernels: ex2.approx / rcp.approx forms (abs error ~1e-7, two orders of magnitude below the
// TF32 operand rounding those kernels already carry); the FP32 strict-parity kernels keep the accurate forms.  The gate
// epilogues are issue- and LSU-bound (RAMNET_PROF: the MMA warp waits for the accumulator buffer on the full-resolution
// GRU layers), and expf + IEEE division cost ~25 instructions per element against ~6 for these.
template <bool FAST>
__device__ __forceinline__ float sigmoid_t(float x) {
    if constexpr (FAST) return __fdividef(1.0f, 1.0f + __expf(-x));
    else return sigmoidf_(x);
}
template <bool FAST>
__device__ __forceinline__ float tanh_t(float x) {
    if constexpr (FAST) return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x));   // e -> inf: 1; e -> 0: -1
    else return tanhf(x);
}

__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ---- fused epilogue -------------------------------------------------------------
struct EpiParams {
    const float *bias;  // [Cout] or nullptr
    const float *aux0;  // residual / h / c_prev
    const float *aux1;  // u
    float *y0;
    float *y1;
    float *y2;  // training stash: GRU_RU -> r, GRU_OUT -> o (candidate), LSTM -> gates [M][C][4]; may be nullptr
    int Cout;   // GEMM N
    int flags;  // RAMNET_FLAG_*
};

// The epilogue of NV consecutive GEMM columns [n0, n0+NV) of output pixel m, in two phases so that
// callers can software-pipeline it: `epilogue_prefetch` issues every global load the epilogue needs
// (bias, residual / h / u / c) into registers, `epilogue_finish` does the math and the stores.
// NV is a multiple of 4 and n0 a multiple of NV, so every access is a 16-byte vector access.
template <int NV>
struct EpiAux {
    float4 bias[NV / 4];
    float4 a[NV / 4];   // aux0: residual / h / c_prev (LSTM uses the first NV/4 scalars)
    float4 b[NV / 4];   // aux1: u
};

template <int EPI, int NV>
__device__ __forceinline__ void epilogue_prefetch(const EpiParams &p, int64_t m, int n0, EpiAux<NV> &x) {
    static_assert(NV % 4 == 0, "NV must be a multiple of 4");
#pragma unroll
    for (int j = 0; j < NV / 4; ++j)
        x.bias[j] = p.bias ? __ldg(reinterpret_cast<const float4 *>(p.bias + n0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (EPI == RAMNET_EPI_BIAS_RES_RELU || EPI == RAMNET_EPI_BIAS_RELU_ADD || EPI == RAMNET_EPI_BIAS_ADD) {
#pragma unroll
        for (int j = 0; j < NV / 4; ++j) x.a[j] = *(reinterpret_cast<const float4 *>(p.aux0 + m * p.Cout + n0) + j);
    } else if constexpr (EPI == RAMNET_EPI_GRU_RU) {
        const int C = p.Cout >> 1;
        if (n0 < C) {
#pragma unroll
            for (int j = 0; j < NV / 4; ++j) x.a[j] = *(reinterpret_cast<const float4 *>(p.aux0 + m * C + n0) + j);
        }
    } else if constexpr (EPI == RAMNET_EPI_GRU_OUT) {
#pragma unroll
        for (int j = 0; j < NV / 4; ++j) {
            x.a[j] = *(reinterpret_cast<const float4 *>(p.aux0 + m * p.Cout + n0) + j);
            x.b[j] = *(reinterpret_cast<const float4 *>(p.aux1 + m * p.Cout + n0) + j);
        }
    } else if constexpr (EPI == RAMNET_EPI_LSTM) {
        const int C = p.Cout >> 2;
        const float *cp = p.aux0 + m * C + (n0 >> 2);
        if constexpr (NV % 16 == 0) {
#pragma unroll
            for (int j = 0; j < NV / 16; ++j) x.a[j] = *(reinterpret_cast<const float4 *>(cp) + j);
        } else {
            float *dst = reinterpret_cast<float *>(&x.a[0]);
#pragma unroll
            for (int q = 0; q < NV / 4; ++q) dst[q] = cp[q];
        }
    }
}

// 256-bit store (sm_100: STG.E.256): one instruction per full 32-byte sector.  The epilogue's access pattern is one
// thread = one pixel row, so a warp store touches 32 different 128-byte lines whatever its width; halving the number of
// store instructions halves the LSU / L1 cycles the epilogue takes away from the shared-memory-bound mainloop.
__device__ __forceinline__ void st_global_v8(float *dst, const float (&a)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(a[0]), "f"(a[1]), "f"(a[2]),
                 "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7])
                 : "memory");
}

template <int EPI, int NV, bool FAST = false>
__device__ __forceinline__ void epilogue_finish(const EpiParams &p, int64_t m, int n0, float (&v)[NV],
                                                const EpiAux<NV> &x) {
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
        v[4 * j] += x.bias[j].x; v[4 * j + 1] += x.bias[j].y; v[4 * j + 2] += x.bias[j].z; v[4 * j + 3] += x.bias[j].w;
    }
    const bool rnd = (p.flags & RAMNET_FLAG_ROUND_TF32) != 0;
    auto st4 = [&](float *dst, float a, float b, float c, float d) {
        if (rnd) { a = round_tf32(a); b = round_tf32(b); c = round_tf32(c); d = round_tf32(d); }
        *reinterpret_cast<float4 *>(dst) = make_float4(a, b, c, d);
    };
    // NV consecutive outputs of one pixel: 32-byte stores when NV allows (rows are 64-byte aligned: Cout % 16 == 0)
    auto store_row = [&](float *dst, const float (&o)[NV]) {
        if constexpr (NV % 8 == 0) {
#pragma unroll
            for (int j = 0; j < NV; j += 8) {
                float t[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) t[q] = rnd ? round_tf32(o[j + q]) : o[j + q];
                st_global_v8(dst + j, t);
            }
        } else {
#pragma unroll
            for (int j = 0; j < NV; j += 4) st4(dst + j, o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
    };
    if constexpr (EPI == RAMNET_EPI_BIAS) {
        store_row(p.y0 + m * p.Cout + n0, v);
    } else if constexpr (EPI == RAMNET_EPI_BIAS_RELU) {
        float o[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) o[j] = fmaxf(v[j], 0.f);
        store_row(p.y0 + m * p.Cout + n0, o);
    } else if constexpr (EPI == RAMNET_EPI_BIAS_RES_RELU) {
        float o[NV];
#pragma unroll
        for (int j = 0; j < NV; j += 4) {
            const float4 r = x.a[j / 4];
            o[j] = fmaxf(v[j] + r.x, 0.f); o[j + 1] = fmaxf(v[j + 1] + r.y, 0.f);
            o[j + 2] = fmaxf(v[j + 2] + r.z, 0.f); o[j + 3] = fmaxf(v[j + 3] + r.w, 0.f);
        }
        store_row(p.y0 + m * p.Cout + n0, o);
    } else if constexpr (EPI == RAMNET_EPI_BIAS_ADD) {
        float o[NV];
#pragma unroll
        for (int j = 0; j < NV; j += 4) {
            const float4 r = x.a[j / 4];
            o[j] = v[j] + r.x; o[j + 1] = v[j + 1] + r.y; o[j + 2] = v[j + 2] + r.z; o[j + 3] = v[j + 3] + r.w;
        }
        store_row(p.y0 + m * p.Cout + n0, o);
    } else if constexpr (EPI == RAMNET_EPI_BIAS_RELU_ADD) {
        float o[NV];
#pragma unroll
        for (int j = 0; j < NV; j += 4) {
            const float4 r = x.a[j / 4];
            o[j] = fmaxf(v[j], 0.f) + r.x; o[j + 1] = fmaxf(v[j + 1], 0.f) + r.y;
            o[j + 2] = fmaxf(v[j + 2], 0.f) + r.z; o[j + 3] = fmaxf(v[j + 3], 0.f) + r.w;
        }
        store_row(p.y0 + m * p.Cout + n0, o);
    } else if constexpr (EPI == RAMNET_EPI_GRU_RU) {
        const int C = p.Cout >> 1;
        if (n0 < C) {  // reset gate -> y1 = h * r
#pragma unroll
            for (int j = 0; j < NV; j += 4) {
                const float4 h = x.a[j / 4];
                const float r0 = sigmoid_t<FAST>(v[j]), r1 = sigmoid_t<FAST>(v[j + 1]), r2 = sigmoid_t<FAST>(v[j + 2]), r3 = sigmoid_t<FAST>(v[j + 3]);
                if (p.y2) *reinterpret_cast<float4 *>(p.y2 + m * C + n0 + j) = make_float4(r0, r1, r2, r3);
                st4(p.y1 + m * C + n0 + j, h.x * r0, h.y * r1, h.z * r2, h.w * r3);
            }
        } else {  // update gate -> y0 = u   (kept full fp32: it is a pointwise operand only)
#pragma unroll
            for (int j = 0; j < NV; j += 4)
                *reinterpret_cast<float4 *>(p.y0 + m * C + (n0 - C) + j) =
                    make_float4(sigmoid_t<FAST>(v[j]), sigmoid_t<FAST>(v[j + 1]), sigmoid_t<FAST>(v[j + 2]), sigmoid_t<FAST>(v[j + 3]));
        }
    } else if constexpr (EPI == RAMNET_EPI_GRU_OUT) {
        float hn[NV];
#pragma unroll
        for (int j = 0; j < NV; j += 4) {
            const float4 h = x.a[j / 4], u = x.b[j / 4];
            const float o0 = tanh_t<FAST>(v[j]), o1 = tanh_t<FAST>(v[j + 1]), o2 = tanh_t<FAST>(v[j + 2]), o3 = tanh_t<FAST>(v[j + 3]);
            if (p.y2) *reinterpret_cast<float4 *>(p.y2 + m * p.Cout + n0 + j) = make_float4(o0, o1, o2, o3);
            hn[j] = h.x * (1.f - u.x) + o0 * u.x; hn[j + 1] = h.y * (1.f - u.y) + o1 * u.y;
            hn[j + 2] = h.z * (1.f - u.z) + o2 * u.z; hn[j + 3] = h.w * (1.f - u.w) + o3 * u.w;
        }
        store_row(p.y0 + m * p.Cout + n0, hn);
    } else if constexpr (EPI == RAMNET_EPI_LSTM) {
        const int C = p.Cout >> 2;
        const float *cprev = reinterpret_cast<const float *>(&x.a[0]);
        float hn[NV / 4], cn[NV / 4];
#pragma unroll
        for (int q = 0; q < NV / 4; ++q) {
            const float gi = sigmoid_t<FAST>(v[4 * q]), gf = sigmoid_t<FAST>(v[4 * q + 1]);
            const float go = sigmoid_t<FAST>(v[4 * q + 2]), gc = tanh_t<FAST>(v[4 * q + 3]);
            cn[q] = gf * cprev[q] + gi * gc;
            hn[q] = go * tanh_t<FAST>(cn[q]);
            if (rnd) hn[q] = round_tf32(hn[q]);
            if (p.y2) *reinterpret_cast<float4 *>(p.y2 + (m * C + (n0 >> 2) + q) * 4) = make_float4(gi, gf, go, gc);
        }
        if constexpr (NV % 16 == 0) {
#pragma unroll
            for (int q = 0; q < NV / 4; q += 4) {
                *reinterpret_cast<float4 *>(p.y0 + m * C + (n0 >> 2) + q) = make_float4(hn[q], hn[q + 1], hn[q + 2], hn[q + 3]);
                *reinterpret_cast<float4 *>(p.y1 + m * C + (n0 >> 2) + q) = make_float4(cn[q], cn[q + 1], cn[q + 2], cn[q + 3]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < NV / 4; ++q) {
                p.y0[m * C + (n0 >> 2) + q] = hn[q];
                p.y1[m * C + (n0 >> 2) + q] = cn[q];
            }
        }
    }
}

template <int EPI, int NV>
__device__ __forceinline__ void epilogue_store(const EpiParams &p, int64_t m, int n0, float (&v)[NV]) {
    EpiAux<NV> x;
    epilogue_prefetch<EPI, NV>(p, m, n0, x);
    epilogue_finish<EPI, NV>(p, m, n0, v, x);
}

static inline int conv_out_dim(int in, int stride) { return (in - 1) / stride + 1; }
static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

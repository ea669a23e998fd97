// Handle, error reporting and the small pointwise entry points of the C ABI.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

int ramnet_set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" int ramnet_version(void) { return RAMNET_ABI_VERSION; }
extern "C" const char *ramnet_last_error(void) { return g_err; }

extern "C" int ramnet_create(int device, ramnet_handle **out) {
    RAMNET_CHECK_ARG(out != nullptr, "ramnet_create: out is NULL");
    int count = 0;
    RAMNET_CUDA(cudaGetDeviceCount(&count));
    RAMNET_CHECK_ARG(device >= 0 && device < count, "ramnet_create: device %d out of range (%d visible)", device, count);
    cudaDeviceProp prop;
    RAMNET_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return ramnet_set_error(RAMNET_EDEVICE, "ramnet_create: device %d is sm_%d%d; this library is sm_100a only",
                                device, prop.major, prop.minor);
    // the handle is bound to `device`, but the caller's current device is left as it was: every entry point switches
    // to h->device for the duration of the call (RAMNET_DEVICE_GUARD) and restores the previous one
    ramnet_device_guard guard(device);
    ramnet_handle *h = new ramnet_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->launches = 0;
    h->encode_tiled = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &h->encode_tiled, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || h->encode_tiled == nullptr) {
        delete h;
        return ramnet_set_error(RAMNET_ECUDA, "ramnet_create: cuTensorMapEncodeTiled entry point unavailable");
    }
    *out = h;
    return RAMNET_OK;
}

bool ramnet_pdl_enabled() {
    static const int on = [] {
        const char *e = getenv("RAMNET_PDL");
        return (e && e[0] == '1') ? 1 : 0;   // measured on B200 inside CUDA graphs: no gain (19.81 vs 19.55 ms/step), off by default
    }();
    return on != 0;
}

extern "C" int ramnet_destroy(ramnet_handle *h) {
    delete h;
    return RAMNET_OK;
}

extern "C" int ramnet_sm_count(const ramnet_handle *h) { return h ? h->sm_count : 0; }
extern "C" int64_t ramnet_launch_count(const ramnet_handle *h) { return h ? h->launches : 0; }

// ---------------------------------------------------------------------------------
// decoder prologue: (x + skip) -> bilinear x2, align_corners=False, NHWC
//   out[2m]   = .25 in[m-1] + .75 in[m];  out[2m+1] = .75 in[m] + .25 in[m+1], indices clamped
// HBM-bound: reads C*4 B per input pixel (+ skip), writes 4x that.  One thread owns 4 channels
// of one input row and slides a 3x3 window along x: 3 (6 with skip) 16-byte loads and 4 16-byte
// stores per input pixel instead of 4-8 loads per OUTPUT pixel; consecutive threads are
// consecutive channel quads, so every warp access is a contiguous run of the NHWC pixel.
// Association matches ATen's upsample_bilinear2d: w_y0*(w_x0*a + w_x1*b) + w_y1*(w_x0*c + w_x1*d).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float4 f4_lerp(float wa, const float4 &a, float wb, const float4 &b) {
    return make_float4(wa * a.x + wb * b.x, wa * a.y + wb * b.y, wa * a.z + wb * b.z, wa * a.w + wb * b.w);
}

constexpr int kUpStrip = 8;   // input pixels per thread along x

__global__ void __launch_bounds__(256) upsample2x_add_kernel(const float4 *__restrict__ x,
                                                             const float4 *__restrict__ skip,
                                                             float4 *__restrict__ y, int N, int H, int W,
                                                             int C4, int round) {
    const int strips = (W + kUpStrip - 1) / kUpStrip;
    const int64_t total = (int64_t)N * H * strips * C4;
    pdl_wait();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        int64_t p = i / C4;
        const int xs = (int)(p % strips);
        p /= strips;
        const int yi = (int)(p % H);
        const int n = (int)(p / H);
        const int64_t img = (int64_t)n * H * W;
        const int rows[3] = {max(yi - 1, 0), yi, min(yi + 1, H - 1)};
        auto ld = [&](int r, int xx) {
            const int64_t o = (img + (int64_t)rows[r] * W + xx) * C4 + c;
            float4 a = __ldg(x + o);
            if (skip != nullptr) {
                const float4 s = __ldg(skip + o);
                a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
            }
            return a;
        };
        const int xbeg = xs * kUpStrip, xend = min(xbeg + kUpStrip, W);
        float4 prev[3], cur[3], nxt[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            prev[r] = ld(r, max(xbeg - 1, 0));
            cur[r] = ld(r, xbeg);
        }
        float4 *out0 = y + (((int64_t)n * 2 * H + 2 * yi) * 2 * W) * C4 + c;   // output row 2*yi
        float4 *out1 = out0 + (int64_t)2 * W * C4;                              // output row 2*yi + 1
        for (int xi = xbeg; xi < xend; ++xi) {
#pragma unroll
            for (int r = 0; r < 3; ++r) nxt[r] = ld(r, min(xi + 1, W - 1));
            float4 he[3], ho[3];   // horizontal blends: even / odd output column of each of the 3 input rows
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                he[r] = f4_lerp(0.25f, prev[r], 0.75f, cur[r]);
                ho[r] = f4_lerp(0.75f, cur[r], 0.25f, nxt[r]);
            }
            float4 o00 = f4_lerp(0.25f, he[0], 0.75f, he[1]), o01 = f4_lerp(0.25f, ho[0], 0.75f, ho[1]);
            float4 o10 = f4_lerp(0.75f, he[1], 0.25f, he[2]), o11 = f4_lerp(0.75f, ho[1], 0.25f, ho[2]);
            if (round) {
                o00 = make_float4(round_tf32(o00.x), round_tf32(o00.y), round_tf32(o00.z), round_tf32(o00.w));
                o01 = make_float4(round_tf32(o01.x), round_tf32(o01.y), round_tf32(o01.z), round_tf32(o01.w));
                o10 = make_float4(round_tf32(o10.x), round_tf32(o10.y), round_tf32(o10.z), round_tf32(o10.w));
                o11 = make_float4(round_tf32(o11.x), round_tf32(o11.y), round_tf32(o11.z), round_tf32(o11.w));
            }
            out0[(int64_t)(2 * xi) * C4] = o00;
            out0[(int64_t)(2 * xi + 1) * C4] = o01;
            out1[(int64_t)(2 * xi) * C4] = o10;
            out1[(int64_t)(2 * xi + 1) * C4] = o11;
#pragma unroll
            for (int r = 0; r < 3; ++r) { prev[r] = cur[r]; cur[r] = nxt[r]; }
        }
    }
}

extern "C" int ramnet_upsample2x_add(ramnet_handle *h, const float *x, const float *skip, float *y, int N, int H,
                                     int W, int C, int flags, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && x && y, "ramnet_upsample2x_add: NULL argument");
    RAMNET_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "ramnet_upsample2x_add: bad shape N=%d H=%d W=%d C=%d (C%%4)", N, H, W, C);
    const int64_t total = (int64_t)N * H * ((W + kUpStrip - 1) / kUpStrip) * (C / 4);
    const int blocks = (int)imin64((total + 255) / 256, (int64_t)h->sm_count * 16);
    RAMNET_CUDA(ramnet_launch(upsample2x_add_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, true,
                              (const float4 *)x, (const float4 *)skip, (float4 *)y, N, H, W, C / 4,
                              (flags & RAMNET_FLAG_ROUND_TF32) != 0));
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

// ---------------------------------------------------------------------------------
// prediction head: logits[m] = sum_c (x[m,c] (+ skip[m,c])) * w[c] + b ; depth = sigmoid(logits)
// 8 lanes per pixel, float4 per lane per iteration -> fully coalesced 128 B rows for C=32.
// ---------------------------------------------------------------------------------
// w_skip != nullptr: skip_type 'concat' (unet.py:11-12,129): logits = x . w + skip . w_skip instead of (x + skip) . w
__global__ void __launch_bounds__(256) pred_sigmoid_kernel(const float *__restrict__ x, const float *__restrict__ skip,
                                                           const float *__restrict__ w, const float *__restrict__ w_skip,
                                                           const float *__restrict__ bias,
                                                           float *__restrict__ logits, float *__restrict__ depth,
                                                           int64_t M, int C) {
    const int lane8 = threadIdx.x & 7;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 3;
    const float b = bias ? __ldg(bias) : 0.f;
    for (int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3; m < ((M + stride - 1) / stride) * stride;
         m += stride) {
        float acc = 0.f;
        if (m < M) {
            for (int c = lane8 * 4; c < C; c += 32) {
                float4 a = *reinterpret_cast<const float4 *>(x + m * C + c);
                if (skip) {
                    float4 s = *reinterpret_cast<const float4 *>(skip + m * C + c);
                    if (w_skip) {
                        const float4 ws = __ldg(reinterpret_cast<const float4 *>(w_skip + c));
                        acc += s.x * ws.x + s.y * ws.y + s.z * ws.z + s.w * ws.w;
                    } else {
                        a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
                    }
                }
                const float4 ww = __ldg(reinterpret_cast<const float4 *>(w + c));
                acc += a.x * ww.x + a.y * ww.y + a.z * ww.z + a.w * ww.w;
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (m < M && lane8 == 0) {
            acc += b;
            if (logits) logits[m] = acc;
            if (depth) depth[m] = sigmoidf_(acc);
        }
    }
}

extern "C" int ramnet_pred_sigmoid(ramnet_handle *h, const float *x, const float *skip, const float *w,
                                   const float *w_skip, const float *bias, float *logits, float *depth, int64_t M, int C,
                                   void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && x && w && (logits || depth) && (!w_skip || skip), "ramnet_pred_sigmoid: NULL argument");
    RAMNET_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0, "ramnet_pred_sigmoid: bad shape M=%lld C=%d (C%%4)", (long long)M, C);
    const int blocks = (int)imin64((M * 8 + 255) / 256, (int64_t)h->sm_count * 16);
    pred_sigmoid_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, skip, w, w_skip, bias, logits, depth, M, C);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

// ---------------------------------------------------------------------------------
// layout helpers (tests + generic-Cin inputs).  Tiled 32x32 transpose through smem.
// ---------------------------------------------------------------------------------
__global__ void transpose_kernel(const float *__restrict__ in, float *__restrict__ out, int R, int Ccols, int round) {
    // in: [batch][R][Ccols] -> out: [batch][Ccols][R]
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const float *src = in + (int64_t)b * R * Ccols;
    float *dst = out + (int64_t)b * R * Ccols;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < R && c < Ccols) tile[j][threadIdx.x] = src[(int64_t)r * Ccols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < R && c < Ccols) {
            float v = tile[threadIdx.x][j];
            dst[(int64_t)c * R + r] = round ? round_tf32(v) : v;
        }
    }
}

static int launch_transpose(ramnet_handle *h, const float *x, float *y, int batch, int R, int Ccols, int round,
                            void *stream) {
    dim3 grid((Ccols + 31) / 32, (R + 31) / 32, batch), block(32, 8);
    RAMNET_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "transpose: shape too large for the grid");
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, y, R, Ccols, round);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_nchw_to_nhwc(ramnet_handle *h, const float *x, float *y, int N, int C, int H, int W, int flags,
                                   void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && x && y && N > 0 && C > 0 && H > 0 && W > 0, "ramnet_nchw_to_nhwc: bad argument");
    return launch_transpose(h, x, y, N, C, H * W, (flags & RAMNET_FLAG_ROUND_TF32) != 0, stream);
}

extern "C" int ramnet_nhwc_to_nchw(ramnet_handle *h, const float *x, float *y, int N, int C, int H, int W,
                                   void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && x && y && N > 0 && C > 0 && H > 0 && W > 0, "ramnet_nhwc_to_nchw: bad argument");
    return launch_transpose(h, x, y, N, H * W, C, 0, stream);
}

__global__ void round_tf32_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = round_tf32(x[i]);
}

extern "C" int ramnet_round_tf32(ramnet_handle *h, const float *x, float *y, int64_t n, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && x && y && n >= 0, "ramnet_round_tf32: bad argument");
    if (n == 0) return RAMNET_OK;
    const int blocks = (int)imin64((n + 255) / 256, (int64_t)h->sm_count * 16);
    round_tf32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, n);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

// ---------------------------------------------------------------------------------
// weight packing: [Cout, Cin, k, k] -> FP32: [tap][Cin][Cout] ; TF32: [tap][Cout][Cin] (rna-rounded)
// ---------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float *__restrict__ w, float *__restrict__ out, int Cout, int Cin, int taps,
                                    int kind, int interleave) {
    const int64_t total = (int64_t)Cout * Cin * taps;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int tap = (int)(i % taps);
        const int ci = (int)((i / taps) % Cin);
        const int co = (int)(i / ((int64_t)taps * Cin));
        int col = co;
        if (interleave) {  // reference chunk order: gate g occupies rows [g*C, (g+1)*C) -> column 4c+g
            const int C = Cout >> 2;
            col = (co % C) * 4 + co / C;
        }
        const float v = w[i];
        if (kind == RAMNET_MMA_FP32)
            out[((int64_t)tap * Cin + ci) * Cout + col] = v;
        else
            out[((int64_t)tap * Cout + col) * Cin + ci] = round_tf32(v);
    }
}

extern "C" int ramnet_pack_weights(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin, int ksize,
                                   int mma_kind, int lstm_interleave, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && w_oihw && w_packed, "ramnet_pack_weights: NULL argument");
    RAMNET_CHECK_ARG(Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3 || ksize == 5), "ramnet_pack_weights: bad shape");
    RAMNET_CHECK_ARG(mma_kind == RAMNET_MMA_FP32 || mma_kind == RAMNET_MMA_TF32, "ramnet_pack_weights: bad mma_kind");
    RAMNET_CHECK_ARG(!lstm_interleave || Cout % 4 == 0, "ramnet_pack_weights: lstm_interleave needs Cout%%4==0");
    const int64_t total = (int64_t)Cout * Cin * ksize * ksize;
    const int blocks = (int)imin64((total + 255) / 256, (int64_t)h->sm_count * 8);
    pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, w_packed, Cout, Cin, ksize * ksize, mma_kind,
                                                                  lstm_interleave);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

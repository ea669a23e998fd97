// Head convolution: 5x5 s1 p2 conv + bias + ReLU on the raw NCHW network
// input (Cin = 1, 5 or 6 <= 8), writing the pixel-major NHWC tensor the tensor-core layers read.
//
// Replaces ConvLayer.forward for head_events/head_rgb (RAM_Net/model/statenet.py:139-145,
// submodules.py:26-35) and unet.head (unet.py:93-94).
//
// K = 25*Cin (25..150) is too ragged for a TMA-fed UMMA tile without padding and the layer is
// 0.2-1.0 GFLOP per map writing 16.8 MB, i.e. 12-54 FLOP/B: it is bound by the fp32 FFMA pipe and
// the NHWC store, so it is a direct convolution:
//   - CTA = 256 threads = 8 warps, output tile 8 rows x 32 columns, one pixel per thread, all
//     Cout (<= 32 per pass) accumulators in registers;
//   - input halo tile (12 x 36 x Cin) and the whole weight tensor staged in shared memory once
//     per CTA; weights are read as warp-broadcast LDS.128;
//   - results transposed through an XOR-swizzled per-warp staging buffer so that every store
//     instruction of a warp writes 512 contiguous bytes of the NHWC row.
#include "common.cuh"

namespace {
constexpr int TW = 32, TH = 8, HALO = 2, IW = TW + 2 * HALO, IH = TH + 2 * HALO;
constexpr int COB = 32;            // output channels per pass
constexpr int MAX_CIN = 8;
constexpr int SMEM_FLOATS = IH * IW * MAX_CIN + MAX_CIN * 25 * COB;  // 9856 floats = 38.5 KB

__global__ void __launch_bounds__(256) head_conv_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                        const float *__restrict__ bias, float *__restrict__ y, int N,
                                                        int Cin, int H, int W, int Cout, int co_base, int round) {
    __shared__ __align__(16) float smem[SMEM_FLOATS];
    float *in_s = smem;                       // [Cin][IH][IW]
    float *w_s = smem + IH * IW * MAX_CIN;    // [Cin*25][COB]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_wait();
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, n = blockIdx.z;

    for (int i = tid; i < Cin * IH * IW; i += 256) {
        const int c = i / (IH * IW), r = (i / IW) % IH, s = i % IW;
        const int gy = y0 + r - HALO, gx = x0 + s - HALO;
        float v = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(x + (((int64_t)n * Cin + c) * H + gy) * W + gx);
        in_s[i] = v;
    }
    for (int i = tid; i < Cin * 25 * COB; i += 256) {
        const int co = i % COB, k = i / COB;  // k = c*25 + tap ; w is [Cout][Cin][5][5]
        const int gco = co_base + co;
        w_s[i] = gco < Cout ? __ldg(w + (int64_t)gco * Cin * 25 + k) : 0.f;
    }
    __syncthreads();

    float acc[COB];
#pragma unroll
    for (int j = 0; j < COB; ++j) acc[j] = 0.f;
    for (int c = 0; c < Cin; ++c) {
        const float *ip = in_s + (c * IH + warp) * IW + lane;
        const float4 *wp = reinterpret_cast<const float4 *>(w_s + c * 25 * COB);
#pragma unroll
        for (int r = 0; r < 5; ++r) {
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const float v = ip[r * IW + s];
#pragma unroll
                for (int q = 0; q < COB / 4; ++q) {
                    const float4 ww = wp[(r * 5 + s) * (COB / 4) + q];
                    acc[4 * q] = fmaf(v, ww.x, acc[4 * q]);
                    acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
                }
            }
        }
    }
    __syncthreads();  // all reads of in_s / w_s done: reuse the region as the store staging buffer

    float4 *stage = reinterpret_cast<float4 *>(smem) + warp * (32 * COB / 4);  // [32 px][8 quads], swizzled
#pragma unroll
    for (int q = 0; q < COB / 4; ++q) {
        float4 v;
        const int co = co_base + 4 * q;
        v.x = acc[4 * q] + (bias && co < Cout ? __ldg(bias + co) : 0.f);
        v.y = acc[4 * q + 1] + (bias && co + 1 < Cout ? __ldg(bias + co + 1) : 0.f);
        v.z = acc[4 * q + 2] + (bias && co + 2 < Cout ? __ldg(bias + co + 2) : 0.f);
        v.w = acc[4 * q + 3] + (bias && co + 3 < Cout ? __ldg(bias + co + 3) : 0.f);
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        if (round) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
        stage[lane * (COB / 4) + (q ^ (lane & 7))] = v;
    }
    __syncwarp();
    const int oy = y0 + warp;
    if (oy < H) {
        float *yrow = y + (((int64_t)n * H + oy) * W) * Cout;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int px = j * 4 + (lane >> 3), q = lane & 7;
            const int ox = x0 + px, co = co_base + 4 * q;
            if (ox < W && co < Cout) {  // Cout % 4 == 0 is checked on the host
                const float4 v = stage[px * (COB / 4) + (q ^ (px & 7))];
                *reinterpret_cast<float4 *>(yrow + (int64_t)ox * Cout + co) = v;
            }
        }
    }
}
}  // namespace

extern "C" int ramnet_head_conv(ramnet_handle *h, const float *x_nchw, const float *w_oihw, const float *bias,
                                float *y_nhwc, int N, int Cin, int H, int W, int Cout, int flags, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && x_nchw && w_oihw && y_nhwc, "ramnet_head_conv: NULL argument");
    RAMNET_CHECK_ARG(N > 0 && H > 0 && W > 0, "ramnet_head_conv: bad shape N=%d H=%d W=%d", N, H, W);
    RAMNET_CHECK_ARG(Cin >= 1 && Cin <= MAX_CIN, "ramnet_head_conv: Cin=%d not in [1,%d]", Cin, MAX_CIN);
    RAMNET_CHECK_ARG(Cout > 0 && Cout % 4 == 0, "ramnet_head_conv: Cout=%d must be a positive multiple of 4", Cout);
    RAMNET_CHECK_ARG(N <= 65535 && (H + TH - 1) / TH <= 65535, "ramnet_head_conv: grid too large");
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, N);
    for (int co_base = 0; co_base < Cout; co_base += COB) {
        RAMNET_CUDA(ramnet_launch(head_conv_kernel, grid, dim3(256), 0, (cudaStream_t)stream, true, x_nchw, w_oihw, bias,
                                  y_nhwc, N, Cin, H, W, Cout, co_base, (flags & RAMNET_FLAG_ROUND_TF32) != 0));
        RAMNET_LAUNCH_CHECK(h);
    }
    return RAMNET_OK;
}

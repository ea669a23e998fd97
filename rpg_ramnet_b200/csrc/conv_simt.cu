// fp32-exact implicit-GEMM convolution on the CUDA cores (RAMNET_MMA_FP32).
//
// The strict-parity arithmetic mode of ramnet_conv_fwd: same data layout (NHWC activations,
// virtual [x0|x1] channel concat, fused epilogues) as the tcgen05 path, but fp32 FFMA with fp32
// accumulation so results match the reference's fp32 cuDNN/MKL-DNN convolutions to rounding
// order.  Used by the parity tests to separate "wrong" from "TF32-rounded", and selectable by
// users through mma_kind.  Classic 64 x BN x 16 smem-tiled SGEMM, register-prefetched.
//
// Reference call sites replaced: see ramnet_conv_fwd in include/ramnet_b200.h.
#include "common.cuh"

namespace {
constexpr int BM = 64, BK = 16, ALD = BK + 4;

struct ConvGeom {
    int N, H, W, C0, C1, Cout, ks, stride, pad, Ho, Wo;
    int64_t M;
};

template <int EPI, int BN>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvGeom g, const float *__restrict__ x0,
                                                        const float *__restrict__ x1, const float *__restrict__ wp,
                                                        EpiParams ep) {
    constexpr int TM = BM * BN / (256 * 4);  // pixels per thread (4 for BN=64, 2 for BN=32)
    constexpr int TXN = BN / 4;              // threads along N
    __shared__ __align__(16) float As[BM * ALD];
    __shared__ __align__(16) float Bs[BK * BN];
    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A loader: pixel a_px, channel quad a_q of the current 16-channel chunk
    const int a_px = tid >> 2, a_q = tid & 3;
    const int64_t am = m0 + a_px;
    const bool a_ok = am < g.M;
    int an = 0, aoy = 0, aox = 0;
    if (a_ok) {
        aox = (int)(am % g.Wo);
        aoy = (int)((am / g.Wo) % g.Ho);
        an = (int)(am / ((int64_t)g.Wo * g.Ho));
    }
    // B loader: k row b_k, column quad b_q  (only the first BK*BN/4 threads load when BN=32)
    const int b_k = tid / TXN, b_q = tid % TXN;
    const bool b_thread = tid < BK * TXN;

    const int Ct = g.C0 + g.C1;
    const int chunks = Ct / BK;
    const int steps = g.ks * g.ks * chunks;

    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 ra = make_float4(0, 0, 0, 0), rb = make_float4(0, 0, 0, 0);
    auto fetch = [&](int step) {
        const int tap = step / chunks, c = (step % chunks) * BK;
        const int r = tap / g.ks, s = tap % g.ks;
        ra = make_float4(0, 0, 0, 0);
        if (a_ok) {
            const int iy = aoy * g.stride + r - g.pad, ix = aox * g.stride + s - g.pad;
            if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
                const int64_t pix = ((int64_t)an * g.H + iy) * g.W + ix;
                const float *src = (c < g.C0) ? x0 + pix * g.C0 + c : x1 + pix * g.C1 + (c - g.C0);
                ra = *reinterpret_cast<const float4 *>(src + a_q * 4);
            }
        }
        rb = make_float4(0, 0, 0, 0);
        if (b_thread) {
            const int col = n0 + b_q * 4;
            if (col < g.Cout)  // packed [tap][Cin][Cout]
                rb = __ldg(reinterpret_cast<const float4 *>(wp + ((int64_t)tap * Ct + c + b_k) * g.Cout + col));
        }
    };

    fetch(0);
    for (int step = 0; step < steps; ++step) {
        *reinterpret_cast<float4 *>(&As[a_px * ALD + a_q * 4]) = ra;
        if (b_thread) *reinterpret_cast<float4 *>(&Bs[b_k * BN + b_q * 4]) = rb;
        __syncthreads();
        if (step + 1 < steps) fetch(step + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k * BN + tx * 4]);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const float a = As[(ty * TM + i) * ALD + k];
                acc[i][0] = fmaf(a, b.x, acc[i][0]);
                acc[i][1] = fmaf(a, b.y, acc[i][1]);
                acc[i][2] = fmaf(a, b.z, acc[i][2]);
                acc[i][3] = fmaf(a, b.w, acc[i][3]);
            }
        }
        __syncthreads();
    }

    const int col = n0 + tx * 4;
    if (col < g.Cout) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int64_t m = m0 + ty * TM + i;
            if (m < g.M) epilogue_store<EPI, 4>(ep, m, col, acc[i]);
        }
    }
}

template <int EPI>
int launch_simt(ramnet_handle *h, const ConvGeom &g, const float *x0, const float *x1, const float *wp,
                const EpiParams &ep, cudaStream_t s) {
    const int64_t mt = (g.M + BM - 1) / BM;
    RAMNET_CHECK_ARG(mt <= 0x7fffffff, "conv_fwd: too many pixel tiles");
    if (g.Cout <= 32) {
        dim3 grid((unsigned)mt, (g.Cout + 31) / 32);
        conv_simt_kernel<EPI, 32><<<grid, 256, 0, s>>>(g, x0, x1, wp, ep);
    } else {
        dim3 grid((unsigned)mt, (g.Cout + 63) / 64);
        conv_simt_kernel<EPI, 64><<<grid, 256, 0, s>>>(g, x0, x1, wp, ep);
    }
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}
}  // namespace

int conv_fwd_tf32(ramnet_handle *h, const ramnet_conv_desc *d, const float *x0, const float *x1, const float *wp,
                  const EpiParams &ep, void *workspace, size_t ws_bytes, cudaStream_t s);
size_t conv_tf32_workspace_bytes(const ramnet_conv_desc *d);

static int validate_desc(const ramnet_conv_desc *d) {
    RAMNET_CHECK_ARG(d != nullptr, "conv: desc is NULL");
    RAMNET_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0, "conv: bad N/H/W = %d/%d/%d", d->N, d->H, d->W);
    RAMNET_CHECK_ARG(d->ksize == 1 || d->ksize == 3 || d->ksize == 5, "conv: ksize %d not in {1,3,5}", d->ksize);
    RAMNET_CHECK_ARG(d->stride == 1 || d->stride == 2, "conv: stride %d not in {1,2}", d->stride);
    RAMNET_CHECK_ARG(d->C0 > 0 && d->C0 % 16 == 0 && d->C1 >= 0 && d->C1 % 16 == 0,
                     "conv: channel counts C0=%d C1=%d must be multiples of 16", d->C0, d->C1);
    RAMNET_CHECK_ARG(d->Cout > 0 && d->Cout % 4 == 0, "conv: Cout=%d must be a positive multiple of 4", d->Cout);
    RAMNET_CHECK_ARG(d->epilogue >= RAMNET_EPI_BIAS && d->epilogue <= RAMNET_EPI_BIAS_ADD, "conv: bad epilogue %d", d->epilogue);
    RAMNET_CHECK_ARG(d->mma_kind == RAMNET_MMA_FP32 || d->mma_kind == RAMNET_MMA_TF32, "conv: bad mma_kind %d", d->mma_kind);
    RAMNET_CHECK_ARG(!(d->flags & RAMNET_FLAG_HPACK) || d->mma_kind == RAMNET_MMA_TF32, "conv: RAMNET_FLAG_HPACK needs mma_kind=TF32");
    RAMNET_CHECK_ARG(!(d->flags & RAMNET_FLAG_UPCONV) || d->mma_kind == RAMNET_MMA_TF32, "conv: RAMNET_FLAG_UPCONV needs mma_kind=TF32");
    if (d->epilogue == RAMNET_EPI_GRU_RU) RAMNET_CHECK_ARG(d->Cout % 8 == 0, "conv: GRU_RU needs Cout = 2C with C%%4 == 0");
    if (d->epilogue == RAMNET_EPI_LSTM) RAMNET_CHECK_ARG(d->Cout % 16 == 0, "conv: LSTM needs Cout = 4C with C%%4 == 0");
    return RAMNET_OK;
}

extern "C" size_t ramnet_conv_workspace_bytes(const ramnet_conv_desc *d) {
    if (d == nullptr || d->mma_kind != RAMNET_MMA_TF32) return 0;
    return conv_tf32_workspace_bytes(d);
}

extern "C" int ramnet_conv_fwd(ramnet_handle *h, const ramnet_conv_desc *d, const float *x0, const float *x1,
                               const float *w_packed, const float *bias, const float *aux0, const float *aux1,
                               float *y0, float *y1, float *y2, void *workspace, size_t workspace_bytes,
                               void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h != nullptr, "conv_fwd: handle is NULL");
    int rc = validate_desc(d);
    if (rc) return rc;
    RAMNET_CHECK_ARG(x0 && w_packed && y0, "conv_fwd: x0 / w_packed / y0 must not be NULL");
    RAMNET_CHECK_ARG((d->C1 == 0) == (x1 == nullptr), "conv_fwd: x1 and C1 disagree");
    switch (d->epilogue) {
        case RAMNET_EPI_BIAS_RES_RELU: RAMNET_CHECK_ARG(aux0, "conv_fwd: residual epilogue needs aux0"); break;
        case RAMNET_EPI_BIAS_ADD:
        case RAMNET_EPI_BIAS_RELU_ADD:
            RAMNET_CHECK_ARG(aux0, "conv_fwd: add epilogues need aux0");
            if (d->mma_kind != RAMNET_MMA_TF32)
                return ramnet_set_error(RAMNET_EUNSUPPORTED, "conv_fwd: the relu+add epilogue is implemented on the TF32 path only");
            break;
        case RAMNET_EPI_GRU_RU: RAMNET_CHECK_ARG(aux0 && y1, "conv_fwd: GRU_RU needs aux0 (h) and y1"); break;
        case RAMNET_EPI_GRU_OUT: RAMNET_CHECK_ARG(aux0 && aux1, "conv_fwd: GRU_OUT needs aux0 (h) and aux1 (u)"); break;
        case RAMNET_EPI_LSTM: RAMNET_CHECK_ARG(aux0 && y1, "conv_fwd: LSTM needs aux0 (c) and y1"); break;
        case RAMNET_EPI_BIAS_RELU_PRED:
            RAMNET_CHECK_ARG(aux0 && aux1, "conv_fwd: PRED needs aux0 (pred weight) and aux1 (pred bias)");
            if (d->mma_kind != RAMNET_MMA_TF32 || d->stride != 1 || d->ksize == 1 || d->Cout % 32 || d->Cout > 256)
                return ramnet_set_error(RAMNET_EUNSUPPORTED, "conv_fwd: the fused prediction epilogue needs mma_kind=TF32, "
                                        "stride 1, ksize 3/5 and Cout %% 32 == 0, Cout <= 256");
            break;
        default: break;
    }
    EpiParams ep{bias, aux0, aux1, y0, y1, y2, d->Cout, d->flags};
    cudaStream_t s = (cudaStream_t)stream;
    if (d->mma_kind == RAMNET_MMA_TF32) return conv_fwd_tf32(h, d, x0, x1, w_packed, ep, workspace, workspace_bytes, s);

    ConvGeom g;
    g.N = d->N; g.H = d->H; g.W = d->W; g.C0 = d->C0; g.C1 = d->C1; g.Cout = d->Cout;
    g.ks = d->ksize; g.stride = d->stride; g.pad = d->ksize / 2;
    g.Ho = conv_out_dim(d->H, d->stride); g.Wo = conv_out_dim(d->W, d->stride);
    g.M = (int64_t)g.N * g.Ho * g.Wo;
    switch (d->epilogue) {
        case RAMNET_EPI_BIAS: return launch_simt<RAMNET_EPI_BIAS>(h, g, x0, x1, w_packed, ep, s);
        case RAMNET_EPI_BIAS_RELU: return launch_simt<RAMNET_EPI_BIAS_RELU>(h, g, x0, x1, w_packed, ep, s);
        case RAMNET_EPI_BIAS_RES_RELU: return launch_simt<RAMNET_EPI_BIAS_RES_RELU>(h, g, x0, x1, w_packed, ep, s);
        case RAMNET_EPI_GRU_RU: return launch_simt<RAMNET_EPI_GRU_RU>(h, g, x0, x1, w_packed, ep, s);
        case RAMNET_EPI_GRU_OUT: return launch_simt<RAMNET_EPI_GRU_OUT>(h, g, x0, x1, w_packed, ep, s);
        case RAMNET_EPI_LSTM: return launch_simt<RAMNET_EPI_LSTM>(h, g, x0, x1, w_packed, ep, s);
    }
    return ramnet_set_error(RAMNET_EINVAL, "conv_fwd: unreachable");
}

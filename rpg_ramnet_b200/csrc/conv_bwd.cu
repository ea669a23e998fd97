// Backward pass building blocks: what autograd's cuDNN/ATen backward kernels do
// for the reference (loss.backward() at trainer/lstm_trainer.py:450), as explicit kernels.
//
//  * data gradient of a convolution = the forward implicit-GEMM kernel (tcgen05 halo kernel) run on
//    dZ with tap-flipped, channel-transposed weights (ramnet_pack_weights_dgrad); a stride-2 conv's
//    data gradient first zero-inserts dZ to the input resolution (ramnet_zero_insert2x);
//  * weight gradient dW[co][ci][r][s] = sum_pixels dZ[p][co] * X[p*stride + (r,s) - pad][ci] is a GEMM
//    whose K dimension is the pixel axis: first version = fp32 FFMA, 64x64 tiles, split-K over pixel
//    ranges with fp32 atomics, accumulating straight into the nn.Conv2d-layout .grad buffer (so
//    BPTT's sum over timesteps costs nothing extra);
//  * the pointwise adjoints of the fused epilogues (ReLU mask, GRU / LSTM gates, sigmoid head,
//    bilinear x2 + skip).
// All tensors NHWC fp32 as in the forward pass.
#include "common.cuh"

namespace {
constexpr int WB = 64, WK = 16;   // wgrad tile: 64 cout x 64 cin, 16 pixels per K step

struct WgradGeom {
    int N, H, W, C0, C1, Cout, ks, stride, pad, Ho, Wo;
    int64_t M;            // output pixels
    int pix_per_block;    // split-K range
};

__global__ void __launch_bounds__(256) conv_wgrad_kernel(WgradGeom g, const float *__restrict__ dz,
                                                         const float *__restrict__ x0, const float *__restrict__ x1,
                                                         float *__restrict__ dw) {
    __shared__ __align__(16) float As[WK * WB];   // [k = pixel][co]
    __shared__ __align__(16) float Bs[WK * WB];   // [k = pixel][ci]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int Ct = g.C0 + g.C1;
    const int ci_tiles = (Ct + WB - 1) / WB;
    const int tap = blockIdx.y / ci_tiles, ci0 = (blockIdx.y % ci_tiles) * WB;
    const int co0 = blockIdx.z * WB;
    const int r = tap / g.ks, s = tap % g.ks;
    const int64_t p_begin = (int64_t)blockIdx.x * g.pix_per_block;
    const int64_t p_end = p_begin + g.pix_per_block < g.M ? p_begin + g.pix_per_block : g.M;
    // loader mapping: 16 pixels x 16 channel quads
    const int lp = tid >> 4, lq = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t p0 = p_begin; p0 < p_end; p0 += WK) {
        const int64_t p = p0 + lp;
        float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
        if (p < p_end) {
            const int co = co0 + lq * 4;
            if (co < g.Cout) a = *reinterpret_cast<const float4 *>(dz + p * g.Cout + co);
            const int ox = (int)(p % g.Wo), oy = (int)((p / g.Wo) % g.Ho), n = (int)(p / ((int64_t)g.Wo * g.Ho));
            const int iy = oy * g.stride + r - g.pad, ix = ox * g.stride + s - g.pad;
            const int ci = ci0 + lq * 4;
            if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W && ci < Ct) {
                const int64_t ip = ((int64_t)n * g.H + iy) * g.W + ix;
                b = (ci < g.C0) ? *reinterpret_cast<const float4 *>(x0 + ip * g.C0 + ci)
                                : *reinterpret_cast<const float4 *>(x1 + ip * g.C1 + (ci - g.C0));
            }
        }
        __syncthreads();
        *reinterpret_cast<float4 *>(&As[lp * WB + lq * 4]) = a;
        *reinterpret_cast<float4 *>(&Bs[lp * WB + lq * 4]) = b;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < WK; ++k) {
            const float4 av = *reinterpret_cast<const float4 *>(&As[k * WB + ty * 4]);
            const float4 bv = *reinterpret_cast<const float4 *>(&Bs[k * WB + tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
    // dW is [Cout][Ct][ks][ks] (nn.Conv2d layout), accumulated
    const int taps = g.ks * g.ks;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= g.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ci = ci0 + tx * 4 + j;
            if (ci < Ct) atomicAdd(dw + ((int64_t)co * Ct + ci) * taps + tap, acc[i][j]);
        }
    }
}

// db[c] += sum over rows of dz[M][C]  (generic fallback: one thread per column, serial over the block's rows)
__global__ void __launch_bounds__(256) colsum_kernel(const float *__restrict__ dz, int64_t M, int C,
                                                     float *__restrict__ db, int rows_per_block) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (int64_t r = r0; r < r1; ++r) acc += dz[r * C + c];
        atomicAdd(db + c, acc);
    }
}

// Same, for C/4 a power of two <= 256: dz is read as one flat float4 stream (coalesced, every thread busy).  The grid
// stride is a multiple of C/4, so a thread always lands on the same four columns; the block folds its 256 partial sums
// through shared memory and issues one atomic per column.
__global__ void __launch_bounds__(256) colsum4_kernel(const float4 *__restrict__ dz, int64_t total4, int C4,
                                                      float *__restrict__ db) {
    __shared__ float4 part[256];
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + stride < total4; i += 2 * stride) {          // two independent loads in flight per thread
        const float4 v0 = dz[i], v1 = dz[i + stride];
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
    }
    if (i < total4) {
        const float4 v0 = dz[i];
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
    part[threadIdx.x] = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
    __syncthreads();
    if ((int)threadIdx.x < C4) {
        float4 acc = part[threadIdx.x];
        for (int j = threadIdx.x + C4; j < 256; j += C4) {
            const float4 v = part[j];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float *o = db + 4 * threadIdx.x;
        atomicAdd(o, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
    }
}

static void launch_colsum(const ramnet_handle *h, const float *dz, int64_t M, int C, float *db, cudaStream_t s) {
    const int C4 = C / 4;
    if (C % 4 == 0 && C4 <= 256 && (C4 & (C4 - 1)) == 0 && (((uintptr_t)dz) & 15) == 0) {
        const int64_t total4 = M * C4;
        int64_t blocks = (total4 + 256 * 8 - 1) / (256 * 8);           // >= 8 float4 per thread
        if (blocks > (int64_t)h->sm_count * 8) blocks = (int64_t)h->sm_count * 8;
        if (blocks < 1) blocks = 1;
        colsum4_kernel<<<(unsigned)blocks, 256, 0, s>>>((const float4 *)dz, total4, C4, db);
    } else {
        const int rpb = 2048;
        colsum_kernel<<<(unsigned)((M + rpb - 1) / rpb), 256, 0, s>>>(dz, M, C, db, rpb);
    }
}

// y[n, 2h, 2w, c] = x[n, h, w, c], zeros elsewhere (the input of a stride-2 conv's data gradient)
__global__ void __launch_bounds__(256) zero_insert_kernel(const float4 *__restrict__ x, const float4 *__restrict__ skip,
                                                          float4 *__restrict__ y, float4 *__restrict__ sum_out, int N, int H,
                                                          int W, int C4, int Hout, int Wout) {
    const int64_t total = (int64_t)N * Hout * Wout * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        int64_t p = i / C4;
        const int ox = (int)(p % Wout);
        p /= Wout;
        const int oy = (int)(p % Hout);
        const int n = (int)(p / Hout);
        float4 v = make_float4(0, 0, 0, 0);
        if (!(ox & 1) && !(oy & 1) && (oy >> 1) < H && (ox >> 1) < W) {
            const int64_t o = (((int64_t)n * H + (oy >> 1)) * W + (ox >> 1)) * C4 + c;
            v = x[o];
            if (skip) {
                const float4 sv = skip[o];
                v.x += sv.x; v.y += sv.y; v.z += sv.z; v.w += sv.w;
            }
            if (sum_out) sum_out[o] = v;       // the dense x + skip (training: the operand of the weight gradient)
        }
        y[i] = v;
    }
}

// [Cout, Cin, k, k] -> data-gradient weights for the FORWARD kernels: a conv with Cin' = Cout, Cout' = Cin,
// taps flipped.  FP32: [tap'][Cout][Cin]   TF32: [tap'][Cin][Cout] (K-major B operand), rna-rounded.
// ci_begin/ci_count select the slice of input channels (one launch per source of a virtual concat).
__global__ void pack_dgrad_kernel(const float *__restrict__ w, float *__restrict__ out, int Cout, int Cin, int taps,
                                  int kind, int ci_begin, int ci_count) {
    const int64_t total = (int64_t)Cout * ci_count * taps;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int tap = (int)(i % taps);
        const int cil = (int)((i / taps) % ci_count);
        const int co = (int)(i / ((int64_t)taps * ci_count));
        const float v = w[((int64_t)co * Cin + ci_begin + cil) * taps + tap];
        const int ftap = taps - 1 - tap;   // flip both axes
        if (kind == RAMNET_MMA_FP32)
            out[((int64_t)ftap * Cout + co) * ci_count + cil] = v;          // [tap][K = Cout][N = ci]
        else
            out[((int64_t)ftap * ci_count + cil) * Cout + co] = round_tf32(v);  // [tap][N = ci][K = Cout]
    }
}

// ---------------------------------------------------------------- pointwise adjoints (float4 over NHWC)
__device__ __forceinline__ float4 ld4(const float *p, int64_t i) { return reinterpret_cast<const float4 *>(p)[i]; }
__device__ __forceinline__ void st4(float *p, int64_t i, float4 v) { reinterpret_cast<float4 *>(p)[i] = v; }
#define F4_MAP2(a, b, expr)                                                                         \
    make_float4(([&](float A, float B) { return expr; })(a.x, b.x), ([&](float A, float B) { return expr; })(a.y, b.y), \
                ([&](float A, float B) { return expr; })(a.z, b.z), ([&](float A, float B) { return expr; })(a.w, b.w))

// dz = dy * (y > 0)
__device__ __forceinline__ float4 rnd4(float4 v, int round) {
    return round ? make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w)) : v;
}

// Bias gradient fused into the pointwise adjoints: db[c] += sum over pixels of dz[., c].  The kernels below walk the
// [M, C] tensor as a flat float4 stream with a grid stride that is a multiple of 256, so when 256 % (C/4) == 0 a thread
// only ever sees ONE column quad (threadIdx.x % (C/4)): it keeps a float4 running sum in registers, the block folds the
// 256 sums through shared memory and issues one atomicAdd per column.  This replaces a separate pass over dz per
// backward call (colsum4_kernel: 34 launches x 22 us per timestep of the bench's training step, ncu round 2).
__device__ __forceinline__ void f4_acc(float4 &a, const float4 &v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
__device__ __forceinline__ void block_colsum_atomic(float4 *part, const float4 &a, int C4, float *__restrict__ db) {
    __syncthreads();                      // `part` may still be read by a previous fold
    part[threadIdx.x] = a;
    __syncthreads();
    if ((int)threadIdx.x < C4) {
        float4 acc = part[threadIdx.x];
        for (int j = threadIdx.x + C4; j < 256; j += C4) f4_acc(acc, part[j]);
        float *o = db + 4 * threadIdx.x;
        atomicAdd(o, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
    }
}
static inline bool colsum_fusable(int C) { return C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0; }

// Atomic-free variant: every block stores its C column sums to scratch, the LAST block to finish (ticket counter) adds
// them up in a fixed order and accumulates into db.  The per-block atomics above still cost ~14 us per launch at 4 blocks
// per SM (592 x C adds onto 1 KB); here the tail is one block reading blocks x C floats (~2 us) and the result no longer
// depends on the order in which blocks retire.  scratch layout: [0] ticket counter (returns to 0), [64...] partials
// [gridDim.x][C]; one scratch region per accumulator (kColsumScratchFloats apart).
constexpr int kColsumMaxBlocks = 1024, kColsumMaxC = 1024;
constexpr size_t kColsumScratchFloats = 64 + (size_t)kColsumMaxBlocks * kColsumMaxC;
__device__ __forceinline__ void block_colsum_lastblock(float4 *part, const float4 &a, int C4, float *__restrict__ db,
                                                       float *__restrict__ scratch) {
    __shared__ unsigned ticket;
    const int C = 4 * C4;
    float *partials = scratch + 64;
    __syncthreads();
    part[threadIdx.x] = a;
    __syncthreads();
    if ((int)threadIdx.x < C4) {
        float4 acc = part[threadIdx.x];
        for (int j = threadIdx.x + C4; j < 256; j += C4) f4_acc(acc, part[j]);
        reinterpret_cast<float4 *>(partials + (size_t)blockIdx.x * C)[threadIdx.x] = acc;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket = atomicAdd(reinterpret_cast<unsigned *>(scratch), 1u);
    __syncthreads();
    if (ticket != gridDim.x - 1) return;
    __threadfence();
    // last block: thread (g, c) sums blocks g, g + G, ... of column c; the G partial sums are folded through shared memory
    const int G = 256 / C > 0 ? 256 / C : 1;
    float *fold = reinterpret_cast<float *>(part);      // 256 float4 = 1024 floats of shared memory
    for (int c0 = 0; c0 < C; c0 += 256) {
        const int c = c0 + (int)threadIdx.x % (C < 256 ? C : 256), g = C < 256 ? (int)threadIdx.x / C : 0;
        float acc = 0.f;
        if (g < G && c < C) {
            unsigned b = g;
            for (; b + 7 * G < gridDim.x; b += 8 * G) {        // eight independent loads in flight (L2 latency ~700 clk each)
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = __ldcg(partials + (size_t)(b + q * G) * C + c);
                acc += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
            }
            for (; b < gridDim.x; b += G) acc += __ldcg(partials + (size_t)b * C + c);
        }
        __syncthreads();
        fold[threadIdx.x] = acc;
        __syncthreads();
        if (g == 0 && c < C) {
            for (int k = 1; k < G; ++k) acc += fold[threadIdx.x + k * C];
            db[c] += acc;
        }
    }
    if (threadIdx.x == 0) *reinterpret_cast<unsigned *>(scratch) = 0u;
}
__device__ __forceinline__ void block_colsum(float4 *part, const float4 &a, int C4, float *__restrict__ db,
                                             float *__restrict__ scratch) {
    if (scratch) block_colsum_lastblock(part, a, C4, db, scratch);
    else block_colsum_atomic(part, a, C4, db);
}

__global__ void __launch_bounds__(256) relu_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ y,
                                                       float *__restrict__ dz, int64_t n4, int round, int C4,
                                                       float *__restrict__ db, float *__restrict__ scratch) {
    __shared__ float4 part[256];
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {          // four independent load pairs in flight per thread
        float4 g[4], v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { g[q] = ld4(dy, i + q * stride); v[q] = ld4(y, i + q * stride); }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 z = rnd4(F4_MAP2(g[q], v[q], B > 0.f ? A : 0.f), round);
            st4(dz, i + q * stride, z);
            f4_acc(sum, z);
        }
    }
    for (; i < n4; i += stride) {
        const float4 g = ld4(dy, i), v = ld4(y, i);
        const float4 z = rnd4(F4_MAP2(g, v, B > 0.f ? A : 0.f), round);
        st4(dz, i, z);
        f4_acc(sum, z);
    }
    if (db) block_colsum(part, sum, C4, db, scratch);
}

// ConvGRU candidate+blend adjoint.  in: dh' (dhn), h, u, o.  out: dzo = dh'*u*(1-o^2), dzu = dh'*(o-h)*u*(1-u)
// written into the update half of dzru [M, 2C] (columns [C, 2C)), dh = dh'*(1-u).
__global__ void __launch_bounds__(256) gru_out_bwd_kernel(const float *__restrict__ dhn, const float *__restrict__ h,
                                                          const float *__restrict__ u, const float *__restrict__ o,
                                                          float *__restrict__ dzo, float *__restrict__ dzru,
                                                          float *__restrict__ dh, int64_t M, int C, int round,
                                                          float *__restrict__ db_o, float *__restrict__ db_ru,
                                                          float *__restrict__ scratch) {
    __shared__ float4 part[256];
    float4 sum_o = make_float4(0.f, 0.f, 0.f, 0.f), sum_u = sum_o;
    const int C4 = C >> 2;
    const int64_t n4 = M * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / C4;
        const int c = (int)(i % C4);
        const float4 g = ld4(dhn, i), hv = ld4(h, i), uv = ld4(u, i), ov = ld4(o, i);
        float4 a, b, d;
        a.x = g.x * uv.x * (1.f - ov.x * ov.x); a.y = g.y * uv.y * (1.f - ov.y * ov.y);
        a.z = g.z * uv.z * (1.f - ov.z * ov.z); a.w = g.w * uv.w * (1.f - ov.w * ov.w);
        b.x = g.x * (ov.x - hv.x) * uv.x * (1.f - uv.x); b.y = g.y * (ov.y - hv.y) * uv.y * (1.f - uv.y);
        b.z = g.z * (ov.z - hv.z) * uv.z * (1.f - uv.z); b.w = g.w * (ov.w - hv.w) * uv.w * (1.f - uv.w);
        d.x = g.x * (1.f - uv.x); d.y = g.y * (1.f - uv.y); d.z = g.z * (1.f - uv.z); d.w = g.w * (1.f - uv.w);
        a = rnd4(a, round); b = rnd4(b, round);
        st4(dzo, i, a);
        st4(dzru, m * (2 * C4) + C4 + c, b);
        st4(dh, i, d);
        f4_acc(sum_o, a); f4_acc(sum_u, b);
    }
    if (db_o) block_colsum(part, sum_o, C4, db_o, scratch);
    if (db_ru) block_colsum(part, sum_u, C4, db_ru + C, scratch ? scratch + kColsumScratchFloats : nullptr);   // update gate = rows [C, 2C)
}

// ConvGRU reset adjoint.  in: drh (grad of h*r), h, r.  out: dzr = drh*h*r*(1-r) into columns [0, C) of dzru,
// dh += drh*r.
__global__ void __launch_bounds__(256) gru_ru_bwd_kernel(const float *__restrict__ drh, const float *__restrict__ h,
                                                         const float *__restrict__ r, float *__restrict__ dzru,
                                                         float *__restrict__ dh, int64_t M, int C, int round,
                                                         float *__restrict__ db_ru, float *__restrict__ scratch) {
    __shared__ float4 part[256];
    float4 sum_r = make_float4(0.f, 0.f, 0.f, 0.f);
    const int C4 = C >> 2;
    const int64_t n4 = M * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / C4;
        const int c = (int)(i % C4);
        const float4 g = ld4(drh, i), hv = ld4(h, i), rv = ld4(r, i);
        float4 a, d = ld4(dh, i);
        a.x = g.x * hv.x * rv.x * (1.f - rv.x); a.y = g.y * hv.y * rv.y * (1.f - rv.y);
        a.z = g.z * hv.z * rv.z * (1.f - rv.z); a.w = g.w * hv.w * rv.w * (1.f - rv.w);
        d.x += g.x * rv.x; d.y += g.y * rv.y; d.z += g.z * rv.z; d.w += g.w * rv.w;
        a = rnd4(a, round);
        st4(dzru, m * (2 * C4) + c, a);
        st4(dh, i, d);
        f4_acc(sum_r, a);
    }
    if (db_ru) block_colsum(part, sum_r, C4, db_ru, scratch);          // reset gate = rows [0, C)
}

// ConvLSTM adjoint (submodules.py:341-356).  gates: post-activation [M][C][4] = (i, f, o, g) as stashed by the forward
// epilogue.  dz is written in the ORIGINAL nn.Conv2d row order (gate-major blocks: in | remember | out | cell) so that
// the weight / data gradients use the unpermuted Gates.weight.
__global__ void __launch_bounds__(256) lstm_bwd_kernel(const float *__restrict__ dh, const float *__restrict__ dc,
                                                       const float *__restrict__ gates, const float *__restrict__ c_prev,
                                                       const float *__restrict__ c_new, float *__restrict__ dz,
                                                       float *__restrict__ dc_prev, int64_t M, int C, int round,
                                                       float *__restrict__ db) {
    __shared__ float4 part[256];
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);      // (in, remember, out, cell) of this thread's channel (256 % C == 0)
    const int64_t n = M * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / C;
        const int c = (int)(i % C);
        const float4 g = *reinterpret_cast<const float4 *>(gates + i * 4);   // i, f, o, gc
        const float tc = tanhf(c_new[i]);
        const float gh = dh ? dh[i] : 0.f;
        const float dct = (dc ? dc[i] : 0.f) + gh * g.z * (1.f - tc * tc);
        float zi = dct * g.w * g.x * (1.f - g.x);
        float zf = dct * c_prev[i] * g.y * (1.f - g.y);
        float zo = gh * tc * g.z * (1.f - g.z);
        float zg = dct * g.x * (1.f - g.w * g.w);
        if (round) { zi = round_tf32(zi); zf = round_tf32(zf); zo = round_tf32(zo); zg = round_tf32(zg); }
        float *row = dz + m * 4 * C;
        row[c] = zi; row[C + c] = zf; row[2 * C + c] = zo; row[3 * C + c] = zg;
        dc_prev[i] = dct * g.y;
        sum.x += zi; sum.y += zf; sum.z += zo; sum.w += zg;
    }
    if (db) {
        part[threadIdx.x] = sum;
        __syncthreads();
        if ((int)threadIdx.x < C) {
            float4 acc = part[threadIdx.x];
            for (int j = threadIdx.x + C; j < 256; j += C) f4_acc(acc, part[j]);
            atomicAdd(db + threadIdx.x, acc.x); atomicAdd(db + C + threadIdx.x, acc.y);
            atomicAdd(db + 2 * C + threadIdx.x, acc.z); atomicAdd(db + 3 * C + threadIdx.x, acc.w);
        }
    }
}

// prediction head adjoint: dlogit = ddepth * s(1-s); dx[m, c] = dlogit[m] * w[c]; dw[c] += sum_m dlogit*x[m,c]; db += sum dlogit
// w_skip != nullptr (skip_type 'concat'): the skip has its own weights: dskip[m, c] = dlogit[m] * w_skip[c],
// dw[C + c] += sum_m dlogit * skip[m, c], and dw[c] sums x alone.
__global__ void __launch_bounds__(256) pred_bwd_kernel(const float *__restrict__ ddepth, const float *__restrict__ depth,
                                                       const float *__restrict__ x, const float *__restrict__ skip,
                                                       const float *__restrict__ w, const float *__restrict__ w_skip,
                                                       float *__restrict__ dx, float *__restrict__ dskip, float *__restrict__ dw,
                                                       float *__restrict__ db, int64_t M, int C) {
    extern __shared__ float sacc[];   // [2C + 1]: dw (x part), dw (skip part, concat only), db
    for (int c = threadIdx.x; c <= 2 * C; c += blockDim.x) sacc[c] = 0.f;
    __syncthreads();
    const int lane8 = threadIdx.x & 7;
    float wreg[8], wsreg[8], dwreg[8], dwsreg[8];   // C <= 64: lane handles channels lane8*4.. (+32)
#pragma unroll
    for (int j = 0; j < 8; ++j) { wreg[j] = 0.f; wsreg[j] = 0.f; dwreg[j] = 0.f; dwsreg[j] = 0.f; }
    for (int c = lane8 * 4, j = 0; c < C; c += 32, j += 4)
        for (int e = 0; e < 4; ++e) {
            wreg[j + e] = w[c + e];
            if (w_skip) wsreg[j + e] = w_skip[c + e];
        }
    float dbacc = 0.f;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 3;
    for (int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3; m < M; m += stride) {
        float dl = ddepth[m];
        if (depth) {
            const float s = depth[m];
            dl *= s * (1.f - s);
        }
        if (lane8 == 0) dbacc += dl;
        for (int c = lane8 * 4, j = 0; c < C; c += 32, j += 4) {
            float4 xv = *reinterpret_cast<const float4 *>(x + m * C + c);
            if (skip) {
                const float4 sv = *reinterpret_cast<const float4 *>(skip + m * C + c);
                if (w_skip) {
                    *reinterpret_cast<float4 *>(dskip + m * C + c) =
                        make_float4(dl * wsreg[j], dl * wsreg[j + 1], dl * wsreg[j + 2], dl * wsreg[j + 3]);
                    dwsreg[j] += dl * sv.x; dwsreg[j + 1] += dl * sv.y; dwsreg[j + 2] += dl * sv.z; dwsreg[j + 3] += dl * sv.w;
                } else {
                    xv.x += sv.x; xv.y += sv.y; xv.z += sv.z; xv.w += sv.w;
                }
            }
            *reinterpret_cast<float4 *>(dx + m * C + c) = make_float4(dl * wreg[j], dl * wreg[j + 1], dl * wreg[j + 2], dl * wreg[j + 3]);
            dwreg[j] += dl * xv.x; dwreg[j + 1] += dl * xv.y; dwreg[j + 2] += dl * xv.z; dwreg[j + 3] += dl * xv.w;
        }
    }
    for (int c = lane8 * 4, j = 0; c < C; c += 32, j += 4)
        for (int e = 0; e < 4; ++e) {
            atomicAdd(&sacc[c + e], dwreg[j + e]);
            if (w_skip) atomicAdd(&sacc[C + c + e], dwsreg[j + e]);
        }
    if (lane8 == 0) atomicAdd(&sacc[2 * C], dbacc);
    __syncthreads();
    for (int c = threadIdx.x; c < (w_skip ? 2 * C : C); c += blockDim.x) atomicAdd(dw + c, sacc[c]);
    if (threadIdx.x == 0 && db) atomicAdd(db, sacc[2 * C]);
}

// adjoint of (x + skip) -> bilinear x2: dx[m] gathers from the <= 3x3 output pixels it feeds; dskip = dx.
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const float4 *__restrict__ dy, float4 *__restrict__ dx, int N,
                                                             int H, int W, int C4) {
    const int64_t total = (int64_t)N * H * W * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        int64_t p = i / C4;
        const int xi = (int)(p % W);
        p /= W;
        const int yi = (int)(p % H);
        const int n = (int)(p / H);
        // output index o reads input floor-index f(o) = (o>>1) - (o even) with weight (o even ? .25 : .75) and f+1 with the
        // complement, both clamped to [0, n-1].  Enumerate the outputs 2m-2 .. 2m+2 that can touch input m.
        float wy[6];
        int oyv[6], ny = 0;
        for (int o = 2 * yi - 2; o <= 2 * yi + 3; ++o) {
            if (o < 0 || o >= 2 * H) continue;
            const int f = (o >> 1) - ((o & 1) ? 0 : 1);
            const float w1 = (o & 1) ? 0.25f : 0.75f;       // weight of tap f+1
            float wgt = 0.f;
            if (max(f, 0) == yi) wgt += 1.f - w1;
            if (min(f + 1, H - 1) == yi) wgt += w1;
            if (wgt != 0.f) { wy[ny] = wgt; oyv[ny] = o; ++ny; }
        }
        float4 acc = make_float4(0, 0, 0, 0);
        for (int o = 2 * xi - 2; o <= 2 * xi + 3; ++o) {
            if (o < 0 || o >= 2 * W) continue;
            const int f = (o >> 1) - ((o & 1) ? 0 : 1);
            const float w1 = (o & 1) ? 0.25f : 0.75f;
            float wx = 0.f;
            if (max(f, 0) == xi) wx += 1.f - w1;
            if (min(f + 1, W - 1) == xi) wx += w1;
            if (wx == 0.f) continue;
            for (int k = 0; k < ny; ++k) {
                const float4 g = dy[(((int64_t)n * 2 * H + oyv[k]) * 2 * W + o) * C4 + c];
                const float ww = wx * wy[k];
                acc.x += ww * g.x; acc.y += ww * g.y; acc.z += ww * g.z; acc.w += ww * g.w;
            }
        }
        dx[i] = acc;
    }
}


// Head conv weight gradient: dW[co][ci][r][s] = sum_{n,y,x} dz[n,y,x,co] * x[n,ci,y+r-2,x+s-2]   (NCHW input,
// Cin <= 8, 5x5, NHWC dz with Cout <= 32).  CTA = one 8x32 pixel tile: input halo and dz tile in shared
// memory, thread = one (ci, r, s) x 16 output channels, 256-pixel reduction in registers, then atomics.
constexpr int HW_TW = 32, HW_TH = 8;
__global__ void __launch_bounds__(256) head_wgrad_kernel(const float *__restrict__ x, const float *__restrict__ dz,
                                                         float *__restrict__ dw, int N, int Cin, int H, int W,
                                                         int Cout) {
    __shared__ float in_s[8 * (HW_TH + 4) * (HW_TW + 4)];
    __shared__ __align__(16) float dz_s[HW_TH * HW_TW * 32];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * HW_TW, y0 = blockIdx.y * HW_TH, n = blockIdx.z;
    constexpr int IW = HW_TW + 4, IH = HW_TH + 4;
    for (int i = tid; i < Cin * IH * IW; i += 256) {
        const int c = i / (IH * IW), r = (i / IW) % IH, sx = i % IW;
        const int gy = y0 + r - 2, gx = x0 + sx - 2;
        in_s[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(x + (((int64_t)n * Cin + c) * H + gy) * W + gx) : 0.f;
    }
    for (int i = tid; i < HW_TH * HW_TW * 8; i += 256) {   // float4 granules: pixel p, quad q
        const int p = i >> 3, q = i & 7;
        const int gy = y0 + p / HW_TW, gx = x0 + p % HW_TW;
        float4 v = make_float4(0, 0, 0, 0);
        if (gy < H && gx < W && q * 4 < Cout) v = *reinterpret_cast<const float4 *>(dz + (((int64_t)n * H + gy) * W + gx) * Cout + q * 4);
        *reinterpret_cast<float4 *>(&dz_s[p * 32 + q * 4]) = v;
    }
    __syncthreads();
    const int half = tid & 1;
    for (int combo = tid >> 1; combo < Cin * 25; combo += 128) {      // 128 (ci, r, s) combos per sweep (Cin <= 8: 2 sweeps)
        const int ci = combo / 25, r = (combo % 25) / 5, sx = combo % 5;
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        for (int p = 0; p < HW_TH * HW_TW; ++p) {
            const float v = in_s[(ci * IH + p / HW_TW + r) * IW + p % HW_TW + sx];
            const float4 *d4 = reinterpret_cast<const float4 *>(&dz_s[p * 32 + half * 16]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 g = d4[q];
                acc[4 * q] = fmaf(v, g.x, acc[4 * q]);
                acc[4 * q + 1] = fmaf(v, g.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(v, g.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(v, g.w, acc[4 * q + 3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int co = half * 16 + j;
            if (co < Cout) atomicAdd(dw + (((int64_t)co * Cin + ci) * 5 + r) * 5 + sx, acc[j]);
        }
    }
}

int grid_for(ramnet_handle *h, int64_t n) { return (int)imin64((n + 255) / 256, (int64_t)h->sm_count * 16); }
// Launches that fold a bias gradient in end with one atomicAdd per column per BLOCK; with the wide grid above that was
// 2368 blocks x C adds onto 1 KB of addresses and cost ~50 us per launch (ncu, round 2: relu_bwd 13 -> 67 us).  Four
// blocks per SM keep the kernels at the HBM rate (several independent loads in flight per thread) with 4x fewer atomics.
int grid_for_colsum(ramnet_handle *h, int64_t n, bool fused) {
    return fused ? (int)imin64((n + 255) / 256, (int64_t)h->sm_count * 4) : grid_for(h, n);
}
}  // namespace

int conv_wgrad_tf32(ramnet_handle *h, const ramnet_conv_desc *d, const float *dz, const float *x0, const float *x1,
                    float *dw, void *workspace, size_t workspace_bytes, cudaStream_t s, int head_cin, int mode);
size_t conv_wgrad_tf32_workspace(const ramnet_handle *h, const ramnet_conv_desc *d, int head_cin = 0);

extern "C" size_t ramnet_conv_wgrad_workspace_bytes(const ramnet_handle *h, const ramnet_conv_desc *d) {
    RAMNET_DEVICE_GUARD(h);
    if (!h || !d || d->mma_kind != RAMNET_MMA_TF32) return 0;
    return conv_wgrad_tf32_workspace(h, d);
}

// Head conv weight gradient on the tensor cores: X is ramnet_head_im2row's unrolled tensor, so this is the tap-packed
// wgrad of a 5x1 conv over 32 channels whose scatter writes the head's [Cout][Cin][5][5] layout.
static ramnet_conv_desc head_wgrad_desc(int N, int H, int W, int Cout) {
    ramnet_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.N = N; d.H = H; d.W = W; d.C0 = 32; d.C1 = 0; d.Cout = Cout; d.ksize = 5; d.stride = 1;
    d.epilogue = RAMNET_EPI_BIAS; d.mma_kind = RAMNET_MMA_TF32;
    return d;
}

extern "C" size_t ramnet_head_conv_wgrad_tc_workspace_bytes(const ramnet_handle *h, int N, int Cin, int H, int W, int Cout) {
    RAMNET_DEVICE_GUARD(h);
    if (!h || Cin < 1 || 5 * Cin > 32 || Cout <= 0 || Cout % 32) return 0;
    const ramnet_conv_desc d = head_wgrad_desc(N, H, W, Cout);
    return conv_wgrad_tf32_workspace(h, &d, Cin);
}

extern "C" int ramnet_head_conv_wgrad_tc(ramnet_handle *h, const float *xe_nhwc32, const float *dz_nhwc, float *dw_oihw,
                                         float *db, int N, int Cin, int H, int W, int Cout, void *workspace,
                                         size_t workspace_bytes, int mode, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && (mode == RAMNET_WGRAD_FINALIZE || (xe_nhwc32 && dz_nhwc)) &&
                         (mode == RAMNET_WGRAD_PARTIAL_FIRST || mode == RAMNET_WGRAD_PARTIAL_ADD || dw_oihw),
                     "head_conv_wgrad_tc: NULL argument");
    RAMNET_CHECK_ARG(Cin >= 1 && 5 * Cin <= 32 && Cout > 0 && Cout % 32 == 0, "head_conv_wgrad_tc: needs 5*Cin <= 32 and Cout %% 32 == 0 (got %d, %d)", Cin, Cout);
    if (db && dz_nhwc && mode != RAMNET_WGRAD_FINALIZE) {
        launch_colsum(h, dz_nhwc, (int64_t)N * H * W, Cout, db, (cudaStream_t)stream);
        RAMNET_LAUNCH_CHECK(h);
    }
    const ramnet_conv_desc d = head_wgrad_desc(N, H, W, Cout);
    return conv_wgrad_tf32(h, &d, dz_nhwc, xe_nhwc32, nullptr, dw_oihw, workspace, workspace_bytes, (cudaStream_t)stream, Cin,
                           mode);
}

extern "C" int ramnet_conv_wgrad(ramnet_handle *h, const ramnet_conv_desc *d, const float *dz, const float *x0,
                                 const float *x1, float *dw_oihw, float *db, void *workspace, size_t workspace_bytes,
                                 int mode, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && d && mode >= RAMNET_WGRAD_FULL && mode <= RAMNET_WGRAD_FINALIZE, "conv_wgrad: NULL argument / bad mode");
    RAMNET_CHECK_ARG(d->C0 % 4 == 0 && d->C1 % 4 == 0 && d->Cout % 4 == 0, "conv_wgrad: channel counts must be multiples of 4");
    if (mode != RAMNET_WGRAD_FULL) {        // deferred modes: the tap-packed TF32 kernel only, no bias gradient here
        RAMNET_CHECK_ARG(d->mma_kind == RAMNET_MMA_TF32 && !db, "conv_wgrad: deferred modes need mma_kind=TF32 and db=NULL");
        RAMNET_CHECK_ARG(mode == RAMNET_WGRAD_FINALIZE ? dw_oihw != nullptr : (dz && x0 && (d->C1 == 0) == (x1 == nullptr)),
                         "conv_wgrad: NULL argument");
        return conv_wgrad_tf32(h, d, dz, x0, x1, dw_oihw, workspace, workspace_bytes, (cudaStream_t)stream, 0, mode);
    }
    RAMNET_CHECK_ARG(dz && x0 && dw_oihw, "conv_wgrad: NULL argument");
    RAMNET_CHECK_ARG((d->C1 == 0) == (x1 == nullptr), "conv_wgrad: x1 and C1 disagree");
    if (db) {
        const int64_t Mrows = (int64_t)d->N * conv_out_dim(d->H, d->stride) * conv_out_dim(d->W, d->stride);
        launch_colsum(h, dz, Mrows, d->Cout, db, (cudaStream_t)stream);
        RAMNET_LAUNCH_CHECK(h);
    }
    if (d->mma_kind == RAMNET_MMA_TF32) {
        const int rc = conv_wgrad_tf32(h, d, dz, x0, x1, dw_oihw, workspace, workspace_bytes, (cudaStream_t)stream, 0,
                                       RAMNET_WGRAD_FULL);
        if (rc != RAMNET_EUNSUPPORTED) return rc;      // unsupported shape: fp32 FFMA kernel below
    }
    db = nullptr;
    WgradGeom g;
    g.N = d->N; g.H = d->H; g.W = d->W; g.C0 = d->C0; g.C1 = d->C1; g.Cout = d->Cout; g.ks = d->ksize;
    g.stride = d->stride; g.pad = d->ksize / 2;
    g.Ho = conv_out_dim(d->H, d->stride); g.Wo = conv_out_dim(d->W, d->stride);
    g.M = (int64_t)g.N * g.Ho * g.Wo;
    const int Ct = d->C0 + d->C1, taps = d->ksize * d->ksize;
    const int tiles = taps * ((Ct + WB - 1) / WB) * ((d->Cout + WB - 1) / WB);
    // split K so that the launch has ~8 CTAs per SM
    int64_t splits = ((int64_t)h->sm_count * 8 + tiles - 1) / tiles;
    int64_t ppb = (g.M + splits - 1) / splits;
    ppb = ((ppb + WK - 1) / WK) * WK;
    if (ppb < 256) ppb = 256;
    g.pix_per_block = (int)ppb;
    dim3 grid((unsigned)((g.M + ppb - 1) / ppb), (unsigned)(taps * ((Ct + WB - 1) / WB)), (unsigned)((d->Cout + WB - 1) / WB));
    conv_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, dz, x0, x1, dw_oihw);
    RAMNET_LAUNCH_CHECK(h);
    if (db) {
        launch_colsum(h, dz, g.M, d->Cout, db, (cudaStream_t)stream);
        RAMNET_LAUNCH_CHECK(h);
    }
    return RAMNET_OK;
}

extern "C" int ramnet_zero_insert2x(ramnet_handle *h, const float *x, const float *skip, float *y, float *sum_out, int N,
                                    int H, int W, int C, int Hout, int Wout, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && x && y && C % 4 == 0 && Hout >= 2 * H - 1 && Wout >= 2 * W - 1, "zero_insert2x: bad argument");
    const int64_t total = (int64_t)N * Hout * Wout * (C / 4);
    zero_insert_kernel<<<grid_for(h, total), 256, 0, (cudaStream_t)stream>>>((const float4 *)x, (const float4 *)skip, (float4 *)y,
                                                                            (float4 *)sum_out, N, H, W, C / 4, Hout, Wout);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_pack_weights_dgrad(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin,
                                         int ksize, int mma_kind, int ci_begin, int ci_count, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && w_oihw && w_packed && ci_begin >= 0 && ci_count > 0 && ci_begin + ci_count <= Cin,
                     "pack_weights_dgrad: bad argument");
    const int64_t total = (int64_t)Cout * ci_count * ksize * ksize;
    pack_dgrad_kernel<<<grid_for(h, total), 256, 0, (cudaStream_t)stream>>>(w_oihw, w_packed, Cout, Cin, ksize * ksize,
                                                                           mma_kind, ci_begin, ci_count);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" size_t ramnet_colsum_scratch_bytes(void) { return 2 * kColsumScratchFloats * sizeof(float); }

extern "C" int ramnet_relu_bwd(ramnet_handle *h, const float *dy, const float *y, float *dz, int64_t n, int C, float *db,
                               float *scratch, int flags, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && dy && y && dz && n > 0 && n % 4 == 0, "relu_bwd: bad argument");
    RAMNET_CHECK_ARG(!db || (C > 0 && n % C == 0), "relu_bwd: db needs the channel count C (n %% C == 0)");
    const bool fuse = db && colsum_fusable(C);
    relu_bwd_kernel<<<grid_for_colsum(h, n / 4, fuse), 256, 0, (cudaStream_t)stream>>>(dy, y, dz, n / 4, flags & RAMNET_FLAG_ROUND_TF32,
                                                                         fuse ? C / 4 : 1, fuse ? db : nullptr, scratch);
    RAMNET_LAUNCH_CHECK(h);
    if (db && !fuse) {
        launch_colsum(h, dz, n / C, C, db, (cudaStream_t)stream);
        RAMNET_LAUNCH_CHECK(h);
    }
    return RAMNET_OK;
}

extern "C" int ramnet_gru_out_bwd(ramnet_handle *h, const float *dhn, const float *hprev, const float *u, const float *o,
                                  float *dzo, float *dzru, float *dh, float *db_o, float *db_ru, float *scratch, int64_t M,
                                  int C, int flags, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && dhn && hprev && u && o && dzo && dzru && dh && M > 0 && C % 4 == 0, "gru_out_bwd: bad argument");
    RAMNET_CHECK_ARG((!db_o && !db_ru) || colsum_fusable(C), "gru_out_bwd: fused bias gradients need 256 %% (C/4) == 0");
    gru_out_bwd_kernel<<<grid_for_colsum(h, M * (C / 4), db_o || db_ru), 256, 0, (cudaStream_t)stream>>>(
        dhn, hprev, u, o, dzo, dzru, dh, M, C, flags & RAMNET_FLAG_ROUND_TF32, db_o, db_ru, scratch);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_gru_ru_bwd(ramnet_handle *h, const float *drh, const float *hprev, const float *r, float *dzru,
                                 float *dh, float *db_ru, float *scratch, int64_t M, int C, int flags, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && drh && hprev && r && dzru && dh && M > 0 && C % 4 == 0, "gru_ru_bwd: bad argument");
    RAMNET_CHECK_ARG(!db_ru || colsum_fusable(C), "gru_ru_bwd: the fused bias gradient needs 256 %% (C/4) == 0");
    gru_ru_bwd_kernel<<<grid_for_colsum(h, M * (C / 4), db_ru != nullptr), 256, 0, (cudaStream_t)stream>>>(
        drh, hprev, r, dzru, dh, M, C, flags & RAMNET_FLAG_ROUND_TF32, db_ru, scratch);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_pred_bwd(ramnet_handle *h, const float *ddepth, const float *depth, const float *x,
                               const float *skip, const float *w, const float *w_skip, float *dx, float *dskip, float *dw,
                               float *db, int64_t M, int C, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && ddepth && x && w && dx && dw && M > 0 && C % 4 == 0 && C <= 64, "pred_bwd: bad argument (C <= 64)");
    RAMNET_CHECK_ARG(!w_skip || (skip && dskip), "pred_bwd: the concat form needs skip and dskip");
    const int blocks = (int)imin64((M * 8 + 255) / 256, (int64_t)h->sm_count * 8);
    pred_bwd_kernel<<<blocks, 256, (2 * C + 1) * sizeof(float), (cudaStream_t)stream>>>(ddepth, depth, x, skip, w, w_skip, dx, dskip,
                                                                                        dw, db, M, C);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_upsample2x_bwd(ramnet_handle *h, const float *dy, float *dx, int N, int H, int W, int C,
                                     void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && dy && dx && N > 0 && H > 0 && W > 0 && C % 4 == 0, "upsample2x_bwd: bad argument");
    const int64_t total = (int64_t)N * H * W * (C / 4);
    upsample2x_bwd_kernel<<<grid_for(h, total), 256, 0, (cudaStream_t)stream>>>((const float4 *)dy, (float4 *)dx, N, H, W, C / 4);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_head_conv_wgrad(ramnet_handle *h, const float *x_nchw, const float *dz_nhwc, float *dw_oihw,
                                      float *db, int N, int Cin, int H, int W, int Cout, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && x_nchw && dz_nhwc && dw_oihw, "head_conv_wgrad: NULL argument");
    RAMNET_CHECK_ARG(Cin >= 1 && Cin <= 8 && Cout % 4 == 0 && Cout <= 32, "head_conv_wgrad: Cin in [1,8], Cout <= 32 (multiple of 4)");
    dim3 grid((W + HW_TW - 1) / HW_TW, (H + HW_TH - 1) / HW_TH, N);
    head_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_nchw, dz_nhwc, dw_oihw, N, Cin, H, W, Cout);
    RAMNET_LAUNCH_CHECK(h);
    if (db) {
        const int64_t M = (int64_t)N * H * W;
        launch_colsum(h, dz_nhwc, M, Cout, db, (cudaStream_t)stream);
        RAMNET_LAUNCH_CHECK(h);
    }
    return RAMNET_OK;
}

extern "C" int ramnet_lstm_bwd(ramnet_handle *h, const float *dh, const float *dc, const float *gates, const float *c_prev,
                               const float *c_new, float *dz, float *dc_prev, float *db, int64_t M, int C, int flags,
                               void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && gates && c_prev && c_new && dz && dc_prev && (dh || dc) && M > 0 && C > 0, "lstm_bwd: bad argument");
    const bool fuse = db && C <= 256 && 256 % C == 0;
    lstm_bwd_kernel<<<grid_for_colsum(h, M * C, fuse), 256, 0, (cudaStream_t)stream>>>(dh, dc, gates, c_prev, c_new, dz, dc_prev, M, C,
                                                                                      flags & RAMNET_FLAG_ROUND_TF32, fuse ? db : nullptr);
    RAMNET_LAUNCH_CHECK(h);
    if (db && !fuse) {
        launch_colsum(h, dz, M, 4 * C, db, (cudaStream_t)stream);
        RAMNET_LAUNCH_CHECK(h);
    }
    return RAMNET_OK;
}

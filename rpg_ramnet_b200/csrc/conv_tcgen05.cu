// Implicit-GEMM convolution on the 5th-gen tensor cores (RAMNET_MMA_TF32): TMA -> smem ->
// tcgen05.mma kind::tf32 -> TMEM -> fused epilogue.  sm_100a only.
//
// GEMM view:  D[m, n] = sum_{tap, c} A[pix(m, tap), c] * W[tap, n, c]
//   m = output pixel of a TH x TW patch of one image (TH*TW = 128 = UMMA M),
//   n = output channel (GEMM column; BN per CTA, UMMA N),
//   k = (filter tap, input channel), walked as taps x 32-channel chunks (32 fp32 = one 128-byte
//       swizzle row, 4 UMMA K-steps of 8).
//
// Data movement: activations are NHWC fp32, so the 32 channels of one pixel are one contiguous
// 128-byte row.  For a stride-1 conv the A tile of tap (r, s) is the TMA box
// {32 ch, TW px, TH rows, 1 image} at (c, x0+s-pad, y0+r-pad, n) of the 4-D tensor (C, W, H, N):
// the hardware zero-fills out-of-bounds pixels, which IS the convolution's zero padding, so the
// im2col matrix never exists anywhere.  Stride-2 convs view the same memory as the 5-D tensor
// (2C, W/2, 2, H/2, N) (column parity folded into the channel axis, row parity its own axis) so
// that "every other pixel" is again a dense box.  The virtual concat [x0 | x1] of the recurrent
// cells is two tensor maps walked back to back.  Weights are packed [tap][Cout][Cin] (K-major)
// and fetched as {32 ch, BN rows} boxes.  Both operands land in 128B-swizzled K-major smem tiles
// that tcgen05.mma consumes through shared-memory descriptors; accumulators live in TMEM.
//
// Roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane),
// warps 2-5 = epilogue (TMEM lane quarter = warp_id % 4).  mbarrier rings: full[s] (TMA ->
// MMA), empty[s] (tcgen05.commit -> TMA), accum (last commit -> epilogue).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

#ifndef RAMNET_EPI_WARPS
#define RAMNET_EPI_WARPS 8
#endif

namespace {

constexpr int kTileM = 128;        // UMMA M
constexpr int kChunk = 32;         // channels per K-chunk (128 bytes of fp32)
constexpr int kABytes = kTileM * kChunk * 4;  // 16 KB
constexpr int kThreads = 192;
constexpr uint32_t kSpinLimit = 1u << 28;     // bring-up guard: trap instead of hanging the GPU

struct TcGeom {
    int N, Ho, Wo, Cout;
    int C0, C1;          // channels of the two sources
    int ks, stride, pad;
    int TW, TH;          // output patch (TW*TH = 128)
    int tiles_x, tiles_y;
    int BN;              // GEMM columns per CTA
    int stages;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a fully-converged warp (deterministic leader): control flow stays warp-uniform so the
// compiler keeps descriptors in uniform registers and emits a single predicated UTCHMMA / UTMALDG.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && ++spins > kSpinLimit) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
        "[%2];" ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                     // leading byte offset (unused for swizzled K-major) = 16 B
    d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset: next 8-row atom
    d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                     // layout: SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = n.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_m(int n, int m) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------- kernel
template <int EPI>
__global__ void __launch_bounds__(kThreads) conv_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x0,
                                                                const __grid_constant__ CUtensorMap map_x1,
                                                                const __grid_constant__ CUtensorMap map_w, TcGeom g,
                                                                EpiParams ep) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = g.BN * kChunk * 4;
    const int stage_bytes = kABytes + b_bytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)g.stages * stage_bytes);
    uint64_t *empty_bar = full_bar + g.stages;
    uint64_t *accum_bar = empty_bar + g.stages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates
    int t = blockIdx.x;
    const int tx = t % g.tiles_x;
    t /= g.tiles_x;
    const int ty = t % g.tiles_y;
    const int img = t / g.tiles_y;
    const int x0 = tx * g.TW, y0 = ty * g.TH;
    const int n0 = blockIdx.y * g.BN;

    const int chunks0 = g.C0 / kChunk, chunks = (g.C0 + g.C1) / kChunk;
    const int ksteps = g.ks * g.ks * chunks;
    // TMEM columns: power of two >= 32
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < g.BN) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x0);
        prefetch_tmap(&map_x1);
        prefetch_tmap(&map_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int r = 0; r < g.ks; ++r) {
            for (int s = 0; s < g.ks; ++s) {
                const int tap = r * g.ks + s;
                for (int ch = 0; ch < chunks; ++ch) {
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    if (elect_one()) {
                        uint8_t *sa = smem + (size_t)stage * stage_bytes;
                        uint8_t *sb = sa + kABytes;
                        mbar_expect_tx(full_bar + stage, (uint32_t)stage_bytes);
                        const bool second = ch >= chunks0;
                        const CUtensorMap *mx = second ? &map_x1 : &map_x0;
                        const int c = (second ? ch - chunks0 : ch) * kChunk;
                        if (g.stride == 1) {
                            tma_load_4d(sa, mx, full_bar + stage, c, x0 + s - g.pad, y0 + r - g.pad, img);
                        } else {
                            // input row 2*oy + (r - pad) = 2*(oy + dy) + py, likewise for columns
                            const int ry = r - g.pad, rx = s - g.pad;
                            const int dy = ry >> 1, py = ry & 1, dx = rx >> 1, px = rx & 1;  // arithmetic shift = floor
                            const int Csrc = second ? g.C1 : g.C0;
                            tma_load_5d(sa, mx, full_bar + stage, px * Csrc + c, x0 + dx, py, y0 + dy, img);
                        }
                        tma_load_3d(sb, &map_w, full_bar + stage, ch * kChunk, n0, tap);
                    }
                    __syncwarp();
                    if (++stage == g.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        const uint32_t idesc = make_idesc_tf32(g.BN);
        int stage = 0;
        uint32_t phase = 0;
        for (int k = 0; k < ksteps; ++k) {
            mbar_wait(full_bar + stage, phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + kABytes);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kChunk / 8; ++kk)  // 8 tf32 = 32 bytes per UMMA K-step: +2 in the >>4 address field
                    umma_tf32(tmem_base, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0);
                umma_commit(empty_bar + stage);  // frees the smem slot once these MMAs have read it
                if (k == ksteps - 1) umma_commit(accum_bar);  // accumulator complete
            }
            __syncwarp();
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> fused math -> global =====================
        const int quarter = warp & 3;             // TMEM lanes [32q, 32q+32) are visible to warps with id%4 == q
        const int row = quarter * 32 + lane;      // GEMM row inside the tile
        const int oy = y0 + row / g.TW, ox = x0 + row % g.TW;
        const bool valid = oy < g.Ho && ox < g.Wo;
        const int64_t m = ((int64_t)img * g.Ho + oy) * g.Wo + ox;
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        for (int c = 0; c < g.BN; c += 16) {
            float v[16];
            tmem_ld16(lane_addr + (uint32_t)c, v);   // warp-collective: every lane participates
            if (valid) epilogue_store<EPI, 16>(ep, m, n0 + c, v);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}


// ================================================================================================
// Persistent "halo" kernel for stride-1 convolutions: tap reuse out of shared memory.
//
// The per-tap kernel above re-reads every input pixel ks*ks times from L2 and is pinned at the
// ~15 TB/s L2->SM ceiling (profiles/r01_layers_v1_per_tap.txt, ncu: l1tex__m_xbar2l1tex_read_bytes).
// Here one TMA box brings the patch PLUS its halo, (8*PTX + ks-1) x (16*PTY + ks-1) pixels x 32
// channels, into smem once per 32-channel chunk, and all ks*ks taps are issued from it: the A operand
// of tap (r, s) for the M-tile at (tx, ty) is the SAME buffer addressed through a descriptor whose
// start is shifted by ((16*ty + r) * HX + 8*tx + s) pixel rows (128 B each) and whose 8-row-group
// stride (SBO) is one halo row (HX * 128 B): an M-tile is 8 pixels wide (one swizzle atom) by 16 tall.
// This works because the 128B swizzle of both TMA and tcgen05.mma is a function of the ABSOLUTE
// shared-memory address bits (measured: any 128-byte-aligned start and any SBO multiple of 128 B read
// back what TMA wrote; the descriptor's base_offset field must stay 0), and it costs nothing
// (tools/microbench/umma_rate.cu: 40/48/64/128 cycles per 128xNx8 MMA for N = 32/64/128/256 either way).
// Up to 4 M-tiles share the halo and every weight tile, so weights are re-read 2-4x less often too.
//
// Persistent: one CTA per SM walks work items (patch, Cout slice) round-robin.  Accumulators are
// double-buffered in TMEM (2 x ntiles x BN columns) so the fused epilogue of item i runs under the
// MMAs of item i+1.  Roles (384 threads): warp 0 = halo (A) producer, warp 1 = MMA issuer,
// warp 2 = weight (B) producer, warp 3 = TMEM allocator, warps 4-11 = epilogue (TMEM lane quarter =
// warp % 4, two warps per quarter split the 16-column chunks).  The epilogue is software-pipelined:
// the tcgen05.ld and the global aux loads (bias, h, u, c, residual) of chunk k+1 are in flight while
// chunk k is computed and stored.
// ================================================================================================
// Optional generalisation of a stride-1 halo launch: a kh x kw tap set anchored at (lo_y, lo_x) and an output written
// through a strided view (used by the sub-pixel data gradient of stride-2 convolutions: one launch per input parity).
struct RectSpec {
    int kh, kw, lo_y, lo_x;
    int out_sy, out_sx, out_oy, out_ox, out_H, out_W;
};

// Up-conv mode (RAMNET_FLAG_UPCONV): bilinear x2 (align_corners=False) + 5x5 conv as ONE 5x5 convolution over the
// LOW-resolution tensor with 4 * Cout output columns (phase-major [py][px][co], weights collapsed at pack time) and a
// depth-to-space store.  Zero padding applies in the HIGH-resolution domain and the bilinear clamp replicates the edge,
// so near the image border the collapsed weights differ; the difference only involves the first / last row and column
// of the input and is added by extra K segments that read edge-only VIEWS of the same tensor (a tensor map whose
// bounds are that row / column / corner pixel: everything else is TMA zero fill) with their own collapsed taps.
constexpr int kMaxUpSeg = 9;      // main, top, bottom, left, right, 4 corners
struct UpSeg {
    int tap0;                // first tap of this segment in the packed weight tensor
    int ntaps;               // taps (multiple of tpg; corners are padded with zero taps)
    int kh, kw, lo_y, lo_x;  // tap window: tap i -> (lo_y + i / kw, lo_x + i % kw) relative to the output pixel
    int vx, vy;              // origin of the view in full-tensor coordinates
    int cond;                // border bits a patch / tile must touch: 1 top, 2 bottom, 4 left, 8 right
};
struct UpMaps { CUtensorMap m[kMaxUpSeg - 1]; };

struct HaloGeom {
    int N, H, W, Cout;       // input height / width
    int Ho, Wo;              // output height / width
    int C0, C1;
    int ks, pad, stride;
    int kh, kw;              // taps along y / x (ks x ks unless a RectSpec narrows a stride-1 launch)
    int lo_y, lo_x;          // stride 1: offset of tap (0, 0) relative to the output pixel (-pad for a centred filter)
    int out_sy, out_sx, out_oy, out_ox, out_H, out_W;   // output pixel (oy, ox) lands at (oy*out_sy + out_oy, ox*out_sx + out_ox)
    int nplanes;             // 1 (stride 1) or 4 (stride 2: input parity planes (py, px))
    int plane_stride;        // bytes between parity planes inside one halo stage (multiple of 1024)
    int lo;                  // first halo row/column relative to the patch origin (floor(-pad/stride))
    int PTX, PTY, ptx_log2;  // M-tiles per patch along x / y (tile = 8 x 16 pixels); PTX is a power of two
    int HX, HY;              // halo buffer pitch (pixels) and rows
    int patches_x, patches_y;
    int BN, n_slices;        // GEMM columns per item, Cout / BN
    int a_stages, b_stages;
    int tpg;                 // filter taps per weight stage: one TMA box / one barrier round trip / one commit per tpg taps
    int hpack, cs, kwp;      // hpack: the kwp horizontal taps are GEMM columns (N = kwp * cs) of kh row-shifted MMAs over a
                             // 32 px x 4 row tile; the epilogue sums the taps across the lanes of an image row (see fill_hpack)
    int nbuf;                // TMEM accumulator buffers
    int up;                  // 1: up-conv mode; Cout is then 4 * up_cout GEMM columns
    int up_cout;             // output channels of the up-conv (columns per phase)
    int nseg;                // K segments (1 + border views)
    int total_taps;          // taps of all segments
    UpSeg seg[kMaxUpSeg];
    int pair;                // 1: CTA-pair mode (cta_group::2, M = 256): a work item is two patches x one BN-wide slice
    int items;               // patches (pair mode: patch pairs) * n_slices
    unsigned long long *prof; // RAMNET_PROF=1: per-role wait-cycle counters (debug), else nullptr
    int *sched;              // RAMNET_FLAG_DYNAMIC: {next-item counter, finished workers} in global memory (both 0 at launch,
                             // reset by the last worker); nullptr = static round-robin item assignment
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int kEpiWarps = RAMNET_EPI_WARPS;          // epilogue warps (multiple of 4: one or more per TMEM lane quarter)
constexpr int kHaloThreads = 128 + 32 * kEpiWarps;

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// The "+r" operands make every later use of the loaded registers depend on the wait.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

__device__ __forceinline__ void mbar_wait_t(uint64_t *bar, uint32_t parity, bool prof, unsigned long long &acc) {
    if (prof) {
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        acc += (unsigned long long)(clock64() - t0);
    } else {
        mbar_wait(bar, parity);
    }
}

// ---- CTA-pair (cta_group::2) variants: the two CTAs of a cluster run ONE M=256 UMMA per instruction.  Each CTA owns
// 128 accumulator rows (its own pixel tiles) and stages half of the weight tile; the tensor cores of both SMs read
// both halves, so the weight bytes each SM pulls out of its shared memory per MMA are halved.  Only the leader
// (cluster rank 0) issues MMAs; every TMA load of either CTA reports to the LEADER's full barrier, and every
// tcgen05.commit is multicast to the barrier at the same offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t leader_smem_addr(const void *p) {     // shared::cluster address of p in CTA rank 0
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(p)));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <bool PAIR>
__device__ __forceinline__ void halo_tma_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    if constexpr (PAIR) {
        asm volatile(
            "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(dst)), "l"(map), "r"(leader_smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2)
            : "memory");
    } else {
        tma_load_3d(dst, map, bar, c0, c1, c2);
    }
}
template <bool PAIR>
__device__ __forceinline__ void halo_tma_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    if constexpr (PAIR) {
        asm volatile(
            "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(smem_u32(dst)), "l"(map), "r"(leader_smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
            : "memory");
    } else {
        tma_load_4d(dst, map, bar, c0, c1, c2, c3);
    }
}
template <bool PAIR>
__device__ __forceinline__ void halo_tma_5d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    if constexpr (PAIR) {
        asm volatile(
            "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
            "[%2];" ::"r"(smem_u32(dst)), "l"(map), "r"(leader_smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
            : "memory");
    } else {
        tma_load_5d(dst, map, bar, c0, c1, c2, c3, c4);
    }
}
template <bool PAIR>
__device__ __forceinline__ void halo_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (PAIR) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
    }
}
template <bool PAIR>
__device__ __forceinline__ void halo_commit(uint64_t *bar) {
    if constexpr (PAIR) {
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            ::"r"(smem_u32(bar)), "h"((uint16_t)3)
            : "memory");
    } else {
        umma_commit(bar);
    }
}
// epilogue -> MMA issuer hand-back of an accumulator buffer: the issuer lives in the leader CTA
// RELAXED arrive: what the MMA warp must not overtake are this warp's tcgen05.ld reads of the accumulator, and those
// have completed (tcgen05.wait::ld returned their data) before the arrive in program order.  The default / .release
// forms order ALL of the warp's prior memory operations first -- SASS: MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in pair mode --
// i.e. they wait until every global store of the item's epilogue has been acknowledged by L2 before the accumulator
// buffer is handed back (ncu source view, round 2: ~9 % of all warp-stall samples of the gru0 kernel sat on that
// sequence).  The stores need no ordering with the next item's MMAs.
template <bool PAIR>
__device__ __forceinline__ void halo_arrive_leader(uint64_t *bar) {
    if constexpr (PAIR) {
        asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_smem_addr(bar)) : "memory");
    } else {
        asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
}

// ---- dynamic work distribution (RAMNET_FLAG_DYNAMIC) ------------------------------------------------------------------
// Static assignment (item = worker, worker + nworkers, ...) gives every worker the same share whenever it starts.  When two
// streams overlap (engine.GraphRunner), a kernel's CTAs start at different times -- as the SMs of the other stream's kernel
// free up -- so the early ones finish early and idle.  Here warp 3 of the (leader) CTA, idle after the TMEM allocation,
// draws items from a global counter and feeds them to the roles through a 4-slot shared-memory ring (one full / one empty
// mbarrier per slot; in pair mode the leader also writes the peer's ring and the peer's warps release slots on the
// leader's barriers).  The first item stays static (no start-up latency); the counter pair resets itself.
// MEASURED NEGATIVE (profiles/r02_dynamic_items.txt): bit-identical and ~1 % slower per kernel in isolation, but the
// two-stream forward step got SLOWER with it (12.6 -> 13.0 ms back graphs only, 13.8 front only, 13.45 both), so
// engine.GraphRunner leaves it off (RAMNET_DYNAMIC=1 / front / back turns it on for A/B runs).
constexpr int kSchedRing = 4;
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && ++spins > kSpinLimit) __trap();
    }
}
template <bool PAIR>
__device__ __forceinline__ void sched_publish(int *ring, uint64_t *full, int s, int item) {
    asm volatile("st.volatile.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(ring + s)), "r"(item) : "memory");
    mbar_arrive(full + s);
    if constexpr (PAIR) {
        uint32_t pr, pf;
        asm volatile("mapa.shared::cluster.u32 %0, %1, 1;" : "=r"(pr) : "r"(smem_u32(ring + s)));
        asm volatile("mapa.shared::cluster.u32 %0, %1, 1;" : "=r"(pf) : "r"(smem_u32(full + s)));
        asm volatile("st.volatile.shared::cluster.s32 [%0], %1;" ::"r"(pr), "r"(item) : "memory");
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(pf) : "memory");
    }
}

template <int EPI, bool PAIR, bool HPACK = false, bool UP = false>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_tcgen05_halo_kernel(const __grid_constant__ CUtensorMap map_x0,
                                                                            const __grid_constant__ CUtensorMap map_x1,
                                                                            const __grid_constant__ CUtensorMap map_w,
                                                                            const __grid_constant__ UpMaps upm,
                                                                            const __grid_constant__ HaloGeom g, EpiParams ep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int plane_bytes = g.HX * g.HY * kChunk * 4;
    const int a_bytes = g.nplanes * plane_bytes;             // TMA transaction bytes per halo stage
    const int a_stride = g.nplanes * g.plane_stride;
    const int bn_local = PAIR ? g.BN / 2 : g.BN;             // weight rows staged by this CTA
    const int b_tile = bn_local * kChunk * 4;                // one tap's weight tile
    const int b_bytes = g.tpg * b_tile;                      // one weight stage = tpg consecutive taps
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // persistent worker = CTA or CTA pair
    const int nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    uint8_t *smem_b = smem + (size_t)g.a_stages * a_stride;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem_b + (size_t)g.b_stages * b_bytes);
    uint64_t *a_empty = a_full + g.a_stages;
    uint64_t *b_full = a_empty + g.a_stages;
    uint64_t *b_empty = b_full + g.b_stages;
    uint64_t *acc_full = b_empty + g.b_stages;   // [2]
    uint64_t *acc_empty = acc_full + 2;          // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    uint32_t *tap_tab = tmem_slot + 2;           // [96] halo offset of each filter tap, in 16-byte units
    int *sched_ring = reinterpret_cast<int *>(tap_tab + 96);                           // [kSchedRing] item ids (-1 = no more)
    uint64_t *sched_full = reinterpret_cast<uint64_t *>(sched_ring + kSchedRing);      // [kSchedRing]
    uint64_t *sched_empty = sched_full + kSchedRing;                                   // [kSchedRing]
    const bool dyn = g.sched != nullptr;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (g.prof && threadIdx.x == 0) atomicMax(g.prof + 8, ~globaltimer_ns());        // ~min over CTAs of the entry time
    // Persistent grid (<= one CTA per SM, all resident): let the next kernel of the stream start its prologue on the
    // SMs this grid leaves idle / as its CTAs retire.
    pdl_launch_dependents();
    const int ntiles = g.PTX * g.PTY;
    const int chunks0 = g.C0 / kChunk, chunks = (g.C0 + g.C1) / kChunk;
    const int taps = g.kh * g.kw;
    const int patches = g.patches_x * g.patches_y * g.N;
    const int acc_cols = ntiles * g.BN;          // TMEM columns of one accumulator buffer
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < g.nbuf * acc_cols) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x0);
        prefetch_tmap(&map_x1);
        prefetch_tmap(&map_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < g.a_stages; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
        for (int s = 0; s < g.b_stages; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, kEpiWarps * (PAIR ? 2 : 1)); }
        // a ring slot is released by every role warp that reads it: halo producer, weight producer, epilogue warps of
        // each CTA + the leader's MMA warp
        for (int s = 0; s < kSchedRing; ++s) { mbar_init(sched_full + s, 1); mbar_init(sched_empty + s, (2 + kEpiWarps) * (PAIR ? 2 : 1) + 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        // tap (r, s) -> where its A tile starts inside a halo stage.  Stride 1: one plane, shift (r, s).
        // Stride 2: input row 2*oy + (r - pad) = 2*(oy + dy) + py -> parity plane (py, px), shift (dy, dx).
        if constexpr (UP) {
            // every segment's box is anchored at its own tap window, so tap i of a segment sits (i / kw, i % kw) pixels
            // into the halo stage; padding taps (corner segments) point at the origin and carry zero weights
            for (int i = lane; i < g.total_taps; i += 32) {
                int sg = 0;
                while (sg + 1 < g.nseg && i >= g.seg[sg + 1].tap0) ++sg;
                const int loc = i - g.seg[sg].tap0;
                const int r = loc / g.seg[sg].kw, sx = loc % g.seg[sg].kw;
                tap_tab[i] = loc < g.seg[sg].kh * g.seg[sg].kw ? (uint32_t)((r * g.HX + sx) * 8) : 0u;
            }
        } else if (lane < g.kh * g.kw) {
            const int r = lane / g.kw, sx = lane % g.kw;
            int off;
            if (g.stride == 1) {
                off = (r * g.HX + sx) * 8;
            } else {
                const int ry = r - g.pad, rx = sx - g.pad;
                const int dy = ry >> 1, py = ry & 1, dx = rx >> 1, px = rx & 1;     // arithmetic shift = floor
                off = (py * 2 + px) * (g.plane_stride >> 4) + ((dy - g.lo) * g.HX + (dx - g.lo)) * 8;
            }
            tap_tab[lane] = (uint32_t)off;
        }
        if constexpr (PAIR) {     // collective over the same warp of both CTAs; both get the same TMEM address
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above touched only kernel parameters, shared memory and TMEM

    const bool prof = g.prof != nullptr;
    unsigned long long w0 = 0, w1 = 0, w2 = 0;
    const long long t_start = clock64();
    if (prof && threadIdx.x == 0) {
        const unsigned long long t = globaltimer_ns();
        atomicMax(g.prof + 9, t);
        atomicMax(g.prof + 12, ~t);
    }

    // item -> (Cout slice, image, patch origin); consecutive items share the weight slice (L2 reuse)
    const int patches_w = PAIR ? (patches + 1) >> 1 : patches;   // patches (pair mode: patch pairs) per Cout slice
    auto decode_r = [&](int item, int rank, int &n0, int &img, int &x0, int &y0) {
        const int slice = item / patches_w;
        int t = item - slice * patches_w;
        if constexpr (PAIR) t = 2 * t + rank;                    // an odd patch count leaves img == N: all out of bounds
        const int pxi = t % g.patches_x;
        t /= g.patches_x;
        const int pyi = t % g.patches_y;
        img = t / g.patches_y;
        if constexpr (HPACK) {   // tiles of 32 columns overlap by kwp - 1; the first starts kwp/2 columns left of the image
            x0 = pxi * (32 - (g.kwp - 1)) - (g.kwp >> 1);
            y0 = pyi * g.PTY * 4;
        } else {
            x0 = pxi * g.PTX * 8;
            y0 = pyi * g.PTY * 16;
        }
        n0 = slice * g.BN;
    };
    auto decode = [&](int item, int &n0, int &img, int &x0, int &y0) { decode_r(item, (int)cta_rank, n0, img, x0, y0); };
    // up-conv: which image borders a rectangle touches (bit 1 top, 2 bottom, 4 left, 8 right): the rows / columns whose
    // outputs need the border segments are the first two and the last two of the low-resolution image
    auto border_bits = [&](int x0, int y0, int w, int h) {
        int f = 0;
        if (y0 == 0) f |= 1;
        if (y0 + h > g.H - 2 && y0 < g.H) f |= 2;
        if (x0 == 0) f |= 4;
        if (x0 + w > g.W - 2 && x0 < g.W) f |= 8;
        return f;
    };
    // Border bits of a work item's patch and of each of its tiles.  In pair mode ONE instruction stream serves both
    // CTAs, so every role of both CTAs uses the union over the two patches; the CTA whose patch does not touch that
    // border reads nothing but TMA zero fill from the edge view, i.e. contributes zero.
    auto item_bits = [&](int item, int &pf, int (&tf)[4]) {
        pf = 0;
#pragma unroll
        for (int tl = 0; tl < 4; ++tl) tf[tl] = 0;
        for (int rk = 0; rk < (PAIR ? 2 : 1); ++rk) {
            int n0, img, x0, y0;
            decode_r(item, rk, n0, img, x0, y0);
            if (img >= g.N) continue;
            pf |= border_bits(x0, y0, g.PTX * 8, g.PTY * 16);
#pragma unroll
            for (int tl = 0; tl < 4; ++tl)
                if (tl < g.PTX * g.PTY)
                    tf[tl] |= border_bits(x0 + (tl & (g.PTX - 1)) * 8, y0 + (tl >> g.ptx_log2) * 16, 8, 16);
        }
    };
    const int nseg = UP ? g.nseg : 1;

    // next work item of this worker (-1: none): static round robin, or the scheduler's ring (every lane of the calling warp
    // reads the slot, one lane releases it)
    auto feed = [&](int &fli, int cur) -> int {
        if (!dyn) {
            const int nx = cur < 0 ? worker : cur + nworkers;
            return nx < g.items ? nx : -1;
        }
        const int sl = fli & (kSchedRing - 1);
        if constexpr (PAIR) mbar_wait_cluster(sched_full + sl, (uint32_t)(fli / kSchedRing) & 1u);
        else mbar_wait(sched_full + sl, (uint32_t)(fli / kSchedRing) & 1u);
        int it;
        asm volatile("ld.volatile.shared::cta.s32 %0, [%1];" : "=r"(it) : "r"(smem_u32(sched_ring + sl)) : "memory");
        __syncwarp();
        if (elect_one()) halo_arrive_leader<PAIR>(sched_empty + sl);
        __syncwarp();
        ++fli;
        return it;
    };

    if (warp == 3 && dyn && leader) {
        // ---------------- scheduler: draws items and publishes them kSchedRing ahead of the slowest role ----------------
        if (lane == 0) {
            for (int li = 0;; ++li) {
                const int sl = li & (kSchedRing - 1);
                mbar_wait(sched_empty + sl, ((uint32_t)(li / kSchedRing) & 1u) ^ 1u);
                int item = li == 0 ? worker : nworkers + atomicAdd(g.sched, 1);
                if (item >= g.items) item = -1;
                sched_publish<PAIR>(sched_ring, sched_full, sl, item);
                if (item < 0) break;
            }
            __threadfence();
            if (atomicAdd(g.sched + 1, 1) == nworkers - 1) {     // every worker has drawn its last item: reset for the next launch
                g.sched[0] = 0;
                g.sched[1] = 0;
                __threadfence();
            }
        }
    } else if (warp == 0) {
        // ---------------- halo producer: one box per (item, 32-channel chunk) ----------------
        int stage = 0, fli = 0;
        uint32_t phase = 0;
        for (int item = feed(fli, -1); item >= 0; item = feed(fli, item)) {
            int n0, img, x0, y0;
            decode(item, n0, img, x0, y0);
            int pf = 0, tf[4];
            if constexpr (UP) item_bits(item, pf, tf);
            for (int sg = 0; sg < nseg; ++sg) {
            if (UP && sg > 0 && (g.seg[sg].cond & ~pf)) continue;
            for (int ch = 0; ch < chunks; ++ch) {
                mbar_wait_t(a_empty + stage, phase ^ 1, prof, w0);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(a_full + stage, (uint32_t)a_bytes * (PAIR ? 2u : 1u));
                    const bool second = ch >= chunks0;
                    const CUtensorMap *mx = second ? &map_x1 : &map_x0;
                    const int c = (second ? ch - chunks0 : ch) * kChunk;
                    uint8_t *dst = smem + (size_t)stage * a_stride;
                    if constexpr (UP) {
                        const UpSeg &us = g.seg[sg];
                        halo_tma_4d<PAIR>(dst, sg == 0 ? &map_x0 : &upm.m[sg - 1], a_full + stage, c, x0 + us.lo_x - us.vx,
                                          y0 + us.lo_y - us.vy, img);
                    } else if (g.stride == 1) {
                        halo_tma_4d<PAIR>(dst, mx, a_full + stage, c, x0 + g.lo_x, y0 + g.lo_y, img);
                    } else {
                        const int Csrc = second ? g.C1 : g.C0;
#pragma unroll
                        for (int pl = 0; pl < 4; ++pl)   // (py, px) parity planes of the 5-D view (2C, W/2, 2, H/2, N)
                            halo_tma_5d<PAIR>(dst + (size_t)pl * g.plane_stride, mx, a_full + stage, (pl & 1) * Csrc + c,
                                              x0 + g.lo, pl >> 1, y0 + g.lo, img);
                    }
                }
                __syncwarp();
                if (++stage == g.a_stages) { stage = 0; phase ^= 1; }
            }
            }
        }
        if (prof && lane == 0) atomicAdd(g.prof + 4, w0);
    } else if (warp == 2) {
        // ---------------- weight producer: one [BN x 32] tile per (item, chunk, tap) ----------------
        int stage = 0, fli = 0;
        uint32_t phase = 0;
        for (int item = feed(fli, -1); item >= 0; item = feed(fli, item)) {
            const int n0 = (item / patches_w) * g.BN + (int)cta_rank * bn_local;   // pair mode: this CTA's half of the slice
            int pf = 0, tf[4];
            if constexpr (UP) item_bits(item, pf, tf);
            for (int sg = 0; sg < nseg; ++sg) {
            if (UP && sg > 0 && (g.seg[sg].cond & ~pf)) continue;
            const int seg_taps = UP ? g.seg[sg].ntaps : taps, seg_tap0 = UP ? g.seg[sg].tap0 : 0;
            for (int ch = 0; ch < chunks; ++ch) {
                for (int tap = 0; tap < seg_taps; tap += g.tpg) {     // one box = tpg taps x bn_local rows x 32 channels
                    mbar_wait_t(b_empty + stage, phase ^ 1, prof, w0);
                    if (elect_one()) {
                        if (leader) mbar_expect_tx(b_full + stage, (uint32_t)b_bytes * (PAIR ? 2u : 1u));
                        halo_tma_3d<PAIR>(smem_b + (size_t)stage * b_bytes, &map_w, b_full + stage, ch * kChunk, n0,
                                          seg_tap0 + tap);
                    }
                    __syncwarp();
                    if (++stage == g.b_stages) { stage = 0; phase ^= 1; }
                }
            }
            }
        }
        if (prof && lane == 0) atomicAdd(g.prof + 5, w0);
    } else if (warp == 1) {
      if (leader) {
        // ---------------- MMA issuer: warp-uniform loops, no divisions, one elected lane issues ----------------
        const uint32_t idesc = make_idesc_tf32_m(g.BN, PAIR ? 256 : 128);
        // descriptor bits above the 14-bit start-address field: LBO=16 B, SBO = one halo row, version 1, SWIZZLE_128B
        // hpack: an M tile is 4 image rows of 32 pixels = 16 consecutive 8-pixel groups of the dense 32-wide box (SBO = 1 KB)
        const uint32_t sbo = HPACK ? 1024u : (uint32_t)g.HX * 128u;
        const uint64_t a_hi = ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        uint32_t tile_off16[4];   // (tile origin inside the halo buffer) / 16 bytes
#pragma unroll
        for (int tl = 0; tl < 4; ++tl) {
            const int tx = tl & (g.PTX - 1), ty = tl >> g.ptx_log2;
            tile_off16[tl] = HPACK ? (uint32_t)(tl * 4 * g.HX * 8) : (uint32_t)((ty * 16 * g.HX + tx * 8) * 8);
        }
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        int li = 0, fli = 0;   // local item counter
        int ready = 0;   // weight stages known to be full, starting at sb
        for (int item = feed(fli, -1); item >= 0; item = feed(fli, item), ++li) {
            const int buf = (g.nbuf == 2) ? (li & 1) : 0;
            const uint32_t use = (uint32_t)((g.nbuf == 2) ? (li >> 1) : li);
            mbar_wait_t(acc_empty + buf, (use & 1) ^ 1, prof, w2);          // epilogue has drained this buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc_base = tmem_base + (uint32_t)(buf * acc_cols);
            int pf = 0, tf[4] = {0, 0, 0, 0};
            if constexpr (UP) item_bits(item, pf, tf);
            for (int sg = 0; sg < nseg; ++sg) {
            if (UP && sg > 0 && (g.seg[sg].cond & ~pf)) continue;
            const int seg_taps = UP ? g.seg[sg].ntaps : taps, seg_tap0 = UP ? g.seg[sg].tap0 : 0;
            uint32_t tmask = 0xFu;     // tiles of the patch this segment contributes to (border segments: border tiles only)
            if (UP && sg > 0) {
                tmask = 0;
#pragma unroll
                for (int tl = 0; tl < 4; ++tl)
                    if (!(g.seg[sg].cond & ~tf[tl])) tmask |= 1u << tl;
            }
            for (int ch = 0; ch < chunks; ++ch) {
                mbar_wait_t(a_full + sa, pa, prof, w0);
                const uint32_t a16 = (smem_u32(smem + (size_t)sa * a_stride) & 0x3FFFFu) >> 4;
                for (int tap0 = 0; tap0 < seg_taps; tap0 += g.tpg) {
                    // One barrier round trip (~100+ cycles even when complete) tells us about EVERY weight stage:
                    // lane i polls the stage i slots ahead; the ballot gives the run of ready stages.
                    if (ready == 0) {
                        const long long tw = prof ? clock64() : 0;
                        int j = sb + lane;
                        uint32_t pj = pb;
                        if (j >= g.b_stages) { j -= g.b_stages; pj ^= 1; }
                        uint32_t mask;
                        do {
                            const bool ok = lane < g.b_stages && mbar_test(b_full + j, pj);
                            mask = __ballot_sync(0xffffffffu, ok);
                        } while (!(mask & 1u));
                        ready = __ffs(~mask) - 1;          // consecutive ready stages starting at sb
                        if (prof) w1 += (unsigned long long)(clock64() - tw);
                    }
                    --ready;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t b_base = smem_u32(smem_b + (size_t)sb * b_bytes);
                    if (elect_one()) {
                        // the whole tap group from one thread: no barrier, fence or commit between its taps (measured: the
                        // per-tap round trip cost ~300 cycles in pair mode, more than the MMAs of a 1-tile tap)
                        for (int t = 0; t < g.tpg; ++t) {
                            const int tap = tap0 + t;
                            const uint64_t bdesc = make_smem_desc(b_base + (uint32_t)(t * b_tile));
                            const uint32_t tap16 = a16 + tap_tab[seg_tap0 + tap];
                            const uint32_t first = (uint32_t)(sg | ch | tap);
#pragma unroll
                            for (int tl = 0; tl < 4; ++tl) {
                                if (tl < ntiles && ((tmask >> tl) & 1u)) {
                                    const uint64_t adesc = a_hi | (uint64_t)(tap16 + tile_off16[tl]);
#pragma unroll
                                    for (int kk = 0; kk < kChunk / 8; ++kk)
                                        halo_umma<PAIR>(acc_base + (uint32_t)(tl * g.BN), adesc + 2 * kk, bdesc + 2 * kk,
                                                        idesc, (first | (uint32_t)kk) != 0);
                                }
                            }
                        }
                        halo_commit<PAIR>(b_empty + sb);
                    }
                    __syncwarp();
                    if (++sb == g.b_stages) { sb = 0; pb ^= 1; }
                }
                if (elect_one()) halo_commit<PAIR>(a_empty + sa);
                __syncwarp();
                if (++sa == g.a_stages) { sa = 0; pa ^= 1; }
            }
            }
            if (elect_one()) halo_commit<PAIR>(acc_full + buf);      // every MMA of the item has been issued
            __syncwarp();
        }
        if (prof && lane == 0) {
            atomicAdd(g.prof + 0, (unsigned long long)(clock64() - t_start));
            atomicAdd(g.prof + 1, w0);
            atomicAdd(g.prof + 2, w1);
            atomicAdd(g.prof + 3, w2);
        }
      }
    } else if (warp >= 4) {
        // ---------------- epilogue: software-pipelined TMEM -> registers -> fused math -> global ----------------
        constexpr int kParts = kEpiWarps / 4;     // warps per lane quarter; part p takes chunks c = 16*(p + kParts*j)
        const int quarter = warp & 3, half = (warp - 4) >> 2;
        const int row = quarter * 32 + lane;      // M row: 8 pixels along x per group, 16 groups along y
        const int nch = (g.BN / 16 - half + kParts - 1) / kParts;
        const int units = ntiles * nch;
        int li = 0, fli = 0;
        for (int item = feed(fli, -1); item >= 0; item = feed(fli, item), ++li) {
            int n0, img, x0, y0;
            decode(item, n0, img, x0, y0);
            const int buf = (g.nbuf == 2) ? (li & 1) : 0;
            const uint32_t use = (uint32_t)((g.nbuf == 2) ? (li >> 1) : li);
            mbar_wait_t(acc_full + buf, use & 1, prof, w0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * acc_cols);

            // unit u -> (tile tl, chunk j); walked incrementally
            struct Unit { int tl, j; int64_t m; bool valid; int col; int nn; };
            auto make_unit = [&](int tl, int j) {
                Unit u;
                u.tl = tl; u.j = j;
                const int ox = x0 + (tl & (g.PTX - 1)) * 8 + (row & 7), oy = y0 + (tl >> g.ptx_log2) * 16 + (row >> 3);
                u.valid = oy < g.Ho && ox < g.Wo && img < g.N;
                u.col = (half + kParts * j) * 16;
                if (UP && g.up) {
                    // depth-to-space: GEMM column (phase, co) of low-resolution pixel (oy, ox) is channel co of the
                    // high-resolution pixel (2 oy + py, 2 ox + px); a 16-column unit never straddles two phases
                    const int n = n0 + u.col, phase = n / g.up_cout;
                    u.nn = n - phase * g.up_cout;
                    u.m = ((int64_t)img * (2 * g.Ho) + 2 * oy + (phase >> 1)) * (2 * g.Wo) + 2 * ox + (phase & 1);
                } else {
                    u.nn = n0 + u.col;
                    u.m = ((int64_t)img * g.out_H + oy * g.out_sy + g.out_oy) * g.out_W + ox * g.out_sx + g.out_ox;
                }
                return u;
            };
            auto next_unit = [&](const Unit &u) { return (u.j + 1 < nch) ? make_unit(u.tl, u.j + 1) : make_unit(u.tl + 1, 0); };

            if constexpr (HPACK) {
                // thread = output pixel (image row `quarter` of the tile, column `lane`); accumulator column s * cs + co holds
                // tap s of channel co for the pixel of the SAME lane: the output is the sum over s of lane + s - kwp/2
                if constexpr (EPI == RAMNET_EPI_BIAS || EPI == RAMNET_EPI_BIAS_RELU || EPI == RAMNET_EPI_BIAS_RES_RELU ||
                              EPI == RAMNET_EPI_BIAS_RELU_PRED) {
                    constexpr bool kPred = EPI == RAMNET_EPI_BIAS_RELU_PRED;
                    const int padx = g.kwp >> 1, cs = g.cs;
                    const int nch0 = (n0 / g.BN) * cs;                      // first output channel of this slice
                    if (!kPred || half == 0) {
                        for (int tl = 0; tl < ntiles; ++tl) {
                            const int ox = x0 + lane, oy = y0 + tl * 4 + quarter;
                            const bool valid = lane >= padx && lane < 32 - padx && ox >= 0 && ox < g.Wo && oy < g.Ho && img < g.N;
                            const int64_t m = ((int64_t)img * g.Ho + oy) * g.Wo + ox;
                            float dot = 0.f;
                            for (int c0 = kPred ? 0 : half * 16; c0 < cs; c0 += kPred ? 16 : 16 * kParts) {
                                float acc[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) acc[j] = 0.f;
                                for (int sx = 0; sx < g.kwp; ++sx) {
                                    uint32_t r[16];
                                    tmem_ld16_issue(lane_base + (uint32_t)(tl * g.BN + sx * cs + c0), r);
                                    tmem_ld_wait(r);
                                    const int src = (lane + sx - padx) & 31;   // out-of-tile sources only feed invalid lanes
#pragma unroll
                                    for (int j = 0; j < 16; ++j) acc[j] += __shfl_sync(0xffffffffu, __uint_as_float(r[j]), src);
                                }
                                if constexpr (kPred) {
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const float4 bb = ep.bias ? __ldg(reinterpret_cast<const float4 *>(ep.bias + nch0 + c0) + q)
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                                        const float4 ww = __ldg(reinterpret_cast<const float4 *>(ep.aux0 + nch0 + c0) + q);
                                        dot = fmaf(fmaxf(acc[4 * q] + bb.x, 0.f), ww.x, dot);
                                        dot = fmaf(fmaxf(acc[4 * q + 1] + bb.y, 0.f), ww.y, dot);
                                        dot = fmaf(fmaxf(acc[4 * q + 2] + bb.z, 0.f), ww.z, dot);
                                        dot = fmaf(fmaxf(acc[4 * q + 3] + bb.w, 0.f), ww.w, dot);
                                    }
                                } else if (valid) {
                                    EpiAux<16> xa;
                                    epilogue_prefetch<EPI, 16>(ep, m, nch0 + c0, xa);
                                    epilogue_finish<EPI, 16, true>(ep, m, nch0 + c0, acc, xa);
                                }
                            }
                            if constexpr (kPred) {
                                if (valid) {
                                    const float logit = dot + __ldg(ep.aux1);
                                    if (ep.y1) ep.y1[m] = logit;
                                    ep.y0[m] = sigmoid_t<true>(logit);
                                }
                            }
                        }
                    }
                }
            } else if constexpr (EPI == RAMNET_EPI_BIAS_RELU_PRED && UP) {
                // up-conv + fused 1x1 prediction head: the BN = 4 * Cout columns of a low-resolution pixel are its four
                // high-resolution phases; each phase's Cout activations are reduced to one depth value in registers
                const int nph = kParts >= 2 ? 2 : 4;        // two warps per lane quarter split the phases
                if (half < 2) {
                    for (int tl = 0; tl < ntiles; ++tl) {
                        const int ox = x0 + (tl & (g.PTX - 1)) * 8 + (row & 7), oy = y0 + (tl >> g.ptx_log2) * 16 + (row >> 3);
                        for (int ph = half * nph; ph < 4 && ph < half * nph + nph; ++ph) {
                            float dot = 0.f;
                            for (int c = 0; c < g.up_cout; c += 16) {
                                uint32_t r[16];
                                tmem_ld16_issue(lane_base + (uint32_t)(tl * g.BN + ph * g.up_cout + c), r);
                                tmem_ld_wait(r);
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float4 bb = ep.bias ? __ldg(reinterpret_cast<const float4 *>(ep.bias + c) + q)
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                                    const float4 ww = __ldg(reinterpret_cast<const float4 *>(ep.aux0 + c) + q);
                                    dot = fmaf(fmaxf(__uint_as_float(r[4 * q]) + bb.x, 0.f), ww.x, dot);
                                    dot = fmaf(fmaxf(__uint_as_float(r[4 * q + 1]) + bb.y, 0.f), ww.y, dot);
                                    dot = fmaf(fmaxf(__uint_as_float(r[4 * q + 2]) + bb.z, 0.f), ww.z, dot);
                                    dot = fmaf(fmaxf(__uint_as_float(r[4 * q + 3]) + bb.w, 0.f), ww.w, dot);
                                }
                            }
                            if (oy < g.Ho && ox < g.Wo && img < g.N) {
                                const int64_t m = ((int64_t)img * (2 * g.Ho) + 2 * oy + (ph >> 1)) * (2 * g.Wo) + 2 * ox + (ph & 1);
                                const float logit = dot + __ldg(ep.aux1);
                                if (ep.y1) ep.y1[m] = logit;
                                ep.y0[m] = sigmoid_t<true>(logit);
                            }
                        }
                    }
                }
            } else if constexpr (EPI == RAMNET_EPI_BIAS_RELU_PRED) {
                // fused 1x1 prediction head: every row's BN = Cout activations are reduced in registers by the
                // first warp of each lane quarter; the 32-channel decoder output never leaves the SM
                if (half == 0) {
                    for (int tl = 0; tl < ntiles; ++tl) {
                        const int ox = x0 + (tl & (g.PTX - 1)) * 8 + (row & 7), oy = y0 + (tl >> g.ptx_log2) * 16 + (row >> 3);
                        float dot = 0.f;
                        for (int c = 0; c < g.BN; c += 16) {
                            uint32_t r[16];
                            tmem_ld16_issue(lane_base + (uint32_t)(tl * g.BN + c), r);
                            tmem_ld_wait(r);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 bb = ep.bias ? __ldg(reinterpret_cast<const float4 *>(ep.bias + c) + q)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                                const float4 ww = __ldg(reinterpret_cast<const float4 *>(ep.aux0 + c) + q);
                                dot = fmaf(fmaxf(__uint_as_float(r[4 * q]) + bb.x, 0.f), ww.x, dot);
                                dot = fmaf(fmaxf(__uint_as_float(r[4 * q + 1]) + bb.y, 0.f), ww.y, dot);
                                dot = fmaf(fmaxf(__uint_as_float(r[4 * q + 2]) + bb.z, 0.f), ww.z, dot);
                                dot = fmaf(fmaxf(__uint_as_float(r[4 * q + 3]) + bb.w, 0.f), ww.w, dot);
                            }
                        }
                        if (oy < g.Ho && ox < g.Wo && img < g.N) {
                            const int64_t m = ((int64_t)img * g.out_H + oy * g.out_sy + g.out_oy) * g.out_W + ox * g.out_sx + g.out_ox;
                            const float logit = dot + __ldg(ep.aux1);
                            if (ep.y1) ep.y1[m] = logit;
                            ep.y0[m] = sigmoid_t<true>(logit);
                        }
                    }
                }
            } else {
            uint32_t ra[16], rb[16];
            EpiAux<16> xa, xb;
            if (units > 0) {
                Unit ua = make_unit(0, 0), ub = ua;
                tmem_ld16_issue(lane_base + (uint32_t)(ua.tl * g.BN + ua.col), ra);
                if (ua.valid) epilogue_prefetch<EPI, 16>(ep, ua.m, ua.nn, xa);
                for (int u = 0; u < units; u += 2) {
                    tmem_ld_wait(ra);
                    if (u + 1 < units) {
                        ub = next_unit(ua);
                        tmem_ld16_issue(lane_base + (uint32_t)(ub.tl * g.BN + ub.col), rb);
                        if (ub.valid) epilogue_prefetch<EPI, 16>(ep, ub.m, ub.nn, xb);
                    }
                    if (ua.valid) {
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(ra[i]);
                        epilogue_finish<EPI, 16, true>(ep, ua.m, ua.nn, v, xa);
                    }
                    if (u + 1 < units) {
                        tmem_ld_wait(rb);
                        if (u + 2 < units) {
                            ua = next_unit(ub);
                            tmem_ld16_issue(lane_base + (uint32_t)(ua.tl * g.BN + ua.col), ra);
                            if (ua.valid) epilogue_prefetch<EPI, 16>(ep, ua.m, ua.nn, xa);
                        }
                        if (ub.valid) {
                            float v[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rb[i]);
                            epilogue_finish<EPI, 16, true>(ep, ub.m, ub.nn, v, xb);
                        }
                    }
                }
            }
            }
            // all tcgen05.ld of this warp have completed (waited above): hand the buffer back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (elect_one()) halo_arrive_leader<PAIR>(acc_empty + buf);
            __syncwarp();
        }
        if (prof && lane == 0) {
            atomicAdd(g.prof + 6, w0);
            atomicAdd(g.prof + 7, (unsigned long long)(clock64() - t_start));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (prof && threadIdx.x == 0) {
        const unsigned long long t = globaltimer_ns();
        atomicMax(g.prof + 10, t);
        atomicMax(g.prof + 13, ~t);
    }
    if constexpr (PAIR) cluster_sync_all();   // neither CTA retires while the other may still read its smem / signal its barriers
    if (prof && threadIdx.x == 0) atomicMax(g.prof + 11, globaltimer_ns());
    if (warp == 3) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if constexpr (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}


// ================================================================================================
// Weight gradient on the tensor cores:  dW[tap][co][ci] += sum_pixels dZ[p][co] * X[p*stride + tap - pad][ci]
//
// The contraction (K) dimension is the PIXEL axis.  Both operands already sit in HBM as pixel rows of
// 32 contiguous channels, so a TMA box {32 ch, 8 px, 8 rows} lands in shared memory as 64 K-rows of
// 128 bytes: exactly the MN-major, 128B-swizzled UMMA operand layout (32 M/N elements contiguous,
// K-groups of 8 rows = 1024 B).  Wider M/N are further 32-channel boxes LBO bytes apart.  One
// tcgen05.mma (M=128, N<=128, K=8 pixels, a_major = b_major = MN) therefore consumes them directly —
// no transposes anywhere.  The "M" operand is whichever of (dZ, X) has more channels, so small-Cout
// layers do not waste the 128 rows.  CTA = (tap, M block, N block, pixel range): it streams its pixel
// tiles through a TMA ring, accumulates in TMEM and flushes once with fp32 atomics into the
// nn.Conv2d-layout gradient buffer (BPTT's sum over timesteps is the same += ).
// ================================================================================================
struct WgGeom {
    int N, Ho, Wo;           // output pixel grid (dZ); X boxes are addressed through stride / tap shift
    int Cout, C0, C1;
    int ks, pad, stride;
    int m_from_x;            // 1: M operand = X (input channels), N operand = dZ;  0: M = dZ, N = X
    int Mch, Nch;            // channels of the M / N operand (Cout or C0+C1)
    int BN;                  // N channels per CTA (multiple of 32, <= 128)
    int m_blocks, n_blocks;
    int tiles_x, tiles_y;    // 8 x TR pixel tiles over (Wo, Ho)
    int tiles_per_cta, total_tiles;
    int stages;
    int T;                   // filter taps per CTA: 1 (per-tap mode) or ks (one filter row: the X box carries a halo)
    int TR;                  // image rows per K tile (8 or 4): a tile is 8 pixels x TR rows
    int HXw;                 // width in pixels of the X box: 8 + T - 1
    int a_box, b_box;        // bytes of one 32-channel box of the M / N operand
};


__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;      // leading byte offset: next 32-channel block of M/N
    d |= (uint64_t)(512 >> 4) << 32;            // stride byte offset: next group of 4 K rows (32-byte-atom swizzle)
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                     // SWIZZLE_128B_BASE32B: the only MN-major layout for 32-bit operands
    return d;
}

__global__ void __launch_bounds__(kThreads) conv_wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap map_dz,
                                                                      const __grid_constant__ CUtensorMap map_x0,
                                                                      const __grid_constant__ CUtensorMap map_x1,
                                                                      WgGeom g, float *__restrict__ part) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_boxes = 4, b_boxes = g.BN / kChunk;
    const int stage_bytes = a_boxes * g.a_box + b_boxes * g.b_box;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)g.stages * stage_bytes);
    uint64_t *empty_bar = full_bar + g.stages;
    uint64_t *accum_bar = empty_bar + g.stages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // blockIdx.y -> (tap, m block, n block)
    int t = blockIdx.y;
    const int nb = t % g.n_blocks;
    t /= g.n_blocks;
    const int mb = t % g.m_blocks;
    const int tap = t / g.m_blocks;                    // per-tap mode: tap index; row mode: filter row
    const int r = g.T == 1 ? tap / g.ks : tap, sx = g.T == 1 ? tap % g.ks : 0;
    const int tile_begin = blockIdx.x * g.tiles_per_cta;
    const int tile_end = min(tile_begin + g.tiles_per_cta, g.total_tiles);
    const int ntile = tile_end - tile_begin;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < g.T * g.BN) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_dz);
        prefetch_tmap(&map_x0);
        prefetch_tmap(&map_x1);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < g.stages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        int stage = 0;
        uint32_t phase = 0;
        const int ry = r - g.pad, rx = sx - g.pad;
        for (int it = 0; it < ntile; ++it) {
            int tt = tile_begin + it;
            const int txi = tt % g.tiles_x;
            tt /= g.tiles_x;
            const int tyi = tt % g.tiles_y;
            const int img = tt / g.tiles_y;
            const int ox0 = txi * 8, oy0 = tyi * g.TR;
            mbar_wait(empty_bar + stage, phase ^ 1);
            if (elect_one()) {
                uint8_t *sa = smem + (size_t)stage * stage_bytes;
                uint8_t *sb = sa + a_boxes * g.a_box;
                mbar_expect_tx(full_bar + stage, (uint32_t)stage_bytes);
                auto load_dz = [&](uint8_t *dst, int ch) {       // ch beyond Cout: zero-filled by TMA
                    tma_load_4d(dst, &map_dz, full_bar + stage, ch, ox0, oy0, img);
                };
                auto load_x = [&](uint8_t *dst, int ch) {        // ch indexes the virtual concat [x0 | x1]
                    const bool second = ch >= g.C0 && g.C1 > 0;
                    const CUtensorMap *mx = second ? &map_x1 : &map_x0;
                    const int c = second ? ch - g.C0 : ch;
                    const int Csrc = second ? g.C1 : g.C0;
                    if (ch >= g.C0 + g.C1) {                     // past the last channel: any OOB channel coordinate
                        if (g.stride == 1) tma_load_4d(dst, mx, full_bar + stage, Csrc, ox0 + rx, oy0 + ry, img);
                        else tma_load_5d(dst, mx, full_bar + stage, 2 * Csrc, ox0 + (rx >> 1), ry & 1, oy0 + (ry >> 1), img);
                    } else if (g.stride == 1) {
                        tma_load_4d(dst, mx, full_bar + stage, c, ox0 + rx, oy0 + ry, img);
                    } else {
                        tma_load_5d(dst, mx, full_bar + stage, (rx & 1) * Csrc + c, ox0 + (rx >> 1), ry & 1, oy0 + (ry >> 1), img);
                    }
                };
                for (int q = 0; q < a_boxes; ++q) {
                    const int ch = (mb * 4 + q) * kChunk;
                    if (g.m_from_x) load_x(sa + q * g.a_box, ch); else load_dz(sa + q * g.a_box, ch);
                }
                for (int q = 0; q < b_boxes; ++q) {
                    const int ch = nb * g.BN + q * kChunk;
                    if (g.m_from_x) load_dz(sb + q * g.b_box, ch); else load_x(sb + q * g.b_box, ch);
                }
            }
            __syncwarp();
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        // idesc: kind::tf32, fp32 accumulate, A and B MN-major (bits 15, 16), M = 128, N = BN
        const uint32_t idesc = make_idesc_tf32(g.BN) | (1u << 15) | (1u << 16);
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < ntile; ++it) {
            mbar_wait(full_bar + stage, phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint64_t adesc = make_smem_desc_mn(sa, (uint32_t)g.a_box);
            const uint64_t bdesc = make_smem_desc_mn(sa + a_boxes * g.a_box, (uint32_t)g.b_box);
            // address-field (>>4) steps: the X box rows are HXw pixels wide (halo in row mode), the dZ box rows 8 pixels
            const uint32_t a_row = (uint32_t)(g.m_from_x ? g.HXw : 8) * 8u, b_row = (uint32_t)(g.m_from_x ? 8 : g.HXw) * 8u;
            const uint32_t a_tap = g.m_from_x ? 8u : 0u, b_tap = g.m_from_x ? 0u : 8u;     // +1 pixel = 128 B per tap
            if (elect_one()) {
                for (int tp = 0; tp < g.T; ++tp) {
#pragma unroll 4
                    for (int kk = 0; kk < g.TR; ++kk)             // one image row of 8 pixels (K = 8) per MMA
                        umma_tf32(tmem_base + (uint32_t)(tp * g.BN), adesc + (uint64_t)(kk * a_row + tp * a_tap),
                                  bdesc + (uint64_t)(kk * b_row + tp * b_tap), idesc, (it | kk) != 0);
                }
                umma_commit(empty_bar + stage);
                if (it == ntile - 1) umma_commit(accum_bar);
            }
            __syncwarp();
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
    } else if (ntile > 0) {
        // ---------------- epilogue: TMEM -> this CTA's partial tile part[split][group][128][BN] (plain stores) ----------------
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int ncols = g.T * g.BN;
        float *dst = part + (((size_t)blockIdx.x * gridDim.y + blockIdx.y) * 128 + row) * ncols;
        for (int c = 0; c < ncols; c += 16) {
            float v[16];
            tmem_ld16(lane_addr + (uint32_t)c, v);
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4 *>(dst + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ================================================================================================
// Tap-packed weight gradient (stride-1 convolutions).
//
// The kernel above spends one MMA per filter tap.  Here the operand that is shifted by the tap (the "N" operand: one
// 32-channel box of X or of dZ carrying a 2-D halo) feeds ALL ks taps of a filter row to one MMA: N-block j of an
// MN-major operand starts LBO bytes after block j-1, so LBO = 128 B (one pixel) makes block j the same box shifted by
// j pixels -> N = ks*32 columns = (tap s, channel c).  The filter row is a descriptor start shifted by whole halo
// rows and owns its own ks*32 accumulator columns.  One CTA = (128 channels of the M operand, one 32-channel box of
// the N operand, a group of <= 3 filter rows, a pixel range): every loaded byte is used for ks * rows taps, the
// MMAs are N = 96 / 160 wide (math-bound instead of shared-memory-bound) and the operand traffic per FLOP drops
// ~3x against the filter-row kernel.
//   m_from_x = 0:  M = dZ tile (no halo),  N = X box at (x0 - pad, y0 - pad + u0);   tap (r, s) = (u, j)
//   m_from_x = 1:  M = X tile (no halo),   N = dZ box at (x0 + pad - (ks-1), y0 + pad - (ks-1) + u0); tap = (ks-1-u, ks-1-j)
// (u = halo row shift, j = N block).  Out-of-bounds pixels of either box are zero-filled by TMA = zero padding.
// ================================================================================================
struct WpGeom {
    int N, H, W;             // pixel grid of dZ (= of X for stride 1, = of each input parity plane for stride 2)
    int Cout, C0, C1;
    int ks;                  // filter size of the convolution (dW is [Cout][Ct][ks][ks])
    int kh, kw;              // taps of THIS problem: ks x ks (stride 1) or the 3x3 / 3x2 / 2x3 / 2x2 sub-filter of one parity class
    int oy, ox;              // origin of the N box relative to the tile origin (before the row-group offset)
    int r0, dr, s0, ds;      // filter tap of (row shift u, N block j): (r0 + dr*u, s0 + ds*j)
    int x5d, py, px;         // stride 2: X is read through the 5-D parity view (2C, W/2, 2, H/2, N), plane (py, px)
    int head_cin;            // > 0: X is ramnet_head_im2row's tensor (channel = dx*Cin + ci) and dW is the head's [Cout][Cin][5][5]
    int mfold;               // 1: the M operand has only 64 channels; MMA rows 64..127 hold the SAME channels read RG image rows
                             // higher, i.e. the filter rows u + RG: one CTA covers 2*RG row shifts and no MMA row is wasted
    int m_from_x;
    int Mch, Nch;            // channels of the M / N operand
    int m_blocks, n_boxes;   // ceil(Mch / 128), Nch / 32
    int RG, row_groups;      // halo row shifts per CTA, ceil(kh / RG)
    int TR;                  // image rows per K tile (tile = 8 px x TR rows)
    int HXw, HYw;            // N-operand box: (8 + kw - 1) px x (TR + RG - 1) rows
    int tiles_x, tiles_y, tiles_per_cta, total_tiles;
    int stages, stage_bytes;
    int ncols;               // accumulator columns per CTA = RG * kw * 32 (workspace row pitch)
    unsigned long long *prof; // RAMNET_PROF=1: cycle counters (debug), else nullptr
};

// The problems of one layer (1, or the 4 parity classes of a stride-2 layer) run as ONE launch: blockIdx.z selects the
// problem.  They share the tensor maps (N boxes sized for the largest sub-filter), the split / group counts and the
// workspace pitch, so the three launches (MMA, split sum, scatter) are paid once per layer.
struct WpBatch {
    WpGeom g[4];
    int n;
    int accumulate;          // 1: the epilogue ADDS this launch's partial tiles to what the workspace holds (BPTT: the same
                             // layer's weight gradient is produced once per pass; the split sum + scatter then runs once per step)
    long long part_stride;   // floats between the workspaces of consecutive problems
};

__global__ void __launch_bounds__(kThreads) conv_wgrad_packed_kernel(const __grid_constant__ CUtensorMap map_dz,
                                                                     const __grid_constant__ CUtensorMap map_x0,
                                                                     const __grid_constant__ CUtensorMap map_x1,
                                                                     const __grid_constant__ WpBatch batch,
                                                                     float *__restrict__ part) {
    const WpGeom &g = batch.g[blockIdx.z];
    part += (size_t)blockIdx.z * batch.part_stride;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int m_box = g.TR * 8 * kChunk * 4;              // one 32-channel box of the M operand (TR KB)
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)g.stages * g.stage_bytes);
    uint64_t *empty_bar = full_bar + g.stages;
    uint64_t *accum_bar = empty_bar + g.stages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // blockIdx.y -> (row group, m block, n box)
    int t = blockIdx.y;
    const int nb = t % g.n_boxes;
    t /= g.n_boxes;
    const int mb = t % g.m_blocks;
    const int rg = t / g.m_blocks;
    const int u0 = rg * g.RG;
    const int rows = min(g.RG, g.kh - u0);                 // halo row shifts handled here
    const int tile_begin = blockIdx.x * g.tiles_per_cta;
    const int tile_end = min(tile_begin + g.tiles_per_cta, g.total_tiles);
    const int ntile = tile_end - tile_begin;
    const int ncb = g.kw * kChunk;                         // accumulator columns per row shift = UMMA N
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < g.ncols) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_dz);
        prefetch_tmap(&map_x0);
        prefetch_tmap(&map_x1);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < g.stages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const bool prof = g.prof != nullptr;
    unsigned long long wcyc = 0;
    const long long t_start = clock64();

    if (warp == 0) {
        // ---------------- TMA producer: 4 M boxes + 1 haloed N box per K tile ----------------
        int stage = 0;
        uint32_t phase = 0;
        const int nbytes = g.HXw * g.HYw * kChunk * 4;     // the box always carries RG - 1 halo rows
        // origin of the N box relative to the tile origin
        const int nx = g.ox, ny = g.oy + u0;
        for (int it = 0; it < ntile; ++it) {
            int tt = tile_begin + it;
            const int txi = tt % g.tiles_x;
            tt /= g.tiles_x;
            const int tyi = tt % g.tiles_y;
            const int img = tt / g.tiles_y;
            const int x0 = txi * 8, y0 = tyi * g.TR;
            mbar_wait_t(empty_bar + stage, phase ^ 1, prof, wcyc);
            if (elect_one()) {
                uint8_t *sa = smem + (size_t)stage * g.stage_bytes;
                uint8_t *sb = sa + 4 * m_box;
                mbar_expect_tx(full_bar + stage, (uint32_t)(4 * m_box + nbytes));
                auto load_x = [&](uint8_t *dst, int ch, int x, int y) {   // ch indexes the virtual concat [x0 | x1]; past its end: zeros
                    const bool second = ch >= g.C0 && g.C1 > 0;
                    const CUtensorMap *mx = second ? &map_x1 : &map_x0;
                    const int Csrc = second ? g.C1 : g.C0;
                    const bool past = ch >= g.C0 + g.C1;
                    const int c = past ? Csrc : (second ? ch - g.C0 : ch);
                    if (g.x5d) tma_load_5d(dst, mx, full_bar + stage, past ? 2 * Csrc : g.px * Csrc + c, x, g.py, y, img);
                    else tma_load_4d(dst, mx, full_bar + stage, c, x, y, img);
                };
                for (int q = 0; q < 4; ++q) {
                    int ch = (mb * 4 + q) * kChunk, ym = y0;
                    if (g.mfold && q >= 2) { ch -= 2 * kChunk; ym -= g.RG; }      // second copy of the 64 channels, RG rows up
                    if (g.m_from_x) load_x(sa + q * m_box, ch, x0, ym);
                    else tma_load_4d(sa + q * m_box, &map_dz, full_bar + stage, ch, x0, ym, img);   // ch >= Cout: zeros
                }
                if (g.m_from_x) tma_load_4d(sb, &map_dz, full_bar + stage, nb * kChunk, x0 + nx, y0 + ny, img);
                else load_x(sb, nb * kChunk, x0 + nx, y0 + ny);
            }
            __syncwarp();
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
        if (prof && lane == 0) atomicAdd(g.prof + 0, wcyc);
    } else if (warp == 1) {
        // ---------------- MMA issuer: rows x TR MMAs of N = ks*32 per K tile ----------------
        const uint32_t idesc = make_idesc_tf32(ncb) | (1u << 15) | (1u << 16);     // A and B MN-major
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < ntile; ++it) {
            mbar_wait_t(full_bar + stage, phase, prof, wcyc);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(smem + (size_t)stage * g.stage_bytes);
            const uint64_t adesc = make_smem_desc_mn(sa, (uint32_t)m_box);          // M blocks: next 32-channel box
            const uint64_t bdesc = make_smem_desc_mn(sa + 4 * m_box, 128u);         // N blocks: the box shifted by one pixel
            if (elect_one()) {
                for (int u = 0; u < rows; ++u) {
#pragma unroll 4
                    for (int kk = 0; kk < g.TR; ++kk)          // one image row of 8 pixels (K = 8) per MMA
                        umma_tf32(tmem_base + (uint32_t)(u * ncb), adesc + (uint64_t)(kk * 64),
                                  bdesc + (uint64_t)((kk + u) * g.HXw * 8), idesc, (it | kk) != 0);
                }
                umma_commit(empty_bar + stage);
                if (it == ntile - 1) umma_commit(accum_bar);
            }
            __syncwarp();
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
        if (prof && lane == 0) {
            atomicAdd(g.prof + 1, wcyc);
            atomicAdd(g.prof + 2, (unsigned long long)(clock64() - t_start));
        }
    } else if (ntile > 0) {
        // ---------------- epilogue: TMEM -> part[split][group][128][ncols] (plain stores) ----------------
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        mbar_wait(accum_bar, 0);
        const long long t_epi = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // column-major tile [col][128 rows]: the 32 lanes of a warp hold 32 consecutive rows of one column -> every store
        // instruction writes one full 128-byte line
        float *dst = part + ((size_t)blockIdx.x * gridDim.y + blockIdx.y) * 128 * g.ncols + row;
        uint32_t ra[16], rb[16];
        const int nchunk = rows * ncb / 16;
        tmem_ld16_issue(lane_addr, ra);
        for (int ci = 0; ci < nchunk; ci += 2) {
            tmem_ld_wait(ra);
            if (ci + 1 < nchunk) tmem_ld16_issue(lane_addr + (uint32_t)((ci + 1) * 16), rb);
            if (batch.accumulate) {      // 16 coalesced 128-byte loads in flight, then add + store
                float old[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) old[j] = dst[(size_t)(ci * 16 + j) * 128];
#pragma unroll
                for (int j = 0; j < 16; ++j) dst[(size_t)(ci * 16 + j) * 128] = old[j] + __uint_as_float(ra[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) dst[(size_t)(ci * 16 + j) * 128] = __uint_as_float(ra[j]);
            }
            if (ci + 1 < nchunk) {
                tmem_ld_wait(rb);
                if (ci + 2 < nchunk) tmem_ld16_issue(lane_addr + (uint32_t)((ci + 2) * 16), ra);
                if (batch.accumulate) {
                    float old[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) old[j] = dst[(size_t)((ci + 1) * 16 + j) * 128];
#pragma unroll
                    for (int j = 0; j < 16; ++j) dst[(size_t)((ci + 1) * 16 + j) * 128] = old[j] + __uint_as_float(rb[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) dst[(size_t)((ci + 1) * 16 + j) * 128] = __uint_as_float(rb[j]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (prof && threadIdx.x == 64) {
            atomicAdd(g.prof + 3, (unsigned long long)(clock64() - t_epi));
            atomicAdd(g.prof + 4, (unsigned long long)(clock64() - t_start));
            atomicAdd(g.prof + 5, 1ull);
        }
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode(ramnet_handle *h, CUtensorMap *map, const void *base, int rank, const cuuint64_t *dims,
           const cuuint64_t *strides_bytes, const cuuint32_t *box,
           CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    cuuint32_t elem[5] = {1, 1, 1, 1, 1};
    CUresult r = ((EncodeTiledFn)h->encode_tiled)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                                                  const_cast<void *>(base), dims, strides_bytes, box, elem,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ramnet_set_error(RAMNET_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return RAMNET_OK;
}

// activation map: stride 1 -> (C, W, H, N); stride 2 -> (2C, W/2, 2, H/2, N)
int encode_activation(ramnet_handle *h, CUtensorMap *map, const float *x, int N, int H, int W, int C, int stride,
                      int TW, int TH, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    if (stride == 1) {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t str[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
        cuuint32_t box[4] = {kChunk, (cuuint32_t)TW, (cuuint32_t)TH, 1};
        return encode(h, map, x, 4, dims, str, box, swizzle);
    }
    cuuint64_t dims[5] = {(cuuint64_t)2 * C, (cuuint64_t)W / 2, 2, (cuuint64_t)H / 2, (cuuint64_t)N};
    cuuint64_t str[4] = {(cuuint64_t)2 * C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)2 * W * C * 4,
                         (cuuint64_t)H * W * C * 4};
    cuuint32_t box[5] = {kChunk, (cuuint32_t)TW, 1, (cuuint32_t)TH, 1};
    return encode(h, map, x, 5, dims, str, box, swizzle);
}

void pick_tile(int Ho, int Wo, int *TW, int *TH) {
    int64_t best = -1;
    for (int tw = 128; tw >= 8; tw >>= 1) {
        const int th = kTileM / tw;
        const int64_t covered = (int64_t)((Wo + tw - 1) / tw) * tw * ((Ho + th - 1) / th) * th;
        if (best < 0 || covered < best) { best = covered; *TW = tw; *TH = th; }
    }
}

template <int EPI>
int launch(ramnet_handle *h, const CUtensorMap &m0, const CUtensorMap &m1, const CUtensorMap &mw, const TcGeom &g,
           const EpiParams &ep, cudaStream_t s) {
    const size_t smem = (size_t)g.stages * (kABytes + (size_t)g.BN * kChunk * 4) + (2 * g.stages + 1) * 8 + 16 + 1024;
    static size_t configured = 0;   // per template instance
    if (smem > configured) {
        RAMNET_CUDA(cudaFuncSetAttribute(conv_tcgen05_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((unsigned)(g.tiles_x * g.tiles_y * g.N), (unsigned)(g.Cout / g.BN));
    conv_tcgen05_kernel<EPI><<<grid, kThreads, smem, s>>>(m0, m1, mw, g, ep);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

const UpMaps &no_up_maps() {
    static UpMaps z = {};
    return z;
}

template <int EPI, bool PAIR, bool HP = false, bool UP = false>
int launch_halo_pair(ramnet_handle *h, const CUtensorMap &m0, const CUtensorMap &m1, const CUtensorMap &mw,
                     const HaloGeom &g, const EpiParams &ep, size_t smem, cudaStream_t s, const UpMaps &um = no_up_maps()) {
    static size_t configured = 0;
    if (smem > configured) {
        RAMNET_CUDA(cudaFuncSetAttribute(conv_tcgen05_halo_kernel<EPI, true, HP, UP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        configured = smem;
    }
    const int pairs = g.items < h->sm_count / 2 ? g.items : h->sm_count / 2;   // persistent: one CTA pair per TPC
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kHaloThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // RAMNET_PDL=1: prologue under the predecessor's tail
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ramnet_pdl_enabled() ? 2 : 1;
    static const bool do_prof = getenv("RAMNET_PROF") != nullptr;      // debug only: synchronises and prints
    if (do_prof) {
        static unsigned long long *buf = nullptr;
        if (!buf) cudaMalloc(&buf, 128);
        cudaMemsetAsync(buf, 0, 128, s);
        HaloGeom gp = g;
        gp.prof = buf;
        RAMNET_CUDA(cudaLaunchKernelEx(&cfg, conv_tcgen05_halo_kernel<EPI, true, HP, UP>, m0, m1, mw, um, gp, ep));
        unsigned long long hbuf[16];
        cudaMemcpyAsync(hbuf, buf, 128, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        {
            const double e0 = (double)~hbuf[8];
            fprintf(stderr, "[ramnet-prof] wall clock (us after the first CTA's entry): setup done %.1f .. %.1f | loops done %.1f .. %.1f | "
                            "last cluster sync %.1f\n", ((double)~hbuf[12] - e0) / 1e3, ((double)hbuf[9] - e0) / 1e3,
                    ((double)~hbuf[13] - e0) / 1e3, ((double)hbuf[10] - e0) / 1e3, ((double)hbuf[11] - e0) / 1e3);
        }
        const double n = pairs;     // MMA counters: one issuer per pair; producer / epilogue counters: two CTAs per pair
        fprintf(stderr, "[ramnet-prof] pairs=%d items=%d per-pair kcycles: mma_total=%.1f wait_a_full=%.1f wait_b_full=%.1f "
                        "wait_acc_empty=%.1f | prodA_wait_empty=%.1f prodB_wait_empty=%.1f | epi(avg of warps) "
                        "wait_acc_full=%.1f total=%.1f\n",
                pairs, g.items, hbuf[0] / n / 1e3, hbuf[1] / n / 1e3, hbuf[2] / n / 1e3, hbuf[3] / n / 1e3, hbuf[4] / n / 2e3,
                hbuf[5] / n / 2e3, hbuf[6] / n / 16e3, hbuf[7] / n / 16e3);
    } else {
        RAMNET_CUDA(cudaLaunchKernelEx(&cfg, conv_tcgen05_halo_kernel<EPI, true, HP, UP>, m0, m1, mw, um, g, ep));
    }
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

// RAMNET_FLAG_DYNAMIC: the {counter, finished} pair of the launch being issued (set by conv_fwd_tf32_rect around its
// dispatch; every tensor-core forward launch goes through launch_halo)
static thread_local int *tl_sched_slot = nullptr;

template <int EPI, bool HP = false, bool UP = false>
int launch_halo(ramnet_handle *h, const CUtensorMap &m0, const CUtensorMap &m1, const CUtensorMap &mw,
                const HaloGeom &g_in, const EpiParams &ep, cudaStream_t s, const UpMaps &um = no_up_maps()) {
    HaloGeom g = g_in;
    const int workers = g.pair ? h->sm_count / 2 : h->sm_count;
    g.sched = (tl_sched_slot && g.items > workers) ? tl_sched_slot : nullptr;     // a single wave has nothing to redistribute
    const size_t a_stride = (size_t)g.nplanes * g.plane_stride;
    const size_t smem = g.a_stages * a_stride + (size_t)g.b_stages * g.tpg * (g.pair ? g.BN / 2 : g.BN) * kChunk * 4 +
                        (2 * g.a_stages + 2 * g.b_stages + 4) * 8 + 16 + 96 * 4 + kSchedRing * 4 + 2 * kSchedRing * 8 + 1024;
    if (g.pair) return launch_halo_pair<EPI, true, HP, UP>(h, m0, m1, mw, g, ep, smem, s, um);
    static size_t configured = 0;
    if (smem > configured) {
        RAMNET_CUDA(cudaFuncSetAttribute(conv_tcgen05_halo_kernel<EPI, false, HP, UP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        configured = smem;
    }
    const int grid = g.items < h->sm_count ? g.items : h->sm_count;   // persistent: one CTA per SM
    static const bool do_prof = getenv("RAMNET_PROF") != nullptr;      // debug only: synchronises and prints
    if (do_prof) {
        static unsigned long long *buf = nullptr;
        if (!buf) cudaMalloc(&buf, 64);
        cudaMemsetAsync(buf, 0, 64, s);
        HaloGeom gp = g;
        gp.prof = buf;
        conv_tcgen05_halo_kernel<EPI, false, HP, UP><<<grid, kHaloThreads, smem, s>>>(m0, m1, mw, um, gp, ep);
        unsigned long long hbuf[8];
        cudaMemcpyAsync(hbuf, buf, 64, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        const double n = grid;
        fprintf(stderr, "[ramnet-prof] grid=%d items=%d per-CTA kcycles: mma_total=%.1f wait_a_full=%.1f wait_b_full=%.1f "
                        "wait_acc_empty=%.1f | prodA_wait_empty=%.1f prodB_wait_empty=%.1f | epi(avg of 8 warps) "
                        "wait_acc_full=%.1f total=%.1f\n",
                grid, g.items, hbuf[0] / n / 1e3, hbuf[1] / n / 1e3, hbuf[2] / n / 1e3, hbuf[3] / n / 1e3, hbuf[4] / n / 1e3,
                hbuf[5] / n / 1e3, hbuf[6] / n / 8e3, hbuf[7] / n / 8e3);
    } else {
        RAMNET_CUDA(ramnet_launch(conv_tcgen05_halo_kernel<EPI, false, HP, UP>, dim3(grid), dim3(kHaloThreads), smem, s, true, m0, m1, mw, um, g, ep));
    }
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

int halo_mode_env() {
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("RAMNET_HALO_MODE");
        mode = e ? atoi(e) : 0;
    }
    return mode;
}

bool fill_halo(const ramnet_conv_desc *d, const RectSpec *rect, HaloGeom *g, int ptx, int pty, int bn, int a_st, int b_st,
               int pair = 0, int max_tpg = 0) {
    if (d->Cout % bn || ptx * pty > 4 || ptx * pty * bn > 512 || (ptx & (ptx - 1))) return false;
    if (pair && (bn % 16 || bn < 32)) return false;          // cta_group::2: N in steps of 16; each CTA stages bn/2 rows
    g->pair = pair;
    g->N = d->N; g->H = d->H; g->W = d->W; g->Cout = d->Cout; g->C0 = d->C0; g->C1 = d->C1;
    g->Ho = conv_out_dim(d->H, d->stride); g->Wo = conv_out_dim(d->W, d->stride);
    g->ks = d->ksize; g->pad = d->ksize / 2; g->stride = d->stride; g->prof = nullptr; g->sched = nullptr;
    g->nplanes = d->stride == 1 ? 1 : 4;
    // halo extent along one axis, in plane pixels: shifts floor((r - pad)/stride) for r = 0..ks-1
    auto fdiv = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
    g->lo = fdiv(-g->pad, d->stride);
    const int hi = fdiv(d->ksize - 1 - g->pad, d->stride);
    g->PTX = ptx; g->PTY = pty; g->ptx_log2 = ptx == 4 ? 2 : (ptx == 2 ? 1 : 0);
    g->HX = ptx * 8 + (hi - g->lo); g->HY = pty * 16 + (hi - g->lo);
    g->hpack = 0; g->cs = 0; g->kwp = 0;
    g->up = 0; g->up_cout = 0; g->nseg = 1; g->total_taps = d->ksize * d->ksize;
    g->kh = g->kw = d->ksize; g->lo_y = g->lo_x = g->lo;
    g->out_sy = g->out_sx = 1; g->out_oy = g->out_ox = 0; g->out_H = g->Ho; g->out_W = g->Wo;
    if (rect) {               // stride-1 launch with a rectangular tap set and / or a strided output view
        if (d->stride != 1) return false;
        g->kh = rect->kh; g->kw = rect->kw; g->lo_y = rect->lo_y; g->lo_x = rect->lo_x;
        g->out_sy = rect->out_sy; g->out_sx = rect->out_sx; g->out_oy = rect->out_oy; g->out_ox = rect->out_ox;
        g->out_H = rect->out_H; g->out_W = rect->out_W;
        g->HX = ptx * 8 + rect->kw - 1; g->HY = pty * 16 + rect->kh - 1;
    }
    if (g->HX > 256 || g->HY > 256) return false;
    g->BN = bn; g->n_slices = d->Cout / bn;
    g->patches_x = (g->Wo + ptx * 8 - 1) / (ptx * 8); g->patches_y = (g->Ho + pty * 16 - 1) / (pty * 16);
    const int64_t npatch = (int64_t)g->patches_x * g->patches_y * d->N;
    const int64_t items = (pair ? (npatch + 1) / 2 : npatch) * g->n_slices;
    if (items > 0x7fffffff) return false;
    g->items = (int)items;
    g->nbuf = (2 * ptx * pty * bn <= 512) ? 2 : 1;
    g->plane_stride = (int)(((size_t)g->HX * g->HY * kChunk * 4 + 1023) & ~(size_t)1023);
    const size_t a_stride = (size_t)g->nplanes * g->plane_stride;
    const size_t b_tile = (size_t)(pair ? bn / 2 : bn) * kChunk * 4;
    const size_t budget = 222 * 1024;
    const bool auto_depth = a_st <= 0;
    if (auto_depth) {
        a_st = 2;
        if (budget < a_st * a_stride + 3 * b_tile) a_st = 1;
        if (budget < a_st * a_stride + 3 * b_tile) return false;
    }
    // taps per weight stage: all of them, one filter row, or one -- the largest that keeps a stage <= 40 KB and
    // leaves room for three stages (RAMNET_HALO_TPG overrides for A/B runs)
    const int taps = g->kh * g->kw;
    static const int tpg_env = [] { const char *e = getenv("RAMNET_HALO_TPG"); return e ? atoi(e) : 0; }();
    const int cand[3] = {taps, g->kw, 1};
    int tpg = 1;
    for (int c : cand) {
        if (c < 1 || taps % c) continue;
        if (tpg_env > 0 && c > tpg_env) continue;
        if (max_tpg > 0 && c > max_tpg) continue;
        if ((size_t)c * b_tile <= 40 * 1024 && a_st * a_stride + 3 * (size_t)c * b_tile <= budget) { tpg = c; break; }
    }
    g->tpg = tpg;
    const size_t b_bytes = (size_t)tpg * b_tile;
    if (auto_depth) {
        b_st = (int)((budget - a_st * a_stride) / b_bytes);
        if (b_st > 10) b_st = 10;
    } else if (a_st * a_stride + b_st * b_bytes > budget) {
        b_st = (int)((budget - a_st * a_stride) / b_bytes);     // forced depths are in single-tap stages: clamp
    }
    if (a_st * a_stride + b_st * b_bytes > budget || b_st < 2) return false;
    g->a_stages = a_st; g->b_stages = b_st;
    return true;
}

// hpack configuration (RAMNET_FLAG_HPACK: weights packed by ramnet_pack_weights_hpack).  For a stride-1 ks x ks conv with
// few output channels the ks horizontal taps become GEMM columns: per 32-channel chunk ks row-shifted MMAs of
// N = ks * cs over a 32 px x 4 row tile instead of ks*ks MMAs of N = cs, and the epilogue adds the taps up across the
// lanes of an image row (warp shuffles).  Index algebra: tests/test_hpack_index_algebra.py.
bool fill_hpack(const ramnet_conv_desc *d, HaloGeom *g, int pty, int cs, int pair) {
    const int ks = d->ksize, bn = ks * cs;
    if (d->stride != 1 || (ks != 3 && ks != 5) || cs % 16 || d->Cout % cs || bn > 256 || pty < 1 || pty > 4) return false;
    if (pair && (bn % 16 || (bn / 2) % 8)) return false;
    if (pty * bn > 512) return false;
    g->pair = pair;
    g->N = d->N; g->H = d->H; g->W = d->W; g->Cout = d->Cout; g->C0 = d->C0; g->C1 = d->C1;
    g->Ho = d->H; g->Wo = d->W;
    g->ks = ks; g->pad = ks / 2; g->stride = 1; g->prof = nullptr; g->sched = nullptr; g->nplanes = 1; g->lo = -g->pad;
    g->hpack = 1; g->cs = cs; g->kwp = ks;
    g->up = 0; g->up_cout = 0; g->nseg = 1; g->total_taps = ks;
    g->kh = ks; g->kw = 1;                      // "taps" of the weight pipeline = filter rows
    g->lo_y = -g->pad; g->lo_x = 0;
    g->out_sy = g->out_sx = 1; g->out_oy = g->out_ox = 0; g->out_H = g->Ho; g->out_W = g->Wo;
    g->PTX = 1; g->PTY = pty; g->ptx_log2 = 0;
    g->HX = 32; g->HY = 4 * pty + ks - 1;
    g->BN = bn; g->n_slices = d->Cout / cs;
    g->patches_x = (g->Wo + (32 - (ks - 1)) - 1) / (32 - (ks - 1));
    g->patches_y = (g->Ho + 4 * pty - 1) / (4 * pty);
    const int64_t npatch = (int64_t)g->patches_x * g->patches_y * d->N;
    const int64_t items = (pair ? (npatch + 1) / 2 : npatch) * g->n_slices;
    if (items > 0x7fffffff) return false;
    g->items = (int)items;
    g->nbuf = (2 * pty * bn <= 512) ? 2 : 1;
    g->plane_stride = (int)(((size_t)g->HX * g->HY * kChunk * 4 + 1023) & ~(size_t)1023);
    const size_t a_stride = g->plane_stride, b_tile = (size_t)(pair ? bn / 2 : bn) * kChunk * 4, budget = 222 * 1024;
    const int a_st = 2;
    int tpg = 1;
    const int cand[2] = {ks, 1};                // a whole chunk's filter rows in one weight stage when it fits
    for (int c : cand)
        if ((size_t)c * b_tile <= 56 * 1024 && a_st * a_stride + 3 * (size_t)c * b_tile <= budget) { tpg = c; break; }
    g->tpg = tpg;
    int b_st = (int)((budget - a_st * a_stride) / ((size_t)tpg * b_tile));
    if (b_st > 10) b_st = 10;
    if (b_st < 2) return false;
    g->a_stages = a_st; g->b_stages = b_st;
    return true;
}

bool plan_hpack(const ramnet_handle *h, const ramnet_conv_desc *d, HaloGeom *g) {
    static const int pair_mode = [] { const char *e = getenv("RAMNET_PAIR"); return e ? atoi(e) : 1; }();
    const int cs = d->Cout % 32 == 0 ? 32 : 16;
    if (d->epilogue == RAMNET_EPI_BIAS_RELU_PRED && cs != d->Cout) return false;    // one slice: whole-row reduction
    if (pair_mode >= 1 && fill_hpack(d, g, 1, cs, 1)) return true;
    return fill_hpack(d, g, 1, cs, 0);
}

// Cycle model of one launch (per SM), from measurements on B200:
//  - shared memory moves 128 B/clk/SM and is shared between the UMMA operand reads and the TMA fills:
//    a 128xBNx8 tf32 MMA reads (128+BN)*32 B (tools/microbench/umma_rate.cu: max(BN/2,(128+BN)/4) clk),
//    each tap additionally writes one BN x 128 B weight tile and 1/taps of the halo box
//    (RAMNET_PROF counters: 678 clk per tap measured vs 676 predicted for BN=128, 2 tiles);
//  - L2 -> SM delivers ~15 TB/s chip-wide (~50 B/clk/SM);
//  - the epilogue is hidden under the next item when TMEM is double-buffered;
//  - persistent scheduling => ceil(items / SMs) rounds.
double halo_cost(const ramnet_handle *h, const ramnet_conv_desc *d, const HaloGeom &g, int taps_override = 0,
                 bool flat_issue = false) {
    const int ntiles = g.PTX * g.PTY, chunks = (d->C0 + d->C1) / kChunk, taps = taps_override ? taps_override : g.kh * g.kw;
    const double halo_bytes = (double)g.nplanes * g.HX * g.HY * 128.0 * (taps_override ? 4.0 : 1.0);   // s2seg: 4 planes per chunk
    const double bn_l = g.pair ? g.BN / 2.0 : (double)g.BN;     // weight rows read from / written to THIS SM's shared memory
    const double smem_tap = (ntiles * 4.0 * (128 + bn_l) * 32.0 + bn_l * 128.0 + halo_bytes / taps) / 128.0;
    const double math_tap = ntiles * 4.0 * g.BN / 2.0;
    // The issuing thread runs beside the (asynchronous) tensor pipe: per weight stage a barrier poll + fence + commit
    // (multicast across the pair: ~300 clk measured, ~60 single), shared by tpg taps, plus ~26 clk per MMA issued.
    // A tap costs whichever is slower (RAMNET_PROF: 410 / 514 clk per tap for 1 / 2 tiles of N = 64 in pair mode).
    // Round 2 (RAMNET_PROF on the parity-plane stride-2 layers, profiles/r02_s2seg.txt): in pair mode the fixed part is
    // ~220 clk per TAP whatever tpg is (380 / 483 / 796 clk per tap for 1 / 2 / 4 tiles of N = 64 at tpg = 3).  The s2seg
    // planner uses that (flat_issue); the stride-1 plans keep the round-1 term unless RAMNET_ISSUE_MODEL=1.
    static const bool issue_v2 = [] { const char *e = getenv("RAMNET_ISSUE_MODEL"); return e && e[0] == '1'; }();
    const double issue_tap = ((flat_issue || issue_v2) && g.pair ? 220.0 : (g.pair ? 300.0 : 60.0) / g.tpg) + 26.0 * 4.0 * ntiles;
    double tap = smem_tap > math_tap ? smem_tap : math_tap;
    if (issue_tap > tap) tap = issue_tap;
    tap += 10.0;
    const double mma = (double)chunks * taps * tap;
    // L2 -> SM bytes per clock per SM the model assumes (RAMNET_L2_BPC overrides for tuning runs)
    static const double l2_bpc = [] { const char *e = getenv("RAMNET_L2_BPC"); return e ? atof(e) : 50.0; }();
    const double l2 = (double)chunks * (halo_bytes + (double)taps * bn_l * 128.0) / l2_bpc;
    const double main = mma > l2 ? mma : l2;
    // epilogue of one item, calibrated on RAMNET_PROF (total - mma_total of single-round layers, round 2): cycles per
    // accumulator column of a 128-row tile: ~57 for bias / relu / residual, ~115 for the GRU reset/update gates
    // (two outputs + h), ~150 for the GRU candidate / LSTM (two aux operands, blend).  Opt-in (RAMNET_EPI_MODEL=1): measured
    // on B200 the plans it picks are SLOWER (one pass 895 -> 941 us, profiles/r02_epilogue_model_ab.txt), so the
    // round-1 model stays the default.
    static const bool epi_v2 = [] { const char *e = getenv("RAMNET_EPI_MODEL"); return e && e[0] == '1'; }();
    const double k_epi = d->epilogue == RAMNET_EPI_GRU_RU ? 115.0
                         : (d->epilogue == RAMNET_EPI_GRU_OUT || d->epilogue == RAMNET_EPI_LSTM) ? 150.0 : 57.0;
    const double epi = epi_v2 ? (double)ntiles * g.BN * k_epi + 1500.0
                              : (double)ntiles * (g.BN / 32.0 + 0.5) * 900.0 + 500.0;
    const double per_item = g.nbuf == 2 ? (main > epi ? main : epi) + 300.0 : main + epi;
    const int workers = g.pair ? h->sm_count / 2 : h->sm_count;
    // Weight of the wave quantisation in the plan cost.  1 = makespan of a kernel that has the GPU to itself:
    // ceil(items / workers) rounds.  0 (RAMNET_FLAG_SM_TIME) = SM time: items / workers rounds -- the right measure when
    // another stream's kernel takes the SMs this one leaves idle in its last wave (engine.GraphRunner overlaps passes;
    // measured 5020 -> 5137 maps/s).  RAMNET_PLAN_WAVES overrides for A/B runs.
    static const double wave_env = [] { const char *e = getenv("RAMNET_PLAN_WAVES"); return e ? atof(e) : -1.0; }();
    const double wave_w = wave_env >= 0.0 ? wave_env : ((d->flags & RAMNET_FLAG_SM_TIME) ? 0.0 : 1.0);
    const double rounds_up = (double)((g.items + workers - 1) / workers), rounds_fr = (double)g.items / workers;
    const double rounds = wave_w * rounds_up + (1.0 - wave_w) * (rounds_fr < 1.0 ? 1.0 : rounds_fr);
    double cost = rounds * per_item + (g.nbuf == 2 ? epi : 0.0) + 4000.0;
    // stride 2: one 4-plane halo per item and a single K chunk for the first encoder -- the halo fetch is exposed and the
    // model above underestimates both modes by ~2x; measured (enc0, 32->64): pair 2x1 55 us vs single 1x1 67 us
    if (d->stride == 2 && !g.pair) cost *= 1.25;
    if (flat_issue && !g.pair) cost *= 1.15;       // s2seg, measured: enc0 2x1 50.7 us single vs 45.2 us pair
    return cost;
}

// Chooses patch shape, BN and pipeline depths for the halo kernel; returns false when the layer
// should stay on the per-tap kernel (stride 2, 1x1, or no configuration fits shared memory).
bool plan_halo(const ramnet_handle *h, const ramnet_conv_desc *d, const RectSpec *rect, HaloGeom *g) {
    if ((halo_mode_env() & 4) && !rect) return false;   // RAMNET_HALO_MODE=4: force the per-tap kernel (A/B tests)
    if (d->ksize == 1 && !rect) return false;
    if (d->stride == 2 && ((d->H | d->W) & 1)) return false;
    if (const char *f = getenv("RAMNET_HALO_FORCE")) {   // tuning aid: "PTX,PTY,BN,a_stages,b_stages[,pair]"
        int ptx, pty, bn, ast, bst, pr = 0;
        if (sscanf(f, "%d,%d,%d,%d,%d,%d", &ptx, &pty, &bn, &ast, &bst, &pr) >= 5 && fill_halo(d, rect, g, ptx, pty, bn, ast, bst, pr))
            return true;
    }
    // RAMNET_PAIR: 0 = single-CTA MMAs only, 1 (default) = cost model decides, 2 = CTA pairs wherever a configuration fits
    static const int pair_mode = [] { const char *e = getenv("RAMNET_PAIR"); return e ? atoi(e) : 1; }();
    static const int shapes[][2] = {{2, 1}, {1, 1}, {4, 1}, {2, 2}, {1, 2}};
    double best = -1;
    HaloGeom cand;
    for (int pair = (pair_mode == 2 ? 1 : 0); pair <= (pair_mode >= 1 ? 1 : 0); ++pair)
        for (const auto &sh : shapes)
            for (int bn = 256; bn >= 16; bn >>= 1) {
                if (d->epilogue == RAMNET_EPI_BIAS_RELU_PRED && bn != d->Cout) continue;   // one slice: whole-row reduction
                if (!fill_halo(d, rect, &cand, sh[0], sh[1], bn, 0, 0, pair)) continue;
                const double c = halo_cost(h, d, cand);
                if (best < 0 || c < best) { best = c; *g = cand; }
            }
    if (best < 0 && pair_mode == 2) {     // nothing fits as a pair: single-CTA configurations
        for (const auto &sh : shapes)
            for (int bn = 256; bn >= 16; bn >>= 1) {
                if (d->epilogue == RAMNET_EPI_BIAS_RELU_PRED && bn != d->Cout) continue;
                if (!fill_halo(d, rect, &cand, sh[0], sh[1], bn, 0, 0, 0)) continue;
                const double c = halo_cost(h, d, cand);
                if (best < 0 || c < best) { best = c; *g = cand; }
            }
    }
    return best >= 0;
}
}  // namespace


// Sums the per-CTA partial tiles over the pixel splits and accumulates into dW [Cout][Ct][ks][ks] (deterministic:
// every gradient element is owned by one thread, no atomics).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ part, float *__restrict__ dw, WgGeom g,
                                                           int splits, int groups) {
    const int taps = g.ks * g.ks, Ct = g.C0 + g.C1, ncols = g.T * g.BN;
    const int64_t total = (int64_t)groups * 128 * ncols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cc = (int)(i % ncols);
        const int row = (int)((i / ncols) % 128);
        int grp = (int)(i / ((int64_t)ncols * 128));
        const int nb = grp % g.n_blocks;
        grp /= g.n_blocks;
        const int mb = grp % g.m_blocks, tr = grp / g.m_blocks;
        const int tap = g.T == 1 ? tr : tr * g.ks + cc / g.BN;      // row mode: group = filter row, column block = tap in row
        const int mch = mb * 128 + row, nch = nb * g.BN + cc % g.BN;
        if (mch >= g.Mch || nch >= g.Nch) continue;
        float acc = 0.f;
        for (int sp = 0; sp < splits; ++sp) acc += part[(size_t)sp * total + i];
        const int co = g.m_from_x ? nch : mch, ci = g.m_from_x ? mch : nch;
        dw[((int64_t)co * Ct + ci) * taps + tap] += acc;
    }
}

// Reduction of the tap-packed partial tiles, two passes (deterministic: fixed summation order, no atomics).
// Pass 1 sums the pixel splits element-wise into the split-0 tile: a block owns 64 consecutive elements and its four
// thread rows each take every fourth split, so the column-major tiles are read as full 256-byte runs.
__global__ void __launch_bounds__(256) wgrad_packed_sum_kernel(float *__restrict__ part, int64_t total, int splits,
                                                               long long part_stride) {
    part += (size_t)blockIdx.y * part_stride;      // blockIdx.y = problem of the batch
    __shared__ float red[4][64];
    const int e = threadIdx.x & 63, q = threadIdx.x >> 6;
    for (int64_t base = (int64_t)blockIdx.x * 64; base < total; base += (int64_t)gridDim.x * 64) {
        const int64_t i = base + e;
        float acc = 0.f;
        if (i < total) {
            float a0 = 0.f, a1 = 0.f;
            int sp = q;
            for (; sp + 4 < splits; sp += 8) {
                a0 += part[(size_t)sp * total + i];
                a1 += part[(size_t)(sp + 4) * total + i];
            }
            if (sp < splits) a0 += part[(size_t)sp * total + i];
            acc = a0 + a1;
        }
        red[q][e] = acc;
        __syncthreads();
        if (q == 0 && i < total) part[i] = (red[0][e] + red[1][e]) + (red[2][e] + red[3][e]);
        __syncthreads();
    }
}

// Pass 2 adds the summed tiles into dW [Cout][Ct][ks][ks].  A block takes 32 accumulator rows of one group, stages
// them in shared memory and writes them back ordered (output channel, input channel, tap), where for a stride-1 layer
// they form contiguous runs (for one output channel: 32 input channels x the group's filter rows).
constexpr int kWpPitch = 32 * 33 + 1;   // shared-memory words per (row shift, tap) plane: conflict-free both ways
__global__ void __launch_bounds__(1024) wgrad_packed_scatter_kernel(const float *__restrict__ sum, float *__restrict__ dw,
                                                                   const __grid_constant__ WpBatch batch, int groups,
                                                                   int splits) {
    // splits > 1: the pixel splits have not been summed by pass 1 (few splits: summing them while staging the tile
    // saves a launch and a pass over the workspace); the order of the additions is fixed either way
    const WpGeom &g = batch.g[blockIdx.y];
    sum += (size_t)blockIdx.y * batch.part_stride;
    const size_t split_stride = (size_t)groups * 128 * g.ncols;
    extern __shared__ float plane[];    // [RG * kw][32 c32][33] (+1 per plane)
    const int taps = g.ks * g.ks, Ct = g.C0 + g.C1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int unit = blockIdx.x; unit < groups * 4; unit += gridDim.x) {
        const int rb = unit & 3;
        int grp = unit >> 2;
        const float *tile = sum + (size_t)grp * 128 * g.ncols + rb * 32;
        const int nb = grp % g.n_boxes;
        grp /= g.n_boxes;
        const int mb = grp % g.m_blocks, rg = grp / g.m_blocks;
        const int u0 = rg * g.RG;
        const int nu = min(g.RG, g.kh - u0);              // valid row shifts of this group
        const int nt = nu * g.kw;                         // taps held by this tile
        __syncthreads();
        for (int col = warp; col < nt * kChunk; col += (int)(blockDim.x >> 5)) {  // col = (u * kw + j) * 32 + c32
            const float *src = tile + (size_t)col * 128 + lane;
            float acc = src[0];
            for (int sp = 1; sp < splits; ++sp) acc += src[(size_t)sp * split_stride];
            plane[(col >> 5) * kWpPitch + (col & 31) * 33 + lane] = acc;
        }
        __syncthreads();
        const int m0 = mb * 128 + rb * 32, n0 = nb * kChunk;
        const int count = 32 * 32 * nt;
        for (int ob = threadIdx.x; ob < count; ob += (int)blockDim.x * 4) {   // 4 independent read-modify-writes in flight
            float v[4], dcur[4];
            int64_t addr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int o = ob + k * (int)blockDim.x;
                addr[k] = -1;
                if (o < count) {
                    const int t = o % nt;                     // (row shift, N block) of the tile, fastest index
                    const int mid = (o / nt) & 31, outer = o / (nt * 32);
                    const int row = g.m_from_x ? mid : outer, c32 = g.m_from_x ? outer : mid;   // -> (co, ci, tap) order
                    int mch = m0 + row, nch = n0 + c32, ufold = 0;
                    if (g.mfold && mch >= 64) { mch -= 64; ufold = g.RG; }        // rows 64..127: same channels, filter rows u + RG
                    if (mch < g.Mch && nch < g.Nch && u0 + t / g.kw + ufold < g.kh) {
                        const int u = u0 + t / g.kw + ufold, j = t % g.kw;
                        const int tap = (g.r0 + g.dr * u) * g.ks + g.s0 + g.ds * j;
                        const int co = g.m_from_x ? nch : mch, ci = g.m_from_x ? mch : nch;
                        if (g.head_cin > 0) {          // unrolled head input: channel ci = dx*Cin + ci'
                            const int dx = ci / g.head_cin, cih = ci - dx * g.head_cin;
                            if (dx < 5) addr[k] = ((int64_t)co * g.head_cin + cih) * taps + tap + dx;
                        } else {
                            addr[k] = ((int64_t)co * Ct + ci) * taps + tap;
                        }
                        v[k] = plane[t * kWpPitch + c32 * 33 + row];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) dcur[k] = addr[k] >= 0 ? dw[addr[k]] : 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (addr[k] >= 0) dw[addr[k]] = dcur[k] + v[k];
        }
    }
}

// Number of tap-packed problems a layer decomposes into: 1 (stride 1), 4 (stride 2: one per input parity class), 0 = not covered.
int wgrad_packed_problems(const ramnet_conv_desc *d) {
    static const int version = [] { const char *e = getenv("RAMNET_WGRAD_V"); return e ? atoi(e) : 2; }();
    if (version < 2) return 0;                           // RAMNET_WGRAD_V=1: filter-row kernel everywhere (A/B runs)
    if (d->ksize != 3 && d->ksize != 5) return 0;
    if (d->C0 % kChunk || d->C1 % kChunk || d->Cout % kChunk) return 0;
    if (d->stride == 1) return 1;
    if (d->stride == 2 && !((d->H | d->W) & 1) && !getenv("RAMNET_WGRAD_S2_OLD")) return 4;
    return 0;
}

// Plans problem `cls` of a layer.  Stride 2: dW[r][s] = sum dZ[oy][ox] * X[2 oy + r - pad][2 ox + s - pad]; with
// r - pad = 2 dy + py the taps of one row parity py read the input parity plane P_py[y'][x'] = X[2 y' + py][..] at
// oy + dy, i.e. a stride-1 problem between dZ and the plane with the sub-filter {r : (r - pad) mod 2 = py}.
bool plan_wgrad_packed(const ramnet_handle *h, const ramnet_conv_desc *d, int cls, int nprob, WpGeom *gp, int *splits_out,
                       int *groups_out, int head_cin = 0) {
    WpGeom &g = *gp;
    g.head_cin = head_cin;
    g.N = d->N; g.Cout = d->Cout; g.C0 = d->C0; g.C1 = d->C1; g.ks = d->ksize;
    const int pad = d->ksize / 2, Ct = d->C0 + d->C1;
    g.m_from_x = Ct > d->Cout ? 1 : 0;                   // the operand with more channels fills the 128 MMA rows
    if (d->stride == 1) {
        g.H = d->H; g.W = d->W; g.kh = g.kw = g.ks; g.x5d = 0; g.py = g.px = 0;
        if (g.m_from_x) { g.oy = g.ox = pad - (g.ks - 1); g.r0 = g.s0 = g.ks - 1; g.dr = g.ds = -1; }
        else { g.oy = g.ox = -pad; g.r0 = g.s0 = 0; g.dr = g.ds = 1; }
        if (head_cin > 0) {   // horizontal taps already live in the channel axis: a ks x 1 problem, no x shift
            if (g.m_from_x) return false;
            g.kw = 1; g.ox = 0; g.s0 = 0; g.ds = 0;
        }
    } else {
        g.H = d->H / 2; g.W = d->W / 2; g.x5d = 1; g.py = cls >> 1; g.px = cls & 1;
        // taps of parity q along one axis: r = rmin + 2 i, shifts dy = (r - pad - q) / 2 = dmin + i
        auto axis = [&](int q, int *k, int *rmin, int *dmin) {
            *rmin = ((pad + q) & 1);                      // smallest r with (r - pad) mod 2 == q
            *k = (g.ks - 1 - *rmin) / 2 + 1;
            const int a = *rmin - pad - q;                // even
            *dmin = a >= 0 ? a / 2 : -((-a) / 2);
        };
        int rmin, smin, dymin, dxmin;
        axis(g.py, &g.kh, &rmin, &dymin);
        axis(g.px, &g.kw, &smin, &dxmin);
        if (g.m_from_x) {     // N = dZ shifted by -(dy, dx): box origin = -(dmax), shift u <-> dy = dmax - u
            g.oy = -(dymin + g.kh - 1); g.ox = -(dxmin + g.kw - 1);
            g.r0 = rmin + 2 * (g.kh - 1); g.dr = -2; g.s0 = smin + 2 * (g.kw - 1); g.ds = -2;
        } else {              // N = plane shifted by (dy, dx): box origin = dmin, shift u <-> dy = dmin + u
            g.oy = dymin; g.ox = dxmin;
            g.r0 = rmin; g.dr = 2; g.s0 = smin; g.ds = 2;
        }
    }
    g.Mch = g.m_from_x ? Ct : d->Cout;
    g.Nch = g.m_from_x ? d->Cout : Ct;
    g.m_blocks = (g.Mch + 127) / 128;
    g.n_boxes = g.Nch / kChunk;
    g.RG = 3;                                            // 3 * kw * 32 <= 480 accumulator columns
    g.TR = 8;                                            // 64-pixel K tiles: measured 10-15 % faster than 32 (fewer barrier round trips)
    int ctas_per_sm_total = 2, smem_budget = 200 * 1024;
    if (const char *f = getenv("RAMNET_WGP")) {          // tuning aid: "TR,RG,total CTAs per SM,smem KB"
        int a = 0, b = 0, c = 0, e = 0;
        if (sscanf(f, "%d,%d,%d,%d", &a, &b, &c, &e) == 4) {
            if (a == 4 || a == 8) g.TR = a;
            if (b >= 1 && b <= 3 && b * g.kw * kChunk <= 512) g.RG = b;
            if (c >= 1) ctas_per_sm_total = c;
            if (e >= 32 && e <= 200) smem_budget = e * 1024;
        }
    }
    if (g.RG > g.kh) g.RG = g.kh;
    g.prof = nullptr;
    g.row_groups = (g.kh + g.RG - 1) / g.RG;
    // 64-channel M operand with more filter rows than one CTA holds (dec2: 64 -> 32, 5x5): fold the second row group
    // into the idle upper half of the 128 MMA rows instead of running it as a second, half-empty CTA group
    // (dec2 182 -> 121 us per call on B200, profiles/r02_first_call_experimental_paths.txt; RAMNET_WGRAD_FOLD=0 disables)
    static const bool fold_on = [] { const char *e = getenv("RAMNET_WGRAD_FOLD"); return !(e && e[0] == '0'); }();
    g.mfold = (fold_on && d->stride == 1 && head_cin == 0 && g.Mch == 64 && g.kh > g.RG && g.kh <= 2 * g.RG) ? 1 : 0;
    if (g.mfold) g.row_groups = 1;
    g.HXw = 8 + g.kw - 1; g.HYw = g.TR + g.RG - 1;
    g.ncols = g.RG * g.kw * kChunk;
    // folded rows read the M operand RG rows above the tile: RG extra rows of tiles at the bottom cover its last rows
    g.tiles_x = (g.W + 7) / 8; g.tiles_y = (g.H + (g.mfold ? g.RG : 0) + g.TR - 1) / g.TR;
    const int64_t total = (int64_t)g.tiles_x * g.tiles_y * g.N;
    if (total > 0x7fffffff) return false;
    g.total_tiles = (int)total;
    const int groups = g.row_groups * g.m_blocks * g.n_boxes;
    // One CTA per SM at a time (the accumulators take all of TMEM).  Pick the pixel split that minimises
    //   waves * (K tiles per CTA + fixed prologue/epilogue cost) + reduction traffic,
    // in units of one K tile (~1 us): a wave that fills only a few SMs costs as much as a full one.
    int64_t splits = 1;
    {
        double best = -1;
        const int64_t max_splits = total / 8 > 0 ? total / 8 : 1;
        for (int waves = 1; waves <= ctas_per_sm_total; ++waves) {
            int64_t sp = ((int64_t)h->sm_count * waves) / ((int64_t)groups * nprob);   // the problems of a layer share one launch
            if (sp < 1) sp = 1;
            if (sp > max_splits) sp = max_splits;
            const int64_t tpc = (total + sp - 1) / sp;
            sp = (total + tpc - 1) / tpc;
            const int64_t w = (sp * groups * nprob + h->sm_count - 1) / h->sm_count;
            const double cost = (double)w * ((double)tpc * g.TR / 8.0 + 3.7) + 0.05 * (double)(sp * groups * nprob);
            if (best < 0 || cost < best) { best = cost; splits = sp; }
        }
    }
    g.tiles_per_cta = (int)((total + splits - 1) / splits);
    splits = (total + g.tiles_per_cta - 1) / g.tiles_per_cta;
    const int nbox = g.HXw * g.HYw * kChunk * 4;
    g.stage_bytes = 4 * g.TR * 8 * kChunk * 4 + ((nbox + 1023) & ~1023);
    g.stages = smem_budget / g.stage_bytes;
    if (g.stages > 8) g.stages = 8;
    if (g.stages < 2) return false;
    *splits_out = (int)splits;
    *groups_out = groups;
    return true;
}

// Plans every problem of a layer and unifies what the shared launch needs (N box, workspace pitch, pipeline depth).
bool plan_wgrad_batch(const ramnet_handle *h, const ramnet_conv_desc *d, WpBatch *b, int *splits_out, int *groups_out,
                      int head_cin = 0) {
    b->n = wgrad_packed_problems(d);
    if (b->n == 0) return false;
    int splits = 0, groups = 0;
    for (int cls = 0; cls < b->n; ++cls) {
        int sp, gr;
        if (!plan_wgrad_packed(h, d, cls, b->n, &b->g[cls], &sp, &gr, head_cin)) return false;
        if (cls == 0) { splits = sp; groups = gr; }
        else if (sp != splits || gr != groups || b->g[cls].tiles_per_cta != b->g[0].tiles_per_cta) return false;
    }
    WpGeom &g0 = b->g[0];
    for (int cls = 1; cls < b->n; ++cls) {
        const WpGeom &g = b->g[cls];
        if (g.HXw > g0.HXw) g0.HXw = g.HXw;
        if (g.HYw > g0.HYw) g0.HYw = g.HYw;
        if (g.ncols > g0.ncols) g0.ncols = g.ncols;
        if (g.stage_bytes > g0.stage_bytes) g0.stage_bytes = g.stage_bytes;
        if (g.stages < g0.stages) g0.stages = g.stages;
    }
    const int nbox = g0.HXw * g0.HYw * kChunk * 4;
    g0.stage_bytes = 4 * g0.TR * 8 * kChunk * 4 + ((nbox + 1023) & ~1023);
    while (g0.stages > 2 && (size_t)g0.stages * g0.stage_bytes > 200 * 1024) --g0.stages;
    for (int cls = 1; cls < b->n; ++cls) {
        WpGeom &g = b->g[cls];
        g.HXw = g0.HXw; g.HYw = g0.HYw; g.ncols = g0.ncols; g.stage_bytes = g0.stage_bytes; g.stages = g0.stages;
    }
    b->part_stride = (long long)splits * groups * 128 * g0.ncols;
    *splits_out = splits;
    *groups_out = groups;
    return true;
}

bool plan_wgrad(const ramnet_handle *h, const ramnet_conv_desc *d, WgGeom *gp, int *splits_out, int *groups_out);

size_t conv_wgrad_tf32_workspace(const ramnet_handle *h, const ramnet_conv_desc *d, int head_cin) {
    int splits, groups;
    {
        WpBatch b;
        if (plan_wgrad_batch(h, d, &b, &splits, &groups, head_cin)) return (size_t)b.n * b.part_stride * sizeof(float);
    }
    if (head_cin > 0) return 0;
    WgGeom g;
    if (!plan_wgrad(h, d, &g, &splits, &groups)) return 0;
    return (size_t)splits * groups * 128 * g.T * g.BN * sizeof(float);
}

// dW += dZ^T * im2col(X) on the tensor cores (see conv_wgrad_tcgen05_kernel).  Returns RAMNET_EUNSUPPORTED
// for shapes the kernel does not cover; the caller then uses the fp32 FFMA kernel.
bool plan_wgrad(const ramnet_handle *h, const ramnet_conv_desc *d, WgGeom *gp, int *splits_out, int *groups_out) {
    if (d->C0 % kChunk || d->C1 % kChunk || d->Cout % 16) return false;
    if (d->stride == 2 && ((d->H | d->W) & 1)) return false;
    WgGeom &g = *gp;
    g.N = d->N; g.Ho = conv_out_dim(d->H, d->stride); g.Wo = conv_out_dim(d->W, d->stride);
    g.Cout = d->Cout; g.C0 = d->C0; g.C1 = d->C1; g.ks = d->ksize; g.pad = d->ksize / 2; g.stride = d->stride;
    const int Ct = d->C0 + d->C1;
    g.m_from_x = Ct > d->Cout ? 1 : 0;                 // the operand with more channels fills the 128 MMA rows
    g.Mch = g.m_from_x ? Ct : d->Cout;
    g.Nch = g.m_from_x ? d->Cout : Ct;
    // Row mode (stride 1): one CTA owns a whole filter row; the X box carries a (ks-1)-pixel halo and every tap of the row
    // reads it through a shifted descriptor, so X and dZ are fetched ks times instead of ks*ks times.  The ks accumulators
    // need ks*BN <= 512 TMEM columns.
    g.T = (d->stride == 1 && d->ksize > 1 && !getenv("RAMNET_WGRAD_PER_TAP")) ? d->ksize : 1;
    g.TR = g.T == 1 ? 8 : 4;
    g.HXw = 8 + g.T - 1;
    int bn_cap = g.T == 1 ? 128 : (g.T == 3 ? 128 : 64);
    int ctas_per_sm = g.T == 1 ? 2 : 1;        // row mode: T*BN accumulator columns may take all of TMEM
    int force_stages = 0;
    if (const char *f = getenv("RAMNET_WGRAD_FORCE")) {          // "BN,TR,stages,ctas_per_sm" (tuning / tests)
        int a = 0, b = 0, c = 0, e = 0;
        if (sscanf(f, "%d,%d,%d,%d", &a, &b, &c, &e) == 4 && g.T > 1) {
            if (a >= 32 && a % 32 == 0 && a * g.T <= 512) bn_cap = a;
            if (b == 4 || b == 8) g.TR = b;
            force_stages = c;
            if (e == 1 || e == 2) ctas_per_sm = e;
        }
    }
    g.BN = g.Nch >= bn_cap ? bn_cap : ((g.Nch + 31) / 32) * 32;
    if (g.T * g.BN > 256) ctas_per_sm = 1;
    const int x_box = g.HXw * g.TR * kChunk * 4, dz_box = 8 * g.TR * kChunk * 4;
    g.a_box = g.m_from_x ? x_box : dz_box;
    g.b_box = g.m_from_x ? dz_box : x_box;
    g.m_blocks = (g.Mch + 127) / 128;
    g.n_blocks = (g.Nch + g.BN - 1) / g.BN;
    g.tiles_x = (g.Wo + 7) / 8; g.tiles_y = (g.Ho + g.TR - 1) / g.TR;
    const int64_t total = (int64_t)g.tiles_x * g.tiles_y * g.N;
    if (total > 0x7fffffff) return false;
    g.total_tiles = (int)total;
    const int groups = (g.T == 1 ? d->ksize * d->ksize : d->ksize) * g.m_blocks * g.n_blocks;
    // ~4 CTAs per SM in total, but at least 1024 pixels (8K MMA cycles) per CTA so the partial-tile flush amortises
    const int min_tiles = 1024 / (8 * g.TR);
    int64_t splits = ((int64_t)h->sm_count * 4 + groups - 1) / groups;
    if (splits > total / min_tiles) splits = total / min_tiles;
    if (splits < 1) splits = 1;
    g.tiles_per_cta = (int)((total + splits - 1) / splits);
    splits = (total + g.tiles_per_cta - 1) / g.tiles_per_cta;
    const int stage_bytes = 4 * g.a_box + (g.BN / kChunk) * g.b_box;
    // per-tap mode: two CTAs per SM (measured faster than one CTA with a deeper ring)
    g.stages = ((ctas_per_sm == 2 ? 96 : 200) * 1024) / stage_bytes;
    if (force_stages > 0) g.stages = force_stages;
    if (g.stages < 2) g.stages = 2;
    if (g.stages > 8) g.stages = 8;
    if ((size_t)g.stages * stage_bytes > 220 * 1024) return false;
    *splits_out = (int)splits;
    *groups_out = groups;
    return true;
}

// mode (RAMNET_WGRAD_*): FULL = partial tiles + split sum + scatter into dw (+=); PARTIAL_FIRST / PARTIAL_ADD = only the
// tensor-core kernel, whose epilogue overwrites / adds to the partial tiles in `workspace` (which the caller keeps for the
// layer across the passes of a step); FINALIZE = only the split sum + scatter of what the workspace holds.
int conv_wgrad_tf32(ramnet_handle *h, const ramnet_conv_desc *d, const float *dz, const float *x0, const float *x1,
                    float *dw, void *workspace, size_t workspace_bytes, cudaStream_t s, int head_cin, int mode) {
    if ((((uintptr_t)dz | (uintptr_t)x0 | (uintptr_t)x1 | (uintptr_t)workspace) & 15) != 0) return RAMNET_EUNSUPPORTED;
    WpBatch batch;
    int psplits, pgroups;
    if (plan_wgrad_batch(h, d, &batch, &psplits, &pgroups, head_cin)) {
        batch.accumulate = mode == RAMNET_WGRAD_PARTIAL_ADD ? 1 : 0;
        const WpGeom &p = batch.g[0];
        const size_t need = (size_t)batch.n * batch.part_stride * sizeof(float);
        RAMNET_CHECK_ARG(workspace != nullptr && workspace_bytes >= need,
                         "conv_wgrad(tf32): workspace of %zu bytes required (ramnet_conv_wgrad_workspace_bytes)", need);
        CUtensorMap mdz, m0, m1;
        const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
        static const bool do_prof = getenv("RAMNET_PROF") != nullptr;      // debug only: synchronises and prints
        static unsigned long long *pbuf = nullptr;
        cudaEvent_t ev[3];
        if (mode != RAMNET_WGRAD_FINALIZE) {
        {
            cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)d->N};
            cuuint64_t str[3] = {(cuuint64_t)d->Cout * 4, (cuuint64_t)p.W * d->Cout * 4, (cuuint64_t)p.H * p.W * d->Cout * 4};
            cuuint32_t box[4] = {kChunk, (cuuint32_t)(p.m_from_x ? p.HXw : 8), (cuuint32_t)(p.m_from_x ? p.HYw : p.TR), 1};
            int rc = encode(h, &mdz, dz, 4, dims, str, box, sw);
            if (rc) return rc;
        }
        auto enc_x = [&](CUtensorMap *m, const float *base, int C) {
            const cuuint32_t bw = (cuuint32_t)(p.m_from_x ? 8 : p.HXw), bh = (cuuint32_t)(p.m_from_x ? p.TR : p.HYw);
            if (!p.x5d) {
                cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
                cuuint64_t str[3] = {(cuuint64_t)C * 4, (cuuint64_t)d->W * C * 4, (cuuint64_t)d->H * d->W * C * 4};
                cuuint32_t box[4] = {kChunk, bw, bh, 1};
                return encode(h, m, base, 4, dims, str, box, sw);
            }
            // stride 2: parity view (2C, W/2, 2, H/2, N); one box = a patch of one parity plane
            cuuint64_t dims[5] = {(cuuint64_t)2 * C, (cuuint64_t)d->W / 2, 2, (cuuint64_t)d->H / 2, (cuuint64_t)d->N};
            cuuint64_t str[4] = {(cuuint64_t)2 * C * 4, (cuuint64_t)d->W * C * 4, (cuuint64_t)2 * d->W * C * 4,
                                 (cuuint64_t)d->H * d->W * C * 4};
            cuuint32_t box[5] = {kChunk, bw, 1, bh, 1};
            return encode(h, m, base, 5, dims, str, box, sw);
        };
        int rc = enc_x(&m0, x0, d->C0);
        if (rc) return rc;
        if (x1) {
            rc = enc_x(&m1, x1, d->C1);
            if (rc) return rc;
        } else {
            m1 = m0;
        }
        const size_t smem = (size_t)p.stages * p.stage_bytes + (2 * p.stages + 1) * 8 + 16 + 1024;
        static size_t configured = 0;
        if (smem > configured) {
            RAMNET_CUDA(cudaFuncSetAttribute(conv_wgrad_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        if (getenv("RAMNET_DEBUG"))
            fprintf(stderr, "[ramnet] wgrad packed %dx%d C=%d+%d->%d k%d s%d: problems=%d m_from_x=%d groups=%d splits=%d tiles/cta=%d stages=%d\n",
                    d->H, d->W, d->C0, d->C1, d->Cout, d->ksize, d->stride, batch.n, p.m_from_x, pgroups, psplits, p.tiles_per_cta,
                    p.stages);
        dim3 grid((unsigned)psplits, (unsigned)pgroups, (unsigned)batch.n);
        if (do_prof) {
            if (!pbuf) cudaMalloc(&pbuf, 64);
            cudaMemsetAsync(pbuf, 0, 64, s);
            for (int c = 0; c < batch.n; ++c) batch.g[c].prof = pbuf;
            for (auto &e : ev) cudaEventCreate(&e);
            cudaEventRecord(ev[0], s);
        }
        conv_wgrad_packed_kernel<<<grid, kThreads, smem, s>>>(mdz, m0, m1, batch, (float *)workspace);
        RAMNET_LAUNCH_CHECK(h);
        if (do_prof) cudaEventRecord(ev[1], s);
        if (do_prof && mode != RAMNET_WGRAD_FULL) { cudaStreamSynchronize(s); for (auto &e : ev) cudaEventDestroy(e); }
        }   // mode != FINALIZE
        if (mode == RAMNET_WGRAD_PARTIAL_FIRST || mode == RAMNET_WGRAD_PARTIAL_ADD) return RAMNET_OK;
        RAMNET_CHECK_ARG(dw != nullptr, "conv_wgrad(tf32): dw is NULL");
        const int64_t elems = (int64_t)pgroups * 128 * p.ncols;
        // few splits: summed inside the scatter kernel (one pass 1355 -> 1251 us on B200; RAMNET_WGRAD_FUSED_SUM=0 runs
        // the separate sum pass)
        static const bool fuse_on = [] { const char *e = getenv("RAMNET_WGRAD_FUSED_SUM"); return !(e && e[0] == '0'); }();
        const bool fused_sum = fuse_on && psplits <= 16;
        if (psplits > 1 && !fused_sum) {
            dim3 sgrid((unsigned)imin64((elems + 63) / 64, (int64_t)h->sm_count * 8), (unsigned)batch.n);
            wgrad_packed_sum_kernel<<<sgrid, 256, 0, s>>>((float *)workspace, elems, psplits, batch.part_stride);
            RAMNET_LAUNCH_CHECK(h);
        }
        const size_t sc_smem = (size_t)3 * 5 * kWpPitch * sizeof(float);       // up to 3 row shifts x 5 taps
        static bool sc_configured = false;
        if (!sc_configured) {
            RAMNET_CUDA(cudaFuncSetAttribute(wgrad_packed_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
            sc_configured = true;
        }
        dim3 cgrid((unsigned)imin64((int64_t)pgroups * 4, (int64_t)h->sm_count * 2), (unsigned)batch.n);
        wgrad_packed_scatter_kernel<<<cgrid, 1024, sc_smem, s>>>((const float *)workspace, dw, batch, pgroups,
                                                                 fused_sum ? psplits : 1);
        RAMNET_LAUNCH_CHECK(h);
        if (do_prof && mode == RAMNET_WGRAD_FULL) {
            cudaEventRecord(ev[2], s);
            unsigned long long hb[8];
            cudaMemcpyAsync(hb, pbuf, 64, cudaMemcpyDeviceToHost, s);
            cudaStreamSynchronize(s);
            float t01 = 0, t12 = 0;
            cudaEventElapsedTime(&t01, ev[0], ev[1]);
            cudaEventElapsedTime(&t12, ev[1], ev[2]);
            const double n = hb[5] ? (double)hb[5] : 1.0;
            fprintf(stderr, "[ramnet-prof] wgrad packed %dx%d C=%d+%d->%d k%d s%d problems=%d TR=%d RG=%d groups=%d splits=%d tiles/cta=%d stages=%d | "
                            "kernel %.1f us, reduce %.1f us | per-CTA kcycles: prod_wait_empty=%.1f mma_wait_full=%.1f mma_total=%.1f "
                            "epilogue=%.1f cta_total=%.1f (n=%llu)\n",
                    d->H, d->W, d->C0, d->C1, d->Cout, d->ksize, d->stride, batch.n, p.TR, p.RG, pgroups, psplits, p.tiles_per_cta,
                    p.stages, t01 * 1e3, t12 * 1e3, hb[0] / n / 1e3, hb[1] / n / 1e3, hb[2] / n / 1e3, hb[3] / n / 1e3,
                    hb[4] / n / 1e3, hb[5]);
            for (auto &e : ev) cudaEventDestroy(e);
        }
        return RAMNET_OK;
    }
    if (head_cin > 0 || mode != RAMNET_WGRAD_FULL) return RAMNET_EUNSUPPORTED;      // the deferred modes exist for the tap-packed kernel only
    WgGeom g;
    int splits_i, groups;
    if (!plan_wgrad(h, d, &g, &splits_i, &groups)) return RAMNET_EUNSUPPORTED;
    const int64_t splits = splits_i;
    const size_t need = (size_t)splits * groups * 128 * g.T * g.BN * sizeof(float);
    RAMNET_CHECK_ARG(workspace != nullptr && workspace_bytes >= need,
                     "conv_wgrad(tf32): workspace of %zu bytes required (ramnet_conv_wgrad_workspace_bytes)", need);
    const int stage_bytes = 4 * g.a_box + (g.BN / kChunk) * g.b_box;

    CUtensorMap mdz, m0, m1;
    {
        cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)g.Wo, (cuuint64_t)g.Ho, (cuuint64_t)d->N};
        cuuint64_t str[3] = {(cuuint64_t)d->Cout * 4, (cuuint64_t)g.Wo * d->Cout * 4, (cuuint64_t)g.Ho * g.Wo * d->Cout * 4};
        cuuint32_t box[4] = {kChunk, 8, (cuuint32_t)g.TR, 1};
        int rc = encode(h, &mdz, dz, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
    }
    int rc = encode_activation(h, &m0, x0, d->N, d->H, d->W, d->C0, d->stride, g.HXw, g.TR, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    if (x1) {
        rc = encode_activation(h, &m1, x1, d->N, d->H, d->W, d->C1, d->stride, g.HXw, g.TR, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
    } else {
        m1 = m0;
    }
    const size_t smem = (size_t)g.stages * stage_bytes + (2 * g.stages + 1) * 8 + 16 + 1024;
    static size_t configured = 0;
    if (smem > configured) {
        RAMNET_CUDA(cudaFuncSetAttribute(conv_wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((unsigned)splits, (unsigned)groups);
    conv_wgrad_tcgen05_kernel<<<grid, kThreads, smem, s>>>(mdz, m0, m1, g, (float *)workspace);
    RAMNET_LAUNCH_CHECK(h);
    const int64_t elems = (int64_t)groups * 128 * g.T * g.BN;
    wgrad_reduce_kernel<<<(unsigned)imin64((elems + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, s>>>(
        (const float *)workspace, dw, g, (int)splits, groups);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

size_t conv_tf32_workspace_bytes(const ramnet_conv_desc *) { return 0; }


// ================================================================================================
// Up-conv (RAMNET_FLAG_UPCONV): UpsampleConvLayer.forward (submodules.py:87-97) = bilinear x2 + 5x5 conv in one launch on
// the low-resolution input.  out[2m+py, 2q+px] = sum_{t,u} W[t,u] Upad[2m+py+t, 2q+px+u]; every row of the upsampled
// image U is a fixed 2-tap combination of input rows (.25/.75), so per output phase (py, px) the 5x5 filter collapses
// onto a 5x5 window of the LOW-resolution input (one row and one column of it are zero): one GEMM with
// N = 4 * Cout columns [py][px][co] over the small tensor.  What the 4x tensor cost (written by upsample2x_add, read back
// 25 times by the conv) disappears, and the narrow decoders (Cout = 32 / 64) run N = 128 / 256 wide MMAs.
// Border: U's clamp replicates the edge and the conv's zero padding applies to U, not to the input.  The uniform
// collapse over the zero-padded input is therefore wrong in output rows 0..2 / 2H-3..2H-1 (same for columns), by terms
// that only involve input row 0 / H-1 (column 0 / W-1):
//     corr[0] = .25 (W[0] - W[-1]) x0,   corr[1] = .25 (W[-1] - W[-2]) x0,   corr[2] = .25 W[-2] x0      (1-D, top)
// and mirrored at the bottom; in 2-D  exact = (M + D)y (x) (M + D)x  gives four edge and four corner terms.  Each is
// one more K segment of the same launch: its A operand is the SAME box read through a tensor map whose bounds are just
// that row / column / pixel (everything else is zero fill) and its taps carry the collapsed D coefficients.  Only
// border patches run them, and only their border tiles issue MMAs.  The algebra is checked against
// F.conv2d(F.interpolate(x)) in fp64 by tests/test_upconv_algebra.py (exact to 1e-14, down to 2x2 inputs).
// ================================================================================================
// taps per segment: main 25, edges 10, corners 4 + 1 zero tap (so that every segment is a multiple of tpg in {1, 5})
__host__ __device__ inline int up_seg_taps(int sg) { return sg == 0 ? 25 : (sg <= 4 ? 10 : 5); }
constexpr int kUpTotalTaps = 85;

// coefficient of input row m + r in upsampled row 2m + p + t (uniform bilinear x2 formula)
__host__ __device__ inline float up_cu(int p, int t, int r) {
    const int sft = p + t;
    if ((sft & 1) == 0) {
        const int e = sft >> 1;
        if (r == e - 1) return 0.25f;
        if (r == e) return 0.75f;
    } else {
        const int e = (sft - 1) >> 1;
        if (r == e) return 0.75f;
        if (r == e + 1) return 0.25f;
    }
    return 0.f;
}
// top / left border correction: coefficient of the FIRST input row for output phase p of low-res row r rows below it
__host__ __device__ inline float up_dT(int p, int t, int r) {
    if (r == 0) {
        if (p == 0) return t == 0 ? 0.25f : (t == -1 ? -0.25f : 0.f);
        return t == -1 ? 0.25f : (t == -2 ? -0.25f : 0.f);
    }
    if (r == -1 && p == 0 && t == -2) return 0.25f;
    return 0.f;
}
__host__ __device__ inline float up_dB(int p, int t, int r) { return up_dT(1 - p, -t, -r); }   // mirror image

// segment sg: tap window and which 1-D operators apply along y / x (0 uniform, 1 first-row, 2 last-row correction)
__host__ __device__ inline void up_seg_shape(int sg, int &kh, int &kw, int &lo_y, int &lo_x, int &fy, int &fx) {
    const int ymode = (sg == 1 || sg == 5 || sg == 6) ? 1 : ((sg == 2 || sg == 7 || sg == 8) ? 2 : 0);
    const int xmode = (sg == 3 || sg == 5 || sg == 7) ? 1 : ((sg == 4 || sg == 6 || sg == 8) ? 2 : 0);
    fy = ymode; fx = xmode;
    kh = ymode ? 2 : 5; lo_y = ymode == 0 ? -2 : (ymode == 1 ? -1 : 0);
    kw = xmode ? 2 : 5; lo_x = xmode == 0 ? -2 : (xmode == 1 ? -1 : 0);
}

__global__ void pack_upconv_kernel(const float *__restrict__ w, float *__restrict__ out, int Cout, int Cin) {
    const int64_t total = (int64_t)kUpTotalTaps * 4 * Cout * Cin;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        int64_t q = i / Cin;
        const int row = (int)(q % (4 * Cout));
        int tap = (int)(q / (4 * Cout));
        int sg = 0;
        while (tap >= up_seg_taps(sg)) { tap -= up_seg_taps(sg); ++sg; }
        int kh, kw, lo_y, lo_x, fy, fx;
        up_seg_shape(sg, kh, kw, lo_y, lo_x, fy, fx);
        float acc = 0.f;
        if (tap < kh * kw) {
            const int r = lo_y + tap / kw, sx = lo_x + tap % kw;
            const int phase = row / Cout, co = row % Cout, py = phase >> 1, px = phase & 1;
            const float *wc = w + ((int64_t)co * Cin + ci) * 25;
            for (int t = -2; t <= 2; ++t) {
                const float a = fy == 0 ? up_cu(py, t, r) : (fy == 1 ? up_dT(py, t, r) : up_dB(py, t, r));
                if (a == 0.f) continue;
                for (int u = -2; u <= 2; ++u) {
                    const float b = fx == 0 ? up_cu(px, u, sx) : (fx == 1 ? up_dT(px, u, sx) : up_dB(px, u, sx));
                    if (b != 0.f) acc = fmaf(a * b, wc[(t + 2) * 5 + (u + 2)], acc);     // a * b is exact (powers of two x 3, 9)
                }
            }
        }
        out[i] = round_tf32(acc);
    }
}

int conv_up_fwd_tf32(ramnet_handle *h, const ramnet_conv_desc *d, const float *x0, const float *wp, EpiParams ep,
                     cudaStream_t s) {
    RAMNET_CHECK_ARG(d->ksize == 5 && d->stride == 1 && d->C1 == 0, "conv_fwd(upconv): 5x5 stride-1 single-source layers only");
    RAMNET_CHECK_ARG(d->C0 % kChunk == 0 && d->Cout % 16 == 0 && 4 * d->Cout <= 512, "conv_fwd(upconv): C0 %% 32, Cout %% 16, Cout <= 128");
    RAMNET_CHECK_ARG(d->H >= 2 && d->W >= 2, "conv_fwd(upconv): input must be at least 2 x 2");
    RAMNET_CHECK_ARG(d->epilogue == RAMNET_EPI_BIAS_RELU || d->epilogue == RAMNET_EPI_BIAS_RELU_ADD ||
                         d->epilogue == RAMNET_EPI_BIAS_RELU_PRED || d->epilogue == RAMNET_EPI_BIAS,
                     "conv_fwd(upconv): bias / relu / relu+add / fused-pred epilogues only");
    ramnet_conv_desc du = *d;
    du.Cout = 4 * d->Cout;            // GEMM columns
    HaloGeom hg;
    static const int pair_mode = [] { const char *e = getenv("RAMNET_PAIR"); return e ? atoi(e) : 1; }();
    static const int shapes[][2] = {{2, 1}, {1, 1}, {4, 1}, {2, 2}, {1, 2}};
    double best = -1;
    HaloGeom cand;
    if (const char *f = getenv("RAMNET_UP_FORCE")) {      // tuning aid: "PTX,PTY,BN,pair"
        int ptx, pty, bn, pr = 0;
        if (sscanf(f, "%d,%d,%d,%d", &ptx, &pty, &bn, &pr) >= 3 && fill_halo(&du, nullptr, &cand, ptx, pty, bn, 0, 0, pr, 5)) {
            best = 0; hg = cand;
        }
    }
    for (int pass = 0; pass < 2 && best < 0; ++pass)      // second pass: single-CTA configurations when pairs are forced but none fits
        for (int pair = (pair_mode == 2 && pass == 0 ? 1 : 0); pair <= (pair_mode >= 1 ? 1 : 0); ++pair)
            for (const auto &sh : shapes)
                for (int bn = 128; bn >= 64; bn >>= 1) {      // <= 128 columns per item: a stage then holds a whole filter row
                                                              // (tpg = 5); measured dec1: BN 256 / tpg 1 118 us, BN 128 / tpg 5 107 us
                    if (bn % d->Cout && d->Cout % bn) continue;                       // 16-column units never straddle a phase anyway
                    if (d->epilogue == RAMNET_EPI_BIAS_RELU_PRED && bn != du.Cout) continue;   // all four phases in one item
                    if (sh[0] * sh[1] > 2) continue;          // measured: 4-tile patches lose TMEM double buffering (dec2 116 vs 92 us)
                    if (!fill_halo(&du, nullptr, &cand, sh[0], sh[1], bn, 0, 0, pair, 5)) continue;
                    const double c = halo_cost(h, &du, cand);
                    if (best < 0 || c < best) { best = c; hg = cand; }
                }
    if (best < 0) return ramnet_set_error(RAMNET_EUNSUPPORTED, "conv_fwd(upconv): no halo configuration fits");
    hg.up = 1; hg.up_cout = d->Cout; hg.nseg = kMaxUpSeg; hg.total_taps = kUpTotalTaps;
    int tap0 = 0;
    for (int sg = 0; sg < kMaxUpSeg; ++sg) {
        UpSeg &u = hg.seg[sg];
        int fy, fx;
        up_seg_shape(sg, u.kh, u.kw, u.lo_y, u.lo_x, fy, fx);
        u.tap0 = tap0; u.ntaps = up_seg_taps(sg); tap0 += up_seg_taps(sg);
        u.vy = fy == 2 ? d->H - 1 : 0; u.vx = fx == 2 ? d->W - 1 : 0;
        u.cond = (fy == 1 ? 1 : (fy == 2 ? 2 : 0)) | (fx == 1 ? 4 : (fx == 2 ? 8 : 0));
    }
    if (getenv("RAMNET_DEBUG"))
        fprintf(stderr, "[ramnet] upconv plan %dx%d C=%d->%d (N=%d): tiles %dx%d HX=%d HY=%d BN=%d a_st=%d b_st=%d tpg=%d nbuf=%d items=%d pair=%d\n",
                d->H, d->W, d->C0, d->Cout, du.Cout, hg.PTX, hg.PTY, hg.HX, hg.HY, hg.BN, hg.a_stages, hg.b_stages, hg.tpg,
                hg.nbuf, hg.items, hg.pair);
    CUtensorMap m0, mw;
    UpMaps um;
    const int C = d->C0;
    const cuuint64_t str[3] = {(cuuint64_t)C * 4, (cuuint64_t)d->W * C * 4, (cuuint64_t)d->H * d->W * C * 4};
    const cuuint32_t box[4] = {kChunk, (cuuint32_t)hg.HX, (cuuint32_t)hg.HY, 1};
    for (int sg = 0; sg < kMaxUpSeg; ++sg) {
        const UpSeg &u = hg.seg[sg];
        // the view: rows [vy, vy + vh), columns [vx, vx + vw) of every image; same strides, shifted base
        const int vh = (u.cond & 3) ? 1 : d->H, vw = (u.cond & 12) ? 1 : d->W;
        const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)vw, (cuuint64_t)vh, (cuuint64_t)d->N};
        const float *base = x0 + ((int64_t)u.vy * d->W + u.vx) * C;
        const int rc = encode(h, sg == 0 ? &m0 : &um.m[sg - 1], base, 4, dims, str, box);
        if (rc) return rc;
    }
    const cuuint64_t wd[3] = {(cuuint64_t)C, (cuuint64_t)du.Cout, (cuuint64_t)kUpTotalTaps};
    const cuuint64_t wst[2] = {(cuuint64_t)C * 4, (cuuint64_t)C * du.Cout * 4};
    const cuuint32_t wb[3] = {kChunk, (cuuint32_t)(hg.pair ? hg.BN / 2 : hg.BN), (cuuint32_t)hg.tpg};
    int rc = encode(h, &mw, wp, 3, wd, wst, wb);
    if (rc) return rc;
    ep.Cout = d->Cout;                // pixel pitch of the high-resolution output / aux operands
    switch (d->epilogue) {
        case RAMNET_EPI_BIAS: return launch_halo<RAMNET_EPI_BIAS, false, true>(h, m0, m0, mw, hg, ep, s, um);
        case RAMNET_EPI_BIAS_RELU: return launch_halo<RAMNET_EPI_BIAS_RELU, false, true>(h, m0, m0, mw, hg, ep, s, um);
        case RAMNET_EPI_BIAS_RELU_ADD: return launch_halo<RAMNET_EPI_BIAS_RELU_ADD, false, true>(h, m0, m0, mw, hg, ep, s, um);
        case RAMNET_EPI_BIAS_RELU_PRED: return launch_halo<RAMNET_EPI_BIAS_RELU_PRED, false, true>(h, m0, m0, mw, hg, ep, s, um);
    }
    return ramnet_set_error(RAMNET_EINVAL, "conv_fwd(upconv): unreachable");
}

// ================================================================================================
// Stride-2 5x5 convolution as four parity-plane K segments (RAMNET_FLAG_S2SEG, round 2).  Input row 2*oy + r - pad =
// 2*(oy + dy) + py: the taps with the same (py, px) form a stride-1 3x3 / 3x2 / 2x3 / 2x2 convolution of ONE parity
// plane x[:, py::2, px::2].  The round-1 stride-2 path fetched all four planes of the halo as one stage (166 KB for a
// 2-tile patch: a single stage, the MMA warp idle a third of the time waiting for it); here every plane is its own K
// segment read through its own tensor map (a strided view of the same tensor), so a stage is one plane (41 KB) and the
// stages pipeline.  Same 25 taps; the (1,1) plane's 4 taps are padded to 6 so that every segment is a multiple of the
// 3 taps a weight stage holds.
// ================================================================================================
constexpr int kS2Taps[4] = {9, 6, 6, 6};
constexpr int kS2TotalTaps = 27;
__host__ __device__ inline void s2seg_shape(int sg, int &kh, int &kw) { kh = (sg & 2) ? 2 : 3; kw = (sg & 1) ? 2 : 3; }

__global__ void pack_s2seg_kernel(const float *__restrict__ w, float *__restrict__ out, int Cout, int Cin) {
    const int64_t total = (int64_t)kS2TotalTaps * Cout * Cin;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        const int co = (int)((i / Cin) % Cout);
        int tap = (int)(i / ((int64_t)Cin * Cout));
        int sg = 0;
        while (sg < 3 && tap >= (sg == 0 ? 9 : 6)) { tap -= (sg == 0 ? 9 : 6); ++sg; }
        int kh, kw;
        s2seg_shape(sg, kh, kw);
        float v = 0.f;
        if (tap < kh * kw) {
            const int py = sg >> 1, px = sg & 1;
            const int dy = -1 + tap / kw, dx = -1 + tap % kw;          // plane offsets; every segment starts at -1
            const int r = 2 * dy + py + 2, sx = 2 * dx + px + 2;        // filter tap of the 5x5, pad 2
            v = w[(((int64_t)co * Cin + ci) * 5 + r) * 5 + sx];
        }
        out[i] = round_tf32(v);
    }
}

int conv_s2seg_fwd_tf32(ramnet_handle *h, const ramnet_conv_desc *d, const float *x0, const float *wp, const EpiParams &ep,
                        cudaStream_t s) {
    RAMNET_CHECK_ARG(d->ksize == 5 && d->stride == 2 && d->C1 == 0 && d->H >= 2 && d->W >= 2,
                     "conv_fwd(s2seg): 5x5 stride-2 single-source layers only");
    RAMNET_CHECK_ARG(d->C0 % kChunk == 0 && d->Cout % 16 == 0, "conv_fwd(s2seg): C0 %% 32, Cout %% 16");
    RAMNET_CHECK_ARG(d->epilogue == RAMNET_EPI_BIAS_RELU || d->epilogue == RAMNET_EPI_BIAS, "conv_fwd(s2seg): bias / relu epilogues");
    ramnet_conv_desc dp = *d;           // the planner sees a 3x3 stride-1 convolution over one parity plane
    dp.H = (d->H + 1) / 2; dp.W = (d->W + 1) / 2; dp.stride = 1; dp.ksize = 3;      // = the output size
    static const int pair_mode = [] { const char *e = getenv("RAMNET_PAIR"); return e ? atoi(e) : 1; }();
    static const int shapes[][2] = {{2, 1}, {1, 1}, {4, 1}, {2, 2}, {1, 2}};
    double best = -1;
    HaloGeom hg, cand;
    if (const char *f = getenv("RAMNET_S2_FORCE")) {      // tuning aid: "PTX,PTY,BN,pair"
        int ptx, pty, bn, pr = 0;
        if (sscanf(f, "%d,%d,%d,%d", &ptx, &pty, &bn, &pr) >= 3 && fill_halo(&dp, nullptr, &cand, ptx, pty, bn, 0, 0, pr, 3)) {
            best = 0; hg = cand;
        }
    }
    for (int pass = 0; pass < 2 && best < 0; ++pass)
        for (int pair = (pair_mode == 2 && pass == 0 ? 1 : 0); pair <= (pair_mode >= 1 ? 1 : 0); ++pair)
            for (const auto &sh : shapes)
                for (int bn = 256; bn >= 16; bn >>= 1) {
                    if (!fill_halo(&dp, nullptr, &cand, sh[0], sh[1], bn, 0, 0, pair, 3)) continue;
                    const double c = halo_cost(h, &dp, cand, kS2TotalTaps, true);
                    if (best < 0 || c < best) { best = c; hg = cand; }
                }
    if (best < 0) return RAMNET_EUNSUPPORTED;
    hg.up = 0; hg.up_cout = d->Cout; hg.nseg = 4; hg.total_taps = kS2TotalTaps;
    int tap0 = 0;
    for (int sg = 0; sg < 4; ++sg) {
        UpSeg &u = hg.seg[sg];
        s2seg_shape(sg, u.kh, u.kw);
        u.lo_y = -1; u.lo_x = -1; u.vx = 0; u.vy = 0; u.cond = 0;
        u.tap0 = tap0; u.ntaps = kS2Taps[sg]; tap0 += kS2Taps[sg];
    }
    if (getenv("RAMNET_DEBUG"))
        fprintf(stderr, "[ramnet] s2seg plan %dx%d C=%d->%d: tiles %dx%d HX=%d HY=%d BN=%d a_st=%d b_st=%d tpg=%d nbuf=%d items=%d pair=%d\n",
                d->H, d->W, d->C0, d->Cout, hg.PTX, hg.PTY, hg.HX, hg.HY, hg.BN, hg.a_stages, hg.b_stages, hg.tpg, hg.nbuf,
                hg.items, hg.pair);
    CUtensorMap m0, mw;
    UpMaps um;
    const int C = d->C0;
    const cuuint64_t str[3] = {(cuuint64_t)2 * C * 4, (cuuint64_t)2 * d->W * C * 4, (cuuint64_t)d->H * d->W * C * 4};
    const cuuint32_t box[4] = {kChunk, (cuuint32_t)hg.HX, (cuuint32_t)hg.HY, 1};
    for (int sg = 0; sg < 4; ++sg) {
        const int py = sg >> 1, px = sg & 1;
        // plane (py, px) = x[:, py::2, px::2]; with odd H or W the odd plane is one row / column shorter and the box's
        // out-of-bounds zero fill supplies exactly the padding the full-resolution convolution would read there
        const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)(d->W - px + 1) / 2, (cuuint64_t)(d->H - py + 1) / 2, (cuuint64_t)d->N};
        const float *base = x0 + ((int64_t)py * d->W + px) * C;
        const int rc = encode(h, sg == 0 ? &m0 : &um.m[sg - 1], base, 4, dims, str, box);
        if (rc) return rc;
    }
    const cuuint64_t wd[3] = {(cuuint64_t)C, (cuuint64_t)d->Cout, (cuuint64_t)kS2TotalTaps};
    const cuuint64_t wst[2] = {(cuuint64_t)C * 4, (cuuint64_t)C * d->Cout * 4};
    const cuuint32_t wb[3] = {kChunk, (cuuint32_t)(hg.pair ? hg.BN / 2 : hg.BN), (cuuint32_t)hg.tpg};
    int rc = encode(h, &mw, wp, 3, wd, wst, wb);
    if (rc) return rc;
    if (d->epilogue == RAMNET_EPI_BIAS) return launch_halo<RAMNET_EPI_BIAS, false, true>(h, m0, m0, mw, hg, ep, s, um);
    return launch_halo<RAMNET_EPI_BIAS_RELU, false, true>(h, m0, m0, mw, hg, ep, s, um);
}

int conv_fwd_tf32_rect(ramnet_handle *h, const ramnet_conv_desc *d, const RectSpec *rect, const float *x0, const float *x1,
                       const float *wp, const EpiParams &ep, cudaStream_t s);

int conv_fwd_tf32(ramnet_handle *h, const ramnet_conv_desc *d, const float *x0, const float *x1, const float *wp,
                  const EpiParams &ep, void *workspace, size_t ws_bytes, cudaStream_t s) {
    // RAMNET_FLAG_DYNAMIC: the first 8 bytes of the workspace are this launch's {item counter, finished workers} pair
    tl_sched_slot = ((d->flags & RAMNET_FLAG_DYNAMIC) && workspace && ws_bytes >= 8 && (((uintptr_t)workspace) & 3) == 0)
                        ? static_cast<int *>(workspace) : nullptr;
    const int rc = conv_fwd_tf32_rect(h, d, nullptr, x0, x1, wp, ep, s);
    tl_sched_slot = nullptr;
    return rc;
}

int conv_fwd_tf32_rect(ramnet_handle *h, const ramnet_conv_desc *d, const RectSpec *rect, const float *x0, const float *x1,
                       const float *wp, const EpiParams &ep, cudaStream_t s) {
    RAMNET_CHECK_ARG(d->C0 % kChunk == 0 && d->C1 % kChunk == 0,
                     "conv_fwd(tf32): C0=%d, C1=%d must be multiples of 32 (one 128-byte swizzle row)", d->C0, d->C1);
    RAMNET_CHECK_ARG(d->Cout % 16 == 0, "conv_fwd(tf32): Cout=%d must be a multiple of 16", d->Cout);
    RAMNET_CHECK_ARG(d->stride == 1 || (d->flags & RAMNET_FLAG_S2SEG) || (d->H % 2 == 0 && d->W % 2 == 0),
                     "conv_fwd(tf32): stride 2 needs even H, W (or RAMNET_FLAG_S2SEG weights)");
    RAMNET_CHECK_ARG((((uintptr_t)x0 | (uintptr_t)x1 | (uintptr_t)wp) & 15) == 0, "conv_fwd(tf32): 16-byte alignment");
    RAMNET_CHECK_ARG((((uintptr_t)ep.y0 | (uintptr_t)ep.y1 | (uintptr_t)ep.y2 | (uintptr_t)ep.aux0 | (uintptr_t)ep.aux1) & 31) == 0 ||
                         d->epilogue == RAMNET_EPI_BIAS_RELU_PRED,
                     "conv_fwd(tf32): outputs and epilogue operands must be 32-byte aligned (256-bit stores)");
    if (d->flags & RAMNET_FLAG_UPCONV) {
        RAMNET_CHECK_ARG(!rect && !x1, "conv_fwd(upconv): single source, no rectangular tap set");
        return conv_up_fwd_tf32(h, d, x0, wp, ep, s);
    }
    if (d->flags & RAMNET_FLAG_S2SEG) {
        RAMNET_CHECK_ARG(!rect && !x1, "conv_fwd(s2seg): single source, no rectangular tap set");
        const int rc = conv_s2seg_fwd_tf32(h, d, x0, wp, ep, s);
        if (rc == RAMNET_EUNSUPPORTED) return ramnet_set_error(rc, "conv_fwd(s2seg): no halo configuration fits");
        return rc;
    }
    HaloGeom hg;
    const bool want_hpack = (d->flags & RAMNET_FLAG_HPACK) != 0;
    if (want_hpack) {
        const bool epi_ok = d->epilogue == RAMNET_EPI_BIAS || d->epilogue == RAMNET_EPI_BIAS_RELU ||
                            d->epilogue == RAMNET_EPI_BIAS_RES_RELU || d->epilogue == RAMNET_EPI_BIAS_RELU_PRED;
        if (rect || !epi_ok || !plan_hpack(h, d, &hg))
            return ramnet_set_error(RAMNET_EUNSUPPORTED, "conv_fwd: RAMNET_FLAG_HPACK weights but no hpack configuration for "
                                    "this layer (stride 1, ksize 3/5, bias / relu / residual / pred epilogues only)");
    }
    const bool halo_ok = want_hpack || plan_halo(h, d, rect, &hg);
    if (!halo_ok && rect)
        return ramnet_set_error(RAMNET_EUNSUPPORTED, "conv_fwd: no halo-kernel configuration for a rectangular / strided launch");
    if (!halo_ok && (d->epilogue == RAMNET_EPI_BIAS_RELU_PRED || d->epilogue == RAMNET_EPI_BIAS_RELU_ADD ||
                     d->epilogue == RAMNET_EPI_BIAS_ADD))
        return ramnet_set_error(RAMNET_EUNSUPPORTED, "conv_fwd: no halo-kernel configuration for the fused prediction / add epilogues");
    if (halo_ok) {
        if (getenv("RAMNET_DEBUG"))
            fprintf(stderr, "[ramnet] halo plan s%d %dx%d C=%d+%d->%d k%d: tiles %dx%d HX=%d HY=%d BN=%d a_st=%d b_st=%d tpg=%d nbuf=%d items=%d pair=%d hpack=%d\n",
                    d->stride, d->H, d->W, d->C0, d->C1, d->Cout, d->ksize, hg.PTX, hg.PTY, hg.HX, hg.HY, hg.BN, hg.a_stages,
                    hg.b_stages, hg.tpg, hg.nbuf, hg.items, hg.pair, hg.hpack);
        CUtensorMap m0, m1, mw;
        auto enc_act = [&](CUtensorMap *m, const float *x, int C) {
            if (d->stride == 1) {
                cuuint32_t box[4] = {kChunk, (cuuint32_t)hg.HX, (cuuint32_t)hg.HY, 1};
                cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
                cuuint64_t str[3] = {(cuuint64_t)C * 4, (cuuint64_t)d->W * C * 4, (cuuint64_t)d->H * d->W * C * 4};
                return encode(h, m, x, 4, dims, str, box);
            }
            // stride 2: 5-D parity view (2C, W/2, 2, H/2, N); one box = one parity plane of the halo
            cuuint32_t box[5] = {kChunk, (cuuint32_t)hg.HX, 1, (cuuint32_t)hg.HY, 1};
            cuuint64_t dims[5] = {(cuuint64_t)2 * C, (cuuint64_t)d->W / 2, 2, (cuuint64_t)d->H / 2, (cuuint64_t)d->N};
            cuuint64_t str[4] = {(cuuint64_t)2 * C * 4, (cuuint64_t)d->W * C * 4, (cuuint64_t)2 * d->W * C * 4,
                                 (cuuint64_t)d->H * d->W * C * 4};
            return encode(h, m, x, 5, dims, str, box);
        };
        int rc = enc_act(&m0, x0, d->C0);
        if (rc) return rc;
        if (x1) {
            rc = enc_act(&m1, x1, d->C1);
            if (rc) return rc;
        } else {
            m1 = m0;
        }
        const int Ct = d->C0 + d->C1, taps = hg.kh * hg.kw;
        const int wrows = hg.hpack ? d->Cout * hg.kwp : d->Cout;        // hpack: rows (slice, tap s, channel) per filter row
        cuuint64_t wd[3] = {(cuuint64_t)Ct, (cuuint64_t)wrows, (cuuint64_t)taps};
        cuuint64_t ws[2] = {(cuuint64_t)Ct * 4, (cuuint64_t)Ct * wrows * 4};
        cuuint32_t wb[3] = {kChunk, (cuuint32_t)(hg.pair ? hg.BN / 2 : hg.BN), (cuuint32_t)hg.tpg};
        rc = encode(h, &mw, wp, 3, wd, ws, wb);
        if (rc) return rc;
        if (hg.hpack) {
            switch (d->epilogue) {
                case RAMNET_EPI_BIAS: return launch_halo<RAMNET_EPI_BIAS, true>(h, m0, m1, mw, hg, ep, s);
                case RAMNET_EPI_BIAS_RELU: return launch_halo<RAMNET_EPI_BIAS_RELU, true>(h, m0, m1, mw, hg, ep, s);
                case RAMNET_EPI_BIAS_RES_RELU: return launch_halo<RAMNET_EPI_BIAS_RES_RELU, true>(h, m0, m1, mw, hg, ep, s);
                case RAMNET_EPI_BIAS_RELU_PRED: return launch_halo<RAMNET_EPI_BIAS_RELU_PRED, true>(h, m0, m1, mw, hg, ep, s);
                default: return ramnet_set_error(RAMNET_EUNSUPPORTED, "conv_fwd: hpack epilogue");
            }
        }
        switch (d->epilogue) {
            case RAMNET_EPI_BIAS: return launch_halo<RAMNET_EPI_BIAS>(h, m0, m1, mw, hg, ep, s);
            case RAMNET_EPI_BIAS_RELU: return launch_halo<RAMNET_EPI_BIAS_RELU>(h, m0, m1, mw, hg, ep, s);
            case RAMNET_EPI_BIAS_RES_RELU: return launch_halo<RAMNET_EPI_BIAS_RES_RELU>(h, m0, m1, mw, hg, ep, s);
            case RAMNET_EPI_GRU_RU: return launch_halo<RAMNET_EPI_GRU_RU>(h, m0, m1, mw, hg, ep, s);
            case RAMNET_EPI_GRU_OUT: return launch_halo<RAMNET_EPI_GRU_OUT>(h, m0, m1, mw, hg, ep, s);
            case RAMNET_EPI_LSTM: return launch_halo<RAMNET_EPI_LSTM>(h, m0, m1, mw, hg, ep, s);
            case RAMNET_EPI_BIAS_RELU_PRED: return launch_halo<RAMNET_EPI_BIAS_RELU_PRED>(h, m0, m1, mw, hg, ep, s);
            case RAMNET_EPI_BIAS_RELU_ADD: return launch_halo<RAMNET_EPI_BIAS_RELU_ADD>(h, m0, m1, mw, hg, ep, s);
            case RAMNET_EPI_BIAS_ADD: return launch_halo<RAMNET_EPI_BIAS_ADD>(h, m0, m1, mw, hg, ep, s);
        }
    }
    TcGeom g;
    g.N = d->N; g.Cout = d->Cout; g.C0 = d->C0; g.C1 = d->C1;
    g.ks = d->ksize; g.stride = d->stride; g.pad = d->ksize / 2;
    g.Ho = conv_out_dim(d->H, d->stride); g.Wo = conv_out_dim(d->W, d->stride);
    pick_tile(g.Ho, g.Wo, &g.TW, &g.TH);
    g.tiles_x = (g.Wo + g.TW - 1) / g.TW; g.tiles_y = (g.Ho + g.TH - 1) / g.TH;
    // GEMM columns per CTA: the largest UMMA N dividing Cout that still yields >= one CTA per SM
    const int64_t mtiles = (int64_t)g.tiles_x * g.tiles_y * g.N;
    RAMNET_CHECK_ARG(mtiles <= 0x7fffffff, "conv_fwd(tf32): too many tiles");
    g.BN = 0;
    for (int bn = 256; bn >= 16; bn >>= 1) {
        if (d->Cout % bn) continue;
        if (g.BN == 0) g.BN = bn;
        if (mtiles * (d->Cout / bn) >= h->sm_count) { g.BN = bn; break; }
        g.BN = bn;
        if (bn <= 64) break;   // do not shrink below 64 columns just to fill SMs
    }
    RAMNET_CHECK_ARG(g.BN >= 16, "conv_fwd(tf32): no UMMA N divides Cout=%d", d->Cout);
    if (d->epilogue == RAMNET_EPI_GRU_RU || d->epilogue == RAMNET_EPI_LSTM)
        RAMNET_CHECK_ARG(g.BN % 16 == 0, "conv_fwd(tf32): gate epilogues need 16-column groups");
    const int stage_bytes = kABytes + g.BN * kChunk * 4;
    g.stages = g.BN > 128 ? 4 : (100 * 1024) / stage_bytes;   // <=128 columns: two CTAs per SM
    if (g.stages > 8) g.stages = 8;
    if (g.stages < 2) g.stages = 2;

    CUtensorMap m0, m1, mw;
    int rc = encode_activation(h, &m0, x0, d->N, d->H, d->W, d->C0, d->stride, g.TW, g.TH);
    if (rc) return rc;
    if (x1) {
        rc = encode_activation(h, &m1, x1, d->N, d->H, d->W, d->C1, d->stride, g.TW, g.TH);
        if (rc) return rc;
    } else {
        m1 = m0;
    }
    const int Ct = d->C0 + d->C1, taps = d->ksize * d->ksize;
    cuuint64_t wd[3] = {(cuuint64_t)Ct, (cuuint64_t)d->Cout, (cuuint64_t)taps};
    cuuint64_t ws[2] = {(cuuint64_t)Ct * 4, (cuuint64_t)Ct * d->Cout * 4};
    cuuint32_t wb[3] = {kChunk, (cuuint32_t)g.BN, 1};
    rc = encode(h, &mw, wp, 3, wd, ws, wb);
    if (rc) return rc;

    switch (d->epilogue) {
        case RAMNET_EPI_BIAS: return launch<RAMNET_EPI_BIAS>(h, m0, m1, mw, g, ep, s);
        case RAMNET_EPI_BIAS_RELU: return launch<RAMNET_EPI_BIAS_RELU>(h, m0, m1, mw, g, ep, s);
        case RAMNET_EPI_BIAS_RES_RELU: return launch<RAMNET_EPI_BIAS_RES_RELU>(h, m0, m1, mw, g, ep, s);
        case RAMNET_EPI_GRU_RU: return launch<RAMNET_EPI_GRU_RU>(h, m0, m1, mw, g, ep, s);
        case RAMNET_EPI_GRU_OUT: return launch<RAMNET_EPI_GRU_OUT>(h, m0, m1, mw, g, ep, s);
        case RAMNET_EPI_LSTM: return launch<RAMNET_EPI_LSTM>(h, m0, m1, mw, g, ep, s);
    }
    return ramnet_set_error(RAMNET_EINVAL, "conv_fwd(tf32): unreachable");
}


// ================================================================================================
// Sub-pixel data gradient of a stride-2 convolution (replaces zero insertion + a full-resolution conv, which spends
// 4x the MACs on inserted zeros).  dX[y][x] = sum_{r,s} dZ[(y + pad - r)/2][(x + pad - s)/2] W[r][s] over the taps
// for which the divisions are exact: with y = 2y' + py and r - pad = 2dy + py, the input pixels of row parity py see
// only the filter rows {r : (r - pad) mod 2 = py} and dX_py[y'] = sum_dy dZ[y' - dy] W[r(dy)].  Each of the 4 input
// parities is therefore a stride-1 convolution of dZ with a 3x3 / 3x2 / 2x3 / 2x2 sub-filter (5x5, pad 2) whose output
// is written into one parity plane of dX: four halo-kernel launches with a RectSpec.
// ================================================================================================
namespace {
struct S2Axis { int k[2], rmin[2], dmin[2]; };
__host__ __device__ inline S2Axis s2_axis(int ks) {
    S2Axis a;
    const int pad = ks / 2;
    for (int q = 0; q < 2; ++q) {
        a.rmin[q] = (pad + q) & 1;
        a.k[q] = (ks - 1 - a.rmin[q]) / 2 + 1;
        const int e = a.rmin[q] - pad - q;               // even
        a.dmin[q] = e >= 0 ? e / 2 : -((-e) / 2);
    }
    return a;
}

// w_oihw [Cout][Cin][ks][ks] -> for class c = py*2+px: [tap (a, b)][ci_count][Cout] (K-major B operand of the data
// gradient GEMM, K = Cout), Wsub[a][b] = W[rmin_py + 2(kh-1-a)][rmin_px + 2(kw-1-b)], classes back to back.
__global__ void pack_dgrad_s2_kernel(const float *__restrict__ w, float *__restrict__ out, int Cout, int Cin, int ks,
                                     int ci_begin, int ci_count) {
    const S2Axis ax = s2_axis(ks);
    const int taps = ks * ks;
    const int64_t per_tap = (int64_t)ci_count * Cout, total = per_tap * taps;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int t = (int)(i / per_tap);                       // tap index in the concatenated class order
        const int64_t rem = i - (int64_t)t * per_tap;
        const int cil = (int)(rem / Cout), co = (int)(rem % Cout);
        int cls = 0;
        for (; cls < 4; ++cls) {
            const int n = ax.k[cls >> 1] * ax.k[cls & 1];
            if (t < n) break;
            t -= n;
        }
        const int py = cls >> 1, px = cls & 1, kw = ax.k[px];
        const int a = t / kw, b = t % kw;
        const int r = ax.rmin[py] + 2 * (ax.k[py] - 1 - a), sx = ax.rmin[px] + 2 * (kw - 1 - b);
        out[i] = round_tf32(w[((int64_t)co * Cin + ci_begin + cil) * taps + r * ks + sx]);
    }
}
}  // namespace

extern "C" int ramnet_pack_weights_dgrad_s2(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin,
                                            int ksize, int ci_begin, int ci_count, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && w_oihw && w_packed && Cout > 0 && Cin > 0 && (ksize == 3 || ksize == 5) && ci_begin >= 0 &&
                         ci_count > 0 && ci_begin + ci_count <= Cin,
                     "pack_weights_dgrad_s2: bad argument");
    const int64_t total = (int64_t)Cout * ci_count * ksize * ksize;
    const int blocks = (int)imin64((total + 255) / 256, (int64_t)h->sm_count * 8);
    pack_dgrad_s2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, w_packed, Cout, Cin, ksize, ci_begin, ci_count);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_conv_dgrad_s2(ramnet_handle *h, const float *dz, const float *w_packed_s2, float *dx, int N, int H,
                                    int W, int Cout, int ci_count, int ksize, int flags, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && dz && w_packed_s2 && dx, "conv_dgrad_s2: NULL argument");
    RAMNET_CHECK_ARG(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "conv_dgrad_s2: input H=%d, W=%d must be even", H, W);
    RAMNET_CHECK_ARG((ksize == 3 || ksize == 5) && Cout % 32 == 0 && ci_count % 32 == 0,
                     "conv_dgrad_s2: ksize 3/5, Cout and ci_count multiples of 32 (got k=%d, Cout=%d, ci_count=%d)", ksize, Cout, ci_count);
    const S2Axis ax = s2_axis(ksize);
    ramnet_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.N = N; d.H = H / 2; d.W = W / 2; d.C0 = Cout; d.C1 = 0; d.Cout = ci_count; d.ksize = ksize; d.stride = 1;
    d.epilogue = RAMNET_EPI_BIAS; d.mma_kind = RAMNET_MMA_TF32; d.flags = flags;
    EpiParams ep{nullptr, nullptr, nullptr, dx, nullptr, nullptr, ci_count, flags};
    const float *wp = w_packed_s2;
    for (int cls = 0; cls < 4; ++cls) {
        const int py = cls >> 1, px = cls & 1;
        RectSpec r;
        r.kh = ax.k[py]; r.kw = ax.k[px];
        r.lo_y = -(ax.dmin[py] + r.kh - 1); r.lo_x = -(ax.dmin[px] + r.kw - 1);
        r.out_sy = r.out_sx = 2; r.out_oy = py; r.out_ox = px; r.out_H = H; r.out_W = W;
        const int rc = conv_fwd_tf32_rect(h, &d, &r, dz, nullptr, wp, ep, (cudaStream_t)stream);
        if (rc) return rc;
        wp += (size_t)r.kh * r.kw * ci_count * Cout;
    }
    return RAMNET_OK;
}


// ================================================================================================
// Head convolution on the tensor cores (TF32 mode, 5*Cin <= 32).
// The raw network input is NCHW with 1..6 channels: no 32-channel pixel rows for TMA / UMMA.  ramnet_head_im2row
// unrolls the five HORIZONTAL taps into the channel axis, Xe[n][y][x][dx*Cin + ci] = X[n][ci][y][x + dx - 2] (zero
// outside, zero-padded to 32 channels, rounded to TF32), which turns the 5x5 conv into a 5x1 conv over a 32-channel NHWC
// tensor: the halo kernel with a RectSpec {kh = 5, kw = 1}, K = 5 * 32, weights [r][co][dx*Cin + ci].  The unrolled
// tensor costs one extra 128 B/pixel write + read, far less than the fp32 FFMA pipe the direct kernel is bound by.
// ================================================================================================
namespace {
template <int CIN>
__global__ void __launch_bounds__(256) head_im2row_kernel(const float *__restrict__ x, float *__restrict__ xe, int N, int H,
                                                          int W) {
    __shared__ float tile[8][32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wtiles = (W + 31) / 32;
    const int64_t ntile = (int64_t)N * H * wtiles;
    for (int c = 5 * CIN; c < 32; ++c) tile[warp][lane][c] = 0.f;      // channel padding: written once
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < ntile; t += (int64_t)gridDim.x * 8) {
        const int xt = (int)(t % wtiles);
        const int y = (int)((t / wtiles) % H);
        const int n = (int)(t / ((int64_t)wtiles * H));
        const int px = xt * 32 + lane;
        const float *row = x + ((int64_t)n * CIN * H + y) * W;          // channel ci of this image row: row + ci*H*W
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
            const int sx = px + dx - 2;
            const bool ok = sx >= 0 && sx < W;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
                tile[warp][lane][dx * CIN + ci] = ok ? round_tf32(__ldg(row + (int64_t)ci * H * W + sx)) : 0.f;
        }
        __syncwarp();
        float4 *dst = reinterpret_cast<float4 *>(xe + (((int64_t)n * H + y) * W + (int64_t)xt * 32) * 32);
        const int npx = min(32, W - xt * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int f = i * 32 + lane, p = f >> 3, q = f & 7;      // 16-byte piece q of pixel p: 512 contiguous bytes per store
            if (p < npx) dst[f] = make_float4(tile[warp][p][4 * q], tile[warp][p][4 * q + 1], tile[warp][p][4 * q + 2],
                                              tile[warp][p][4 * q + 3]);
        }
        __syncwarp();
    }
}

// w_oihw [Cout][Cin][5][5] -> [r][Cout][32] with column dx*Cin + ci (K-major B operand), TF32-rounded, zero-padded
__global__ void pack_head_kernel(const float *__restrict__ w, float *__restrict__ out, int Cout, int Cin) {
    const int total = 5 * Cout * 32;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i & 31, co = (i >> 5) % Cout, r = i / (32 * Cout);
        float v = 0.f;
        if (c < 5 * Cin) {
            const int dx = c / Cin, ci = c - dx * Cin;
            v = round_tf32(w[(((int64_t)co * Cin + ci) * 5 + r) * 5 + dx]);
        }
        out[i] = v;
    }
}
}  // namespace

extern "C" int ramnet_head_im2row(ramnet_handle *h, const float *x_nchw, float *xe_nhwc32, int N, int Cin, int H, int W,
                                  void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && x_nchw && xe_nhwc32 && N > 0 && H > 0 && W > 0, "head_im2row: bad argument");
    RAMNET_CHECK_ARG(Cin >= 1 && 5 * Cin <= 32, "head_im2row: 5*Cin = %d must fit one 32-channel pixel row", 5 * Cin);
    const int64_t ntile = (int64_t)N * H * ((W + 31) / 32);
    const int blocks = (int)imin64((ntile + 7) / 8, (int64_t)h->sm_count * 8);
    cudaStream_t s = (cudaStream_t)stream;
    switch (Cin) {
        case 1: head_im2row_kernel<1><<<blocks, 256, 0, s>>>(x_nchw, xe_nhwc32, N, H, W); break;
        case 2: head_im2row_kernel<2><<<blocks, 256, 0, s>>>(x_nchw, xe_nhwc32, N, H, W); break;
        case 3: head_im2row_kernel<3><<<blocks, 256, 0, s>>>(x_nchw, xe_nhwc32, N, H, W); break;
        case 4: head_im2row_kernel<4><<<blocks, 256, 0, s>>>(x_nchw, xe_nhwc32, N, H, W); break;
        case 5: head_im2row_kernel<5><<<blocks, 256, 0, s>>>(x_nchw, xe_nhwc32, N, H, W); break;
        default: head_im2row_kernel<6><<<blocks, 256, 0, s>>>(x_nchw, xe_nhwc32, N, H, W); break;
    }
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_pack_weights_head(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin,
                                        void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && w_oihw && w_packed && Cout > 0 && Cin >= 1 && 5 * Cin <= 32, "pack_weights_head: bad argument");
    pack_head_kernel<<<(5 * Cout * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_oihw, w_packed, Cout, Cin);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_head_conv_tc(ramnet_handle *h, const float *xe_nhwc32, const float *w_packed, const float *bias,
                                   float *y_nhwc, int N, int H, int W, int Cout, int flags, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && xe_nhwc32 && w_packed && y_nhwc && N > 0 && H > 0 && W > 0, "head_conv_tc: bad argument");
    RAMNET_CHECK_ARG(Cout > 0 && Cout % 32 == 0, "head_conv_tc: Cout=%d must be a multiple of 32", Cout);
    ramnet_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.N = N; d.H = H; d.W = W; d.C0 = 32; d.C1 = 0; d.Cout = Cout; d.ksize = 5; d.stride = 1;
    d.epilogue = RAMNET_EPI_BIAS_RELU; d.mma_kind = RAMNET_MMA_TF32; d.flags = flags;
    RectSpec r;
    r.kh = 5; r.kw = 1; r.lo_y = -2; r.lo_x = 0;
    r.out_sy = r.out_sx = 1; r.out_oy = r.out_ox = 0; r.out_H = H; r.out_W = W;
    EpiParams ep{bias, nullptr, nullptr, y_nhwc, nullptr, nullptr, Cout, flags};
    return conv_fwd_tf32_rect(h, &d, &r, xe_nhwc32, nullptr, w_packed, ep, (cudaStream_t)stream);
}


// ================================================================================================
// Planner introspection (host only, no CUDA call): what the forward halo kernel and the tap-packed weight gradient
// would launch for a layer on a device with `sm_count` SMs.  Used by tests/test_planner.py to check the planners'
// invariants (shared memory, TMEM columns, divisibility) over a sweep of shapes without a GPU, and by humans to see
// why a layer got its configuration.  Writes one line of `key=value` pairs; returns its length (0 when nothing fits).
// ================================================================================================
extern "C" int ramnet_plan_describe(const ramnet_conv_desc *d, int sm_count, char *buf, size_t buf_bytes) {
    if (!d || !buf || buf_bytes == 0 || sm_count <= 0) return 0;
    ramnet_handle h;
    memset(&h, 0, sizeof(h));
    h.device = -1;
    h.sm_count = sm_count;
    int n = 0;
    auto put = [&](const char *fmt, auto... args) {
        if ((size_t)n < buf_bytes) n += snprintf(buf + n, buf_bytes - (size_t)n, fmt, args...);
    };
    buf[0] = 0;
    HaloGeom g;
    const bool tf32_ok = d->C0 % kChunk == 0 && d->C1 % kChunk == 0 && d->Cout % 16 == 0 && d->C0 > 0 &&
                         (d->stride == 1 || (d->H % 2 == 0 && d->W % 2 == 0));
    if (tf32_ok && plan_halo(&h, d, nullptr, &g)) {
        const size_t a_stride = (size_t)g.nplanes * g.plane_stride;
        const size_t smem = g.a_stages * a_stride + (size_t)g.b_stages * g.tpg * (g.pair ? g.BN / 2 : g.BN) * kChunk * 4 +
                            (2 * g.a_stages + 2 * g.b_stages + 4) * 8 + 16 + 32 * 4 + 1024;
        put("halo=1 pair=%d ptx=%d pty=%d hx=%d hy=%d bn=%d a_stages=%d b_stages=%d tpg=%d taps=%d nbuf=%d items=%d "
            "tmem_cols=%d smem=%zu nplanes=%d ", g.pair, g.PTX, g.PTY, g.HX, g.HY, g.BN, g.a_stages, g.b_stages, g.tpg,
            g.kh * g.kw, g.nbuf, g.items, g.nbuf * g.PTX * g.PTY * g.BN, smem, g.nplanes);
    } else {
        put("halo=0 ");
    }
    WpBatch b;
    int splits = 0, groups = 0;
    if (plan_wgrad_batch(&h, d, &b, &splits, &groups)) {
        const WpGeom &p = b.g[0];
        const size_t smem = (size_t)p.stages * p.stage_bytes + (2 * p.stages + 1) * 8 + 16 + 1024;
        put("wgrad=1 problems=%d m_from_x=%d rg=%d tr=%d ncols=%d groups=%d splits=%d tiles_per_cta=%d total_tiles=%d stages=%d "
            "smem=%zu workspace=%zu hxw=%d hyw=%d", b.n, p.m_from_x, p.RG, p.TR, p.ncols, groups, splits, p.tiles_per_cta,
            p.total_tiles, p.stages, smem, (size_t)b.n * (size_t)b.part_stride * sizeof(float), p.HXw, p.HYw);
    } else {
        put("wgrad=0");
    }
    return n;
}


// w_oihw [Cout][Cin][ks][ks] -> hpack layout [r][slice * ks * cs + s * cs + co_l][Cin] (K-major B operand, TF32-rounded):
// the weight tile of filter row r holds, for every output-channel slice of cs channels, the ks horizontal taps as
// consecutive groups of cs GEMM columns.
namespace {
__global__ void pack_hpack_kernel(const float *__restrict__ w, float *__restrict__ out, int Cout, int Cin, int ks, int cs) {
    const int64_t total = (int64_t)ks * ks * Cout * Cin;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cin);
        int64_t t = i / Cin;
        const int row = (int)(t % ((int64_t)ks * Cout));      // slice * ks * cs + s * cs + co_l
        const int r = (int)(t / ((int64_t)ks * Cout));
        const int slice = row / (ks * cs), rem = row % (ks * cs), sx = rem / cs, col = rem % cs;
        const int co = slice * cs + col;
        out[i] = round_tf32(w[(((int64_t)co * Cin + c) * ks + r) * ks + sx]);
    }
}
}  // namespace

extern "C" int ramnet_pack_weights_s2seg(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin,
                                         void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && w_oihw && w_packed && Cout > 0 && Cout % 16 == 0 && Cin > 0 && Cin % 32 == 0,
                     "pack_weights_s2seg: bad argument");
    const int64_t total = (int64_t)kS2TotalTaps * Cout * Cin;
    pack_s2seg_kernel<<<(unsigned)imin64((total + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, (cudaStream_t)stream>>>(
        w_oihw, w_packed, Cout, Cin);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int64_t ramnet_upconv_packed_floats(int Cout, int Cin) { return (int64_t)kUpTotalTaps * 4 * Cout * Cin; }

extern "C" int ramnet_pack_weights_upconv(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin,
                                          void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && w_oihw && w_packed && Cout > 0 && Cout % 16 == 0 && Cin > 0 && Cin % 32 == 0,
                     "pack_weights_upconv: bad argument");
    const int64_t total = ramnet_upconv_packed_floats(Cout, Cin);
    pack_upconv_kernel<<<(unsigned)imin64((total + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, (cudaStream_t)stream>>>(
        w_oihw, w_packed, Cout, Cin);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

extern "C" int ramnet_pack_weights_hpack(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin,
                                         int ksize, void *stream) {
    RAMNET_DEVICE_GUARD(h);
    RAMNET_CHECK_ARG(h && w_oihw && w_packed && Cout > 0 && Cout % 16 == 0 && Cin > 0 && (ksize == 3 || ksize == 5),
                     "pack_weights_hpack: bad argument");
    const int cs = Cout % 32 == 0 ? 32 : 16;
    const int64_t total = (int64_t)ksize * ksize * Cout * Cin;
    pack_hpack_kernel<<<(unsigned)imin64((total + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, (cudaStream_t)stream>>>(
        w_oihw, w_packed, Cout, Cin, ksize, cs);
    RAMNET_LAUNCH_CHECK(h);
    return RAMNET_OK;
}

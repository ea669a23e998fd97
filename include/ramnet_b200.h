/* ramnet_b200.h — C ABI of libramnet_sm100a.so (B200 / sm_100a only).
 *
 * The drop-in boundary for the RAM-Net per-timestep hot path.
 * The reference (uzh-rpg/rpg_ramnet) has no native code: every device op it
 * issues is a stock ATen/cuDNN call made from Python.  Each entry point below
 * therefore names the reference Python call site(s) whose device work it
 * replaces.  The Python host side (rpg_ramnet_b200/) binds these with ctypes;
 * INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch types, no C++ in the signatures;
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns all memory (inputs, outputs, workspaces); the library
 *     never allocates on the hot path and never synchronises the host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     all work is enqueued on it, so every call is CUDA-graph capturable;
 *   - return value: 0 on success, negative RAMNET_E* otherwise; the message
 *     is available from ramnet_last_error() (thread-local, valid until the
 *     next failing call on that thread);
 *   - activations are fp32 NHWC ("pixel-major": [N, H, W, C], C contiguous);
 *     network inputs are fp32 NCHW exactly as the reference feeds them.
 */
#ifndef RAMNET_B200_H
#define RAMNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAMNET_ABI_VERSION 3

enum {
    RAMNET_OK = 0,
    RAMNET_EINVAL = -1,       /* bad argument / unsupported shape */
    RAMNET_ECUDA = -2,        /* CUDA runtime / driver error */
    RAMNET_EUNSUPPORTED = -3, /* valid request the library does not implement */
    RAMNET_EDEVICE = -4       /* not an sm_100 device */
};

/* Arithmetic of the dense contraction (ramnet_conv_desc.mma_kind). */
enum {
    RAMNET_MMA_FP32 = 0, /* CUDA-core FFMA, fp32 exact: strict-parity mode */
    RAMNET_MMA_TF32 = 1  /* tcgen05.mma kind::tf32, fp32 accumulate in TMEM */
};

/* Fused epilogues of the implicit-GEMM convolution. `acc` is the fp32
 * accumulator of GEMM column n (output channel) at output pixel m. */
enum {
    /* y0 = acc + b                                   (pred logits, generic) */
    RAMNET_EPI_BIAS = 0,
    /* y0 = relu(acc + b)        ConvLayer.forward, submodules.py:26-35       */
    RAMNET_EPI_BIAS_RELU = 1,
    /* y0 = relu(acc + b + aux0) ResidualBlock.forward tail, submodules.py:213-215 */
    RAMNET_EPI_BIAS_RES_RELU = 2,
    /* ConvGRU gates, submodules.py:446-448.  Cout = 2C, columns [0,C) are the
     * reset gate, [C,2C) the update gate.  aux0 = h (prev state, C channels).
     * y0 = u = sigmoid(acc_u + b_u)            [.., C]
     * y1 = h * sigmoid(acc_r + b_r)            [.., C]                        */
    RAMNET_EPI_GRU_RU = 3,
    /* ConvGRU candidate + blend, submodules.py:449-452.  aux0 = h, aux1 = u.
     * y0 = h*(1-u) + tanh(acc + b)*u                                          */
    RAMNET_EPI_GRU_OUT = 4,
    /* ConvLSTM, submodules.py:341-356.  Cout = 4C with columns interleaved at
     * pack time: column 4c+g, g = 0 in, 1 remember, 2 out, 3 cell (the
     * reference's chunk order, :344).  aux0 = c (prev cell).
     * y1 = c' = s(f)c + s(i)tanh(g);  y0 = h' = s(o)tanh(c')                 */
    RAMNET_EPI_LSTM = 5,
    /* Last decoder + prediction head fused (statenet.py:116-117,313 on top of submodules.py:89-95):
     * t = relu(acc + b) is never written; y0[m] = sigmoid(sum_n t[n]*aux0[n] + aux1[0]) (depth,
     * [N,1,H,W]); y1[m] (may be NULL) = the logit.  aux0 = pred weight [Cout], aux1 = pred bias [1].
     * TF32 path only, Cout a multiple of 32 and <= 256 (one Cout slice per tile). */
    RAMNET_EPI_BIAS_RELU_PRED = 6,
    /* y0 = relu(acc + b) + aux0: a decoder layer that also forms the NEXT decoder's skip sum (statenet.py:15-16,306-308,
     * `x + super_state`), so that the sum never needs its own pass.  aux0 has the output's shape. */
    RAMNET_EPI_BIAS_RELU_ADD = 7,
    /* y0 = acc + b + aux0 (no activation): a data gradient accumulated onto the gradient another path already
     * produced (ConvGRU: dx and dh each receive two contributions); y0 may alias aux0. */
    RAMNET_EPI_BIAS_ADD = 8
};

enum {
    RAMNET_FLAG_ROUND_TF32 = 1, /* round every stored output to TF32 (rna) so a
                                   following kind::tf32 MMA truncates nothing */
    RAMNET_FLAG_HPACK = 2,      /* ramnet_conv_fwd: w_packed is in ramnet_pack_weights_hpack's layout (horizontal taps as
                                   GEMM columns); stride 1, ksize 3/5, bias / relu / residual / pred epilogues.
                                   Default for Cout = 32 layers since round 2 (RAMNET_HPACK=0 on the Python side disables). */
    RAMNET_FLAG_UPCONV = 4,     /* ramnet_conv_fwd: bilinear x2 (align_corners=False) FOLLOWED BY the 5x5 stride-1 convolution
                                   (UpsampleConvLayer.forward, submodules.py:87-97) in one launch on the LOW-resolution
                                   input: desc.H/W are the input's, y0 is [N, Cout, 2H, 2W].  w_packed comes from
                                   ramnet_pack_weights_upconv (collapsed taps of the 4 output phases + border segments).
                                   Epilogues: BIAS_RELU, BIAS_RELU_ADD, BIAS_RELU_PRED.  TF32 path, x1 = NULL. */
    RAMNET_FLAG_S2SEG = 8,      /* ramnet_conv_fwd: 5x5 stride-2 convolution as four parity-plane K segments (multi-stage
                                   halos); w_packed from ramnet_pack_weights_s2seg ([27 taps][Cout][Cin], the (1,1) plane
                                   padded to 6 taps).  Epilogues BIAS, BIAS_RELU; TF32 path, x1 = NULL. */
    RAMNET_FLAG_SM_TIME = 16,   /* ramnet_conv_fwd planning hint: the caller runs this launch concurrently with kernels of
                                   another stream, so the tile plan minimises SM time (items / SMs) instead of the makespan
                                   of a kernel alone on the GPU (ceil(items / SMs) waves).  Results are unaffected. */
    RAMNET_FLAG_DYNAMIC = 32,   /* ramnet_conv_fwd (TF32): work items are drawn from a global counter instead of a static round
                                   robin, so that CTAs that start late (SMs still held by a kernel of another stream) take
                                   less.  `workspace` must then point at 8 bytes that are zero at launch and belong to this
                                   launch site alone (two ints: the kernel resets them when it ends).  Results are unaffected. */
};

/* One implicit-GEMM convolution:  y[m, n] = epi( sum_{tap,c} x[pix(m,tap), c] * w[tap, n, c] ).
 * The K dimension is the virtual concatenation [x0 | x1] along channels
 * (torch.cat(...,1) at submodules.py:340,445,449 without the copy).
 * Zero padding of ksize/2, cross-correlation (nn.Conv2d semantics). */
typedef struct ramnet_conv_desc {
    int32_t N, H, W;   /* input batch / height / width */
    int32_t C0, C1;    /* channels of x0, x1 (C1 = 0 when x1 is NULL) */
    int32_t Cout;      /* GEMM N (2C for GRU_RU, 4C for LSTM) */
    int32_t ksize;     /* 1, 3 or 5 */
    int32_t stride;    /* 1 or 2; Ho = (H-1)/stride+1 */
    int32_t epilogue;  /* RAMNET_EPI_* */
    int32_t mma_kind;  /* RAMNET_MMA_* */
    int32_t flags;     /* RAMNET_FLAG_* */
    int32_t reserved;
} ramnet_conv_desc;

typedef struct ramnet_handle ramnet_handle;

/* ---- library / handle ------------------------------------------------- */
int ramnet_version(void);
const char *ramnet_last_error(void);
/* Binds to CUDA device `device` (must be sm_100); caches SM count, the
 * cuTensorMapEncodeTiled entry point and per-kernel attributes. */
int ramnet_create(int device, ramnet_handle **out);
int ramnet_destroy(ramnet_handle *h);
int ramnet_sm_count(const ramnet_handle *h);
/* Kernels launched through this handle since creation (bench.py's gpu_launches). */
int64_t ramnet_launch_count(const ramnet_handle *h);

/* ---- a-1  events_to_voxel_grid ---------------------------------------- *
 * Replaces utils/event_tensor_utils.py:71-117 (numpy) and :120-187 (torch
 * index_add_ twin); live twin data_loader/dataset_asynchronous.py:253-298.
 * events: [n,4] float64 rows [t, x, y, p] (not modified, unlike the
 * reference).  grid: [bins, height, width] float32, zeroed by the callee.
 * Index arithmetic is int64 and bit-exact with the reference; votes whose
 * pixel falls outside the grid are dropped and counted in *oob_count
 * (device int32, may be NULL; the reference raises IndexError / wraps). */
int ramnet_voxel_grid(ramnet_handle *h, const double *events, int64_t n, int bins, int width,
                      int height, float *grid, int32_t *oob_count, void *stream);
/* Debug/parity twin: writes the un-accumulated vote stream instead of
 * scattering it (idx = -1 for a dropped vote). Arrays of length n. */
/* Same result with two extensions.  (1) `workspace` (nullable; ramnet_voxel_grid_workspace_bytes bytes, 16-byte
 * aligned): for large event counts the votes accumulate in a pixel-major [H*W][8] buffer where both votes of an event
 * are one 128-bit vector reduction, then a second pass lays the grid out as [bins, H, W] (bins <= 8).  (2) `stats`
 * (nullable, 3 doubles, zeroed by the callee): sum, sum of squares and count of the NON-ZERO voxels, gathered while the
 * grid is produced; with RAMNET_VOXEL_NORMALIZE the grid is then normalised in place exactly as the loaders do
 * (data_loader/event_dataset.py:144-151, dataset_asynchronous.py:300-308): scatter -> normalise in one call. */
#define RAMNET_VOXEL_NORMALIZE 1
size_t ramnet_voxel_grid_workspace_bytes(int num_bins, int width, int height);
int ramnet_voxel_grid_ex(ramnet_handle *h, const double *events, int64_t n, int num_bins, int width, int height,
                         float *grid, int32_t *oob_count, void *workspace, size_t workspace_bytes, double *stats,
                         int flags, void *stream);
int ramnet_voxel_votes(ramnet_handle *h, const double *events, int64_t n, int bins, int width,
                       int height, int64_t *idx_left, float *val_left, int64_t *idx_right,
                       float *val_right, void *stream);

/* ---- a-2  head convolution -------------------------------------------- *
 * Replaces ConvLayer.forward for head_events / head_rgb / unet.head
 * (statenet.py:139-145, unet.py:93-94): 5x5 s1 p2 conv + bias + ReLU on the
 * raw NCHW network input with Cin <= 8; writes NHWC [N,H,W,Cout].  fp32 FFMA
 * (HBM-bound: 12-54 FLOP/B).  w: [Cout, Cin, 5, 5] exactly as nn.Conv2d holds it. */
int ramnet_head_conv(ramnet_handle *h, const float *x_nchw, const float *w_oihw, const float *bias,
                     float *y_nhwc, int N, int Cin, int H, int W, int Cout, int flags, void *stream);

/* ---- a-3..a-7  implicit-GEMM convolution with fused epilogue ----------- *
 * Replaces nn.Conv2d + torch.cat + the pointwise gate kernels at
 * submodules.py:27-33 (ConvLayer), :340-356 (ConvLSTM), :445-452 (ConvGRU),
 * :202-215 (ResidualBlock), :89-95 (UpsampleConvLayer after upsample).
 * w_packed: see ramnet_pack_weights.  bias: [Cout] (NULL = 0).
 * workspace: ramnet_conv_workspace_bytes(desc) bytes (may be 0/NULL). */
size_t ramnet_conv_workspace_bytes(const ramnet_conv_desc *d);
int ramnet_conv_fwd(ramnet_handle *h, const ramnet_conv_desc *d, const float *x0, const float *x1,
                    const float *w_packed, const float *bias, const float *aux0, const float *aux1,
                    float *y0, float *y1, float *y2, void *workspace, size_t workspace_bytes, void *stream);
/* y2 (may be NULL) is the training stash: RAMNET_EPI_GRU_RU writes r = sigmoid(reset), RAMNET_EPI_GRU_OUT
 * writes o = tanh(candidate), RAMNET_EPI_LSTM writes the four post-activation gates [M, C, 4] — the values autograd
 * would have kept for backward. */
/* [Cout, Cin, k, k] fp32 (nn.Conv2d layout) -> the layout `mma_kind` consumes:
 *   FP32: [k*k][Cin][Cout]           TF32: [k*k][Cout][Cin], values rounded to TF32 (rna).
 * `lstm_interleave` != 0 permutes output channels to 4c+g (RAMNET_EPI_LSTM). */
/* [Cout, Cin, k, k] -> [r][slice * k * cs + s * cs + co_l][Cin] with cs = 32 (16 when Cout % 32 != 0), TF32-rounded:
 * the layout RAMNET_FLAG_HPACK launches read (csrc/conv_tcgen05.cu fill_hpack; tests/test_hpack_index_algebra.py). */
int ramnet_pack_weights_hpack(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin,
                              int ksize, void *stream);
/* [Cout, Cin, 5, 5] -> the RAMNET_FLAG_UPCONV layout [tap][4 * Cout][Cin], TF32-rounded: for every tap of the main
 * segment (5x5 window on the low-resolution input) and of the 8 border segments (first / last row, first / last column,
 * 4 corners) the 5x5 filter collapsed through the bilinear x2 coefficients, one block of Cout rows per output phase
 * (py, px).  ramnet_upconv_packed_floats gives the size of w_packed in floats. */
int ramnet_pack_weights_s2seg(ramnet_handle *h, const float *w_oihw, float *w_packed /* 27 * Cout * Cin floats */, int Cout,
                              int Cin, void *stream);
int64_t ramnet_upconv_packed_floats(int Cout, int Cin);
int ramnet_pack_weights_upconv(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin, void *stream);
int ramnet_pack_weights(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin,
                        int ksize, int mma_kind, int lstm_interleave, void *stream);

/* ---- a-7  decoder prologue --------------------------------------------- *
 * Replaces skip_sum (statenet.py:15-16,306-308) + f.interpolate(scale 2,
 * bilinear, align_corners=False) (submodules.py:88).  NHWC in [N,H,W,C]
 * (+ optional skip of the same shape) -> NHWC out [N,2H,2W,C]. */
int ramnet_upsample2x_add(ramnet_handle *h, const float *x, const float *skip, float *y, int N,
                          int H, int W, int C, int flags, void *stream);

/* ---- a-8  prediction head ---------------------------------------------- *
 * Replaces pred (1x1 conv, no activation, statenet.py:116-117) + torch.sigmoid
 * (:313); `skip` (may be NULL) is unet.py:129's x + head.  x: NHWC [M, C];
 * w: [C]; logits / depth: [M] (= [N,1,H,W]); either output may be NULL.
 * w_skip (nullable, [C]): skip_type 'concat' (unet.py:11-12): logits = x . w + skip . w_skip instead of (x + skip) . w,
 * i.e. the 1x1 conv over cat([x, skip]) without the concatenated tensor. */
int ramnet_pred_sigmoid(ramnet_handle *h, const float *x, const float *skip, const float *w, const float *w_skip,
                        const float *bias, float *logits, float *depth, int64_t M, int C, void *stream);

/* Tensor-core head conv (TF32 mode, 5*Cin <= 32): ramnet_head_im2row unrolls the five horizontal taps into a
 * 32-channel NHWC tensor xe[n][y][x][dx*Cin + ci] = x[n][ci][y][x+dx-2]; ramnet_head_conv_tc then runs the remaining
 * 5x1 conv (+bias, ReLU) on tcgen05 with ramnet_pack_weights_head weights.  Same reference call sites as
 * ramnet_head_conv. */
int ramnet_head_im2row(ramnet_handle *h, const float *x_nchw, float *xe_nhwc32, int N, int Cin, int H, int W,
                       void *stream);
int ramnet_pack_weights_head(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin,
                             void *stream);
int ramnet_head_conv_tc(ramnet_handle *h, const float *xe_nhwc32, const float *w_packed, const float *bias,
                        float *y_nhwc, int N, int H, int W, int Cout, int flags, void *stream);

/* ---- f-2 / f-3  loader wire format and trainer metrics on the device ---- *
 * ramnet_voxel_normalize: in place, mean / stddev of the NON-ZERO voxels -> (x - mean) / stddev on them
 *   (data_loader/event_dataset.py:144-151, dataset_asynchronous.py:300-308, utils/event_tensor_utils.py:52-66);
 *   `batch` grids of n voxels each, normalised independently in one launch pair; stats = 3 doubles per grid of
 *   device scratch (sum, sum of squares, count), left filled.
 * ramnet_depth_to_label: metric depth -> clip(1 + log(clip(d, 0, clip)/clip) / reg_factor, 0, 1), NaN preserved
 *   (data_loader/dataset.py:296-305).
 * ramnet_depth_metrics: masked error sums of model/metric.py:8-57 per sample over hw pixels, out[s*8 + k]:
 *   0 count(~nan |t-p|), 1 sum |d|/(t+eps), 2 sum d^2/(t^2+eps), 3 sum d^2, 4 sum |d|, 5 count(~nan t),
 *   6 sum (p-t)^2 over ~nan t, 7 unused (replaces the full-map D2H + numpy of trainer/lstm_trainer.py:100-106). */
int ramnet_voxel_normalize(ramnet_handle *h, float *grid, int64_t n, int batch, double *stats, void *stream);
int ramnet_depth_to_label(ramnet_handle *h, const float *depth, float *label, int64_t n, float clip_distance,
                          float reg_factor, void *stream);
int ramnet_depth_metrics(ramnet_handle *h, const float *pred, const float *target, int N, int64_t hw, float eps,
                         double *out, void *stream);

/* Planner introspection, host only (no CUDA call, no handle): the configuration the forward halo kernel and the
 * tap-packed weight gradient would use for `d` on a device with `sm_count` SMs, as one line of key=value pairs
 * written to buf; returns its length.  For tests (planner invariants without a GPU) and tuning. */
int ramnet_plan_describe(const ramnet_conv_desc *d, int sm_count, char *buf, size_t buf_bytes);

/* ---- layout helpers ----------------------------------------------------- */
int ramnet_nchw_to_nhwc(ramnet_handle *h, const float *x, float *y, int N, int C, int H, int W,
                        int flags, void *stream);
int ramnet_nhwc_to_nchw(ramnet_handle *h, const float *x, float *y, int N, int C, int H, int W,
                        void *stream);
int ramnet_round_tf32(ramnet_handle *h, const float *x, float *y, int64_t n, void *stream);

/* ---- a-13  backward building blocks -------------------------------------- *
 * Replace what loss.backward() (trainer/lstm_trainer.py:450) dispatches to cuDNN/ATen:
 * data gradient = ramnet_conv_fwd on dZ with ramnet_pack_weights_dgrad weights (stride 2: after
 * ramnet_zero_insert2x); weight/bias gradient = ramnet_conv_wgrad, ACCUMULATED (+=) into buffers in
 * nn.Conv2d layout [Cout, C0+C1, k, k] / [Cout]; the rest are pointwise adjoints of the fused epilogues.
 * dz: gradient w.r.t. the GEMM output (pre-activation), NHWC [N, Ho, Wo, Cout]. */
size_t ramnet_conv_wgrad_workspace_bytes(const ramnet_handle *h, const ramnet_conv_desc *d);
/* `mode`: BPTT produces the weight gradient of one layer once per pass (L * (K+1) times per step, trainer/lstm_trainer.py:
 * 256-272 then :450).  FULL does everything per call.  The deferred modes keep the layer's partial tiles in a
 * caller-owned `workspace` that lives for the whole step: PARTIAL_FIRST overwrites it, PARTIAL_ADD accumulates into it
 * (both run only the tensor-core kernel; dw / db are not touched), FINALIZE (dz, x0, x1 ignored) runs the split sum +
 * scatter once and accumulates into dw.  The deferred modes return RAMNET_EUNSUPPORTED for shapes the tap-packed
 * TF32 kernel does not cover (use FULL there). */
enum { RAMNET_WGRAD_FULL = 0, RAMNET_WGRAD_PARTIAL_FIRST = 1, RAMNET_WGRAD_PARTIAL_ADD = 2, RAMNET_WGRAD_FINALIZE = 3 };
int ramnet_conv_wgrad(ramnet_handle *h, const ramnet_conv_desc *d, const float *dz, const float *x0,
                      const float *x1, float *dw_oihw, float *db, void *workspace, size_t workspace_bytes,
                      int mode, void *stream);
int ramnet_head_conv_wgrad(ramnet_handle *h, const float *x_nchw, const float *dz_nhwc, float *dw_oihw,
                           float *db, int N, int Cin, int H, int W, int Cout, void *stream);
/* Head conv weight gradient on the tensor cores (TF32 path): xe = ramnet_head_im2row's tensor, dw in the head's
 * [Cout][Cin][5][5] layout (Cout % 32 == 0), accumulated (+=). */
size_t ramnet_head_conv_wgrad_tc_workspace_bytes(const ramnet_handle *h, int N, int Cin, int H, int W, int Cout);
int ramnet_head_conv_wgrad_tc(ramnet_handle *h, const float *xe_nhwc32, const float *dz_nhwc, float *dw_oihw,
                              float *db, int N, int Cin, int H, int W, int Cout, void *workspace,
                              size_t workspace_bytes, int mode, void *stream);
int ramnet_pack_weights_dgrad(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin,
                              int ksize, int mma_kind, int ci_begin, int ci_count, void *stream);
/* Sub-pixel data gradient of a stride-2 conv (TF32 path): dX [N, H, W, ci_count] from dZ [N, H/2, W/2, Cout] as four
 * stride-1 convolutions of dZ, one per input-pixel parity, with the 3x3 / 3x2 / 2x3 / 2x2 (5x5) sub-filters that
 * ramnet_pack_weights_dgrad_s2 lays out back to back -- no zero insertion, a quarter of its MACs. */
int ramnet_pack_weights_dgrad_s2(ramnet_handle *h, const float *w_oihw, float *w_packed, int Cout, int Cin,
                                 int ksize, int ci_begin, int ci_count, void *stream);
int ramnet_conv_dgrad_s2(ramnet_handle *h, const float *dz, const float *w_packed_s2, float *dx, int N, int H,
                         int W, int Cout, int ci_count, int ksize, int flags, void *stream);
/* y[n, 2h, 2w, :] = x[n, h, w, :] (+ skip), zeros elsewhere: input of a stride-2 conv's data gradient and of the
 * TransposedConvLayer decoder (submodules.py:38-66; skip = statenet.py:306-308's skip sum). */
int ramnet_zero_insert2x(ramnet_handle *h, const float *x, const float *skip, float *y,
                         float *sum_out /* nullable: dense x + skip [N,H,W,C] */, int N, int H, int W, int C, int Hout,
                         int Wout, void *stream);
/* dz = dy * (y > 0).  flags & RAMNET_FLAG_ROUND_TF32 (here and in the GRU adjoints): round the dz outputs to TF32
 * so that the tcgen05 dgrad / wgrad GEMMs that consume them truncate nothing. */
/* db (nullable, [C]) here and in the gate adjoints: the bias gradient db[c] += sum_pixels dz[., c] fused into the same
 * pass (accumulating, so it can point straight at bias.grad across the passes of BPTT).
 * scratch (nullable, ramnet_colsum_scratch_bytes bytes, zeroed ONCE by the caller and then reused on one stream): the
 * column sums are combined by the last block to finish instead of by atomics (cheaper, and deterministic). */
size_t ramnet_colsum_scratch_bytes(void);
int ramnet_relu_bwd(ramnet_handle *h, const float *dy, const float *y, float *dz, int64_t n, int C, float *db,
                    float *scratch, int flags, void *stream);
/* ConvGRU adjoints (submodules.py:446-452).  gru_out_bwd: dzo = dh'*u*(1-o^2); columns [C,2C) of dzru =
 * dh'*(o-h)*u*(1-u); dh = dh'*(1-u).  gru_ru_bwd: columns [0,C) of dzru = drh*h*r*(1-r); dh += drh*r. */
int ramnet_gru_out_bwd(ramnet_handle *h, const float *dhn, const float *hprev, const float *u, const float *o,
                       float *dzo, float *dzru, float *dh, float *db_o /* [C] */, float *db_ru /* [2C], rows [C,2C) */,
                       float *scratch, int64_t M, int C, int flags, void *stream);
int ramnet_gru_ru_bwd(ramnet_handle *h, const float *drh, const float *hprev, const float *r, float *dzru,
                      float *dh, float *db_ru /* [2C], rows [0,C) */, float *scratch, int64_t M, int C, int flags,
                      void *stream);
/* ConvLSTM adjoint (submodules.py:341-356): gates = post-activation (i,f,o,g) [M,C,4] stashed by RAMNET_EPI_LSTM (y2);
 * dz [M,4C] comes out in nn.Conv2d row order (gate-major); dh / dc may be NULL (no gradient through that output). */
int ramnet_lstm_bwd(ramnet_handle *h, const float *dh, const float *dc, const float *gates, const float *c_prev,
                    const float *c_new, float *dz, float *dc_prev, float *db /* [4C], nn.Conv2d row order */, int64_t M,
                    int C, int flags, void *stream);
/* pred + sigmoid adjoint: dx[m,c] = g*w[c] (= dskip), dw[c] += sum g*(x+skip)[m,c], db += sum g, g = ddepth*s(1-s);
 * skip may be NULL; depth NULL: ddepth is the gradient of the LOGITS (g = ddepth; a norm layer follows the pred conv).
 * w_skip (nullable; concat form, see ramnet_pred_sigmoid): dskip[m,c] = g*w_skip[c] goes to its own buffer, dw is [2C]
 * with dw[C+c] += sum g*skip[m,c] and dw[c] += sum g*x[m,c]. */
int ramnet_pred_bwd(ramnet_handle *h, const float *ddepth, const float *depth, const float *x, const float *skip,
                    const float *w, const float *w_skip, float *dx, float *dskip, float *dw, float *db, int64_t M, int C,
                    void *stream);
/* adjoint of ramnet_upsample2x_add: dx (= dskip) [N,H,W,C] from dy [N,2H,2W,C] */
int ramnet_upsample2x_bwd(ramnet_handle *h, const float *dy, float *dx, int N, int H, int W, int C,
                          void *stream);

/* ---- live normalisation layers (config key norm = 'BN' | 'IN') ------------------------------------------- *
 * Replaces nn.BatchNorm2d / nn.InstanceNorm2d where their statistics cannot be folded into the conv weights
 * (model/submodules.py:21-24,29-30 ConvLayer, :52-55,60-61 TransposedConvLayer, :82-85,91-92 UpsampleConvLayer,
 * :188-194,203-211 ResidualBlock): train mode (batch / instance statistics, running-statistics update with momentum
 * and the unbiased variance), the ResidualBlock's InstanceNorm2d without running statistics (also in eval mode), and
 * eval-mode norms that gradients flow through (RAMNET_NORM_RUNNING).  z = conv output [N, HW, C] NHWC;
 * y = act((z - mean) * invstd * gamma + beta (+ res)).  gamma / beta / res / running_* may be NULL.
 * sums: ramnet_norm_scratch_bytes(N, C) bytes of scratch; stats (out, [G, C, 2] floats = mean, invstd with G = N for
 * RAMNET_NORM_INSTANCE else 1) is what ramnet_norm_bwd needs back. */
#define RAMNET_NORM_INSTANCE 1   /* statistics per (sample, channel) instead of per channel over the batch */
#define RAMNET_NORM_RUNNING 2    /* normalise with running_mean / running_var; nothing is updated */
#define RAMNET_NORM_RELU 4
#define RAMNET_NORM_SIGMOID 8
#define RAMNET_NORM_ROUND_TF32 16 /* round the stored output (forward: y, backward: dz) to TF32 */
size_t ramnet_norm_scratch_bytes(int N, int C);
int ramnet_norm_fwd(ramnet_handle *h, const float *z, const float *res, const float *gamma, const float *beta,
                    float *running_mean, float *running_var, double momentum, double eps, int N, int64_t HW, int C,
                    int flags, double *sums, float *stats, float *y, void *stream);
/* dz = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * act'(y), xhat = (z - mean) * invstd (the two
 * means are zero under RAMNET_NORM_RUNNING); dres (nullable) = g; dgamma / dbeta (nullable, [C]) accumulate (+=).
 * coef: [G, C, 2] floats of scratch. */
int ramnet_norm_bwd(ramnet_handle *h, const float *dy, const float *y, const float *z, const float *stats,
                    const float *gamma, int N, int64_t HW, int C, int flags, double *sums, float *coef, float *dz,
                    float *dres, float *dgamma, float *dbeta, void *stream);

/* ---- a-12  scale_invariant_loss ---------------------------------------- *
 * Replaces model/loss.py:6-9 (boolean-mask gathers + host sync).
 * stats[0..2] = sum(d), sum(d^2), count over non-NaN d = pred - target, in
 * float64 (zeroed by the callee).  ramnet_si_loss_grad then writes
 * grad[i] = scale * (2w/n)(d_i - lambda*mean(d)), 0 where d is NaN, reading n
 * and mean from `stats` on the device (no host sync).  For data-parallel
 * exact-loss mode all-reduce `stats` between the two calls. */
#define RAMNET_LOSS_LOG_SPACE 1 /* d = log(pred) - log(target): scale_invariant_log_loss, model/loss.py:12-15 */
int ramnet_si_loss_stats(ramnet_handle *h, const float *pred, const float *target, int64_t n,
                         double *stats, int flags, void *stream);
int ramnet_si_loss_value(ramnet_handle *h, const double *stats, float weight, float n_lambda,
                         float *loss_out, void *stream);
/* scale_dev (nullable): device scalar multiplied into `scale` (autograd's grad_output, never read on the host). */
int ramnet_si_loss_grad(ramnet_handle *h, const float *pred, const float *target, int64_t n,
                        const double *stats, float weight, float n_lambda, float scale,
                        const float *scale_dev, int flags, float *grad, void *stream);

/* ---- §8f-1  MultiScaleGradient loss -------------------------------------- *
 * Replaces model/loss.py:22-70 (4 x AvgPool2d + kornia spatial_gradient + boolean-mask sums) for C = 1 maps
 * [N,1,H,W]: stats[2*s] = sum |g|, stats[2*s+1] = count of non-NaN gradient entries at scale s (float64, zeroed by
 * the callee); value = (1/S) sum_s stats[2s]/stats[2s+1] * N * 2; grad = d value / d pred * scale (0 at NaN).
 * ramnet_msg_pooled_count = P, the pooled pixels over all scales.  workspace: ramnet_msg_workspace_bytes bytes of 16-byte
 * aligned scratch (pooled maps + replicated statistics in _stats, pooled-pixel gradients in _grad).  signs (nullable in _stats; P int8 pairs): sign of the two Sobel components per
 * pooled pixel, (0, 0) where NaN -- all the backward pass needs from the forward pass. */
int64_t ramnet_msg_pooled_count(int N, int H, int W, int start_scale, int scales);
size_t ramnet_msg_workspace_bytes(int N, int H, int W, int start_scale, int scales);
int ramnet_msg_loss_stats(ramnet_handle *h, const float *pred, const float *target, int N, int H, int W,
                          int start_scale, int scales, double *stats, float *workspace, signed char *signs, void *stream);
int ramnet_msg_loss_value(ramnet_handle *h, const double *stats, int N, int scales, float *loss_out, void *stream);
/* n_batch: batch size the value was normalised with (0 = N; the global batch when `stats` were all-reduced). */
int ramnet_msg_loss_grad(ramnet_handle *h, const signed char *signs, int N, int H, int W, int start_scale, int scales,
                         const double *stats, int n_batch, float scale, const float *scale_dev, float *workspace,
                         float *grad, void *stream);
/* preview=True branch (model/loss.py:46-47, TensorBoard only): out [N,1,H/pool,W/pool] = kornia sobel magnitude
 * sqrt(gx^2 + gy^2 + 1e-6) of AvgPool2d(pool)(pred - target). */
int ramnet_msg_sobel_preview(ramnet_handle *h, const float *pred, const float *target, int N, int H, int W,
                             int pool, float *out, void *stream);

/* ---- a-14  Adam --------------------------------------------------------- *
 * Replaces torch.optim.Adam.step (built at base/base_trainer.py:36-37,
 * stepped at trainer/lstm_trainer.py:453) with one multi-tensor launch over a
 * flat fp32 buffer.  step is 1-based; L2 weight decay as torch.optim.Adam.  Hyper-parameters
 * are doubles: torch derives 1-beta, the bias corrections and the step size in double. */
int ramnet_adam_step(ramnet_handle *h, float *p, const float *g, float *m, float *v, int64_t n,
                     double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                     void *stream);
/* Same update with the 1-based step number kept in device memory: *step_counter is incremented (when `increment`),
 * then used.  Lets a whole training step (forward, backward, optimiser) be captured in one CUDA graph and replayed.
 * A step split into several launches over slices of the flat buffer (bucketed gradient all-reduce) passes
 * increment = 1 for its first slice only. */
int ramnet_adam_step_dev(ramnet_handle *h, float *p, const float *g, float *m, float *v, int64_t n, double lr,
                         double beta1, double beta2, double eps, double weight_decay, int *step_counter,
                         int increment, void *stream);

/* ---- §8f-4  test.py output stage ------------------------------------------- *
 * Replaces the per-map host work of RAM_Net/test.py:259-360,365-379 for N maps [N, 1, H, W] (hw = H*W): grey (nullable,
 * [N, hw] uint8) = cv2.imwrite's 8-bit conversion of depth * 255; bgr (nullable, [N, hw, 3] uint8) = make_colormap
 * (test.py:31-38) through the caller's 256 x 3 RGB LUT (the reference's matplotlib color mapper sampled at i/255);
 * scale_sums (nullable, [N, 2] doubles, needs target) = (sum p t, sum p p) in metric space, p = clip * exp(reg (d - 1))
 * (test.py:370-376).  scratch: N * 4 uint32. */
int ramnet_depth_output(ramnet_handle *h, const float *depth, const float *target, int N, int64_t hw,
                        const float *lut_rgb256, unsigned char *grey, unsigned char *bgr, double *scale_sums,
                        float reg_factor, float clip_distance, unsigned *scratch, void *stream);

/* ---- measurement utility (not on the path) ---------------------------------- *
 * Live ceiling of the tensor pipe the convolutions use: tcgen05.mma kind::tf32 128x256x8 issued back to back from
 * shared memory, one CTA per SM, timed with CUDA events on the legacy stream (synchronises).  bench.py reports the
 * roofline fraction against this number next to the bf16 figure of MEASURED_PEAKS.json. */
int ramnet_tf32_pipe_rate(ramnet_handle *h, double *tflops_out, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* RAMNET_B200_H */

// placeholder: replaced by the tcgen05/TMA implicit-GEMM kernel
#include "common.cuh"
size_t conv_tf32_workspace_bytes(const ramnet_conv_desc *) { return 0; }
int conv_fwd_tf32(ramnet_handle *, const ramnet_conv_desc *, const float *, const float *, const float *,
                  const EpiParams &, void *, size_t, cudaStream_t) {
    return ramnet_set_error(RAMNET_EUNSUPPORTED, "conv_fwd: RAMNET_MMA_TF32 path not built");
}

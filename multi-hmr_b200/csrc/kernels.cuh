// Host-callable launchers of the non-GEMM kernels (one translation unit each).
#pragma once
#include "common.cuh"

namespace mhmr {

// attn_tc.cu
int attention_forward(const __half* qkv, int64_t ld_qkv, __half* out, int64_t ldo, int B, int T, int D,
                      cudaStream_t stream);

}  // namespace mhmr

// Stage-level C-ABI entry points (unit parity + ncu targets). See include/mhmr.h.
#include "../../include/mhmr.h"
#include "gemm_tc.cuh"
#include "kernels.cuh"

using namespace mhmr;

extern "C" {

const char* mhmr_last_error(void) { return get_last_error(); }

int mhmr_op_gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                     int epilogue, const float* bias, const float* gamma, const float* rowadd,
                     void* out, int64_t ldo, int rows_in, int rows_out, int row_off, int block_n,
                     void* stream) {
  GemmEpi ep;
  ep.bias = bias;
  ep.gamma = gamma;
  ep.rowadd = rowadd;
  ep.out = out;
  ep.ldo = ldo;
  ep.rows_in = rows_in;
  ep.rows_out = rows_out;
  ep.row_off = row_off;
  GemmPlan plan;
  int rc = gemm_plan_init(&plan, static_cast<const __half*>(A), lda, static_cast<const __half*>(W),
                          ldw, M, N, K, epilogue, ep, block_n);
  if (rc != MHMR_OK) return rc;
  return gemm_plan_run(&plan, static_cast<cudaStream_t>(stream));
}

int mhmr_op_attention(const void* qkv, int64_t ld_qkv, void* out, int64_t ldo, int B, int T, int D,
                      void* stream) {
  return attention_forward(static_cast<const __half*>(qkv), ld_qkv, static_cast<__half*>(out), ldo, B, T,
                           D, static_cast<cudaStream_t>(stream));
}

int mhmr_op_normalize_u8(const void* img_u8, const float* lut, float* out, int B, int H, int W, void* stream) {
  MHMR_REQUIRE(img_u8 != nullptr && lut != nullptr && out != nullptr, "null argument");
  return normalize_u8(static_cast<const uint8_t*>(img_u8), lut, out, B, H, W, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

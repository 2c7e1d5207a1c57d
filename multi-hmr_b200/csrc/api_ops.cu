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
  MHMR_REQUIRE(epilogue >= 0 && epilogue < EPI_NUM_PUBLIC_KINDS, "gemm: bad epilogue kind");
  GemmPlan plan;
  int rc = gemm_plan_init(&plan, static_cast<const __half*>(A), lda, static_cast<const __half*>(W),
                          ldw, M, N, K, epilogue, ep, block_n);
  if (rc != MHMR_OK) return rc;
  return gemm_plan_run(&plan, static_cast<cudaStream_t>(stream));
}

int mhmr_op_resid_ln_linear_f16(const void* A, int64_t lda, const void* Wp, int64_t ldwp, const float* bp,
                                const float* ls, float* X, int M, int D, int Ka, const float* ln_g,
                                const float* ln_b, const float* W, const float* b, int N, int gelu, void* out16,
                                int64_t ldo, void* stream) {
  MHMR_REQUIRE(A && Wp && bp && ls && X && ln_g && ln_b && W && b && out16, "null argument");
  MHMR_REQUIRE(D % 128 == 0 && D <= 1024 && N % 32 == 0, "resid_ln_linear: D must be a multiple of 128 (<= 1024)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int bn_d = (D % 256 == 0) ? 512 : 128, bn_n = (N % 256 == 0) ? 512 : 128;
  const int slots = gemm_stat_slots(D, bn_d);
  __half *x16 = nullptr, *xlo = nullptr, *W16 = nullptr;
  float2* stats = nullptr;
  float* bias2 = nullptr;
  auto release = [&]() {
    cudaStreamSynchronize(st);
    cudaFree(x16); cudaFree(xlo); cudaFree(W16); cudaFree(stats); cudaFree(bias2);
  };
  int rc = MHMR_OK;
  auto run = [&]() -> int {
    MHMR_CUDA_CHECK(cudaMalloc(&x16, static_cast<size_t>(M) * D * 2));
    MHMR_CUDA_CHECK(cudaMalloc(&xlo, static_cast<size_t>(M) * D * 2));
    MHMR_CUDA_CHECK(cudaMalloc(&W16, static_cast<size_t>(N) * D * 2));
    MHMR_CUDA_CHECK(cudaMalloc(&stats, static_cast<size_t>(M) * slots * sizeof(float2)));
    MHMR_CUDA_CHECK(cudaMalloc(&bias2, static_cast<size_t>(N) * 4));
    int r = fold_ln_linear(W, b, ln_g, ln_b, W16, bias2, N, D, st);
    if (r != MHMR_OK) return r;
    r = split_rowstats(X, x16, xlo, D, stats, slots, M, D, st);  // the stream enters as (hi, lo)
    if (r != MHMR_OK) return r;
    GemmEpi p;
    p.bias = bp; p.gamma = ls;
    p.x16 = x16; p.xlo = xlo; p.ldx16 = D; p.stats = stats; p.stat_slots = slots;
    GemmPlan pp;
    r = gemm_plan_init(&pp, static_cast<const __half*>(A), lda, static_cast<const __half*>(Wp), ldwp, M, D, Ka,
                       EPI_LS_RESID_SPLIT, p, bn_d);
    if (r != MHMR_OK) return r;
    r = gemm_plan_run(&pp, st);
    if (r != MHMR_OK) return r;
    GemmEpi c;
    c.bias = bias2; c.stats = stats; c.stat_slots = slots; c.out = out16; c.ldo = ldo;
    GemmPlan cp;
    r = gemm_plan_init(&cp, x16, D, W16, D, M, N, D, gelu ? EPI_LN_GELU_F16 : EPI_LN_BIAS_F16, c, bn_n);
    if (r != MHMR_OK) return r;
    r = gemm_plan_run(&cp, st);
    if (r != MHMR_OK) return r;
    return merge_split(x16, xlo, X, static_cast<int64_t>(M) * D, st);
  };
  rc = run();
  release();
  return rc;
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

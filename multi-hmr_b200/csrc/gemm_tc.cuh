// tcgen05 / TMA / TMEM GEMM family for the Multi-HMR hot path (sm_100a only).
//
//   C[M,N] = epilogue( A[M,K] (fp16, K-major) x W[N,K]^T (fp16, K-major, i.e. torch Linear.weight) )
//
// fp32 accumulation in TMEM.  The epilogues cover every Linear of the reference's ViT backbone and
// of the dense part of the head (SURVEY.md §2.4 k1,k3,k5,k6,k7,k9,k15):
//   EPI_BIAS_F16       out16 = acc + bias                       (qkv: dinov2 Attention.qkv)
//   EPI_BIAS_GELU_F16  out16 = gelu_erf(acc + bias)             (mlp.fc1 + nn.GELU)
//   EPI_BIAS_RELU_F16  out16 = relu(acc + bias)                 (regression_mlp hidden, model.py:596-609)
//   EPI_LS_RESID_F32   out32 += gamma * (acc + bias)  in place  (attn.proj / mlp.fc2 + LayerScale + residual)
//   EPI_ROWADD_F32     out32[remap(m)] = acc + rowadd[m % rows_in]  (patch-embed + bias + pos-embed scatter)
//   EPI_BIAS_F32       out32 = acc + bias (bias may be null)    (HPH to_kv, cross_attn_transformer.py:187)
#pragma once
#include "common.cuh"

namespace mhmr {

enum GemmEpiKind : int {
  EPI_BIAS_F16 = 0,
  EPI_BIAS_GELU_F16 = 1,
  EPI_BIAS_RELU_F16 = 2,
  EPI_LS_RESID_F32 = 3,
  EPI_ROWADD_F32 = 4,
  EPI_BIAS_F32 = 5,
  EPI_NUM_KINDS = 6,
};

struct GemmEpi {
  const float* bias = nullptr;    // [N]
  const float* gamma = nullptr;   // [N]            (EPI_LS_RESID_F32)
  const float* rowadd = nullptr;  // [rows_in, N]   (EPI_ROWADD_F32)
  void* out = nullptr;            // fp16 or fp32, row pitch ldo elements
  int64_t ldo = 0;
  // Row remap: out_row = (m / rows_in) * rows_out + row_off + (m % rows_in); rows_in == 0 => identity.
  int rows_in = 0, rows_out = 0, row_off = 0;
};

struct GemmPlan {
  CUtensorMap tmA, tmB;
  int M = 0, N = 0, K = 0;
  int bn = 256;          // 128 / 256: single-CTA tile width; 512: CTA pair (cta_group::2), 256 x 256 tile
  int epi = EPI_BIAS_F16;
  GemmEpi ep;
  int grid = 0;
};

// Build the TMA descriptors + launch geometry once (weights and workspaces are persistent).
int gemm_plan_init(GemmPlan* plan, const __half* A, int64_t lda, const __half* W, int64_t ldw, int M,
                   int N, int K, int epi_kind, const GemmEpi& ep, int bn);
int gemm_plan_run(const GemmPlan* plan, cudaStream_t stream);
int gemm_plan_run_2cta(const GemmPlan* plan, cudaStream_t stream);  // gemm_tc2.cu
// launch geometry for `M` rows with this plan's tile shape
int gemm_plan_grid(const GemmPlan* plan, int M);

}  // namespace mhmr

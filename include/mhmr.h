/*
 * libmhmr_sm100.so — C-ABI of the B200-native Multi-HMR inference path.
 *
 * Drop-in boundary for the ONE hot path of naver/multi-hmr: `Model.forward(x, K)` as called by
 * `demo.py:forward_model` (reference demo.py:108-126, model.py:205-349).  The reference has no native
 * layer (pure PyTorch); what this library replaces are the ATen/cuBLAS/cuDNN dispatches listed in
 * SURVEY.md §2.4 (k1..k19).  Signatures use plain pointers and sizes only (no torch types): device
 * buffers are borrowed for the duration of a call, `stream` is a `cudaStream_t` passed as void*.
 *
 * Every function returns 0 on success and a negative code on failure; `mhmr_last_error()` returns a
 * thread-local message.  Nothing here falls back to a CPU path: without a CUDA device the compute
 * entry points fail loudly.
 */
#ifndef MHMR_H_
#define MHMR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MHMR_OK 0
#define MHMR_ERR_CUDA (-1)
#define MHMR_ERR_ARG (-2)
#define MHMR_ERR_STATE (-3)
#define MHMR_ERR_CAPACITY (-4)
#define MHMR_ERR_UNSUPPORTED (-5)

const char* mhmr_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Stage-level operators (unit parity + ncu targets)
 * ---------------------------------------------------------------------------------------------- */

/* Epilogue kinds of mhmr_op_gemm_f16 */
#define MHMR_EPI_BIAS_F16 0      /* out16 = acc + bias                — dinov2 Attention.qkv              */
#define MHMR_EPI_BIAS_GELU_F16 1 /* out16 = gelu_erf(acc + bias)      — dinov2 Mlp.fc1 + nn.GELU          */
#define MHMR_EPI_BIAS_RELU_F16 2 /* out16 = relu(acc + bias)          — regression_mlp, model.py:596-609  */
#define MHMR_EPI_LS_RESID_F32 3  /* out32 += gamma*(acc + bias)       — attn.proj / mlp.fc2 + LayerScale  */
#define MHMR_EPI_ROWADD_F32 4    /* out32[remap(m)] = acc + rowadd[m % rows_in] — patch-embed + pos-embed */
#define MHMR_EPI_BIAS_F32 5      /* out32 = acc (+ bias)              — HPH to_kv, cross_attn_transformer.py:187 */

/* C = epilogue(A[M,K] x W[N,K]^T): fp16 operands (K contiguous, torch nn.Linear weight layout), fp32
 * accumulation on the tcgen05 tensor cores.  Replaces torch.nn.functional.linear on the hot path
 * (reference blocks/dinov2.py:25 -> dinov2 Attention/Mlp; model.py:135; cross_attn_transformer.py:187).
 * Row remap for MHMR_EPI_ROWADD_F32: out_row = (m / rows_in) * rows_out + row_off + m % rows_in. */
int mhmr_op_gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                     int epilogue, const float* bias, const float* gamma, const float* rowadd,
                     void* out, int64_t ldo, int rows_in, int rows_out, int row_off, int block_n,
                     void* stream);

/* Multi-head self-attention of the ViT backbone, head dim 64: out[:, h*64:(h+1)*64] =
 * softmax(q_h k_h^T / 8) v_h per image.  qkv is [B*T, 3*D] fp16 (q | k | v column blocks, the layout the
 * qkv Linear produces), out is [B*T, D] fp16.  Replaces dinov2 Attention.forward's
 * `q*scale @ k^T -> softmax -> @ v` (reached from reference blocks/dinov2.py:25). */
int mhmr_op_attention(const void* qkv, int64_t ld_qkv, void* out, int64_t ldo, int B, int T, int D,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MHMR_H_ */

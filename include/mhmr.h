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
 * Row remap for MHMR_EPI_ROWADD_F32: out_row = (m / rows_in) * rows_out + row_off + m % rows_in.
 * block_n: 128 or 256 = single-CTA 128 x block_n tiles; 512 = CTA pairs (tcgen05 cta_group::2), 256 x 256 tiles. */
int mhmr_op_gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                     int epilogue, const float* bias, const float* gamma, const float* rowadd,
                     void* out, int64_t ldo, int rows_in, int rows_out, int row_off, int block_n,
                     void* stream);

/* One seam of a dinov2 Block with the LayerNorm folded into the GEMMs on both sides of it -- what the engine runs
 * between attention and the MLP (and between one block's MLP and the next block's attention):
 *     X += ls * (A @ Wp^T + bp)                       attn.proj / mlp.fc2 + LayerScale + residual (layers/block.py)
 *     out16 = act(LayerNorm(X; ln_g, ln_b, eps 1e-6) @ W^T + b)      norm2 -> mlp.fc1 (+ GELU) / norm1 -> attn.qkv
 * Between the two the residual stream is a two-term fp16 split X = hi + lo (22 significant bits, 4 bytes per element
 * like fp32): the first GEMM's epilogue updates (hi, lo) in place and leaves per-row partial (sum, sum of squares); the
 * second GEMM runs on the hi plane with the row-centred W * diag(ln_g) and applies 1/sigma in its epilogue.  A, Wp fp16
 * (K contiguous); X fp32 [M, D] in place (split on entry, merged on exit); W fp32 [N, D] (folded and rounded inside);
 * D multiple of 128 (<= 1024), N multiple of 32.
 * Unit-test entry: allocates its temporaries and synchronises the stream. */
int mhmr_op_resid_ln_linear_f16(const void* A, int64_t lda, const void* Wp, int64_t ldwp, const float* bp,
                                const float* ls, float* X, int M, int D, int Ka, const float* ln_g,
                                const float* ln_b, const float* W, const float* b, int N, int gelu, void* out16,
                                int64_t ldo, void* stream);

/* Multi-head self-attention of the ViT backbone, head dim 64: out[:, h*64:(h+1)*64] =
 * softmax(q_h k_h^T / 8) v_h per image.  qkv is [B*T, 3*D] fp16 (q | k | v column blocks, the layout the
 * qkv Linear produces), out is [B*T, D] fp16.  Replaces dinov2 Attention.forward's
 * `q*scale @ k^T -> softmax -> @ v` (reached from reference blocks/dinov2.py:25). */
int mhmr_op_attention(const void* qkv, int64_t ld_qkv, void* out, int64_t ldo, int B, int T, int D,
                      void* stream);

/* Image preprocessing on the device: uint8 [B,H,W,3] (RGB, HWC) -> fp32 [B,3,H,W] through a [3][256] fp32 table.
 * Replaces reference utils/image.py:12-24 `normalize_rgb` (called from demo.py:48 `open_image`); with the table the
 * host derives from that function the output is bit-identical, and the upload is 4x smaller.  W % 4 == 0. */
int mhmr_op_normalize_u8(const void* img_u8, const float* lut, float* out, int B, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Engine: the whole `Model.forward(x, K)` path (reference model.py:205-349) behind one handle
 * ---------------------------------------------------------------------------------------------- */
typedef struct mhmr_engine mhmr_engine;

typedef struct mhmr_config {
  int arch;              /* 0 = dinov2_vits14, 1 = dinov2_vitb14, 2 = dinov2_vitl14 (Model(backbone=...), model.py:35) */
  int img_size;          /* Model(img_size=...), multiple of 14 (model.py:37,65) */
  int max_batch;         /* capacity of the activation workspaces */
  int max_persons;       /* capacity of the per-person buffers; more detections => MHMR_ERR_CAPACITY */
  int xat_depth;         /* Model(xat_depth=...)     model.py:42 */
  int xat_num_heads;     /* Model(xat_num_heads=...) model.py:43 */
  int num_betas;         /* Model(num_betas=...)     model.py:47 (10 or 11) */
  int person_center_idx; /* index of Model(person_center=...) in smplx JOINT_NAMES ('head' = 15) */
  int num_verts;         /* body-model vertices (SMPL-X: 10475) */
  int refine_central;    /* 1: fp32 refinement of the detected tokens' residual streams; 0: bulk fp16 pass only */
} mhmr_config;

/* Per-call outputs: device buffers owned by the caller (torch tensors), sized for max_persons.
 * Person order = torch.where order (b, y, x) (model.py:149,616).  Nullable: v2d, z. */
typedef struct mhmr_outputs {
  float* scores_map;   /* [B, res, res]    'scores' (after NMS in inference mode, model.py:145-157)      */
  int32_t* count;      /* [1]              number of detected persons P                                   */
  int32_t* det_idx;    /* [3, max_persons] image index b, row y, col x of each person                     */
  float* det_score;    /* [max_persons]    person 'scores'                                                */
  float* offset;       /* [max_persons, 2] mlp_offset output (model.py:258)                               */
  float* loc;          /* [max_persons, 2] 'loc' (model.py:272-275)                                       */
  float* dist_pp;      /* [max_persons]    'dist_postprocessed' (raw pred_cam[:,0])                       */
  float* dist;         /* [max_persons]    'dist' (model.py:189-203)                                      */
  float* rotmat;       /* [max_persons, 53, 3, 3]                                                         */
  float* rotvec;       /* [max_persons, 53, 3]                                                            */
  float* shape;        /* [max_persons, num_betas]                                                        */
  float* expression;   /* [max_persons, 10]                                                               */
  float* transl;       /* [max_persons, 3]                                                                */
  float* transl_pelvis;/* [max_persons, 3]                                                                */
  float* v3d;          /* [max_persons, V, 3]                                                             */
  float* v2d;          /* [max_persons, V, 2]   nullable (not part of the inference person dict)          */
  float* j3d;          /* [max_persons, 127, 3]                                                           */
  float* j2d;          /* [max_persons, 127, 2]                                                           */
  float* z;            /* [B, N, D] backbone features (blocks/dinov2.py:25), nullable (stage parity)      */
} mhmr_outputs;

/* Replaces `Model(**ckpt_args)` (reference demo.py:98-100, model.py:33-131). */
int mhmr_create(const mhmr_config* cfg, mhmr_engine** out);
int mhmr_destroy(mhmr_engine* h);

/* Replaces `model.load_state_dict(ckpt['model_state_dict'], strict=False)` (demo.py:103): one call per
 * fp32 tensor, `key` = the reference's state_dict key.  `data` may be a host or a
 * device pointer; the engine keeps its own device copy.  Extra keys the C-ABI expects:
 *   backbone.encoder.pos_embed  must already be interpolated to the working grid: [1, 1+N, D]
 *   camera.freq_bands           [16] = torch.linspace(1, 32, 16)   (blocks/camera_embed.py:46)
 *   smplx.{v_template,shapedirs,expr_dirs,posedirs,J_regressor,lbs_weights,lmk_bary_coords}
 *                               the buffers smplx.create() registers (blocks/smpl_layer.py:38). */
int mhmr_set_weight(mhmr_engine* h, const char* key, const float* data, int64_t numel);
/* Integer body-model tables: smplx.parents [55], smplx.extra_joints_idxs [21],
 * smplx.lmk_tri [51*3] (= faces[lmk_faces_idx]). */
int mhmr_set_table_i32(mhmr_engine* h, const char* key, const int32_t* data, int64_t numel);
/* Repack (fp16 K-major weight tiles, fused tables, folded joint regressor), allocate workspaces, build
 * TMA descriptors.  Fails if a required key is missing. */
int mhmr_finalize(mhmr_engine* h);

/* One forward of `Model.forward(x, idx, det_thresh, nms_kernel_size, K, is_training)` (model.py:205-349):
 * x [B,3,S,S] fp32 NCHW and K [B,3,3] fp32 are DEVICE pointers.  forced_idx (nullable) is a device
 * int64 [4, forced_P] tensor (b, y, x, c) = the reference's `idx=` argument (training-style path,
 * model.py:150-151: no NMS/threshold).  Asynchronous on `stream`; no host synchronisation. */
int mhmr_forward(mhmr_engine* h, const float* x, const float* K, int B, float det_thresh,
                 int nms_kernel_size, const int64_t* forced_idx, int forced_P, const mhmr_outputs* out,
                 void* stream);
/* The same forward fed by the fused image loader (replaces `normalize_rgb` of
 * utils/image.py:12-24 + the fp32 upload of demo.py:48-50): img_u8 [B,S,S,3] uint8 RGB (HWC, what PIL yields after
 * ImageOps.pad) and the [3][256] fp32 table of normalize_rgb; uint8 -> normalised fp16 patch rows in one kernel. */
int mhmr_forward_u8(mhmr_engine* h, const uint8_t* img_u8, const float* lut, const float* K, int B, float det_thresh,
                    int nms_kernel_size, const int64_t* forced_idx, int forced_P, const mhmr_outputs* out,
                    void* stream);
/* Waits for the forward enqueued last on `stream` and returns the person count; MHMR_ERR_CAPACITY if it
 * exceeded max_persons (outputs are then incomplete — never silently truncated). */
int mhmr_sync_count(mhmr_engine* h, void* stream, int* num_persons);

/* Stage-level entry: backbone only (blocks/dinov2.py:16-26): x [B,3,S,S] -> z [B,N,D] fp32. */
int mhmr_vit_forward(mhmr_engine* h, const float* x, int B, float* z, void* stream);
/* Stage-level entry: SMPL-X layer only (blocks/smpl_layer.py:47-155) for P persons (device pointers):
 * rotvec [P,53,3], shape [P,nb], expression [P,10], loc [P,2], dist [P], K_det [P,3,3]. */
int mhmr_smplx_forward(mhmr_engine* h, int P, const float* rotvec, const float* shape,
                       const float* expression, const float* loc, const float* dist, const float* K_det,
                       float* v3d, float* v2d, float* j3d, float* j2d, float* transl, float* transl_pelvis,
                       void* stream);
/* ------------------------------------------------------------------------------------------------
 * Sharded batches (one process per GPU, image shards per rank).  The reference is single-GPU
 * (README.md:107); these entry points are what a multi-GPU caller of `forward_model` binds.
 *   block  = header (8 x int32: persons detected, persons packed, capacity, floats per record, first global
 *            image index of the rank, 0, 0, 0) | capacity x record (fp32)
 *   record = [global image index, score, loc 2, transl 3, transl_pelvis 3, rotvec 159, expression 10,
 *             shape nb, v3d 3V, j3d 381, j2d 254]  — the person dict of model.py:329-347
 * ---------------------------------------------------------------------------------------------- */
int mhmr_record_floats(int num_betas, int num_verts);
int64_t mhmr_record_block_bytes(int num_betas, int num_verts, int capacity);
/* ONE kernel: packs the (device-resident) outputs of a forward into `block`; the person count is read on the
 * device and travels in the header, unused record slots are zero-filled. */
int mhmr_pack_records(const mhmr_outputs* out, int max_persons, int num_betas, int num_verts, int image_offset,
                      int capacity, float* block, void* stream);
/* NCCL communicator of the record exchange (libnccl.so.2 is resolved at run time from the process). */
typedef struct mhmr_comm mhmr_comm;
int mhmr_nccl_unique_id(void* id128);                                   /* ncclGetUniqueId (rank 0) */
int mhmr_comm_create(const void* id128, int world, int rank, mhmr_comm** out);  /* ncclCommInitRank  */
int mhmr_comm_destroy(mhmr_comm* c);
/* ONE ncclAllGather of the per-rank blocks (all_blocks = world x block_bytes, rank order) on `stream`. */
int mhmr_allgather_records(mhmr_comm* c, const void* block, void* all_blocks, int64_t block_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluation metrics: what the reference's `Trainer.evaluate` (train.py:336-482) computes
 * from the outputs of Model.forward.  All pointers are DEVICE pointers; nothing synchronises with the host.
 * ---------------------------------------------------------------------------------------------- */
/* Replaces utils/training.py:25-147 `match_2d_greedy(pred_kps, gtkp, valid_mask)` (valid=None, as train.py:364
 * calls it) including the IoU gate of :149-193: pred_j2d [P,J,2], gt_j2d [G,J,2] pixels, valid_mask [G,J] bytes or
 * NULL (all valid).  Outputs: pairs [min(P,G),2] = (pred, gt) in the order the greedy loop finds them, n_pairs [1],
 * pred_to_gt [P] (-1 = false positive), gt_to_pred [G] (-1 = miss).  P, G <= 48. */
int mhmr_eval_match_2d(const float* pred_j2d, const float* gt_j2d, const uint8_t* valid_mask, int P, int G, int J,
                       float iou_thresh, int32_t* pairs, int32_t* n_pairs, int32_t* pred_to_gt, int32_t* gt_to_pred,
                       void* stream);
/* Replaces train.py:373-394 (PVE / PA-PVE) and :411-427 (MPJPE / PA-MPJPE) for the matched pairs: point sets
 * pred [*,n,3], gt [*,n,3] (metres), optional per-person centres [*,3] subtracted first (the pelvis, train.py:376-382);
 * err_mm[m] = mean |gt - pred| * 1000, pa_err_mm[m] = the same after the similarity alignment of pred onto gt
 * (roma.rigid_points_registration(compute_scaling=True)).  One CTA per pair m < *n_pairs (m < max_pairs). */
int mhmr_eval_points_error(const float* pred, const float* pred_center, const float* gt, const float* gt_center,
                           const int32_t* pairs, const int32_t* n_pairs, int max_pairs, int n_points, float* err_mm,
                           float* pa_err_mm, void* stream);

/* Kernel launches enqueued by the last mhmr_forward (bench.py's `gpu_launches`). */
int mhmr_last_launch_count(mhmr_engine* h);

/* Live per-kernel-family timing with CUDA events on the launching stream (bench.py's `roofline`):
 * while enabled, every launch of mhmr_forward is bracketed by an event pair; mhmr_get_profile waits for
 * them, returns the summed device milliseconds and launch counts per category and resets the record. */
#define MHMR_CAT_MISC 0        /* im2col, cls rows                          */
#define MHMR_CAT_LAYERNORM 1
#define MHMR_CAT_GEMM_QKV 2
#define MHMR_CAT_ATTENTION 3
#define MHMR_CAT_GEMM_PROJ 4
#define MHMR_CAT_GEMM_FC1 5
#define MHMR_CAT_GEMM_FC2 6
#define MHMR_CAT_GEMM_OTHER 7  /* patch-embed, detection hidden, HPH to_kv */
#define MHMR_CAT_HEAD 8        /* detection / HPH / post-processing kernels */
#define MHMR_CAT_SMPLX 9       /* prep + vertex + joints kernels            */
#define MHMR_CAT_REFINE 10     /* fp32 refinement of the detected tokens' residual streams */
#define MHMR_NUM_CATEGORIES 11
int mhmr_set_profiling(mhmr_engine* h, int enable);
int mhmr_get_profile(mhmr_engine* h, float* ms_by_category, int* launches_by_category);

#ifdef __cplusplus
}
#endif
#endif /* MHMR_H_ */

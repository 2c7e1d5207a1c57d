// Host-callable launchers of the non-GEMM kernels (one translation unit each).
#pragma once
#include "common.cuh"

namespace mhmr {

// ---- attn_tc.cu ------------------------------------------------------------------------------
int attention_forward(const __half* qkv, int64_t ld_qkv, __half* out, int64_t ldo, int B, int T, int D,
                      cudaStream_t stream);

// ---- vit_misc.cu -----------------------------------------------------------------------------
int im2col_patch14(const float* x, __half* A, int B, int S, int ldA, cudaStream_t stream);
int normalize_u8(const uint8_t* img, const float* lut, float* out, int B, int H, int W, cudaStream_t stream);
int im2col_u8_patch14(const uint8_t* img, const float* lut, __half* A, int B, int S, int ldA, cudaStream_t stream);
int cls_rows(float* X, const float* cls_pos, int B, int T, int D, cudaStream_t stream);
int layernorm(const float* X, const float* gamma, const float* beta, __half* out16, int64_t ld16,
              float* out32, int64_t ld32, int M, int D, float eps, int rows_in, int skip,
              cudaStream_t stream);
int f32_to_f16_2d(const float* src, int64_t lds, __half* dst, int64_t ldd, int rows, int cols,
                  cudaStream_t stream);
// folded LayerNorm (gemm_tc.cuh): entry kernel of the chain and the load-time weight folding
int split_rowstats(const float* X, __half* xhi, __half* xlo, int64_t ld16, float2* stats, int slots, int M, int D,
                   cudaStream_t stream);
int merge_split(const __half* xhi, const __half* xlo, float* X, int64_t n, cudaStream_t stream);
// LayerNorm of a two-term fp16 stream (X = hi plane, Xlo = lo plane; Xlo == nullptr: X is fp32)
int layernorm_split(const void* X, const __half* Xlo, const float* gamma, const float* beta, __half* out16,
                    int64_t ld16, float* out32, int64_t ld32, int M, int D, float eps, int rows_in, int skip,
                    cudaStream_t stream);
int fold_ln_linear(const float* W, const float* bias, const float* ln_g, const float* ln_b, __half* W16,
                   float* bias2, int N, int K, cudaStream_t stream);
int repack_f32(const float* src, int64_t lds, int scol, float* dst, int64_t ldd, int dcol, int rows,
               int cols, bool zero_fill, cudaStream_t stream);
int add_vec(const float* a, const float* b, float* out, int64_t n, int64_t b_period, cudaStream_t stream);

// ---- head.cu ---------------------------------------------------------------------------------
int rowdot_sigmoid(const __half* hid, int64_t ld, const float* w, const float* b, float* scores, int M,
                   int D, cudaStream_t st);
int nms_compact(const float* scores, float* scores_out, int B, int res, int nms_k, float thresh,
                int max_persons, int* det_b, int* det_y, int* det_x, float* det_score, int* count,
                int* count_clamped, int* img_off, cudaStream_t st);
int forced_detections(const float* scores, float* scores_out, int B, int res, const int64_t* idx4, int P,
               int* det_b, int* det_y, int* det_x, float* det_score, int* count, int* count_clamped,
               int* img_off, cudaStream_t st);
int loc_to_transl(const float* loc, const float* dist, const float* K_det, int P, float* transl,
                  cudaStream_t st);
int invert_K(const float* K, float* Kinv, int B, cudaStream_t st);
int ctx_fourier(const float* Kinv, const float* freqs, __half* ctx, int64_t ld, int B, int res, int D,
                int pad_cols, cudaStream_t st);
int person_gather(const float* z32, const float* xr, const float* norm_g, const float* norm_b, const float* Kinv,
                  const float* freqs, const float* cq_x, const float* cq_y, const float* cv_x, const float* cv_y,
                  const int* det_b, const int* det_y, const int* det_x, const int* count, int max_persons, int res,
                  int D, float* zc, float* query, float* vals, int ldq, cudaStream_t st);
// central-stream refinement (engine.cu:refine_streams): row indices, input patches and pos-embed rows of the
// detected cells
int refine_prepare(const float* img, const uint8_t* img_u8, const float* lut, int S, const float* rowadd, int D, const int* det_b, const int* det_y,
                   const int* det_x, const int* count, int max_persons, int res, int* rowidx, float* patch,
                   int ldp, float* xr, cudaStream_t st);
int kv_add_rows(float* KV, int64_t ldkv, const float* dKV, int ncols, const int* det_b, const int* det_y,
                const int* det_x, const int* count, int max_persons, int res, cudaStream_t st);
int skinny_linear(const float* x, int ldx, const int* count, int max_persons, int K, const float* W, int ldw,
                  const float* bias, int Nout, const float* ln_g, const float* ln_b, float ln_eps, int act,
                  const float* resid, int ldr, float* out, int ldo, cudaStream_t st);
// extras of the skinny linear: fp16 gathered input rows, LayerScale, columns per CTA (0 = pick for occupancy)
struct SkinnyExtra {
  const __half* x16 = nullptr;
  int64_t ldx16 = 0;
  const int* rowidx = nullptr;
  const float* gamma = nullptr;
  int cols = 0;
};
int skinny_linear_ex(const float* x, int ldx, const SkinnyExtra& ex, const int* count, int max_persons, int K,
                     const float* W, int ldw, const float* bias, int Nout, const float* ln_g, const float* ln_b,
                     float ln_eps, int act, const float* resid, int ldr, float* out, int ldo, cudaStream_t st);
// ---- refine.cu: central-stream refinement -----------------------------------------
struct RefineLayer {
  const __half* O16;  // this block's attention output of the bulk pass [B*T, D]
  const float *Wproj, *bproj, *ls1, *ln2_g, *ln2_b, *Wfc1, *bfc1, *Wfc2, *bfc2, *ls2;  // fp32 masters
};
// term[l][p][:] = ls1_l * (W_proj_l . O16_l[rowidx[p], :] + b_proj_l) for every block l: ONE launch
int refine_proj_terms(const RefineLayer* layers, int depth, const int* rowidx, const int* count, int D,
                      int max_persons, float* term, cudaStream_t st);
// x[p][:] <- for every block: (x + term_l) + ls2_l * MLP_l(LN2_l(x + term_l)): ONE persistent cooperative launch
int refine_mlp_chain(const RefineLayer* layers, int depth, const int* count, int D, int max_persons, const float* term,
                     float* x, float* h, unsigned int* barrier, cudaStream_t st);
int hph_self_attn(const float* qkv, int ld, const int* det_b, const int* img_off, const int* count,
                  int max_persons, int heads, float* out, int ldo, cudaStream_t st);
int hph_cross_attn(const float* q, int ldq, const float* KV, int64_t ldkv, int k_col, int v_col,
                   const int* det_b, const int* count, int max_persons, int heads, int N, float* out,
                   int ldo, cudaStream_t st);
int person_post(const float* dec, int ld_dec, int num_betas, const float* offset, const float* K,
                const float* Kinv, const int* det_b, const int* det_y, const int* det_x, const int* count,
                int max_persons, float focal_norm, float* rotmat, float* rotvec, float* shape, float* expr,
                float* dist_pp, float* dist, float* loc, float* transl, float* K_det, cudaStream_t st);

// ---- smplx_lbs.cu ----------------------------------------------------------------------------
struct SmplxDeviceModel {
  int V = 0;            // vertices
  int L = 0;            // num_betas + 10 shape/expression coefficients
  int num_betas = 10;
  int center_idx = 15;  // person_center joint ('head')
  int ldp = 0;          // PDX / vt row pitch: 3V rounded up to 4
  const float* PDX = nullptr;          // [486 + L, ldp]
  const float* vt = nullptr;           // [ldp] v_template flattened
  const float* lbs_weights_padded = nullptr;  // [ceil(V/72)*72, 55], zero rows beyond V
  CUtensorMap tmPDX;                   // TMA descriptor of PDX (boxes of 16 rows x 220 columns)
  const float* Jt = nullptr;           // [55, 3]   J_regressor . v_template
  const float* Jdirs = nullptr;        // [55*3, L] J_regressor . shapedirs
  const int* parents = nullptr;        // [55]
  const int* extra_idx = nullptr;      // [21] vertex-picked joints
  const int* lmk_tri = nullptr;        // [51, 3] vertex ids of the landmark faces
  const float* lmk_bary = nullptr;     // [51, 3]
};
struct SmplxScratch {
  float* cf = nullptr;      // [max_persons, 486 + L]
  float* Amat = nullptr;    // [max_persons, 55, 12]
  float* xf = nullptr;      // [max_persons, 16]
  float* jposed = nullptr;  // [max_persons, 55, 3]
};
int smplx_make_tmap(SmplxDeviceModel* bm);
int smplx_tile_verts();
int smplx_build_pdx(const float* posedirs, const float* sdirs_full, int L, int V, int ldp, float* PDX,
                    cudaStream_t st);
int smplx_fold_jreg(const float* Jr, const float* M, int V, int Q, float* out, cudaStream_t st);
int smplx_forward(const SmplxDeviceModel& bm, const float* rotvec, const float* shape, const float* expr,
                  const float* transl, const float* K_det, const int* count, int max_persons,
                  SmplxScratch& ws, float* v3d, float* v2d, float* j3d, float* j2d, float* transl_pelvis,
                  cudaStream_t st);

}  // namespace mhmr

// Engine: weights, workspaces, TMA plans and the launch sequence of the whole Multi-HMR forward
// (reference model.py:205-349) behind the C-ABI of include/mhmr.h.
#include <cmath>
#include <map>
#include <memory>
#include <cstdlib>
#include <string>
#include <vector>

#include "gemm_tc.cuh"
#include "kernels.cuh"

using namespace mhmr;

namespace {

#define TRY(expr)                       \
  do {                                  \
    int rc_ = (expr);                   \
    if (rc_ != MHMR_OK) return rc_;     \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct ArchSpec { int D, depth, heads; };
const ArchSpec kArch[3] = {{384, 12, 6}, {768, 12, 12}, {1024, 24, 16}};
constexpr int kHphDim = 1024;
constexpr int kCamDim = 99;

struct VitLayer {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *bqkv, *bproj, *ls1, *bfc1, *bfc2, *ls2;
  __half *Wqkv, *Wproj, *Wfc1, *Wfc2;
  const float *Wproj32, *Wfc1_32, *Wfc2_32;  // fp32 masters (central-stream refinement)
  __half* O16;                               // this layer's attention output [max_batch*T, D]
  // norm1 / norm2 folded into qkv / fc1 (gemm_tc.cuh): biases with the LayerNorm beta folded in
  float *bqkv_f = nullptr, *bfc1_f = nullptr;
  GemmPlan qkv, proj, fc1, fc2;
};
struct HphLayer {
  const float *ln0_g, *ln0_b, *Wqkv, *Wsa_out, *bsa_out;
  const float *ln1_g, *ln1_b, *Wq, *Wca_out, *bca_out;
  const float *ln2_g, *ln2_b, *Wff0, *bff0, *Wff3, *bff3;
};

}  // namespace

struct mhmr_engine {
  mhmr_config cfg{};
  int D = 0, depth = 0, heads = 0, res = 0, N = 0, T = 0, C = 0, Cp = 0, Cq = 0, nkv = 0, ndec = 0;
  bool finalized = false;
  std::map<std::string, DevBuf> weights;   // raw fp32 device copies, keyed like the state_dict
  std::map<std::string, DevBuf> tables;    // int32 tables
  std::vector<void*> owned;                // everything cudaMalloc'ed (freed in destroy)
  int launches = 0;
  bool profiling = false;
  struct ProfEntry { int cat; cudaEvent_t a, b; };
  std::vector<ProfEntry> prof;
  std::vector<cudaEvent_t> event_pool;
  size_t events_used = 0;
  cudaEvent_t next_event() {
    if (events_used == event_pool.size()) {
      cudaEvent_t ev;
      cudaEventCreate(&ev);
      event_pool.push_back(ev);
    }
    return event_pool[events_used++];
  }

  // packed weights
  __half* Wpatch = nullptr;   // [D, 592]
  float* rowadd = nullptr;    // [N, D]  pos_embed[1:] + patch bias
  float* cls_pos = nullptr;   // [D]
  std::vector<VitLayer> vit;
  __half* Wcls0 = nullptr;    // [D, D]
  __half* Wkv16 = nullptr;    // [nkv, Cp]
  float* Wkv32 = nullptr;     // [nkv, Cq]
  float* Wte_q = nullptr;     // [1024, Cq]
  float* te_const = nullptr;  // [1024]
  float* Wdec = nullptr;      // [ndec, 1024]
  float* bdec = nullptr;      // [ndec]
  std::vector<HphLayer> hph;
  SmplxDeviceModel bm;

  // workspaces
  __half *A16 = nullptr, *Xn16 = nullptr, *QKV16 = nullptr, *O16 = nullptr, *H16 = nullptr, *ctx16 = nullptr;
  bool ln_fold = true;       // MHMR_LN_FOLD=0: separate LayerNorm kernels (A/B measurements)
  float2* ln_stats = nullptr;  // [max_batch*T, ln_slots] partial row statistics of the residual stream
  __half* Xlo = nullptr;       // lo plane of the two-term fp16 residual stream (hi plane = Xn16), gemm_tc.cuh
  int ln_slots = 0;
  float *X = nullptr, *z32 = nullptr, *scores_raw = nullptr, *KV32 = nullptr, *Kinv = nullptr;
  int *det = nullptr, *count = nullptr, *img_off = nullptr;
  // central-stream refinement: token rows, input patches, residual streams and MLP hidden of the detected persons
  int* r_rowidx = nullptr;
  float *r_patch = nullptr, *r_x = nullptr, *r_h = nullptr, *r_term = nullptr;
  RefineLayer* r_layers = nullptr;   // device array [depth]
  unsigned int* r_barrier = nullptr;
  const float* Wpatch32 = nullptr;
  float *zc = nullptr, *query = nullptr, *vals = nullptr, *dKV = nullptr, *offh = nullptr, *xa = nullptr,
        *qkvp = nullptr, *att = nullptr, *qca = nullptr, *ffh = nullptr, *dec = nullptr, *K_det = nullptr,
        *one_count_x = nullptr;
  int* one = nullptr;  // device int == 1 (count for load-time skinny launches)
  SmplxScratch sx;
  GemmPlan patch_plan, cls0_plan, kv_plan;
  int* h_count = nullptr;  // pinned host copy of the person count

  ~mhmr_engine() {
    for (void* p : owned) cudaFree(p);
    for (auto& kv : weights) cudaFree(kv.second.p);
    for (auto& kv : tables) cudaFree(kv.second.p);
    if (h_count) cudaFreeHost(h_count);
    for (cudaEvent_t ev : event_pool) cudaEventDestroy(ev);
  }

  template <typename T>
  int alloc(T** out, size_t count_elems) {
    void* p = nullptr;
    const size_t bytes = count_elems * sizeof(T);
    MHMR_CUDA_CHECK(cudaMalloc(&p, bytes > 0 ? bytes : 16));
    MHMR_CUDA_CHECK(cudaMemset(p, 0, bytes > 0 ? bytes : 16));
    owned.push_back(p);
    *out = static_cast<T*>(p);
    return MHMR_OK;
  }
  const float* w(const std::string& key, int64_t expect_numel = -1) {
    auto it = weights.find(key);
    if (it == weights.end()) {
      set_last_error("missing weight: " + key);
      return nullptr;
    }
    if (expect_numel >= 0 && static_cast<int64_t>(it->second.bytes / 4) != expect_numel) {
      set_last_error("weight " + key + " has " + std::to_string(it->second.bytes / 4) + " elements, expected " +
                     std::to_string(expect_numel));
      return nullptr;
    }
    return static_cast<const float*>(it->second.p);
  }
  const int* tab(const std::string& key, int64_t expect_numel) {
    auto it = tables.find(key);
    if (it == tables.end() || static_cast<int64_t>(it->second.bytes / 4) != expect_numel) {
      set_last_error("missing or mis-sized table: " + key);
      return nullptr;
    }
    return static_cast<const int*>(it->second.p);
  }
};

namespace {

#define NEEDW(var, key, numel)                          \
  const float* var = e->w(key, numel);                  \
  if (var == nullptr) return MHMR_ERR_STATE;

// CTA-pair 256x256 tiles (cta_group::2) whenever N is a multiple of 256, else single-CTA 128x128 tiles
int pick_bn(int N) { return (N % 256 == 0) ? 512 : 128; }

int to_f16(mhmr_engine* e, const float* src, int64_t lds, int rows, int cols, int64_t ldd, __half** out,
           cudaStream_t st) {
  TRY(e->alloc(out, static_cast<size_t>(rows) * ldd));
  return f32_to_f16_2d(src, lds, *out, ldd, rows, cols, st);
}

int finalize_vit(mhmr_engine* e, cudaStream_t st) {
  const int D = e->D, N = e->N, T = e->T, Bm = e->cfg.max_batch;
  const std::string enc = "backbone.encoder.";
  NEEDW(pw, enc + "patch_embed.proj.weight", static_cast<int64_t>(D) * 588);
  NEEDW(pb, enc + "patch_embed.proj.bias", D);
  NEEDW(cls, enc + "cls_token", D);
  const float* pos = e->w(enc + "pos_embed", static_cast<int64_t>(1 + N) * D);
  if (pos == nullptr) {
    set_last_error(std::string(get_last_error()) + " (pos_embed must be interpolated to the working grid [1,1+N,D])");
    return MHMR_ERR_STATE;
  }
  TRY(to_f16(e, pw, 588, D, 588, 592, &e->Wpatch, st));
  e->Wpatch32 = pw;
  TRY(e->alloc(&e->rowadd, static_cast<size_t>(N) * D));
  TRY(add_vec(pos + D, pb, e->rowadd, static_cast<int64_t>(N) * D, D, st));
  TRY(e->alloc(&e->cls_pos, D));
  TRY(add_vec(cls, pos, e->cls_pos, D, D, st));

  const size_t M = static_cast<size_t>(Bm) * T;
  TRY(e->alloc(&e->A16, static_cast<size_t>(Bm) * N * 592));
  TRY(e->alloc(&e->X, M * D));
  TRY(e->alloc(&e->Xn16, M * D));
  TRY(e->alloc(&e->QKV16, M * 3 * D));
  // one attention-output buffer per layer when the refinement pass needs the rows of every layer afterwards
  const size_t o_layers = e->cfg.refine_central ? static_cast<size_t>(e->depth) : 1;
  TRY(e->alloc(&e->O16, o_layers * M * D));
  TRY(e->alloc(&e->H16, M * 4 * D));
  {
    const char* lf = std::getenv("MHMR_LN_FOLD");
    e->ln_fold = !(lf != nullptr && lf[0] == '0');
  }
  e->ln_slots = gemm_stat_slots(D, pick_bn(D));
  if (e->ln_fold) {
    TRY(e->alloc(&e->ln_stats, M * e->ln_slots));
    TRY(e->alloc(&e->Xlo, M * D));
  }

  GemmEpi ep;
  ep.rowadd = e->rowadd; ep.out = e->X; ep.ldo = D; ep.rows_in = N; ep.rows_out = T; ep.row_off = 1;
  TRY(gemm_plan_init(&e->patch_plan, e->A16, 592, e->Wpatch, 592, Bm * N, D, 588, EPI_ROWADD_F32, ep, pick_bn(D)));

  e->vit.resize(e->depth);
  for (int l = 0; l < e->depth; ++l) {
    VitLayer& L = e->vit[l];
    const std::string b = enc + "blocks." + std::to_string(l) + ".";
    NEEDW(n1g, b + "norm1.weight", D) NEEDW(n1b, b + "norm1.bias", D)
    NEEDW(n2g, b + "norm2.weight", D) NEEDW(n2b, b + "norm2.bias", D)
    NEEDW(wqkv, b + "attn.qkv.weight", 3ll * D * D) NEEDW(bqkv, b + "attn.qkv.bias", 3 * D)
    NEEDW(wproj, b + "attn.proj.weight", static_cast<int64_t>(D) * D) NEEDW(bproj, b + "attn.proj.bias", D)
    NEEDW(ls1, b + "ls1.gamma", D) NEEDW(ls2, b + "ls2.gamma", D)
    NEEDW(wfc1, b + "mlp.fc1.weight", 4ll * D * D) NEEDW(bfc1, b + "mlp.fc1.bias", 4 * D)
    NEEDW(wfc2, b + "mlp.fc2.weight", 4ll * D * D) NEEDW(bfc2, b + "mlp.fc2.bias", D)
    L.ln1_g = n1g; L.ln1_b = n1b; L.ln2_g = n2g; L.ln2_b = n2b;
    L.Wproj32 = wproj; L.Wfc1_32 = wfc1; L.Wfc2_32 = wfc2;
    L.O16 = e->O16 + (e->cfg.refine_central ? static_cast<size_t>(l) * M * D : 0);
    L.bqkv = bqkv; L.bproj = bproj; L.ls1 = ls1; L.bfc1 = bfc1; L.bfc2 = bfc2; L.ls2 = ls2;
    TRY(to_f16(e, wproj, D, D, D, D, &L.Wproj, st));
    TRY(to_f16(e, wfc2, 4 * D, D, 4 * D, 4 * D, &L.Wfc2, st));
    GemmEpi a; a.out = e->QKV16; a.ldo = 3 * D;
    GemmEpi p; p.bias = bproj; p.gamma = ls1; p.out = e->X; p.ldo = D;
    GemmEpi f1; f1.out = e->H16; f1.ldo = 4 * D;
    GemmEpi f2; f2.bias = bfc2; f2.gamma = ls2; f2.out = e->X; f2.ldo = D;
    int epi_qkv = EPI_BIAS_F16, epi_fc1 = EPI_BIAS_GELU_F16, epi_proj = EPI_LS_RESID_F32, epi_fc2 = EPI_LS_RESID_F32;
    if (e->ln_fold) {
      // qkv / fc1 read the hi plane of the residual stream that the previous proj / fc2 epilogue (layer 0:
      // split_rowstats) left in Xn16, and normalise in their epilogue from the row statistics
      TRY(e->alloc(&L.Wqkv, static_cast<size_t>(3) * D * D));
      TRY(e->alloc(&L.bqkv_f, 3 * D));
      TRY(fold_ln_linear(wqkv, bqkv, n1g, n1b, L.Wqkv, L.bqkv_f, 3 * D, D, st));
      TRY(e->alloc(&L.Wfc1, static_cast<size_t>(4) * D * D));
      TRY(e->alloc(&L.bfc1_f, 4 * D));
      TRY(fold_ln_linear(wfc1, bfc1, n2g, n2b, L.Wfc1, L.bfc1_f, 4 * D, D, st));
      a.bias = L.bqkv_f; a.stats = e->ln_stats; a.stat_slots = e->ln_slots;
      f1.bias = L.bfc1_f; f1.stats = e->ln_stats; f1.stat_slots = e->ln_slots;
      epi_qkv = EPI_LN_BIAS_F16;
      epi_fc1 = EPI_LN_GELU_F16;
      // the residual stream lives in (Xn16, Xlo) = (hi, lo); hi is the A operand of qkv / fc1
      p.out = nullptr; p.x16 = e->Xn16; p.xlo = e->Xlo; p.ldx16 = D; p.stats = e->ln_stats; p.stat_slots = e->ln_slots;
      f2.out = nullptr; f2.x16 = e->Xn16; f2.xlo = e->Xlo; f2.ldx16 = D; f2.stats = e->ln_stats; f2.stat_slots = e->ln_slots;
      epi_proj = epi_fc2 = EPI_LS_RESID_SPLIT;
    } else {
      TRY(to_f16(e, wqkv, D, 3 * D, D, D, &L.Wqkv, st));
      TRY(to_f16(e, wfc1, D, 4 * D, D, D, &L.Wfc1, st));
      a.bias = bqkv;
      f1.bias = bfc1;
    }
    TRY(gemm_plan_init(&L.qkv, e->Xn16, D, L.Wqkv, D, static_cast<int>(M), 3 * D, D, epi_qkv, a, pick_bn(3 * D)));
    TRY(gemm_plan_init(&L.proj, L.O16, D, L.Wproj, D, static_cast<int>(M), D, D, epi_proj, p, pick_bn(D)));
    TRY(gemm_plan_init(&L.fc1, e->Xn16, D, L.Wfc1, D, static_cast<int>(M), 4 * D, D, epi_fc1, f1, pick_bn(4 * D)));
    TRY(gemm_plan_init(&L.fc2, e->H16, 4 * D, L.Wfc2, 4 * D, static_cast<int>(M), D, 4 * D, epi_fc2, f2, pick_bn(D)));
  }
  if (e->w(enc + "norm.weight", D) == nullptr || e->w(enc + "norm.bias", D) == nullptr) return MHMR_ERR_STATE;
  return MHMR_OK;
}

int finalize_head(mhmr_engine* e, cudaStream_t st) {
  const int D = e->D, N = e->N, Bm = e->cfg.max_batch, Pm = e->cfg.max_persons, C = e->C, Cp = e->Cp,
            Cq = e->Cq, nb = e->cfg.num_betas, depth = e->cfg.xat_depth, inner = e->cfg.xat_num_heads * 32;
  const int res = e->res;
  const size_t BN = static_cast<size_t>(Bm) * N;
  TRY(e->alloc(&e->z32, BN * D));
  TRY(e->alloc(&e->ctx16, BN * Cp));
  TRY(e->alloc(&e->scores_raw, BN));
  TRY(e->alloc(&e->KV32, BN * e->nkv));
  TRY(e->alloc(&e->Kinv, static_cast<size_t>(Bm) * 9));
  TRY(e->alloc(&e->det, static_cast<size_t>(3) * Pm));
  TRY(e->alloc(&e->count, 4));
  TRY(e->alloc(&e->img_off, static_cast<size_t>(Bm) + 1));
  TRY(e->alloc(&e->one, 4));
  const int one_h = 1;
  MHMR_CUDA_CHECK(cudaMemcpyAsync(e->one, &one_h, sizeof(int), cudaMemcpyHostToDevice, st));
  MHMR_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&e->h_count), sizeof(int)));

  // detection
  NEEDW(c0w, "mlp_classif.0.weight", static_cast<int64_t>(D) * D) NEEDW(c0b, "mlp_classif.0.bias", D)
  NEEDW(c2w, "mlp_classif.2.weight", D) NEEDW(c2b, "mlp_classif.2.bias", 1)
  (void)c2w; (void)c2b;
  TRY(to_f16(e, c0w, D, D, D, D, &e->Wcls0, st));
  GemmEpi ce; ce.bias = c0b; ce.out = e->H16; ce.ldo = D;
  TRY(gemm_plan_init(&e->cls0_plan, e->ctx16, Cp, e->Wcls0, D, static_cast<int>(BN), D, D, EPI_BIAS_RELU_F16, ce, pick_bn(D)));
  NEEDW(o0w, "mlp_offset.0.weight", static_cast<int64_t>(D) * D) NEEDW(o0b, "mlp_offset.0.bias", D)
  NEEDW(o2w, "mlp_offset.2.weight", 2ll * D) NEEDW(o2b, "mlp_offset.2.bias", 2)
  (void)o0w; (void)o0b; (void)o2w; (void)o2b;

  // HPH
  const std::string h = "x_attention_head.";
  for (const char* nm : {"cross_queries_x", "cross_queries_y", "cross_values_x", "cross_values_y"})
    if (e->w(h + nm, static_cast<int64_t>(res) * C) == nullptr) return MHMR_ERR_STATE;
  if (e->w("camera.freq_bands", 16) == nullptr) return MHMR_ERR_STATE;
  const int token_dim = 318 + nb + 3 + C;
  const std::string t = h + "transformer.";
  NEEDW(tew, t + "to_token_embedding.weight", static_cast<int64_t>(kHphDim) * token_dim)
  NEEDW(teb, t + "to_token_embedding.bias", kHphDim)
  NEEDW(pe, t + "pos_embedding", kHphDim)
  NEEDW(ipose, h + "init_body_pose", 318) NEEDW(ibetas, h + "init_betas", nb) NEEDW(icam, h + "init_cam", 3)
  TRY(e->alloc(&e->Wte_q, static_cast<size_t>(kHphDim) * Cq));
  TRY(repack_f32(tew, token_dim, 0, e->Wte_q, Cq, 0, kHphDim, C, true, st));
  // te_const = W_te[:, C:] . [init_pose | init_betas | init_cam] + bias + pos_embedding
  const int ni = 318 + nb + 3, nip = (ni + 3) & ~3;
  float *Wte_i = nullptr, *init_vec = nullptr, *tmpb = nullptr;
  TRY(e->alloc(&Wte_i, static_cast<size_t>(kHphDim) * nip));
  TRY(repack_f32(tew, token_dim, C, Wte_i, nip, 0, kHphDim, ni, true, st));
  TRY(e->alloc(&init_vec, nip));
  MHMR_CUDA_CHECK(cudaMemcpyAsync(init_vec, ipose, 318 * 4, cudaMemcpyDeviceToDevice, st));
  MHMR_CUDA_CHECK(cudaMemcpyAsync(init_vec + 318, ibetas, nb * 4, cudaMemcpyDeviceToDevice, st));
  MHMR_CUDA_CHECK(cudaMemcpyAsync(init_vec + 318 + nb, icam, 3 * 4, cudaMemcpyDeviceToDevice, st));
  TRY(e->alloc(&tmpb, kHphDim));
  TRY(add_vec(teb, pe, tmpb, kHphDim, kHphDim, st));
  TRY(e->alloc(&e->te_const, kHphDim));
  TRY(skinny_linear(init_vec, nip, e->one, 1, ni, Wte_i, nip, tmpb, kHphDim, nullptr, nullptr, 0.f, 0, nullptr, 0,
                    e->te_const, kHphDim, st));

  e->hph.resize(depth);
  TRY(e->alloc(&e->Wkv32, static_cast<size_t>(e->nkv) * Cq));
  for (int l = 0; l < depth; ++l) {
    HphLayer& L = e->hph[l];
    const std::string p = t + "transformer.layers." + std::to_string(l) + ".";
    NEEDW(a, p + "0.norm.weight", kHphDim) NEEDW(b, p + "0.norm.bias", kHphDim)
    NEEDW(c, p + "0.fn.to_qkv.weight", 3ll * inner * kHphDim)
    NEEDW(d, p + "0.fn.to_out.0.weight", static_cast<int64_t>(kHphDim) * inner) NEEDW(f, p + "0.fn.to_out.0.bias", kHphDim)
    NEEDW(g, p + "1.norm.weight", kHphDim) NEEDW(hh, p + "1.norm.bias", kHphDim)
    NEEDW(kvw, p + "1.fn.to_kv.weight", 2ll * inner * C)
    NEEDW(qw, p + "1.fn.to_q.weight", static_cast<int64_t>(inner) * kHphDim)
    NEEDW(co, p + "1.fn.to_out.0.weight", static_cast<int64_t>(kHphDim) * inner) NEEDW(cb, p + "1.fn.to_out.0.bias", kHphDim)
    NEEDW(n2g, p + "2.norm.weight", kHphDim) NEEDW(n2b, p + "2.norm.bias", kHphDim)
    NEEDW(f0, p + "2.fn.net.0.weight", static_cast<int64_t>(kHphDim) * kHphDim) NEEDW(f0b, p + "2.fn.net.0.bias", kHphDim)
    NEEDW(f3, p + "2.fn.net.3.weight", static_cast<int64_t>(kHphDim) * kHphDim) NEEDW(f3b, p + "2.fn.net.3.bias", kHphDim)
    L.ln0_g = a; L.ln0_b = b; L.Wqkv = c; L.Wsa_out = d; L.bsa_out = f;
    L.ln1_g = g; L.ln1_b = hh; L.Wq = qw; L.Wca_out = co; L.bca_out = cb;
    L.ln2_g = n2g; L.ln2_b = n2b; L.Wff0 = f0; L.bff0 = f0b; L.Wff3 = f3; L.bff3 = f3b;
    TRY(repack_f32(kvw, C, 0, e->Wkv32 + static_cast<size_t>(l) * 2 * inner * Cq, Cq, 0, 2 * inner, C, true, st));
  }
  TRY(e->alloc(&e->Wkv16, static_cast<size_t>(e->nkv) * Cp));
  TRY(f32_to_f16_2d(e->Wkv32, Cq, e->Wkv16, Cp, e->nkv, C, st));  // cols >= C stay zero
  GemmEpi ke; ke.out = e->KV32; ke.ldo = e->nkv;
  TRY(gemm_plan_init(&e->kv_plan, e->ctx16, Cp, e->Wkv16, Cp, static_cast<int>(BN), e->nkv, Cp, EPI_BIAS_F32, ke, pick_bn(e->nkv)));

  // decoders stacked: [pose6 318 | betas nb | cam 3 | expression 10], bias + init (model.py:571-575)
  e->ndec = 318 + nb + 3 + 10;
  NEEDW(dpw, h + "decpose.weight", 318ll * kHphDim) NEEDW(dpb, h + "decpose.bias", 318)
  NEEDW(dsw, h + "decshape.weight", static_cast<int64_t>(nb) * kHphDim) NEEDW(dsb, h + "decshape.bias", nb)
  NEEDW(dcw, h + "deccam.weight", 3ll * kHphDim) NEEDW(dcb, h + "deccam.bias", 3)
  NEEDW(dew, h + "decexpression.weight", 10ll * kHphDim) NEEDW(deb, h + "decexpression.bias", 10)
  TRY(e->alloc(&e->Wdec, static_cast<size_t>(e->ndec) * kHphDim));
  TRY(e->alloc(&e->bdec, e->ndec));
  const size_t rowb = kHphDim * sizeof(float);
  MHMR_CUDA_CHECK(cudaMemcpyAsync(e->Wdec, dpw, 318 * rowb, cudaMemcpyDeviceToDevice, st));
  MHMR_CUDA_CHECK(cudaMemcpyAsync(e->Wdec + 318 * kHphDim, dsw, nb * rowb, cudaMemcpyDeviceToDevice, st));
  MHMR_CUDA_CHECK(cudaMemcpyAsync(e->Wdec + (318 + nb) * kHphDim, dcw, 3 * rowb, cudaMemcpyDeviceToDevice, st));
  MHMR_CUDA_CHECK(cudaMemcpyAsync(e->Wdec + (321 + nb) * kHphDim, dew, 10 * rowb, cudaMemcpyDeviceToDevice, st));
  TRY(add_vec(dpb, ipose, e->bdec, 318, 318, st));
  TRY(add_vec(dsb, ibetas, e->bdec + 318, nb, nb, st));
  TRY(add_vec(dcb, icam, e->bdec + 318 + nb, 3, 3, st));
  MHMR_CUDA_CHECK(cudaMemcpyAsync(e->bdec + 321 + nb, deb, 10 * 4, cudaMemcpyDeviceToDevice, st));  // init_expression = 0

  // per-person buffers
  TRY(e->alloc(&e->zc, static_cast<size_t>(Pm) * D));
  TRY(e->alloc(&e->query, static_cast<size_t>(Pm) * Cq));
  TRY(e->alloc(&e->vals, static_cast<size_t>(Pm) * Cq));
  TRY(e->alloc(&e->dKV, static_cast<size_t>(Pm) * e->nkv));
  TRY(e->alloc(&e->offh, static_cast<size_t>(Pm) * D));
  TRY(e->alloc(&e->xa, static_cast<size_t>(Pm) * kHphDim));
  TRY(e->alloc(&e->qkvp, static_cast<size_t>(Pm) * 3 * inner));
  TRY(e->alloc(&e->att, static_cast<size_t>(Pm) * inner));
  TRY(e->alloc(&e->qca, static_cast<size_t>(Pm) * inner));
  TRY(e->alloc(&e->ffh, static_cast<size_t>(Pm) * kHphDim));
  TRY(e->alloc(&e->dec, static_cast<size_t>(Pm) * e->ndec));
  TRY(e->alloc(&e->K_det, static_cast<size_t>(Pm) * 9));
  if (e->cfg.refine_central) {
    TRY(e->alloc(&e->r_rowidx, Pm));
    TRY(e->alloc(&e->r_patch, static_cast<size_t>(Pm) * 592));
    TRY(e->alloc(&e->r_x, static_cast<size_t>(Pm) * D));
    TRY(e->alloc(&e->r_h, static_cast<size_t>(Pm) * 4 * D));
    TRY(e->alloc(&e->r_term, static_cast<size_t>(e->depth) * Pm * D));
    TRY(e->alloc(&e->r_barrier, 4));
    std::vector<RefineLayer> rl(e->depth);
    for (int l = 0; l < e->depth; ++l) {
      const VitLayer& L = e->vit[l];
      rl[l] = RefineLayer{L.O16, L.Wproj32, L.bproj, L.ls1, L.ln2_g, L.ln2_b, L.Wfc1_32, L.bfc1, L.Wfc2_32, L.bfc2, L.ls2};
    }
    TRY(e->alloc(&e->r_layers, rl.size()));
    MHMR_CUDA_CHECK(cudaMemcpyAsync(e->r_layers, rl.data(), rl.size() * sizeof(RefineLayer), cudaMemcpyHostToDevice, st));
    MHMR_CUDA_CHECK(cudaStreamSynchronize(st));  // rl lives on this stack frame
  }
  return MHMR_OK;
}

int finalize_body(mhmr_engine* e, cudaStream_t st) {
  const int V = e->cfg.num_verts, nb = e->cfg.num_betas, L = nb + 10, Pm = e->cfg.max_persons;
  NEEDW(vt, "smplx.v_template", 3ll * V)
  NEEDW(sd, "smplx.shapedirs", 3ll * V * nb)
  NEEDW(ed, "smplx.expr_dirs", 30ll * V)
  NEEDW(pd, "smplx.posedirs", 486ll * 3 * V)
  NEEDW(jr, "smplx.J_regressor", 55ll * V)
  NEEDW(lw, "smplx.lbs_weights", 55ll * V)
  NEEDW(bary, "smplx.lmk_bary_coords", 51 * 3)
  const int* parents = e->tab("smplx.parents", 55);
  const int* extra = e->tab("smplx.extra_joints_idxs", 21);
  const int* tri = e->tab("smplx.lmk_tri", 51 * 3);
  if (!parents || !extra || !tri) return MHMR_ERR_STATE;
  SmplxDeviceModel& bm = e->bm;
  bm.V = V; bm.L = L; bm.num_betas = nb; bm.center_idx = e->cfg.person_center_idx;
  bm.ldp = (3 * V + 3) & ~3;
  float *sfull = nullptr, *PDX = nullptr, *vtp = nullptr, *Jt = nullptr, *Jd = nullptr;
  TRY(e->alloc(&sfull, static_cast<size_t>(3) * V * L));
  TRY(repack_f32(sd, nb, 0, sfull, L, 0, 3 * V, nb, false, st));
  TRY(repack_f32(ed, 10, 0, sfull, L, nb, 3 * V, 10, false, st));
  TRY(e->alloc(&PDX, static_cast<size_t>(486 + L) * bm.ldp));
  TRY(smplx_build_pdx(pd, sfull, L, V, bm.ldp, PDX, st));
  TRY(e->alloc(&vtp, bm.ldp));
  MHMR_CUDA_CHECK(cudaMemcpyAsync(vtp, vt, 3ll * V * 4, cudaMemcpyDeviceToDevice, st));
  TRY(e->alloc(&Jt, 55 * 3));
  TRY(smplx_fold_jreg(jr, vt, V, 3, Jt, st));
  TRY(e->alloc(&Jd, static_cast<size_t>(55) * 3 * L));
  TRY(smplx_fold_jreg(jr, sfull, V, 3 * L, Jd, st));
  // skinning weights padded to whole 72-vertex tiles (the vertex kernel bulk-copies one tile per CTA)
  const int tv = smplx_tile_verts();
  const int Vpad = (V + tv - 1) / tv * tv;
  float* lwp = nullptr;
  TRY(e->alloc(&lwp, static_cast<size_t>(Vpad) * 55));
  MHMR_CUDA_CHECK(cudaMemcpyAsync(lwp, lw, 55ll * V * 4, cudaMemcpyDeviceToDevice, st));
  bm.PDX = PDX; bm.vt = vtp; bm.lbs_weights_padded = lwp; bm.Jt = Jt; bm.Jdirs = Jd;
  TRY(smplx_make_tmap(&bm));
  bm.parents = parents; bm.extra_idx = extra; bm.lmk_tri = tri; bm.lmk_bary = bary;
  TRY(e->alloc(&e->sx.cf, static_cast<size_t>(Pm) * (486 + L)));
  TRY(e->alloc(&e->sx.Amat, static_cast<size_t>(Pm) * 55 * 12));
  TRY(e->alloc(&e->sx.xf, static_cast<size_t>(Pm) * 16));
  TRY(e->alloc(&e->sx.jposed, static_cast<size_t>(Pm) * 55 * 3));
  return MHMR_OK;
}

// image of a forward: normalised fp32 NCHW (the reference's Model.forward input) or uint8 NHWC + the [3][256]
// normalize_rgb table (fused loader)
struct ImgSrc {
  const float* f32 = nullptr;
  const uint8_t* u8 = nullptr;
  const float* lut = nullptr;
};

struct ProfScope {
  mhmr_engine* e; int cat; cudaStream_t st; cudaEvent_t a{};
  ProfScope(mhmr_engine* e_, int cat_, cudaStream_t st_) : e(e_), cat(cat_), st(st_) {
    if (e->profiling) { a = e->next_event(); cudaEventRecord(a, st); }
  }
  ~ProfScope() {
    if (e->profiling) { cudaEvent_t b = e->next_event(); cudaEventRecord(b, st); e->prof.push_back({cat, a, b}); }
  }
};

#define LAUNCH(cat, expr)               \
  do {                                  \
    ProfScope ps_(e, cat, st);          \
    TRY(expr);                          \
    ++e->launches;                      \
  } while (0)

int run_plan(mhmr_engine* e, int cat, GemmPlan& plan, int M, cudaStream_t st) {
  ProfScope ps_(e, cat, st);
  GemmPlan p = plan;  // tensor maps were built for the maximum M; the kernel bounds rows by p.M
  p.M = M;
  p.grid = gemm_plan_grid(&p, M);
  TRY(gemm_plan_run(&p, st));
  ++e->launches;
  return MHMR_OK;
}

int vit_forward(mhmr_engine* e, const ImgSrc& x, int B, float* z_out, cudaStream_t st) {
  const int D = e->D, N = e->N, T = e->T, M = B * T;
  if (x.u8 != nullptr) {
    LAUNCH(MHMR_CAT_MISC, im2col_u8_patch14(x.u8, x.lut, e->A16, B, e->cfg.img_size, 592, st));
  } else {
    LAUNCH(MHMR_CAT_MISC, im2col_patch14(x.f32, e->A16, B, e->cfg.img_size, 592, st));
  }
  LAUNCH(MHMR_CAT_MISC, cls_rows(e->X, e->cls_pos, B, T, D, st));
  TRY(run_plan(e, MHMR_CAT_GEMM_OTHER, e->patch_plan, B * N, st));
  for (int l = 0; l < e->depth; ++l) {
    VitLayer& L = e->vit[l];
    // norm1 / norm2: folded into qkv / fc1 (statistics + raw fp16 rows come from the previous epilogue), or kernels
    if (!e->ln_fold) {
      LAUNCH(MHMR_CAT_LAYERNORM, layernorm(e->X, L.ln1_g, L.ln1_b, e->Xn16, D, nullptr, 0, M, D, 1e-6f, 0, 0, st));
    } else if (l == 0) {
      LAUNCH(MHMR_CAT_LAYERNORM, split_rowstats(e->X, e->Xn16, e->Xlo, D, e->ln_stats, e->ln_slots, M, D, st));
    }
    TRY(run_plan(e, MHMR_CAT_GEMM_QKV, L.qkv, M, st));
    LAUNCH(MHMR_CAT_ATTENTION, attention_forward(e->QKV16, 3 * D, L.O16, D, B, T, D, st));
    TRY(run_plan(e, MHMR_CAT_GEMM_PROJ, L.proj, M, st));
    if (!e->ln_fold)
      LAUNCH(MHMR_CAT_LAYERNORM, layernorm(e->X, L.ln2_g, L.ln2_b, e->Xn16, D, nullptr, 0, M, D, 1e-6f, 0, 0, st));
    TRY(run_plan(e, MHMR_CAT_GEMM_FC1, L.fc1, M, st));
    TRY(run_plan(e, MHMR_CAT_GEMM_FC2, L.fc2, M, st));
  }
  // final norm, cls dropped: fp32 features (head query side, optional user copy) + fp16 context columns
  const float* ng = e->w("backbone.encoder.norm.weight");
  const float* nb = e->w("backbone.encoder.norm.bias");
  if (e->ln_fold) {
    LAUNCH(MHMR_CAT_LAYERNORM,
           layernorm_split(e->Xn16, e->Xlo, ng, nb, e->ctx16, e->Cp, e->z32, D, M, D, 1e-6f, T, 1, st));
  } else {
    LAUNCH(MHMR_CAT_LAYERNORM, layernorm(e->X, ng, nb, e->ctx16, e->Cp, e->z32, D, M, D, 1e-6f, T, 1, st));
  }
  if (z_out != nullptr)
    MHMR_CUDA_CHECK(cudaMemcpyAsync(z_out, e->z32, static_cast<size_t>(B) * N * D * 4, cudaMemcpyDeviceToDevice, st));
  return MHMR_OK;
}

// Central-stream refinement.  The bulk pass computes every token with fp16 tensor-core operands;
// what the per-person outputs are sensitive to is the residual stream of the DETECTED tokens themselves (their
// own patch embedding and their own MLP / projection branches: 92 % of the feature error variance,
// tools/precision_study.py).  Those few rows are recomputed here in fp32 with the fp32 master weights:
//   x = patch-embed(pixels) + pos;  per block: x += ls1 * (Wproj . O16[row] + b);  x += ls2 * MLP(LN2(x))
// where O16[row] is the attention output of the bulk pass for that token (kept per layer).  The final norm is
// applied by person_gather.  Same arithmetic as dinov2 Block.forward (reached from blocks/dinov2.py:25).  Four
// launches: patches / row indices, patch embedding, every projection term at once, the MLP chain of all blocks in one
// persistent cooperative kernel (refine.cu).
int refine_streams(mhmr_engine* e, const ImgSrc& x, const int* det_b, const int* det_y, const int* det_x,
                   const int* count, cudaStream_t st) {
  const int D = e->D, Pm = e->cfg.max_persons;
  LAUNCH(MHMR_CAT_REFINE, refine_prepare(x.f32, x.u8, x.lut, e->cfg.img_size, e->rowadd, D, det_b, det_y, det_x, count, Pm, e->res,
                                         e->r_rowidx, e->r_patch, 592, e->r_x, st));
  SkinnyExtra none;
  LAUNCH(MHMR_CAT_REFINE, skinny_linear_ex(e->r_patch, 592, none, count, Pm, 588, e->Wpatch32, 588, nullptr, D, nullptr,
                                           nullptr, 0.f, 0, e->r_x, D, e->r_x, D, st));
  LAUNCH(MHMR_CAT_REFINE, refine_proj_terms(e->r_layers, e->depth, e->r_rowidx, count, D, Pm, e->r_term, st));
  LAUNCH(MHMR_CAT_REFINE, refine_mlp_chain(e->r_layers, e->depth, count, D, Pm, e->r_term, e->r_x, e->r_h, e->r_barrier, st));
  return MHMR_OK;
}

int head_forward(mhmr_engine* e, const ImgSrc& x, const float* K, int B, float det_thresh, int nms, const int64_t* forced_idx,
                 int forced_P, const mhmr_outputs* o, cudaStream_t st) {
  const int D = e->D, N = e->N, res = e->res, Pm = e->cfg.max_persons, Cq = e->Cq, nb = e->cfg.num_betas;
  const int heads = e->cfg.xat_num_heads, inner = heads * 32, BN = B * N;
  int* det_b = o->det_idx; int* det_y = o->det_idx + Pm; int* det_x = o->det_idx + 2 * Pm;
  int* count_true = o->count;   // true number of detections (may exceed max_persons: reported as an error)
  int* count = e->count + 2;    // clamped to max_persons: what the per-person kernels iterate over
  LAUNCH(MHMR_CAT_HEAD, invert_K(K, e->Kinv, B, st));
  LAUNCH(MHMR_CAT_HEAD, ctx_fourier(e->Kinv, e->w("camera.freq_bands"), e->ctx16, e->Cp, B, res, D, e->Cp - D, st));
  // detection (model.py:133-158)
  TRY(run_plan(e, MHMR_CAT_GEMM_OTHER, e->cls0_plan, BN, st));
  LAUNCH(MHMR_CAT_HEAD, rowdot_sigmoid(e->H16, D, e->w("mlp_classif.2.weight"), e->w("mlp_classif.2.bias"), e->scores_raw, BN, D, st));
  if (forced_idx != nullptr) {
    LAUNCH(MHMR_CAT_HEAD, forced_detections(e->scores_raw, o->scores_map, B, res, forced_idx, forced_P, det_b, det_y, det_x,
                         o->det_score, count_true, count, e->img_off, st));
  } else {
    LAUNCH(MHMR_CAT_HEAD, nms_compact(e->scores_raw, o->scores_map, B, res, nms, det_thresh, Pm, det_b, det_y, det_x,
                       o->det_score, count_true, count, e->img_off, st));
  }
  MHMR_CUDA_CHECK(cudaMemcpyAsync(e->h_count, count_true, sizeof(int), cudaMemcpyDeviceToHost, st));
  // keys / values of both decoder layers for every token (to_kv, cross_attn_transformer.py:187)
  TRY(run_plan(e, MHMR_CAT_GEMM_OTHER, e->kv_plan, BN, st));
  const std::string h = "x_attention_head.";
  const float* xr = nullptr;
  if (e->cfg.refine_central) {
    TRY(refine_streams(e, x, det_b, det_y, det_x, count, st));
    xr = e->r_x;
  }
  LAUNCH(MHMR_CAT_HEAD, person_gather(e->z32, xr, e->w("backbone.encoder.norm.weight"), e->w("backbone.encoder.norm.bias"),
                       e->Kinv, e->w("camera.freq_bands"), e->w(h + "cross_queries_x"),
                       e->w(h + "cross_queries_y"), e->w(h + "cross_values_x"), e->w(h + "cross_values_y"),
                       det_b, det_y, det_x, count, Pm, res, D, e->zc, e->query, e->vals, Cq, st));
  // offset head (model.py:258)
  LAUNCH(MHMR_CAT_HEAD, skinny_linear(e->zc, D, count, Pm, D, e->w("mlp_offset.0.weight"), D, e->w("mlp_offset.0.bias"), D,
                       nullptr, nullptr, 0.f, 1, nullptr, 0, e->offh, D, st));
  LAUNCH(MHMR_CAT_HEAD, skinny_linear(e->offh, D, count, Pm, D, e->w("mlp_offset.2.weight"), D, e->w("mlp_offset.2.bias"), 2,
                       nullptr, nullptr, 0.f, 0, nullptr, 0, o->offset, 2, st));
  // learned value embeddings injected at the detected cells (model.py:514-517)
  LAUNCH(MHMR_CAT_HEAD, skinny_linear(e->vals, Cq, count, Pm, e->C, e->Wkv32, Cq, nullptr, e->nkv, nullptr, nullptr, 0.f, 0,
                       nullptr, 0, e->dKV, e->nkv, st));
  LAUNCH(MHMR_CAT_HEAD, kv_add_rows(e->KV32, e->nkv, e->dKV, e->nkv, det_b, det_y, det_x, count, Pm, res, st));
  // token embedding (cross_attn_transformer.py:352-357)
  LAUNCH(MHMR_CAT_HEAD, skinny_linear(e->query, Cq, count, Pm, e->C, e->Wte_q, Cq, e->te_const, kHphDim, nullptr, nullptr, 0.f,
                       0, nullptr, 0, e->xa, kHphDim, st));
  for (int l = 0; l < e->cfg.xat_depth; ++l) {
    HphLayer& L = e->hph[l];
    LAUNCH(MHMR_CAT_HEAD, skinny_linear(e->xa, kHphDim, count, Pm, kHphDim, L.Wqkv, kHphDim, nullptr, 3 * inner, L.ln0_g, L.ln0_b,
                         1e-5f, 0, nullptr, 0, e->qkvp, 3 * inner, st));
    LAUNCH(MHMR_CAT_HEAD, hph_self_attn(e->qkvp, 3 * inner, det_b, e->img_off, count, Pm, heads, e->att, inner, st));
    LAUNCH(MHMR_CAT_HEAD, skinny_linear(e->att, inner, count, Pm, inner, L.Wsa_out, inner, L.bsa_out, kHphDim, nullptr, nullptr,
                         0.f, 0, e->xa, kHphDim, e->xa, kHphDim, st));
    LAUNCH(MHMR_CAT_HEAD, skinny_linear(e->xa, kHphDim, count, Pm, kHphDim, L.Wq, kHphDim, nullptr, inner, L.ln1_g, L.ln1_b, 1e-5f,
                         0, nullptr, 0, e->qca, inner, st));
    LAUNCH(MHMR_CAT_HEAD, hph_cross_attn(e->qca, inner, e->KV32, e->nkv, l * 2 * inner, l * 2 * inner + inner, det_b, count, Pm,
                          heads, N, e->att, inner, st));
    LAUNCH(MHMR_CAT_HEAD, skinny_linear(e->att, inner, count, Pm, inner, L.Wca_out, inner, L.bca_out, kHphDim, nullptr, nullptr,
                         0.f, 0, e->xa, kHphDim, e->xa, kHphDim, st));
    LAUNCH(MHMR_CAT_HEAD, skinny_linear(e->xa, kHphDim, count, Pm, kHphDim, L.Wff0, kHphDim, L.bff0, kHphDim, L.ln2_g, L.ln2_b,
                         1e-5f, 2, nullptr, 0, e->ffh, kHphDim, st));
    LAUNCH(MHMR_CAT_HEAD, skinny_linear(e->ffh, kHphDim, count, Pm, kHphDim, L.Wff3, kHphDim, L.bff3, kHphDim, nullptr, nullptr,
                         0.f, 0, e->xa, kHphDim, e->xa, kHphDim, st));
  }
  LAUNCH(MHMR_CAT_HEAD, skinny_linear(e->xa, kHphDim, count, Pm, kHphDim, e->Wdec, kHphDim, e->bdec, e->ndec, nullptr, nullptr, 0.f,
                       0, nullptr, 0, e->dec, e->ndec, st));
  const float focal_norm = static_cast<float>(e->cfg.img_size / (2.0 * tan(30.0 * 3.14159265358979323846 / 180.0)));
  LAUNCH(MHMR_CAT_HEAD, person_post(e->dec, e->ndec, nb, o->offset, K, e->Kinv, det_b, det_y, det_x, count, Pm, focal_norm,
                     o->rotmat, o->rotvec, o->shape, o->expression, o->dist_pp, o->dist, o->loc, o->transl,
                     e->K_det, st));
  {
    ProfScope ps_(e, MHMR_CAT_SMPLX, st);
    TRY(smplx_forward(e->bm, o->rotvec, o->shape, o->expression, o->transl, e->K_det, count, Pm, e->sx, o->v3d,
                      o->v2d, o->j3d, o->j2d, o->transl_pelvis, st));
  }
  e->launches += 3;
  return MHMR_OK;
}

}  // namespace

extern "C" {

int mhmr_create(const mhmr_config* cfg, mhmr_engine** out) {
  MHMR_REQUIRE(cfg != nullptr && out != nullptr, "null argument");
  MHMR_REQUIRE(cfg->arch >= 0 && cfg->arch <= 2, "arch must be 0 (S), 1 (B) or 2 (L)");
  MHMR_REQUIRE(cfg->img_size > 0 && cfg->img_size % 14 == 0, "Invalid img size");  // model.py:65
  MHMR_REQUIRE(cfg->max_batch > 0 && cfg->max_persons > 0, "capacities must be positive");
  MHMR_REQUIRE(cfg->num_betas == 10 || cfg->num_betas == 11, "num_betas must be 10 or 11");  // model.py:384
  MHMR_REQUIRE(cfg->xat_depth >= 1 && cfg->xat_depth <= 8 && cfg->xat_num_heads >= 1, "bad HPH geometry");
  MHMR_REQUIRE(cfg->person_center_idx >= 0 && cfg->person_center_idx < 55,
               "person_center must be one of the 55 kinematic joints");
  MHMR_REQUIRE(cfg->num_verts > 0, "num_verts must be positive");
  auto e = std::make_unique<mhmr_engine>();
  e->cfg = *cfg;
  const ArchSpec& a = kArch[cfg->arch];
  e->D = a.D; e->depth = a.depth; e->heads = a.heads;
  e->res = cfg->img_size / 14;
  e->N = e->res * e->res;
  e->T = e->N + 1;
  e->C = e->D + kCamDim;
  e->Cp = e->D + 128;                 // fp16 context pitch: D feature columns + 99 camera columns + zero pad
  e->Cq = (e->C + 3) & ~3;            // fp32 per-person pitch
  e->nkv = cfg->xat_depth * 2 * cfg->xat_num_heads * 32;
  MHMR_REQUIRE(e->nkv % 32 == 0, "HPH inner dim must be a multiple of 32");
  *out = e.release();
  return MHMR_OK;
}

int mhmr_destroy(mhmr_engine* h) {
  delete h;
  return MHMR_OK;
}

static int store_copy(std::map<std::string, DevBuf>& m, const char* key, const void* data, size_t bytes) {
  MHMR_REQUIRE(key != nullptr && data != nullptr && bytes > 0, "null/empty tensor");
  auto it = m.find(key);
  if (it != m.end()) {
    cudaFree(it->second.p);
    m.erase(it);
  }
  DevBuf b;
  b.bytes = bytes;
  MHMR_CUDA_CHECK(cudaMalloc(&b.p, bytes));
  MHMR_CUDA_CHECK(cudaMemcpy(b.p, data, bytes, cudaMemcpyDefault));
  m[key] = b;
  return MHMR_OK;
}

int mhmr_set_weight(mhmr_engine* h, const char* key, const float* data, int64_t numel) {
  MHMR_REQUIRE(h != nullptr, "null engine");
  if (h->finalized) { set_last_error("engine already finalized"); return MHMR_ERR_STATE; }
  return store_copy(h->weights, key, data, static_cast<size_t>(numel) * 4);
}

int mhmr_set_table_i32(mhmr_engine* h, const char* key, const int32_t* data, int64_t numel) {
  MHMR_REQUIRE(h != nullptr, "null engine");
  if (h->finalized) { set_last_error("engine already finalized"); return MHMR_ERR_STATE; }
  return store_copy(h->tables, key, data, static_cast<size_t>(numel) * 4);
}

int mhmr_finalize(mhmr_engine* h) {
  MHMR_REQUIRE(h != nullptr, "null engine");
  if (h->finalized) return MHMR_OK;
  cudaStream_t st = nullptr;
  TRY(finalize_vit(h, st));
  TRY(finalize_head(h, st));
  TRY(finalize_body(h, st));
  MHMR_CUDA_CHECK(cudaStreamSynchronize(st));
  h->finalized = true;
  return MHMR_OK;
}

static int forward_impl(mhmr_engine* h, const ImgSrc& x, const float* K, int B, float det_thresh,
                        int nms_kernel_size, const int64_t* forced_idx, int forced_P, const mhmr_outputs* out,
                        void* stream) {
  MHMR_REQUIRE(h != nullptr && K != nullptr && out != nullptr, "null argument");
  if (!h->finalized) { set_last_error("mhmr_forward before mhmr_finalize"); return MHMR_ERR_STATE; }
  MHMR_REQUIRE(B >= 1 && B <= h->cfg.max_batch, "batch exceeds max_batch");
  MHMR_REQUIRE(forced_idx == nullptr || (forced_P >= 0 && forced_P <= h->cfg.max_persons),
               "forced_P exceeds max_persons");
  MHMR_REQUIRE(out->scores_map && out->count && out->det_idx && out->det_score && out->offset && out->loc &&
                   out->dist_pp && out->dist && out->rotmat && out->rotvec && out->shape && out->expression &&
                   out->transl && out->transl_pelvis && out->v3d && out->j3d && out->j2d,
               "a required output buffer is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  h->launches = 0;
  TRY(vit_forward(h, x, B, out->z, st));
  return head_forward(h, x, K, B, det_thresh, nms_kernel_size, forced_idx, forced_P, out, st);
}

int mhmr_forward(mhmr_engine* h, const float* x, const float* K, int B, float det_thresh,
                 int nms_kernel_size, const int64_t* forced_idx, int forced_P, const mhmr_outputs* out,
                 void* stream) {
  MHMR_REQUIRE(x != nullptr, "null image");
  ImgSrc src;
  src.f32 = x;
  return forward_impl(h, src, K, B, det_thresh, nms_kernel_size, forced_idx, forced_P, out, stream);
}

int mhmr_forward_u8(mhmr_engine* h, const uint8_t* img_u8, const float* lut, const float* K, int B, float det_thresh,
                    int nms_kernel_size, const int64_t* forced_idx, int forced_P, const mhmr_outputs* out,
                    void* stream) {
  MHMR_REQUIRE(img_u8 != nullptr && lut != nullptr, "null image / table");
  ImgSrc src;
  src.u8 = img_u8;
  src.lut = lut;
  return forward_impl(h, src, K, B, det_thresh, nms_kernel_size, forced_idx, forced_P, out, stream);
}

int mhmr_sync_count(mhmr_engine* h, void* stream, int* num_persons) {
  MHMR_REQUIRE(h != nullptr && num_persons != nullptr, "null argument");
  MHMR_CUDA_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  *num_persons = *h->h_count;
  if (*h->h_count > h->cfg.max_persons) {
    set_last_error("detected " + std::to_string(*h->h_count) + " persons > max_persons " +
                   std::to_string(h->cfg.max_persons));
    return MHMR_ERR_CAPACITY;
  }
  return MHMR_OK;
}

int mhmr_vit_forward(mhmr_engine* h, const float* x, int B, float* z, void* stream) {
  MHMR_REQUIRE(h != nullptr && x != nullptr && z != nullptr, "null argument");
  if (!h->finalized) { set_last_error("mhmr_vit_forward before mhmr_finalize"); return MHMR_ERR_STATE; }
  MHMR_REQUIRE(B >= 1 && B <= h->cfg.max_batch, "batch exceeds max_batch");
  h->launches = 0;
  ImgSrc src;
  src.f32 = x;
  return vit_forward(h, src, B, z, static_cast<cudaStream_t>(stream));
}

int mhmr_smplx_forward(mhmr_engine* h, int P, const float* rotvec, const float* shape,
                       const float* expression, const float* loc, const float* dist, const float* K_det,
                       float* v3d, float* v2d, float* j3d, float* j2d, float* transl, float* transl_pelvis,
                       void* stream) {
  MHMR_REQUIRE(h != nullptr, "null engine");
  if (!h->finalized) { set_last_error("mhmr_smplx_forward before mhmr_finalize"); return MHMR_ERR_STATE; }
  MHMR_REQUIRE(P >= 1 && P <= h->cfg.max_persons, "P exceeds max_persons");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MHMR_CUDA_CHECK(cudaMemcpyAsync(h->count + 1, &P, sizeof(int), cudaMemcpyHostToDevice, st));
  TRY(loc_to_transl(loc, dist, K_det, P, transl, st));
  return smplx_forward(h->bm, rotvec, shape, expression, transl, K_det, h->count + 1, P, h->sx, v3d, v2d, j3d, j2d,
                       transl_pelvis, st);
}

int mhmr_last_launch_count(mhmr_engine* h) { return h != nullptr ? h->launches : 0; }

int mhmr_set_profiling(mhmr_engine* h, int enable) {
  MHMR_REQUIRE(h != nullptr, "null engine");
  h->profiling = enable != 0;
  h->prof.clear();
  h->events_used = 0;
  return MHMR_OK;
}

int mhmr_get_profile(mhmr_engine* h, float* ms_by_category, int* launches_by_category) {
  MHMR_REQUIRE(h != nullptr && ms_by_category != nullptr && launches_by_category != nullptr, "null argument");
  for (int c = 0; c < MHMR_NUM_CATEGORIES; ++c) { ms_by_category[c] = 0.f; launches_by_category[c] = 0; }
  for (const auto& p : h->prof) {
    MHMR_CUDA_CHECK(cudaEventSynchronize(p.b));
    float ms = 0.f;
    MHMR_CUDA_CHECK(cudaEventElapsedTime(&ms, p.a, p.b));
    ms_by_category[p.cat] += ms;
    launches_by_category[p.cat] += 1;
  }
  h->prof.clear();
  h->events_used = 0;
  return MHMR_OK;
}

}  // extern "C"

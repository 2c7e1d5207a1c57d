// tcgen05 + TMA + TMEM persistent GEMM (see gemm_tc.cuh for the contract).
//
// CTA = 384 threads, one CTA per SM, persistent over output tiles (128 x BN):
//   warp 0    : TMA producer   (one elected lane)   global -> 128B-swizzled smem ring
//   warp 1    : MMA issuer     (one elected lane)   tcgen05.mma kind::f16, M=128, N=BN, K=16
//   warp 2    : TMEM allocator (2 x BN fp32 columns: double-buffered accumulator)
//   warps 4-11: epilogue       tcgen05.ld 32x32b -> registers -> smem transpose -> fused epilogue with
//                              coalesced 128-bit global accesses (two warps per TMEM sub-partition)
// Pipelines: smem full/empty ring (TMA <-> MMA) and TMEM full/empty pair (MMA <-> epilogue), so the
// epilogue of tile i overlaps the main loop of tile i+1.
#include "gemm_tc.cuh"
#include "gemm_epilogue.cuh"

namespace mhmr {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 fp16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + 32 * kEpiWarps;  // TMA, MMA, TMEM-alloc, spare + epilogue warps
template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 3 : 5;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kEpiWarps * kScratchBytes + kBarBytes + 1024;  // +1024: manual align
  static constexpr int kTmemCols = 2 * BN;                                  // 256 or 512
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               int M, int N, int K, GemmEpi ep) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (128B swizzle atoms) by POINTER arithmetic, so that the compiler keeps the
  // shared address space (a round trip through uintptr_t degrades every access to generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* scratch_base = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes + kEpiWarps * kScratchBytes);
  uint64_t* full_bar = bars;                  // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;       // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;   // [2]        MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BK - 1) / BK;

  griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // the previous kernel's outputs (A rows, residual, statistics) are complete and visible

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    // (whole warp on the warp-uniform loop, one elected lane issues -- see elect_one_sync)
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc = make_idesc_f16(BM, BN, false, false);
    uint32_t stage = 0, phase = 0, acc_iter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++acc_iter) {
      const uint32_t as = acc_iter & 1u;
      const uint32_t aphase = (acc_iter >> 1) & 1u;
      mbar_wait(&tempty_bar[as], aphase ^ 1u);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kABytes;
        const uint64_t a_desc = make_sw128_desc(sa, 16, 1024);
        const uint64_t b_desc = make_sw128_desc(sb, 16, 1024);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance K by 16 fp16 = 32 B inside the 128-B swizzle row: +2 in (addr >> 4) units
            umma_f16_ss(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                       // smem slot free once MMAs retire
          if (kb == num_kb - 1) umma_commit(&tfull_bar[as]);    // accumulator complete
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ Epilogue -----------------------------------
    const int ew = warp & 3;             // TMEM sub-partition of this warp: lanes [32*ew, 32*ew+32)
    const int par = (warp - 4) >> 2;     // two warps per sub-partition split the column chunks
    float* scratch = scratch_base + (warp - 4) * (kScratchBytes / 4);
    uint32_t acc_iter = 0;
    EpiStatsPrefetch pf;
    if (static_cast<int>(blockIdx.x) < num_tiles)
      epilogue_load_row_stats<EPI>(pf, ep, M, (static_cast<int>(blockIdx.x) / num_n) * BM + ew * 32, lane);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++acc_iter) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const uint32_t as = acc_iter & 1u;
      const uint32_t aphase = (acc_iter >> 1) & 1u;
      const int m_base = m_blk * BM + ew * 32;
      EpiRowState rowst;
      epilogue_tile_begin<EPI>(rowst, pf, ep, K);
      if (tile + static_cast<int>(gridDim.x) < num_tiles)  // next tile's row statistics: under this tile's epilogue
        epilogue_load_row_stats<EPI>(pf, ep, M, ((tile + static_cast<int>(gridDim.x)) / num_n) * BM + ew * 32, lane);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
      constexpr int kChunksPerWarp = BN / 32 / (kEpiWarps / 4);
#pragma unroll 1
      for (int ci = 0; ci < kChunksPerWarp; ++ci) {
        const int c = ci * (kEpiWarps / 4) + par;
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tmem_ld_wait();
        if (ci == kChunksPerWarp - 1) {  // accumulator fully read by this warp: hand it back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        const int n0 = n_blk * BN + c * 32;
        if (n0 < N) epilogue_chunk<EPI>(r, scratch, ep, M, N, m_base, n0, lane, rowst);
      }
      epilogue_tile_end<EPI>(rowst, ep, M, m_base, n_blk * (kEpiWarps / 4) + par, lane);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN, int EPI>
int launch_one(const GemmPlan* p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_tc_kernel<BN, EPI>;
  static PerDeviceOnce once;
  if (once.first()) {
    MHMR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p->grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  MHMR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p->tmA, p->tmB, p->M, p->N, p->K, p->ep));
  return MHMR_OK;
}

template <int BN>
int launch_bn(const GemmPlan* p, cudaStream_t stream) {
  switch (p->epi) {
    case EPI_BIAS_F16: return launch_one<BN, EPI_BIAS_F16>(p, stream);
    case EPI_BIAS_GELU_F16: return launch_one<BN, EPI_BIAS_GELU_F16>(p, stream);
    case EPI_BIAS_RELU_F16: return launch_one<BN, EPI_BIAS_RELU_F16>(p, stream);
    case EPI_LS_RESID_F32: return launch_one<BN, EPI_LS_RESID_F32>(p, stream);
    case EPI_ROWADD_F32: return launch_one<BN, EPI_ROWADD_F32>(p, stream);
    case EPI_BIAS_F32: return launch_one<BN, EPI_BIAS_F32>(p, stream);
    case EPI_LS_RESID_SPLIT: return launch_one<BN, EPI_LS_RESID_SPLIT>(p, stream);
    case EPI_LN_BIAS_F16: return launch_one<BN, EPI_LN_BIAS_F16>(p, stream);
    case EPI_LN_GELU_F16: return launch_one<BN, EPI_LN_GELU_F16>(p, stream);
    default: break;
  }
  set_last_error("gemm: unknown epilogue kind");
  return MHMR_ERR_ARG;
}

}  // namespace

int gemm_plan_init(GemmPlan* plan, const __half* A, int64_t lda, const __half* W, int64_t ldw, int M,
                   int N, int K, int epi_kind, const GemmEpi& ep, int bn) {
  MHMR_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem");
  MHMR_REQUIRE(N % 32 == 0, "gemm: N must be a multiple of 32");
  MHMR_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "gemm: row pitches must be multiples of 8 fp16 (16 B, TMA)");
  MHMR_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  MHMR_REQUIRE(bn == 128 || bn == 256 || bn == 512, "gemm: block_n must be 128, 256 or 512 (CTA pair)");
  MHMR_REQUIRE(epi_kind >= 0 && epi_kind < EPI_NUM_KINDS, "gemm: bad epilogue kind");
  if (epi_kind != EPI_LS_RESID_SPLIT)
    MHMR_REQUIRE(ep.out != nullptr && ep.ldo % 8 == 0, "gemm: output missing or pitch not multiple of 8");
  if (epi_kind != EPI_BIAS_F32 && epi_kind != EPI_ROWADD_F32)
    MHMR_REQUIRE(ep.bias != nullptr, "gemm: bias required for this epilogue");
  if (epi_kind == EPI_LS_RESID_F32 || epi_kind == EPI_LS_RESID_SPLIT)
    MHMR_REQUIRE(ep.gamma != nullptr, "gemm: gamma required");
  if (epi_kind == EPI_LS_RESID_SPLIT)
    MHMR_REQUIRE(ep.x16 != nullptr && ep.xlo != nullptr && ep.ldx16 % 8 == 0 && ep.stats != nullptr &&
                     ep.stat_slots == gemm_stat_slots(N, bn) && N % (bn == 128 ? 128 : 256) == 0,
                 "gemm: split-stream epilogue needs both planes, stats, whole column tiles and matching slot count");
  if (epi_kind == EPI_LN_BIAS_F16 || epi_kind == EPI_LN_GELU_F16)
    MHMR_REQUIRE(ep.stats != nullptr && ep.stat_slots > 0 && ep.stat_slots % 2 == 0 && ep.stat_slots <= 8,
                 "gemm: folded-LN consumer needs row statistics (even slot count, at most 8)");
  if (epi_kind == EPI_ROWADD_F32)
    MHMR_REQUIRE(ep.rowadd != nullptr && ep.rows_in > 0, "gemm: rowadd/rows_in required");
  plan->M = M; plan->N = N; plan->K = K; plan->bn = bn; plan->epi = epi_kind; plan->ep = ep;
  int rc = make_tmap_2d(&plan->tmA, A, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, M, K, lda * 2, BM, BK, true);
  if (rc != MHMR_OK) return rc;
  // CTA pair: every CTA stages half (128 rows) of the 256-row weight tile
  rc = make_tmap_2d(&plan->tmB, W, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, N, K, ldw * 2, bn == 512 ? 128 : bn, BK,
                    true);
  if (rc != MHMR_OK) return rc;
  plan->grid = gemm_plan_grid(plan, M);
  return MHMR_OK;
}

int gemm_plan_grid(const GemmPlan* plan, int M) {
  const int sms = device_sm_count();
  if (plan->bn == 512) {
    const int tiles = ((M + 255) / 256) * ((plan->N + 255) / 256);
    const int pairs = sms / 2;
    return 2 * (tiles < pairs ? tiles : pairs);
  }
  const int tiles = ((M + BM - 1) / BM) * ((plan->N + plan->bn - 1) / plan->bn);
  return tiles < sms ? tiles : sms;
}

int gemm_stat_slots(int N, int bn) {
  const int tile_n = (bn == 128) ? 128 : 256;  // bn 512 = CTA pair with 256-column tiles
  return ((N + tile_n - 1) / tile_n) * (kEpiWarps / 4);
}

int gemm_plan_run(const GemmPlan* plan, cudaStream_t stream) {
  if (plan->bn == 512) return gemm_plan_run_2cta(plan, stream);
  return plan->bn == 256 ? launch_bn<256>(plan, stream) : launch_bn<128>(plan, stream);
}

}  // namespace mhmr

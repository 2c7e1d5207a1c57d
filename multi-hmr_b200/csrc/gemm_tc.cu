// tcgen05 + TMA + TMEM persistent GEMM (see gemm_tc.cuh for the contract).
//
// CTA = 256 threads, one CTA per SM, persistent over output tiles (128 x BN):
//   warp 0   : TMA producer   (one elected lane)   global -> 128B-swizzled smem ring
//   warp 1   : MMA issuer     (one elected lane)   tcgen05.mma kind::f16, M=128, N=BN, K=16
//   warp 2   : TMEM allocator (2 x BN fp32 columns: double-buffered accumulator)
//   warps 4-7: epilogue       tcgen05.ld 32x32b -> registers -> fused epilogue -> global
// Pipelines: smem full/empty ring (TMA <-> MMA) and TMEM full/empty pair (MMA <-> epilogue), so the
// epilogue of tile i overlaps the main loop of tile i+1.
#include "gemm_tc.cuh"

namespace mhmr {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 fp16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 256;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +1024: manual align
  static constexpr int kTmemCols = 2 * BN;                                    // 256 or 512
};

// One thread owns one output row; `r` holds 32 consecutive fp32 accumulator columns [n0, n0+32).
template <int EPI>
__device__ __forceinline__ void epilogue_store_chunk(const uint32_t (&r)[32], const GemmEpi& ep,
                                                     int N, int64_t out_row, int row_in_group,
                                                     int n0) {
  if constexpr (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16 || EPI == EPI_BIAS_RELU_F16) {
    __half* out = reinterpret_cast<__half*>(ep.out) + out_row * ep.ldo + n0;
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float x[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 b = __ldg(b4 + q * 2 + h);
        x[h * 4 + 0] = __uint_as_float(r[q * 8 + h * 4 + 0]) + b.x;
        x[h * 4 + 1] = __uint_as_float(r[q * 8 + h * 4 + 1]) + b.y;
        x[h * 4 + 2] = __uint_as_float(r[q * 8 + h * 4 + 2]) + b.z;
        x[h * 4 + 3] = __uint_as_float(r[q * 8 + h * 4 + 3]) + b.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if constexpr (EPI == EPI_BIAS_GELU_F16) x[i] = gelu_erf(x[i]);
        if constexpr (EPI == EPI_BIAS_RELU_F16) x[i] = fmaxf(x[i], 0.0f);
      }
      const __half2 h0 = __floats2half2_rn(x[0], x[1]);
      const __half2 h1 = __floats2half2_rn(x[2], x[3]);
      const __half2 h2 = __floats2half2_rn(x[4], x[5]);
      const __half2 h3 = __floats2half2_rn(x[6], x[7]);
      uint4 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&h0);
      pk.y = *reinterpret_cast<const uint32_t*>(&h1);
      pk.z = *reinterpret_cast<const uint32_t*>(&h2);
      pk.w = *reinterpret_cast<const uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(out + q * 8) = pk;
    }
  } else {
    float* out = reinterpret_cast<float*>(ep.out) + out_row * ep.ldo + n0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 a = make_float4(__uint_as_float(r[q * 4 + 0]), __uint_as_float(r[q * 4 + 1]),
                             __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3]));
      if constexpr (EPI == EPI_LS_RESID_F32) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + q);
        const float4 g = __ldg(reinterpret_cast<const float4*>(ep.gamma + n0) + q);
        const float4 x = *reinterpret_cast<const float4*>(out + q * 4);
        a.x = x.x + g.x * (a.x + b.x);
        a.y = x.y + g.y * (a.y + b.y);
        a.z = x.z + g.z * (a.z + b.z);
        a.w = x.w + g.w * (a.w + b.w);
      } else if constexpr (EPI == EPI_ROWADD_F32) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(
                                   ep.rowadd + static_cast<int64_t>(row_in_group) * N + n0) + q);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      } else {  // EPI_BIAS_F32
        if (ep.bias != nullptr) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + q);
          a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
      }
      *reinterpret_cast<float4*>(out + q * 4) = a;
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               int M, int N, int K, GemmEpi ep) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
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
      mbar_init(&tempty_bar[s], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (lane == 0) {
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
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance K by 16 fp16 = 32 B inside the 128-B swizzle row: +2 in (addr >> 4) units
            umma_f16_ss(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                       // smem slot free once MMAs retire
          if (kb == num_kb - 1) umma_commit(&tfull_bar[as]);    // accumulator complete
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ Epilogue -----------------------------------
    const int ew = warp & 3;  // TMEM sub-partition of this warp: lanes [32*ew, 32*ew+32)
    uint32_t acc_iter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++acc_iter) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const uint32_t as = acc_iter & 1u;
      const uint32_t aphase = (acc_iter >> 1) & 1u;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int m = m_blk * BM + ew * 32 + lane;
      int64_t out_row = m;
      int row_in_group = 0;
      if (ep.rows_in > 0) {
        const int g = m / ep.rows_in;
        row_in_group = m - g * ep.rows_in;
        out_row = static_cast<int64_t>(g) * ep.rows_out + ep.row_off + row_in_group;
      }
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tmem_ld_wait();
        const int n0 = n_blk * BN + c * 32;
        if (m < M && n0 < N) epilogue_store_chunk<EPI>(r, ep, N, out_row, row_in_group, n0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
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
  static bool attr_set = false;
  if (!attr_set) {
    MHMR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
    attr_set = true;
  }
  kern<<<p->grid, kThreads, Cfg::kSmemBytes, stream>>>(p->tmA, p->tmB, p->M, p->N, p->K, p->ep);
  MHMR_CUDA_CHECK(cudaGetLastError());
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
  MHMR_REQUIRE(bn == 128 || bn == 256, "gemm: BN must be 128 or 256");
  MHMR_REQUIRE(epi_kind >= 0 && epi_kind < EPI_NUM_KINDS, "gemm: bad epilogue kind");
  MHMR_REQUIRE(ep.out != nullptr && ep.ldo % 8 == 0, "gemm: output missing or pitch not multiple of 8");
  if (epi_kind != EPI_BIAS_F32 && epi_kind != EPI_ROWADD_F32)
    MHMR_REQUIRE(ep.bias != nullptr, "gemm: bias required for this epilogue");
  if (epi_kind == EPI_LS_RESID_F32) MHMR_REQUIRE(ep.gamma != nullptr, "gemm: gamma required");
  if (epi_kind == EPI_ROWADD_F32)
    MHMR_REQUIRE(ep.rowadd != nullptr && ep.rows_in > 0, "gemm: rowadd/rows_in required");
  plan->M = M; plan->N = N; plan->K = K; plan->bn = bn; plan->epi = epi_kind; plan->ep = ep;
  int rc = make_tmap_2d(&plan->tmA, A, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, M, K, lda * 2, BM, BK, true);
  if (rc != MHMR_OK) return rc;
  rc = make_tmap_2d(&plan->tmB, W, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, N, K, ldw * 2, bn, BK, true);
  if (rc != MHMR_OK) return rc;
  const int tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
  const int sms = device_sm_count();
  plan->grid = tiles < sms ? tiles : sms;
  return MHMR_OK;
}

int gemm_plan_run(const GemmPlan* plan, cudaStream_t stream) {
  return plan->bn == 256 ? launch_bn<256>(plan, stream) : launch_bn<128>(plan, stream);
}

}  // namespace mhmr

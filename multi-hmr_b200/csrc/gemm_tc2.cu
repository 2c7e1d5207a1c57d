// tcgen05 GEMM with CTA pairs (cta_group::2): 256 x 256 output tile per pair of SMs.
//
// A 128x256 single-CTA tile needs 96 B/clk of shared-memory reads for the MMA plus 96 B/clk of TMA writes
// — more than the 128 B/clk an SM's shared memory delivers, which caps the tensor pipe near 2/3 (measured:
// 69 %).  In a CTA pair each SM stages its own 128 rows of A and only HALF of the B tile (128 of the 256
// weight rows); the tensor cores of both SMs read the two halves through the pair datapath, so the per-SM
// traffic drops to 64 + 64 B/clk and the L2 -> SM traffic for B halves.
//
// Cluster (2,1,1), persistent over pair-tiles.  Per CTA (384 threads):
//   warp 0    : TMA producer — both CTAs load their own A rows / B half, completion lands on the LEADER's
//               full barrier (cp.async.bulk.tensor ... cta_group::2, peer bit of the barrier address cleared)
//   warp 1    : MMA issuer — leader CTA only: tcgen05.mma.cta_group::2, M = 256, N = 256, K = 16;
//               tcgen05.commit multicast frees the smem stage / publishes the accumulator in BOTH CTAs
//   warp 2    : TMEM allocator (cta_group::2, 2 x 256 fp32 columns: double-buffered accumulator per CTA)
//   warps 4-11: epilogue of this CTA's 128 rows (gemm_epilogue.cuh); the peer's epilogue warps hand the
//               accumulator back with remote arrives on the leader's barrier.
#include "gemm_epilogue.cuh"

namespace mhmr {

namespace {

constexpr int BM2 = 256;   // rows per pair tile (128 per CTA)
constexpr int BN2 = 256;   // columns per pair tile (each CTA stages 128 weight rows)
constexpr int BK2 = 64;
constexpr int kEpiWarps2 = 8;
constexpr int kThreads2 = 128 + 32 * kEpiWarps2;
constexpr int kABytes2 = 128 * BK2 * 2;         // 16 KB: this CTA's A rows
constexpr int kBBytes2 = (BN2 / 2) * BK2 * 2;   // 16 KB: this CTA's half of the B tile
constexpr int kStageBytes2 = kABytes2 + kBBytes2;
// smem ring depth: 5 stages.  PRE (attn.proj on the split stream: short K, epilogue-bound) trades one stage for the
// residual prefetch buffers of its epilogue warps; mlp.fc2 (K = 4D, tensor-bound) keeps 5 stages and direct loads
// (measured: fc2 +16 us per launch with 4 stages).
template <bool PRE>
constexpr int stages2() { return PRE ? 4 : 5; }
template <bool PRE>
constexpr int smem_bytes2() {
  return stages2<PRE>() * kStageBytes2 + kEpiWarps2 * kScratchBytes + 256 + 1024 + (PRE ? kEpiWarps2 * kPrefetchBytes : 0);
}
constexpr int kTmemCols2 = 2 * BN2;              // 512
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the barrier at the same smem offset in the leader CTA (rank 0) of the pair.
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  const uint32_t addr = smem_u32(bar) & kPeerBitMask;
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                                int32_t c_inner, int32_t c_outer) {
  const uint32_t mbar = smem_u32(bar) & kPeerBitMask;  // transaction bytes are counted by the leader
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(mbar), "r"(c_inner),
        "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all previously issued MMAs retire) on the barrier at this offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      :
      : "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(NCOLS)
               : "memory");
}

template <int EPI, bool PRE = false>
__global__ void __launch_bounds__(kThreads2, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                int M, int N, int K, GemmEpi ep) {
  constexpr int kStages2 = stages2<PRE>();
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (128B swizzle atoms) by POINTER arithmetic, so that the compiler keeps the
  // shared address space (a round trip through uintptr_t degrades every access to generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* scratch_base = reinterpret_cast<float*>(smem + kStages2 * kStageBytes2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages2 * kStageBytes2 + kEpiWarps2 * kScratchBytes);
  uint64_t* full_bar = bars;                   // [kStages2]  used in the leader: both CTAs' TMA -> MMA
  uint64_t* empty_bar = bars + kStages2;       // [kStages2]  MMA (multicast commit) -> this CTA's TMA
  uint64_t* tfull_bar = bars + 2 * kStages2;   // [2]         MMA (multicast commit) -> this CTA's epilogue
  uint64_t* tempty_bar = tfull_bar + 2;        // [2]         used in the leader: both epilogues -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* prefetch_base = reinterpret_cast<uint8_t*>(bars) + 256;   // [kEpiWarps2][kPrefetchBytes] (EPI_LS_RESID_SPLIT)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  const int num_m = (M + BM2 - 1) / BM2;
  const int num_n = (N + BN2 - 1) / BN2;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BK2 - 1) / BK2;

  griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's arrive.expect_tx + the peer producer's remote arrive
      mbar_init(&empty_bar[s], 1);  // one multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * kEpiWarps2);  // epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta<kTmemCols2>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  griddep_wait();  // the previous kernel's outputs (A rows, residual, statistics) are complete and visible
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------
    // The whole warp runs the (warp-uniform) loop and one elected lane issues: under `if (lane == 0)` the
    // compiler wraps every TMA / MMA / commit in an elect-and-retry loop (~80 clk each), see elect_one_sync.
    uint32_t stage = 0, phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int m0 = m_blk * BM2 + static_cast<int>(rank) * 128;
      const int n0 = n_blk * BN2 + static_cast<int>(rank) * (BN2 / 2);
      if constexpr (EPI == EPI_LS_RESID_F32 || EPI == EPI_LS_RESID_SPLIT) {
        // The epilogue of proj / fc2 reads 128 KB of the residual stream per tile straight from HBM (written a
        // whole layer ago) and is latency-bound on it (r01: proj at 44 % DRAM).  The producer warp runs at most the
        // smem ring ahead of the tensor pipe: when it starts a tile it pulls that tile's residual rows into L2, a
        // whole main loop before the epilogue asks for them (prefetch.global.L2 is a hint: no hazard with the
        // in-place update).
        const int nc0 = n_blk * BN2;
        const int ncols = min(BN2, N - nc0);
        for (int r = lane; r < 128; r += 32) {
          if (m0 + r >= M) break;
          if constexpr (EPI == EPI_LS_RESID_F32) {
            const char* row = reinterpret_cast<const char*>(reinterpret_cast<const float*>(ep.out) +
                                                            static_cast<int64_t>(m0 + r) * ep.ldo + nc0);
            for (int b = 0; b < ncols * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + b));
          } else {
            const int64_t off = static_cast<int64_t>(m0 + r) * ep.ldx16 + nc0;
            const char* rh = reinterpret_cast<const char*>(ep.x16 + off);
            const char* rl = reinterpret_cast<const char*>(ep.xlo + off);
            for (int b = 0; b < ncols * 2; b += 128) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(rh + b));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(rl + b));
            }
          }
        }
        __syncwarp();
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem + stage * kStageBytes2;
        uint8_t* sb = sa + kABytes2;
        if (elect_one_sync()) {
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes2);
          else mbar_arrive_leader(&full_bar[stage]);
          tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * BK2, m0);
          tma_load_2d_2sm(sb, &tmB, &full_bar[stage], kb * BK2, n0);
        }
        __syncwarp();
        if (++stage == kStages2) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader only) ------------------
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16(BM2, BN2, false, false);
      uint32_t stage = 0, phase = 0, acc_iter = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++acc_iter) {
        const uint32_t as = acc_iter & 1u;
        const uint32_t aphase = (acc_iter >> 1) & 1u;
        mbar_wait(&tempty_bar[as], aphase ^ 1u);  // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN2;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes2);
          const uint32_t sb = sa + kABytes2;
          const uint64_t a_desc = make_sw128_desc(sa, 16, 1024);
          const uint64_t b_desc = make_sw128_desc(sb, 16, 1024);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK2 / 16; ++k)
              umma_f16_ss_2cta(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_2cta(&empty_bar[stage]);                     // smem stage free in both CTAs
            if (kb == num_kb - 1) umma_commit_2cta(&tfull_bar[as]);  // accumulator complete in both CTAs
          }
          __syncwarp();
          if (++stage == kStages2) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ Epilogue (this CTA's 128 rows) ------------
    const int ew = warp & 3;
    const int par = (warp - 4) >> 2;
    float* scratch = scratch_base + (warp - 4) * (kScratchBytes / 4);
    uint32_t acc_iter = 0;
    const int m_off = static_cast<int>(rank) * 128 + ew * 32;
    EpiStatsPrefetch pf;
    if (pair < num_tiles) epilogue_load_row_stats<EPI>(pf, ep, M, (pair / num_n) * BM2 + m_off, lane);
    uint8_t* pre = PRE ? prefetch_base + (warp - 4) * kPrefetchBytes : nullptr;
    if (PRE && pair < num_tiles)   // residual values of this warp's first chunk: in flight under the first main loop
      epilogue_prefetch_resid<EPI>(pre, ep, M, (pair / num_n) * BM2 + m_off, (pair % num_n) * BN2 + par * 32, lane);
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++acc_iter) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const uint32_t as = acc_iter & 1u;
      const uint32_t aphase = (acc_iter >> 1) & 1u;
      const int m_base = m_blk * BM2 + m_off;
      EpiRowState rowst;
      epilogue_tile_begin<EPI>(rowst, pf, ep, K);
      if (tile + num_pairs < num_tiles)  // the next tile's row statistics travel under this tile's epilogue
        epilogue_load_row_stats<EPI>(pf, ep, M, ((tile + num_pairs) / num_n) * BM2 + m_off, lane);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN2;
      constexpr int kChunksPerWarp = BN2 / 32 / (kEpiWarps2 / 4);
#pragma unroll 1
      for (int ci = 0; ci < kChunksPerWarp; ++ci) {
        const int c = ci * (kEpiWarps2 / 4) + par;
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tmem_ld_wait();
        if (ci == kChunksPerWarp - 1) {  // accumulator fully read by this warp: hand it back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(&tempty_bar[as]);
            else mbar_arrive_leader(&tempty_bar[as]);
          }
        }
        const int n0 = n_blk * BN2 + c * 32;
        // next chunk of this warp: same tile, or the first chunk of its next tile
        int next_m = m_base, next_n0 = n0 + (kEpiWarps2 / 4) * 32;
        if (ci == kChunksPerWarp - 1) {
          const int nt = tile + num_pairs;
          next_m = (nt / num_n) * BM2 + m_off;
          next_n0 = (nt < num_tiles) ? (nt % num_n) * BN2 + par * 32 : -1;
        }
        if (n0 < N) epilogue_chunk<EPI>(r, scratch, ep, M, N, m_base, n0, lane, rowst, pre, next_m, next_n0);
      }
      epilogue_tile_end<EPI>(rowst, ep, M, m_base, n_blk * (kEpiWarps2 / 4) + par, lane);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still address it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta<kTmemCols2>(tmem_base);
  }
}

template <int EPI, bool PRE>
int launch_2cta_pre(const GemmPlan* p, cudaStream_t stream) {
  auto kern = gemm_tc2_kernel<EPI, PRE>;
  static PerDeviceOnce once;
  if (once.first()) {
    MHMR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes2<PRE>()));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p->grid);
  cfg.blockDim = dim3(kThreads2);
  cfg.dynamicSmemBytes = smem_bytes2<PRE>();
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  MHMR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p->tmA, p->tmB, p->M, p->N, p->K, p->ep));
  return MHMR_OK;
}

template <int EPI>
int launch_2cta(const GemmPlan* p, cudaStream_t stream) {
  if constexpr (EPI == EPI_LS_RESID_SPLIT) {
    if (p->K <= p->N) return launch_2cta_pre<EPI, true>(p, stream);   // attn.proj: K = N = D
  }
  return launch_2cta_pre<EPI, false>(p, stream);
}

}  // namespace

int gemm_plan_run_2cta(const GemmPlan* p, cudaStream_t stream) {
  switch (p->epi) {
    case EPI_BIAS_F16: return launch_2cta<EPI_BIAS_F16>(p, stream);
    case EPI_BIAS_GELU_F16: return launch_2cta<EPI_BIAS_GELU_F16>(p, stream);
    case EPI_BIAS_RELU_F16: return launch_2cta<EPI_BIAS_RELU_F16>(p, stream);
    case EPI_LS_RESID_F32: return launch_2cta<EPI_LS_RESID_F32>(p, stream);
    case EPI_ROWADD_F32: return launch_2cta<EPI_ROWADD_F32>(p, stream);
    case EPI_BIAS_F32: return launch_2cta<EPI_BIAS_F32>(p, stream);
    case EPI_LS_RESID_SPLIT: return launch_2cta<EPI_LS_RESID_SPLIT>(p, stream);
    case EPI_LN_BIAS_F16: return launch_2cta<EPI_LN_BIAS_F16>(p, stream);
    case EPI_LN_GELU_F16: return launch_2cta<EPI_LN_GELU_F16>(p, stream);
    default: break;
  }
  set_last_error("gemm: unknown epilogue kind");
  return MHMR_ERR_ARG;
}

}  // namespace mhmr

// Fused GEMM epilogue shared by the 1-CTA and the 2-CTA (cta_group::2) tcgen05 kernels.
#pragma once
#include "gemm_tc.cuh"

namespace mhmr {

constexpr int kScratchStride = 36;               // floats per scratch row (32 + 4: conflict-free float4)
constexpr int kScratchBytes = 32 * kScratchStride * 4;


// Per-(warp, tile) state of the fused LayerNorm (GemmEpi::ln_stats / stats_out).  Statically indexed only.
//   consumer (fp16 epilogues): v[k] = -mu, v[4 + k] = rstd of row m_base + k*8 + (lane >> 2)
//   producer (fp32 epilogues): v[k] = sum, v[8 + k] = sum of squares over this warp's columns of row
//                              m_base + k*4 + (lane >> 3), this lane's 4 columns per chunk
struct EpiTileCtx {
  float v[16];
};

template <int EPI>
__device__ __forceinline__ void epilogue_tile_begin(EpiTileCtx& ctx, const GemmEpi& ep, int M, int m_base, int lane) {
  constexpr bool kF16 = (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16 || EPI == EPI_BIAS_RELU_F16);
#pragma unroll
  for (int i = 0; i < 16; ++i) ctx.v[i] = 0.f;
  if constexpr (kF16) {
    if (ep.ln_stats == nullptr) return;
    // The slot partials are summed in a fixed order (slots cg, cg+4, ... then a 2-level butterfly), so the
    // statistics -- and with them the whole forward -- are bit-reproducible run to run.
    const int cg = lane & 3, rs = lane >> 2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int m = m_base + k * 8 + rs;
      float s1 = 0.f, s2 = 0.f;
      if (m < M) {
        const float2* p = reinterpret_cast<const float2*>(ep.ln_stats) + static_cast<int64_t>(m) * ep.ln_slots;
        for (int sl = cg; sl < ep.ln_slots; sl += 4) {
          const float2 t = __ldg(p + sl);
          s1 += t.x;
          s2 += t.y;
        }
      }
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
      const float mu = s1 * ep.ln_inv_dim;
      ctx.v[k] = -mu;
      ctx.v[4 + k] = rsqrtf(fmaxf(s2 * ep.ln_inv_dim - mu * mu, 0.f) + ep.ln_eps);
    }
  }
}

// Producer side: one plain store of (sum, sum of squares) per row into this warp's slot -- no atomics,
// nothing to zero beforehand.  slot = n_blk * (epilogue warps per TMEM quarter) + par.
template <int EPI>
__device__ __forceinline__ void epilogue_tile_end(EpiTileCtx& ctx, const GemmEpi& ep, int M, int m_base, int slot,
                                                  int lane) {
  if constexpr (EPI == EPI_LS_RESID_F32 || EPI == EPI_ROWADD_F32) {
    if (ep.stats_out == nullptr) return;
    const int cg = lane & 7, rs = lane >> 3;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float s1 = ctx.v[k], s2 = ctx.v[8 + k];
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      const int m = m_base + k * 4 + rs;
      if (cg == 0 && m < M) {
        int64_t orow = m;
        if constexpr (EPI == EPI_ROWADD_F32) {
          const int grp = m / ep.rows_in, rin = m - grp * ep.rows_in;
          orow = static_cast<int64_t>(grp) * ep.rows_out + ep.row_off + rin;
        }
        reinterpret_cast<float2*>(ep.stats_out)[orow * ep.stat_slots + slot] = make_float2(s1, s2);
      }
    }
  }
}

// The accumulator chunk (32 rows x 32 columns, one row per thread after tcgen05.ld) is transposed through
// a per-warp smem scratch so that global memory is accessed with lanes along the contiguous dimension:
//   fp32 outputs: 8 lanes x float4 cover one 128-byte row segment, 4 rows per warp instruction;
//   fp16 outputs: 4 lanes x (8 halves) cover one 64-byte row segment, 8 rows per warp instruction.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], float* scratch, const GemmEpi& ep,
                                               EpiTileCtx& ctx, int M, int N, int m_base, int n0, int lane) {
  constexpr bool kF16 = (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16 || EPI == EPI_BIAS_RELU_F16);
  // (a) rows -> scratch
  float* my = scratch + lane * kScratchStride;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(my + q * 4) =
        make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]),
                    __uint_as_float(r[q * 4 + 3]));
  __syncwarp();
  if constexpr (kF16) {
    const int cg = lane & 3, rs = lane >> 2;  // 8 columns per lane, 8 rows per instruction
    const int n = n0 + cg * 8;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + n + 4));
    const bool ln = (ep.ln_stats != nullptr);  // uniform: LayerNorm folded into this GEMM
    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
    if (ln) {
      c0 = __ldg(reinterpret_cast<const float4*>(ep.colsum + n));
      c1 = __ldg(reinterpret_cast<const float4*>(ep.colsum + n + 4));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int rl = k * 8 + rs;
      const int m = m_base + rl;
      const float4 v0 = *reinterpret_cast<const float4*>(scratch + rl * kScratchStride + cg * 8);
      const float4 v1 = *reinterpret_cast<const float4*>(scratch + rl * kScratchStride + cg * 8 + 4);
      float2 y[4];
      if (ln) {
        // y = rstd * (acc - mu * colsum) + bias'
        const float2 nmu = make_float2(ctx.v[k], ctx.v[k]), rs2 = make_float2(ctx.v[4 + k], ctx.v[4 + k]);
        y[0] = __ffma2_rn(rs2, __ffma2_rn(nmu, make_float2(c0.x, c0.y), make_float2(v0.x, v0.y)), make_float2(b0.x, b0.y));
        y[1] = __ffma2_rn(rs2, __ffma2_rn(nmu, make_float2(c0.z, c0.w), make_float2(v0.z, v0.w)), make_float2(b0.z, b0.w));
        y[2] = __ffma2_rn(rs2, __ffma2_rn(nmu, make_float2(c1.x, c1.y), make_float2(v1.x, v1.y)), make_float2(b1.x, b1.y));
        y[3] = __ffma2_rn(rs2, __ffma2_rn(nmu, make_float2(c1.z, c1.w), make_float2(v1.z, v1.w)), make_float2(b1.z, b1.w));
      } else {
        y[0] = __fadd2_rn(make_float2(v0.x, v0.y), make_float2(b0.x, b0.y));
        y[1] = __fadd2_rn(make_float2(v0.z, v0.w), make_float2(b0.z, b0.w));
        y[2] = __fadd2_rn(make_float2(v1.x, v1.y), make_float2(b1.x, b1.y));
        y[3] = __fadd2_rn(make_float2(v1.z, v1.w), make_float2(b1.z, b1.w));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if constexpr (EPI == EPI_BIAS_GELU_F16) y[i] = gelu_erf_fast2(y[i]);
        if constexpr (EPI == EPI_BIAS_RELU_F16) y[i] = make_float2(fmaxf(y[i].x, 0.0f), fmaxf(y[i].y, 0.0f));
      }
      const float x[8] = {y[0].x, y[0].y, y[1].x, y[1].y, y[2].x, y[2].y, y[3].x, y[3].y};
      const __half2 h0 = __floats2half2_rn(x[0], x[1]), h1 = __floats2half2_rn(x[2], x[3]);
      const __half2 h2 = __floats2half2_rn(x[4], x[5]), h3 = __floats2half2_rn(x[6], x[7]);
      uint4 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&h0);
      pk.y = *reinterpret_cast<const uint32_t*>(&h1);
      pk.z = *reinterpret_cast<const uint32_t*>(&h2);
      pk.w = *reinterpret_cast<const uint32_t*>(&h3);
      if (m < M)
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(ep.out) + static_cast<int64_t>(m) * ep.ldo + n) = pk;
    }
  } else {
    const int cg = lane & 7, rs = lane >> 3;  // 4 columns per lane, 4 rows per instruction
    const int n = n0 + cg * 4;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f), g = make_float4(1.f, 1.f, 1.f, 1.f);
    if constexpr (EPI == EPI_LS_RESID_F32) {
      b = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
      g = __ldg(reinterpret_cast<const float4*>(ep.gamma + n));
    } else if constexpr (EPI == EPI_BIAS_F32) {
      if (ep.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
    }
    float* outp[8];
    int64_t orows[8];
    float4 xres[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int m = m_base + k * 4 + rs;
      int64_t orow = m;
      if constexpr (EPI == EPI_ROWADD_F32) {
        const int grp = m / ep.rows_in, rin = m - grp * ep.rows_in;
        orow = static_cast<int64_t>(grp) * ep.rows_out + ep.row_off + rin;
        xres[k] = (m < M) ? __ldg(reinterpret_cast<const float4*>(ep.rowadd + static_cast<int64_t>(rin) * N + n))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      orows[k] = orow;
      outp[k] = reinterpret_cast<float*>(ep.out) + orow * ep.ldo + n;
      if constexpr (EPI == EPI_LS_RESID_F32)
        xres[k] = (m < M) ? *reinterpret_cast<const float4*>(outp[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int rl = k * 4 + rs;
      const float4 v = *reinterpret_cast<const float4*>(scratch + rl * kScratchStride + cg * 4);
      float4 a;
      if constexpr (EPI == EPI_LS_RESID_F32) {
        a.x = xres[k].x + g.x * (v.x + b.x);
        a.y = xres[k].y + g.y * (v.y + b.y);
        a.z = xres[k].z + g.z * (v.z + b.z);
        a.w = xres[k].w + g.w * (v.w + b.w);
      } else if constexpr (EPI == EPI_ROWADD_F32) {
        a.x = v.x + xres[k].x; a.y = v.y + xres[k].y; a.z = v.z + xres[k].z; a.w = v.w + xres[k].w;
      } else {
        a.x = v.x + b.x; a.y = v.y + b.y; a.z = v.z + b.z; a.w = v.w + b.w;
      }
      const bool row_ok = (m_base + rl) < M;
      if (row_ok) *reinterpret_cast<float4*>(outp[k]) = a;
      if constexpr (EPI == EPI_LS_RESID_F32 || EPI == EPI_ROWADD_F32) {
        if (ep.out16 != nullptr) {  // uniform: producer side of the fused LayerNorm
          const int64_t orow = orows[k];
          if (row_ok) {
            const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&h0);
            pk.y = *reinterpret_cast<const uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(ep.out16 + orow * ep.ld16 + n) = pk;
          }
          if (row_ok) {  // row partials, reduced across lanes once per tile (epilogue_tile_end)
            ctx.v[k] += (a.x + a.y) + (a.z + a.w);
            ctx.v[8 + k] += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
          }
        }
      }
    }
  }
  __syncwarp();
}

}  // namespace mhmr

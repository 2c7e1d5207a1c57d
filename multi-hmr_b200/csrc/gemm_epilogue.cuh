// Fused GEMM epilogue shared by the 1-CTA and the 2-CTA (cta_group::2) tcgen05 kernels.
#pragma once
#include "gemm_tc.cuh"

namespace mhmr {

constexpr int kScratchStride = 36;               // floats per scratch row (32 + 4: conflict-free float4)
constexpr int kScratchBytes = 32 * kScratchStride * 4;


// The accumulator chunk (32 rows x 32 columns, one row per thread after tcgen05.ld) is transposed through
// a per-warp smem scratch so that global memory is accessed with lanes along the contiguous dimension:
//   fp32 outputs: 8 lanes x float4 cover one 128-byte row segment, 4 rows per warp instruction;
//   fp16 outputs: 4 lanes x (8 halves) cover one 64-byte row segment, 8 rows per warp instruction.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], float* scratch, const GemmEpi& ep,
                                               int M, int N, int m_base, int n0, int lane) {
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
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int rl = k * 8 + rs;
      const int m = m_base + rl;
      const float4 v0 = *reinterpret_cast<const float4*>(scratch + rl * kScratchStride + cg * 8);
      const float4 v1 = *reinterpret_cast<const float4*>(scratch + rl * kScratchStride + cg * 8 + 4);
      float2 y[4] = {__fadd2_rn(make_float2(v0.x, v0.y), make_float2(b0.x, b0.y)),
                     __fadd2_rn(make_float2(v0.z, v0.w), make_float2(b0.z, b0.w)),
                     __fadd2_rn(make_float2(v1.x, v1.y), make_float2(b1.x, b1.y)),
                     __fadd2_rn(make_float2(v1.z, v1.w), make_float2(b1.z, b1.w))};
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
      if (m_base + rl < M) *reinterpret_cast<float4*>(outp[k]) = a;
    }
  }
  __syncwarp();
}

}  // namespace mhmr

// Micro-benchmark: MUFU.EX2 (and mixes) throughput per SM, to calibrate the softmax roofline.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i + 1) * 1e-3f - 1.0f;
  uint32_t sink = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[i]));
      if (MODE == 1) {  // + pack to half2 (F2FP) per pair
        if (i & 1) {
          __half2 h = __floats2half2_rn(a[i - 1], y);
          sink ^= *reinterpret_cast<uint32_t*>(&h);
        }
      }
      if (MODE == 2) y = fmaf(y, 0.999f, -1.0f);  // + 1 FFMA per exp
      a[i] = (MODE == 2) ? y : (y - 1.5f);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(sink & 1);
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
  int sms = 148, iters = 4096;
  float* out;
  cudaMalloc(&out, sms * 1024 * 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<sms, warps_per_sm * 32>>>(out, 16, 1.0f);
  cudaEventRecord(a);
  k<MODE><<<sms, warps_per_sm * 32>>>(out, iters, 1.0f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double exps = double(sms) * warps_per_sm * 32 * iters * 8;
  printf("%-28s warps/SM=%2d  %.3f ms  %.2f Texp/s  = %.2f exp/clk/SM at %.0f MHz (nominal max clock)\n", name,
         warps_per_sm, ms, exps / ms / 1e9, exps / (ms * 1e-3) / sms / (clk * 1e3), clk / 1e3);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16, 32}) run<0>("ex2 only", w);
  for (int w : {8, 16}) run<1>("ex2 + F2FP per pair", w);
  for (int w : {8, 16}) run<2>("ex2 + FFMA", w);
  return 0;
}

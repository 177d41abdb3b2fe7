// Microbenchmark: cost of the GroupNorm + SiLU element math with ONE warp per SM sub-partition (the situation of the conv
// kernel's transform warps) and with 4 warps per sub-partition.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe mufu_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int V>
__global__ void k(uint32_t* out, long long* cyc, int iters, float s0, float b0) {
  uint32_t w[4];
  for (int q = 0; q < 4; ++q) w[q] = 0x3c003800u + threadIdx.x * 17u + q * 1234567u;
  float hs[8], hb[8];
  for (int q = 0; q < 8; ++q) { hs[q] = s0 + q * 0.01f; hb[q] = b0 - q * 0.02f; }
  __half2 hs2[4], hb2[4];
  for (int q = 0; q < 4; ++q) { hs2[q] = __floats2half2_rn(hs[2*q], hs[2*q+1]); hb2[q] = __floats2half2_rn(hb[2*q], hb[2*q+1]); }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (V == 0) {  // fp32, tanh
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
        const float z0 = fmaf(f.x, hs[2*q], hb[2*q]), z1 = fmaf(f.y, hs[2*q+1], hb[2*q+1]);
        float t0_, t1_;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t0_) : "f"(z0));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t1_) : "f"(z1));
        __half2 hv = __floats2half2_rn(fmaf(z0, t0_, z0), fmaf(z1, t1_, z1));
        w[q] = *reinterpret_cast<uint32_t*>(&hv) ^ (uint32_t)it;
      } else if (V == 1) {  // fp32, ex2 + rcp
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
        const float z0 = fmaf(f.x, hs[2*q], hb[2*q]), z1 = fmaf(f.y, hs[2*q+1], hb[2*q+1]);
        const float e0 = exp2f(-z0 * 1.44f), e1 = exp2f(-z1 * 1.44f);
        __half2 hv = __floats2half2_rn(__fdividef(z0, 1.f + e0), __fdividef(z1, 1.f + e1));
        w[q] = *reinterpret_cast<uint32_t*>(&hv) ^ (uint32_t)it;
      } else if (V == 2) {  // half2, tanh.f16x2
        const __half2 x = *reinterpret_cast<const __half2*>(&w[q]);
        const __half2 z = __hfma2(x, hs2[q], hb2[q]);
        uint32_t zi = *reinterpret_cast<const uint32_t*>(&z), ti;
        asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(zi));
        const __half2 v = __hfma2(z, *reinterpret_cast<__half2*>(&ti), z);
        w[q] = *reinterpret_cast<const uint32_t*>(&v) ^ (uint32_t)it;
      } else if (V == 3) {  // MUFU only
        float a = __uint_as_float(w[q]), b;
        asm("tanh.approx.f32 %0, %1;" : "=f"(b) : "f"(a));
        float c;
        asm("tanh.approx.f32 %0, %1;" : "=f"(c) : "f"(b + 1.0f));
        w[q] = __float_as_uint(c) ^ (uint32_t)it;
      } else if (V == 4) {  // FMA only (same count as V0 without MUFU / cvt)
        float a = __uint_as_float(w[q]);
        a = fmaf(a, hs[2*q], hb[2*q]); a = fmaf(a, hs[2*q+1], hb[2*q+1]); a = fmaf(a, hs[2*q], hb[2*q+1]); a = fmaf(a, hs[2*q+1], hb[2*q]);
        w[q] = __float_as_uint(a) ^ (uint32_t)it;
      } else if (V == 5) {  // half2, ex2.f16x2 + rcp via half2
        const __half2 x = *reinterpret_cast<const __half2*>(&w[q]);
        const __half2 z = __hfma2(x, hs2[q], hb2[q]);
        const __half2 e = h2exp2(__hneg2(z));
        const __half2 v = __h2div(z, __hadd2(e, __float2half2_rn(1.f)));
        w[q] = *reinterpret_cast<const uint32_t*>(&v) ^ (uint32_t)it;
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = w[0] ^ w[1] ^ w[2] ^ w[3];
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}

template <int V>
void run(const char* name, int threads) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<V><<<148, threads>>>(out, cyc, 16, 0.5f, 0.1f);
  k<V><<<148, threads>>>(out, cyc, iters, 0.5f, 0.1f);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s threads/SM %4d: %7.1f cycles per 8-element row per warp, %6.2f elements/clk/SM  (%s)\n", name, threads, (double)h / iters,
         8.0 * threads * iters / (double)h, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int t : {128, 256, 512, 1024}) {
    run<0>("fp32 fma+tanh+fma", t);
    run<1>("fp32 ex2+rcp", t);
    run<2>("half2 hfma2+tanh.f16x2+hfma2", t);
    run<3>("2x tanh.f32 only", t);
    run<4>("4x fma only", t);
    run<5>("half2 h2exp2 + h2div", t);
  }
  return 0;
}

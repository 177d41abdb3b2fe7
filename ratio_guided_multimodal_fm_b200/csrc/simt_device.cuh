// Device helpers shared by the fp32 CUDA-core kernels (ops_simt.cu) and the training-side kernels (train.cu).
#pragma once
#include "common.h"
#include <cuda_fp16.h>

namespace rgfm {

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_SILU) return silu_f(x);
  if (act == ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}

// load `N` consecutive channel values of a raw activation stored as fp32 (fmt 0), bf16 (1) or fp16 (2)
__device__ __forceinline__ void load_raw4(const float* base, size_t off, int fmt, float* v) {
  if (fmt == 0) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(base + off));
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  } else {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + off));
    if (fmt == 1) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
  }
}
__device__ __forceinline__ void load_raw8(const float* base, size_t off, int fmt, float* v) {
  if (fmt == 0) {
    load_raw4(base, off, 0, v);
    load_raw4(base, off + 4, 0, v + 4);
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + off));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = fmt == 1 ? __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]))
                                : __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      v[2 * j] = a.x;
      v[2 * j + 1] = a.y;
    }
  }
}
__device__ __forceinline__ float load_raw1(const float* base, size_t off, int fmt) {
  if (fmt == 0) return base[off];
  const uint16_t u = reinterpret_cast<const uint16_t*>(base)[off];
  if (fmt == 1) return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&u));
  return __half2float(*reinterpret_cast<const __half*>(&u));
}


constexpr int PACK_PIX = 32;
constexpr int PACK_MAXC = 512;

// per-channel scale / shift of the normalisation of image n into shared memory (all threads of the CTA call this)
template <int BATCH = 4>
__device__ __forceinline__ void pack_scale_shift(const PackArgs& a, int n, float* s_scale, float* s_shift, float (*s_red)[PACK_MAXC],
                                                 double (*s_grp)[2]) {
  const int tid = threadIdx.x;
  const int C0 = a.src[0].C;
  if (a.norm == NORM_GN) {
    for (int c = tid; c < a.C; c += blockDim.x) {
      const int si = (c < C0) ? 0 : 1;
      float s, ss;
      stats_channel_sums<BATCH>(a.src[si].st, n, si ? c - C0 : c, s, ss);
      if (a.spatial == SP_D2S) {  // an output channel is four source channels (one per output-pixel parity)
        for (int q = 1; q < 4; ++q) {
          float s2, ss2;
          stats_channel_sums<BATCH>(a.src[0].st, n, c + q * a.C, s2, ss2);
          s += s2;
          ss += ss2;
        }
      }
      s_red[0][c] = s;
      s_red[1][c] = ss;
    }
    __syncthreads();
    const int cpg = a.C / a.groups;
    // one warp per group: lane-strided partial sums in double, butterfly reduction (a fixed order, so deterministic)
    for (int g = tid >> 5; g < a.groups; g += blockDim.x >> 5) {
      double s = 0.0, ss = 0.0;
      for (int c = g * cpg + (tid & 31); c < (g + 1) * cpg; c += 32) {
        s += (double)s_red[0][c];
        ss += (double)s_red[1][c];
      }
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
      }
      if ((tid & 31) == 0) {
        const double cnt = (double)cpg * (a.spatial == SP_D2S ? a.Ho * a.Wo : a.Hs * a.Ws);
        const double mean = s / cnt;
        double var = ss / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_grp[g][0] = mean;
        s_grp[g][1] = 1.0 / sqrt(var + 1e-5);
      }
    }
    __syncthreads();
    for (int c = tid; c < a.C; c += blockDim.x) {
      const int g = c / cpg;
      const float sc = (float)s_grp[g][1] * a.gamma[c];
      s_scale[c] = sc;
      s_shift[c] = a.beta[c] - (float)s_grp[g][0] * sc;
    }
  } else {
    for (int c = tid; c < a.C; c += blockDim.x) {
      s_scale[c] = a.norm == NORM_AFFINE ? a.gamma[c] : 1.0f;
      s_shift[c] = a.norm == NORM_AFFINE ? a.beta[c] : 0.0f;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float block_sum256(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < 8; ++w) r += red[w];
  return r;
}

}  // namespace rgfm

// fp32 CUDA-core kernels: operand packing (GroupNorm/BatchNorm-affine + SiLU + pool/upsample + concat), the
// exact-fp32 implicit-GEMM convolution used by the fp32 precision mode, and the small dense heads.
// The bf16 tensor-core convolution lives in conv_tc.cu and plugs into the same op program.
#include "common.h"
#include "simt_device.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace rgfm {

thread_local int64_t g_launch_counter = 0;
thread_local int g_tc_cta_cap = 0;

bool pdl_enabled() {
  // measured on B200 (round 1): 890 pairs/s with programmatic edges vs 896 without -- off unless RGFM_PDL=1
  static const bool on = getenv("RGFM_PDL") && atoi(getenv("RGFM_PDL")) != 0;
  return on;
}

static inline void post_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(std::string(what) + " launch failed: " + cudaGetErrorString(e));
  ++g_launch_counter;
}

// store 4 consecutive operand values as fp32 (FMT 0), bf16 (FMT 1) or fp16 (FMT 2)
template <int FMT>
__device__ __forceinline__ void store4(void* out, size_t idx, float a, float b, float c, float d) {
  if (FMT == 0) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + idx) = make_float4(a, b, c, d);
  } else if (FMT == 1) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + idx) = u;
  } else {
    const float L = 65504.f;  // saturate instead of overflowing to inf
    __half2 lo = __floats2half2_rn(fminf(fmaxf(a, -L), L), fminf(fmaxf(b, -L), L));
    __half2 hi = __floats2half2_rn(fminf(fmaxf(c, -L), L), fminf(fmaxf(d, -L), L));
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + idx) = u;
  }
}

// ------------------------------------------------------------------------------------------------
// NCHW fp32 image -> NHWC operand with channels zero-padded to Cp (input conv operand)
// ------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void pack_nchw_kernel(const float* __restrict__ x, int C, int HW, int Cp, void* __restrict__ out,
                                 size_t total_px) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total_px) return;
  size_t n = p / HW;
  int q = (int)(p - n * HW);
  const float* src = x + n * (size_t)C * HW + q;
  for (int c = 0; c < Cp; c += 4) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (c + j < C) ? __ldg(src + (size_t)(c + j) * HW) : 0.0f;
    store4<FMT>(out, p * Cp + c, v[0], v[1], v[2], v[3]);
  }
}

void launch_pack_nchw(const float* x, int C, int H, int W, int Cp, void* out, int out_bf16, int nimg,
                      cudaStream_t s) {
  size_t total = (size_t)nimg * H * W;
  int threads = 256;
  unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (out_bf16 == 1)
    pack_nchw_kernel<1><<<blocks, threads, 0, s>>>(x, C, H * W, Cp, out, total);
  else if (out_bf16 == 2)
    pack_nchw_kernel<2><<<blocks, threads, 0, s>>>(x, C, H * W, Cp, out, total);
  else
    pack_nchw_kernel<0><<<blocks, threads, 0, s>>>(x, C, H * W, Cp, out, total);
  post_launch("pack_nchw");
}

__device__ __forceinline__ float fast_act(float x, int act) {
  if (act == ACT_SILU) return __fdividef(x, 1.0f + __expf(-x));
  if (act == ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}

template <int FMT>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (FMT == 1) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  const float L = 65504.f;
  __half2 v = __floats2half2_rn(fminf(fmaxf(a, -L), L), fminf(fmaxf(b, -L), L));
  return *reinterpret_cast<uint32_t*>(&v);
}

// NCHW fp32 image -> padded-linear 16-bit operand for the tensor-core input conv: 16 channels (32 bytes) per slot holding
// [x_hi (C), x_lo (C), 0...] with x_hi = round16(x), x_lo = round16(x - x_hi); the conv reads it through a 64-channel TMA box
// whose upper 48 channels are out of bounds (zero).  The weights are duplicated over the two
// halves, so the state x_t enters the first conv with ~22 mantissa bits instead of 11 (8 for bf16) at no extra cost
// (the K block is 64 channels wide anyway).
template <int FMT>
__global__ void __launch_bounds__(256) pack_nchw_padded_kernel(const float* __restrict__ x, int C, int H, int W,
                                                               uint4* __restrict__ out) {
  const int n = blockIdx.y;
  const int P = W + 2;
  const int slots_img = (H + 1) * P;
  const int pp = blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0 && blockIdx.x == 0)
    for (int i = threadIdx.x; i < P * 2; i += blockDim.x) out[i] = make_uint4(0u, 0u, 0u, 0u);
  if (pp >= slots_img) return;
  const int ho = pp / P, ww = pp - ho * P;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (ho < H && ww >= 1 && ww <= W) {
    for (int c = 0; c < C; ++c) {
      const float xv = __ldg(x + ((size_t)(n * C + c) * H + ho) * W + (ww - 1));
      float hi;
      if (FMT == 1) hi = __bfloat162float(__float2bfloat16(xv));
      else hi = __half2float(__float2half_rn(xv));
      v[c] = hi;
      v[C + c] = xv - hi;
    }
  }
  uint4* dst = out + ((size_t)(1 + n * (H + 1)) * P + pp) * 2;
  dst[0] = make_uint4(pack2<FMT>(v[0], v[1]), pack2<FMT>(v[2], v[3]), pack2<FMT>(v[4], v[5]), pack2<FMT>(v[6], v[7]));
  dst[1] = make_uint4(0u, 0u, 0u, 0u);
}

void launch_pack_nchw_padded(const float* x, int C, int H, int W, void* out, int fmt, int nimg, cudaStream_t s) {
  RGFM_CHECK(2 * C <= 8 && (fmt == 1 || fmt == 2), "pack_nchw_padded: at most 4 image channels, 16-bit output");
  dim3 grid(((H + 1) * (W + 2) + 255) / 256, nimg);
  if (fmt == 1)
    pack_nchw_padded_kernel<1><<<grid, 256, 0, s>>>(x, C, H, W, reinterpret_cast<uint4*>(out));
  else
    pack_nchw_padded_kernel<2><<<grid, 256, 0, s>>>(x, C, H, W, reinterpret_cast<uint4*>(out));
  post_launch("pack_nchw_padded");
}

// ------------------------------------------------------------------------------------------------
// PACK: raw NHWC (1 or 2 concatenated sources) -> normalised + activated (+pooled / upsampled) conv operand.
// GroupNorm statistics come from the producers' {sum, sumsq} partials (per output tile for the SIMT conv, per
// 128-row unit for the tcgen05 conv); they are combined here in a fixed order (deterministic), group moments in
// double.  reference: F.group_norm at unet.py:83,91,274; BatchNorm(eval)/max_pool2d at ratio_flexible.py:217-226.
// ------------------------------------------------------------------------------------------------

template <int FMT>
__global__ void __launch_bounds__(256) pack_kernel(PackArgs a) {
  __shared__ float s_scale[PACK_MAXC];
  __shared__ float s_shift[PACK_MAXC];
  __shared__ float s_red[2][PACK_MAXC];
  __shared__ double s_grp[32][2];
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int C0 = a.src[0].C;
  pack_scale_shift(a, n, s_scale, s_shift, s_red, s_grp);

  const int C4 = a.C >> 2;
  const int npix = a.Ho * a.Wo;
  const int p0 = blockIdx.x * PACK_PIX;
  const bool plain = (a.norm == NORM_NONE && a.act == ACT_NONE);
  for (int i = tid; i < PACK_PIX * C4; i += blockDim.x) {
    const int p = p0 + i / C4;
    if (p >= npix) break;
    const int c = (i % C4) << 2;
    const int ho = p / a.Wo, wo = p - ho * a.Wo;
    const int si = (c < C0) ? 0 : 1;
    const int cl = si ? c - C0 : c;
    const int Cs = a.src[si].Cs;
    const int spad = a.src[si].padded;
    float4 r;
    if (a.spatial == SP_POOL2) {
      r = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          float lv[4];
          load_raw4(a.src[si].p, raw_offset(spad, n, 2 * ho + dy, 2 * wo + dx, a.Hs, a.Ws, Cs) + cl, a.src[si].fmt, lv);
          float4 v = make_float4(lv[0], lv[1], lv[2], lv[3]);
          v.x = apply_act(v.x * s_scale[c] + s_shift[c], a.act);
          v.y = apply_act(v.y * s_scale[c + 1] + s_shift[c + 1], a.act);
          v.z = apply_act(v.z * s_scale[c + 2] + s_shift[c + 2], a.act);
          v.w = apply_act(v.w * s_scale[c + 3] + s_shift[c + 3], a.act);
          r.x = fmaxf(r.x, v.x); r.y = fmaxf(r.y, v.y); r.z = fmaxf(r.z, v.z); r.w = fmaxf(r.w, v.w);
        }
    } else {
      const bool half_res = a.spatial == SP_UP2 || a.spatial == SP_D2S;
      const int hs = half_res ? (ho >> 1) : ho;
      const int ws = half_res ? (wo >> 1) : wo;
      const int csub = a.spatial == SP_D2S ? ((ho & 1) * 2 + (wo & 1)) * a.C : 0;  // parity sub-channel block
      float lv[4];
      load_raw4(a.src[si].p, raw_offset(spad, n, hs, ws, a.Hs, a.Ws, Cs) + csub + cl, a.src[si].fmt, lv);
      r = make_float4(lv[0], lv[1], lv[2], lv[3]);
      if (!plain) {
        r.x = apply_act(r.x * s_scale[c] + s_shift[c], a.act);
        r.y = apply_act(r.y * s_scale[c + 1] + s_shift[c + 1], a.act);
        r.z = apply_act(r.z * s_scale[c + 2] + s_shift[c + 2], a.act);
        r.w = apply_act(r.w * s_scale[c + 3] + s_shift[c + 3], a.act);
      }
    }
    const size_t o = ((size_t)n * npix + p) * a.C + c;
    store4<FMT>(a.out, o, r.x, r.y, r.z, r.w);
  }
}

// ------------------------------------------------------------------------------------------------
// PACK (padded-linear layout) for the tcgen05 convolution.  The operand is a 2-D tensor [slot][Cpad] of 16-bit
// values where slot = (1 + n*(H+1) + h) * P + (w + 1), P = W + 2: every image row carries one zero pixel on each
// side and every image is followed by one zero row (row 0 is zero too), so the nine taps of a 3x3 convolution are
// nine ROW SHIFTS (ky*P + kx) of one shared-memory tile and the conv's zero padding is physically in memory.
// This kernel writes the pad slots (zeros) together with the valid pixels, so arena blocks can be recycled.
//
// Streaming kernel: grid (slabs, images).  Each thread owns one 8-channel octet (its scale / shift live in
// registers) and walks pixel slots with stride 256 / (Cpad / 8); every load and store is 16 bytes and consecutive
// threads touch consecutive addresses.  GroupNorm statistics come from the producer's partials (no second pass
// over the image).
// ------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256) pack_padded_kernel(PackArgs a) {
  __shared__ float s_scale[PACK_MAXC];
  __shared__ float s_shift[PACK_MAXC];
  __shared__ float s_red[2][PACK_MAXC];
  __shared__ double s_grp[32][2];
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int C0 = a.src[0].C;
  pack_scale_shift(a, n, s_scale, s_shift, s_red, s_grp);

  const int P = a.Wo + 2;
  const int CP8 = a.Cpad >> 3;
  const int slots_img = (a.Ho + 1) * P;
  uint4* __restrict__ out = reinterpret_cast<uint4*>(a.out);
  if (n == 0 && blockIdx.x == 0)  // the leading zero row
    for (int i = tid; i < P * CP8; i += blockDim.x) out[i] = make_uint4(0u, 0u, 0u, 0u);
  const int PL = 256 / CP8;  // pixel lanes
  const int co = tid % CP8, pl = tid / CP8;
  if (pl >= PL) return;
  const int c = co << 3;
  const bool cvalid = c < a.C;
  const int si = (c < C0) ? 0 : 1;
  const int cl = si ? c - C0 : c;
  const int Cs = a.src[si].Cs;
  const int spad = a.src[si].padded, sfmt = a.src[si].fmt;
  const float* __restrict__ sp = a.src[si].p;
  const bool plain = (a.norm == NORM_NONE && a.act == ACT_NONE);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = cvalid ? s_scale[c + j] : 0.f;
    sh[j] = cvalid ? s_shift[c + j] : 0.f;
  }
  // this CTA's slab of pixel slots
  const int per = (slots_img + gridDim.x - 1) / gridDim.x;
  const int s_begin = blockIdx.x * per;
  const int s_end = min(slots_img, s_begin + per);
  const size_t slot0 = (size_t)(1 + n * (a.Ho + 1)) * P;
  int pp = s_begin + pl;
  int ho = pp / P, ww = pp - ho * P;
  const int dh = PL / P, dw = PL - dh * P;
  for (; pp < s_end; pp += PL) {
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (cvalid && ho < a.Ho && ww >= 1 && ww <= a.Wo) {
      const int wo = ww - 1;
      float v[8];
      if (a.spatial == SP_POOL2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = -INFINITY;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            float xv[8];
            load_raw8(sp, raw_offset(spad, n, 2 * ho + dy, 2 * wo + dx, a.Hs, a.Ws, Cs) + cl, sfmt, xv);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], fast_act(fmaf(xv[j], sc[j], sh[j]), a.act));
          }
      } else {
        const int hs = (a.spatial == SP_UP2) ? (ho >> 1) : ho;
        const int ws = (a.spatial == SP_UP2) ? (wo >> 1) : wo;
        float xv[8];
        load_raw8(sp, raw_offset(spad, n, hs, ws, a.Hs, a.Ws, Cs) + cl, sfmt, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = plain ? xv[j] : fast_act(fmaf(xv[j], sc[j], sh[j]), a.act);
      }
      o.x = pack2<FMT>(v[0], v[1]);
      o.y = pack2<FMT>(v[2], v[3]);
      o.z = pack2<FMT>(v[4], v[5]);
      o.w = pack2<FMT>(v[6], v[7]);
    }
    out[(slot0 + pp) * CP8 + co] = o;
    ho += dh;
    ww += dw;
    if (ww >= P) {
      ww -= P;
      ++ho;
    }
  }
}

// Fast path of the padded-linear pack: 16-bit sources, no pooling.  Same mapping as pack_padded_kernel (one 8-channel
// octet per thread, scale / shift in registers), but every thread keeps FOUR independent 16-byte loads in flight per
// iteration: the kernel is a pure stream and one dependent load per thread leaves it latency-bound at ~2 TB/s.
template <int FMT>
__global__ void __launch_bounds__(256) pack_padded_fast_kernel(PackArgs a) {
  pdl_trigger();
  pdl_wait();  // the statistics and the raw tensor come from the previous kernel
  __shared__ float s_scale[PACK_MAXC];
  __shared__ float s_shift[PACK_MAXC];
  __shared__ float s_red[2][PACK_MAXC];
  __shared__ double s_grp[32][2];
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int C0 = a.src[0].C;
  pack_scale_shift(a, n, s_scale, s_shift, s_red, s_grp);

  const int P = a.Wo + 2;
  const int CP8 = a.Cpad >> 3;
  const int slots_img = (a.Ho + 1) * P;
  uint4* __restrict__ out = reinterpret_cast<uint4*>(a.out);
  if (n == 0 && blockIdx.x == 0)  // the leading zero row
    for (int i = tid; i < P * CP8; i += blockDim.x) out[i] = make_uint4(0u, 0u, 0u, 0u);
  const int PL = 256 / CP8;  // pixel lanes
  const int co = tid % CP8, pl = tid / CP8;
  if (pl >= PL) return;
  const int c = co << 3;
  const bool cvalid = c < a.C;
  const int si = (c < C0) ? 0 : 1;
  const int cl = si ? c - C0 : c;
  const int Cs = a.src[si].Cs;
  const int sfmt = a.src[si].fmt;  // 1 = bf16, 2 = fp16
  const uint16_t* __restrict__ sp = reinterpret_cast<const uint16_t*>(a.src[si].p);
  const bool plain = (a.norm == NORM_NONE && a.act == ACT_NONE);
  const bool up2 = a.spatial == SP_UP2;
  const int act = a.act;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = cvalid ? s_scale[c + j] : 0.f;
    sh[j] = cvalid ? s_shift[c + j] : 0.f;
  }
  const int per = (slots_img + gridDim.x - 1) / gridDim.x;
  const int s_begin = blockIdx.x * per;
  const int s_end = min(slots_img, s_begin + per);
  const size_t slot0 = (size_t)(1 + n * (a.Ho + 1)) * P;
  const size_t src_img = (size_t)n * (a.Hs + 1) * (a.Ws + 2);  // first padded-linear row of the source image
  const int Ps = a.Ws + 2;
  int pp = s_begin + pl;
  int ho = pp / P, ww = pp - ho * P;
  const int dh = PL / P, dw = PL - dh * P;
  for (; pp < s_end; pp += 4 * PL) {
    uint4 rawv[4];
    bool ok[4];
    int pps[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      pps[u] = pp + u * PL;
      ok[u] = cvalid && pps[u] < s_end && ho < a.Ho && ww >= 1 && ww <= a.Wo;
      rawv[u] = make_uint4(0u, 0u, 0u, 0u);
      if (ok[u]) {
        const int wo = ww - 1;
        const int hs = up2 ? (ho >> 1) : ho, ws = up2 ? (wo >> 1) : wo;
        rawv[u] = __ldg(reinterpret_cast<const uint4*>(sp + (src_img + (size_t)hs * Ps + ws) * Cs + cl));
      }
      ho += dh;
      ww += dw;
      if (ww >= P) {
        ww -= P;
        ++ho;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (pps[u] >= s_end) break;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (ok[u]) {
        if (plain && sfmt == FMT) {
          o = rawv[u];
        } else {
          const uint32_t w4[4] = {rawv[u].x, rawv[u].y, rawv[u].z, rawv[u].w};
          float v[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = sfmt == 1 ? __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[j]))
                                       : __half22float2(*reinterpret_cast<const __half2*>(&w4[j]));
            v[2 * j] = plain ? f.x : fast_act(fmaf(f.x, sc[2 * j], sh[2 * j]), act);
            v[2 * j + 1] = plain ? f.y : fast_act(fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]), act);
          }
          o.x = pack2<FMT>(v[0], v[1]);
          o.y = pack2<FMT>(v[2], v[3]);
          o.z = pack2<FMT>(v[4], v[5]);
          o.w = pack2<FMT>(v[6], v[7]);
        }
      }
      out[(slot0 + pps[u]) * CP8 + co] = o;
    }
  }
}

// The hot pack of the U-Nets: GroupNorm + SiLU of fp16 padded-linear raw activations (1 or 2 concatenated sources) -> 16-bit
// padded-linear operand.  The generic kernels above are issue-bound (~22 instructions per element: run-time mode checks,
// saturating converts, a 5-instruction sigmoid around the normalisation FMA); this one is specialised down to ~9:
//   z = fma(x, sc, sh);  e = ex2(fma(x, -sc*log2e, -sh*log2e));  silu = z * rcp(1 + e)
// with both coefficient pairs per channel in registers, four 16-byte loads in flight per thread and no saturation (a
// GroupNorm output cannot leave the fp16 range).
template <int FMT, int U, int STREAM>
__global__ void __launch_bounds__(256) pack_gn_silu_kernel(PackArgs a) {
  pdl_trigger();
  pdl_wait();  // the statistics and the raw tensor come from the previous kernel
  __shared__ float s_scale[PACK_MAXC];
  __shared__ float s_shift[PACK_MAXC];
  __shared__ float s_red[2][PACK_MAXC];
  __shared__ double s_grp[32][2];
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int C0 = a.src[0].C;
  pack_scale_shift(a, n, s_scale, s_shift, s_red, s_grp);

  const int P = a.Wo + 2;
  const int CP8 = a.Cpad >> 3;
  const int slots_img = (a.Ho + 1) * P;
  uint4* __restrict__ out = reinterpret_cast<uint4*>(a.out);
  if (n == 0 && blockIdx.x == 0)  // the leading zero row
    for (int i = tid; i < P * CP8; i += blockDim.x) out[i] = make_uint4(0u, 0u, 0u, 0u);
  const int PL = 256 / CP8;  // pixel lanes
  const int co = tid % CP8, pl = tid / CP8;
  if (pl >= PL) return;
  const int c = co << 3;
  const bool cvalid = c < a.C;
  const int si = (c < C0) ? 0 : 1;
  const int cl = si ? c - C0 : c;
  const int Cs = a.src[si].Cs;
  const uint16_t* __restrict__ sp = reinterpret_cast<const uint16_t*>(a.src[si].p);
  constexpr float LOG2E = 1.4426950408889634f;
  float sc[8], sh[8], nsc[8], nsh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = cvalid ? s_scale[c + j] : 0.f;
    sh[j] = cvalid ? s_shift[c + j] : 0.f;
    nsc[j] = -sc[j] * LOG2E;
    nsh[j] = -sh[j] * LOG2E;
  }
  const int per = (slots_img + gridDim.x - 1) / gridDim.x;
  const int s_begin = blockIdx.x * per;
  const int s_end = min(slots_img, s_begin + per);
  const size_t slot0 = (size_t)(1 + n * (a.Ho + 1)) * P;
  // source and operand share the geometry: operand slot pp = (h, w + 1) <-> raw row pp - 1
  const uint16_t* __restrict__ src_img = sp + ((size_t)n * slots_img) * Cs + cl;
  int pp = s_begin + pl;
  int ho = pp / P, ww = pp - ho * P;
  const int dh = PL / P, dw = PL - dh * P;
  for (; pp < s_end; pp += U * PL) {
    uint4 rawv[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int q = pp + u * PL;
      ok[u] = cvalid && q < s_end && ho < a.Ho && ww >= 1 && ww <= a.Wo;
      rawv[u] = make_uint4(0u, 0u, 0u, 0u);
      if (ok[u]) rawv[u] = STREAM ? __ldcs(reinterpret_cast<const uint4*>(src_img + (size_t)(q - 1) * Cs)) : __ldg(reinterpret_cast<const uint4*>(src_img + (size_t)(q - 1) * Cs));
      ho += dh;
      ww += dw;
      if (ww >= P) {
        ww -= P;
        ++ho;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int q = pp + u * PL;
      if (q >= s_end) break;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (ok[u]) {
        const uint32_t w4[4] = {rawv[u].x, rawv[u].y, rawv[u].z, rawv[u].w};
        uint32_t r4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[j]));
          const float z0 = fmaf(f.x, sc[2 * j], sh[2 * j]), z1 = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
          const float e0 = exp2f(fmaf(f.x, nsc[2 * j], nsh[2 * j])), e1 = exp2f(fmaf(f.y, nsc[2 * j + 1], nsh[2 * j + 1]));
          const float v0 = __fdividef(z0, 1.0f + e0), v1 = __fdividef(z1, 1.0f + e1);
          if (FMT == 1) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
            r4[j] = *reinterpret_cast<uint32_t*>(&h);
          } else {
            __half2 h = __floats2half2_rn(v0, v1);
            r4[j] = *reinterpret_cast<uint32_t*>(&h);
          }
        }
        o = make_uint4(r4[0], r4[1], r4[2], r4[3]);
      }
      if (STREAM) __stcs(&out[(slot0 + q) * CP8 + co], o);
      else out[(slot0 + q) * CP8 + co] = o;
    }
  }
}

void launch_pack_padded(const PackArgs& a, cudaStream_t s) {
  RGFM_CHECK(a.C <= PACK_MAXC && a.C % 8 == 0 && a.Cpad % 8 == 0 && a.Cpad >= a.C && a.Cpad <= 2048, "pack_padded: channel counts");
  RGFM_CHECK(a.src[0].C % 8 == 0, "pack_padded: first source channels must be a multiple of 8");
  RGFM_CHECK(a.out_bf16 == 1 || a.out_bf16 == 2, "pack_padded: 16-bit output only");
  // one CTA per image when the batch alone fills the machine, otherwise split images into slabs
  const int slots_img = (a.Ho + 1) * (a.Wo + 2);
  static const int slab_mul = getenv("RGFM_PACK_SLABS") ? atoi(getenv("RGFM_PACK_SLABS")) : 4;
  int slabs = (slab_mul * 148 + a.nimg - 1) / a.nimg;
  const int max_slabs = (slots_img * (a.Cpad / 8) + 2047) / 2048;  // >= 8 iterations of 256 threads per CTA
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  dim3 grid(slabs, a.nimg);
  const bool fast = a.spatial != SP_POOL2 && a.src[0].fmt != 0 && a.src[0].padded && (a.nsrc == 1 || (a.src[1].fmt != 0 && a.src[1].padded));
  const bool gn_silu = fast && a.norm == NORM_GN && a.act == ACT_SILU && a.spatial == SP_ID && a.src[0].fmt == 2 &&
                       (a.nsrc == 1 || a.src[1].fmt == 2);
  if (gn_silu) {
    // measured slower on B200 and removed: 8 loads in flight, evict-first hints, software-pipelined batches, more slabs
    if (a.out_bf16 == 1)
      launch_pdl(pack_gn_silu_kernel<1, 4, 0>, grid, dim3(256), 0, s, a);
    else
      launch_pdl(pack_gn_silu_kernel<2, 4, 0>, grid, dim3(256), 0, s, a);
  } else if (fast) {
    if (a.out_bf16 == 1)
      launch_pdl(pack_padded_fast_kernel<1>, grid, dim3(256), 0, s, a);
    else
      launch_pdl(pack_padded_fast_kernel<2>, grid, dim3(256), 0, s, a);
  } else if (a.out_bf16 == 1) {
    pack_padded_kernel<1><<<grid, 256, 0, s>>>(a);
  } else {
    pack_padded_kernel<2><<<grid, 256, 0, s>>>(a);
  }
  post_launch("pack_padded");
}

// GroupNorm coefficients for the conv kernel's fused operand path: per (image, channel) {scale, shift} with
// operand = silu(raw * scale + shift), from the producers' partial statistics (same reduction as the pack kernels).
// Per-image GroupNorm {scale, shift} for the conv kernel's fused GroupNorm + SiLU segments: the prologue of the pack
// kernels on its own (one block per image), written once per layer instead of being re-derived by every consumer CTA.
// coef_h (optional): the same coefficients halved and rounded to fp16, laid out per (image, 8-channel octet) as 8 x hs then
// 8 x hb (32 bytes): silu(z) = hz + hz * tanh(hz) with hz = x * hs + hb is what the conv's transform warps evaluate in
// packed half2 arithmetic (conv_tc.cu), two LDG.128 per octet instead of four and no conversions in the hot loop
__device__ __forceinline__ void store_coef_h(__half* coef_h, int n, int C, int c, float sc, float sh) {
  __half* q = coef_h + ((size_t)n * (C >> 3) + (c >> 3)) * 16 + (c & 7);
  q[0] = __float2half_rn(0.5f * sc);
  q[8] = __float2half_rn(0.5f * sh);
}
__global__ void __launch_bounds__(256) gn_coef_kernel(PackArgs a, float2* __restrict__ coef, __half* __restrict__ coef_h) {
  __shared__ float s_scale[PACK_MAXC];
  __shared__ float s_shift[PACK_MAXC];
  __shared__ float s_red[2][PACK_MAXC];
  __shared__ double s_grp[32][2];
  const int n = blockIdx.x;
  pack_scale_shift<12>(a, n, s_scale, s_shift, s_red, s_grp);
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    coef[(size_t)n * a.C + c] = make_float2(s_scale[c], s_shift[c]);
    if (coef_h) store_coef_h(coef_h, n, a.C, c, s_scale[c], s_shift[c]);
  }
}
// C <= 256: one thread per channel, one barrier.  gamma / beta and every statistics partial are requested before anything
// is consumed (one global round trip), each thread sums its own group from shared memory in double, and the count enters
// as a host-computed reciprocal (no fp64 division or square root on the critical path of a ~3 us kernel that sits between
// two convolutions 60 times per Euler step).
__global__ void __launch_bounds__(256) gn_coef256_kernel(PackArgs a, float2* __restrict__ coef, __half* __restrict__ coef_h, double inv_cnt, int ipb) {
  __shared__ float s_red[2][256];
  pdl_trigger();  // the consuming conv may set up (barriers, TMEM, tensor maps) while this grid runs
  pdl_wait();     // the statistics come from the previous kernel
  // ipb = 256 / C images per block, thread -> (image, channel): every thread of the block has work
  const int il = threadIdx.x / a.C, c = threadIdx.x - il * a.C;
  const int n = blockIdx.x * ipb + il;
  const bool on = il < ipb && n < a.nimg;
  const int C0 = a.src[0].C;
  const int cpg = a.C / a.groups;
  float gam = 0.f, bet = 0.f, s = 0.f, ss = 0.f;
  if (on) {
    gam = __ldg(a.gamma + c);
    bet = __ldg(a.beta + c);
    const int si = (c < C0) ? 0 : 1;
    stats_channel_sums<12>(a.src[si].st, n, si ? c - C0 : c, s, ss);
  }
  s_red[0][threadIdx.x] = s;
  s_red[1][threadIdx.x] = ss;
  __syncthreads();
  if (!on) return;
  const int g0 = (int)threadIdx.x - c + (c / cpg) * cpg;
  double gs = 0.0, gss = 0.0;
  for (int k = 0; k < cpg; ++k) {  // ascending channel order, like pack_scale_shift
    gs += (double)s_red[0][g0 + k];
    gss += (double)s_red[1][g0 + k];
  }
  const double mean = gs * inv_cnt;
  double var = gss * inv_cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  const float sc = rsqrtf((float)var + 1e-5f) * gam;
  const float sh = bet - (float)mean * sc;
  coef[(size_t)n * a.C + c] = make_float2(sc, sh);
  if (coef_h) store_coef_h(coef_h, n, a.C, c, sc, sh);
}
void launch_gn_coef(const PackArgs& a, float* coef, void* coef_h, cudaStream_t s) {
  RGFM_CHECK(a.C <= PACK_MAXC && a.norm == NORM_GN && a.C % a.groups == 0, "gn_coef: GroupNorm over at most 512 channels");
#ifdef RGFM_EXPERIMENTS  // timing-only switch that knowingly corrupts results: compiled out of the shipped library
  static const bool skip = getenv("RGFM_EXPERIMENT_SKIP_GN_COEF") != nullptr;
  if (skip) return;
#endif
  if (a.C <= 256) {
    const int ipb = 256 / a.C;
    launch_pdl(gn_coef256_kernel, dim3((a.nimg + ipb - 1) / ipb), dim3(256), 0, s, a, reinterpret_cast<float2*>(coef),
               reinterpret_cast<__half*>(coef_h), 1.0 / ((double)(a.C / a.groups) * a.Hs * a.Ws), ipb);
  }
  else
    gn_coef_kernel<<<a.nimg, 256, 0, s>>>(a, reinterpret_cast<float2*>(coef), reinterpret_cast<__half*>(coef_h));
  post_launch("gn_coef");
}

void launch_pack(const PackArgs& a, cudaStream_t s) {
  RGFM_CHECK(a.C <= PACK_MAXC && a.C % 4 == 0, "pack: channel count");
  RGFM_CHECK(a.groups <= 32, "pack: groups");
  dim3 grid((a.Ho * a.Wo + PACK_PIX - 1) / PACK_PIX, a.nimg);
  if (a.out_bf16 == 1)
    pack_kernel<1><<<grid, 256, 0, s>>>(a);
  else if (a.out_bf16 == 2)
    pack_kernel<2><<<grid, 256, 0, s>>>(a);
  else
    pack_kernel<0><<<grid, 256, 0, s>>>(a);
  post_launch("pack");
}

// ------------------------------------------------------------------------------------------------
// fp32 implicit-GEMM convolution (3x3 / 1x1, stride 1 / 2) on CUDA cores: the exact-fp32 precision mode.
// M = 128 output pixels (TileGeom), N = TN output channels, K = taps x Cin in chunks of 16.
// Epilogue: + bias + per-step time bias + residual, raw NHWC store (or NCHW velocity / fused Euler update),
// per-tile GroupNorm partial sums.  reference ops: nn.Conv2d at unet.py:55,66,70,104,115,164,210.
// ------------------------------------------------------------------------------------------------
template <int TN>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs a) {
  constexpr int KC = 16;
  constexpr int CPT = TN / 16;
  __shared__ __align__(16) float As[KC][TM + 4];
  __shared__ __align__(16) float Bs[KC][TN];
  __shared__ float Ts[TM][TN + 1];

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int n0 = blockIdx.y * TN;
  const int pg = tid >> 4, cg = tid & 15;
  const float* in = reinterpret_cast<const float*>(a.in);

  // A-operand load assignment: two (pixel, channel-quad) items per thread
  int ln[2], lh[2], lw[2];
  bool lv[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int idx = tid + j * 256;
    int n, ho, wo, part;
    lv[j] = tile_row(a.g, tile, idx >> 2, a.nimg, n, ho, wo, part);
    ln[j] = n;
    lh[j] = ho * a.stride - a.pad;
    lw[j] = wo * a.stride - a.pad;
  }

  float acc[8][CPT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[i][j] = 0.0f;

  // K loop over (tap, 16-channel chunk) with a register prefetch: the global loads of step it + 1 are in flight while
  // step it is multiplied out of shared memory (same summation order as a plain loop, so results do not change)
  const int ntaps = a.ks * a.ks;
  const int nchunk = a.Cin / KC, nit = ntaps * nchunk;
  auto load_g = [&](int it, float4 (&va)[2], float4& vb) {
    const int tap = it / nchunk, c0 = (it - tap * nchunk) * KC;
    const int ky = tap / a.ks, kx = tap - ky * a.ks;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c4 = (tid + j * 256) & 3;
      const int hi = lh[j] + ky, wi = lw[j] + kx;
      va[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lv[j] && hi >= 0 && hi < a.Hin && wi >= 0 && wi < a.Win)
        va[j] = __ldg(reinterpret_cast<const float4*>(in + (((size_t)ln[j] * a.Hin + hi) * a.Win + wi) * a.Cin + c0 + c4 * 4));
    }
    if (tid < KC * TN / 4) {
      const int k = tid / (TN / 4), c4b = tid % (TN / 4);
      vb = __ldg(reinterpret_cast<const float4*>(a.w + ((size_t)tap * a.Cin + c0 + k) * a.CoutP + n0 + c4b * 4));
    }
  };
  auto store_s = [&](const float4 (&va)[2], const float4& vb) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int idx = tid + j * 256;
      const int px = idx >> 2, c4 = idx & 3;
      As[c4 * 4 + 0][px] = va[j].x;
      As[c4 * 4 + 1][px] = va[j].y;
      As[c4 * 4 + 2][px] = va[j].z;
      As[c4 * 4 + 3][px] = va[j].w;
    }
    if (tid < KC * TN / 4) {
      const int k = tid / (TN / 4), c4b = tid % (TN / 4);
      *reinterpret_cast<float4*>(&Bs[k][c4b * 4]) = vb;
    }
  };
  float4 pa[2], pb = make_float4(0.f, 0.f, 0.f, 0.f);
  load_g(0, pa, pb);
  store_s(pa, pb);
  __syncthreads();
  for (int it = 0; it < nit; ++it) {
    const bool more = it + 1 < nit;
    if (more) load_g(it + 1, pa, pb);
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][pg * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][pg * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) bv[j] = Bs[k][cg * CPT + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
    if (more) {
      store_s(pa, pb);
      __syncthreads();
    }
  }

  // epilogue
  const float* tb = a.temb ? a.temb + (size_t)(a.step ? *a.step : 0) * a.temb_stride + a.temb_off : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int px = pg * 8 + i;
    int n, ho, wo, part;
    const bool valid = tile_row(a.g, tile, px, a.nimg, n, ho, wo, part);
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int co = n0 + cg * CPT + j;
      float val = 0.0f;
      if (valid && co < a.Cout) {
        val = acc[i][j] + a.bias[co];
        if (tb) val += tb[co];
        const size_t o = (((size_t)n * a.g.H + ho) * a.g.W + wo) * a.Cout + co;
        if (a.res) val += a.res[o];
        if (a.out_mode == OUT_RAW) {
          const size_t oo = a.out_padded ? raw_offset(1, n, ho, wo, a.g.H, a.g.W, a.Cout) + co : o;
          if (a.out_fmt == 0) a.out[oo] = val;
          else if (a.out_fmt == 1) reinterpret_cast<__nv_bfloat16*>(a.out)[oo] = __float2bfloat16(val);
          else reinterpret_cast<__half*>(a.out)[oo] = __float2half_rn(fminf(fmaxf(val, -65504.f), 65504.f));
        } else {
          const size_t oc = (((size_t)n * a.Cout + co) * a.g.H + ho) * a.g.W + wo;
          if (a.out_mode == OUT_NCHW)
            a.out[oc] = val;
          else
            a.out[oc] = __fadd_rn(a.out[oc], __fmul_rn(val, a.dt));
        }
      }
      Ts[px][cg * CPT + j] = val;
    }
  }
  if (a.stats == nullptr) return;
  __syncthreads();
  const int hw = a.g.H * a.g.W;
  for (int item = tid; item < a.g.imgs * TN; item += 256) {
    const int il = item / TN, c = item - il * TN;
    const int r0 = (a.g.imgs == 1) ? 0 : il * hw;
    const int r1 = (a.g.imgs == 1) ? a.g.th * a.g.W : r0 + hw;
    float s = 0.f, ss = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float v = Ts[r][c];
      s += v;
      ss = fmaf(v, v, ss);
    }
    const int n = (a.g.imgs == 1) ? tile / a.g.parts : tile * a.g.imgs + il;
    const int part = (a.g.imgs == 1) ? tile % a.g.parts : 0;
    const int co = n0 + c;
    if (n < a.nimg && co < a.Cout) {
      float* st = a.stats + (((size_t)n * a.g.parts + part) * a.Cout + co) * 2;
      st[0] = s;
      st[1] = ss;
    }
  }
}

void launch_conv_simt(const ConvArgs& a, cudaStream_t s) {
  RGFM_CHECK(a.Cin % 16 == 0, "conv_simt: Cin must be a multiple of 16");
  RGFM_CHECK(a.CoutP % 32 == 0, "conv_simt: padded Cout must be a multiple of 32");
  const int tiles = a.g.tiles(a.nimg);
  // 64-channel tiles unless they would leave SMs idle (deep layers at 4x4 / 2x2 with small batches): the 32-channel
  // variant doubles the CTAs; every output's summation order is the same, so results do not depend on the choice
  if (a.CoutP % 64 == 0 && tiles * (a.CoutP / 64) >= 148) {
    dim3 grid(tiles, a.CoutP / 64);
    conv_simt_kernel<64><<<grid, 256, 0, s>>>(a);
  } else {
    dim3 grid(tiles, a.CoutP / 32);
    conv_simt_kernel<32><<<grid, 256, 0, s>>>(a);
  }
  post_launch("conv_simt");
}

// ------------------------------------------------------------------------------------------------
// Global average pool + Linear: one CTA per image.  reference: ratio_estimator.py:86-91, ratio_flexible.py:228-230
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gap_fc_kernel(const float* __restrict__ in, int HW, int C,
                                                     const float* __restrict__ w, const float* __restrict__ b,
                                                     int Cout, float* __restrict__ out, int ld) {
  __shared__ float m[PACK_MAXC];
  const int n = blockIdx.x;
  const float* src = in + (size_t)n * HW * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += src[(size_t)p * C + c];
    m[c] = s / (float)HW;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < Cout; j += 8) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(w[(size_t)j * C + c], m[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[(size_t)n * ld + j] = s + b[j];
  }
}

void launch_gap_fc(const float* in, int HW, int C, const float* w, const float* b, int Cout, float* out, int ld,
                   int nimg, cudaStream_t s) {
  RGFM_CHECK(C <= PACK_MAXC, "gap_fc: C");
  gap_fc_kernel<<<nimg, 256, 0, s>>>(in, HW, C, w, b, Cout, out, ld);
  post_launch("gap_fc");
}

// ------------------------------------------------------------------------------------------------
// Dense head: up to 4 x (Linear [-> LayerNorm -> SiLU] | [-> ReLU]); one CTA per sample.
// reference: score_net at ratio_estimator.py:125-135, ratio_flexible.py:332-345
// ------------------------------------------------------------------------------------------------
struct MlpPack {
  MlpLayer l[4];
  int n;
};
constexpr int MLP_MAXW = 8192;

__global__ void __launch_bounds__(256) mlp_kernel(const float* __restrict__ in, int ld_in, MlpPack mp,
                                                  float* __restrict__ out, int ld_out) {
  extern __shared__ float sm[];  // two ping-pong activation vectors
  __shared__ float red[2];
  const int n = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* cur = sm;
  float* nxt = sm + MLP_MAXW;
  for (int i = tid; i < mp.l[0].in; i += 256) cur[i] = in[(size_t)n * ld_in + i];
  __syncthreads();
  for (int li = 0; li < mp.n; ++li) {
    const MlpLayer L = mp.l[li];
    for (int j = warp; j < L.out; j += 8) {
      const float* wr = L.w + (size_t)j * L.in;
      float s = 0.f;
      for (int c = lane; c < L.in; c += 32) s = fmaf(wr[c], cur[c], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) nxt[j] = s + L.b[j];
    }
    __syncthreads();
    if (L.ln_g) {
      if (warp == 0) {
        float s = 0.f;
        for (int c = lane; c < L.out; c += 32) s += nxt[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)L.out;
        float v = 0.f;
        for (int c = lane; c < L.out; c += 32) {
          const float d = nxt[c] - mean;
          v = fmaf(d, d, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
          red[0] = mean;
          red[1] = rsqrtf(v / (float)L.out + 1e-5f);
        }
      }
      __syncthreads();
      for (int c = tid; c < L.out; c += 256) nxt[c] = (nxt[c] - red[0]) * red[1] * L.ln_g[c] + L.ln_b[c];
      __syncthreads();
    }
    if (L.act != ACT_NONE) {
      for (int c = tid; c < L.out; c += 256) nxt[c] = apply_act(nxt[c], L.act);
      __syncthreads();
    }
    float* t = cur;
    cur = nxt;
    nxt = t;
  }
  const int nout = mp.l[mp.n - 1].out;
  for (int c = tid; c < nout; c += 256) out[(size_t)n * ld_out + c] = cur[c];
}

void launch_mlp(const float* in, int ld_in, const MlpLayer* layers, int nlayers, float* out, int ld_out, int nimg,
                cudaStream_t s) {
  RGFM_CHECK(nlayers >= 1 && nlayers <= 4, "mlp: layer count");
  MlpPack mp;
  mp.n = nlayers;
  for (int i = 0; i < nlayers; ++i) {
    mp.l[i] = layers[i];
    RGFM_CHECK(layers[i].in <= MLP_MAXW && layers[i].out <= MLP_MAXW, "mlp: width");
  }
  const size_t smem = 2 * MLP_MAXW * sizeof(float);
  static thread_local int attr_dev = -1;  // function attributes are per device
  int dev = 0;
  RGFM_CUDA(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    RGFM_CUDA(cudaFuncSetAttribute(mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev = dev;
  }
  mlp_kernel<<<nimg, 256, smem, s>>>(in, ld_in, mp, out, ld_out);
  post_launch("mlp");
}

// log r = logsigmoid(T) - logsigmoid(-T) (disc) | log(softplus(T) + 1e-8) (rulsif); r = exp(log r)
// reference: ratio_estimator.py:182-189, flow_utils.py:256-257
__global__ void ratio_head_kernel(const float* __restrict__ scores, int loss_type, float* __restrict__ lr,
                                  float* __restrict__ r, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float T = scores[i];
  float v;
  if (loss_type == 0) {
    const float l1p = log1pf(expf(-fabsf(T)));
    const float a = fminf(T, 0.f) - l1p;
    const float b = fminf(-T, 0.f) - l1p;
    v = a - b;
  } else {
    const float sp = (T > 20.f) ? T : log1pf(expf(T));
    v = logf(sp + 1e-8f);
  }
  if (lr) lr[i] = v;
  if (r) r[i] = expf(v);
}

void launch_ratio_head(const float* scores, int loss_type, float* log_ratio, float* ratios, int n, cudaStream_t s) {
  ratio_head_kernel<<<(n + 127) / 128, 128, 0, s>>>(scores, loss_type, log_ratio, ratios, n);
  post_launch("ratio_head");
}

// ================================================================================================
// Input-gradient (dgrad-only) chain of the ratio estimators: d log r(x, y) / d(x, y), no weight gradients.
// README.md:159-164 (v = v_ind + gamma * grad log r); oracle: autograd through the restated estimator.
// ================================================================================================

// ---- score-net backward: one CTA per sample recomputes the MLP forward keeping every intermediate in shared memory,
// applies the head (disc: log r = T; rulsif: log(softplus(T) + 1e-8)) and back-propagates to the 512 features.

struct MlpPackB {
  MlpLayer l[4];
  int n;
};

__global__ void __launch_bounds__(256) mlp_bwd_kernel(const float* __restrict__ in, int ld_in, MlpPackB mp, int loss_type,
                                                      float* __restrict__ scores, float* __restrict__ log_ratio,
                                                      float* __restrict__ dfeat, int ld_d) {
  extern __shared__ float sm[];
  __shared__ float red[8];
  __shared__ float s_mean[4], s_rstd[4];
  const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // layout: a[0] = input (in0), then per layer: h (pre-LN), a (post-activation); finally two gradient buffers
  float* act[5];
  float* pre[4];
  int off = 0;
  act[0] = sm;
  off += mp.l[0].in;
  for (int li = 0; li < mp.n; ++li) {
    pre[li] = sm + off; off += mp.l[li].out;
    act[li + 1] = sm + off; off += mp.l[li].out;
  }
  float* gcur = sm + off; off += 512;
  float* gnext = sm + off;
  for (int i = tid; i < mp.l[0].in; i += 256) act[0][i] = in[(size_t)n * ld_in + i];
  __syncthreads();
  // ---- forward ----
  for (int li = 0; li < mp.n; ++li) {
    const MlpLayer L = mp.l[li];
    for (int j = warp; j < L.out; j += 8) {
      const float* wr = L.w + (size_t)j * L.in;
      float s = 0.f;
      for (int c = lane; c < L.in; c += 32) s = fmaf(wr[c], act[li][c], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) pre[li][j] = s + L.b[j];
    }
    __syncthreads();
    if (L.ln_g) {
      float s = 0.f;
      for (int c = tid; c < L.out; c += 256) s += pre[li][c];
      const float mean = block_sum256(s, red) / (float)L.out;
      float v = 0.f;
      for (int c = tid; c < L.out; c += 256) {
        const float d = pre[li][c] - mean;
        v = fmaf(d, d, v);
      }
      const float rstd = rsqrtf(block_sum256(v, red) / (float)L.out + 1e-5f);
      if (tid == 0) {
        s_mean[li] = mean;
        s_rstd[li] = rstd;
      }
      for (int c = tid; c < L.out; c += 256) {
        const float z = (pre[li][c] - mean) * rstd * L.ln_g[c] + L.ln_b[c];
        act[li + 1][c] = apply_act(z, L.act);
      }
    } else {
      for (int c = tid; c < L.out; c += 256) act[li + 1][c] = apply_act(pre[li][c], L.act);
    }
    __syncthreads();
  }
  // ---- head ----
  const float T = act[mp.n][0];
  float dT, lr;
  if (loss_type == 0) {
    const float l1p = log1pf(expf(-fabsf(T)));
    lr = (fminf(T, 0.f) - l1p) - (fminf(-T, 0.f) - l1p);
    dT = 1.0f;  // logsigmoid(T) - logsigmoid(-T) = T
  } else {
    const float sp = (T > 20.f) ? T : log1pf(expf(T));
    lr = logf(sp + 1e-8f);
    dT = (1.0f / (1.0f + expf(-T))) / (sp + 1e-8f);
  }
  if (tid == 0) {
    if (scores) scores[n] = T;
    if (log_ratio) log_ratio[n] = lr;
  }
  // ---- backward ----
  for (int c = tid; c < mp.l[mp.n - 1].out; c += 256) gcur[c] = (c == 0) ? dT : 0.f;
  __syncthreads();
  for (int li = mp.n - 1; li >= 0; --li) {
    const MlpLayer L = mp.l[li];
    // gcur = d / d act[li+1]  ->  dh (in place)
    if (L.ln_g) {
      const float mean = s_mean[li], rstd = s_rstd[li];
      float s1 = 0.f, s2 = 0.f;
      for (int c = tid; c < L.out; c += 256) {
        const float xh = (pre[li][c] - mean) * rstd;
        const float z = xh * L.ln_g[c] + L.ln_b[c];
        float dz = gcur[c];
        if (L.act == ACT_SILU) {
          const float sg = 1.0f / (1.0f + expf(-z));
          dz *= sg * (1.0f + z * (1.0f - sg));
        } else if (L.act == ACT_RELU) {
          dz = z > 0.f ? dz : 0.f;
        }
        const float dg = dz * L.ln_g[c];
        gcur[c] = dg;  // gamma * dz
        s1 += dg;
        s2 = fmaf(dg, xh, s2);
      }
      const float m1 = block_sum256(s1, red) / (float)L.out;
      const float m2 = block_sum256(s2, red) / (float)L.out;
      for (int c = tid; c < L.out; c += 256) {
        const float xh = (pre[li][c] - mean) * rstd;
        gcur[c] = rstd * (gcur[c] - m1 - xh * m2);
      }
    } else if (L.act != ACT_NONE) {
      for (int c = tid; c < L.out; c += 256) {
        const float z = pre[li][c];
        float dz = gcur[c];
        if (L.act == ACT_SILU) {
          const float sg = 1.0f / (1.0f + expf(-z));
          dz *= sg * (1.0f + z * (1.0f - sg));
        } else {
          dz = z > 0.f ? dz : 0.f;
        }
        gcur[c] = dz;
      }
    }
    __syncthreads();
    // d / d act[li] = W^T dh   (thread i walks column i of W: consecutive threads read consecutive addresses)
    for (int i = tid; i < L.in; i += 256) {
      float s = 0.f;
      for (int j = 0; j < L.out; ++j) s = fmaf(L.w[(size_t)j * L.in + i], gcur[j], s);
      gnext[i] = s;
    }
    __syncthreads();
    float* t = gcur;
    gcur = gnext;
    gnext = t;
  }
  for (int i = tid; i < mp.l[0].in; i += 256) dfeat[(size_t)n * ld_d + i] = gcur[i];
}

void launch_mlp_bwd(const float* in, int ld_in, const MlpLayer* layers, int nlayers, int loss_type, float* scores, float* log_ratio,
                    float* dfeat, int ld_d, int nimg, cudaStream_t s) {
  RGFM_CHECK(nlayers >= 1 && nlayers <= 4, "mlp_bwd: layer count");
  MlpPackB mp;
  mp.n = nlayers;
  size_t floats = layers[0].in + 2 * 512;
  for (int i = 0; i < nlayers; ++i) {
    mp.l[i] = layers[i];
    RGFM_CHECK(layers[i].in <= 512 && layers[i].out <= 512, "mlp_bwd: width");
    floats += 2 * (size_t)layers[i].out;
  }
  const size_t smem = floats * sizeof(float);
  RGFM_CHECK(smem <= 48 * 1024, "mlp_bwd: shared memory");
  mlp_bwd_kernel<<<nimg, 256, smem, s>>>(in, ld_in, mp, loss_type, scores, log_ratio, dfeat, ld_d);
  post_launch("mlp_bwd");
}

// ---- GAP + fc backward: feat = fc(mean_p act[p][c])  ->  dAct[p][c] = gscale * (W^T dfeat)[c] / HW for every pixel p
// gscale: the loss scale of the 16-bit gradient chain (1 in the fp32 mode); every later backward op is linear in it
__global__ void __launch_bounds__(256) gap_fc_bwd_kernel(const float* __restrict__ dfeat, int ld, const float* __restrict__ w, int Cout,
                                                         int HW, int C, float gscale, float* __restrict__ dact) {
  __shared__ float dm[PACK_MAXC];
  __shared__ float df[512];
  const int n = blockIdx.x;
  for (int j = threadIdx.x; j < Cout; j += blockDim.x) df[j] = dfeat[(size_t)n * ld + j];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < Cout; ++j) s = fmaf(w[(size_t)j * C + c], df[j], s);
    dm[c] = gscale * s / (float)HW;
  }
  __syncthreads();
  float* dst = dact + (size_t)n * HW * C;
  for (int i = threadIdx.x; i < HW * C; i += blockDim.x) dst[i] = dm[i % C];
}
void launch_gap_fc_bwd(const float* dfeat, int ld, const float* w, int Cout, int HW, int C, float gscale, float* dact, int nimg, cudaStream_t s) {
  RGFM_CHECK(C <= PACK_MAXC && Cout <= 512, "gap_fc_bwd: sizes");
  gap_fc_bwd_kernel<<<nimg, 256, 0, s>>>(dfeat, ld, w, Cout, HW, C, gscale, dact);
  post_launch("gap_fc_bwd");
}

// ---- pack backward: cur = pool?(act(norm(raw)))  ->  dRaw from dCur.
// Max-pool routes the gradient to the first maximum of the 2x2 window of act(norm(raw)) (recomputed); GroupNorm needs the
// two per-(image, group) moments of gamma * dz, so one CTA per image makes two passes (the image stays in L2).
// Every tensor is either dense fp32 NHWC (exact mode, and the layers the tensor-core kernel does not take) or the
// tensor-core kernel's 16-bit padded-linear raw layout (PackBwdIO): raw activations as the forward conv stored them,
// incoming gradients as the next layer's dgrad conv stored them, and the outgoing gradient with its pad rows, columns and
// channels written as zeros, so that the dgrad conv can use it as a zero-padded operand.
__global__ void __launch_bounds__(256) pack_bwd_kernel(PackArgs a, PackBwdIO io, const float* __restrict__ dcur, float* __restrict__ draw) {
  __shared__ float s_scale[PACK_MAXC];
  __shared__ float s_shift[PACK_MAXC];
  __shared__ float s_red[2][PACK_MAXC];
  __shared__ double s_grp[32][2];
  __shared__ float s_m1[32], s_m2[32];
  const int n = blockIdx.x, tid = threadIdx.x;
  pack_scale_shift(a, n, s_scale, s_shift, s_red, s_grp);  // scale = rstd * gamma (GN) | BN scale ; shift likewise
  const int C = a.C, H = a.Hs, Wd = a.Ws;
  const bool pool = a.spatial == SP_POOL2;
  const PackSrc& S = a.src[0];
  const int cpg = C / a.groups;
  auto raw_at = [&](int h, int w, int c) -> float { return load_raw1(S.p, raw_offset(S.padded, n, h, w, H, Wd, S.Cs) + c, S.fmt); };
  auto dcur_at = [&](int ho, int wo, int c) -> float {
    return load_raw1(dcur, raw_offset(io.dcur_padded, n, ho, wo, a.Ho, a.Wo, io.dcur_Cs) + c, io.dcur_fmt);
  };
  // gradient arriving at act(norm(raw))[h][w][c]
  auto grad_out = [&](int h, int w, int c) -> float {
    if (!pool) return dcur_at(h, w, c);
    const int ho = h >> 1, wo = w >> 1;
    if (ho >= a.Ho || wo >= a.Wo) return 0.f;  // odd sizes: the last row / column is not pooled
    // first maximum of the window in scan order
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int hh = 2 * ho + (k >> 1), ww = 2 * wo + (k & 1);
      const float z = fmaf(raw_at(hh, ww, c), s_scale[c], s_shift[c]);
      const float v = apply_act(z, a.act);
      if (v > best) {
        best = v;
        bi = k;
      }
    }
    return (bi == ((h & 1) * 2 + (w & 1))) ? dcur_at(ho, wo, c) : 0.f;
  };
  auto dact = [&](float z) -> float {
    if (a.act == ACT_SILU) {
      const float sg = 1.0f / (1.0f + expf(-z));
      return sg * (1.0f + z * (1.0f - sg));
    }
    if (a.act == ACT_RELU) return z > 0.f ? 1.0f : 0.f;
    return 1.0f;
  };
  // d raw of one real element
  const float cnt = (float)cpg * H * Wd;
  auto draw_of = [&](int h, int w, int c) -> float {
    const float z = fmaf(raw_at(h, w, c), s_scale[c], s_shift[c]);
    const float dz = grad_out(h, w, c) * dact(z);
    if (a.norm != NORM_GN) return dz * s_scale[c];
    const int g = c / cpg;
    const float xh = (a.gamma[c] != 0.f) ? (z - a.beta[c]) / a.gamma[c] : 0.f;
    const float rstd = (float)s_grp[g][1];
    return rstd * (dz * a.gamma[c] - s_m1[g] / cnt - xh * s_m2[g] / cnt);
  };
  const int total = H * Wd * C;
  if (a.norm == NORM_GN) {
    // pass A: per-group sums of gamma*dz and gamma*dz*xhat
    for (int g = tid; g < a.groups; g += blockDim.x) s_m1[g] = s_m2[g] = 0.f;
    __syncthreads();
    if (io.chan_sums) {
      // group moments of gamma * dz from the per-channel sums the training step already has
      for (int c = tid; c < C; c += blockDim.x) {
        float a0 = 0.f, a1 = 0.f;
        for (int sl = 0; sl < io.nslices; ++sl) {
          const float2 q = __ldg(reinterpret_cast<const float2*>(io.chan_sums + (((size_t)n * io.nslices + sl) * C + c) * 2));
          a0 += q.x;
          a1 += q.y;
        }
        s_red[0][c] = a.gamma[c] * a1;
        s_red[1][c] = a.gamma[c] * a0;
      }
      __syncthreads();
      for (int g = tid; g < a.groups; g += blockDim.x) {
        float u1 = 0.f, u2 = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
          u1 += s_red[0][c];
          u2 += s_red[1][c];
        }
        s_m1[g] = u1;
        s_m2[g] = u2;
      }
    } else if (256 % C == 0 && blockDim.x == 256) {
      // a thread owns ONE channel (i = tid + 256 k keeps i % C): register sums, then a fixed-order combine per group --
      // deterministic, unlike the shared-memory atomics of the general case below
      const int c = tid % C;
      float t1 = 0.f, t2 = 0.f;
      for (int pix = tid / C; pix < H * Wd; pix += 256 / C) {
        const int h = pix / Wd, w = pix - h * Wd;
        const float z = fmaf(raw_at(h, w, c), s_scale[c], s_shift[c]);
        const float dz = grad_out(h, w, c) * dact(z);
        const float dg = dz * a.gamma[c];
        const float xh = (a.gamma[c] != 0.f) ? (z - a.beta[c]) / a.gamma[c] : 0.f;
        t1 += dg;
        t2 = fmaf(dg, xh, t2);
      }
      s_red[0][tid] = t1;
      s_red[1][tid] = t2;
      __syncthreads();
      for (int g = tid; g < a.groups; g += blockDim.x) {
        float u1 = 0.f, u2 = 0.f;
        for (int j = 0; j < 256; ++j)
          if ((j % C) / cpg == g) {
            u1 += s_red[0][j];
            u2 += s_red[1][j];
          }
        s_m1[g] = u1;
        s_m2[g] = u2;
      }
    } else {
      for (int i = tid; i < total; i += blockDim.x) {
        const int c = i % C, pix = i / C, h = pix / Wd, w = pix - h * Wd;
        const float z = fmaf(raw_at(h, w, c), s_scale[c], s_shift[c]);
        const float dz = grad_out(h, w, c) * dact(z);
        const float dg = dz * a.gamma[c];
        const float xh = (a.gamma[c] != 0.f) ? (z - a.beta[c]) / a.gamma[c] : 0.f;
        atomicAdd(&s_m1[c / cpg], dg);
        atomicAdd(&s_m2[c / cpg], dg * xh);
      }
    }
    __syncthreads();
  }
  if (!io.draw_padded) {
    float* dr = draw + (size_t)n * H * Wd * C;
    for (int i = tid + blockIdx.y * blockDim.x; i < total; i += blockDim.x * gridDim.y) {
      const int c = i % C, pix = i / C, h = pix / Wd, w = pix - h * Wd;
      dr[i] = draw_of(h, w, c);
    }
  } else {
    // 16-bit padded-linear raw layout: every element of the image's (H + 1) x (W + 2) x Cs block is written (pads = 0)
    const int Cs = io.draw_Cs, P = Wd + 2;
    __half* dr = reinterpret_cast<__half*>(draw) + (size_t)n * (H + 1) * P * Cs;
    const int total_p = (H + 1) * P * Cs;
    for (int i = tid; i < total_p; i += blockDim.x) {
      const int c = i % Cs, pix = i / Cs, h = pix / P, w = pix - h * P;
      float v = 0.f;
      if (h < H && w < Wd && c < C) v = draw_of(h, w, c);
      dr[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    }
  }
}
// Streaming pack backward of the 16-bit gradient chain, per-channel affine norm (BatchNorm-eval) + SiLU, optional 2x2 max-pool:
// raw activation and outgoing gradient in the fp16 padded-linear layout, one thread per (pooling window | pixel, 8-channel
// octet), 16-byte loads and stores, the outgoing tensor's pad row / columns written as zeros by the threads of the
// window grid's last row / column.  Algorithmic bytes per element: 2 (raw) + 2 (d raw) + 0.5..4 (incoming gradient).
template <bool POOL>
__global__ void __launch_bounds__(256) pack_bwd_fast_kernel(PackArgs a, PackBwdIO io, const float* __restrict__ dcur, uint16_t* __restrict__ draw) {
  const int C = a.C, H = a.Hs, Wd = a.Ws, P = Wd + 2, Cs = io.draw_Cs, oct = Cs >> 3;
  const PackSrc& S = a.src[0];
  // index space: (image, window row, window column, octet); windows are 2x2 pixels (POOL) or single pixels
  const int gh = POOL ? H / 2 + 1 : H + 1, gw = POOL ? Wd / 2 + 1 : P;
  const long long total = (long long)a.nimg * gh * gw * oct;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int o8 = (int)(i % oct);
    long long r = i / oct;
    const int wx = (int)(r % gw);
    r /= gw;
    const int hy = (int)(r % gh), n = (int)(r / gh);
    const int c = o8 * 8;
    float sc[8], sh[8];
    const bool cval = c < C;  // channel padding of the tensors (C < Cs): zeros
    if (cval) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { sc[j] = __ldg(a.gamma + c + j); sh[j] = __ldg(a.beta + c + j); }
    }
    constexpr int NP = POOL ? 4 : 1;
    float g[8];  // incoming gradient of the window / pixel
    bool have_g = false;
    const int ho = hy, wo = wx;
    if (cval && ho < a.Ho && wo < a.Wo) {
      load_raw8(dcur, raw_offset(io.dcur_padded, n, ho, wo, a.Ho, a.Wo, io.dcur_Cs) + c, io.dcur_fmt, g);
      have_g = true;
    }
    float z[NP][8];
    bool real[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int h = POOL ? 2 * hy + (k >> 1) : hy, w = POOL ? 2 * wx + (k & 1) : wx;
      real[k] = cval && h < H && w < Wd;
      if (real[k]) {
        float x[8];
        load_raw8(S.p, raw_offset(S.padded, n, h, w, H, Wd, S.Cs) + c, S.fmt, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[k][j] = fmaf(x[j], sc[j], sh[j]);
      }
    }
    int win[8];  // POOL: the first maximum of act(z) in scan order takes the window's gradient
    if (POOL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float best = -INFINITY;
        int bi = 0;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          const float v = real[k] ? apply_act(z[k][j], a.act) : -INFINITY;
          if (v > best) { best = v; bi = k; }
        }
        win[j] = bi;
      }
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int h = POOL ? 2 * hy + (k >> 1) : hy, w = POOL ? 2 * wx + (k & 1) : wx;
      if (h > H || w >= P) continue;  // outside the padded image block
      uint32_t pk[4] = {0u, 0u, 0u, 0u};
      if (real[k] && have_g) {
        float d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float zz = z[k][j];
          float da = 1.0f;
          if (a.act == ACT_SILU) {
            const float sg = 1.0f / (1.0f + __expf(-zz));
            da = sg * (1.0f + zz * (1.0f - sg));
          } else if (a.act == ACT_RELU) {
            da = zz > 0.f ? 1.0f : 0.f;
          }
          const float go = (!POOL || win[j] == k) ? g[j] : 0.f;
          d[j] = fminf(fmaxf(go * da * sc[j], -65504.f), 65504.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __half2 hv = __floats2half2_rn(d[2 * j], d[2 * j + 1]);
          pk[j] = *reinterpret_cast<const uint32_t*>(&hv);
        }
      }
      *reinterpret_cast<uint4*>(draw + raw_offset(1, n, h, w, H, Wd, Cs) + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

void launch_pack_bwd(const PackArgs& a, const PackBwdIO& io, const float* dcur, float* draw, cudaStream_t s) {
  // streaming path: affine norm, fp16 padded-linear outgoing gradient, whole octets, even sizes when pooled
  const bool fast = a.nsrc == 1 && a.norm == NORM_AFFINE && ((a.src[0].fmt == 2 && a.src[0].padded) || a.src[0].fmt == 0) && io.draw_padded && io.draw_Cs % 8 == 0 &&
                    a.src[0].Cs % 8 == 0 && io.dcur_Cs % 8 == 0 && (io.dcur_fmt == 0 || io.dcur_fmt == 2) && a.C % 8 == 0 &&
                    (a.spatial == SP_ID || (a.spatial == SP_POOL2 && a.Hs % 2 == 0 && a.Ws % 2 == 0)) && a.spatial != SP_UP2;
  static const bool fast_off = getenv("RGFM_PACK_BWD_FAST") && atoi(getenv("RGFM_PACK_BWD_FAST")) == 0;  // A/B switch (same results)
  if (fast && !fast_off) {
    const bool pool = a.spatial == SP_POOL2;
    const long long total = (long long)a.nimg * (pool ? (a.Hs / 2 + 1) * (a.Ws / 2 + 1) : (a.Hs + 1) * (a.Ws + 2)) * (io.draw_Cs / 8);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    if (pool) pack_bwd_fast_kernel<true><<<blocks, 256, 0, s>>>(a, io, dcur, reinterpret_cast<uint16_t*>(draw));
    else pack_bwd_fast_kernel<false><<<blocks, 256, 0, s>>>(a, io, dcur, reinterpret_cast<uint16_t*>(draw));
    post_launch("pack_bwd_fast");
    return;
  }
  RGFM_CHECK(a.nsrc == 1 && (a.src[0].fmt == 0 || a.src[0].fmt == 2) && a.C <= PACK_MAXC && a.spatial != SP_UP2,
             "pack_bwd: one fp32 or fp16 source, no upsampling");
  RGFM_CHECK((io.dcur_fmt == 0 || io.dcur_fmt == 2) && (!io.draw_padded || io.draw_Cs >= a.C), "pack_bwd: gradient tensor formats");
  // several CTAs per image only where no CTA needs a pass over the whole image (no GroupNorm, or its sums are supplied)
  const int ys = (!io.draw_padded && (a.norm != NORM_GN || io.chan_sums)) ? std::max(1, io.yslices) : 1;
  pack_bwd_kernel<<<dim3(a.nimg, ys), 256, 0, s>>>(a, io, dcur, draw);
  post_launch("pack_bwd");
}

// x <- x + dt * (v + gamma_t * g)   (README.md:163: v_guided = v_ind + gamma * grad log r); gamma_t is read from the plan's
// per-step scalar table, so a constant strength, a sweep and a time-dependent schedule gamma(t) are the same kernel
__global__ void euler_grad_kernel(float* __restrict__ x, const float* __restrict__ v, const float* __restrict__ g,
                                  const StepScalars* __restrict__ sc, const int* __restrict__ step, float dt, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float gamma = sc[step ? *step : 0].gamma;
  float4 a = reinterpret_cast<float4*>(x)[i];
  const float4 b = __ldg(reinterpret_cast<const float4*>(v) + i);
  const float4 c = __ldg(reinterpret_cast<const float4*>(g) + i);
  a.x = __fadd_rn(a.x, __fmul_rn(__fadd_rn(b.x, __fmul_rn(gamma, c.x)), dt));
  a.y = __fadd_rn(a.y, __fmul_rn(__fadd_rn(b.y, __fmul_rn(gamma, c.y)), dt));
  a.z = __fadd_rn(a.z, __fmul_rn(__fadd_rn(b.z, __fmul_rn(gamma, c.z)), dt));
  a.w = __fadd_rn(a.w, __fmul_rn(__fadd_rn(b.w, __fmul_rn(gamma, c.w)), dt));
  reinterpret_cast<float4*>(x)[i] = a;
}
void launch_euler_grad(float* x, const float* v, const float* g, const StepScalars* sc, const int* step, float dt, size_t n,
                       cudaStream_t s) {
  RGFM_CHECK(n % 4 == 0, "euler_grad: n % 4");
  const size_t n4 = n / 4;
  euler_grad_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(x, v, g, sc, step, dt, n4);
  post_launch("euler_grad");
}

__global__ void step_inc_kernel(int* step) { *step += 1; }
void launch_step_inc(int* step, cudaStream_t s) {
  step_inc_kernel<<<1, 1, 0, s>>>(step);
  post_launch("step_inc");
}

// x <- x + v*dt, rounded like the reference (mul then add, no FMA).  reference: flow_utils.py:372-373
__global__ void euler_kernel(float* __restrict__ x, const float* __restrict__ v, float dt, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = reinterpret_cast<float4*>(x)[i];
  const float4 b = __ldg(reinterpret_cast<const float4*>(v) + i);
  a.x = __fadd_rn(a.x, __fmul_rn(b.x, dt));
  a.y = __fadd_rn(a.y, __fmul_rn(b.y, dt));
  a.z = __fadd_rn(a.z, __fmul_rn(b.z, dt));
  a.w = __fadd_rn(a.w, __fmul_rn(b.w, dt));
  reinterpret_cast<float4*>(x)[i] = a;
}
void launch_euler(float* x, const float* v, float dt, size_t n, cudaStream_t s) {
  RGFM_CHECK(n % 4 == 0, "euler: n % 4");
  size_t n4 = n / 4;
  euler_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(x, v, dt, n4);
  post_launch("euler");
}

}  // namespace rgfm

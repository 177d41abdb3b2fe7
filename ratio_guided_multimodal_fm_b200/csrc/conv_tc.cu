// tcgen05 implicit-GEMM convolution for sm_100a: bf16 operands staged in shared memory by TMA, fp32 accumulators
// in TMEM, one elected thread issuing tcgen05.mma, epilogue reading TMEM with tcgen05.ld.
//
// GEMM view of a 3x3 (or 1x1) stride-1 convolution over an NHWC bf16 operand:
//   M = 128 output pixels (TileGeom: whole rows of one image, or several whole small images)
//   N = CoutP (16..256, all output channels in one tile -> the accumulator is 128 lanes x N TMEM columns)
//   K = taps x Cin, walked in K-blocks of BK = 64 (or 32) channels of one tap.
// The A tile of K-block (tap=(ky,kx), c0) is the SAME pixel tile shifted by (ky-1, kx-1): one 4-D TMA box
// {BK, W, th, imgs} at coordinates (c0, kx-1, h0+ky-1, n0).  TMA zero-fills out-of-bounds coordinates, which
// implements the conv's zero padding for free and lands the rows in exactly the K-major 128B-swizzled layout the
// UMMA shared-memory descriptor expects.  The B tile is a {BK, CoutP, 1} box of the repacked weights
// [tap][CoutP][Cin].  Stride-2 convolutions are computed at stride 1 and stored subsampled.
//
// Persistent kernel, one CTA per SM, static round-robin over tiles.  Warp roles (320 threads): warp 0 = TMA
// producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 and 6..9 = two epilogue groups, each owning one of
// two TMEM accumulators, so the MMAs of tile i+1 overlap the epilogue of tile i (TMEM -> registers -> + bias +
// time bias + residual -> fp32 NHWC store, per-tile GroupNorm partial sums by a butterfly transpose-reduce, or
// NCHW velocity / fused Euler update for the output conv).  The smem ring runs across tile boundaries.
#include "net.h"
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstring>

namespace rgfm {

static inline void post_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(std::string(what) + " launch failed: " + cudaGetErrorString(e));
  ++g_launch_counter;
}

constexpr int TC_MAX_STAGES = 6;
constexpr int TC_THREADS = 64 + 2 * 128;  // producer warp, MMA warp, two epilogue groups of 4 warps

struct TcArgs {
  ConvArgs c;       // epilogue description; c.g is the tile geometry of the operand (= stride-1 output) space
  int nkb;          // K-blocks = taps * cblocks
  int cblocks;      // Cin / BK
  int BK;           // 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
  int ks, pad;
  int stages;
  uint32_t a_bytes; // TMA box bytes of one A stage (rows actually loaded)
  uint32_t a_stage; // reserved bytes per A stage (128 rows)
  uint32_t b_bytes; // bytes of one B stage
  uint32_t layout;  // UMMA layout type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
  uint32_t sbo;     // stride between 8-row groups in bytes
  uint32_t idesc;
  uint32_t tmem_cols;
  int sub;          // stride-2: store rows with even (ho, wo) only
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  // bounded spin: a protocol bug becomes a trap ("unspecified launch failure"), never a hung GPU
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}

// 32 lanes x 32 values -> lane L ends with the sum over lanes of value L (31 shuffles)
__device__ __forceinline__ float transpose_reduce32(float* f, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = hi ? f[i] : f[i + off];
      const float keep = hi ? f[i + off] : f[i];
      f[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return f[0];
}

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[256];
  __shared__ float s_part[2][4][256][2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = p.a_stage + p.b_bytes;
  const ConvArgs& a = p.c;
  const int tiles = a.g.tiles(a.nimg);

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    // bias + this step's time bias: identical for every tile of the launch (t is uniform over the batch)
    const float* tb = a.temb ? a.temb + (size_t)(a.step ? *a.step : 0) * a.temb_stride + a.temb_off : nullptr;
    for (int c = threadIdx.x - 64; c < a.CoutP; c += TC_THREADS - 64) {
      float b = a.bias[c];
      if (tb && c < a.Cout) b += tb[c];
      s_bias[c] = b;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      uint32_t g = 0;  // running K-block counter across tiles (ring position)
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int n0, h0;
        if (a.g.imgs == 1) {
          n0 = tile / a.g.parts;
          h0 = (tile - n0 * a.g.parts) * a.g.th;
        } else {
          n0 = tile * a.g.imgs;
          h0 = 0;
        }
        for (int kb = 0; kb < p.nkb; ++kb, ++g) {
          const uint32_t s = g % p.stages;
          const uint32_t ph = (g / p.stages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
          const int ky = tap / p.ks, kx = tap - ky * p.ks;
          mbar_expect_tx(&full_bar[s], p.a_bytes + p.b_bytes);
          const uint32_t sa = smem_base + s * stage_bytes;
          tma_load_4d(sa, &tmA, &full_bar[s], cb * p.BK, kx - p.pad, h0 + ky - p.pad, n0);
          tma_load_3d(sa + p.a_stage, &tmB, &full_bar[s], cb * p.BK, 0, tap);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    if (lane == 0) {
      uint32_t g = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&tmem_empty_bar[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t td = tmem_d + (uint32_t)(buf * a.CoutP);
        for (int kb = 0; kb < p.nkb; ++kb, ++g) {
          const uint32_t s = g % p.stages;
          const uint32_t ph = (g / p.stages) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * stage_bytes;
          const uint64_t da = make_desc(sa, p.sbo, p.layout);
          const uint64_t db = make_desc(sa + p.a_stage, p.sbo, p.layout);
          const int nk = p.BK >> 4;
          for (int k = 0; k < nk; ++k)  // advance 16 elements = 32 bytes inside the swizzle atom: +2 in the >>4 address field
            tc_mma_bf16(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (uint32_t)((kb | k) != 0));
          tc_commit(&empty_bar[s]);
        }
        tc_commit(&tmem_full_bar[buf]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------- epilogue: two groups of 4 warps, one per accumulator -------------------------------
    const int eg = (warp - 2) >> 2;  // epilogue group == TMEM accumulator index
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64 - eg * 128;  // thread index within the group
    int it = eg;
    for (int tile = blockIdx.x + eg * gridDim.x; tile < tiles; tile += 2 * gridDim.x, it += 2) {
      int n, ho, wo, part;
      bool valid = tile_row(a.g, tile, r, a.nimg, n, ho, wo, part);
      int Ho = a.g.H, Wo = a.g.W;
      if (p.sub) {
        valid = valid && !(ho & 1) && !(wo & 1);
        ho >>= 1; wo >>= 1; Ho >>= 1; Wo >>= 1;
      }
      mbar_wait(&tmem_full_bar[eg], ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(eg * a.CoutP);
      if (a.out_mode == OUT_RAW) {
        const size_t o = valid ? (((size_t)n * Ho + ho) * Wo + wo) * a.Cout : 0;
        for (int c0 = 0; c0 < a.CoutP; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + s_bias[c0 + j];
          if (valid) {
            if (a.res) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 rv = *reinterpret_cast<const float4*>(a.res + o + c0 + j);
                f[j] += rv.x; f[j + 1] += rv.y; f[j + 2] += rv.z; f[j + 3] += rv.w;
              }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(a.out + o + c0 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          }
          if (a.stats) {
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              f[j] = valid ? f[j] : 0.f;
              sq[j] = f[j] * f[j];
            }
            const float s1 = transpose_reduce32(f, lane);
            const float s2 = transpose_reduce32(sq, lane);
            s_part[eg][q][c0 + lane][0] = s1;
            s_part[eg][q][c0 + lane][1] = s2;
          }
        }
        // accumulator fully read: hand it back to the MMA warp before the (global-memory) statistics tail
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[eg]);
        if (a.stats) {
          if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
          for (int c = et; c < a.Cout; c += 128) {
            if (a.g.imgs == 1) {
              const int ni = tile / a.g.parts, pi = tile - ni * a.g.parts;
              if (ni < a.nimg) {
                float* st = a.stats + (((size_t)ni * a.g.parts + pi) * a.Cout + c) * 2;
                st[0] = (s_part[eg][0][c][0] + s_part[eg][1][c][0]) + (s_part[eg][2][c][0] + s_part[eg][3][c][0]);
                st[1] = (s_part[eg][0][c][1] + s_part[eg][1][c][1]) + (s_part[eg][2][c][1] + s_part[eg][3][c][1]);
              }
            } else {  // two 64-pixel images per tile: quarters {0,1} and {2,3}
#pragma unroll
              for (int il = 0; il < 2; ++il) {
                const int ni = tile * 2 + il;
                if (ni < a.nimg) {
                  float* st = a.stats + ((size_t)ni * a.Cout + c) * 2;
                  st[0] = s_part[eg][2 * il][c][0] + s_part[eg][2 * il + 1][c][0];
                  st[1] = s_part[eg][2 * il][c][1] + s_part[eg][2 * il + 1][c][1];
                }
              }
            }
          }
          // s_part[eg] is rewritten by the next tile of this group: wait until everyone has read it
          if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
        }
      } else {
        // output conv: Cout <= 16 real channels -> NCHW velocity, or fused Euler update x += dt * v
        uint32_t v[16];
        tmem_ld16(taddr, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[eg]);
        if (valid) {
          for (int c = 0; c < a.Cout; ++c) {
            const float val = __uint_as_float(v[c]) + s_bias[c];
            const size_t oc = (((size_t)n * a.Cout + c) * Ho + ho) * Wo + wo;
            if (a.out_mode == OUT_NCHW)
              a.out[oc] = val;
            else
              a.out[oc] = __fadd_rn(a.out[oc], __fmul_rn(val, a.dt));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(p.tmem_cols) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    RGFM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    RGFM_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct TcConv {
  CUtensorMap tmA, tmB;
  TcArgs p;
  size_t smem;
  uint16_t* w = nullptr;
};
void TcDeleter::operator()(TcConv* p) const { delete p; }

bool tc_available() {
  int dev = 0;
  cudaDeviceProp pr{};
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&pr, dev) != cudaSuccess) return false;
  return pr.major == 10;
}

TcPtr make_tc_conv(const TcConvDesc& d, DevPool& pool) {
  RGFM_CHECK(d.Cin % 32 == 0, "tc conv: operand channels must be a multiple of 32");
  RGFM_CHECK(d.CoutP % 16 == 0 && d.CoutP >= 16 && d.CoutP <= 256, "tc conv: padded Cout must be 16..256, multiple of 16");
  RGFM_CHECK(d.ks == 1 || d.ks == 3, "tc conv: 1x1 or 3x3");
  TcPtr t(new TcConv());
  TcArgs& p = t->p;
  std::memset(&p, 0, sizeof(p));
  const int BK = (d.Cin % 64 == 0) ? 64 : 32;
  const TileGeom g = make_geom(d.Hin, d.Win);
  p.BK = BK;
  p.cblocks = d.Cin / BK;
  p.ks = d.ks;
  p.pad = d.ks / 2;
  p.nkb = d.ks * d.ks * p.cblocks;
  p.a_stage = TM * BK * 2;
  p.a_bytes = (uint32_t)(g.W * g.th * g.imgs) * BK * 2;
  p.b_bytes = (uint32_t)d.CoutP * BK * 2;
  p.layout = BK == 64 ? 2u : 4u;
  p.sbo = 8u * BK * 2;
  const uint32_t fmt = d.fp16 ? 0u : 1u;  // UMMA F16F32Format: 0 = F16, 1 = BF16
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(d.CoutP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)d.CoutP) cols <<= 1;  // two accumulators (MMA of tile i+1 overlaps the epilogue of tile i)
  p.tmem_cols = cols;
  const uint32_t stage = p.a_stage + p.b_bytes;
  int stages = (int)((176u * 1024u) / stage);  // one persistent CTA per SM
  if (stages < 3) stages = 3;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages > p.nkb) stages = p.nkb;
  p.stages = stages;
  t->smem = (size_t)stages * stage + 1024;

  // weights: torch [Cout][Cin_real][ks][ks] fp32 -> [tap][CoutP][Cin] bf16 (zero padded)
  const int taps = d.ks * d.ks;
  std::vector<uint16_t> hw((size_t)taps * d.CoutP * d.Cin, 0);
  for (int o = 0; o < d.Cout; ++o)
    for (int c = 0; c < d.Cin_real; ++c)
      for (int tp = 0; tp < taps; ++tp) {
        const float wv = d.w_oihw[((size_t)o * d.Cin_real + c) * taps + tp];
        uint16_t bits;
        if (d.fp16) {
          const __half h = __float2half_rn(wv);
          std::memcpy(&bits, &h, 2);
        } else {
          const __nv_bfloat16 h = __float2bfloat16(wv);
          std::memcpy(&bits, &h, 2);
        }
        hw[((size_t)tp * d.CoutP + o) * d.Cin + c] = bits;
      }
  t->w = pool.upload(hw);

  const CUtensorMapDataType dt16 = d.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d.Cin, (cuuint64_t)d.Win, (cuuint64_t)d.Hin, (cuuint64_t)d.chunk};
    cuuint64_t strides[3] = {(cuuint64_t)d.Cin * 2, (cuuint64_t)d.Win * d.Cin * 2, (cuuint64_t)d.Hin * d.Win * d.Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)g.W, (cuuint32_t)g.th, (cuuint32_t)g.imgs};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode_fn()(&t->tmA, dt16, 4, const_cast<void*>(d.in), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RGFM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed with code " + std::to_string((int)r));
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)d.Cin, (cuuint64_t)d.CoutP, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)d.Cin * 2, (cuuint64_t)d.CoutP * d.Cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)d.CoutP, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_fn()(&t->tmB, dt16, 3, t->w, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RGFM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed with code " + std::to_string((int)r));
  }
  static bool attr = false;
  if (!attr) {
    RGFM_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  return t;
}

void launch_conv_tc(TcConv& tc, const ConvArgs& a, cudaStream_t s) {
  TcArgs p = tc.p;
  p.c = a;
  p.sub = (a.stride == 2) ? 1 : 0;
  p.c.g = make_geom(a.Hin, a.Win);  // tiles walk the operand (stride-1 output) space
  if (a.stats) RGFM_CHECK(p.c.g.imgs == 1 || (p.c.g.imgs == 2 && !p.sub), "tc conv: GroupNorm partials need <=2 images per tile");
  if (a.out_mode != OUT_RAW) RGFM_CHECK(a.CoutP == 16, "tc conv: the output conv uses a 16-column accumulator");
  else RGFM_CHECK(a.CoutP % 32 == 0 && a.Cout == a.CoutP, "tc conv: raw outputs need Cout % 32 == 0");
  const int tiles = p.c.g.tiles(a.nimg);
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    RGFM_CUDA(cudaGetDevice(&dev));
    RGFM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = tiles < num_sms ? tiles : num_sms;  // persistent: one CTA per SM, static round-robin over tiles
  conv_tc_kernel<<<grid, TC_THREADS, tc.smem, s>>>(tc.tmA, tc.tmB, p);
  post_launch("conv_tc");
}

}  // namespace rgfm

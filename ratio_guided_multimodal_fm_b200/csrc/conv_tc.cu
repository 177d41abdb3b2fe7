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
#include <cstdlib>

namespace rgfm {

static inline void post_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(std::string(what) + " launch failed: " + cudaGetErrorString(e));
  ++g_launch_counter;
}

constexpr int TC_MAX_STAGES = 12;
constexpr int TC_THREADS = 96 + 2 * 128;  // producer warp, two MMA warps (one per accumulator half), two epilogue groups of 4 warps
constexpr int TC_SUPER = 256;             // output rows per super-tile (two M=128 accumulators sharing every B tile)
constexpr uint32_t TC_STAGE_BYTES = 128 * 128;  // epilogue staging buffer: 128 rows x 32 fp32, 128B-swizzled

struct TcArgs {
  ConvArgs c;           // epilogue description (bias, time bias, output, nimg, out_mode, dt)
  int H, W, P;          // operand image geometry and padded pitch P = W + 2
  int R;                // GEMM rows per image: (H + 1) * P
  int nseg;             // K segments: [0] the conv proper, [1..2] the folded 1x1 skip conv / identity residual
  int cblocks[3];       // operand channels / 64 per segment
  int ksteps[3][4];     // K16 steps that hold real channels in each 64-channel block (all-zero steps are never issued)
  int ntaps[3];         // 9 (3x3) or 1 (1x1)
  int tap_shift[3][9];  // row shift of each tap inside the A tile: ky*P + kx (3x3), P + 1 (1x1)
  int row_off[3];       // TMA row coordinate of GEMM row 0: 0 for packed operands, -(P + 1) for raw activation tensors
  int a_box_rows;       // rows per TMA box; an A tile is two boxes
  uint32_t a_slot;      // bytes of one A slot (2 * a_box_rows * 128)
  uint32_t b_bytes;     // bytes of one weight tap tile (nmma * 128)
  uint32_t b_stage;     // bytes of one weight-ring stage (taps_per_stage * b_bytes)
  int tps[3];           // taps per weight-ring stage, per segment: barrier round trips are amortised over tps*4 MMAs
  int b_stages;
  int nmma;             // N of the MMA / TMEM columns per accumulator (Cout rounded up to 16)
  int nbuf;             // TMEM double buffering (2) when 4 * nmma <= 512 columns, else 1
  uint32_t idesc[3];    // per segment: packed operands use the net's 16-bit format, raw activations are always fp16
  uint32_t tmem_cols;
  int sub;              // stride-2: store rows with even (h, w) only, into the half-resolution raw tensor
  int tma_store;        // raw output leaves through swizzled smem staging + TMA store (fully coalesced)
  int out_H, out_W;     // output spatial size (H, W or halved when sub)
  int out_fmt;          // raw output storage: 1 = bf16, 2 = fp16
  float* stats;         // GroupNorm partials [unit][nslots][CoutP][2] of the rounded output (null = none)
  int nslots;
  int dbg;              // experiment switches (RGFM_TC_DBG): 1 = epilogue does no work, 2 = producer loads nothing
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
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
__device__ __forceinline__ void tc_mma_bf16_acc(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.u32 p, 1, 1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc)
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

__device__ __forceinline__ uint32_t cvt2(float a, float b, int fmt) {
  if (fmt == 1) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  const float L = 65504.f;
  __half2 v = __floats2half2_rn(fminf(fmaxf(a, -L), L), fminf(fmaxf(b, -L), L));
  return *reinterpret_cast<uint32_t*>(&v);
}
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


__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1) : "memory");
}

// decode padded-linear output index m' -> (image, h, w); valid iff a real pixel of an image < nimg
__device__ __forceinline__ bool slot_pixel(int m, int H, int P, int nimg, int& n, int& h, int& w) {
  const int hv = m / P;
  w = m - hv * P;
  n = hv / (H + 1);
  h = hv - n * (H + 1);
  return (w < P - 2) && (h < H) && (n < nimg);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
               const __grid_constant__ CUtensorMap tmO, const TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[2], a_empty[2];
  __shared__ __align__(8) uint64_t b_full[TC_MAX_STAGES], b_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[256];
  __shared__ __align__(16) float s_part[2][4][3][64][2];  // per epilogue group / warp / image slot: {sum, sumsq} of a 64-channel chunk

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_stage = smem_base + 2u * p.a_slot;       // epilogue staging tiles (one per epilogue group)
  const uint32_t smem_b = smem_stage + 2u * TC_STAGE_BYTES;    // weight ring
  const ConvArgs& a = p.c;
  const int N = p.nmma;
  const int total_rows = a.nimg * (p.H + 1) * p.P;
  const int supers = (total_rows + TC_SUPER - 1) / TC_SUPER;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB0)) : "memory");
    if (p.nseg > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA1)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB1)) : "memory");
    }
    if (p.nseg > 2) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA2)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB2)) : "memory");
    }
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 2);         // both MMA warps
      mbar_init(&tmem_full_bar[s], 2);   // both MMA warps
      mbar_init(&tmem_empty_bar[s], 256);
    }
    for (int s = 0; s < p.b_stages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 3) {
    // bias + this step's time bias: identical for every tile of the launch (t is uniform over the batch)
    const float* tb = a.temb ? a.temb + (size_t)(a.step ? *a.step : 0) * a.temb_stride + a.temb_off : nullptr;
    for (int c = threadIdx.x - 96; c < N; c += TC_THREADS - 96) {
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
    // The whole warp runs the (warp-uniform) loop; one elected lane issues.  Ring positions are kept as
    // (index, phase) counters: no divisions on the critical path.
    uint32_t as = 0, aph = 0, bs = 0, bph = 0;
    for (int st = blockIdx.x; st < supers; st += gridDim.x) {
      const int row0 = st * TC_SUPER;
      for (int seg = 0; seg < p.nseg; ++seg) {
        const CUtensorMap* tA = seg == 0 ? &tmA0 : (seg == 1 ? &tmA1 : &tmA2);
        const CUtensorMap* tB = seg == 0 ? &tmB0 : (seg == 1 ? &tmB1 : &tmB2);
        const int ncb = p.cblocks[seg], ntp = p.ntaps[seg];
        for (int cb = 0; cb < ncb; ++cb) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          if (elect_one()) {
            if (p.dbg & 2) {
              mbar_arrive(&a_full[as]);
            } else {
              mbar_expect_tx(&a_full[as], p.a_slot);
              const uint32_t sa = smem_base + as * p.a_slot;
              const int ra = row0 + p.row_off[seg];  // negative / past-the-end rows are zero-filled by TMA
              tma_load_2d(sa, tA, &a_full[as], cb * 64, ra);
              tma_load_2d(sa + (p.a_slot >> 1), tA, &a_full[as], cb * 64, ra + p.a_box_rows);
            }
          }
          __syncwarp();
          as ^= 1u;
          if (as == 0u) aph ^= 1u;
          const int tps = p.tps[seg];
          for (int tap = 0; tap < ntp; tap += tps) {
            mbar_wait(&b_empty[bs], bph ^ 1u);
            if (elect_one()) {
              if (p.dbg & 2) {
                mbar_arrive(&b_full[bs]);
              } else {
                mbar_expect_tx(&b_full[bs], (uint32_t)tps * p.b_bytes);
                for (int j = 0; j < tps; ++j)
                  tma_load_3d(smem_b + bs * p.b_stage + (uint32_t)j * p.b_bytes, tB, &b_full[bs], cb * 64, 0, tap + j);
              }
            }
            __syncwarp();
            if (++bs == (uint32_t)p.b_stages) {
              bs = 0u;
              bph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp <= 2) {
    // ------------------- MMA issuers: warp 1 accumulates rows 0..127, warp 2 rows 128..255 of each super-tile -------------------
    // Small-N MMAs (32-64 tensor-pipe cycles each) are issue-bound from a single thread, so the two independent
    // accumulators are driven by two warps and every descriptor is a base + immediate.
    const uint32_t hf = (uint32_t)(warp - 1);
    uint32_t as = 0, aph = 0, bs = 0, bph = 0;
    const uint64_t dhi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;  // SBO = 1024 B, version 1, SWIZZLE_128B
    int it = 0;
    for (int st = blockIdx.x; st < supers; st += gridDim.x, ++it) {
      const int buf = (p.nbuf == 2) ? (it & 1) : 0;
      const uint32_t use = (p.nbuf == 2) ? ((uint32_t)it >> 1) : (uint32_t)it;
      mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u);  // both epilogue groups drained this accumulator pair
      tc_fence_after();
      const uint32_t td = tmem_d + (uint32_t)((buf * 2 + (int)hf) * N);
      uint32_t first = 0;
      for (int seg = 0; seg < p.nseg; ++seg) {
        const int ncb = p.cblocks[seg], ntp = p.ntaps[seg];
        for (int cb = 0; cb < ncb; ++cb) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t sa = smem_base + as * p.a_slot + hf * (128u * 128u);
          const int tps = p.tps[seg];
          const int ks = p.ksteps[seg][cb];
          const uint32_t idesc = p.idesc[seg];
          for (int tap = 0; tap < ntp; tap += tps) {
            if (!(p.dbg & 4)) {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
            }
            if (elect_one()) {
              for (int j = 0; j < tps; ++j) {
                // the tap is a ROW SHIFT of the shared A tile: descriptor start = tile base + shift * 128 B;
                // +32 bytes (+2 in the >>4 address field) per K16 step inside the swizzle atom
                const uint64_t da = dhi | (uint64_t)((sa + (uint32_t)p.tap_shift[seg][tap + j] * 128u) >> 4);
                const uint64_t db = dhi | (uint64_t)((smem_b + bs * p.b_stage + (uint32_t)j * p.b_bytes) >> 4);
                if (!(p.dbg & 8)) {
                  tc_mma_bf16(td, da, db, idesc, first);
                  if (ks > 1) tc_mma_bf16_acc(td, da + 2, db + 2, idesc);
                  if (ks > 2) tc_mma_bf16_acc(td, da + 4, db + 4, idesc);
                  if (ks > 3) tc_mma_bf16_acc(td, da + 6, db + 6, idesc);
                }
                first = 1u;
              }
              tc_commit(&b_empty[bs]);
              if (tap + tps >= ntp) tc_commit(&a_empty[as]);
            }
            __syncwarp();
            first = 1u;
            if (++bs == (uint32_t)p.b_stages) {
              bs = 0u;
              bph ^= 1u;
            }
          }
          as ^= 1u;
          if (as == 0u) aph ^= 1u;
        }
      }
      if (elect_one()) tc_commit(&tmem_full_bar[buf]);
      __syncwarp();
    }
  } else {
    // ---------------- epilogue: group eg owns rows [128*eg, 128*eg+128) of every super-tile ----------------
    const int eg = (warp - 3) >> 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int et = threadIdx.x - 96 - eg * 128;
    const int wq = et >> 5;  // warp inside the group (row quarter for the column sums)
    const int r = q * 32 + lane;  // row inside the 128-row unit
    const int CP = a.CoutP;  // channel stride of the raw output (multiple of 64; channels >= nmma are written as zeros)
    const uint32_t stg = smem_stage + (uint32_t)eg * TC_STAGE_BYTES;
    int it = 0;
    for (int st = blockIdx.x; st < supers; st += gridDim.x, ++it) {
      const int buf = (p.nbuf == 2) ? (it & 1) : 0;
      const uint32_t use = (p.nbuf == 2) ? ((uint32_t)it >> 1) : (uint32_t)it;
      const int unit = st * 2 + eg;  // 128-row unit index
      const int m = unit * 128 + r;
      int n, h, w;
      bool valid = slot_pixel(m, p.H, p.P, a.nimg, n, h, w);  // a real pixel of an image < nimg
      // where this thread's row goes in a stride-2 (subsampled) output: even (h, w) rows map to (h/2, w/2) -- including
      // the pad row h == H and pad column w == W, which are written as zeros -- and the otherwise idle column
      // w == W + 1 writes the second pad column, so the half-resolution raw tensor is zero padded like every other one
      bool wr = false;
      size_t o = 0;
      if (p.sub) {
        const bool inimg = (n < a.nimg) && (h <= p.H);
        valid = valid && !(h & 1) && !(w & 1);
        if (inimg && !(h & 1) && (!(w & 1) || w == p.W + 1)) {
          wr = true;
          const int w2 = (w == p.W + 1) ? p.out_W + 1 : (w >> 1);
          o = ((size_t)((size_t)n * (p.out_H + 1) + (h >> 1)) * (p.out_W + 2) + w2) * CP;
        }
      }
      mbar_wait(&tmem_full_bar[buf], use & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * 2 + eg) * N);
      if (p.dbg & 1) {
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[buf]);
        continue;
      }
      if (a.out_mode == OUT_RAW) {
        // raw activation, 16-bit, channel stride CP; rows that are not real pixels are stored as exact zeros, so the
        // tensor doubles as a zero-padded conv operand and the per-channel sums below need no masking
        uint16_t* out16 = reinterpret_cast<uint16_t*>(a.out);
        // image boundaries inside this unit (for the GroupNorm partials): rows >= b1 belong to slot 1, >= b2 to slot 2
        const int nf = (unit * 128) / p.R;
        const int b1 = (nf + 1) * p.R - unit * 128;
        const int b2 = b1 + p.R;
        for (int c0 = 0; c0 < CP; c0 += 64) {
          if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous store has left the buffer
          if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t pk[16];
            if (c0 + hh * 32 < N) {
              uint32_t v[32];
              tmem_ld32(taddr + c0 + hh * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const uint32_t pv = cvt2(__uint_as_float(v[2 * j]) + s_bias[c0 + hh * 32 + 2 * j],
                                         __uint_as_float(v[2 * j + 1]) + s_bias[c0 + hh * 32 + 2 * j + 1], p.out_fmt);
                pk[j] = valid ? pv : 0u;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = 0u;
            }
            // registers -> 128B-swizzled staging tile [128 rows x 64 channels x 2 B]
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t dst = stg + (uint32_t)r * 128u + (uint32_t)(((hh * 4 + j) ^ (r & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * j]), "r"(pk[4 * j + 1]),
                           "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3]) : "memory");
            }
            if (!p.tma_store && wr) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(out16 + o + c0 + hh * 32 + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            }
          }
          if (c0 + 64 >= N) {  // last chunk that reads TMEM: hand the accumulator back to the MMA warps
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[buf]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
          if (p.tma_store && et == 0) {
            tma_store_2d(&tmO, stg, c0, unit * 128);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (p.stats != nullptr && c0 < N) {
            // GroupNorm partials of the ROUNDED values: thread (wq, lane) sums channel pair `lane` over the 32 rows
            // of quarter wq, split by image slot (warp-uniform), then the four quarters are combined in order
            float s0a = 0.f, s0b = 0.f, q0a = 0.f, q0b = 0.f, s1a = 0.f, s1b = 0.f, q1a = 0.f, q1b = 0.f;
            float s2a = 0.f, s2b = 0.f, q2a = 0.f, q2b = 0.f;
            const uint32_t colb = (uint32_t)((lane & 3) << 2);
            const uint32_t chunk16 = (uint32_t)(lane >> 2);
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
              const int row = wq * 32 + rr;
              uint32_t u;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(stg + (uint32_t)row * 128u + ((chunk16 ^ (uint32_t)(row & 7)) << 4) + colb));
              float2 f;
              if (p.out_fmt == 1) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
              else f = __half22float2(*reinterpret_cast<const __half2*>(&u));
              if (row < b1) {
                s0a += f.x; s0b += f.y; q0a = fmaf(f.x, f.x, q0a); q0b = fmaf(f.y, f.y, q0b);
              } else if (row < b2) {
                s1a += f.x; s1b += f.y; q1a = fmaf(f.x, f.x, q1a); q1b = fmaf(f.y, f.y, q1b);
              } else {
                s2a += f.x; s2b += f.y; q2a = fmaf(f.x, f.x, q2a); q2b = fmaf(f.y, f.y, q2b);
              }
            }
            float* sp = &s_part[eg][wq][0][2 * lane][0];
            *reinterpret_cast<float4*>(sp) = make_float4(s0a, q0a, s0b, q0b);
            *reinterpret_cast<float4*>(sp + 128) = make_float4(s1a, q1a, s1b, q1b);
            *reinterpret_cast<float4*>(sp + 256) = make_float4(s2a, q2a, s2b, q2b);
            if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            for (int i = et; i < p.nslots * 128; i += 128) {
              const int slot = i >> 7, rem = i & 127;
              const float* pp = &s_part[eg][0][slot][0][0] + rem;
              const float tot = ((pp[0] + pp[3 * 128]) + pp[2 * 3 * 128]) + pp[3 * 3 * 128];
              p.stats[((size_t)(unit * p.nslots + slot) * CP + c0) * 2 + rem] = tot;
            }
          }
        }
      } else {
        // output conv: Cout <= 16 real channels -> NCHW velocity, or fused Euler update x += dt * v
        uint32_t v[16];
        tmem_ld16(taddr, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[buf]);
        if (valid) {
          for (int c = 0; c < a.Cout; ++c) {
            const float val = __uint_as_float(v[c]) + s_bias[c];
            const size_t oc = (((size_t)n * a.Cout + c) * p.out_H + h) * p.out_W + w;
            if (a.out_mode == OUT_NCHW)
              a.out[oc] = val;
            else
              a.out[oc] = __fadd_rn(a.out[oc], __fmul_rn(val, a.dt));
          }
        }
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // smem must outlive the stores
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
  CUtensorMap tmA[3], tmB[3], tmO;
  TcArgs p;
  size_t smem;
};
void TcDeleter::operator()(TcConv* p) const { delete p; }

bool tc_available() {
  int dev = 0;
  cudaDeviceProp pr{};
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&pr, dev) != cudaSuccess) return false;
  return pr.major == 10;
}

size_t tc_operand_slots(int H, int W, int chunk) { return (size_t)(1 + (size_t)chunk * (H + 1)) * (W + 2); }
size_t tc_raw_rows(int H, int W, int chunk) { return (size_t)chunk * (H + 1) * (W + 2); }

static void encode_2d(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, uint64_t inner, uint64_t rows,
                      uint32_t box_inner, uint32_t box_rows, const char* what) {
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {inner * esize};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(tm, dt, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RGFM_CHECK(r == CUDA_SUCCESS, std::string("cuTensorMapEncodeTiled(") + what + ") failed with code " + std::to_string((int)r));
}

size_t tc_stats_floats(int H, int W, int chunk, int CoutP) {
  const int R = (H + 1) * (W + 2);
  const size_t supers = ((size_t)chunk * R + TC_SUPER - 1) / TC_SUPER;
  return supers * 2 * (size_t)stats_slots_for(R) * CoutP * 2;
}

TcPtr make_tc_conv(const TcConvDesc& d, DevPool& pool) {
  const int nmma = (d.Cout + 15) / 16 * 16;
  RGFM_CHECK(nmma >= 16 && nmma <= 256, "tc conv: Cout must be 1..256");
  RGFM_CHECK(d.out == nullptr || (d.CoutP % 64 == 0 && d.CoutP >= nmma && nmma % 32 == 0), "tc conv: raw outputs need Cout % 32 == 0 and a channel stride % 64 == 0");
  RGFM_CHECK(d.Win + 2 <= 39, "tc conv: image width must be <= 37");
  RGFM_CHECK(d.nseg >= 1 && d.nseg <= 3, "tc conv: one to three K segments");
  TcPtr t(new TcConv());
  TcArgs& p = t->p;
  std::memset(&p, 0, sizeof(p));
  const int P = d.Win + 2;
  p.H = d.Hin; p.W = d.Win; p.P = P;
  p.R = (d.Hin + 1) * P;
  p.nseg = d.nseg;
  p.nmma = nmma;
  const int rows_needed = TC_SUPER + 2 * P + 2;
  p.a_box_rows = ((rows_needed + 1) / 2 + 7) / 8 * 8;
  RGFM_CHECK(p.a_box_rows <= 256, "tc conv: A box too tall");
  p.a_slot = 2u * p.a_box_rows * 128u;
  p.b_bytes = (uint32_t)nmma * 128u;
  p.nbuf = (4 * nmma <= 512) ? 2 : 1;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * p.nbuf * nmma)) cols <<= 1;
  p.tmem_cols = cols;
  p.tma_store = (d.out != nullptr) ? 1 : 0;
  p.out_fmt = 2;  // raw activations are stored in fp16 in both 16-bit modes (11-bit mantissa; the epilogue saturates)
  p.stats = d.stats;
  p.nslots = stats_slots_for(p.R);
  RGFM_CHECK(d.stats == nullptr || p.nslots <= 3, "tc conv: GroupNorm partials need images of at least 64 GEMM rows");
  const uint32_t fixed = 2u * p.a_slot + 2u * TC_STAGE_BYTES;
  // weight-ring stage = tps taps of the 3x3 filter (tps in {9, 3, 1}): as many as keep >= 3 stages in the budget
  const uint32_t ring_budget = 196u * 1024u - fixed;
  int tps = 1;
  if (d.seg[0].ks == 3) {
    if (3u * 9u * p.b_bytes <= ring_budget) tps = 9;
    else if (3u * 3u * p.b_bytes <= ring_budget) tps = 3;
  }
  p.tps[0] = tps;
  p.tps[1] = p.tps[2] = 1;
  for (int sg = 1; sg < d.nseg; ++sg)
    if (d.seg[sg].ks == 3) p.tps[sg] = tps;
  p.b_stage = (uint32_t)tps * p.b_bytes;
  int stages = (int)(ring_budget / p.b_stage);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  RGFM_CHECK(stages >= 2, "tc conv: not enough shared memory for the weight ring");
  p.b_stages = stages;
  t->smem = (size_t)fixed + (size_t)stages * p.b_stage + 1024;

  for (int sg = 0; sg < d.nseg; ++sg) {
    const TcSegment& S = d.seg[sg];
    const bool seg_fp16 = d.fp16 || S.raw;
    const CUtensorMapDataType dt16 = seg_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const uint32_t fmt = seg_fp16 ? 0u : 1u;  // UMMA F16F32Format: 0 = F16, 1 = BF16
    p.idesc[sg] = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(nmma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    RGFM_CHECK(S.Cin % 64 == 0 && S.Cin <= 256, "tc conv: operand channels must be a multiple of 64, at most 256");
    RGFM_CHECK(S.ks == 1 || S.ks == 3, "tc conv: 1x1 or 3x3");
    RGFM_CHECK(S.Cin_real >= 1 && S.Cin_real <= S.Cin, "tc conv: real channel count");
    p.cblocks[sg] = (S.Cin_real + 63) / 64;  // blocks that hold no real channel are skipped altogether
    for (int cb = 0; cb < p.cblocks[sg]; ++cb) {
      const int left = S.Cin_real - cb * 64;
      p.ksteps[sg][cb] = left >= 64 ? 4 : (left + 15) / 16;
    }
    p.ntaps[sg] = S.ks * S.ks;
    // a raw activation tensor is indexed by the GEMM row itself: operand slot = raw row + P + 1
    p.row_off[sg] = S.raw ? -(P + 1) : 0;
    if (S.ks == 3) {
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) p.tap_shift[sg][ky * 3 + kx] = ky * P + kx;
    } else {
      p.tap_shift[sg][0] = P + 1;
    }
    // weights: torch [Cout][Cin_real][ks][ks] fp32 (or identity when w_oihw == nullptr) -> [tap][nmma][Cin] 16-bit
    const int taps = p.ntaps[sg];
    std::vector<uint16_t> hw((size_t)taps * nmma * S.Cin, 0);
    for (int o = 0; o < d.Cout; ++o)
      for (int c = 0; c < S.Cin_real; ++c)
        for (int tp = 0; tp < taps; ++tp) {
          const int wct = S.w_cin_total > 0 ? S.w_cin_total : S.Cin_real;
          const float wv = S.w_oihw ? S.w_oihw[((size_t)o * wct + S.w_c0 + c) * taps + tp] : (o == c ? 1.0f : 0.0f);
          uint16_t bits;
          if (seg_fp16) {
            const __half hv = __float2half_rn(wv);
            std::memcpy(&bits, &hv, 2);
          } else {
            const __nv_bfloat16 hv = __float2bfloat16(wv);
            std::memcpy(&bits, &hv, 2);
          }
          hw[((size_t)tp * nmma + o) * S.Cin + c] = bits;
        }
    uint16_t* dw = pool.upload(hw);
    // operand: 2-D [rows][Cin]; rows before the first / after the last are zero-filled by TMA
    encode_2d(&t->tmA[sg], dt16, 2, S.in, (uint64_t)S.Cin,
              S.raw ? tc_raw_rows(d.Hin, d.Win, d.chunk) : tc_operand_slots(d.Hin, d.Win, d.chunk), 64, (uint32_t)p.a_box_rows, "A");
    {
      cuuint64_t dims[3] = {(cuuint64_t)S.Cin, (cuuint64_t)nmma, (cuuint64_t)taps};
      cuuint64_t strides[2] = {(cuuint64_t)S.Cin * 2, (cuuint64_t)nmma * S.Cin * 2};
      cuuint32_t box[3] = {64, (cuuint32_t)nmma, 1};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = encode_fn()(&t->tmB[sg], dt16, 3, dw, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      RGFM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed with code " + std::to_string((int)r));
    }
  }
  for (int sg = d.nseg; sg < 3; ++sg) {
    t->tmA[sg] = t->tmA[0];
    t->tmB[sg] = t->tmB[0];
  }
  if (p.tma_store)
    encode_2d(&t->tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d.out, (uint64_t)d.CoutP, tc_raw_rows(d.Hin, d.Win, d.chunk), 64, 128, "out");
  else
    t->tmO = t->tmA[0];
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
  p.out_H = p.sub ? p.H / 2 : p.H;
  p.out_W = p.sub ? p.W / 2 : p.W;
  if (p.sub) {
    RGFM_CHECK(p.H % 2 == 0 && p.W % 2 == 0, "tc conv: stride 2 needs even image sizes");
    p.tma_store = 0;
  }
  RGFM_CHECK(a.stats == nullptr && a.res == nullptr, "tc conv: statistics and residuals are configured at plan build time");
  if (a.out_mode != OUT_RAW) {
    RGFM_CHECK(p.nmma == 16, "tc conv: the output conv uses a 16-column accumulator");
    p.tma_store = 0;
  } else {
    RGFM_CHECK(a.CoutP % 64 == 0 && a.CoutP >= p.nmma && a.out_padded && a.out_fmt == p.out_fmt && a.out != nullptr,
               "tc conv: raw outputs are 16-bit, padded-linear, channel stride % 64 == 0");
  }
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("RGFM_TC_DBG");
    dbg = e ? atoi(e) : 0;
  }
  p.dbg = dbg;
  const int total_rows = a.nimg * (p.H + 1) * p.P;
  const int supers = (total_rows + TC_SUPER - 1) / TC_SUPER;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    RGFM_CUDA(cudaGetDevice(&dev));
    RGFM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = supers < num_sms ? supers : num_sms;  // persistent: one CTA per SM, static round-robin
  conv_tc_kernel<<<grid, TC_THREADS, tc.smem, s>>>(tc.tmA[0], tc.tmB[0], tc.tmA[1], tc.tmB[1], tc.tmA[2], tc.tmB[2], tc.tmO, p);
  post_launch("conv_tc");
}

}  // namespace rgfm

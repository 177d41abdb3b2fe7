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
  int nseg;             // K segments: [0] the conv proper, [1] the folded 1x1 skip conv / identity residual
  int cblocks[2];       // operand channels / 64 per segment
  int ntaps[2];         // 9 (3x3) or 1 (1x1)
  int tap_shift[2][9];  // row shift of each tap inside the A tile: ky*P + kx (3x3), P+1 (1x1)
  int a_box_rows;       // rows per TMA box; an A tile is two boxes
  uint32_t a_slot;      // bytes of one A slot (2 * a_box_rows * 128)
  uint32_t b_bytes;     // bytes of one weight tap tile (CoutP * 128)
  uint32_t b_stage;     // bytes of one weight-ring stage (taps_per_stage * b_bytes)
  int tps[2];           // taps per weight-ring stage, per segment: barrier round trips are amortised over tps*4 MMAs
  int b_stages;
  int nbuf;             // TMEM double buffering (2) when 4 * CoutP <= 512 columns, else 1
  uint32_t idesc;
  uint32_t tmem_cols;
  int sub;              // stride-2: store rows with even (h, w) only
  int tma_store;        // raw output leaves through swizzled smem staging + TMA store (fully coalesced)
  int out_H, out_W;     // output spatial size (H, W or halved when sub)
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
               const __grid_constant__ CUtensorMap tmO, const TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[2], a_empty[2];
  __shared__ __align__(8) uint64_t b_full[TC_MAX_STAGES], b_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_stage = smem_base + 2u * p.a_slot;                       // epilogue staging (if tma_store)
  const uint32_t smem_b = smem_stage + (p.tma_store ? 2u * TC_STAGE_BYTES : 0u);  // weight ring
  const ConvArgs& a = p.c;
  const int N = a.CoutP;
  const int total_rows = a.nimg * (p.H + 1) * p.P;
  const int supers = (total_rows + TC_SUPER - 1) / TC_SUPER;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB0)) : "memory");
    if (p.nseg > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA1)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB1)) : "memory");
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
        const CUtensorMap* tA = seg ? &tmA1 : &tmA0;
        const CUtensorMap* tB = seg ? &tmB1 : &tmB0;
        const int ncb = p.cblocks[seg], ntp = p.ntaps[seg];
        for (int cb = 0; cb < ncb; ++cb) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          if (elect_one()) {
            if (p.dbg & 2) {
              mbar_arrive(&a_full[as]);
            } else {
              mbar_expect_tx(&a_full[as], p.a_slot);
              const uint32_t sa = smem_base + as * p.a_slot;
              tma_load_2d(sa, tA, &a_full[as], cb * 64, row0);
              tma_load_2d(sa + (p.a_slot >> 1), tA, &a_full[as], cb * 64, row0 + p.a_box_rows);
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
                  tc_mma_bf16(td, da, db, p.idesc, first);
                  tc_mma_bf16_acc(td, da + 2, db + 2, p.idesc);
                  tc_mma_bf16_acc(td, da + 4, db + 4, p.idesc);
                  tc_mma_bf16_acc(td, da + 6, db + 6, p.idesc);
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
    const int r = q * 32 + lane;  // row inside the 128-row unit
    uint32_t sc = 0;              // staging-buffer use counter of this group
    int it = 0;
    for (int st = blockIdx.x; st < supers; st += gridDim.x, ++it) {
      const int buf = (p.nbuf == 2) ? (it & 1) : 0;
      const uint32_t use = (p.nbuf == 2) ? ((uint32_t)it >> 1) : (uint32_t)it;
      const int unit = st * 2 + eg;  // 128-row unit index
      const int m = unit * 128 + r;
      int n, h, w;
      bool valid = slot_pixel(m, p.H, p.P, a.nimg, n, h, w);
      if (p.sub) {
        valid = valid && !(h & 1) && !(w & 1);
        h >>= 1; w >>= 1;
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
        const size_t o = valid ? raw_offset(a.out_padded, n, h, w, p.out_H, p.out_W, a.Cout) : 0;
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + s_bias[c0 + j];
          if (p.tma_store) {
            // registers -> 128B-swizzled staging tile -> one TMA store of [128 rows x 32 channels]
            const uint32_t stg = smem_stage + (uint32_t)eg * TC_STAGE_BYTES;
            if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous store has left the buffer
            if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t dst = stg + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(f[4 * j]), "f"(f[4 * j + 1]),
                           "f"(f[4 * j + 2]), "f"(f[4 * j + 3]) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (et == 0) {
              tma_store_2d(&tmO, stg, c0, unit * 128);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            ++sc;
          } else if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(a.out + o + c0 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[buf]);
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
    if (p.tma_store && et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // smem must outlive the stores
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
  CUtensorMap tmA[2], tmB[2], tmO;
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

TcPtr make_tc_conv(const TcConvDesc& d, DevPool& pool) {
  RGFM_CHECK(d.CoutP % 16 == 0 && d.CoutP >= 16 && d.CoutP <= 256, "tc conv: padded Cout must be 16..256, multiple of 16");
  RGFM_CHECK(d.Win + 2 <= 39, "tc conv: image width must be <= 37");
  RGFM_CHECK(d.nseg == 1 || d.nseg == 2, "tc conv: one or two K segments");
  TcPtr t(new TcConv());
  TcArgs& p = t->p;
  std::memset(&p, 0, sizeof(p));
  const int P = d.Win + 2;
  p.H = d.Hin; p.W = d.Win; p.P = P;
  p.nseg = d.nseg;
  const int rows_needed = TC_SUPER + 2 * P + 2;
  p.a_box_rows = ((rows_needed + 1) / 2 + 7) / 8 * 8;
  RGFM_CHECK(p.a_box_rows <= 256, "tc conv: A box too tall");
  p.a_slot = 2u * p.a_box_rows * 128u;
  p.b_bytes = (uint32_t)d.CoutP * 128u;
  const uint32_t fmt = d.fp16 ? 0u : 1u;  // UMMA F16F32Format: 0 = F16, 1 = BF16
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(d.CoutP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.nbuf = (4 * d.CoutP <= 512) ? 2 : 1;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * p.nbuf * d.CoutP)) cols <<= 1;
  p.tmem_cols = cols;
  p.tma_store = (d.out != nullptr && d.CoutP <= 128 && d.Cout == d.CoutP && d.Cout % 32 == 0) ? 1 : 0;
  const uint32_t fixed = 2u * p.a_slot + (p.tma_store ? 2u * TC_STAGE_BYTES : 0u);
  // weight-ring stage = tps taps of the 3x3 filter (tps in {9, 3, 1}): as many as keep >= 3 stages in the budget
  const uint32_t ring_budget = 198u * 1024u - fixed;
  int tps = 1;
  if (d.seg[0].ks == 3) {
    if (3u * 9u * p.b_bytes <= ring_budget) tps = 9;
    else if (3u * 3u * p.b_bytes <= ring_budget) tps = 3;
  }
  p.tps[0] = tps;
  p.tps[1] = 1;
  p.b_stage = (uint32_t)tps * p.b_bytes;
  int stages = (int)(ring_budget / p.b_stage);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  RGFM_CHECK(stages >= 2, "tc conv: not enough shared memory for the weight ring");
  p.b_stages = stages;
  t->smem = (size_t)fixed + (size_t)stages * p.b_stage + 1024;

  const CUtensorMapDataType dt16 = d.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  for (int sg = 0; sg < d.nseg; ++sg) {
    const TcSegment& S = d.seg[sg];
    RGFM_CHECK(S.Cin % 64 == 0, "tc conv: operand channels must be a multiple of 64");
    RGFM_CHECK(S.ks == 1 || S.ks == 3, "tc conv: 1x1 or 3x3");
    p.cblocks[sg] = S.Cin / 64;
    p.ntaps[sg] = S.ks * S.ks;
    if (S.ks == 3) {
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) p.tap_shift[sg][ky * 3 + kx] = ky * P + kx;
    } else {
      p.tap_shift[sg][0] = P + 1;
    }
    // weights: torch [Cout][Cin_real][ks][ks] fp32 (or identity when w_oihw == nullptr) -> [tap][CoutP][Cin] 16-bit
    const int taps = p.ntaps[sg];
    std::vector<uint16_t> hw((size_t)taps * d.CoutP * S.Cin, 0);
    for (int o = 0; o < d.Cout; ++o)
      for (int c = 0; c < S.Cin_real; ++c)
        for (int tp = 0; tp < taps; ++tp) {
          const float wv = S.w_oihw ? S.w_oihw[((size_t)o * S.Cin_real + c) * taps + tp] : (o == c ? 1.0f : 0.0f);
          uint16_t bits;
          if (d.fp16) {
            const __half hv = __float2half_rn(wv);
            std::memcpy(&bits, &hv, 2);
          } else {
            const __nv_bfloat16 hv = __float2bfloat16(wv);
            std::memcpy(&bits, &hv, 2);
          }
          hw[((size_t)tp * d.CoutP + o) * S.Cin + c] = bits;
        }
    uint16_t* dw = pool.upload(hw);
    // operand: 2-D [slots][Cin]; out-of-range rows of the last tile are zero-filled by TMA
    encode_2d(&t->tmA[sg], dt16, 2, S.in, (uint64_t)S.Cin, tc_operand_slots(d.Hin, d.Win, d.chunk), 64, (uint32_t)p.a_box_rows, "A");
    {
      cuuint64_t dims[3] = {(cuuint64_t)S.Cin, (cuuint64_t)d.CoutP, (cuuint64_t)taps};
      cuuint64_t strides[2] = {(cuuint64_t)S.Cin * 2, (cuuint64_t)d.CoutP * S.Cin * 2};
      cuuint32_t box[3] = {64, (cuuint32_t)d.CoutP, 1};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = encode_fn()(&t->tmB[sg], dt16, 3, dw, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      RGFM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed with code " + std::to_string((int)r));
    }
  }
  if (d.nseg == 1) {
    t->tmA[1] = t->tmA[0];
    t->tmB[1] = t->tmB[0];
  }
  if (p.tma_store)
    encode_2d(&t->tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d.out, (uint64_t)d.Cout, tc_raw_rows(d.Hin, d.Win, d.chunk), 32, 128, "out");
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
  if (p.sub) p.tma_store = 0;
  RGFM_CHECK(a.stats == nullptr && a.res == nullptr, "tc conv: statistics and residuals are handled by pack / the folded K segment");
  if (a.out_mode != OUT_RAW) {
    RGFM_CHECK(a.CoutP == 16, "tc conv: the output conv uses a 16-column accumulator");
    p.tma_store = 0;
  } else {
    RGFM_CHECK(a.CoutP % 32 == 0 && a.Cout == a.CoutP && a.out_padded, "tc conv: raw outputs need Cout % 32 == 0 and the padded-linear layout");
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
  conv_tc_kernel<<<grid, TC_THREADS, tc.smem, s>>>(tc.tmA[0], tc.tmB[0], tc.tmA[1], tc.tmB[1], tc.tmO, p);
  post_launch("conv_tc");
}

}  // namespace rgfm

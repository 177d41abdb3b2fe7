// tcgen05 implicit-GEMM convolution for sm_100a: 16-bit operands staged in shared memory by TMA, fp32 accumulators in
// TMEM, tcgen05.mma issued by one thread per accumulator half, epilogue reading TMEM with tcgen05.ld.
//
// GEMM view of a 3x3 (or 1x1) convolution over padded-linear NHWC tensors:
//   M = 256 GEMM rows per super-tile = two M=128 accumulators that share every weight tile
//   N = Cout rounded up to 16 (<= 256): the accumulator is 128 lanes x N TMEM columns per half
//   K = (segment, 64-channel block, tap, K16 step); K16 steps that hold only zero padding are never issued.
// In the padded-linear layout every tap of the filter is a ROW SHIFT (ky * P + kx, P = W + 2) of ONE shared-memory tile,
// so one 2-box TMA load of 256 + 2P + 2 rows per (super-tile, channel block) serves all nine taps and the conv's zero
// padding is physically in memory (plus TMA's out-of-bounds zero fill before the first / after the last row).  Raw
// activation tensors are themselves valid operands (row offset -(P + 1)): 1x1 skip convs, identity residuals and
// stride-2 convs read them directly.  Stride-2 convolutions are computed at stride 1 and stored subsampled.
//
// Persistent kernel, one CTA per SM, static round-robin over super-tiles, 608 threads: warps 0..7 = two epilogue groups
// (one per 128-row unit; TMEM lane quarter = warp & 3), warp 8 = TMA producer (one thread; it also initialises the barriers), warps
// 9..10 = MMA issuers (one thread each; warp 9 allocates TMEM), warps 11..18 = transform group.  Weights are RESIDENT in shared memory when
// they fit (loaded once per CTA), otherwise they stream through a two-stage ring; A tiles use 2..4 slots; TMEM
// accumulators are double buffered when 4 * N <= 512 columns so the epilogue of tile i overlaps the MMAs of tile i + 1.
// Fused GroupNorm + SiLU: a K segment may be a RAW activation with per-image {scale, shift} coefficients; the transform
// group rewrites the TMA-loaded tile in place (silu(x * scale + shift), one MUFU.TANH per element, swizzled layout) and
// releases it to the MMA threads, so the normalised operand never exists in HBM.
// Epilogue: TMEM -> + bias (+ this step's time bias) -> saturating fp16 -> 128B-swizzled staging tile -> TMA store of
// the raw activation (invalid rows as zeros), GroupNorm partials of the staged tile by ldmatrix.trans + mma.sync; the
// output conv writes the NCHW velocity or applies the fused Euler update.
#include "net.h"
#include <cuda.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace rgfm {

static inline void post_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(std::string(what) + " launch failed: " + cudaGetErrorString(e));
  ++g_launch_counter;
}

constexpr int TC_MAX_STAGES = 12;
// row-dense neighbour exchange as packed fp16 pairs (half the shuffles, one extra fp16 rounding per neighbour term):
// measured < 0.5 % of a step; off = exact fp32 exchange.  Compile time: the kernels carry one exchange loop.
constexpr bool TC_XCHG16 = false;
// REGSTATS kernels: GroupNorm partials from the accumulator registers (channel-quad sums of the thread's row + one 31-shuffle
// transpose-reduce per unit, after the accumulator went back to the MMA warps) instead of ldmatrix.trans + mma.sync on the
// staged tile.  It removes the legacy HMMAs that queue behind the UMMA stream but costs ~160 issue slots per thread and
// unit, so it pays where the epilogue's LATENCY bounds a conv (one operand stage per super-tile: few MMAs, few instructions
// in flight) and loses where the SM is short of issue slots (multi-stage convs; measured 179 -> 196 us, profiles/notes_r02.md).
// Selected per conv by tc_reg_stats().
constexpr int TC_XF_THREADS = 256;        // transform group: fused GroupNorm + SiLU on the A tiles in shared memory (12 warps: no faster)
// two epilogue groups of 4 warps, producer warp, two MMA warps (one per accumulator half), transform group
constexpr int TC_THREADS = 2 * 128 + 96 + TC_XF_THREADS;
constexpr uint32_t TC_STAGE_BYTES = 128 * 128;
// dynamic shared memory for A slots + epilogue staging + weights (static: ~20 KB; trace builds add their stamp table)
#ifdef RGFM_TC_TRACE
constexpr uint32_t TC_SMEM_BUDGET = 205u * 1024u;
#else
constexpr uint32_t TC_SMEM_BUDGET = 206u * 1024u;
#endif

struct TcArgs {
  ConvArgs c;           // epilogue description (bias, time bias, output, nimg, out_mode, dt)
  int H, W, P;          // operand image geometry and padded pitch P = W + 2
  int R;                // GEMM rows per image: (H + 1) * P
  uint32_t magic_P, magic_H1, magic_R;  // ceil(2^32 / d) for d = P, H + 1, R (fast_div)
  int nseg;             // K segments: [0] the conv proper, [1..2] the folded 1x1 skip conv / identity residual
  int cblocks[3];       // operand channels / 64 per segment
  int ksteps[3][4];     // K16 steps that hold real channels in each 64-channel block (all-zero steps are never issued)
  int ntaps[3];         // 9 (3x3) or 1 (1x1)
  int row_off[3];       // TMA row coordinate of tile row 0 relative to the super-tile's first GEMM row (see make_tc_conv)
  int box_rows[3];      // rows per TMA box of the segment: a_box_rows for 3x3 segments, 128 for 1x1 (no halo, aligned)
  int a_box_rows;       // rows per TMA box; an A tile is two boxes
  uint32_t a_slot;      // bytes of one A slot (2 * a_box_rows * 128)
  uint32_t b_bytes;     // bytes of one weight tap tile (nmma * 128)
  uint32_t b_stage;     // bytes of one weight-ring stage (taps_per_stage * b_bytes)
  int tps[3];           // taps per weight-ring stage, per segment: barrier round trips are amortised over tps*4 MMAs
  int b_stages;
  int a_slots;          // A-tile pipeline depth (2..4)
  int resident;         // all weights of the conv fit next to the A slots: loaded ONCE per CTA, no weight ring traffic
  uint32_t b_total;     // bytes of all weight tiles (resident mode)
  int nmma;             // output channels of the conv rounded up to 16
  // Tap stacking: the three kx taps of a filter row read the SAME A rows shifted by 0 / 1 / 2, so
  // they can share one MMA whose B tile stacks their weights on N ([kx * nmma + co][K], which is simply three consecutive
  // tap tiles) and whose accumulator holds nstk column groups D_j; the epilogue forms out[m] = sum_j D_j[m + j] with
  // lane shuffles.  An M=128 MMA costs max(N / 2, ~50) cycles and 4 KB + N * 32 B of shared-memory reads whatever N is,
  // so stacking cuts both for N <= 64 (measured: N=96 57 cycles, N=192 97 cycles against 3 x 49.7).
  //   nstk = 3 (nmma <= 32): one MMA per filter row, N = 3 * nmma
  //   nstk = 2 (nmma <= 64): kx = 0, 1 stacked (N = 2 * nmma) + kx = 2 on its own (N = nmma, A shifted by 2 rows) -- a
  //                          third column group would not leave room for double-buffered accumulators in TMEM
  //   nstk = 1: the classic nine taps
  // A unit yields unit_rows = 128 - (nstk - 1) output rows (the last rows of the M=128 tile lack their D_j[m + j]).
  int nstk;
  // Row-dense mode: the GEMM row space drops the two pad COLUMNS of the padded-linear layout --
  // row m = image_row * Wt + w, image_row = n * (H + 1) + h, Wt = 8 / 16 / 32 >= W -- by loading the tile through a 3-D
  // view [C][P][image rows] of the same tensor (box [64][Wt][rows]).  Wt divides 32, so every warp of the epilogue holds
  // whole image rows: with all three kx taps stacked on N (one MMA per filter row, A read once, N = 3 * nmma) the output is
  //   out[l] = D_0[l - 1] (w > 0) + D_1[l] + D_2[l + 1] (w < Wt - 1)
  // -- intra-warp shuffles with zero fill at the image-row ends (= the conv's own zero padding), no boundary fix-up, all
  // 128 rows of a unit are outputs, and only the pad ROW of each image is wasted (3 / 6 / 11 % at 32 / 16 / 8 pixels
  // against 10 / 20 / 41 % of the padded-linear row space).  ky is a shift of Wt rows of the shared tile (atom aligned).
  int dense;
  // One-unit tiles (row-dense convs with 64 output channels and resident weights): three column groups of 64 are 192 TMEM
  // columns per unit, so the two units of a 256-row super-tile fill TMEM single-buffered and the MMAs of tile i + 1 wait
  // for the epilogue's reads of tile i.  With upt == 1 a tile is ONE unit (128 rows + halo), MMA warp / epilogue group
  // (tile & 1) owns it and TMEM buffer (tile & 1): the accumulators are double buffered again.  The halo costs 1.5x instead
  // of 1.25x operand rows at 32 pixels; the weights must be resident (a ring would stream them per 128 rows).
  // An MMA warp WAITS on the operand stages of the other warp's tiles without using them: a parity wait only tells apart
  // two consecutive phases of a barrier, so every phase of every slot has to be observed by both warps.
  int upt;              // units per tile: 2, or 1
  int stages_total;     // operand stages (64-channel blocks of all segments) per tile
  int reg_stats;        // GroupNorm partials from the accumulator registers (REGSTATS kernels; tc_reg_stats())
  int Wt, wt_shift;     // tile width and its log2
  int irs;              // image rows per super-tile (256 / Wt)
  int w0[3];            // per segment: TMA w coordinate of tile column 0 (0 = raw tensor, 1 = packed operand)
  int ir_off[3];        // per segment: image-row coordinate of the tile's first image row relative to the super-tile's
  uint32_t a_bytes[3];  // per segment: bytes of the A box
  int unit_rows;        // output rows per accumulator half
  int super_rows;       // 2 * unit_rows: output rows per super-tile
  int ncols;            // TMEM columns per accumulator: nstk * nmma rounded up to 32
  uint32_t idesc3[3][3];  // [segment][kx]: descriptor (N = taps in the group * nmma) of the MMA group that starts at tap kx
                          // of a filter row, 0 = the tap is covered by the group before it
  int nbuf;             // TMEM double buffering (2) when 4 * ncols <= 512 columns, else 1
  uint32_t idesc[3];    // per segment, N = nmma (1x1 segments): packed operands use the net's 16-bit format, raw activations are always fp16
  uint32_t tmem_cols;
  int sub;              // stride-2: store rows with even (h, w) only, into the half-resolution raw tensor
  int tma_store;        // raw output leaves through swizzled smem staging + TMA store (fully coalesced)
  int out_H, out_W;     // output spatial size (H, W or halved when sub)
  int rows_alloc;       // rows of the raw output tensor (chunk * R): bound for direct row stores
  int any_xform;        // some segment applies GroupNorm + SiLU to its raw A tiles in shared memory (transform warps active)
  int xform[3];         // per segment
  const float2* coef[3];  // per segment: {scale, shift}[image][coef_C]
  const uint4* coef_h[3]; // per segment: fp16 {scale / 2 x 8, shift / 2 x 8} per (image, octet) = two uint4 (fp16 mode's half2 transform)
  int coef_C[3], coef_c0[3];
  int cin_real[3];      // real channels of the segment (transform touches only those)
  int xf_bf16;          // format the transform writes: 1 = bf16, 0 = fp16 (the net's operand format)
  int exp_no_xform;     // timing experiments only (RGFM_EXPERIMENT_NO_XFORM): transform warps keep the protocol but do no work
  int exp_no_bias;      // timing experiments only (RGFM_EXPERIMENT_NO_BIAS): the epilogue skips its shared-memory bias loads
  int l2_prefetch;      // prefetch the next super-tile's A boxes into L2 (RGFM_TC_L2_PREFETCH, default on)
  int descend;          // walk the super-tiles from the last to the first: every tensor is a little larger than L2, the
                        // packs stream in ascending order, so a conv that starts at the END of its operand finds the rows
                        // the pack wrote last still in L2 (and leaves its own first rows there for the next pack)
  int out_fmt;          // raw output storage: 1 = bf16, 2 = fp16
  unsigned int* sat;    // counter of output rows that hit the fp16 saturation bound (null = not counted)
  float* stats;         // GroupNorm partials [unit][nslots][CoutP][2] of the rounded output (null = none)
  int nslots;
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// All mbarrier wrappers take 32-bit shared-window addresses computed ONCE per thread (bar_base + 8 * index): deriving
// the address from a generic pointer at every use costs a special-register read chain on the single-thread roles.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  // try_wait with a SUSPEND-TIME HINT: the hardware parks the thread until the phase completes (or ~20 us pass) instead
  // of returning after a few cycles, so a waiting role costs no issue slots.  Without the hint the polling loops of
  // the single-thread roles were 17 % of all executed instructions (ncu, profiles/notes_r02.md).
  // Bounded: a protocol bug becomes a trap ("unspecified launch failure") after ~1 s, never a hung GPU.
  // unroll 1: left alone, the compiler unrolls this loop at every call site, and the resulting code size makes the
  // single-warp producer / MMA roles miss in the instruction cache on every step.
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 16); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) return;
  }
  __trap();
}
// (the epilogue's pollers used to back off with nanosleep; the suspend hint does that in hardware with no wake-up delay)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t addr, uint32_t parity) { mbar_wait(addr, parity); }
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// pull a box into L2 only (no shared-memory slot needed): the later tma_load_2d of the same box pays L2 instead of HBM latency
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tm, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
// n / d for n * (d - 1) < 2^32 with magic = ceil(2^32 / d) (checked on the host)
__device__ __forceinline__ int fast_div(int n, uint32_t magic) { return (int)__umulhi((uint32_t)n, magic); }

// two fp32 -> packed fp16, saturated to +-65504 by the conversion itself (F2FP.SATFINITE: one instruction), zeroed for
// rows that are written as zeros (vm).  `seen` keeps the running max |value| of the thread's row in fp32 (the compiler
// folds the chain into three-input FMNMX3 with |.| modifiers: half an instruction per value, where tracking the packed
// halves cost one): a stored row whose max reaches 65504 saturated something, and the epilogue counts it (diag[12]) --
// saturation is never silent.
__device__ __forceinline__ uint32_t cvt2_f16_sat(float a, float b, uint32_t vm, float& seen) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  seen = fmaxf(seen, fmaxf(fabsf(a), fabsf(b)));
  return r & vm;
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& d0, uint32_t& d1, uint32_t& d2, uint32_t& d3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3) : "r"(addr));
}
// D(16x8, fp32) += A(16x16, row) * B(16x8, col) on the legacy tensor path (HMMA: ~16 cycles each on B200 and it shares
// the tensor pipe with tcgen05, so the epilogue spends as few of them as it can)
__device__ __forceinline__ void mma_f16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// packed fp16 pair -> packed fp16 pair of (x / 16)^2: two HMUL2 (the bf16 route through fp32 cost five instructions per
// pair and the epilogue is issue-bound).  The prescale keeps x^2 inside fp16 for |x| < 4096 (beyond that the variance
// becomes inf -- loud, not silent) and the sums are scaled back by 256 when the partials are written; each square
// carries 11 mantissa bits (bf16 had 8) and the HMMA accumulates in fp32.
__device__ __forceinline__ uint32_t sq2_f16_scaled(uint32_t h2) {
  const __half2 k = __half2half2(__ushort_as_half((unsigned short)0x2C00));  // 2^-4
  const __half2 t = __hmul2(*reinterpret_cast<const __half2*>(&h2), k);
  const __half2 q = __hmul2(t, t);
  return *reinterpret_cast<const uint32_t*>(&q);
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

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// decode padded-linear output index m' -> (image, h, w); valid iff a real pixel of an image < nimg
__device__ __forceinline__ bool slot_pixel(int m, int H, int P, uint32_t magic_P, uint32_t magic_H1, int nimg, int& n, int& h, int& w) {
  const int hv = fast_div(m, magic_P);
  w = m - hv * P;
  n = fast_div(hv, magic_H1);
  h = hv - n * (H + 1);
  return (w < P - 2) && (h < H) && (n < nimg);
}


// row-dense decode: m = image_row * Wt + w
__device__ __forceinline__ bool dense_pixel(int m, int H, int W, int Wt, int wt_shift, uint32_t magic_H1, int nimg, int& n, int& h, int& w) {
  const int ir = m >> wt_shift;
  w = m & (Wt - 1);
  n = fast_div(ir, magic_H1);
  h = ir - n * (H + 1);
  return (w < W) && (h < H) && (n < nimg);
}

__device__ __forceinline__ void st_shared_f4x4(float* dst, const uint32_t* v) {
#pragma unroll
  for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dst + 4 * k) = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}
// Sixteen output channels [c, c + 16) of this thread's row from the accumulator's column groups (TcArgs::nstk):
//   out[m] = D_0[m] + D_1[m + 1] + D_2[m + 2]
// Row m + j lives j lanes further along the warp (a rotating shuffle).  For the last nstk - 1 lanes of a warp it is a
// row of the NEXT warp's TMEM lane quarter, which this warp cannot read: those lanes get a wrapped (wrong) value here,
// are stored as zeros by the caller, and are finished by the epilogue's fix-up pass from the raw pieces that the
// boundary lanes of both warps leave in `xw` (this warp's slice of s_xchg, n = nmma floats per vector):
//   nstk = 2: [0] D_0 of lane 31, [1] D_1 of lane 0
//   nstk = 3: [0] D_0 of lane 30, [1] D_0 of lane 31, [2] D_1 of lane 31, [3] D_1 of lane 0, [4] D_2 of lane 0, [5] D_2 of lane 1
// Plain vector stores by one or two lanes: no extra TMEM loads, no selects in the per-element path.
// Row-dense mode: sixteen output channels [c, c + 16) from the three kx column groups, neighbours by shuffle with zero
// fill at the ends of the image row (ml: the row has a left neighbour, mr: a right neighbour; TcArgs::dense).
// x16: the neighbours' partial sums travel as packed fp16 pairs (one SHFL per two channels; SHFL throughput is what bounds
// this loop) and are added to the centre tap's fp32 sum -- each neighbour term carries one extra fp16 rounding, the same
// size as the rounding of the stored activation itself.  The output conv (fp32 velocity) exchanges fp32.
__device__ __forceinline__ void acc16_dense(uint32_t taddr, int c, int n, bool ml, bool mr, bool x16, float* o) {
  uint32_t v0[16], v1[16], v2[16];
  tmem_ld16(taddr + c, v0);
  tmem_ld16(taddr + n + c, v1);
  tmem_ld16(taddr + 2 * n + c, v2);
  tmem_ld_wait();
  if (x16) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t lp, rp;
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lp) : "f"(__uint_as_float(v0[2 * j + 1])), "f"(__uint_as_float(v0[2 * j])));
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(rp) : "f"(__uint_as_float(v2[2 * j + 1])), "f"(__uint_as_float(v2[2 * j])));
      lp = __shfl_up_sync(0xffffffffu, lp, 1);
      rp = __shfl_down_sync(0xffffffffu, rp, 1);
      lp = ml ? lp : 0u;
      rp = mr ? rp : 0u;
      const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lp));
      const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rp));
      o[2 * j] = (__uint_as_float(v1[2 * j]) + lf.x) + rf.x;
      o[2 * j + 1] = (__uint_as_float(v1[2 * j + 1]) + lf.y) + rf.y;
    }
    return;
  }
  // the zero fill at the row ends as PREDICATED adds (no select per value: the loop is issue-bound)
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float l = __uint_as_float(__shfl_up_sync(0xffffffffu, v0[j], 1));
    const float r = __uint_as_float(__shfl_down_sync(0xffffffffu, v2[j], 1));
    float acc = __uint_as_float(v1[j]);
    if (ml) acc += l;
    if (mr) acc += r;
    o[j] = acc;
  }
}

// 32 values per lane -> lane L ends with the sum over the 32 lanes of value L; lanes with keep == false contribute zeros.
// The input array is left intact (the first butterfly writes into g), so a second pass with another mask can follow.
__device__ __forceinline__ float transpose_reduce32_masked(const float* f, bool keep, int lane) {
  float g[16];
  {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float a = keep ? f[i] : 0.f, b = keep ? f[i + 16] : 0.f;
      g[i] = (hi ? b : a) + __shfl_xor_sync(0xffffffffu, hi ? a : b, 16);
    }
  }
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = hi ? g[i] : g[i + off];
      const float kp = hi ? g[i + off] : g[i];
      g[i] = kp + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return g[0];
}

__device__ __forceinline__ void acc16(uint32_t taddr, int c, int nstk, int n, int lane, float* xw, float* o) {
  uint32_t v0[16];
  tmem_ld16(taddr + c, v0);
  if (nstk == 1) {
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v0[j]);
    return;
  }
  uint32_t v1[16];
  tmem_ld16(taddr + n + c, v1);
  if (nstk == 2) {
    tmem_ld_wait();
    if (lane == 31) st_shared_f4x4(xw + c, v0);
    if (lane == 0) st_shared_f4x4(xw + n + c, v1);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      o[j] = __uint_as_float(v0[j]) + __uint_as_float(__shfl_sync(0xffffffffu, v1[j], (lane + 1) & 31));
    return;
  }
  uint32_t v2[16];
  tmem_ld16(taddr + 2 * n + c, v2);
  tmem_ld_wait();
  if (lane >= 30) st_shared_f4x4(xw + (lane - 30) * n + c, v0);
  if (lane == 31) st_shared_f4x4(xw + 2 * n + c, v1);
  if (lane == 0) {
    st_shared_f4x4(xw + 3 * n + c, v1);
    st_shared_f4x4(xw + 4 * n + c, v2);
  }
  if (lane == 1) st_shared_f4x4(xw + 5 * n + c, v2);
#pragma unroll
  for (int j = 0; j < 16; ++j)
    o[j] = (__uint_as_float(v0[j]) + __uint_as_float(__shfl_sync(0xffffffffu, v1[j], (lane + 1) & 31))) +
           __uint_as_float(__shfl_sync(0xffffffffu, v2[j], (lane + 2) & 31));
}

// The same for 16 values (convs with at most 32 output channels): 16 shuffles, lane L ends with the total of value L >> 1.
__device__ __forceinline__ float transpose_reduce16_masked(const float* f, bool keep, int lane) {
  float g[8];
  {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = keep ? f[i] : 0.f, b = keep ? f[i + 8] : 0.f;
      g[i] = (hi ? b : a) + __shfl_xor_sync(0xffffffffu, hi ? a : b, 16);
    }
  }
#pragma unroll
  for (int off = 8; off >= 2; off >>= 1) {
    const int half = off >> 1;  // values kept: 4, 2, 1
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = hi ? g[i] : g[i + half];
      const float kp = hi ? g[i + half] : g[i];
      g[i] = kp + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return g[0] + __shfl_xor_sync(0xffffffffu, g[0], 1);
}

// -DRGFM_TC_TRACE: block 0 records clock64 stamps of the first 24 super-tiles per role and prints them (debug builds only)
#ifdef RGFM_TC_TRACE
#define TRC(role, idx) do { if (blockIdx.x == 0 && (idx) < 12) s_trace[role][idx] = clock64() - t_start; } while (0)
#else
#define TRC(role, idx) do { } while (0)
#endif

// DENSE: the row-dense instantiation (TcArgs::dense) -- its own register allocation, none of the padded-linear stacking code
// XBF16: format the transform group writes (bf16 mode: fp32 arithmetic; fp16 mode: packed half2) -- compile time, so that a
// kernel carries one transform loop: code size is felt by the single-thread roles, whose instruction-cache misses nobody shares
// FINAL: the output conv (NCHW velocity / fused Euler update / scaled gradient) -- its kernels carry no raw-activation epilogue
// STACKED: padded-linear units that stack kx taps (lane-shifted sums, boundary fix-up pass); nstk == 1 kernels carry none of it
template <bool DENSE, bool REGSTATS, bool XBF16, bool FINAL, bool STACKED>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmZ, const TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bars[16 + 2 * TC_MAX_STAGES];
  const uint32_t bar_base = smem_u32(s_bars);
#define a_full(i) (bar_base + 8u * (uint32_t)(i))
#define a_empty(i) (bar_base + 8u * (4u + (uint32_t)(i)))
#define tmem_full_bar(i) (bar_base + 8u * (8u + (uint32_t)(i)))
#define tmem_empty_bar(i) (bar_base + 8u * (10u + (uint32_t)(i)))
// a_ready(slot): transform group -> MMA warps (convs with a fused GroupNorm + SiLU segment)
#define a_ready(i) (bar_base + 8u * (12u + (uint32_t)(i)))
#define b_full(i) (bar_base + 8u * (16u + (uint32_t)(i)))
#define b_empty(i) (bar_base + 8u * (16u + TC_MAX_STAGES + (uint32_t)(i)))
  __shared__ uint32_t tmem_slot;
#ifdef RGFM_TC_TRACE
  __shared__ long long s_trace[8][12];
  __shared__ long long s_tp[2];  // prologue end (after the block-wide barrier), resident weights landed
  const long long t_start = clock64();
#endif
  __shared__ uint8_t s_rown[400];  // transform group: image (relative to the tile's first) of each A-tile row, 0xff = not a pixel
  __shared__ __align__(16) float s_bias[256];
  // tap stacking: raw pieces of the rows that straddle two warps' lane quarters (acc16); row-dense mode: 4 KB of zeros, the
  // source of the TMA store that writes the output's two pad columns
  __shared__ __align__(128) float s_xchg[2][4][192];
  __shared__ __align__(16) float s_part[2][4][3][64][2];
  // per epilogue group / warp / image slot: {sum, sumsq} of a 64-channel chunk

  pdl_trigger();  // the next kernel of the stream may be scheduled as this grid's CTAs retire
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_stage = smem_base + (uint32_t)p.a_slots * p.a_slot;  // epilogue staging tiles (one per epilogue group)
  const uint32_t smem_b = smem_stage + 2u * TC_STAGE_BYTES;    // weights (resident or ring)
  const ConvArgs& a = p.c;
  const int N = p.nmma;
  const int total_rows = a.nimg * (p.H + 1) * (DENSE ? p.Wt : p.P);
  const int SR = p.super_rows, UR = p.unit_rows;
  const int supers = (total_rows + SR - 1) / SR;

  // Barriers are initialised by the PRODUCER's own thread (warp 8, lane 0), which then leaves the block-wide barrier below
  // with a non-blocking arrive and issues the launch's first operand load at once: the other 18 warps reach that barrier
  // ~1350 cycles into the kernel and the producer's cold start (instruction and constant-cache misses of a role one
  // thread walks) used to begin only there -- first TMA at ~3000 cycles (trace build; profiles/notes_r02.md section 15).
  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB0)) : "memory");
    for (int s = 0; s < 4; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), p.upt == 1 ? 1 : 2);   // both MMA warps (one-unit tiles: the tile's own MMA warp)
      mbar_init(a_ready(s), 1);  // transform group -> MMA warps
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tmem_full_bar(s), p.upt == 1 ? 1 : 2);   // both MMA warps (one-unit tiles: per MMA warp / epilogue group)
      mbar_init(tmem_empty_bar(s), p.upt == 1 ? 4 : 8);  // one arrival per epilogue warp
    }
    for (int s = 0; s < (p.resident ? 1 : p.b_stages); ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    if (p.nseg > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA1)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB1)) : "memory");
    }
    if (p.nseg > 2) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA2)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB2)) : "memory");
    }
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
    if (p.tma_store && DENSE) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmZ)) : "memory");
  }
  // Warp roles.  The hardware arbiter favours HIGHER warp ids, so the latency-critical single-thread roles sit at the
  // top: warps 0..7 = two epilogue groups (TMEM lane quarter = warp & 3), warp 8 = TMA producer (+ barrier init),
  // warps 9, 10 = MMA issuers of the two accumulator halves.
  if (warp == 9) {  // TMEM allocator = the first MMA warp (the producer has its first load to issue)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // everything above (barrier init, TMEM allocation, tensor-map prefetch) overlaps the previous kernel's tail; nothing
  // below may run before its writes are visible
  if (DENSE && warp < 8) {  // the zero tile of the pad-column store (before the barrier below; generic -> async proxy fence)
    for (int i = threadIdx.x; i < 2 * 4 * 192; i += 256) (&s_xchg[0][0][0])[i] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  pdl_wait();
  tc_fence_before();
  // the producer warp arrives without waiting (its barriers are its own work and it never touches TMEM); everybody else
  // needs the barriers and the TMEM address published
  // (named barrier 5, not barrier 0: the kernel's closing __syncthreads() must never be able to count towards this one)
  if (warp == 8) asm volatile("bar.arrive 5, %0;" ::"n"(TC_THREADS) : "memory");
  else asm volatile("bar.sync 5, %0;" ::"n"(TC_THREADS) : "memory");
  tc_fence_after();
  const uint32_t tmem_d = warp == 8 ? 0u : tmem_slot;  // (the producer left the barrier early and never touches TMEM)
#ifdef RGFM_TC_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) s_tp[0] = clock64() - t_start;
#endif
  if (warp < 8) {
    // bias + this step's time bias: identical for every tile of the launch (t is uniform over the batch).  AFTER the
    // block-wide barrier, among the epilogue warps only: the chain step -> time-bias row -> value is ~1500 cycles of
    // dependent global loads, and the producer's first operand load used to wait for it at the barrier (the first
    // accumulator is thousands of cycles away: the epilogue has the time).
    const float* tb = a.temb ? a.temb + (size_t)(a.step ? *a.step : 0) * a.temb_stride + a.temb_off : nullptr;
    for (int c = threadIdx.x; c < N; c += 256) {
      float b = a.bias[c];
      if (tb && c < a.Cout) b += tb[c];
      s_bias[c] = b;
    }
    asm volatile("bar.sync 4, 256;" ::: "memory");
  }

  if (warp == 8) {
    // ------------------------------- TMA producer -------------------------------
    // ONE thread runs the whole loop (waits included).  Ring positions are (index, phase) counters: no divisions.
    // The single-thread roles keep their loops rolled and tiny on purpose: each is walked by one warp only, so every
    // instruction-cache line it touches is a miss nobody else amortises (a 64 KB kernel cost ~480 cycles per
    // pipeline stage against ~120 for the barrier operations themselves).
    if (lane == 0) {
      uint32_t as = 0, aph = 0, bs = 0, bph = 0;
      const uint32_t a_slots = (uint32_t)p.a_slots;
      // resident weights (every weight tile of the conv, once, (segment, channel block, tap) order) are requested right
      // AFTER the launch's first operand tile: the transform group needs that tile long before the MMA warps need weights
      bool want_w = p.resident != 0;
#pragma unroll 1
      for (int sti = blockIdx.x; sti < supers; sti += gridDim.x) {
        const int st = p.descend ? supers - 1 - sti : sti;  // see TcArgs::descend
        const int row0 = st * SR;
#pragma unroll 1
        for (int seg = 0; seg < p.nseg; ++seg) {
          const CUtensorMap* tA = seg == 0 ? &tmA0 : (seg == 1 ? &tmA1 : &tmA2);
          const CUtensorMap* tB = seg == 0 ? &tmB0 : (seg == 1 ? &tmB1 : &tmB2);
          const int ncb = p.cblocks[seg], ntp = p.ntaps[seg];
          const int tps = p.tps[seg];
          const int ra = row0 + p.row_off[seg];  // negative / past-the-end rows are zero-filled by TMA
#pragma unroll 1
          for (int cb = 0; cb < ncb; ++cb) {
            mbar_wait(a_empty(as), aph ^ 1u);
            const uint32_t half_bytes = (uint32_t)p.box_rows[seg] * 128u;
            const uint32_t sa = smem_base + as * p.a_slot;
            if (DENSE) {  // one 3-D box [64 channels][Wt pixels][image rows] of the padded-linear tensor: no pad columns
              mbar_expect_tx(a_full(as), p.a_bytes[seg]);
              tma_load_3d(sa, tA, a_full(as), cb * 64, p.w0[seg], st * p.irs + p.ir_off[seg]);
              if (p.l2_prefetch && sti + (int)gridDim.x < supers) {
                const int st2 = p.descend ? supers - 1 - (sti + (int)gridDim.x) : sti + (int)gridDim.x;
                tma_prefetch_3d(tA, cb * 64, p.w0[seg], st2 * p.irs + p.ir_off[seg]);
              }
            } else {
            mbar_expect_tx(a_full(as), 2u * half_bytes);
            tma_load_2d(sa, tA, a_full(as), cb * 64, ra);
            tma_load_2d(sa + half_bytes, tA, a_full(as), cb * 64, ra + p.box_rows[seg]);
            }
            if (!DENSE && p.l2_prefetch && sti + (int)gridDim.x < supers) {
              // The same tile of this CTA's NEXT super-tile, into L2 now: with two A slots the load -> (transform) -> MMA
              // chain of a slot is serial, and the ~1.9 us HBM latency of the load was the largest term in it.
              const int st2 = p.descend ? supers - 1 - (sti + (int)gridDim.x) : sti + (int)gridDim.x;
              const int ra2 = st2 * SR + p.row_off[seg];
              tma_prefetch_2d(tA, cb * 64, ra2);
              tma_prefetch_2d(tA, cb * 64, ra2 + p.box_rows[seg]);
            }
            if (seg == 0 && cb == 0) TRC(0, (sti - (int)blockIdx.x) / (int)gridDim.x);
            if (++as == a_slots) {
              as = 0u;
              aph ^= 1u;
            }
            if (want_w) {
              want_w = false;
              mbar_expect_tx(b_full(0), p.b_total);
              uint32_t dst = smem_b;
#pragma unroll 1
              for (int sg = 0; sg < p.nseg; ++sg) {
                const CUtensorMap* tW = sg == 0 ? &tmB0 : (sg == 1 ? &tmB1 : &tmB2);
#pragma unroll 1
                for (int c2 = 0; c2 < p.cblocks[sg]; ++c2)
#pragma unroll 1
                  for (int tap = 0; tap < p.ntaps[sg]; ++tap, dst += p.b_bytes) tma_load_3d(dst, tW, b_full(0), c2 * 64, 0, tap);
              }
            }
            if (p.resident) continue;
#pragma unroll 1
            for (int tap = 0; tap < ntp; tap += tps) {
              mbar_wait(b_empty(bs), bph ^ 1u);
              mbar_expect_tx(b_full(bs), (uint32_t)tps * p.b_bytes);
              uint32_t dst = smem_b + bs * p.b_stage;
#pragma unroll 1
              for (int j = 0; j < tps; ++j, dst += p.b_bytes) tma_load_3d(dst, tB, b_full(bs), cb * 64, 0, tap + j);
              if (++bs == (uint32_t)p.b_stages) {
                bs = 0u;
                bph ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ------------------- MMA issuers: warp 9 accumulates rows 0..127, warp 10 rows 128..255 of each super-tile -------------------
    // One thread per warp runs the role.  A tap is a ROW SHIFT of the shared A tile: descriptor start + (ky * P + kx)
    // rows * 8 sixteen-byte units; K16 step k adds 2 units to both descriptors; tap j of a ring stage sits j * b_tap16
    // units into the stage.  All of it is register arithmetic in a rolled loop (~8 instructions per MMA).
    const uint32_t hf = (uint32_t)(warp - 9);
    if (lane == 0) {
      uint32_t as = 0, aph = 0, bs = 0, bph = 0;
      // 16-byte units: a filter row is P rows of the padded-linear tile, or Wt rows of the row-dense one (whose kx taps
      // never shift the A descriptor: they are column groups of one MMA)
      const uint32_t rowstep = DENSE ? (uint32_t)p.Wt * 8u : (uint32_t)p.P * 8u - 16u;
      const uint32_t kxstep = DENSE ? 0u : 8u;
      const uint32_t td_1x1 = DENSE ? (uint32_t)p.nmma : 0u;  // row-dense: a 1x1 segment adds to the centre column group
      const uint32_t b_tap16 = p.b_bytes >> 4, b_stage16 = p.b_stage >> 4;
      const uint32_t b_stages = (uint32_t)p.b_stages;
      const uint64_t dhi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;  // SBO = 1024 B, version 1, SWIZZLE_128B
      const uint32_t a_slots = (uint32_t)p.a_slots;
      const bool resident = p.resident != 0;
      if (resident) {
        mbar_wait(b_full(0), 0u);
        tc_fence_after();
#ifdef RGFM_TC_TRACE
        if (blockIdx.x == 0 && hf == 0) s_tp[1] = clock64() - t_start;
#endif
      }
      const bool upt1 = p.upt == 1;
      const uint32_t a_unit = upt1 ? 0u : hf * ((uint32_t)UR * 128u);  // this warp's rows inside the A tile
      int it = 0;
#pragma unroll 1
      for (int st = blockIdx.x; st < supers; st += gridDim.x, ++it) {
        if (upt1 && ((uint32_t)it & 1u) != hf) {
          // The other MMA warp's tile: step over its operand stages -- but OBSERVE each of them.  A parity wait only tells
          // apart two consecutive phases of a barrier: a warp that skipped a slot's first fill and then waited for its
          // second one could pass while the first is still in flight (seen as a deadlock with the slower bf16 transform).
          for (int k = 0; k < p.stages_total; ++k) {
            mbar_wait(p.any_xform ? a_ready(as) : a_full(as), aph);
            if (++as == a_slots) {
              as = 0u;
              aph ^= 1u;
            }
          }
          continue;
        }
        const int buf = (!upt1 && p.nbuf == 2) ? (it & 1) : 0;
        const uint32_t use = (upt1 || p.nbuf == 2) ? ((uint32_t)it >> 1) : (uint32_t)it;
        const uint32_t tb = upt1 ? hf : (uint32_t)buf;  // barrier pair: per buffer, or per MMA warp with one-unit tiles
        mbar_wait(tmem_empty_bar(tb), (use & 1u) ^ 1u);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t td = tmem_d + (uint32_t)((buf * 2 + (int)hf) * p.ncols);
        uint32_t acc = 0;  // the first MMA of the super-tile overwrites the accumulator
        uint32_t b_res = smem_b >> 4;  // resident mode: running position in the (segment, channel block, tap) order
#pragma unroll 1
        for (int seg = 0; seg < p.nseg; ++seg) {
          const int ncb = p.cblocks[seg], ntp = p.ntaps[seg];
          const int tps = resident ? ntp : p.tps[seg];
          const uint32_t idesc = p.idesc[seg];
          // descriptors of the MMA groups that start at tap kx = 0, 1, 2 of a filter row (0 = covered by the group before);
          // a 1x1 segment is one group of one tap
          const uint32_t id0 = ntp == 1 ? idesc : p.idesc3[seg][0], id1 = p.idesc3[seg][1], id2 = p.idesc3[seg][2];
#pragma unroll 1
          for (int cb = 0; cb < ncb; ++cb) {
            mbar_wait(p.any_xform ? a_ready(as) : a_full(as), aph);
            tc_fence_after();
            if (hf == 0 && seg == 0 && cb == 0) TRC(1, it);
            const int ks = p.ksteps[seg][cb];
            // descriptor low word of the current filter row (a 1x1 segment's tile starts at the super-tile's own first row)
            uint32_t a_lo = (smem_base + as * p.a_slot + a_unit) >> 4;
            uint32_t kx = 0;
#pragma unroll 1
            for (int tap = 0; tap < ntp; tap += tps) {
              uint32_t b_lo = b_res;
              if (!resident) {
                mbar_wait(b_full(bs), bph);
                tc_fence_after();
                b_lo = (smem_b >> 4) + bs * b_stage16;
              }
#pragma unroll 1
              for (int j = 0; j < tps; ++j) {
                // the MMA group that STARTS at this tap (N = taps in the group * nmma: their weight tiles are consecutive in
                // shared memory); nothing to issue when the tap belongs to the previous group
                const uint32_t idg = kx == 0u ? id0 : (kx == 1u ? id1 : id2);
                if (idg != 0u) {
                  const uint32_t tdd = ntp == 1 ? td + td_1x1 : td;
#pragma unroll 1
                  for (int k = 0; k < ks; ++k) {
                    tc_mma_bf16(tdd, dhi | (uint64_t)(a_lo + 2u * k), dhi | (uint64_t)(b_lo + 2u * k), idg, acc);
                    acc = 1u;
                  }
                }
                b_lo += b_tap16;
                if (++kx == 3u) {  // next filter row
                  kx = 0u;
                  a_lo += rowstep;
                } else {
                  a_lo += kxstep;
                }
              }
              if (resident) {
                b_res = b_lo;
              } else {
                tc_commit(b_empty(bs));  // the ring stage is free once these MMAs have read it
                if (++bs == b_stages) {
                  bs = 0u;
                  bph ^= 1u;
                }
              }
            }
            tc_commit(a_empty(as));
            if (++as == a_slots) {
              as = 0u;
              aph ^= 1u;
            }
          }
        }
        tc_commit(tmem_full_bar(tb));
        if (hf == 0) TRC(2, it);
      }
    }
  } else if (warp >= 11) {
    // ---------------- transform group: fused GroupNorm + SiLU on the raw A tiles ----------------
    // The tile was loaded by TMA from a RAW activation tensor; these 256 threads rewrite its real pixels in place as
    // silu(x * scale + shift) (per image and channel, coefficients from gn_coef_kernel) in the swizzled layout the MMA
    // reads, then hand the slot to the MMA threads.  Pad rows / columns are zeros in the raw tensor (and out-of-range
    // rows are zero-filled by TMA), so they are simply left alone.  Segments without a transform are passed through.
    // Measured alternatives (profiles/notes_r01.md): reading the raw rows with plain global loads instead of TMA, and
    // splitting the tile into three boxes with their own barriers so that the first accumulator half starts early,
    // were both slower; a single warp of this kernel runs at ~0.2 IPC, hence eight of them.
    if (p.any_xform) {
      const int tt = threadIdx.x - 11 * 32;
      const int rows_used = DENSE ? (p.irs + 2) * p.Wt : SR + 2 * p.P + 2;
      [[maybe_unused]] const bool wcheck = DENSE && p.Wt != p.W;
      uint32_t as = 0, aph = 0;
      const uint32_t a_slots = (uint32_t)p.a_slots;
#pragma unroll 1
      for (int sti = blockIdx.x; sti < supers; sti += gridDim.x) {
        const int st = p.descend ? supers - 1 - sti : sti;
        // raw row of tile row 0 (every raw 3x3 segment has this offset): one image row (dense) / P + 1 rows before the super-tile
        const int rbase = DENSE ? (st * p.irs - 1) * p.Wt : st * SR - (p.P + 1);
        const int nbase = fast_div(rbase < 0 ? 0 : rbase, p.magic_R);
        // row table of the super-tile, shared by all its (segment, channel block) tiles: one division chain per row here
        // instead of one per row and tile in the hot loop
        // (the previous super-tile's readers are behind the closing barrier of its last tile; the opening barrier of the
        // first tile below publishes the new table)
        if constexpr (DENSE) {  // one entry per image row of the tile (at most 34)
          if (tt < p.irs + 2) {
            const int ir = st * p.irs - 1 + tt;
            const int n = fast_div(ir < 0 ? 0 : ir, p.magic_H1);
            const bool ok = ir >= 0 && ir - n * (p.H + 1) < p.H && n < a.nimg;
            s_rown[tt] = ok ? (uint8_t)(n - nbase) : (uint8_t)0xff;
          }
        } else {
        for (int r = tt; r < 400; r += TC_XF_THREADS) {
          int n = -1, h, w;
          const int R = rbase + r;
          const bool ok = r < rows_used && R >= 0 && slot_pixel(R, p.H, p.P, p.magic_P, p.magic_H1, a.nimg, n, h, w);
          s_rown[r] = ok ? (uint8_t)(n - nbase) : (uint8_t)0xff;
        }
        }
#pragma unroll 1
        for (int seg = 0; seg < p.nseg; ++seg) {
          const int ncb = p.cblocks[seg];
#pragma unroll 1
          for (int cb = 0; cb < ncb; ++cb) {
            if (!p.xform[seg]) {  // a TMA tile used as it is (1x1 skip / identity segment): one thread passes it on
              if (tt == 0) {
                mbar_wait(a_full(as), aph);
                mbar_arrive(a_ready(as));
              }
              if (++as == a_slots) {
                as = 0u;
                aph ^= 1u;
              }
              continue;
            }
            // a 32-channel block uses 4 octets x 64 row lanes so that every thread works
            const int left = p.cin_real[seg] - cb * 64;
            const int ol = left == 32 ? 2 : 3;
            const int oct = tt & ((1 << ol) - 1), rl = tt >> ol, rstride = TC_XF_THREADS >> ol;
            const bool work = oct * 8 < left && p.exp_no_xform != 1;
            const float2* cf = p.coef[seg] + (size_t)nbase * p.coef_C[seg] + p.coef_c0[seg] + cb * 64 + oct * 8;
            const int cC = p.coef_C[seg];
            const uint32_t sa = smem_base + as * p.a_slot;
            constexpr bool to_bf16 = XBF16;
#ifdef RGFM_EXPERIMENTS  // accuracy experiment (exp build only): fp16 mode with the fp32-arithmetic transform (one rounding per operand)
            const bool xf_fp32 = p.exp_no_xform == 3;
#else
            constexpr bool xf_fp32 = false;
#endif
            if constexpr (DENSE) {
              if (!to_bf16 && !xf_fp32) {
                // ---- row-dense tile, fp16 mode: the lean loop ----
                // rstride (32 or 64 rows) is a multiple of the tile width, so a thread's rows all have the same pixel
                // column w (threads of columns >= W idle) and its k-th row lies in tile image row a0 + k * da.  The rows of
                // one image are one contiguous k range: loop over the tile's images (2..5), load that image's coefficients
                // once, and run a branch-free inner loop over its rows -- no row table, no per-row validity / image tests
                // (pad rows fall between the ranges).  ~25 instead of ~67 thread-instructions per 8-channel item, in a
                // kernel that is short of issue slots (profiles/notes_r02.md).
                const int sh = p.wt_shift;
                const int a0 = rl >> sh, da = rstride >> sh;
                const int dsh = 31 - __clz(da);
                const bool mine = work && (rl & (p.Wt - 1)) < p.W;
                const int H1 = p.H + 1;
                const int ir0 = st * p.irs - 1;
                const int ir_lo = ir0 < 0 ? 0 : ir0;
                const int ir_hi = min(ir0 + p.irs + 2, a.nimg * H1);
                const int ostride = (cC >> 3) * 2;  // uint4 per image
                const uint4* ch = p.coef_h[seg] + (size_t)(((p.coef_c0[seg] + cb * 64) >> 3) + oct) * 2;
                int n = nbase;
                uint4 ns = make_uint4(0u, 0u, 0u, 0u), nb = ns;
                if (mine && n * H1 < ir_hi) {  // first image: in flight while the tile is still loading
                  ns = __ldg(ch + (size_t)n * ostride);
                  nb = __ldg(ch + (size_t)n * ostride + 1);
                }
                if (tt == 0) mbar_wait(a_full(as), aph);
                asm volatile("bar.sync 3, %0;" ::"n"(TC_XF_THREADS) : "memory");
                if (tt == 0 && seg == 0 && cb == 0) TRC(4, (sti - (int)blockIdx.x) / (int)gridDim.x);
                if (mine) {
                  const uint32_t abase = sa + (uint32_t)rl * 128u + (uint32_t)((oct ^ (rl & 7)) << 4);
                  const uint32_t astep = (uint32_t)rstride * 128u;
#pragma unroll 1
                  for (; n * H1 < ir_hi; ++n) {
                    const uint4 hs = ns, hb = nb;
                    if ((n + 1) * H1 < ir_hi) {  // the next image's coefficients, one image ahead
                      ns = __ldg(ch + (size_t)(n + 1) * ostride);
                      nb = __ldg(ch + (size_t)(n + 1) * ostride + 1);
                    }
                    const int lo = max(n * H1, ir_lo) - ir0, hi = min(n * H1 + p.H, ir_hi) - ir0;  // tile image rows of image n
                    const int k_lo = lo <= a0 ? 0 : (lo - a0 + da - 1) >> dsh;
                    const int k_hi = hi <= a0 ? 0 : (hi - a0 + da - 1) >> dsh;
                    const uint32_t hsv[4] = {hs.x, hs.y, hs.z, hs.w}, hbv[4] = {hb.x, hb.y, hb.z, hb.w};
#pragma unroll 1
                    for (int k = k_lo; k < k_hi; k += 4) {
                      uint32_t w4[4][4];
                      const uint32_t ad = abase + (uint32_t)k * astep;
#pragma unroll
                      for (int j = 0; j < 4; ++j)  // all loads in flight before the math
                        if (k + j < k_hi)
                          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                       : "=r"(w4[j][0]), "=r"(w4[j][1]), "=r"(w4[j][2]), "=r"(w4[j][3]) : "r"(ad + (uint32_t)j * astep));
#pragma unroll
                      for (int j = 0; j < 4; ++j) {
                        if (k + j >= k_hi) break;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                          uint32_t z, t;
                          asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(z) : "r"(w4[j][q]), "r"(hsv[q]), "r"(hbv[q]));
                          asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(z));
                          asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(w4[j][q]) : "r"(z), "r"(t));
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ad + (uint32_t)j * astep), "r"(w4[j][0]), "r"(w4[j][1]),
                                     "r"(w4[j][2]), "r"(w4[j][3]) : "memory");
                      }
                    }
                  }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core (async proxy) reads
                asm volatile("bar.sync 3, %0;" ::"n"(TC_XF_THREADS) : "memory");
                if (tt == 0) mbar_arrive(a_ready(as));
                if (tt == 0 && seg == 0 && cb == 0) TRC(5, (sti - (int)blockIdx.x) / (int)gridDim.x);
                if (++as == a_slots) {
                  as = 0u;
                  aph ^= 1u;
                }
                continue;
              }
            }
            if (!to_bf16 && !xf_fp32) {
              // ---- fp16 mode: packed half2 arithmetic, four rows in flight per thread ----
              // silu(z) = hz + hz * tanh(hz), hz = x * (scale / 2) + shift / 2: per PAIR of channels one HFMA2, one
              // MUFU.TANH (f16x2) and one HFMA2, no conversions (the fp32 path below spends ~4.5 instructions per element;
              // the kernel is issue-bound, profiles/notes_r02.md).  The operand is rounded to fp16 either way; the extra
              // fp16 roundings of hz and tanh are inside the 16-bit tolerance (tests/test_gpu_parity.py).
              const int ostride = (cC >> 3) * 2;  // uint4 per image
              const uint4* ch = p.coef_h[seg] + (size_t)nbase * ostride + (size_t)(((p.coef_c0[seg] + cb * 64) >> 3) + oct) * 2;
              uint32_t cur = 0xffu, nxt = 0xffu;
              uint4 hs = make_uint4(0u, 0u, 0u, 0u), hb = hs, ns = hs, nb = hs;
              if (work) {
                cur = 0u;
                hs = __ldg(ch);
                hb = __ldg(ch + 1);
                if (nbase + 1 < a.nimg) {
                  nxt = 1u;
                  ns = __ldg(ch + ostride);
                  nb = __ldg(ch + ostride + 1);
                }
              }
              if (tt == 0) mbar_wait(a_full(as), aph);
              asm volatile("bar.sync 3, %0;" ::"n"(TC_XF_THREADS) : "memory");
              if (tt == 0 && seg == 0 && cb == 0) TRC(4, (sti - (int)blockIdx.x) / (int)gridDim.x);
              if (work) {
#pragma unroll 1
                for (int r0 = rl; r0 < rows_used; r0 += 4 * rstride) {
                  uint32_t w4[4][4], addr[4], nr[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {  // all loads in flight before the math
                    const int r = r0 + j * rstride;
                    if constexpr (DENSE) {
                      nr[j] = r < rows_used ? (uint32_t)s_rown[r >> p.wt_shift] : 0xffu;
                      if (wcheck && (r & (p.Wt - 1)) >= p.W) nr[j] = 0xffu;  // tile wider than the image: pad / out-of-bounds pixels
                    } else {
                      nr[j] = r < rows_used ? (uint32_t)s_rown[r] : 0xffu;
                    }
                    addr[j] = sa + (uint32_t)r * 128u + (uint32_t)((oct ^ (r & 7)) << 4);
                    if (nr[j] != 0xffu)
                      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                   : "=r"(w4[j][0]), "=r"(w4[j][1]), "=r"(w4[j][2]), "=r"(w4[j][3]) : "r"(addr[j]));
                  }
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    if (nr[j] == 0xffu) continue;  // not a pixel: the raw row is zero already (conv padding)
                    if (nr[j] != cur) {
                      cur = nr[j];
                      if (cur != nxt) {  // not the prefetched image (a thread whose rows skip one): load it now
                        ns = __ldg(ch + (size_t)cur * ostride);
                        nb = __ldg(ch + (size_t)cur * ostride + 1);
                      }
                      hs = ns;
                      hb = nb;
                      nxt = 0xffu;
                      if (nbase + (int)cur + 1 < a.nimg) {  // and request the one after it
                        nxt = cur + 1u;
                        ns = __ldg(ch + (size_t)nxt * ostride);
                        nb = __ldg(ch + (size_t)nxt * ostride + 1);
                      }
                    }
                    const uint32_t hsv[4] = {hs.x, hs.y, hs.z, hs.w}, hbv[4] = {hb.x, hb.y, hb.z, hb.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      uint32_t z, t;
                      asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(z) : "r"(w4[j][q]), "r"(hsv[q]), "r"(hbv[q]));
                      asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(z));
                      asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(w4[j][q]) : "r"(z), "r"(t));
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr[j]), "r"(w4[j][0]), "r"(w4[j][1]), "r"(w4[j][2]), "r"(w4[j][3]) : "memory");
                  }
                }
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core (async proxy) reads
              asm volatile("bar.sync 3, %0;" ::"n"(TC_XF_THREADS) : "memory");
              if (tt == 0) mbar_arrive(a_ready(as));
              if (tt == 0 && seg == 0 && cb == 0) TRC(5, (sti - (int)blockIdx.x) / (int)gridDim.x);
              if (++as == a_slots) {
                as = 0u;
                aph ^= 1u;
              }
              continue;
            }
            uint32_t cur = 0xffu, nxt = 0xffu;
            float hs[8], hb[8];  // half scale / half shift: silu(z) = hz + hz * tanh(hz), hz = z / 2
            float4 nv[4];        // raw {scale, shift} of the NEXT image of the tile, requested one image ahead: a tile of
                                 // 16x16 / 8x8 images spans 2..5 of them and a load at every switch cost ~4 % of the step
            if (work) {  // coefficients of the tile's first image: in flight while the tile itself is still loading
              cur = 0u;
              const float4* c4 = reinterpret_cast<const float4*>(cf);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 v = __ldg(c4 + q);  // {scale, shift} of two channels
                hs[2 * q] = 0.5f * v.x; hb[2 * q] = 0.5f * v.y; hs[2 * q + 1] = 0.5f * v.z; hb[2 * q + 1] = 0.5f * v.w;
              }
              if (nbase + 1 < a.nimg) {
                nxt = 1u;
                const float4* n4 = reinterpret_cast<const float4*>(cf + (size_t)cC);
#pragma unroll
                for (int q = 0; q < 4; ++q) nv[q] = __ldg(n4 + q);
              }
            }
            // ONE polling thread + a group barrier: one polling lane per warp (and per-warp arrivals below) was measured
            // 15 % slower -- concurrent pollers slow the mbarrier unit the single-thread roles depend on
            if (tt == 0) mbar_wait(a_full(as), aph);
            asm volatile("bar.sync 3, %0;" ::"n"(TC_XF_THREADS) : "memory");
            if (tt == 0 && seg == 0 && cb == 0) TRC(4, (sti - (int)blockIdx.x) / (int)gridDim.x);
            if (work) {
#pragma unroll 1
              for (int r0 = rl; r0 < rows_used; r0 += 2 * rstride) {
                uint32_t w4[2][4], addr[2], nr[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {  // both loads in flight before the math
                  const int r = r0 + j * rstride;
                  if constexpr (DENSE) {
                    nr[j] = r < rows_used ? (uint32_t)s_rown[r >> p.wt_shift] : 0xffu;
                    if (wcheck && (r & (p.Wt - 1)) >= p.W) nr[j] = 0xffu;
                  } else {
                    nr[j] = s_rown[r];
                  }
                  addr[j] = sa + (uint32_t)r * 128u + (uint32_t)((oct ^ (r & 7)) << 4);
                  if (r < rows_used)
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(w4[j][0]), "=r"(w4[j][1]), "=r"(w4[j][2]), "=r"(w4[j][3]) : "r"(addr[j]));
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  if (nr[j] == 0xffu) continue;  // not a pixel: the raw row is zero already (conv padding)
                  if (nr[j] != cur) {
                    cur = nr[j];
                    if (cur != nxt) {  // not the prefetched image (a thread whose rows skip one): load it now
                      const float4* c4 = reinterpret_cast<const float4*>(cf + (size_t)cur * cC);
#pragma unroll
                      for (int q = 0; q < 4; ++q) nv[q] = __ldg(c4 + q);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      hs[2 * q] = 0.5f * nv[q].x; hb[2 * q] = 0.5f * nv[q].y; hs[2 * q + 1] = 0.5f * nv[q].z; hb[2 * q + 1] = 0.5f * nv[q].w;
                    }
                    nxt = 0xffu;
                    if (nbase + (int)cur + 1 < a.nimg) {  // and request the one after it
                      nxt = cur + 1u;
                      const float4* n4 = reinterpret_cast<const float4*>(cf + (size_t)nxt * cC);
#pragma unroll
                      for (int q = 0; q < 4; ++q) nv[q] = __ldg(n4 + q);
                    }
                  }
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[j][q]));
                    const float z0 = fmaf(f.x, hs[2 * q], hb[2 * q]), z1 = fmaf(f.y, hs[2 * q + 1], hb[2 * q + 1]);
                    float t0, t1;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(z0));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(z1));
                    const float v0 = fmaf(z0, t0, z0), v1 = fmaf(z1, t1, z1);
                    const __half2 hv = __floats2half2_rn(v0, v1);
                    const __nv_bfloat162 bv = __floats2bfloat162_rn(v0, v1);
                    w4[j][q] = to_bf16 ? *reinterpret_cast<const uint32_t*>(&bv) : *reinterpret_cast<const uint32_t*>(&hv);
                  }
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr[j]), "r"(w4[j][0]), "r"(w4[j][1]), "r"(w4[j][2]), "r"(w4[j][3]) : "memory");
                }
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core (async proxy) reads
            asm volatile("bar.sync 3, %0;" ::"n"(TC_XF_THREADS) : "memory");
            if (tt == 0) mbar_arrive(a_ready(as));
            if (tt == 0 && seg == 0 && cb == 0) TRC(5, (sti - (int)blockIdx.x) / (int)gridDim.x);
            if (++as == a_slots) {
              as = 0u;
              aph ^= 1u;
            }
          }
        }
      }
    }
  } else {
    // ---------------- epilogue: group eg owns the eg-th unit (UR output rows) of every super-tile ----------------
    const int eg = warp >> 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int et = threadIdx.x - eg * 128;
    const int wq = et >> 5;  // warp inside the group (row quarter for the column sums)
    const int r = q * 32 + lane;  // row inside the unit's M=128 tile
    const int CP = a.CoutP;  // channel stride of the raw output (32, or a multiple of 64; channels >= nmma are written as zeros)
    const uint32_t stg = smem_stage + (uint32_t)eg * TC_STAGE_BYTES;
    const int nstk = STACKED ? p.nstk : 1;
    float* xw = &s_xchg[eg][q][0];
    // rows whose D_j[m + j] sits in the next warp's lane quarter: stored as zeros by the main path, finished by the fix-up
    const bool bnd = !DENSE && nstk > 1 && lane >= 33 - nstk;
    constexpr bool dense = DENSE;
    // row-dense: lane % Wt is the pixel's w (a unit starts on an image-row boundary); the row ends have no neighbour
    const bool ml = (lane & (p.Wt - 1)) != 0, mr = (lane & (p.Wt - 1)) != p.Wt - 1;
    const int g = lane >> 2, t4 = lane & 3;  // mma.sync fragment coordinates of the GroupNorm partials
    int it = 0;
    for (int sti = blockIdx.x; sti < supers; sti += gridDim.x, ++it) {
      const int st = p.descend ? supers - 1 - sti : sti;
      const bool upt1 = p.upt == 1;
      if (upt1 && (it & 1) != eg) continue;  // one-unit tiles: the other group's tile
      const int buf = (!upt1 && p.nbuf == 2) ? (it & 1) : 0;
      const uint32_t use = (upt1 || p.nbuf == 2) ? ((uint32_t)it >> 1) : (uint32_t)it;
      const int tb = upt1 ? eg : buf;  // TMEM barrier pair
      const int unit = upt1 ? st : st * 2 + eg;  // unit index: UR output rows each
      const int m = unit * UR + r;
      int n, h, w;
      // a real pixel of an image < nimg, and one of the UR rows this unit produces (TcArgs::unit_rows)
      bool valid = (dense ? dense_pixel(m, p.H, p.W, p.Wt, p.wt_shift, p.magic_H1, a.nimg, n, h, w)
                          : slot_pixel(m, p.H, p.P, p.magic_P, p.magic_H1, a.nimg, n, h, w)) && r < UR;
      // where this thread's row goes in a stride-2 (subsampled) output: even (h, w) rows map to (h/2, w/2) -- including
      // the pad row h == H and pad column w == W, which are written as zeros -- and the otherwise idle column
      // w == W + 1 writes the second pad column, so the half-resolution raw tensor is zero padded like every other one
      bool wr = false;
      size_t o = 0;
      if (p.sub) {
        const bool inimg = (n < a.nimg) && (h <= p.H) && r < UR;
        valid = valid && !(h & 1) && !(w & 1);
        if (dense) {
          // no pad-column lanes in the row-dense space: the (otherwise idle) lanes w == 1 and w == 3 of an even image row
          // write the two pad columns of the half-resolution tensor; the pad row h == H is part of the row space
          if (inimg && !(h & 1) && w < p.W && (!(w & 1) || w < 4)) {
            wr = true;
            const int w2 = (w & 1) ? p.out_W + (w >> 1) : (w >> 1);
            o = ((size_t)((size_t)n * (p.out_H + 1) + (h >> 1)) * (p.out_W + 2) + w2) * CP;
          }
        } else if (inimg && !(h & 1) && (!(w & 1) || w == p.W + 1)) {
          wr = true;
          const int w2 = (w == p.W + 1) ? p.out_W + 1 : (w >> 1);
          o = ((size_t)((size_t)n * (p.out_H + 1) + (h >> 1)) * (p.out_W + 2) + w2) * CP;
        }
      }
      // ONE thread per group waits for the accumulator (every poll of an mbarrier.try_wait loop is ~8 issued instructions
      // and the pollers were 7 % of all warp-instructions); the group barrier publishes it.  The raw-activation path has that
      // barrier anyway (the staging tile's previous TMA store must have been read out before it is overwritten).
      if constexpr (FINAL) {
        if (et == 0) mbar_wait(tmem_full_bar(tb), use & 1u);
        if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        tc_fence_after();
      }
      if (threadIdx.x == 0) TRC(3, it);
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * 2 + eg) * p.ncols);
      if constexpr (!FINAL) {
        // raw activation, 16-bit, channel stride CP; rows that are not real pixels are stored as exact zeros, so the
        // tensor doubles as a zero-padded conv operand and the per-channel sums below need no masking
        uint16_t* out16 = reinterpret_cast<uint16_t*>(a.out);
        // image boundaries inside this unit (for the GroupNorm partials): rows >= b1 belong to slot 1, >= b2 to slot 2
        const int nf = fast_div(unit * UR, p.magic_R);
        const int b1 = (nf + 1) * p.R - unit * UR;
        const int b2 = b1 + p.R;
        float sat2 = 0.f;  // running max |value| of this row (cvt2_f16_sat)
        float sat_fix = 0.f;  // the same for the boundary rows this thread finishes in the fix-up pass (stacked padded-linear units)
        const uint32_t vm = (valid && !bnd) ? 0xffffffffu : 0u;
        // row-dense statistics: image slot of this row inside the unit, and of the warp's first / last row
        const int my_slot = n - nf;
        const int sl_lo = __shfl_sync(0xffffffffu, my_slot, 0), sl_hi = __shfl_sync(0xffffffffu, my_slot, 31);
        float* s_dpart = &s_part[0][0][0][0][0];  // [group][warp][slot][4 blocks x 8 values]
        [[maybe_unused]] float fq[REGSTATS ? 32 : 1];  // per 16-channel block: sums of its four channel quads, then their sums of squares (this row)
        if constexpr (REGSTATS) {
#pragma unroll
          for (int j = 0; j < 32; ++j) fq[j] = 0.f;
        }
        for (int c0 = 0; c0 < CP; c0 += 64) {
          if (et == 0) {
            if (c0 == 0) mbar_wait(tmem_full_bar(tb), use & 1u);                   // this tile's accumulator is complete
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");         // previous store has left the buffer
          }
          if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
          if (c0 == 0) tc_fence_after();
          // one 16-channel block at a time: accumulator (+ neighbour exchange) -> + bias -> saturating fp16 -> two 16-byte
          // pieces of the 128B-swizzled staging tile [128 rows x 64 channels x 2 B].  The loop stays rolled unless the
          // statistics live in registers (static indexing): the epilogue's code size is felt in the instruction caches.
          auto block = [&](const int blk) {
            const int cc = c0 + blk * 16;
            uint32_t pk[8];
            if (cc < N) {
              float v[16];
              if constexpr (DENSE) acc16_dense(taddr, cc, N, ml, mr, TC_XCHG16, v);
              else acc16(taddr, cc, nstk, N, lane, xw, v);
              [[maybe_unused]] float f[8];  // GroupNorm partials in registers: {sum, sum of squares} of the block's four channel quads
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!p.exp_no_bias) b4 = *reinterpret_cast<const float4*>(&s_bias[cc + 4 * j]);
                const float x0 = v[4 * j] + b4.x, x1 = v[4 * j + 1] + b4.y, x2 = v[4 * j + 2] + b4.z, x3 = v[4 * j + 3] + b4.w;
                pk[2 * j] = cvt2_f16_sat(x0, x1, vm, sat2);
                pk[2 * j + 1] = cvt2_f16_sat(x2, x3, vm, sat2);
                if constexpr (REGSTATS) {
                  f[j] = (x0 + x1) + (x2 + x3);
                  f[4 + j] = fmaf(x3, x3, fmaf(x2, x2, fmaf(x1, x1, x0 * x0)));
                }
              }
              if constexpr (REGSTATS) {  // kept in registers until the accumulator has been released
#pragma unroll
                for (int j = 0; j < 8; ++j) fq[blk * 8 + j] = f[j];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) pk[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint32_t dst = stg + (uint32_t)r * 128u + (uint32_t)(((blk * 2 + j) ^ (r & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * j]), "r"(pk[4 * j + 1]),
                           "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3]) : "memory");
            }
            if (!p.tma_store && wr) {  // stride-2 convs (never stacked): subsampled rows go straight to global memory
#pragma unroll
              for (int j = 0; j < 2; ++j)
                *reinterpret_cast<uint4*>(out16 + o + cc + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            }
          };
          const int nblk = min(4, (CP - c0) >> 4);  // 32-channel outputs: the upper half of the staging tile is never read
          if constexpr (REGSTATS) {
#pragma unroll
            for (int blk = 0; blk < 4; ++blk)
              if (blk < nblk) block(blk);
          } else {
#pragma unroll 1
            for (int blk = 0; blk < nblk; ++blk) block(blk);
          }
          if (c0 + 64 >= N) {  // last chunk that reads TMEM: hand the accumulator back to the MMA warps (one arrival per warp)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(tb));
          }
          if (nstk > 1 && !dense && c0 < N) {
            // fix-up pass: the rows that straddle two warps' lane quarters (3 boundaries x (nstk - 1) rows per unit), from
            // the raw fp32 pieces acc16 left in s_xchg -- 96 threads, one channel pair each
            if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (et < 96) {
              const int b = et >> 5;
              const int half = nstk == 3 ? ((et >> 4) & 1) : 1;  // 0: row 32b + 30 (nstk = 3 only), 1: row 32b + 31
              const int cpair = nstk == 3 ? (et & 15) : (et & 31);
              const int ch = c0 + 2 * cpair;
              if (ch < N) {
                const float* xa = &s_xchg[eg][b][0];      // the warp that owns the row
                const float* xb = &s_xchg[eg][b + 1][0];  // the next warp: D_j of its first lanes
                float2 sum;
                if (nstk == 2) {
                  const float2 d0 = *reinterpret_cast<const float2*>(xa + ch), d1 = *reinterpret_cast<const float2*>(xb + N + ch);
                  sum = make_float2(d0.x + d1.x, d0.y + d1.y);
                } else if (half == 0) {  // D_0[30] + D_1[31] + D_2[next 0]
                  const float2 d0 = *reinterpret_cast<const float2*>(xa + ch), d1 = *reinterpret_cast<const float2*>(xa + 2 * N + ch),
                               d2 = *reinterpret_cast<const float2*>(xb + 4 * N + ch);
                  sum = make_float2((d0.x + d1.x) + d2.x, (d0.y + d1.y) + d2.y);
                } else {                 // D_0[31] + D_1[next 0] + D_2[next 1]
                  const float2 d0 = *reinterpret_cast<const float2*>(xa + N + ch), d1 = *reinterpret_cast<const float2*>(xb + 3 * N + ch),
                               d2 = *reinterpret_cast<const float2*>(xb + 5 * N + ch);
                  sum = make_float2((d0.x + d1.x) + d2.x, (d0.y + d1.y) + d2.y);
                }
                const int rr = 32 * b + 30 + half;
                int n2, h2, w2;
                const bool ok = slot_pixel(unit * UR + rr, p.H, p.P, p.magic_P, p.magic_H1, a.nimg, n2, h2, w2);
                float satb = 0.f;
                const uint32_t pv = cvt2_f16_sat(sum.x + s_bias[ch], sum.y + s_bias[ch + 1], ok ? 0xffffffffu : 0u, satb);
                if (ok) sat_fix = fmaxf(sat_fix, satb);
                const uint32_t dst = stg + (uint32_t)rr * 128u + (uint32_t)((((ch - c0) >> 3) ^ (rr & 7)) << 4) + (uint32_t)(((ch - c0) & 7) << 1);
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(pv) : "memory");
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
          if (p.tma_store && et == 0) {
            if (dense) {
              // box [64 channels][Wt pixels][128 / Wt image rows]; when the tile is as wide as the image the two pad columns
              // of the padded-linear tensor lie outside it and are written as zeros by a second box (w = W, W + 1)
              tma_store_3d(&tmO, stg, c0, 0, unit * (128 >> p.wt_shift));
              if (p.Wt == p.W) tma_store_3d(&tmZ, smem_u32(&s_xchg[0][0][0]), c0, p.W, unit * (128 >> p.wt_shift));
            } else {
              tma_store_2d(&tmO, stg, c0, unit * UR);  // the box has UR rows
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (threadIdx.x == 0 && c0 == 0) TRC(6, it);
          if constexpr (REGSTATS) {
          if (p.stats != nullptr && c0 < N) {
            // combine the four warps' quad sums; every channel of a quad is credited a quarter of the quad's sums, so the
            // consumers' per-channel gather (common.h stats_channel_sums) adds up to the group sums it needs unchanged
            // (GroupNorm groups are whole quads: 4, 8, 12, 16, 24, 32 channels)
            // Statistics from the registers (no ldmatrix, no legacy HMMA queued behind the UMMA stream), after the
            // accumulator went back to the MMA warps: quad sums of the stored rows, summed over the warp's 32 rows by
            // one 31-shuffle transpose-reduce.  A warp's rows lie in at most two images (slots sl_lo / sl_hi of the unit);
            // the second pass runs only for a warp that straddles an image boundary.
            if (N <= 32) {  // two 16-channel blocks: 16 values
              const float ta = transpose_reduce16_masked(fq, valid && my_slot == sl_lo, lane);
              float tb = 0.f;
              if (sl_hi != sl_lo) tb = transpose_reduce16_masked(fq, valid && my_slot == sl_hi, lane);
              if (!(lane & 1)) {
                float* sp = s_dpart + ((eg * 4 + wq) * 3) * 32 + (lane >> 1);
                for (int sl = 0; sl < p.nslots; ++sl) sp[sl * 32] = sl == sl_lo ? ta : (sl == sl_hi ? tb : 0.f);
              }
            } else {
              const float ta = transpose_reduce32_masked(fq, valid && my_slot == sl_lo, lane);
              float tb = 0.f;
              if (sl_hi != sl_lo) tb = transpose_reduce32_masked(fq, valid && my_slot == sl_hi, lane);
              float* sp = s_dpart + ((eg * 4 + wq) * 3) * 32 + lane;
              for (int sl = 0; sl < p.nslots; ++sl) sp[sl * 32] = sl == sl_lo ? ta : (sl == sl_hi ? tb : 0.f);
            }
            if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (et < p.nslots * 32) {
              const int slot = et >> 5, j = et & 31;
              const float* sp = s_dpart + (eg * 4 * 3 + slot) * 32 + j;
              const float tot = 0.25f * (((sp[0] + sp[3 * 32]) + sp[2 * 3 * 32]) + sp[3 * 3 * 32]);
              const int ch = c0 + (j >> 3) * 16 + (j & 3) * 4;
              if (ch < N) {
                float* dst = p.stats + ((size_t)(unit * p.nslots + slot) * CP + ch) * 2 + ((j >> 2) & 1);
                dst[0] = tot; dst[2] = tot; dst[4] = tot; dst[6] = tot;
              }
            }
          }
          } else if (p.stats != nullptr && c0 < N) {
            // GroupNorm partials of the ROUNDED values on the legacy tensor path, channels on M and image slots on N:
            //   S[ch][slot] = sum_row X[row][ch] * Ind[row][slot]   (Ind = 0/1: row belongs to image slot of the unit)
            // per warp = its 32 staged rows: 4 channel blocks x 2 row blocks of mma.sync m16n8k16, A = X^T by
            // ldmatrix.trans from the swizzled staging tile, B = the indicator built in registers.  The squares are
            // fp16 too ((x / 16)^2, sq2_f16_scaled).  16 + 16 HMMAs per warp and 64-channel chunk (a [slot][row] x
            // [row][ch] formulation needs 4x as many, and HMMA is what bounds the epilogue).
            float sm[4][4], sq[4][4];
#pragma unroll
            for (int cb = 0; cb < 4; ++cb)
#pragma unroll
              for (int e = 0; e < 4; ++e) sm[cb][e] = sq[cb][e] = 0.f;
            const int mat = lane >> 3, rrow = lane & 7;
            const bool one_img = b1 >= UR;  // the whole unit lies inside one image (9 of 10 units at 32x32): slot 0 everywhere
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              const int r0 = wq * 32 + kb * 16;
              // B fragment: b0 = Ind[k0, k0+1][g], b1 = Ind[k0+8, k0+9][g], k0 = 2 * t4 (column g = image slot g), fp16 1.0 / 0.0
              uint32_t b0h = g == 0 ? 0x3C003C00u : 0u, b1h = b0h;
              if (!one_img) {
                uint32_t ind[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int row = r0 + (t4 << 1) + (e & 1) + ((e >> 1) << 3);
                  const int sl = (row >= b1) + (row >= b2);
                  ind[e] = (sl == g) ? 1u : 0u;
                }
                b0h = ind[0] * 0x3C00u + ind[1] * 0x3C000000u;
                b1h = ind[2] * 0x3C00u + ind[3] * 0x3C000000u;
              }
              // ldmatrix.x4.trans: matrices (rows r0.., chunk 2cb), (rows r0.., chunk 2cb+1), (rows r0+8.., 2cb), (rows r0+8.., 2cb+1)
              // = A fragments a0 (ch 0-7, k 0-7), a1 (ch 8-15, k 0-7), a2 (ch 0-7, k 8-15), a3 (ch 8-15, k 8-15)
              const int lrow = r0 + ((mat >> 1) << 3) + rrow;
#pragma unroll
              for (int cb = 0; cb < 4; ++cb) {
                if (c0 + cb * 16 >= N) break;  // channel blocks past the real output channels hold zeros
                const uint32_t chunk16 = (uint32_t)(2 * cb + (mat & 1));
                uint32_t af[4];
                ldsm_x4_trans(stg + (uint32_t)lrow * 128u + ((chunk16 ^ (uint32_t)(lrow & 7)) << 4), af[0], af[1], af[2], af[3]);
                mma_f16_16816(sm[cb], af, b0h, b1h);
                uint32_t a2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) a2[e] = sq2_f16_scaled(af[e]);
                mma_f16_16816(sq[cb], a2, b0h, b1h);
              }
            }
            // D fragment: c0 = (ch g, slot 2*t4), c1 = (ch g, slot 2*t4+1), c2 / c3 = the same for ch g + 8
            if (t4 < 2) {
#pragma unroll
              for (int cb = 0; cb < 4; ++cb) {
                if (c0 + cb * 16 >= N) break;
                float* d0 = &s_part[eg][wq][2 * t4][cb * 16 + g][0];
                d0[0] = sm[cb][0]; d0[1] = 256.f * sq[cb][0];       // slot 2*t4, channel g
                d0[16] = sm[cb][2]; d0[17] = 256.f * sq[cb][2];     // slot 2*t4, channel g + 8
                if (t4 == 0) {
                  float* d1 = &s_part[eg][wq][1][cb * 16 + g][0];
                  d1[0] = sm[cb][1]; d1[1] = 256.f * sq[cb][1];     // slot 1
                  d1[16] = sm[cb][3]; d1[17] = 256.f * sq[cb][3];
                }
              }
            }
            if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            for (int i = et; i < p.nslots * 128; i += 128) {
              const int slot = i >> 7, rem = i & 127;
              const float* pp = &s_part[eg][0][slot][0][0] + rem;
              const float tot = ((pp[0] + pp[3 * 128]) + pp[2 * 3 * 128]) + pp[3 * 3 * 128];
              if (c0 + (rem >> 1) < N) p.stats[((size_t)(unit * p.nslots + slot) * CP + c0) * 2 + rem] = tot;
            }
          }
        }
        if (threadIdx.x == 0) TRC(7, it);
        if (p.sat != nullptr) {  // rows of this tile that hit the fp16 bound (reported in diag[12]; must stay 0)
          const unsigned hit = __ballot_sync(0xffffffffu, (vm != 0u && sat2 >= 65504.f) || sat_fix >= 65504.f);
          if (hit != 0u && lane == 0) atomicAdd(p.sat, (unsigned int)__popc(hit));
        }
      } else {
        // output conv (never stacked): Cout <= 16 real channels -> NCHW velocity, or fused Euler update x += dt * v
        float v[16];
        if (dense) {
          acc16_dense(taddr, 0, N, ml, mr, false, v);
        } else {
          uint32_t u[16];
          tmem_ld16(taddr, u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(u[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty_bar(tb));
        if (valid) {
          for (int c = 0; c < a.Cout; ++c) {
            const float val = v[c] + s_bias[c];
            const size_t oc = (((size_t)n * a.Cout + c) * p.out_H + h) * p.out_W + w;
            if (a.out_mode == OUT_NCHW)
              a.out[oc] = val;
            else if (a.out_mode == OUT_NCHW_SCALED)
              a.out[oc] = val * a.dt;
            else
              a.out[oc] = __fadd_rn(a.out[oc], __fmul_rn(val, a.dt));
          }
        }
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // smem must outlive the stores
  }
#ifdef RGFM_TC_TRACE
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    printf("TRACE nmma=%d nstk=%d nseg=%d cblocks=%d taps=%d resident=%d a_slots=%d supers=%d xform=%d end=%lld\n",
           p.nmma, p.nstk, p.nseg, p.cblocks[0], p.ntaps[0], p.resident, p.a_slots, supers, p.any_xform, clock64() - t_start);
    printf("  prologue end %lld, resident weights landed %lld, dense=%d\n", s_tp[0], p.resident ? s_tp[1] : 0ll, (int)DENSE);
    for (int i = 0; i < 12 && blockIdx.x + i * gridDim.x < supers; ++i)
      printf("  st%2d prodA %7lld | xform wake %7lld done %7lld | mma a_ready %7lld committed %7lld | epi tmem_full %7lld stored %7lld done %7lld\n", i,
             s_trace[0][i], s_trace[4][i], s_trace[5][i], s_trace[1][i], s_trace[2][i], s_trace[3][i], s_trace[6][i], s_trace[7][i]);
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
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
  CUtensorMap tmA[3], tmB[3], tmO, tmZ;
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

static void encode_3d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t mid, uint64_t outer, uint32_t box_inner,
                      uint32_t box_mid, uint32_t box_outer, CUtensorMapSwizzle swz, const char* what) {
  cuuint64_t dims[3] = {inner, mid, outer};
  cuuint64_t strides[2] = {inner * 2, inner * mid * 2};
  cuuint32_t box[3] = {box_inner, box_mid, box_outer};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RGFM_CHECK(r == CUDA_SUCCESS, std::string("cuTensorMapEncodeTiled(") + what + ") failed with code " + std::to_string((int)r));
}

// Row-dense mode (TcArgs::dense): convs with at most 64 output channels (three kx column groups of an accumulator are
// 3 * nmma <= 192 TMEM columns) whose image is at most 32 pixels wide.  It trades shared-memory operand traffic and MMA
// instructions (a third of both) for epilogue instructions (three column groups to read, the neighbour exchange), and
// the kernel is short of BOTH the shared-memory port and issue slots (ncu: 55..58 % of the issue slots busy with 19
// warps, profiles/notes_r02.md): measured per layer it wins 10..20 % where a super-tile walks at least two operand
// stages (decoder conv1s, every conv2 with its skip / identity segment) and on the output convs (N = 16: 12 instead of
// 36 MMAs), and loses 8..25 % on single-stage convs, whose epilogue is the critical role.  Hence the rule below.
// RGFM_TC_DENSE=0: padded-linear everywhere, 2: row-dense wherever it is legal (A/B timing).
// stages: 64-channel operand blocks of all K segments.
bool tc_dense(int cout, int W, bool final, int stages) {
  const int mode = getenv("RGFM_TC_DENSE") ? atoi(getenv("RGFM_TC_DENSE")) : 1;  // read at plan build time (tests switch it)
  const int nmma = (cout + 15) / 16 * 16;
  if (mode == 0 || nmma > 64 || W > 32 || W < 4) return false;
  // Cout <= 32 (third session, after the instruction diet: the lean per-image transform loop exists for row-dense tiles
  // only): single-stage 32->32 convs 88.6 -> 78.8 us, the 32->32 stride-2 conv 80.1 -> 68.1 us, the 1->32 input conv
  // 66.2 -> 68.5 us (profiles/policy_matrix_r02.txt)
  return mode == 2 || final || stages >= 2 || nmma <= 32;
}
static int tc_tile_width(int W) { return W <= 8 ? 8 : (W <= 16 ? 16 : 32); }
// GEMM rows per image of the row space a conv's GroupNorm partials are indexed by
int tc_rows_per_image(int cout, int H, int W, bool final, int stages) {
  return tc_dense(cout, W, final, stages) ? (H + 1) * tc_tile_width(W) : (H + 1) * (W + 2);
}

// Column groups per accumulator (TcArgs::nstk).  Stacking shortens the MMA phase of a tile (N = 96: 57 cycles, N = 192: 97,
// N = 128 + 64: 115 against 3 x 49.7) but lengthens its epilogue (three / two accumulator column groups to read, the
// lane-shifted sum, the boundary fix-up), and the epilogue is the critical role of a conv whose tile holds ONE operand
// stage (measured, profiles/notes_r02.md: 64->64 @32x32 136 -> 191 us stacked, 192->64 339 -> 313 us).  So taps are stacked
// only where the MMA phase dominates: stride-1 convs with a raw 16-bit output (their epilogue has the staging tile the
// fix-up writes into) whose super-tile walks at least two operand stages -- for Cout <= 32 at least two 3x3 stages.
// stages3 / stages: 64-channel operand blocks of the 3x3 segments / of all segments.
// RGFM_TC_STACK=0: the classic nine-tap issue everywhere; 2: stack wherever it is legal (A/B timing; same results up to
// fp32 summation order).
int tc_stack_groups(int cout, int stride, bool final, int stages3, int stages) {
  static const int mode = getenv("RGFM_TC_STACK") ? atoi(getenv("RGFM_TC_STACK")) : 1;
  const int nmma = (cout + 15) / 16 * 16;
  if (mode == 0 || stride != 1 || final || nmma > 64) return 1;
  const int nstk = nmma <= 32 ? 3 : 2;
  if (mode == 2) return nstk;
  return (nstk == 3 ? stages3 >= 2 : stages >= 2) ? nstk : 1;
}
int tc_unit_rows(int cout, int W, int stride, bool final, int stages3, int stages) {
  if (tc_dense(cout, W, final, stages)) return 128;
  return 128 - (tc_stack_groups(cout, stride, final, stages3, stages) - 1);
}

size_t tc_stats_floats(int R, int chunk, int unit_rows, int CoutP) {
  const int SR = 2 * unit_rows;
  const size_t supers = ((size_t)chunk * R + SR - 1) / SR;
  return supers * 2 * (size_t)stats_slots_for(R) * CoutP * 2;
}

TcPtr make_tc_conv(const TcConvDesc& d, DevPool& pool) {
  const int nmma = (d.Cout + 15) / 16 * 16;
  RGFM_CHECK(nmma >= 16 && nmma <= 256, "tc conv: Cout must be 1..256");
  RGFM_CHECK(d.out == nullptr || ((d.CoutP % 64 == 0 || d.CoutP == 32) && d.CoutP >= nmma && nmma % 32 == 0),
             "tc conv: raw outputs need Cout % 32 == 0 and a channel stride of 32 or a multiple of 64");
  RGFM_CHECK(d.Win + 2 <= 39, "tc conv: image width must be <= 37");
  RGFM_CHECK(d.nseg >= 1 && d.nseg <= 3, "tc conv: one to three K segments");
  TcPtr t(new TcConv());
  TcArgs& p = t->p;
  std::memset(&p, 0, sizeof(p));
  const int P = d.Win + 2;
  p.H = d.Hin; p.W = d.Win; p.P = P;
  {
    int stall = 0;
    for (int sg = 0; sg < d.nseg; ++sg) stall += (d.seg[sg].Cin_real + 63) / 64;
    p.dense = tc_dense(d.Cout, d.Win, d.out == nullptr && d.stride == 1, stall) ? 1 : 0;
  }
  {
    // register statistics: convs whose super-tile walks ONE operand stage (see REGSTATS above), with at least 64 output
    // channels, and no transform warps competing for issue slots unless the conv is wide (measured per layer, same box:
    // upsamplers.0 128 -> 121 us, downsamplers.0 64->64 129 -> 117, enc2.conv1 64->128 79 -> 67; 32-channel convs and the
    // fused 64->64 conv1s lose 3..8 %).  Never with taps stacked in the padded-linear row space (the boundary rows of a
    // stacked unit are finished after the main pass).  RGFM_TC_REG_STATS=0: never, 2: wherever legal (A/B timing).
    const int mode = getenv("RGFM_TC_REG_STATS") ? atoi(getenv("RGFM_TC_REG_STATS")) : 1;
    int stall = 0;
    bool xf = false;
    for (int sg = 0; sg < d.nseg; ++sg) {
      stall += (d.seg[sg].Cin_real + 63) / 64;
      xf = xf || d.seg[sg].coef != nullptr;
    }
    p.reg_stats = (mode == 2 || (mode == 1 && stall == 1 && nmma >= 64 && (!xf || nmma >= 128))) ? 1 : 0;
  }
  p.upt = 2;
  for (int sg = 0; sg < d.nseg; ++sg) p.stages_total += (d.seg[sg].Cin_real + 63) / 64;
  if (p.dense) {
    RGFM_CHECK(d.seg[0].ks == 3, "tc conv: segment 0 of a row-dense conv is 3x3");
    p.Wt = tc_tile_width(d.Win);
    p.wt_shift = p.Wt == 8 ? 3 : (p.Wt == 16 ? 4 : 5);
    // one-unit tiles (TcArgs::upt): N = 64 and every weight tile resident next to two 128-row A slots and the staging tiles
    uint32_t wbytes = 0;
    for (int sg = 0; sg < d.nseg; ++sg) wbytes += (uint32_t)((d.seg[sg].Cin_real + 63) / 64) * (uint32_t)(d.seg[sg].ks * d.seg[sg].ks) * (uint32_t)nmma * 128u;
    const uint32_t slot1 = (uint32_t)(128 / p.Wt + 2) * p.Wt * 128u;
    const int upt_mode = getenv("RGFM_TC_UPT") ? atoi(getenv("RGFM_TC_UPT")) : 1;  // 0: never (A/B timing)
    // measured per layer (same box): -4..-8 % on the two-stage convs (conv2 + identity residual), +19..24 % on the convs
    // with three or four stages per tile (conv2 + 1x1 skip over a concatenation: more stages than A slots) -> two stages only
    if (upt_mode != 0 && nmma == 64 && p.stages_total == 2 && 2u * slot1 + 2u * TC_STAGE_BYTES + wbytes <= TC_SMEM_BUDGET) p.upt = 1;
    p.irs = (p.upt == 1 ? 128 : 256) / p.Wt;
  }
  p.R = p.dense ? (d.Hin + 1) * p.Wt : (d.Hin + 1) * P;
  {
    auto magic = [&](int dv) {
      const uint64_t nmax = (uint64_t)d.chunk * p.R + 2 * 256;
      RGFM_CHECK(nmax * (uint64_t)(dv > 1 ? dv - 1 : 1) < (1ull << 32), "tc conv: row space too large for the fast division");
      return (uint32_t)(((1ull << 32) + dv - 1) / dv);
    };
    p.magic_P = magic(P);
    p.magic_H1 = magic(d.Hin + 1);
    p.magic_R = magic(p.R);
  }
  p.nseg = d.nseg;
  p.nmma = nmma;
  {
    int st3 = 0, stall = 0;
    for (int sg = 0; sg < d.nseg; ++sg) {
      const int cb = (d.seg[sg].Cin_real + 63) / 64;
      stall += cb;
      if (d.seg[sg].ks == 3) st3 += cb;
    }
    p.nstk = p.dense ? 3 : tc_stack_groups(d.Cout, d.stride, d.out == nullptr && d.stride == 1, st3, stall);
    if (!p.dense && p.nstk > 1) p.reg_stats = 0;
  }
  p.unit_rows = p.dense ? 128 : 128 - (p.nstk - 1);
  p.super_rows = p.upt * p.unit_rows;
  p.ncols = (p.nstk * nmma + 31) / 32 * 32;
  p.sat = d.sat;
  const int rows_needed = p.super_rows + 2 * P + 2;
  p.a_box_rows = ((rows_needed + 1) / 2 + 7) / 8 * 8;
  RGFM_CHECK(p.a_box_rows <= 256, "tc conv: A box too tall");
  p.a_slot = p.dense ? (uint32_t)(p.irs + 2) * p.Wt * 128u : 2u * p.a_box_rows * 128u;
  p.b_bytes = (uint32_t)nmma * 128u;
  p.nbuf = (4 * p.ncols <= 512) ? 2 : 1;  // (one-unit tiles: one accumulator per MMA warp, 2 * ncols columns, see TcArgs::upt)
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * p.nbuf * p.ncols)) cols <<= 1;
  p.tmem_cols = cols;
  p.tma_store = (d.out != nullptr) ? 1 : 0;  // 32-channel raws: the same 64-wide store box, its upper half clipped by TMA
  p.out_fmt = 2;  // raw activations are stored in fp16 in both 16-bit modes (11-bit mantissa; the epilogue saturates)
  p.stats = d.stats;
  p.rows_alloc = d.chunk * p.R;
  p.nslots = stats_slots_for(p.R);
  RGFM_CHECK(d.stats == nullptr || p.nslots <= 3, "tc conv: GroupNorm partials need images of at least 64 GEMM rows");
  // shared-memory plan: [A slots][2 staging tiles][weights].  If every weight tile of the conv fits next to two A slots
  // the weights stay RESIDENT (loaded once per CTA) and whatever is left deepens the A pipeline; otherwise the weights
  // stream through a ring of stages of tps taps (tps in {9, 3, 1}: as many as keep >= 3 stages in the budget).
  const uint32_t budget = TC_SMEM_BUDGET;
  const uint32_t stage = 2u * TC_STAGE_BYTES;  // one epilogue staging tile per accumulator half
  uint32_t b_total = 0;
  for (int sg = 0; sg < d.nseg; ++sg) b_total += (uint32_t)((d.seg[sg].Cin_real + 63) / 64) * (uint32_t)(d.seg[sg].ks * d.seg[sg].ks) * p.b_bytes;
  p.b_total = b_total;
  // Measured (profiles/notes_r01.md): keeping the weights resident is worth far more than a deeper A pipeline -- streaming
  // the weight tiles from L2 for every super-tile is what bounds the ring mode (L2 -> SM bandwidth), not load latency.
  static const int force_slots = getenv("RGFM_TC_A_SLOTS") ? atoi(getenv("RGFM_TC_A_SLOTS")) : 0;
  const uint32_t min_slots = force_slots ? (uint32_t)force_slots : 2u;
  p.resident = (min_slots * p.a_slot + stage + b_total <= budget) ? 1 : 0;
  RGFM_CHECK(p.upt == 2 || p.resident, "tc conv: one-unit tiles need resident weights");
  p.tps[0] = p.tps[1] = p.tps[2] = 1;
  if (p.resident) {
    int slots = (int)((budget - stage - b_total) / p.a_slot);
    p.a_slots = slots > 4 ? 4 : slots;
    p.b_stages = 1;
    p.b_stage = p.b_bytes;
    t->smem = (size_t)p.a_slots * p.a_slot + stage + b_total + 1024;
  } else {
    // weight ring: at least two stages of >= 1 tap must fit next to the A slots, else fall back to two slots
    uint32_t slots = min_slots;
    while (slots > 2u && slots * p.a_slot + stage + 2u * p.b_bytes > budget) --slots;
    p.a_slots = (int)slots;
    const uint32_t fixed = slots * p.a_slot + stage;
    const uint32_t ring_budget = budget - fixed;
    // Every ring stage costs the single producer / MMA threads ~500 cycles of hand-off (wait, fence, commit at one
    // dependent instruction per ~5 cycles), so stages are made as large as two of them fit: measured 6.37 -> 5.99 ms of
    // conv time per Euler step against three-plus smaller stages (profiles/notes_r01.md).
    int tps = 1;
    static const int min_stages = getenv("RGFM_TC_MIN_STAGES") ? atoi(getenv("RGFM_TC_MIN_STAGES")) : 2;
    if (d.seg[0].ks == 3) {
      if ((uint32_t)min_stages * 9u * p.b_bytes <= ring_budget) tps = 9;
      else if ((uint32_t)min_stages * 3u * p.b_bytes <= ring_budget) tps = 3;
    }
    // a stacked MMA group reads the tap tiles of one filter row as ONE B operand: they must sit in the same ring stage
    RGFM_CHECK(p.nstk == 1 || tps % 3 == 0, "tc conv: stacked taps need whole filter rows per weight-ring stage");
    for (int sg = 0; sg < d.nseg; ++sg)
      if (d.seg[sg].ks == 3) p.tps[sg] = tps;
    p.b_stage = (uint32_t)tps * p.b_bytes;
    int stages = (int)(ring_budget / p.b_stage);
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    RGFM_CHECK(stages >= 2, "tc conv: not enough shared memory for the weight ring");
    p.b_stages = stages;
    t->smem = (size_t)fixed + (size_t)stages * p.b_stage + 1024;
  }

  for (int sg = 0; sg < d.nseg; ++sg) {
    const TcSegment& S = d.seg[sg];
    const bool xf = S.coef != nullptr;
    RGFM_CHECK(!xf || (S.raw && S.Cin_real % 16 == 0), "tc conv: a fused GroupNorm + SiLU segment must be a raw tensor with channels % 16 == 0");
    // raw tensors are fp16 in memory; a transformed tile is rewritten in the net's operand format before the MMA reads it
    const bool seg_fp16 = d.fp16 || (S.raw && !xf);
    const CUtensorMapDataType dt16 = seg_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    p.xform[sg] = xf ? 1 : 0;
    p.coef[sg] = reinterpret_cast<const float2*>(S.coef);
    p.coef_h[sg] = xf ? reinterpret_cast<const uint4*>(S.coef + (size_t)d.chunk * S.coef_C * 2) : nullptr;
    RGFM_CHECK(!xf || (S.coef_c0 % 8 == 0 && S.coef_C % 8 == 0), "tc conv: GroupNorm coefficient offsets must be multiples of 8");
    p.coef_C[sg] = S.coef_C;
    p.coef_c0[sg] = S.coef_c0;
    p.cin_real[sg] = S.Cin_real;
    if (xf) {
      p.any_xform = 1;
      p.xf_bf16 = d.fp16 ? 0 : 1;
      RGFM_CHECK(S.coef_c0 % 2 == 0 && S.coef_C % 2 == 0, "tc conv: GroupNorm coefficient offsets must be even");
    }
    const uint32_t fmt = seg_fp16 ? 0u : 1u;  // UMMA F16F32Format: 0 = F16, 1 = BF16
    auto idesc_n = [&](int n) { return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); };
    p.idesc[sg] = idesc_n(nmma);
    // MMA groups of a filter row (TcArgs::nstk): which tap starts a group and how many taps it stacks on N
    const int grp[3][3] = {{1, 1, 1}, {2, 0, 1}, {3, 0, 0}};
    for (int kx = 0; kx < 3; ++kx) p.idesc3[sg][kx] = grp[p.nstk - 1][kx] ? idesc_n(grp[p.nstk - 1][kx] * nmma) : 0u;
    // a 32-channel tensor (64-byte rows) is read through the same 64-channel TMA box: the upper half is out of bounds and
    // zero-filled, so the shared-memory tile and the MMA loop are unchanged while the global traffic halves
    RGFM_CHECK((S.Cin % 64 == 0 || S.Cin == 32 || S.Cin == 16) && S.Cin <= 256, "tc conv: operand channels must be 16, 32 or a multiple of 64, at most 256");
    RGFM_CHECK(S.ks == 1 || S.ks == 3, "tc conv: 1x1 or 3x3");
    RGFM_CHECK(S.Cin_real >= 1 && S.Cin_real <= S.Cin, "tc conv: real channel count");
    p.cblocks[sg] = (S.Cin_real + 63) / 64;  // blocks that hold no real channel are skipped altogether
    for (int cb = 0; cb < p.cblocks[sg]; ++cb) {
      const int left = S.Cin_real - cb * 64;
      p.ksteps[sg][cb] = left >= 64 ? 4 : (left + 15) / 16;
    }
    p.ntaps[sg] = S.ks * S.ks;
    // a raw activation tensor is indexed by the GEMM row itself: operand slot = raw row + P + 1
    // 3x3: the tile starts P + 1 rows before the super-tile's first pixel (halo); 1x1: exactly at it -- 256 rows instead
    // of 256 + 2P + 2, and MMA operand rows aligned to the 8-row swizzle atom
    p.row_off[sg] = (S.raw ? -(P + 1) : 0) + (S.ks == 1 ? P + 1 : 0);
    p.box_rows[sg] = S.ks == 1 ? 128 : p.a_box_rows;
    RGFM_CHECK(!xf || S.ks == 3, "tc conv: fused GroupNorm + SiLU segments are 3x3");
    // weights: torch [Cout][Cin_real][ks][ks] fp32 (or identity when w_oihw == nullptr) -> [tap][nmma][Cin] 16-bit
    const int taps = p.ntaps[sg];
    std::vector<uint16_t> hw((size_t)taps * nmma * S.Cin, 0);
    for (int o = 0; o < d.Cout; ++o)
      for (int c = 0; c < S.Cin_real; ++c)
        for (int tp = 0; tp < taps; ++tp) {
          const int wct = S.w_cin_total > 0 ? S.w_cin_total : S.Cin_real;
          const float wv = S.w_oihw ? S.w_oihw[((size_t)o * wct + S.w_c0 + c) * taps + tp] : (o == c ? 1.0f : 0.0f);
          // an operand outside the 16-bit format's range would become inf and poison every output it touches
          RGFM_CHECK(std::isfinite(wv) && (!seg_fp16 || std::fabs(wv) <= 65504.f), "tc conv: a weight is not finite or outside fp16's range");
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
    if (p.dense) {
      // 3-D view [Cin][P][image rows] of the same padded-linear tensor; the box takes Wt pixels of each image row (the
      // pad columns stay behind; for W < Wt the columns past the row are pads / out of bounds = zeros).  A packed operand
      // holds pixel (image row, w) at (image row + 1, w + 1).  3x3 segments start one image row above the super-tile.
      const int irows = S.ks == 3 ? p.irs + 2 : p.irs;
      p.w0[sg] = S.raw ? 0 : 1;
      p.ir_off[sg] = (S.raw ? 0 : 1) + (S.ks == 3 ? -1 : 0);
      p.a_bytes[sg] = (uint32_t)irows * p.Wt * 128u;
      const uint64_t image_rows = (uint64_t)d.chunk * (d.Hin + 1) + (S.raw ? 0 : 1);
      cuuint64_t dims[3] = {(cuuint64_t)S.Cin, (cuuint64_t)P, image_rows};
      cuuint64_t strides[2] = {(cuuint64_t)S.Cin * 2, (cuuint64_t)S.Cin * P * 2};
      cuuint32_t box[3] = {64, (cuuint32_t)p.Wt, (cuuint32_t)irows};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = encode_fn()(&t->tmA[sg], S.raw ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : dt16, 3, const_cast<void*>(S.in), dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      RGFM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A, row-dense) failed with code " + std::to_string((int)r));
    } else {
    // operand: 2-D [rows][Cin]; rows before the first / after the last are zero-filled by TMA
    encode_2d(&t->tmA[sg], S.raw ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : dt16, 2, S.in, (uint64_t)S.Cin,
              S.raw ? tc_raw_rows(d.Hin, d.Win, d.chunk) : tc_operand_slots(d.Hin, d.Win, d.chunk), 64, (uint32_t)p.box_rows[sg], "A");
    }
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
  t->tmZ = t->tmA[0];
  if (p.tma_store && p.dense) {
    const uint64_t image_rows = (uint64_t)d.chunk * (d.Hin + 1);
    encode_3d(&t->tmO, d.out, (uint64_t)d.CoutP, (uint64_t)P, image_rows, 64, (uint32_t)p.Wt, (uint32_t)(128 / p.Wt), CU_TENSOR_MAP_SWIZZLE_128B, "out, row-dense");
    encode_3d(&t->tmZ, d.out, (uint64_t)d.CoutP, (uint64_t)P, image_rows, 64, 2, (uint32_t)(128 / p.Wt), CU_TENSOR_MAP_SWIZZLE_NONE, "pad columns");
  } else if (p.tma_store)
    encode_2d(&t->tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d.out, (uint64_t)d.CoutP, tc_raw_rows(d.Hin, d.Win, d.chunk), 64,
              (uint32_t)p.unit_rows, "out");
  else
    t->tmO = t->tmA[0];
  // per device (function attributes belong to the current context): cheap enough to set for every plan
#define RGFM_TC_ATTR(D, R, X, F, S) RGFM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<D, R, X, F, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BUDGET + 1024))
#define RGFM_TC_ATTR4(X, F) RGFM_TC_ATTR(false, false, X, F, false); RGFM_TC_ATTR(true, false, X, F, false)
  RGFM_TC_ATTR4(false, false); RGFM_TC_ATTR4(true, false); RGFM_TC_ATTR4(false, true); RGFM_TC_ATTR4(true, true);
  RGFM_TC_ATTR(false, true, false, false, false); RGFM_TC_ATTR(true, true, false, false, false); RGFM_TC_ATTR(false, true, true, false, false); RGFM_TC_ATTR(true, true, true, false, false);
  RGFM_TC_ATTR(false, false, false, false, true); RGFM_TC_ATTR(false, false, true, false, true);  // stacked padded-linear units (RGFM_TC_DENSE=0)
#undef RGFM_TC_ATTR4
#undef RGFM_TC_ATTR
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
    RGFM_CHECK((a.CoutP % 64 == 0 || a.CoutP == 32) && a.CoutP >= p.nmma && a.out_padded && a.out_fmt == p.out_fmt && a.out != nullptr,
               "tc conv: raw outputs are 16-bit, padded-linear, channel stride 32 or a multiple of 64");
  }
  static const int descend = getenv("RGFM_TC_DESCEND") ? atoi(getenv("RGFM_TC_DESCEND")) : 1;
  p.descend = descend;
  static const int l2pf = getenv("RGFM_TC_L2_PREFETCH") ? atoi(getenv("RGFM_TC_L2_PREFETCH")) : 1;
  p.l2_prefetch = l2pf;
#ifdef RGFM_EXPERIMENTS  // timing-only switches that knowingly corrupt results: compiled out of the shipped library
  static const int noxf = getenv("RGFM_EXPERIMENT_NO_XFORM") ? atoi(getenv("RGFM_EXPERIMENT_NO_XFORM")) : 0;  // 1: no work, 2: no coefficient switches
  p.exp_no_xform = noxf;
  static const int nobias = getenv("RGFM_EXPERIMENT_NO_BIAS") ? atoi(getenv("RGFM_EXPERIMENT_NO_BIAS")) : 0;
  p.exp_no_bias = nobias;
  static const bool no_stats = getenv("RGFM_TC_NO_STATS") != nullptr;
  if (no_stats) p.stats = nullptr;
#endif
  const int total_rows = a.nimg * (p.H + 1) * (p.dense ? p.Wt : p.P);
  const int supers = (total_rows + p.super_rows - 1) / p.super_rows;
  static int num_sms = 0;  // B200 only: every device of the box has the same count
  if (!num_sms) {
    int dev = 0;
    RGFM_CUDA(cudaGetDevice(&dev));
    RGFM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int cap = (g_tc_cta_cap > 0 && g_tc_cta_cap < num_sms) ? g_tc_cta_cap : num_sms;  // see common.h
  const int grid = supers < cap ? supers : cap;  // persistent: one CTA per SM, static round-robin
#define RGFM_TC_LAUNCH(D, R, X, F, S) launch_pdl(conv_tc_kernel<D, R, X, F, S>, dim3(grid), dim3(TC_THREADS), tc.smem, s, tc.tmA[0], tc.tmB[0], tc.tmA[1], tc.tmB[1], tc.tmA[2], tc.tmB[2], tc.tmO, tc.tmZ, p)
#define RGFM_TC_LAUNCH2(D, R, F, S) do { if (xb) RGFM_TC_LAUNCH(D, R, true, F, S); else RGFM_TC_LAUNCH(D, R, false, F, S); } while (0)
  const bool final_out = a.out_mode != OUT_RAW;
  const bool regs = p.reg_stats && p.stats != nullptr && !final_out;
  const bool xb = p.any_xform && p.xf_bf16;
  if (final_out) {  // no statistics: one kernel per row space
    if (p.dense) RGFM_TC_LAUNCH2(true, false, true, false); else RGFM_TC_LAUNCH2(false, false, true, false);
  } else if (p.dense) {
    if (regs) RGFM_TC_LAUNCH2(true, true, false, false); else RGFM_TC_LAUNCH2(true, false, false, false);
  } else if (p.nstk > 1) {  // register statistics are never combined with stacked padded-linear units (make_tc_conv)
    RGFM_TC_LAUNCH2(false, false, false, true);
  } else {
    if (regs) RGFM_TC_LAUNCH2(false, true, false, false); else RGFM_TC_LAUNCH2(false, false, false, false);
  }
#undef RGFM_TC_LAUNCH2
#undef RGFM_TC_LAUNCH
  post_launch("conv_tc");
}

}  // namespace rgfm

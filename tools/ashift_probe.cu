// Hardware probe (B200): the A operand of tcgen05.mma held in TENSOR MEMORY for an implicit-GEMM convolution.
//   tcgen05.cp.128x256b   shared memory (SWIZZLE_128B K-major tile, arbitrary row shift) -> TMEM, one K16 slab (32 B per lane)
//   tcgen05.shift.down    moves the 32-byte rows of a slab by one lane (the next kx tap of a filter row)
//   tcgen05.mma [d], [a_tmem], b_desc   reads only the weights from shared memory
// Part 1 checks the semantics against a host model (which lane receives which row; what the shift does at the ends;
// whether cp -> mma -> shift -> mma issued back to back by one thread are ordered).  Part 2 times each instruction and
// the conv-like sequence (per filter row: 4 cp, then 3 x 4 MMAs with 2 x 4 shifts in between) at N = 32 / 64 / 128,
// next to the same taps issued the classic way (A from shared memory), alone and beside eight warps of ld.shared.v4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o ashift_probe tools/ashift_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ROWS = 256, K = 64, NMAX = 256, M = 128;
constexpr int THREADS = 32 * 13;  // warps 0..3 = TMEM readers (lane quarters), warp 4 = issuer, warps 5..12 = background
constexpr uint32_t ACOL = 256;    // TMEM columns of the A operand (4 slabs x 8 columns); the accumulator starts at column 0

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ uint64_t desc128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void shift_down(uint32_t taddr) {
  asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Shared {
  uint64_t bar_full, bar_op;
  uint32_t tmem_slot;
  volatile int go, stop;
};

// dump `ncols` (multiple of 32) TMEM columns starting at `col` of all 128 lanes to out[lane][ncols] (warps 0..3)
__device__ void dump(uint32_t tmem, uint32_t col, int ncols, uint32_t* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 4) {
    for (int c = 0; c < ncols; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + col + c, v);
      for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + lane) * ncols + c + j] = v[j];
    }
  }
}

// mode 0: semantics (dumps), mode 1..: timings (out64[0] = cycles)
__global__ void __launch_bounds__(THREADS, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                    int mode, int N, int reps, int bg, int shift_rows, uint32_t* dumps, long long* out64) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) Shared sh;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + ROWS * K * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.bar_full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.bar_op)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sh.go = 0;
    sh.stop = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sh.tmem_slot;
  if (threadIdx.x == 128) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sh.bar_full)), "r"((uint32_t)((ROWS + NMAX) * K * 2)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sA), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(&sh.bar_full)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sB), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&sh.bar_full)), "r"(0), "r"(0) : "memory");
  }
  mbar_wait(&sh.bar_full, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // fp16 x fp16 -> fp32
  const uint32_t a_addr = sA + (uint32_t)shift_rows * 128u;
  uint32_t ph = 0;

  if (mode == 0) {
    // ---- stage 0: zero the A columns + neighbours through an MMA with zero... simply dump what cp writes ----
    // stage 1: cp 4 slabs
    if (threadIdx.x == 128) {
      for (int k = 0; k < 4; ++k) cp_128x256b(tmem + ACOL + 8u * k, desc128(a_addr + 32u * k));
      commit(&sh.bar_op);
    }
    mbar_wait(&sh.bar_op, ph); ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    dump(tmem, ACOL, 32, dumps);  // dump 0: A after cp
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // stage 2: TS MMA (K = 64 as 4 MMAs), no wait after the cp is needed here (already waited); D columns 0..N
    if (threadIdx.x == 128) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int k = 0; k < 4; ++k) mma_ts(tmem, tmem + ACOL + 8u * k, desc128(sB + 32u * k), idesc, k ? 1u : 0u);
      commit(&sh.bar_op);
    }
    mbar_wait(&sh.bar_op, ph); ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    dump(tmem, 0, 64, dumps + 128 * 32);  // dump 1: D of the TS MMA (64 columns)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // stage 3: one shift of every slab
    if (threadIdx.x == 128) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int k = 0; k < 4; ++k) shift_down(tmem + ACOL + 8u * k);
      commit(&sh.bar_op);
    }
    mbar_wait(&sh.bar_op, ph); ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    dump(tmem, ACOL, 32, dumps + 128 * 32 + 128 * 64);  // dump 2: A after one shift
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // stage 4: back-to-back, no waits in between: cp (fresh) -> MMA (tap 0) -> shift -> MMA acc (tap 1) -> shift -> MMA acc (tap 2),
    // the three taps use B rows [0, N), [N, 2N), [2N, 3N)  (N <= 64 so that 3N <= NMAX... rows of the B tile)
    if (threadIdx.x == 128) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int k = 0; k < 4; ++k) cp_128x256b(tmem + ACOL + 8u * k, desc128(a_addr + 32u * k));
      for (int tap = 0; tap < 3; ++tap) {
        for (int k = 0; k < 4; ++k)
          mma_ts(tmem + 64u, tmem + ACOL + 8u * k, desc128(sB + (uint32_t)(tap * N) * 128u + 32u * k), idesc, (tap | k) ? 1u : 0u);
        if (tap < 2)
          for (int k = 0; k < 4; ++k) shift_down(tmem + ACOL + 8u * k);
      }
      commit(&sh.bar_op);
    }
    mbar_wait(&sh.bar_op, ph); ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    dump(tmem, 64, 64, dumps + 128 * 32 + 128 * 64 + 128 * 32);  // dump 3: D of the chained sequence
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  } else {
    // ---- timings ----
    if (warp >= 5 && bg != 0) {
      const uint32_t scratch = sB + NMAX * K * 2 + (threadIdx.x - 160) * 16;
      while (!sh.go) {}
      long long iters = 0;
      uint32_t acc0 = 0, acc1 = 0;
      while (!sh.stop) {
        uint32_t v[16][4];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i][0]), "=r"(v[i][1]), "=r"(v[i][2]), "=r"(v[i][3]) : "r"(scratch + (i & 7) * 4608));
#pragma unroll
        for (int i = 0; i < 16; ++i) { acc0 ^= v[i][0] ^ v[i][2]; acc1 ^= v[i][1] ^ v[i][3]; }
        ++iters;
      }
      if (lane == 0) out64[2 + warp] = iters + (acc0 == 0x12345u) + (acc1 == 7u);
    }
    if (threadIdx.x == 128) {
      sh.go = 1;
      const long long t0 = clock64();
      if (mode == 1) {  // cp only
        for (int r = 0; r < reps; ++r)
          for (int k = 0; k < 4; ++k) cp_128x256b(tmem + ACOL + 8u * k, desc128(a_addr + 32u * k));
      } else if (mode == 2) {  // shift only
        for (int r = 0; r < reps; ++r)
          for (int k = 0; k < 4; ++k) shift_down(tmem + ACOL + 8u * k);
      } else if (mode == 3) {  // TS MMA only
        for (int r = 0; r < reps; ++r)
          for (int k = 0; k < 4; ++k) mma_ts(tmem, tmem + ACOL + 8u * k, desc128(sB + 32u * k), idesc, 1u);
      } else if (mode == 4) {  // SS MMA only
        for (int r = 0; r < reps; ++r)
          for (int k = 0; k < 4; ++k) mma_ss(tmem, desc128(a_addr + 32u * k), desc128(sB + 32u * k), idesc, 1u);
      } else if (mode == 5) {  // conv filter row through TMEM: 4 cp + 3 x 4 MMA + 2 x 4 shift  (12 MMAs per rep)
        for (int r = 0; r < reps; ++r) {
          for (int k = 0; k < 4; ++k) cp_128x256b(tmem + ACOL + 8u * k, desc128(a_addr + 32u * k));
          for (int tap = 0; tap < 3; ++tap) {
            for (int k = 0; k < 4; ++k) mma_ts(tmem, tmem + ACOL + 8u * k, desc128(sB + 32u * k), idesc, 1u);
            if (tap < 2)
              for (int k = 0; k < 4; ++k) shift_down(tmem + ACOL + 8u * k);
          }
        }
      } else if (mode == 6) {  // the same filter row the classic way: 12 SS MMAs
        for (int r = 0; r < reps; ++r)
          for (int tap = 0; tap < 3; ++tap)
            for (int k = 0; k < 4; ++k) mma_ss(tmem, desc128(a_addr + (uint32_t)tap * 128u + 32u * k), desc128(sB + 32u * k), idesc, 1u);
      } else if (mode == 7) {  // three cp per filter row (one per kx shift) and no shift instruction: 12 cp + 12 TS MMAs
        for (int r = 0; r < reps; ++r)
          for (int tap = 0; tap < 3; ++tap) {
            for (int k = 0; k < 4; ++k) cp_128x256b(tmem + ACOL + 32u * (tap & 1) + 8u * k, desc128(a_addr + (uint32_t)tap * 128u + 32u * k));
            for (int k = 0; k < 4; ++k) mma_ts(tmem, tmem + ACOL + 32u * (tap & 1) + 8u * k, desc128(sB + 32u * k), idesc, 1u);
          }
      } else if (mode == 8) {  // mode 5 with the slab order interleaved: cp k, then per tap: MMA k / shift k alternating
        for (int r = 0; r < reps; ++r) {
          for (int k = 0; k < 4; ++k) cp_128x256b(tmem + ACOL + 8u * k, desc128(a_addr + 32u * k));
          for (int tap = 0; tap < 3; ++tap)
            for (int k = 0; k < 4; ++k) {
              mma_ts(tmem, tmem + ACOL + 8u * k, desc128(sB + 32u * k), idesc, 1u);
              if (tap < 2) shift_down(tmem + ACOL + 8u * k);
            }
        }
      }
      if (reps == 0) while (clock64() - t0 < 400000) {}
      commit(&sh.bar_op);
      mbar_wait(&sh.bar_op, 0);
      out64[0] = clock64() - t0;
      sh.stop = 1;
    }
    mbar_wait(&sh.bar_op, 0);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  CK(cudaSetDevice(0));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  std::vector<__half> hA(ROWS * K), hB(NMAX * K);
  std::vector<float> fA(ROWS * K), fB(NMAX * K);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)((rand() % 9) - 4); hA[i] = __float2half(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)((rand() % 7) - 3); hB[i] = __float2half(fB[i]); }
  __half *dA, *dB; long long* dOut; uint32_t* dDump;
  const size_t dump_words = 128 * 32 + 128 * 64 + 128 * 32 + 128 * 64;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dOut, 128)); CK(cudaMalloc(&dDump, dump_words * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[2] = {K, ROWS}; cuuint64_t str[1] = {K * 2}; cuuint32_t box[2] = {K, ROWS}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {K, NMAX}; cuuint64_t str[1] = {K * 2}; cuuint32_t box[2] = {K, NMAX}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  const size_t smem = (ROWS + NMAX) * K * 2 + 2048 + 48 * 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  // ---------------- part 1: semantics ----------------
  for (int shift_rows : {0, 35}) {
    const int N = 64;
    CK(cudaMemset(dDump, 0xff, dump_words * 4));
    probe<<<1, THREADS, smem>>>(tmA, tmB, 0, N, 0, 0, shift_rows, dDump, dOut);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("semantics kernel error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> d(dump_words);
    CK(cudaMemcpy(d.data(), dDump, dump_words * 4, cudaMemcpyDeviceToHost));
    const uint32_t* d0 = d.data();
    const uint32_t* d1 = d0 + 128 * 32;
    const uint32_t* d2 = d1 + 128 * 64;
    const uint32_t* d3 = d2 + 128 * 32;
    auto a_word = [&](int row, int c) {  // packed fp16 pair (k = 2c, 2c + 1) of tile row `row`
      uint16_t lo, hi;
      memcpy(&lo, &hA[(size_t)row * K + 2 * c], 2);
      memcpy(&hi, &hA[(size_t)row * K + 2 * c + 1], 2);
      return (uint32_t)lo | ((uint32_t)hi << 16);
    };
    // dump 0: lane l <- row shift_rows + l ?
    int bad = 0;
    for (int l = 0; l < 128; ++l)
      for (int c = 0; c < 32; ++c)
        if (d0[l * 32 + c] != a_word(shift_rows + l, c)) ++bad;
    printf("[row shift %2d] cp.128x256b: lane l == tile row (shift + l), column c == channels (2c, 2c+1): %s (%d mismatching words)\n", shift_rows, bad ? "NO" : "yes", bad);
    // dump 1: D[l][n] = sum_k A[shift + l][k] * B[n][k]
    bad = 0;
    for (int l = 0; l < 128; ++l)
      for (int n = 0; n < N; ++n) {
        float ref = 0.f;
        for (int k = 0; k < K; ++k) ref += fA[(size_t)(shift_rows + l) * K + k] * fB[(size_t)n * K + k];
        float got;
        memcpy(&got, &d1[l * 64 + n], 4);
        if (got != ref) ++bad;
      }
    printf("[row shift %2d] TS MMA (A in TMEM) == reference GEMM: %s (%d mismatches)\n", shift_rows, bad ? "NO" : "yes", bad);
    // dump 2: after one shift, which source lane does lane l hold?
    int from_next = 0, from_prev = 0, same = 0, other = 0;
    int first_other = -1;
    for (int l = 0; l < 128; ++l) {
      auto eq = [&](int src) {
        if (src < 0 || src > 127) return false;
        for (int c = 0; c < 32; ++c)
          if (d2[l * 32 + c] != d0[src * 32 + c]) return false;
        return true;
      };
      if (eq(l + 1)) ++from_next;
      else if (eq(l - 1)) ++from_prev;
      else if (eq(l)) ++same;
      else { ++other; if (first_other < 0) first_other = l; }
    }
    printf("[row shift %2d] after tcgen05.shift.down: lanes holding old lane l+1: %d, old lane l-1: %d, unchanged: %d, other: %d (first %d)\n",
           shift_rows, from_next, from_prev, same, other, first_other);
    for (int l : {0, 1, 31, 32, 33, 63, 64, 126, 127}) {
      const char* what = "other";
      auto eq = [&](int src) {
        if (src < 0 || src > 127) return false;
        for (int c = 0; c < 32; ++c)
          if (d2[l * 32 + c] != d0[src * 32 + c]) return false;
        return true;
      };
      if (eq(l + 1)) what = "old l+1";
      else if (eq(l - 1)) what = "old l-1";
      else if (eq(l)) what = "unchanged";
      printf("    lane %3d: %s\n", l, what);
    }
    // dump 3: chained cp -> MMA -> shift -> MMA -> shift -> MMA without waits; D[l][n] = sum_tap sum_k A[shift + l + s*tap][k] * B[tap*N + n][k]
    for (int s : {1, -1}) {
      bad = 0;
      int bad_lanes_hi = 0;
      for (int l = 0; l < 128; ++l) {
        bool lane_bad = false;
        for (int n = 0; n < N; ++n) {
          float ref = 0.f;
          for (int tap = 0; tap < 3; ++tap) {
            const int row = shift_rows + l + s * tap;
            if (row < 0 || row >= ROWS) continue;
            for (int k = 0; k < K; ++k) ref += fA[(size_t)row * K + k] * fB[(size_t)(tap * N + n) * K + k];
          }
          float got;
          memcpy(&got, &d3[l * 64 + n], 4);
          if (got != ref) { ++bad; lane_bad = true; }
        }
        if (lane_bad && (l >= 126 || l < 2)) ++bad_lanes_hi;
      }
      printf("[row shift %2d] chained cp/MMA/shift/MMA/shift/MMA == conv row with lane shift %+d per tap: %d mismatching elements (%d of the lanes are edge lanes 0,1,126,127)\n",
             shift_rows, s, bad, bad_lanes_hi);
    }
  }

  // ---------------- part 2: timings ----------------
  const char* mname[9] = {"", "cp.128x256b", "shift.down", "TS MMA", "SS MMA", "filter row via TMEM (4 cp + 12 MMA + 8 shift)",
                          "filter row classic (12 SS MMA)", "filter row 12 cp + 12 TS MMA", "filter row via TMEM, MMA/shift interleaved"};
  for (int bg : {0, 1})
    for (int mode = 1; mode <= 8; ++mode)
      for (int N : {32, 64, 128}) {
        if ((mode == 1 || mode == 2) && N != 64) continue;
        const int reps = 512;
        CK(cudaMemset(dOut, 0, 128));
        long long best = 1ll << 60;
        double bgrate = 0;
        for (int it = 0; it < 2; ++it) {
          probe<<<1, THREADS, smem>>>(tmA, tmB, mode, N, reps, bg, 35, dDump, dOut);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("kernel error %s (mode %d)\n", cudaGetErrorString(e), mode); return 1; }
          long long h[16]; CK(cudaMemcpy(h, dOut, 128, cudaMemcpyDeviceToHost));
          if (h[0] < best) {
            best = h[0];
            double tot = 0;
            for (int w = 5; w <= 12; ++w) tot += (double)h[2 + w];
            bgrate = tot * 16 / (double)h[0];
          }
        }
        const double per = (mode <= 4) ? 4.0 : 12.0;
        printf("%-48s N=%3d %s: %7.1f cycles per %s", mname[mode], N, bg ? "beside 8 warps of ld.shared.v4" : "alone", (double)best / (reps * per),
               mode <= 4 ? "instruction" : "MMA (tap x K16)");
        if (bg) printf("; background %.3f warp-instructions/clk/SM", bgrate);
        printf("\n");
      }
  {  // background alone
    CK(cudaMemset(dOut, 0, 128));
    probe<<<1, THREADS, smem>>>(tmA, tmB, 1, 64, 0, 1, 35, dDump, dOut);
    CK(cudaDeviceSynchronize());
    long long h[16]; CK(cudaMemcpy(h, dOut, 128, cudaMemcpyDeviceToHost));
    double tot = 0;
    for (int w = 5; w <= 12; ++w) tot += (double)h[2 + w];
    printf("background alone: %.3f warp-instructions/clk/SM\n", tot * 16 / (double)h[0]);
  }
  return 0;
}

// Hardware probe (B200): cost of one tcgen05.mma (M=128, K=16, kind::f16, SWIZZLE_128B K-major operands in shared memory) as
// a function of N and of the ROW SHIFT of the A descriptor's start address -- the halo-reuse convolution reads its nine
// taps as row shifts of one tile, and 8 of the 9 start off the 8-row swizzle atom.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate_probe tools/umma_rate_probe.cu -lcuda
// Prints cycles per MMA (clock64 around `reps` x 4 back-to-back MMAs + commit + wait, one CTA).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ROWS = 256, K = 64, NMAX = 256, M = 128;  // ROWS is the TMA box; shifted / second-half reads run past it into B (timing only)

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

__global__ void __launch_bounds__(256, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                int shift, int N, int reps, int mode, long long* out) {
  // mode 0: one issuing thread; 1: + 3 warps streaming ld/st.shared.v4 over 32 KB; 2: two issuing threads, two accumulators
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + ROWS * K * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_mma)), "r"(mode == 2 ? 2u : 1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;
  __shared__ volatile int s_go, s_stop;
  if (threadIdx.x == 0) { s_go = 0; s_stop = 0; }
  __syncthreads();
  if ((mode == 5 || mode == 6) && warp >= 1) {
    // heavy, INDEPENDENT shared-memory traffic from seven warps (mode 5: 16 ld.shared.v4 in flight per thread, mode 6:
    // st.shared.v4): does the MMA stream slow down when the LSU side wants the whole shared-memory bandwidth?
    const uint32_t scratch = sB + NMAX * K * 2 + (threadIdx.x - 32) * 16;
    while (!s_go) {}
    long long iters = 0;
    uint32_t acc0 = 0, acc1 = 0;
    while (!s_stop) {
      if (mode == 5) {
        uint32_t v[16][4];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i][0]), "=r"(v[i][1]), "=r"(v[i][2]), "=r"(v[i][3]) : "r"(scratch + (i & 11) * 3584));
#pragma unroll
        for (int i = 0; i < 16; ++i) { acc0 ^= v[i][0] ^ v[i][2]; acc1 ^= v[i][1] ^ v[i][3]; }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(scratch + (i & 11) * 3584), "r"(acc0), "r"(acc1), "r"((uint32_t)i), "r"((uint32_t)iters) : "memory");
      }
      ++iters;
    }
    if (lane == 0) out[2 + warp] = iters + (acc0 == 0x12345u) + (acc1 == 7u);
  }
  if ((mode == 3 || mode == 4) && warp >= 1) {
    // How much do warps that SHARE the issuing thread's scheduler (warp 4: same sub-partition as warp 0) progress while
    // the MMA queue is kept full?  mode 3: ld/st.shared loop, mode 4: dependent FMA chain.  Iterations per warp -> out[2 + warp]
    const uint32_t scratch = sB + NMAX * K * 2 + (threadIdx.x - 32) * 16;
    while (!s_go) {}
    long long iters = 0;
    uint32_t a0 = 1, a1 = 2, a2 = 3, a3 = 4;
    float f = 1.0f;
    while (!s_stop) {
      if (mode == 3) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t ad = scratch + i * 4096;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(ad));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) f = fmaf(f, 1.0001f, 0.5f);
      }
      ++iters;
    }
    if (lane == 0) out[2 + warp] = iters + (a0 == 0x12345u) + (f == 3.f);
  }
  if (mode == 1 && warp >= 1 && warp < 4) {  // background shared-memory traffic on a region the MMAs do not touch
    const uint32_t scratch = sB + NMAX * K * 2 + (threadIdx.x - 32) * 16;
    while (!s_go) {}
    uint32_t a0 = 1, a1 = 2, a2 = 3, a3 = 4;
    while (!s_stop) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t ad = scratch + i * 1536;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(ad));
      }
    }
    if (a0 == 0x12345u) out[1] = a1;
  }
  if (mode == 2 && threadIdx.x == 32) {  // second issuing thread: rows 128.., accumulator columns 256..
    while (!s_go) {}
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a_addr = sA + (uint32_t)(shift + 128) * 128u;
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < K / 16; ++k) {
        const uint64_t da = (uint64_t)(((a_addr + k * 32) >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint64_t db = (uint64_t)(((sB + k * 32) >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d + 256u), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_full)), "r"((uint32_t)((ROWS + NMAX) * K * 2)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sA), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(&bar_full)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sB), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&bar_full)), "r"(0), "r"(0) : "memory");
    mbar_wait(&bar_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a_addr = sA + (uint32_t)shift * 128u;
    s_go = 1;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < K / 16; ++k) {
        const uint64_t da = (uint64_t)(((a_addr + k * 32) >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint64_t db = (uint64_t)(((sB + k * 32) >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
      }
    if (reps == 0) while (clock64() - t0 < 400000) {}  // baseline: the background warps alone for the same kind of interval
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    mbar_wait(&bar_mma, 0);
    out[0] = clock64() - t0;
    s_stop = 1;
  }
  mbar_wait(&bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
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
  std::vector<__nv_bfloat16> hA(ROWS * K), hB(NMAX * K);
  srand(1);
  for (auto& v : hA) v = __float2bfloat16((float)((rand() % 15) - 7) * 0.01f);
  for (auto& v : hB) v = __float2bfloat16((float)((rand() % 7) - 3) * 0.01f);
  __nv_bfloat16 *dA, *dB; long long* dOut;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dOut, 128));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[2] = {K, ROWS}; cuuint64_t str[1] = {K * 2}; cuuint32_t box[2] = {K, ROWS}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {K, NMAX}; cuuint64_t str[1] = {K * 2}; cuuint32_t box[2] = {K, NMAX}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  const size_t smem = (ROWS + NMAX) * K * 2 + 2048 + 48 * 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int reps = 64;
  for (int mode = 0; mode < 3; ++mode)
  for (int N : {32, 64, 128, 256}) {
    printf("mode %d N=%3d:", mode, N);
    for (int shift : {0, 1, 35, 70}) {
      long long best = 1ll << 60;
      for (int it = 0; it < 3; ++it) {
        probe<<<1, 256, smem>>>(tmA, tmB, shift, N, reps, mode, dOut);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
        long long c; CK(cudaMemcpy(&c, dOut, 8, cudaMemcpyDeviceToHost));
        if (c < best) best = c;
      }
      printf("  shift %2d: %6.1f", shift, (double)best / (reps * 4 * (mode == 2 ? 2 : 1)));
    }
    printf("   cycles/MMA\n");
  }
  // scheduler sharing: N = 64, full-rate MMA issue from warp 0 (sub-partition 0) for reps2 x 4 MMAs; iterations of warps 1..7
  for (int mode = 3; mode <= 4; ++mode)
    for (int reps2 : {0, 2048}) {
      CK(cudaMemset(dOut, 0, 128));
      probe<<<1, 256, smem>>>(tmA, tmB, 0, 64, reps2, mode, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[16]; CK(cudaMemcpy(h, dOut, 128, cudaMemcpyDeviceToHost));
      printf("mode %d (%s) %5d MMAs in %lld cycles; background iterations per 1000 cycles, warps 1..7 (warp 4 shares the issuer's scheduler):",
             mode, mode == 3 ? "ld/st.shared" : "fma chain", reps2 * 4, h[0]);
      for (int w = 1; w < 8; ++w) printf(" %.1f", 1000.0 * h[2 + w] / (double)h[0]);
      printf("\n");
    }
  for (int mode = 5; mode <= 6; ++mode)
    for (int reps2 : {0, 2048}) {
      CK(cudaMemset(dOut, 0, 128));
      probe<<<1, 256, smem>>>(tmA, tmB, 0, 64, reps2, mode, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[16]; CK(cudaMemcpy(h, dOut, 128, cudaMemcpyDeviceToHost));
      double tot = 0;
      for (int w = 1; w < 8; ++w) tot += (double)h[2 + w];
      printf("mode %d (7 warps of independent %s) %5d MMAs in %lld cycles = %.1f cycles/MMA; LSU side: %.1f bytes/clk\n", mode,
             mode == 5 ? "ld.shared.v4" : "st.shared.v4", reps2 * 4, h[0], reps2 ? (double)h[0] / (reps2 * 4) : 0.0,
             tot * 16 * 32 * 16 / (double)h[0]);
    }
  return 0;
}

// Hardware probe (B200) for the tap-stacked convolution (three kx taps stacked on N, lane-shifted sum in the epilogue):
//   1. cost of one tcgen05.mma (M=128, K=16, kind::f16, SWIZZLE_128B operands in shared memory) at N = 48, 96, 192
//   2. what eight "epilogue-like" warps get done beside a saturated MMA stream of that shape:
//      SHFL chains, tcgen05.ld, ld.shared.v4, st.shared.v4 (operations per clock per SM, with and without the MMAs)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o stack_probe tools/stack_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ROWS = 256, K = 64, NMAX = 256, M = 128;
constexpr int THREADS = 32 * 9;  // warp 0 = issuer, warps 1..8 = background

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

// bg: 0 none, 1 SHFL, 2 tcgen05.ld x16, 3 ld.shared.v4, 4 st.shared.v4, 5 MUFU.TANH + FMA (transform-like math)
__global__ void __launch_bounds__(THREADS, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                    int N, int reps, int bg, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + ROWS * K * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
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
  if (warp >= 1 && bg != 0) {
    const uint32_t scratch = sB + NMAX * K * 2 + (threadIdx.x - 32) * 16;
    while (!s_go) {}
    long long iters = 0;
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = (float)(lane + i);
    uint32_t acc0 = 0, acc1 = 0;
    const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + 256u;  // columns the MMAs do not write
    while (!s_stop) {
      if (bg == 1) {
        // 32 shuffles + adds per iteration, 16 independent chains
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] += __shfl_down_sync(0xffffffffu, f[i], 1);
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] += __shfl_down_sync(0xffffffffu, f[i], 2);
      } else if (bg == 2) {
        uint32_t v[4][16];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[q][0]), "=r"(v[q][1]), "=r"(v[q][2]), "=r"(v[q][3]), "=r"(v[q][4]), "=r"(v[q][5]), "=r"(v[q][6]), "=r"(v[q][7]), "=r"(v[q][8]),
                "=r"(v[q][9]), "=r"(v[q][10]), "=r"(v[q][11]), "=r"(v[q][12]), "=r"(v[q][13]), "=r"(v[q][14]), "=r"(v[q][15])
              : "r"(taddr + q * 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int i = 0; i < 16; ++i) acc0 ^= v[q][i];
      } else if (bg == 3) {
        uint32_t v[16][4];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i][0]), "=r"(v[i][1]), "=r"(v[i][2]), "=r"(v[i][3]) : "r"(scratch + (i & 7) * 4608));
#pragma unroll
        for (int i = 0; i < 16; ++i) { acc0 ^= v[i][0] ^ v[i][2]; acc1 ^= v[i][1] ^ v[i][3]; }
      } else if (bg == 4) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(scratch + (i & 7) * 4608), "r"(acc0), "r"(acc1), "r"((uint32_t)i), "r"((uint32_t)iters) : "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float t;
          const float z = fmaf(f[i], 0.37f, 0.11f);
          asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(z));
          f[i] = fmaf(z, t, z);
        }
      }
      ++iters;
    }
    float fs = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) fs += f[i];
    if (lane == 0) out[2 + warp] = iters + (acc0 == 0x12345u) + (acc1 == 7u) + (fs == 3.25f);
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
    const uint32_t a_addr = sA + 35u * 128u;  // an unaligned row shift, like a filter row of the halo tile
    s_go = 1;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < K / 16; ++k) {
        const uint64_t da = (uint64_t)(((a_addr + k * 32) >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint64_t db = (uint64_t)(((sB + k * 32) >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
      }
    if (reps == 0) while (clock64() - t0 < 400000) {}
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
  for (int N : {32, 48, 64, 96, 128, 192, 256}) {
    long long best = 1ll << 60;
    for (int it = 0; it < 3; ++it) {
      probe<<<1, THREADS, smem>>>(tmA, tmB, N, 64, 0, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
      long long c; CK(cudaMemcpy(&c, dOut, 8, cudaMemcpyDeviceToHost));
      if (c < best) best = c;
    }
    printf("MMA alone N=%3d: %6.1f cycles/MMA\n", N, (double)best / (64 * 4));
  }
  const char* names[6] = {"", "SHFL (32 per iter)", "tcgen05.ld x16 (4 per iter)", "ld.shared.v4 (16 per iter)", "st.shared.v4 (16 per iter)", "tanh+2fma (16 per iter)"};
  const double per_iter[6] = {0, 32, 4, 16, 16, 16};
  for (int bg = 1; bg <= 5; ++bg)
    for (int N : {0, 64, 96, 192}) {
      const int reps2 = N ? 2048 : 0;
      CK(cudaMemset(dOut, 0, 128));
      probe<<<1, THREADS, smem>>>(tmA, tmB, N ? N : 64, reps2, bg, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[16]; CK(cudaMemcpy(h, dOut, 128, cudaMemcpyDeviceToHost));
      double tot = 0;
      for (int w = 1; w <= 8; ++w) tot += (double)h[2 + w];
      printf("bg %-28s MMA N=%3d: %7.1f cycles/MMA; 8 warps: %.3f warp-instructions/clk/SM\n", names[bg], N, reps2 ? (double)h[0] / (reps2 * 4) : 0.0,
             tot * per_iter[bg] / (double)h[0]);
    }
  return 0;
}

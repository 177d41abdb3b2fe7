// Hardware probe (B200): can a K-major SWIZZLE_128B UMMA shared-memory descriptor start at an arbitrary 128-byte ROW
// offset inside a TMA-written tile?  If yes, the nine taps of a 3x3 convolution can read one haloed activation tile
// from shared memory instead of re-fetching it nine times from L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_shift_probe tools/umma_shift_probe.cu -lcuda
// Prints, per row shift r and per descriptor variant (base_offset = 0 | r%8), the max abs error vs a CPU reference.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ROWS = 256, K = 64, N = 16, M = 128;

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

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                int shift, int base_off, int kstep_elems, float* out) {
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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_full)), "r"((uint32_t)((ROWS + N) * K * 2)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sA), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(&bar_full)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sB), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&bar_full)), "r"(0), "r"(0) : "memory");
    mbar_wait(&bar_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a_addr = sA + (uint32_t)shift * 128u;
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t da = (uint64_t)(((a_addr + k * kstep_elems * 2) >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
                          ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
      const uint64_t db = (uint64_t)(((sB + k * 32) >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
      const uint32_t acc = k != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
  }
  mbar_wait(&bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[16];
  const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(32u) : "memory");
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
  std::vector<__nv_bfloat16> hA(ROWS * K), hB(N * K);
  std::vector<float> fA(ROWS * K), fB(N * K);
  srand(1);
  for (int i = 0; i < ROWS * K; ++i) { fA[i] = (float)((rand() % 15) - 7); hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < N * K; ++i) { fB[i] = (float)((rand() % 7) - 3); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB; float* dOut;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dOut, M * N * 4));
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
    cuuint64_t dims[2] = {K, N}; cuuint64_t str[1] = {K * 2}; cuuint32_t box[2] = {K, N}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  const size_t smem = (ROWS + N) * K * 2 + 2048;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> out(M * N);
  const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 17, 34, 35, 36, 68, 70, 100};
  for (int shift : shifts) {
    for (int variant = 0; variant < 2; ++variant) {
      const int bo = variant == 0 ? 0 : (shift & 7);
      CK(cudaMemset(dOut, 0, M * N * 4));
      probe<<<1, 128, smem>>>(tmA, tmB, shift, bo, 16, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d variant %d: kernel error %s\n", shift, variant, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(out.data(), dOut, M * N * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0; int bad = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)fA[(m + shift) * K + k] * fB[n * K + k];
          const double d = fabs(ref - out[m * N + n]);
          if (d > maxerr) maxerr = d;
          if (d > 0.5) ++bad;
        }
      printf("shift %3d  base_offset %d : max err %.1f  wrong %d / %d  %s\n", shift, bo, maxerr, bad, M * N, bad ? "MISMATCH" : "OK");
    }
  }
  return 0;
}

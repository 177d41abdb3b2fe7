// Microbenchmark: cost of mbarrier primitives on sm_100a as seen by ONE thread, alone and with other warps polling.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mbar_probe tools/mbar_probe.cu && tools/mbar_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) {
#pragma unroll 1
  while (!try_wait(bar, par)) {}
}
// mode 0: one thread: arrive + wait (complete) x N.  mode 1: ping-pong between warp 0 lane 0 and warp 1 lane 0.
// extra: number of additional warps that poll a never-completing barrier (all 32 lanes, or lane 0 only if lane0only)
__global__ void probe(int mode, int iters, int extra, int lane0only, int sleep_ns, long long* out) {
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ volatile int stop;
  const uint32_t b0 = smem_u32(&bars[0]), b1 = b0 + 8, b2 = b0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(b0, 1); mbar_init(b1, 1); mbar_init(b2, 1);
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= 2) {  // pollers
    if (warp - 2 < extra && (!lane0only || lane == 0)) {
      while (!stop) {
        try_wait(b2, 0);
        if (sleep_ns) __nanosleep(sleep_ns);
      }
    }
    return;
  }
  if (lane != 0) return;
  if (mode == 0) {
    if (warp == 0) {
      long long t0 = clock64();
      uint32_t ph = 0;
      for (int i = 0; i < iters; ++i) {
        mbar_arrive(b0);
        wait(b0, ph);
        ph ^= 1;
      }
      long long t1 = clock64();
      out[0] = (t1 - t0) / iters;
      stop = 1;
    }
  } else {
    uint32_t ph = 0;
    if (warp == 0) {
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        mbar_arrive(b0);   // signal consumer
        wait(b1, ph);      // wait for its answer
        ph ^= 1;
      }
      long long t1 = clock64();
      out[0] = (t1 - t0) / iters;
      stop = 1;
    } else {
      for (int i = 0; i < iters; ++i) {
        wait(b0, ph);
        mbar_arrive(b1);
        ph ^= 1;
      }
    }
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  long long h;
  for (int mode = 0; mode < 2; ++mode)
    for (int extra : {0, 8})
      for (int l0 : {0, 1})
        for (int sl : {0, 64}) {
          if (extra == 0 && (l0 || sl)) continue;
          probe<<<1, 32 * 10, 0>>>(mode, 2000, extra, l0, sl, d);
          cudaError_t e = cudaDeviceSynchronize();
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          printf("mode %d (%s) pollers %d %s sleep %d ns: %lld cycles per %s  [%s]\n", mode, mode ? "ping-pong 2 warps" : "1 thread arrive+wait", extra,
                 l0 ? "lane0" : "32 lanes", sl, h, mode ? "round trip" : "arrive+wait", cudaGetErrorString(e));
        }
  return 0;
}

// mc_feng guidance (reference flow_utils.py:273-369 == sample_mnist_svhn.py:124-171) fused
// with the blend and the Euler update, in fp32 on CUDA cores.  The reference materialises ~10 [B, N_mc, D]
// temporaries per step; here the block is three tiled kernels that never leave [B, N_mc] + [B, D]:
//   1. mcf_dist   : S[b,i] = sum_d (x[b,d] - fl(t*m[i,d]))^2   (register-tiled like an SGEMM, split over D)
//   2. mcf_weights: logits -> stabilised exp -> p_bar, Z_bar (+1e-10 each) -> w -> normalise (+1e-10)
//   3. mcf_apply  : g[b,d] = (sum_i w[b,i] (m[i,d] - x[b,d])) / sigma ; x += dt*((1-gamma) v + gamma g)
// The direct (x - t m)^2 form is kept (no ||x||^2 - 2 x.m expansion) so late-time logits, which are divided by
// sigma_t^2 ~ 1e-4, do not lose the bits the reference has.
#include "common.h"
#include <algorithm>

namespace rgfm {

static inline void post_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(std::string(what) + " launch failed: " + cudaGetErrorString(e));
  ++g_launch_counter;
}

constexpr int DT = 64;   // tile edge (b and i / d)
constexpr int DK = 16;   // reduction chunk
constexpr int MCF_KMAX = 8;  // longest non-zero weight list the sparse apply path handles

// sparse-weight scratch, laid out after zbuf[B][4]: max_cnt (4 ints), cnt[B], idx[B][KMAX], w[B][KMAX]
__host__ __device__ inline int* sparse_max(const McfArgs& a) { return reinterpret_cast<int*>(a.zbuf + (size_t)a.B * 4); }
__host__ __device__ inline int* sparse_cnt(const McfArgs& a) { return sparse_max(a) + 4; }
__host__ __device__ inline int* sparse_idx(const McfArgs& a) { return sparse_cnt(a) + a.B; }
__host__ __device__ inline float* sparse_w(const McfArgs& a) { return reinterpret_cast<float*>(sparse_idx(a) + (size_t)a.B * MCF_KMAX); }

// grid: (ceil(B/64), ceil(N/64), ksx + ksy).  z < ksx handles a D-range of x, else of y.
__global__ void __launch_bounds__(256) mcf_dist_kernel(McfArgs a) {
  __shared__ __align__(16) float Xs[DK][DT + 4];
  __shared__ __align__(16) float Ms[DK][DT + 4];
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * DT, i0 = blockIdx.y * DT;
  const int z = blockIdx.z;
  const bool isx = z < a.ksplit_x;
  const int D = isx ? a.Dx : a.Dy;
  const int ks = isx ? a.ksplit_x : a.ksplit_y;
  const int zz = isx ? z : z - a.ksplit_x;
  const int chunk = ((D / DK + ks - 1) / ks) * DK;
  const int d_begin = zz * chunk;
  const int d_end = min(D, d_begin + chunk);
  const float* X = isx ? a.x : a.y;
  const float* M = isx ? a.mcx : a.mcy;
  const float t = a.sc[a.step ? *a.step : 0].t;
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0) *sparse_max(a) = 0;  // re-armed for mcf_weights

  const int lb = tid >> 2, l4 = tid & 3;  // load assignment: row, float4 within the 16-wide chunk
  const int tb = tid >> 4, ti = tid & 15; // compute assignment: 4 b x 4 i
  float acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;

  for (int d0 = d_begin; d0 < d_end; d0 += DK) {
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), mv = xv;
    if (b0 + lb < a.B) xv = __ldg(reinterpret_cast<const float4*>(X + (size_t)(b0 + lb) * D + d0 + l4 * 4));
    if (i0 + lb < a.N) mv = __ldg(reinterpret_cast<const float4*>(M + (size_t)(i0 + lb) * D + d0 + l4 * 4));
    Xs[l4 * 4 + 0][lb] = xv.x; Xs[l4 * 4 + 1][lb] = xv.y; Xs[l4 * 4 + 2][lb] = xv.z; Xs[l4 * 4 + 3][lb] = xv.w;
    // mu = t * x1, rounded to fp32 exactly as the reference's tensor-by-scalar multiply (flow_utils.py:291)
    Ms[l4 * 4 + 0][lb] = __fmul_rn(t, mv.x); Ms[l4 * 4 + 1][lb] = __fmul_rn(t, mv.y);
    Ms[l4 * 4 + 2][lb] = __fmul_rn(t, mv.z); Ms[l4 * 4 + 3][lb] = __fmul_rn(t, mv.w);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < DK; ++k) {
      const float4 xq = *reinterpret_cast<const float4*>(&Xs[k][tb * 4]);
      const float4 mq = *reinterpret_cast<const float4*>(&Ms[k][ti * 4]);
      const float xa[4] = {xq.x, xq.y, xq.z, xq.w};
      const float ma[4] = {mq.x, mq.y, mq.z, mq.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float d = __fsub_rn(xa[p], ma[q]);
          acc[p][q] = fmaf(d, d, acc[p][q]);
        }
    }
    __syncthreads();
  }
  float* out = a.logits + (size_t)z * a.B * a.N;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int b = b0 + tb * 4 + p;
    if (b >= a.B) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + ti * 4 + q;
      if (i < a.N) out[(size_t)b * a.N + i] = acc[p][q];
    }
  }
}

__device__ __forceinline__ float block_reduce(float v, float* sh, int op) {  // op 0 sum, 1 max, 2 min; 128 threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = op == 0 ? v + u : (op == 1 ? fmaxf(v, u) : fminf(v, u));
  }
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < 4; ++w) r = op == 0 ? r + sh[w] : (op == 1 ? fmaxf(r, sh[w]) : fminf(r, sh[w]));
  return r;
}

// one CTA (128 threads) per sample b
__global__ void __launch_bounds__(128) mcf_weights_kernel(McfArgs a) {
  extern __shared__ float sw[];  // [N]
  __shared__ float sh[4];
  const int b = blockIdx.x, tid = threadIdx.x;
  const StepScalars sc = a.sc[a.step ? *a.step : 0];
  const size_t BN = (size_t)a.B * a.N;
  float mx = -INFINITY;
  for (int i = tid; i < a.N; i += 128) {
    float sx = 0.f, sy = 0.f;
    for (int z = 0; z < a.ksplit_x; ++z) sx += a.logits[z * BN + (size_t)b * a.N + i];
    for (int z = 0; z < a.ksplit_y; ++z) sy += a.logits[(a.ksplit_x + z) * BN + (size_t)b * a.N + i];
    // log_p = -0.5 * sum / sigma^2 per modality, then added (flow_utils.py:300-302)
    const float l = __fadd_rn(__fdiv_rn(__fmul_rn(-0.5f, sx), sc.sig2), __fdiv_rn(__fmul_rn(-0.5f, sy), sc.sig2));
    sw[i] = l;
    mx = fmaxf(mx, l);
  }
  mx = block_reduce(mx, sh, 1);
  float sp = 0.f, srp = 0.f;
  for (int i = tid; i < a.N; i += 128) {
    const float p = expf(sw[i] - mx);
    sw[i] = p;
    sp += p;
    srp = fmaf(a.ratios[i], p, srp);
  }
  sp = block_reduce(sp, sh, 0);
  srp = block_reduce(srp, sh, 0);
  const float p_bar = sp / (float)a.N + 1e-10f;
  const float z_bar = srp / (float)a.N + 1e-10f;
  float swt = 0.f;
  for (int i = tid; i < a.N; i += 128) {
    const float w = __fmul_rn(__fdiv_rn(a.ratios[i], z_bar), __fdiv_rn(sw[i], p_bar));
    sw[i] = w;
    swt += w;
  }
  swt = block_reduce(swt, sh, 0);
  const float den = swt + 1e-10f;
  float wmin = INFINITY, wmax = -INFINITY;
  for (int i = tid; i < a.N; i += 128) {
    const float w = __fdiv_rn(sw[i], den);
    a.wbuf[(size_t)b * a.N + i] = w;
    if (a.w) a.w[(size_t)b * a.N + i] = w;
    wmin = fminf(wmin, w);
    wmax = fmaxf(wmax, w);
  }
  wmin = block_reduce(wmin, sh, 2);
  wmax = block_reduce(wmax, sh, 1);
  if (tid == 0) {
    a.zbuf[b * 4 + 0] = z_bar;
    a.zbuf[b * 4 + 1] = wmin;
    a.zbuf[b * 4 + 2] = wmax;
  }
  // Once the softmax saturates almost every weight is EXACTLY zero (exp underflow), so the guidance sum has a handful of
  // terms per sample: warp 0 compacts the non-zero (index, weight) pairs in increasing index order (deterministic), and
  // the apply kernel walks the list instead of the dense [B, N_mc] x [N_mc, D] product.
  __syncthreads();
  if (tid < 32) {
    int* nzc = sparse_cnt(a);
    int* nzi = sparse_idx(a) + (size_t)b * MCF_KMAX;
    float* nzw = sparse_w(a) + (size_t)b * MCF_KMAX;
    int cnt = 0;
    for (int base = 0; base < a.N; base += 32) {
      const int i = base + tid;
      const float wv = (i < a.N) ? sw[i] / den : 0.f;  // the value written to wbuf above (same __fdiv_rn)
      const unsigned m = __ballot_sync(0xffffffffu, wv != 0.f);
      const int pos = cnt + __popc(m & ((1u << tid) - 1u));
      if (wv != 0.f && pos < MCF_KMAX) {
        nzi[pos] = i;
        nzw[pos] = wv;
      }
      cnt += __popc(m);
    }
    if (tid == 0) {
      nzc[b] = cnt;
      if (cnt > MCF_KMAX) atomicMax(sparse_max(a), cnt);  // any long row -> the dense kernel handles this launch
    }
  }
}

// grid: (ceil(B/64), ceil(D/64)); one launch per modality
struct ApplyArgs {
  int B, N, D;
  const float *x, *v, *m, *w;
  const StepScalars* sc;
  const int* step;
  float *x_out, *g;
  const int* max_cnt;   // > MCF_KMAX: some sample has a long weight list -> dense path; else sparse path
  const int* nz_cnt;
  const int* nz_idx;
  const float* nz_w;
  int g_step;           // write g only at this step (< 0: always)
};

// Sparse path of the fused guidance sum + blend + Euler update: a pure stream over [B, D] (read x, v; write x) with the
// few non-zero bank rows gathered from L2.  Terms are accumulated in increasing index order with the dense kernel's exact
// arithmetic (an fma with a zero weight leaves the accumulator unchanged), so both paths give the same bits.
__global__ void __launch_bounds__(256) mcf_apply_sparse_kernel(ApplyArgs ax, ApplyArgs ay) {
  if (*ax.max_cnt > MCF_KMAX) return;  // the dense kernel owns this launch
  const StepScalars sc = ax.sc[ax.step ? *ax.step : 0];
  const int Dx4 = ax.D >> 2, D4t = (ax.D + ay.D) >> 2;  // both modalities in one element space: [b][x quads | y quads]
  const size_t total = (size_t)ax.B * D4t;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / D4t);
    const int q4 = (int)(e - (size_t)b * D4t);
    const bool isx = q4 < Dx4;
    const ApplyArgs& a = isx ? ax : ay;
    const int d = (isx ? q4 : q4 - Dx4) << 2;
    const bool write_g = a.g != nullptr && (a.g_step < 0 || (a.step ? *a.step : 0) == a.g_step);
    const float4 xv = *reinterpret_cast<const float4*>(a.x + (size_t)b * a.D + d);
    const float4 vv = __ldg(reinterpret_cast<const float4*>(a.v + (size_t)b * a.D + d));
    const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
    const float va[4] = {vv.x, vv.y, vv.z, vv.w};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int cnt = a.nz_cnt[b];
    for (int k = 0; k < cnt; ++k) {
      const int i = a.nz_idx[(size_t)b * MCF_KMAX + k];
      const float w = a.nz_w[(size_t)b * MCF_KMAX + k];
      const float4 mv = __ldg(reinterpret_cast<const float4*>(a.m + (size_t)i * a.D + d));
      const float ma[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(w, __fsub_rn(ma[q], xr[q]), acc[q]);
    }
    float go[4], xo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      go[q] = __fmul_rn(acc[q], sc.inv_sigma);
      const float vg = __fadd_rn(__fmul_rn(sc.one_minus_gamma, va[q]), __fmul_rn(sc.gamma, go[q]));
      xo[q] = __fadd_rn(xr[q], __fmul_rn(vg, sc.dt));
    }
    *reinterpret_cast<float4*>(a.x_out + (size_t)b * a.D + d) = make_float4(xo[0], xo[1], xo[2], xo[3]);
    if (write_g) *reinterpret_cast<float4*>(a.g + (size_t)b * a.D + d) = make_float4(go[0], go[1], go[2], go[3]);
  }
}

__global__ void __launch_bounds__(256) mcf_apply_kernel(ApplyArgs a) {
  if (*a.max_cnt <= MCF_KMAX) return;  // every sample has a short weight list: the sparse kernel owns this launch
  __shared__ __align__(16) float Ws[DK][DT + 4];  // [i][b]
  __shared__ __align__(16) float Ms[DK][DT];      // [i][d]
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * DT, d0 = blockIdx.y * DT;
  const int tb = tid >> 4, td = tid & 15;
  const StepScalars sc = a.sc[a.step ? *a.step : 0];

  float xr[4][4], acc[4][4];
  const int dcol = d0 + td * 4;
  const bool dok = dcol < a.D;  // D % 4 == 0
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int b = b0 + tb * 4 + p;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dok && b < a.B) xv = *reinterpret_cast<const float4*>(a.x + (size_t)b * a.D + dcol);
    xr[p][0] = xv.x; xr[p][1] = xv.y; xr[p][2] = xv.z; xr[p][3] = xv.w;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
  }
  const int lb = tid >> 2, l4 = tid & 3;   // W load: row b, float4 over i
  const int mi = tid >> 4, m4 = tid & 15;  // M load: row i, float4 over d
  for (int i0 = 0; i0 < a.N; i0 += DK) {
    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b0 + lb < a.B) {
      const float* wp = a.w + (size_t)(b0 + lb) * a.N + i0 + l4 * 4;
      if ((a.N & 3) == 0 && i0 + l4 * 4 + 3 < a.N) {
        wv = *reinterpret_cast<const float4*>(wp);
      } else if (i0 + l4 * 4 + 3 < a.N) {
        wv = make_float4(wp[0], wp[1], wp[2], wp[3]);
      } else {
        if (i0 + l4 * 4 + 0 < a.N) wv.x = wp[0];
        if (i0 + l4 * 4 + 1 < a.N) wv.y = wp[1];
        if (i0 + l4 * 4 + 2 < a.N) wv.z = wp[2];
      }
    }
    const int any = __syncthreads_or((wv.x != 0.f) | (wv.y != 0.f) | (wv.z != 0.f) | (wv.w != 0.f));
    if (!any) continue;  // all 64 x 16 weights are exactly zero (saturated softmax): nothing to add
    Ws[l4 * 4 + 0][lb] = wv.x; Ws[l4 * 4 + 1][lb] = wv.y; Ws[l4 * 4 + 2][lb] = wv.z; Ws[l4 * 4 + 3][lb] = wv.w;
    float4 mv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 + mi < a.N && d0 + m4 * 4 < a.D) mv = __ldg(reinterpret_cast<const float4*>(a.m + (size_t)(i0 + mi) * a.D + d0 + m4 * 4));
    *reinterpret_cast<float4*>(&Ms[mi][m4 * 4]) = mv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < DK; ++k) {
      const float4 wq = *reinterpret_cast<const float4*>(&Ws[k][tb * 4]);
      const float4 mq = *reinterpret_cast<const float4*>(&Ms[k][td * 4]);
      const float wa[4] = {wq.x, wq.y, wq.z, wq.w};
      const float ma[4] = {mq.x, mq.y, mq.z, mq.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(wa[p], __fsub_rn(ma[q], xr[p][q]), acc[p][q]);
    }
    __syncthreads();
  }
  if (!dok) return;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int b = b0 + tb * 4 + p;
    if (b >= a.B) continue;
    const float4 vv = *reinterpret_cast<const float4*>(a.v + (size_t)b * a.D + dcol);
    const float va[4] = {vv.x, vv.y, vv.z, vv.w};
    float go[4], xo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      go[q] = __fmul_rn(acc[p][q], sc.inv_sigma);
      // v = (1-gamma) v + gamma g ; x = x + v*dt   (flow_utils.py:368-373; mul/add, no contraction)
      const float vg = __fadd_rn(__fmul_rn(sc.one_minus_gamma, va[q]), __fmul_rn(sc.gamma, go[q]));
      xo[q] = __fadd_rn(xr[p][q], __fmul_rn(vg, sc.dt));
    }
    *reinterpret_cast<float4*>(a.x_out + (size_t)b * a.D + dcol) = make_float4(xo[0], xo[1], xo[2], xo[3]);
    if (a.g != nullptr && (a.g_step < 0 || (a.step ? *a.step : 0) == a.g_step))
      *reinterpret_cast<float4*>(a.g + (size_t)b * a.D + dcol) = make_float4(go[0], go[1], go[2], go[3]);
  }
}

size_t mcf_scratch_floats(int B, int N, int Dx, int Dy, int* ksx, int* ksy) {
  // split the D reduction so the distance kernel launches >= ~2 waves of CTAs on 148 SMs
  const int tiles = ((B + DT - 1) / DT) * ((N + DT - 1) / DT);
  int want = (2 * 148 + tiles - 1) / tiles;
  if (want < 2) want = 2;
  if (want > 16) want = 16;
  const int unit = Dx < Dy ? Dx : Dy;
  // proportional split, each split a multiple of DK
  int kx = (int)((double)want * Dx / (Dx + Dy) + 0.5), ky = want - kx;
  if (kx < 1) kx = 1;
  if (ky < 1) ky = 1;
  while ((Dx / DK) < kx) --kx;
  while ((Dy / DK) < ky) --ky;
  (void)unit;
  *ksx = kx;
  *ksy = ky;
  // logits partials, weights, zbuf, then the sparse lists: 4 + B + 2 * B * KMAX words
  return (size_t)(kx + ky) * B * N + (size_t)B * N + (size_t)B * 4 + 4 + (size_t)B * (1 + 2 * MCF_KMAX);
}

void launch_mcf(const McfArgs& a, cudaStream_t s, std::vector<ProfRec>* prof) {
  launch_mcf_prep(a, s, prof);
  launch_mcf_apply(a, s, prof);
}

void launch_mcf_prep(const McfArgs& a, cudaStream_t s, std::vector<ProfRec>* prof) {
  RGFM_CHECK(a.Dx % DK == 0 && a.Dy % DK == 0, "mc_feng: D must be a multiple of 16");
  RGFM_CHECK(a.N * sizeof(float) <= 48 * 1024, "mc_feng: mc_batch_size too large (max 12288)");
  const double BN = (double)a.B * a.N, D = (double)a.Dx + a.Dy;
  {
    // algorithmic: 3 flops per (b, i, d); reads x, y once and the bank once, writes the partial sums
    ProfScope ps(prof, s, "mcf_dist", 3.0 * BN * D, 4.0 * (a.B * D + a.N * D + BN * (a.ksplit_x + a.ksplit_y)));
    dim3 g1((a.B + DT - 1) / DT, (a.N + DT - 1) / DT, a.ksplit_x + a.ksplit_y);
    mcf_dist_kernel<<<g1, 256, 0, s>>>(a);
    post_launch("mcf_dist");
  }
  {
    ProfScope ps(prof, s, "mcf_weights", 10.0 * BN, 4.0 * BN * (a.ksplit_x + a.ksplit_y + 1));
    mcf_weights_kernel<<<a.B, 128, a.N * sizeof(float), s>>>(a);
    post_launch("mcf_weights");
  }
}

void launch_mcf_apply(const McfArgs& a, cudaStream_t s, std::vector<ProfRec>* prof) {
  const double BN = (double)a.B * a.N, D = (double)a.Dx + a.Dy;
  {
    // fused guidance-sum + blend + Euler update: algorithmic bytes = read x, v, write x (+ bank and weights once)
    // one of the two paths runs per launch (device-side switch on the longest weight list, no host round trip); the
    // profile times them separately and credits the algorithmic work to whichever did it (bench.py: the longer one)
    ApplyArgs ax{a.B, a.N, a.Dx, a.x, a.vx, a.mcx, a.wbuf, a.sc, a.step, a.x_out, a.gx,
                 sparse_max(a), sparse_cnt(a), sparse_idx(a), sparse_w(a), a.g_step};
    ApplyArgs ay{a.B, a.N, a.Dy, a.y, a.vy, a.mcy, a.wbuf, a.sc, a.step, a.y_out, a.gy,
                 sparse_max(a), sparse_cnt(a), sparse_idx(a), sparse_w(a), a.g_step};
    {
      ProfScope ps(prof, s, "mcf_apply_dense", 2.0 * BN * D, 4.0 * (3.0 * a.B * D + a.N * D + BN));
      dim3 g3x((a.B + DT - 1) / DT, (a.Dx + DT - 1) / DT);
      mcf_apply_kernel<<<g3x, 256, 0, s>>>(ax);
      post_launch("mcf_apply_x");
      dim3 g3y((a.B + DT - 1) / DT, (a.Dy + DT - 1) / DT);
      mcf_apply_kernel<<<g3y, 256, 0, s>>>(ay);
      post_launch("mcf_apply_y");
    }
    // sparse path: read x, v once, write x once (+ the few gathered bank rows from L2)
    ProfScope ps(prof, s, "mcf_apply", 2.0 * a.B * D * MCF_KMAX, 4.0 * 3.0 * a.B * D);
    const size_t e4 = (size_t)a.B * (a.Dx + a.Dy) / 4;
    const unsigned sb = (unsigned)std::min<size_t>((e4 + 255) / 256, 148 * 8);
    mcf_apply_sparse_kernel<<<sb, 256, 0, s>>>(ax, ay);
    post_launch("mcf_apply_sparse");
  }
}

// Diagnostics of the 28x28 reference sampler (flow_utils.py:349-363), written to a device buffer at
// step == int(0.3*num_steps) so the loop stays free of host synchronisation.
__global__ void __launch_bounds__(1024) diag_kernel(DiagArgs a) {
  // which apply kernel did this step's work (the other one returned at once): diag[13] dense steps, diag[14] sparse steps
  if (threadIdx.x == 0) {
    const int longest = *reinterpret_cast<const int*>(a.zbuf + (size_t)a.B * 4);  // sparse_max(), armed by mcf_weights
    a.diag[longest > MCF_KMAX ? 13 : 14] += 1.0f;
  }
  if ((a.step ? *a.step : 0) != a.diag_step) return;
  __shared__ float sh[32][4];
  __shared__ float sz[32][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float nsum[4] = {0.f, 0.f, 0.f, 0.f};
  const float* ptr[4] = {a.vx, a.vy, a.gx, a.gy};
  const int dim[4] = {a.Dx, a.Dy, a.Dx, a.Dy};
  for (int b = warp; b < a.B; b += 32) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float s = 0.f;
      for (int d = lane; d < dim[k]; d += 32) {
        const float v = ptr[k][(size_t)b * dim[k] + d];
        s = fmaf(v, v, s);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      nsum[k] += sqrtf(s);
    }
  }
  float z = 0.f, wmin = INFINITY, wmax = -INFINITY;
  for (int b = tid; b < a.B; b += 1024) {
    z += a.zbuf[b * 4];
    wmin = fminf(wmin, a.zbuf[b * 4 + 1]);
    wmax = fmaxf(wmax, a.zbuf[b * 4 + 2]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    z += __shfl_xor_sync(0xffffffffu, z, o);
    wmin = fminf(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
    wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  }
  if (lane == 0) {
    for (int k = 0; k < 4; ++k) sh[warp][k] = nsum[k];
    sz[warp][0] = z; sz[warp][1] = wmin; sz[warp][2] = wmax;
  }
  __syncthreads();
  if (tid == 0) {
    float t4[4] = {0.f, 0.f, 0.f, 0.f};
    float zz = 0.f, mn = INFINITY, mxv = -INFINITY;
    for (int w = 0; w < 32; ++w) {
      for (int k = 0; k < 4; ++k) t4[k] += sh[w][k];
      zz += sz[w][0]; mn = fminf(mn, sz[w][1]); mxv = fmaxf(mxv, sz[w][2]);
    }
    const StepScalars sc = a.sc[a.diag_step];
    a.diag[0] = sc.t;
    a.diag[1] = sc.sigma;
    for (int k = 0; k < 4; ++k) a.diag[2 + k] = t4[k] / (float)a.B;
    a.diag[6] = mn;
    a.diag[7] = mxv;
    a.diag[8] = zz / (float)a.B;
  }
}

// diag[12] = output rows of the tensor-core convs whose fp16 store saturated during this run (0 for valid samples)
__global__ void sat_report_kernel(const unsigned int* a, const unsigned int* b, const unsigned int* c, float* out) {
  *out = (float)((a ? *a : 0u) + (b ? *b : 0u) + (c ? *c : 0u));
}
void launch_sat_report(const unsigned int* a, const unsigned int* b, const unsigned int* c, float* out, cudaStream_t s) {
  sat_report_kernel<<<1, 1, 0, s>>>(a, b, c, out);
  post_launch("sat_report");
}

// ---- pair coherence (evaluate_mnist_svhn.py:28-57, evaluate.py:56-90): argmax agreement of two classifiers ----
// logits [n][10] each; torch.argmax returns the FIRST maximum, so the scan keeps the earliest index (strict >).
__global__ void __launch_bounds__(256) argmax_match_kernel(const float* __restrict__ la, const float* __restrict__ lb, int n, float* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int hit = 0;
  if (i < n) {
    int ia = 0, ib = 0;
    float ma = la[(size_t)i * 10], mb = lb[(size_t)i * 10];
#pragma unroll
    for (int c = 1; c < 10; ++c) {
      const float va = la[(size_t)i * 10 + c], vb = lb[(size_t)i * 10 + c];
      if (va > ma) { ma = va; ia = c; }
      if (vb > mb) { mb = vb; ib = c; }
    }
    hit = ia == ib;
  }
  const unsigned m = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (float)__popc(m));  // integers below 2^24: exact in fp32
}
void launch_argmax_match(const float* la, const float* lb, int n, float* count, cudaStream_t s) {
  argmax_match_kernel<<<(n + 255) / 256, 256, 0, s>>>(la, lb, n, count);
  post_launch("argmax_match");
}

// inverse of the pair's transform on [n][1][H][H] images (evaluate.py:30-53; quarter turns and flips are pixel
// permutations: TF.rotate(img, 90) is a counter-clockwise quarter turn = torch.rot90(k=1))
__global__ void inverse_transform_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int H, int code) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * H * H) return;
  const int px = (int)(i % ((size_t)H * H)), r = px / H, c = px % H;
  const float* img = in + (i - px);
  float v;
  switch (code) {
    case 1: v = img[c * H + (H - 1 - r)]; break;            // rotate90  -> rot90(k=1):  out[r][c] = in[c][H-1-r]
    case 2: v = img[(H - 1 - r) * H + (H - 1 - c)]; break;  // rotate180 -> rot90(k=2)
    case 3: v = img[(H - 1 - c) * H + r]; break;            // rotate270 -> rot90(k=-1): out[r][c] = in[H-1-c][r]
    case 4: v = -img[px]; break;                            // invert
    case 5: v = img[r * H + (H - 1 - c)]; break;            // flip_h
    case 6: v = img[(H - 1 - r) * H + c]; break;            // flip_v
    default: v = img[px];
  }
  out[i] = v;
}
void launch_inverse_transform(const float* in, float* out, int n, int H, int code, cudaStream_t s) {
  const size_t tot = (size_t)n * H * H;
  inverse_transform_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(in, out, n, H, code);
  post_launch("inverse_transform");
}

void launch_diag(const DiagArgs& a, cudaStream_t s) {
  diag_kernel<<<1, 1024, 0, s>>>(a);
  post_launch("diag");
}

// min / max / mean of the MC ratios (the reference prints them: flow_utils.py:258)
__global__ void __launch_bounds__(256) ratio_stats_kernel(const float* __restrict__ r, int n, float* __restrict__ d3) {
  __shared__ float sh[8][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mn = INFINITY, mx = -INFINITY, sm = 0.f;
  for (int i = tid; i < n; i += 256) {
    const float v = r[i];
    mn = fminf(mn, v); mx = fmaxf(mx, v); sm += v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
  }
  if (lane == 0) { sh[warp][0] = mn; sh[warp][1] = mx; sh[warp][2] = sm; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, sh[w][0]); mx = fmaxf(mx, sh[w][1]); sm += sh[w][2]; }
    d3[0] = mn; d3[1] = mx; d3[2] = sm / (float)n;
  }
}
void launch_ratio_stats(const float* ratios, int n, float* diag3, cudaStream_t s) {
  ratio_stats_kernel<<<1, 256, 0, s>>>(ratios, n, diag3);
  post_launch("ratio_stats");
}

}  // namespace rgfm

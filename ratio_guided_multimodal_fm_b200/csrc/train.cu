// Training-side reuse: the loss and parameter gradients of one ratio-estimator optimisation step,
// i.e. what `loss.backward()` leaves in `param.grad` inside RatioTrainer.train_step (src/utils/trainer.py:26-56) for the
// losses of src/utils/losses.py:27-133 (DiscriminatorLoss, RuLSIFLoss).
//
// The step reuses the sampler's fp32 kernels for everything that already exists -- forward convs and the input-gradient
// (dgrad) convs on conv_simt, GroupNorm / BatchNorm-affine + SiLU + max-pool packs and their backward, GAP + fc and its
// backward -- and adds what only training needs: the loss heads with their score gradients, weight gradients of the
// convs (a deterministic split-K outer-product kernel), of the linears and of the normalisation affines, and the
// score-net backward with caller-supplied dropout masks.
//
// Parameters are read IN PLACE from the caller's live parameter tensors (device pointers in torch layout) and gradients
// are written straight into the caller's gradient tensors: no host round trip and no plan rebuild after an optimiser
// step.  The conv kernels' [tap][Cin][CoutP] weight layouts are refreshed by a repack kernel at the start of every step
// (3.3 M floats at most).  Everything is fp32; summation orders are fixed, so a step is bit-reproducible.
//
// Not covered (stated in DESIGN.md): BatchNorm in training mode (batch statistics) -- RatioEstimatorMNISTSVHN is
// trained here with frozen running statistics; the U-Nets' flow-matching step (needs a U-Net backward chain).
#include "net.h"
#include "simt_device.cuh"
#include <cmath>
#include <algorithm>
#include <cstring>
#include <tuple>

namespace rgfm {

static inline void post_launch_t(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(std::string(what) + " launch failed: " + cudaGetErrorString(e));
  ++g_launch_counter;
}

// ------------------------------------------------------------------------------------------------
// weight refresh: torch [Cout][Cin][3][3] -> conv_simt layouts
//   forward  wf[tap][ci][co]            (ci < CinP, co < CoutP, zero padded)
//   dgrad    wb[tap][co][ci] = W[co][ci][8 - tap]   (rows = Cout, columns padded to CinB; net.cu encoder())
// ------------------------------------------------------------------------------------------------
__global__ void repack_conv_w_kernel(const float* __restrict__ w, int cout, int cin, int CinP, int CoutP, int CinB,
                                     float* __restrict__ wf, float* __restrict__ wb) {
  const int nf = 9 * CinP * CoutP, nb = wb ? 9 * cout * CinB : 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf + nb; i += gridDim.x * blockDim.x) {
    if (i < nf) {
      const int co = i % CoutP, ci = (i / CoutP) % CinP, t = i / (CoutP * CinP);
      wf[i] = (co < cout && ci < cin) ? w[((size_t)co * cin + ci) * 9 + t] : 0.f;
    } else {
      const int j = i - nf;
      const int ci = j % CinB, co = (j / CinB) % cout, t = j / (CinB * cout);
      wb[j] = (ci < cin) ? w[((size_t)co * cin + ci) * 9 + (8 - t)] : 0.f;
    }
  }
}
__global__ void pad_vec_kernel(const float* __restrict__ v, int n, int nP, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nP) out[i] = i < n ? v[i] : 0.f;
}
// BatchNorm with frozen running statistics as a per-channel affine (ratio_flexible.py:217-226 in eval mode)
__global__ void bn_affine_kernel(const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ m,
                                 const float* __restrict__ v, int C, float* __restrict__ scale, float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    const float inv = 1.0f / sqrtf(v[i] + 1e-5f);
    const float sc = g[i] * inv;
    scale[i] = sc;
    shift[i] = b[i] - m[i] * sc;
  }
}

// ------------------------------------------------------------------------------------------------
// Loss heads (losses.py:27-133) over the scores of one batch; is_real[i] in {0, 1}.  One CTA, fixed summation order.
//   disc:   L = mean_real softplus(-s) + mean_fake softplus(s)                    (BCE with logits, targets 1 / 0)
//   rulsif: w = softplus(s);  L = 0.5 mean_all w^2 - mean_real w + lambda (mean_all w - 1)^2   (alpha is unused by the
//           reference's forward: "treat concatenated samples as mixture", losses.py:108-121)
// out[0] = loss; disc: out[1] = acc_real, out[2] = acc_fake; rulsif: out[1] = mean_w_real, out[2] = mean_w_fake,
// out[3] = constraint term.  dscore[i] = dL / ds_i.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // F.softplus (threshold 20)
__device__ __forceinline__ float sigmoid_t(float x) { return 1.0f / (1.0f + expf(-x)); }
// softplus without the threshold shortcut, the way binary_cross_entropy_with_logits evaluates it
__device__ __forceinline__ float softplus_stable(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

__global__ void __launch_bounds__(256) ratio_loss_kernel(const float* __restrict__ s, const float* __restrict__ is_real, int n,
                                                         int loss_type, float lambda, float* __restrict__ out,
                                                         float* __restrict__ dscore) {
  __shared__ float red[8];
  const int tid = threadIdx.x;
  float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // n_real | disc: sum_r sp(-s), sum_f sp(s), #r(s>0), #f(s<0) | rulsif: sum_r w, sum_f w, sum w^2
  for (int i = tid; i < n; i += 256) {
    const bool r = is_real[i] != 0.f;
    const float v = s[i];
    a[0] += r ? 1.f : 0.f;
    if (loss_type == RGFM_LOSS_DISC) {
      if (r) { a[1] += softplus_stable(-v); a[3] += v > 0.f ? 1.f : 0.f; }
      else   { a[2] += softplus_stable(v);  a[4] += v < 0.f ? 1.f : 0.f; }
    } else {
      const float w = softplus_t(v);
      if (r) a[1] += w; else a[2] += w;
      a[3] = fmaf(w, w, a[3]);
    }
  }
  float t[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) t[k] = block_sum256(a[k], red);
  const float nr = t[0], nf = (float)n - t[0], nn = (float)n;
  if (loss_type == RGFM_LOSS_DISC) {
    if (tid == 0) {
      out[0] = t[1] / nr + t[2] / nf;
      out[1] = t[3] / nr;
      out[2] = t[4] / nf;
      out[3] = 0.f;
    }
    for (int i = tid; i < n; i += 256) {
      const float v = s[i];
      dscore[i] = is_real[i] != 0.f ? -sigmoid_t(-v) / nr : sigmoid_t(v) / nf;
    }
  } else {
    const float m = (t[1] + t[2]) / nn;
    const float cons = lambda * (m - 1.f) * (m - 1.f);
    if (tid == 0) {
      out[0] = 0.5f * t[3] / nn - t[1] / nr + cons;
      out[1] = t[1] / nr;
      out[2] = t[2] / nf;
      out[3] = cons;
    }
    const float dm = 2.f * lambda * (m - 1.f) / nn;
    for (int i = tid; i < n; i += 256) {
      const float v = s[i];
      const float w = softplus_t(v);
      const float dw = w / nn - (is_real[i] != 0.f ? 1.f / nr : 0.f) + dm;
      dscore[i] = dw * (v > 20.f ? 1.f : sigmoid_t(v));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Score net in training mode: Linear -> LayerNorm -> SiLU -> Dropout per hidden layer, then Linear -> 1
// (ratio_estimator.py:125-135, ratio_flexible.py:332-345), as BATCHED layer kernels: a tiled GEMM per Linear (forward
// Y = X W^T + b, backward dX = dY W) and one warp per sample row for LayerNorm + SiLU + dropout and their backward.
// Dropout is a caller-supplied mask per hidden layer (values 0 or 1 / (1 - p); null = no dropout).  The backward also
// leaves what the parameter-gradient reductions need: dh (gradient at each Linear output) and, for the LayerNorm affine,
// dz * xhat and dz per element.  Fixed summation orders: deterministic.
// ------------------------------------------------------------------------------------------------
// Y[n][j] = b[j] + sum_i X[n][i] W[j][i]      (32 x 32 outputs per CTA)
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                         const float* __restrict__ b, int I, int J, int N, float* __restrict__ Y, int ldy) {
  __shared__ float Xs[32][33], Ws[32][33];
  const int n0 = blockIdx.y * 32, j0 = blockIdx.x * 32, tid = threadIdx.x;
  const int tj = tid & 31, tn = tid >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < I; k0 += 32) {
    for (int e = tid; e < 1024; e += 256) {
      const int r = e >> 5, c = e & 31;
      Xs[r][c] = (n0 + r < N && k0 + c < I) ? X[(size_t)(n0 + r) * ldx + k0 + c] : 0.f;
      Ws[r][c] = (j0 + r < J && k0 + c < I) ? W[(size_t)(j0 + r) * I + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float w = Ws[tj][k];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(Xs[tn + 8 * q][k], w, acc[q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int n = n0 + tn + 8 * q, j = j0 + tj;
    if (n < N && j < J) Y[(size_t)n * ldy + j] = acc[q] + b[j];
  }
}
// dX[n][i] = sum_j dY[n][j] W[j][i]
__global__ void __launch_bounds__(256) linear_bwd_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ W, int I, int J,
                                                         int N, float* __restrict__ dX, int ldx) {
  __shared__ float Ys[32][33], Ws[32][33];
  const int n0 = blockIdx.y * 32, i0 = blockIdx.x * 32, tid = threadIdx.x;
  const int ti = tid & 31, tn = tid >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < J; k0 += 32) {
    for (int e = tid; e < 1024; e += 256) {
      const int r = e >> 5, c = e & 31;
      Ys[r][c] = (n0 + r < N && k0 + c < J) ? dY[(size_t)(n0 + r) * ldy + k0 + c] : 0.f;
      Ws[r][c] = (k0 + r < J && i0 + c < I) ? W[(size_t)(k0 + r) * I + i0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float w = Ws[k][ti];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(Ys[tn + 8 * q][k], w, acc[q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int n = n0 + tn + 8 * q, i = i0 + ti;
    if (n < N && i < I) dX[(size_t)n * ldx + i] = acc[q];
  }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// one warp per row: act = mask * silu(LN(pre));  stat[n] = {mean, rstd}
__global__ void __launch_bounds__(256) ln_silu_fwd_kernel(const float* __restrict__ pre, int C, int N, const float* __restrict__ g,
                                                          const float* __restrict__ b, const float* __restrict__ mask,
                                                          float* __restrict__ act, float2* __restrict__ stat) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* p = pre + (size_t)n * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += p[c];
  const float mean = warp_sum(s) / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = p[c] - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)C + 1e-5f);
  if (lane == 0) stat[n] = make_float2(mean, rstd);
  for (int c = lane; c < C; c += 32) {
    float a = silu_f((p[c] - mean) * rstd * g[c] + b[c]);
    if (mask) a *= mask[(size_t)n * C + c];
    act[(size_t)n * C + c] = a;
  }
}
// one warp per row: gradient at the activation -> dh (gradient at the Linear output), plus dz * xhat and dz
__global__ void __launch_bounds__(256) ln_silu_bwd_kernel(const float* __restrict__ dact, const float* __restrict__ pre, int C, int N,
                                                          const float* __restrict__ g, const float* __restrict__ b,
                                                          const float* __restrict__ mask, const float2* __restrict__ stat,
                                                          float* __restrict__ dh, float* __restrict__ dzx, float* __restrict__ dzo) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  const float2 st = stat[n];
  const size_t r = (size_t)n * C;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float xh = (pre[r + c] - st.x) * st.y;
    const float z = xh * g[c] + b[c];
    float dz = dact[r + c];
    if (mask) dz *= mask[r + c];
    const float sg = 1.0f / (1.0f + expf(-z));
    dz *= sg * (1.0f + z * (1.0f - sg));
    dzx[r + c] = dz * xh;
    dzo[r + c] = dz;
    const float dg = dz * g[c];
    s1 += dg;
    s2 = fmaf(dg, xh, s2);
  }
  const float m1 = warp_sum(s1) / (float)C, m2 = warp_sum(s2) / (float)C;
  for (int c = lane; c < C; c += 32) {
    const float xh = (pre[r + c] - st.x) * st.y;
    dh[r + c] = st.y * (dzo[r + c] * g[c] - m1 - xh * m2);
  }
}
// scores[n] = pre_last[n][0]
__global__ void copy_col0_kernel(const float* __restrict__ src, int ld, int N, float* __restrict__ dst) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) dst[n] = src[(size_t)n * ld];
}

// ------------------------------------------------------------------------------------------------
// dW[j][i] = sum_n A[n][j] * B[n][i]   (j < J, i < I; row strides lda / ldb): weight gradient of a Linear.
// 32 x 32 outputs per CTA, samples walked in order (deterministic).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) linear_wgrad_kernel(const float* __restrict__ A, int lda, int J, const float* __restrict__ B,
                                                           int ldb, int I, int N, float* __restrict__ dW) {
  __shared__ float As[32][33], Bs[32][33];
  const int j0 = blockIdx.y * 32, i0 = blockIdx.x * 32, tid = threadIdx.x;
  const int ti = tid & 31, tj = tid >> 5;  // outputs (j0 + tj + 8 q, i0 + ti), q = 0..3
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n0 = 0; n0 < N; n0 += 32) {
    for (int e = tid; e < 1024; e += 256) {
      const int r = e >> 5, c = e & 31;
      const int n = n0 + r;
      As[r][c] = (n < N && j0 + c < J) ? A[(size_t)n * lda + j0 + c] : 0.f;
      Bs[r][c] = (n < N && i0 + c < I) ? B[(size_t)n * ldb + i0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float b = Bs[r][ti];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(As[r][tj + 8 * q], b, acc[q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = j0 + tj + 8 * q, i = i0 + ti;
    if (j < J && i < I) dW[(size_t)j * I + i] = acc[q];
  }
}

// out[split][c] = sum over the split's rows of M[r][c]  (column sums, two deterministic stages: call again on `out`)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ M, int ld, int C, long long rows, int rows_per_split,
                                                     float* __restrict__ out) {
  __shared__ float red[8][33];
  const int tc = threadIdx.x & 31, tr = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tc;
  const long long r0 = (long long)blockIdx.y * rows_per_split;
  const long long r1 = r0 + rows_per_split < rows ? r0 + rows_per_split : rows;
  float s = 0.f;
  if (c < C)
    for (long long r = r0 + tr; r < r1; r += 8) s += M[(size_t)r * ld + c];
  red[tr][tc] = s;
  __syncthreads();
  if (tr == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][tc];
    out[(size_t)blockIdx.y * C + c] = t;
  }
}

// mean over the pixels of the last packed activation: gap[n][c]  (the fc layer's input, ratio_estimator.py:88-90)
__global__ void __launch_bounds__(256) gap_kernel(const float* __restrict__ in, int HW, int C, float* __restrict__ out) {
  const int n = blockIdx.x;
  const float* src = in + (size_t)n * HW * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += src[(size_t)p * C + c];
    out[(size_t)n * C + c] = s / (float)HW;
  }
}

// ------------------------------------------------------------------------------------------------
// Normalisation affine gradients of one layer, per image and pixel slice (blockIdx.y): part[n][slice][c] = {sum_p dz * u, sum_p dz} where dz is the gradient
// at the norm's output (incoming gradient routed through the max-pool to the first maximum and through SiLU', exactly
// as pack_bwd_kernel does) and u = xhat (GroupNorm) or the raw activation (per-channel affine).  C must divide 256, so
// that a thread owns one channel; the 256 / C threads of a channel are combined in a fixed order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) norm_param_grad_kernel(PackArgs a, const float* __restrict__ dcur, float* __restrict__ part) {
  __shared__ float s_scale[PACK_MAXC];
  __shared__ float s_shift[PACK_MAXC];
  __shared__ float s_red[2][PACK_MAXC];
  __shared__ double s_grp[32][2];
  const int n = blockIdx.x, tid = threadIdx.x;
  pack_scale_shift(a, n, s_scale, s_shift, s_red, s_grp);
  const int C = a.C, H = a.Hs, Wd = a.Ws;
  const bool pool = a.spatial == SP_POOL2;
  const float* raw = a.src[0].p + (size_t)n * H * Wd * C;
  const float* dc = dcur + (size_t)n * a.Ho * a.Wo * C;
  const int c = tid % C, cpg = C / a.groups;
  const float sc = s_scale[c], sh = s_shift[c];
  const float mean = a.norm == NORM_GN ? (float)s_grp[c / cpg][0] : 0.f, rstd = a.norm == NORM_GN ? (float)s_grp[c / cpg][1] : 1.f;
  float s0 = 0.f, s1 = 0.f;
  for (int pix = blockIdx.y * (256 / C) + tid / C; pix < H * Wd; pix += gridDim.y * (256 / C)) {
    const int h = pix / Wd, w = pix - h * Wd;
    const float r = raw[(size_t)pix * C + c];
    const float z = fmaf(r, sc, sh);
    float g;
    if (!pool) {
      g = dc[(size_t)pix * C + c];
    } else {
      const int ho = h >> 1, wo = w >> 1;
      g = 0.f;
      if (ho < a.Ho && wo < a.Wo) {
        float best = -INFINITY;
        int bi = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int hh = 2 * ho + (k >> 1), ww = 2 * wo + (k & 1);
          const float v = silu_f(fmaf(raw[((size_t)hh * Wd + ww) * C + c], sc, sh));
          if (v > best) {
            best = v;
            bi = k;
          }
        }
        if (bi == ((h & 1) * 2 + (w & 1))) g = dc[((size_t)ho * a.Wo + wo) * C + c];
      }
    }
    const float sg = 1.0f / (1.0f + expf(-z));
    const float dz = g * sg * (1.0f + z * (1.0f - sg));
    const float u = a.norm == NORM_GN ? (r - mean) * rstd : r;
    s0 = fmaf(dz, u, s0);
    s1 += dz;
  }
  __syncthreads();
  s_red[0][tid] = s0;
  s_red[1][tid] = s1;
  __syncthreads();
  if (tid < C) {
    float t0 = 0.f, t1 = 0.f;
    for (int j = tid; j < 256; j += C) {
      t0 += s_red[0][j];
      t1 += s_red[1][j];
    }
    float* q = part + (((size_t)n * gridDim.y + blockIdx.y) * C + tid) * 2;
    q[0] = t0;
    q[1] = t1;
  }
}
// sums[c][2] (over the images) -> the affine's parameter gradients.  GroupNorm: d gamma = sums[c][0], d beta = sums[c][1].
// Frozen BatchNorm y = (x - rm) * inv * w + b:  d w = inv * (sum dz x - rm sum dz), d b = sum dz.
__global__ void norm_grad_finish_kernel(const float* __restrict__ sums, int C, const float* __restrict__ rm, const float* __restrict__ rv,
                                        float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float a = sums[2 * c], b = sums[2 * c + 1];
  if (rm) {
    const float inv = 1.0f / sqrtf(rv[c] + 1e-5f);
    dgamma[c] = inv * (a - rm[c] * b);
  } else {
    dgamma[c] = a;
  }
  dbeta[c] = b;
}

// ------------------------------------------------------------------------------------------------
// Conv weight gradient: dW[co][ci][ky][kx] = sum_{n,h,w} dY[n][h][w][co] * X[n][h + ky - 1][w + kx - 1][ci]
// X: the conv's fp32 NHWC operand (channel stride CinP), dY: dense fp32 NHWC gradient of the raw output.
// A CTA owns a (16 CI x 64) tile of (ci, co) for all nine taps and a contiguous range of (image, row block) units; per unit
// (RB rows, about 32 pixels) it stages the RB + 2 operand rows (zero halo) and the RB gradient rows in shared memory and
// accumulates 9 x CI x 4 outer-product terms per thread and pixel.  Partials part[split][tap][ci][co] are summed in split order by conv_wgrad_reduce_kernel, which also
// transposes to the torch layout: deterministic, no atomics.
// ------------------------------------------------------------------------------------------------
template <int CI>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ X, int CinP, const float* __restrict__ dY, int Cout,
                                                         int H, int W, int RB, int nimg, int units_per_split, float* __restrict__ part) {
  constexpr int TCI = 16 * CI, TCO = 64;
  extern __shared__ __align__(16) float smw[];
  float* Xs = smw;                                // [RB + 2][W + 2][TCI]
  float* Ys = smw + (RB + 2) * (W + 2) * TCI;     // [RB][W][TCO]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int co_tiles = (Cout + TCO - 1) / TCO;
  const int cot = blockIdx.x % co_tiles, cit = blockIdx.x / co_tiles;
  const int ci0 = cit * TCI, co0 = cot * TCO;
  float acc[9][CI][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < CI; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][c][j] = 0.f;
  // work unit = (image, block of RB consecutive rows)
  const int nblk = (H + RB - 1) / RB;
  const int total = nimg * nblk;
  const int u0 = blockIdx.y * units_per_split, u1 = min(u0 + units_per_split, total);
  for (int u = u0; u < u1; ++u) {
    const int n = u / nblk, h0 = (u - n * nblk) * RB;
    __syncthreads();
    for (int i = tid; i < (RB + 2) * (W + 2) * (TCI / 4); i += 256) {
      const int c4 = i % (TCI / 4), q = i / (TCI / 4), wq = q % (W + 2), ky = q / (W + 2);
      const int hh = h0 + ky - 1, ww = wq - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hh >= 0 && hh < H && ww >= 0 && ww < W)
        v = __ldg(reinterpret_cast<const float4*>(X + (((size_t)n * H + hh) * W + ww) * CinP + ci0 + c4 * 4));
      *reinterpret_cast<float4*>(Xs + (size_t)(ky * (W + 2) + wq) * TCI + c4 * 4) = v;
    }
    for (int i = tid; i < RB * W * (TCO / 4); i += 256) {
      const int c4 = i % (TCO / 4), pq = i / (TCO / 4), w = pq % W, rr = pq / W, co = co0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (co < Cout && h0 + rr < H) v = __ldg(reinterpret_cast<const float4*>(dY + (((size_t)n * H + h0 + rr) * W + w) * Cout + co));
      *reinterpret_cast<float4*>(Ys + (size_t)pq * TCO + c4 * 4) = v;
    }
    __syncthreads();
    for (int rr = 0; rr < RB; ++rr)
      for (int w = 0; w < W; ++w) {
        const float4 y = *reinterpret_cast<const float4*>(Ys + (size_t)(rr * W + w) * TCO + tx * 4);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float* xp = Xs + (size_t)((rr + t / 3) * (W + 2) + w + (t % 3)) * TCI + ty * CI;
          float x[CI];
          if (CI == 4) {
            const float4 v = *reinterpret_cast<const float4*>(xp);
            x[0] = v.x; x[1 % CI] = v.y; x[2 % CI] = v.z; x[3 % CI] = v.w;
          } else if (CI == 2) {
            const float2 v = *reinterpret_cast<const float2*>(xp);
            x[0] = v.x; x[1 % CI] = v.y;
          } else {
            x[0] = xp[0];
          }
#pragma unroll
          for (int c = 0; c < CI; ++c) {
            acc[t][c][0] = fmaf(x[c], y.x, acc[t][c][0]);
            acc[t][c][1] = fmaf(x[c], y.y, acc[t][c][1]);
            acc[t][c][2] = fmaf(x[c], y.z, acc[t][c][2]);
            acc[t][c][3] = fmaf(x[c], y.w, acc[t][c][3]);
          }
        }
      }
  }
  const int co = co0 + tx * 4;
  if (co < Cout) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int c = 0; c < CI; ++c) {
        const int ci = ci0 + ty * CI + c;
        *reinterpret_cast<float4*>(part + (((size_t)blockIdx.y * 9 + t) * CinP + ci) * Cout + co) =
            make_float4(acc[t][c][0], acc[t][c][1], acc[t][c][2], acc[t][c][3]);
      }
  }
}
__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ part, int nsplit, int CinP, int Cout, int cin,
                                         float* __restrict__ dW) {
  const int total = Cout * cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    // thread order follows the partials' layout (co fastest) so the reads coalesce; the store is the transposition
    const int co = i % Cout, ci = (i / Cout) % cin, t = i / (Cout * cin);
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += part[(((size_t)k * 9 + t) * CinP + ci) * Cout + co];
    dW[((size_t)co * cin + ci) * 9 + t] = s;
  }
}

void launch_ratio_loss(const float* scores, const float* is_real, int n, int loss_type, float lambda, float* out8, float* dscore,
                       cudaStream_t s) {
  ratio_loss_kernel<<<1, 256, 0, s>>>(scores, is_real, n, loss_type, lambda, out8, dscore);
  post_launch_t("ratio_loss");
}

// ------------------------------------------------------------------------------------------------
// Flow-matching step, data side (train_flow_matching_epoch, flow_utils.py:131-142): the rectified-flow interpolation of
// CFMSchedule.add_noise (flow_utils.py:40-66) in one pass, and F.mse_loss(v, u) with its gradient dL/dv = 2 (v - u) / n.
// HBM-bound streaming kernels: 128-bit loads and stores, each tensor touched once.  x_t is rounded like the reference
// (two products, one sum, no FMA); u = x_1 - x_0 is exact.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cfm_interp_kernel(const float4* __restrict__ x1, const float4* __restrict__ x0,
                                                         const float* __restrict__ t, size_t d4, size_t n4,
                                                         float4* __restrict__ xt, float4* __restrict__ u) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float tb = __ldg(t + i / d4), om = __fsub_rn(1.0f, tb);
    const float4 a = __ldg(x1 + i), b = __ldg(x0 + i);
    float4 r, v;
    r.x = __fadd_rn(__fmul_rn(om, b.x), __fmul_rn(tb, a.x)); v.x = __fsub_rn(a.x, b.x);
    r.y = __fadd_rn(__fmul_rn(om, b.y), __fmul_rn(tb, a.y)); v.y = __fsub_rn(a.y, b.y);
    r.z = __fadd_rn(__fmul_rn(om, b.z), __fmul_rn(tb, a.z)); v.z = __fsub_rn(a.z, b.z);
    r.w = __fadd_rn(__fmul_rn(om, b.w), __fmul_rn(tb, a.w)); v.w = __fsub_rn(a.w, b.w);
    xt[i] = r;
    u[i] = v;
  }
}
// part[block] = sum over the block's elements of (v - u)^2 ; dv = 2 (v - u) / n
__global__ void __launch_bounds__(256) fm_mse_kernel(const float4* __restrict__ v, const float4* __restrict__ u, size_t n4, float scale,
                                                     float4* __restrict__ dv, float* __restrict__ part) {
  __shared__ float red[8];
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(v + i), b = __ldg(u + i);
    const float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    s = fmaf(d.x, d.x, s); s = fmaf(d.y, d.y, s); s = fmaf(d.z, d.z, s); s = fmaf(d.w, d.w, s);
    if (dv) dv[i] = make_float4(scale * d.x, scale * d.y, scale * d.z, scale * d.w);
  }
  const float t = block_sum256(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}
__global__ void __launch_bounds__(256) fm_mse_finish_kernel(const float* __restrict__ part, int nparts, float inv_n, float* __restrict__ loss) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 256) s += part[i];
  const float t = block_sum256(s, red);
  if (threadIdx.x == 0) *loss = t * inv_n;
}
void launch_cfm_interp(const float* x1, const float* x0, const float* t, int batch, size_t dim, float* xt, float* u, cudaStream_t s) {
  RGFM_CHECK(dim % 4 == 0, "cfm_add_noise: sample size must be a multiple of 4");
  const size_t n4 = (size_t)batch * dim / 4;
  const int blocks = (int)std::min<size_t>((n4 + 255) / 256, 148 * 8);
  cfm_interp_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(x1), reinterpret_cast<const float4*>(x0), t, dim / 4, n4,
                                            reinterpret_cast<float4*>(xt), reinterpret_cast<float4*>(u));
  post_launch_t("cfm_interp");
}
// scratch: 148 * 8 floats
void launch_fm_mse(const float* v, const float* u, size_t n, float* loss, float* dv, float* scratch, cudaStream_t s) {
  RGFM_CHECK(n % 4 == 0 && n > 0, "fm_mse: element count must be a positive multiple of 4");
  const size_t n4 = n / 4;
  const int blocks = (int)std::min<size_t>((n4 + 255) / 256, 148 * 8);
  fm_mse_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(v), reinterpret_cast<const float4*>(u), n4, 2.0f / (float)n,
                                        reinterpret_cast<float4*>(dv), scratch);
  post_launch_t("fm_mse");
  fm_mse_finish_kernel<<<1, 256, 0, s>>>(scratch, blocks, 1.0f / (float)n, loss);
  post_launch_t("fm_mse_finish");
}

// ================================================================================================
// host side
// ================================================================================================
struct TrainLayer { std::string conv, norm; int cin, cout; bool pool; };

struct TrainEncoder {
  std::string pre;
  std::vector<TrainLayer> L;
  bool gn = false, from_y = false;
  int in_ch = 1, img = 28, feat_off = 0, fc_in = 128;
  // per layer
  std::vector<int> sp, cinP;
  std::vector<float*> opnd, raw, stats, draw, dopnd, wf, wb, biasP, scale, shift;
  std::vector<PackArgs> pa;
  float *cur_last = nullptr, *dcur_last = nullptr, *gap = nullptr, *zero_bias = nullptr;
  int HWl = 0;
};

class RatioTrainer {
 public:
  RatioTrainer(int kind, const rgfm_weights* params, const rgfm_weights* grads, int max_batch, int device)
      : kind_(kind), B_(max_batch), device_(device), P_(params), G_(grads) {
    RGFM_CHECK(kind == RGFM_NET_RATIO28 || kind == RGFM_NET_RATIO_MNIST_SVHN, "ratio trainer: not a ratio estimator kind");
    RGFM_CHECK(max_batch >= 1, "ratio trainer: max_batch");
    RGFM_CUDA(cudaSetDevice(device));
    if (kind == RGFM_NET_RATIO28) {
      std::vector<TrainLayer> L{{"conv1", "gn1", 1, 32, true}, {"conv2", "gn2", 32, 64, true}, {"conv3", "gn3", 64, 128, true}, {"conv4", "gn4", 128, 128, false}};
      add_encoder("encoder_x.", L, true, 1, 28, false, 0, 128);
      add_encoder("encoder_y.", L, true, 1, 28, true, 256, 128);
      add_mlp({{0, 1}, {4, 5}}, 8, {512, 256});
    } else {
      std::vector<TrainLayer> M{{"conv1", "bn1", 1, 32, true}, {"conv2", "bn2", 32, 64, true}, {"conv3", "bn3", 64, 128, true}, {"conv4", "bn4", 128, 128, false}};
      std::vector<TrainLayer> S{{"conv1a", "bn1a", 3, 64, false},   {"conv1b", "bn1b", 64, 64, true},   {"conv2a", "bn2a", 64, 128, false},
                                {"conv2b", "bn2b", 128, 128, true}, {"conv3a", "bn3a", 128, 256, false}, {"conv3b", "bn3b", 256, 256, true},
                                {"conv4a", "bn4a", 256, 256, false}, {"conv4b", "bn4b", 256, 256, true}};
      add_encoder("encoder_mnist.", M, false, 1, 32, false, 0, 128);
      add_encoder("encoder_svhn.", S, false, 3, 32, true, 256, 256);
      add_mlp({{0, 1}, {4, 5}, {8, 9}}, 11, {512, 512, 256});
    }
    RGFM_CUDA(cudaStreamCreateWithFlags(&cap_, cudaStreamNonBlocking));
    sx_ = dalloc((size_t)B_ * dim_x());
    sy_ = dalloc((size_t)B_ * dim_y());
    sr_ = dalloc(B_);
    for (int i = 0; i < nmlp_ - 1; ++i) smask_[i] = dalloc((size_t)B_ * mlp_[i].out);
    feat_ = dalloc((size_t)B_ * 512);
    dfeat_ = dalloc((size_t)B_ * 512);
    scores_ = dalloc(B_);
    dscore_ = dalloc(B_);
    out_ = dalloc(8);
    // scratch for the two-stage column sums and the conv weight-gradient partials
    colsum_scratch_ = dalloc((size_t)kColSplits * 512);
    size_t wmax = 0;
    for (auto& e : enc_)
      for (size_t i = 0; i < e.L.size(); ++i) wmax = std::max(wmax, (size_t)wgrad_splits(e, (int)i) * 9 * e.cinP[i] * e.L[i].cout);
    wpart_ = dalloc(wmax);
    size_t npmax = 0;
    for (auto& e : enc_)
      for (auto& l : e.L) npmax = std::max(npmax, (size_t)B_ * kMaxSlices * l.cout * 2);
    npart_ = dalloc(npmax);
    nsum_ = dalloc(1024);
  }
  ~RatioTrainer() {
    cudaSetDevice(device_);
    for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second);
    if (cap_) cudaStreamDestroy(cap_);
    for (void* p : owned_) cudaFree(p);
  }
  int device() const { return device_; }
  int max_batch() const { return B_; }
  size_t dim_x() const { return (size_t)enc_[0].in_ch * enc_[0].img * enc_[0].img; }
  size_t dim_y() const { return (size_t)enc_[1].in_ch * enc_[1].img * enc_[1].img; }
  int hidden_layers() const { return nmlp_ - 1; }
  int hidden_width(int i) const { return mlp_[i].out; }

  // One step: forward in training mode, loss, backward into the caller's gradient tensors.  out8 (device, 8 floats) gets
  // {loss, metric1, metric2, metric3} (ratio_loss_kernel).  masks: one device pointer per hidden layer or null.
  void step(const float* x, const float* y, const float* is_real, int nb, int loss_type, float lambda, const float* const* masks,
            int n_masks, float* out8, cudaStream_t s) {
    RGFM_CHECK(nb >= 1 && nb <= B_, "ratio trainer: batch exceeds max_batch");
    RGFM_CHECK(n_masks >= 0 && n_masks <= nmlp_ - 1, "ratio trainer: at most one dropout mask per hidden layer of the score net");
    // The ~250 launches of a step are captured ONCE per (batch, loss, masks) into a CUDA graph over plan-owned staging
    // buffers: a training loop hands in fresh tensors every step, so the inputs are copied in (0.8 .. 2 MB, device to
    // device) and the graph's pointers never change.  Parameters are read through their live pointers either way.
    RGFM_CUDA(cudaMemcpyAsync(sx_, x, (size_t)nb * dim_x() * sizeof(float), cudaMemcpyDeviceToDevice, s));
    RGFM_CUDA(cudaMemcpyAsync(sy_, y, (size_t)nb * dim_y() * sizeof(float), cudaMemcpyDeviceToDevice, s));
    RGFM_CUDA(cudaMemcpyAsync(sr_, is_real, (size_t)nb * sizeof(float), cudaMemcpyDeviceToDevice, s));
    for (int i = 0; i < n_masks; ++i)
      RGFM_CUDA(cudaMemcpyAsync(smask_[i], masks[i], (size_t)nb * mlp_[i].out * sizeof(float), cudaMemcpyDeviceToDevice, s));
    static const bool graphs_on = !(getenv("RGFM_TRAIN_GRAPH") && atoi(getenv("RGFM_TRAIN_GRAPH")) == 0);  // A/B switch (same results)
    if (!graphs_on) {
      body(nb, loss_type, lambda, n_masks, s);
    } else {
      uint32_t lbits;
      memcpy(&lbits, &lambda, 4);
      const std::tuple<int, int, int, uint32_t> key(nb, loss_type, n_masks, lbits);
      auto it = graphs_.find(key);
      if (it == graphs_.end()) {
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ex = nullptr;
        RGFM_CUDA(cudaStreamBeginCapture(cap_, cudaStreamCaptureModeThreadLocal));
        try {
          body(nb, loss_type, lambda, n_masks, cap_);
        } catch (...) {
          cudaStreamEndCapture(cap_, &g);
          if (g) cudaGraphDestroy(g);
          throw;
        }
        RGFM_CUDA(cudaStreamEndCapture(cap_, &g));
        cudaError_t e = cudaGraphInstantiate(&ex, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) throw Error(std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(e));
        it = graphs_.emplace(key, ex).first;
      }
      RGFM_CUDA(cudaGraphLaunch(it->second, s));
    }
    RGFM_CUDA(cudaMemcpyAsync(out8, out_, 8 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }

 private:
  void body(int nb, int loss_type, float lambda, int n_masks, cudaStream_t s) {
    const float *x = sx_, *y = sy_, *is_real = sr_;
    refresh(s);
    for (auto& e : enc_) forward_encoder(e, e.from_y ? y : x, nb, s);
    float* const* mk = smask_;
    auto mask_of = [&](int i) -> const float* { return i < n_masks ? mk[i] : nullptr; };
    // forward: act[0] = feat; act_[i + 1] = dropout(silu(LN(pre_[i]))); the last Linear's output is the score
    for (int i = 0; i < nmlp_; ++i) {
      const MlpLayer& L = mlp_[i];
      const float* ain = i == 0 ? feat_ : act_[i];
      linear_fwd_kernel<<<dim3((L.out + 31) / 32, (nb + 31) / 32), 256, 0, s>>>(ain, i == 0 ? 512 : L.in, L.w, L.b, L.in, L.out, nb, pre_[i], L.out);
      post_launch_t("linear_fwd");
      if (i < nmlp_ - 1) {
        ln_silu_fwd_kernel<<<(nb + 7) / 8, 256, 0, s>>>(pre_[i], L.out, nb, L.ln_g, L.ln_b, mask_of(i), act_[i + 1], reinterpret_cast<float2*>(stat_[i]));
        post_launch_t("ln_silu_fwd");
      }
    }
    copy_col0_kernel<<<(nb + 255) / 256, 256, 0, s>>>(pre_[nmlp_ - 1], 1, nb, scores_);
    post_launch_t("copy_col0");
    ratio_loss_kernel<<<1, 256, 0, s>>>(scores_, is_real, nb, loss_type, lambda, out_, dscore_);
    post_launch_t("ratio_loss");
    // backward: dh_[i] = gradient at the output of Linear i; dact_[i] = gradient at its input
    for (int i = nmlp_ - 1; i >= 0; --i) {
      const MlpLayer& L = mlp_[i];
      const float* dhp = dscore_;  // the last Linear's output IS the score
      if (i < nmlp_ - 1) {
        ln_silu_bwd_kernel<<<(nb + 7) / 8, 256, 0, s>>>(dact_[i + 1], pre_[i], L.out, nb, L.ln_g, L.ln_b, mask_of(i),
                                                        reinterpret_cast<const float2*>(stat_[i]), dh_[i], dzx_[i], dz_[i]);
        post_launch_t("ln_silu_bwd");
        dhp = dh_[i];
      }
      dh_ptr_[i] = dhp;
      linear_bwd_kernel<<<dim3((L.in + 31) / 32, (nb + 31) / 32), 256, 0, s>>>(dhp, L.out, L.w, L.in, L.out, nb, i == 0 ? dfeat_ : dact_[i],
                                                                               i == 0 ? 512 : L.in);
      post_launch_t("linear_bwd");
    }
    // score-net parameter gradients
    for (int i = 0; i < nmlp_; ++i) {
      const MlpLayer& L = mlp_[i];
      const float* ain = i == 0 ? feat_ : act_[i];
      const int lda = L.out, ldb = i == 0 ? 512 : L.in;
      dim3 grid((L.in + 31) / 32, (L.out + 31) / 32);
      linear_wgrad_kernel<<<grid, 256, 0, s>>>(dh_ptr_[i], lda, L.out, ain, ldb, L.in, nb, mlp_gw_[i]);
      post_launch_t("linear_wgrad");
      colsum(dh_ptr_[i], L.out, L.out, nb, mlp_gb_[i], s);
      if (i < nmlp_ - 1) {
        colsum(dzx_[i], L.out, L.out, nb, mlp_gln_g_[i], s);
        colsum(dz_[i], L.out, L.out, nb, mlp_gln_b_[i], s);
      }
    }
    for (auto& e : enc_) backward_encoder(e, nb, s);
  }

  std::map<std::tuple<int, int, int, uint32_t>, cudaGraphExec_t> graphs_;
  cudaStream_t cap_ = nullptr;
  float *sx_ = nullptr, *sy_ = nullptr, *sr_ = nullptr, *smask_[3] = {};
  static constexpr int kColSplits = 256;
  static constexpr int kMaxSlices = 32;

  int kind_, B_, device_;
  WeightMap P_, G_;
  std::vector<void*> owned_;
  std::vector<TrainEncoder> enc_;
  MlpLayer mlp_[4];
  int nmlp_ = 0;
  float *act_[5] = {}, *pre_[4] = {}, *dact_[5] = {}, *stat_[3] = {}, *dh_[4] = {}, *dzx_[3] = {}, *dz_[3] = {};
  const float* dh_ptr_[4] = {};
  float *mlp_gw_[4] = {}, *mlp_gb_[4] = {}, *mlp_gln_g_[3] = {}, *mlp_gln_b_[3] = {};
  float *feat_ = nullptr, *dfeat_ = nullptr, *scores_ = nullptr, *dscore_ = nullptr, *out_ = nullptr;
  float *colsum_scratch_ = nullptr, *wpart_ = nullptr, *npart_ = nullptr, *nsum_ = nullptr;

  float* dalloc(size_t floats) {
    void* p = nullptr;
    floats = std::max<size_t>(floats, 4);
    RGFM_CUDA(cudaMalloc(&p, floats * sizeof(float)));
    RGFM_CUDA(cudaMemset(p, 0, floats * sizeof(float)));
    owned_.push_back(p);
    return (float*)p;
  }
  const float* param(const std::string& name, int64_t n) const { return P_.get(name, n); }
  float* grad(const std::string& name, int64_t n) const { return const_cast<float*>(G_.get(name, n)); }

  void add_encoder(const std::string& pre, const std::vector<TrainLayer>& L, bool gn, int in_ch, int img, bool from_y, int feat_off,
                   int fc_in) {
    TrainEncoder e;
    e.pre = pre; e.L = L; e.gn = gn; e.in_ch = in_ch; e.img = img; e.from_y = from_y; e.feat_off = feat_off; e.fc_in = fc_in;
    int h = img;
    const size_t nl = L.size();
    e.sp.resize(nl); e.cinP.resize(nl);
    e.opnd.resize(nl); e.raw.resize(nl); e.stats.assign(nl, nullptr); e.draw.resize(nl); e.dopnd.assign(nl, nullptr);
    e.wf.resize(nl); e.wb.assign(nl, nullptr); e.biasP.resize(nl); e.scale.assign(nl, nullptr); e.shift.assign(nl, nullptr);
    e.pa.resize(nl);
    int cmax = 0;
    for (size_t i = 0; i < nl; ++i) {
      const TrainLayer& l = L[i];
      RGFM_CHECK(256 % l.cout == 0, "ratio trainer: channel counts must divide 256");
      e.sp[i] = h;
      e.cinP[i] = i == 0 ? 16 : l.cin;
      const int coutP = (l.cout + 31) / 32 * 32, cinB = (l.cin + 31) / 32 * 32;
      e.opnd[i] = dalloc((size_t)B_ * h * h * e.cinP[i]);
      e.raw[i] = dalloc((size_t)B_ * h * h * l.cout);
      e.draw[i] = dalloc((size_t)B_ * h * h * l.cout);
      if (i > 0) e.dopnd[i] = dalloc((size_t)B_ * h * h * l.cin);
      const TileGeom g = make_geom(h, h);
      if (gn) e.stats[i] = dalloc((size_t)B_ * g.parts * l.cout * 2);
      e.wf[i] = dalloc((size_t)9 * e.cinP[i] * coutP);
      if (i > 0) e.wb[i] = dalloc((size_t)9 * l.cout * cinB);
      e.biasP[i] = dalloc(coutP);
      if (!gn) {
        e.scale[i] = dalloc(l.cout);
        e.shift[i] = dalloc(l.cout);
      }
      cmax = std::max(cmax, cinB);
      if (l.pool) h /= 2;
    }
    e.HWl = h * h;
    const int C = L.back().cout;
    RGFM_CHECK(C == fc_in, "ratio trainer: fc input width");
    e.cur_last = dalloc((size_t)B_ * e.HWl * C);
    e.dcur_last = dalloc((size_t)B_ * e.HWl * C);
    e.gap = dalloc((size_t)B_ * C);
    e.zero_bias = dalloc(cmax);
    // pack descriptors (norm + SiLU [+ max-pool] of layer i's raw output into layer i + 1's operand)
    for (size_t i = 0; i < nl; ++i) {
      const TrainLayer& l = L[i];
      PackArgs a{};
      PackSrc ps{};
      ps.p = e.raw[i]; ps.C = l.cout; ps.Cs = l.cout; ps.padded = 0; ps.fmt = 0;
      if (gn) {
        ps.st.p = e.stats[i]; ps.st.mode = ST_PARTS; ps.st.C = l.cout; ps.st.parts = make_geom(e.sp[i], e.sp[i]).parts;
      }
      a.src[0] = ps; a.src[1] = ps; a.nsrc = 1;
      a.Hs = a.Ws = e.sp[i];
      a.Ho = a.Wo = l.pool ? e.sp[i] / 2 : e.sp[i];
      a.C = l.cout; a.norm = gn ? NORM_GN : NORM_AFFINE; a.act = ACT_SILU; a.spatial = l.pool ? SP_POOL2 : SP_ID;
      if (gn) {
        a.gamma = param(pre + l.norm + ".weight", l.cout);
        a.beta = param(pre + l.norm + ".bias", l.cout);
      } else {
        a.gamma = e.scale[i];
        a.beta = e.shift[i];
      }
      a.groups = 8;
      a.out = (i + 1 < nl) ? e.opnd[i + 1] : e.cur_last;
      a.out_bf16 = 0;
      a.Cpad = l.cout;
      e.pa[i] = a;
    }
    enc_.push_back(std::move(e));
  }

  void add_mlp(const std::vector<std::pair<int, int>>& lin_ln, int last, const std::vector<int>& dims) {
    int in = 512;
    for (size_t i = 0; i < lin_ln.size(); ++i) {
      MlpLayer l{};
      const std::string a = "score_net." + std::to_string(lin_ln[i].first), n = "score_net." + std::to_string(lin_ln[i].second);
      l.in = in; l.out = dims[i];
      l.w = param(a + ".weight", (int64_t)l.in * l.out); l.b = param(a + ".bias", l.out);
      l.ln_g = param(n + ".weight", l.out); l.ln_b = param(n + ".bias", l.out);
      l.act = ACT_SILU;
      mlp_gw_[i] = grad(a + ".weight", (int64_t)l.in * l.out); mlp_gb_[i] = grad(a + ".bias", l.out);
      mlp_gln_g_[i] = grad(n + ".weight", l.out); mlp_gln_b_[i] = grad(n + ".bias", l.out);
      mlp_[nmlp_++] = l;
      in = l.out;
    }
    MlpLayer f{};
    const std::string a = "score_net." + std::to_string(last);
    f.in = in; f.out = 1; f.w = param(a + ".weight", in); f.b = param(a + ".bias", 1); f.act = ACT_NONE;
    mlp_gw_[nmlp_] = grad(a + ".weight", in); mlp_gb_[nmlp_] = grad(a + ".bias", 1);
    mlp_[nmlp_++] = f;
    for (int i = 0; i < nmlp_; ++i) {
      RGFM_CHECK(mlp_[i].in <= 512 && mlp_[i].out <= 512, "ratio trainer: score-net width");
      pre_[i] = dalloc((size_t)B_ * mlp_[i].out);
      if (i > 0) {
        act_[i] = dalloc((size_t)B_ * mlp_[i].in);
        dact_[i] = dalloc((size_t)B_ * mlp_[i].in);
      }
      if (i < nmlp_ - 1) {
        dh_[i] = dalloc((size_t)B_ * mlp_[i].out);
        dzx_[i] = dalloc((size_t)B_ * mlp_[i].out);
        dz_[i] = dalloc((size_t)B_ * mlp_[i].out);
        stat_[i] = dalloc((size_t)B_ * 2);
      }
    }
  }

  void refresh(cudaStream_t s) {
    for (auto& e : enc_)
      for (size_t i = 0; i < e.L.size(); ++i) {
        const TrainLayer& l = e.L[i];
        const int coutP = (l.cout + 31) / 32 * 32, cinB = (l.cin + 31) / 32 * 32;
        const int total = 9 * e.cinP[i] * coutP + (e.wb[i] ? 9 * l.cout * cinB : 0);
        repack_conv_w_kernel<<<std::min((total + 255) / 256, 1024), 256, 0, s>>>(param(e.pre + l.conv + ".weight", (int64_t)l.cout * l.cin * 9),
                                                                                  l.cout, l.cin, e.cinP[i], coutP, cinB, e.wf[i], e.wb[i]);
        post_launch_t("repack_conv_w");
        pad_vec_kernel<<<(coutP + 255) / 256, 256, 0, s>>>(param(e.pre + l.conv + ".bias", l.cout), l.cout, coutP, e.biasP[i]);
        post_launch_t("pad_vec");
        if (!e.gn) {
          const std::string p = e.pre + l.norm;
          bn_affine_kernel<<<(l.cout + 255) / 256, 256, 0, s>>>(param(p + ".weight", l.cout), param(p + ".bias", l.cout),
                                                                 param(p + ".running_mean", l.cout), param(p + ".running_var", l.cout),
                                                                 l.cout, e.scale[i], e.shift[i]);
          post_launch_t("bn_affine");
        }
      }
  }

  void forward_encoder(TrainEncoder& e, const float* img, int nb, cudaStream_t s) {
    launch_pack_nchw(img, e.in_ch, e.img, e.img, 16, e.opnd[0], 0, nb, s);
    for (size_t i = 0; i < e.L.size(); ++i) {
      const TrainLayer& l = e.L[i];
      ConvArgs a{};
      a.in = e.opnd[i]; a.Hin = a.Win = e.sp[i]; a.Cin = e.cinP[i];
      a.w = e.wf[i]; a.bias = e.biasP[i];
      a.CoutP = (l.cout + 31) / 32 * 32; a.Cout = l.cout; a.ks = 3; a.stride = 1; a.pad = 1;
      a.g = make_geom(e.sp[i], e.sp[i]);
      a.nimg = nb; a.out_mode = OUT_RAW; a.out = e.raw[i]; a.stats = e.stats[i];
      launch_conv_simt(a, s);
      PackArgs p = e.pa[i];
      p.nimg = nb;
      launch_pack(p, s);
    }
    const int C = e.L.back().cout;
    launch_gap_fc(e.cur_last, e.HWl, C, param(e.pre + "fc.weight", (int64_t)256 * C), param(e.pre + "fc.bias", 256), 256,
                  feat_ + e.feat_off, 512, nb, s);
  }

  // out[c] = sum over `rows` rows of M[r][c]
  void colsum(const float* M, int ld, int C, long long rows, float* out, cudaStream_t s) {
    RGFM_CHECK(C <= 512, "colsum: width");
    int splits = (int)std::min<long long>(kColSplits, (rows + 63) / 64);
    const int rps = (int)((rows + splits - 1) / splits);
    splits = (int)((rows + rps - 1) / rps);
    if (splits == 1) {
      colsum_kernel<<<dim3((C + 31) / 32, 1), 256, 0, s>>>(M, ld, C, rows, rps, out);
    } else {
      colsum_kernel<<<dim3((C + 31) / 32, splits), 256, 0, s>>>(M, ld, C, rows, rps, colsum_scratch_);
      colsum_kernel<<<dim3((C + 31) / 32, 1), 256, 0, s>>>(colsum_scratch_, C, C, splits, splits, out);
    }
    post_launch_t("colsum");
  }

  static int wgrad_rb(int H) { return std::max(1, std::min(H, 32 / H)); }  // rows per unit: about 32 pixels
  int wgrad_splits(const TrainEncoder& e, int i) const {
    const TrainLayer& l = e.L[i];
    const int tci = e.cinP[i] >= 64 ? 64 : e.cinP[i];
    const int tiles = (e.cinP[i] / tci) * ((l.cout + 63) / 64);
    const int rows = B_ * e.sp[i];
    return std::max(1, std::min(rows, (148 * 3 + tiles - 1) / tiles));
  }

  void backward_encoder(TrainEncoder& e, int nb, cudaStream_t s) {
    const int nl = (int)e.L.size(), C = e.L.back().cout;
    const float* fw = param(e.pre + "fc.weight", (int64_t)256 * C);
    launch_gap_fc_bwd(dfeat_ + e.feat_off, 512, fw, 256, e.HWl, C, 1.f, e.dcur_last, nb, s);
    gap_kernel<<<nb, 256, 0, s>>>(e.cur_last, e.HWl, C, e.gap);
    post_launch_t("gap");
    linear_wgrad_kernel<<<dim3((C + 31) / 32, 256 / 32), 256, 0, s>>>(dfeat_ + e.feat_off, 512, 256, e.gap, C, C, nb,
                                                                       grad(e.pre + "fc.weight", (int64_t)256 * C));
    post_launch_t("linear_wgrad");
    colsum(dfeat_ + e.feat_off, 512, 256, nb, grad(e.pre + "fc.bias", 256), s);
    for (int i = nl - 1; i >= 0; --i) {
      const TrainLayer& l = e.L[i];
      const int H = e.sp[i];
      const float* dcur = (i == nl - 1) ? e.dcur_last : e.dopnd[i + 1];
      PackArgs p = e.pa[i];
      p.nimg = nb;
      PackBwdIO io{};
      io.dcur_fmt = 0; io.dcur_padded = 0; io.dcur_Cs = l.cout; io.draw_padded = 0; io.draw_Cs = l.cout;
      // affine gradients of the norm (before pack_bwd only by convention: both read the same tensors)
      // pixel slices per image: enough CTAs to fill the GPU on small batches, at least 64 pixels each
      // pixel slices per image: about eight elements per thread (the kernels are latency-bound: several exp and dependent loads per element)
      const int slices = std::max(1, std::min(kMaxSlices, (H * H * l.cout + 2047) / 2048));
      norm_param_grad_kernel<<<dim3(nb, slices), 256, 0, s>>>(p, dcur, npart_);
      post_launch_t("norm_param_grad");
      colsum(npart_, 2 * l.cout, 2 * l.cout, (long long)nb * slices, nsum_, s);
      if (e.gn) {  // GroupNorm's group moments come from the same sums (u = xhat): pack_bwd needs no pass of its own
        io.chan_sums = npart_;
        io.nslices = slices;
      }
      io.yslices = slices;
      const std::string np = e.pre + l.norm;
      norm_grad_finish_kernel<<<(l.cout + 255) / 256, 256, 0, s>>>(nsum_, l.cout, e.gn ? nullptr : param(np + ".running_mean", l.cout),
                                                                   e.gn ? nullptr : param(np + ".running_var", l.cout),
                                                                   grad(np + ".weight", l.cout), grad(np + ".bias", l.cout));
      post_launch_t("norm_grad_finish");
      launch_pack_bwd(p, io, dcur, e.draw[i], s);
      // conv bias and weight gradients
      colsum(e.draw[i], l.cout, l.cout, (long long)nb * H * H, grad(e.pre + l.conv + ".bias", l.cout), s);
      {
        const int cinP = e.cinP[i];
        const int tci = cinP >= 64 ? 64 : cinP;
        const int tiles = (cinP / tci) * ((l.cout + 63) / 64);
        const int RB = wgrad_rb(H);
        const int units = nb * ((H + RB - 1) / RB);
        int splits = std::max(1, std::min(units, (148 * 3 + tiles - 1) / tiles));
        const int ups = (units + splits - 1) / splits;
        splits = (units + ups - 1) / ups;
        const size_t smem = ((size_t)(RB + 2) * (H + 2) * tci + (size_t)RB * H * 64) * sizeof(float);
        dim3 grid(tiles, splits);
        if (tci == 64) conv_wgrad_kernel<4><<<grid, 256, smem, s>>>(e.opnd[i], cinP, e.draw[i], l.cout, H, H, RB, nb, ups, wpart_);
        else if (tci == 32) conv_wgrad_kernel<2><<<grid, 256, smem, s>>>(e.opnd[i], cinP, e.draw[i], l.cout, H, H, RB, nb, ups, wpart_);
        else conv_wgrad_kernel<1><<<grid, 256, smem, s>>>(e.opnd[i], cinP, e.draw[i], l.cout, H, H, RB, nb, ups, wpart_);
        post_launch_t("conv_wgrad");
        const int total = l.cout * l.cin * 9;
        conv_wgrad_reduce_kernel<<<std::min((total + 255) / 256, 2048), 256, 0, s>>>(wpart_, splits, cinP, l.cout, l.cin,
                                                                                      grad(e.pre + l.conv + ".weight", (int64_t)total));
        post_launch_t("conv_wgrad_reduce");
      }
      if (i > 0) {  // input gradient of the conv = gradient at the previous layer's pack output
        ConvArgs a{};
        a.in = e.draw[i]; a.Hin = a.Win = H; a.Cin = l.cout;
        a.w = e.wb[i]; a.bias = e.zero_bias;
        a.CoutP = (l.cin + 31) / 32 * 32; a.Cout = l.cin; a.ks = 3; a.stride = 1; a.pad = 1;
        a.g = make_geom(H, H);
        a.nimg = nb; a.out_mode = OUT_RAW; a.out = e.dopnd[i];
        launch_conv_simt(a, s);
      }
    }
  }
};


// ------------------------------------------------------------------------------------------------
// Convolution ops for the flow-matching training step (flow_utils.py:103-156 trains the U-Nets of
// unet.py / unet_flexible.py): forward, input gradient, weight gradient and bias gradient of nn.Conv2d(k = 3, padding = 1,
// stride 1 or 2) on dense fp32 NHWC tensors (torch "channels_last" memory) with the weight in torch's own
// [Cout][Cin][3][3] layout, so that a torch.autograd.Function can hand over its tensors as they are (fm_train.py).
// The same kernels as the ratio trainer: conv_simt_kernel for the forward and -- on the flipped / transposed filter --
// the input gradient, conv_wgrad_kernel + a fixed-order reduction for dW, colsum_kernel for db.  A stride-2 conv
// (Downsample, unet.py:99-107) is differentiated on the full-resolution grid: its output gradient with zeros inserted
// at the odd positions turns both gradients into those of a stride-1 conv.  Scratch is stream-ordered
// (cudaMallocAsync / cudaFreeAsync: the driver's pool, no synchronisation).  Exact fp32, deterministic.
// ------------------------------------------------------------------------------------------------
__global__ void pad_channels_kernel(const float* __restrict__ src, int Cs, float* __restrict__ dst, int Cd, size_t rows) {
  const size_t total = rows * (size_t)Cd;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / Cd;
    const int c = (int)(i - r * Cd);
    dst[i] = c < Cs ? src[r * Cs + c] : 0.f;
  }
}
// dst[n][2h][2w][c] = src[n][h][w][c] (c < Cs), zeros elsewhere; dst is [n][2 Ho][2 Wo][Cd]
__global__ void zero_insert_kernel(const float* __restrict__ src, int Ho, int Wo, int Cs, float* __restrict__ dst, int Cd, int nimg) {
  const int H = 2 * Ho, W = 2 * Wo;
  const size_t total = (size_t)nimg * H * W * Cd;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cd);
    size_t q = i / Cd;
    const int w = (int)(q % W);
    q /= W;
    const int h = (int)(q % H);
    const int n = (int)(q / H);
    float v = 0.f;
    if (c < Cs && !(h & 1) && !(w & 1)) v = src[(((size_t)n * Ho + (h >> 1)) * Wo + (w >> 1)) * Cs + c];
    dst[i] = v;
  }
}
// forward filter wf[tap][ci][co] (ci < CinP, co < CoutP, zero padded)
__global__ void repack_w_fwd_kernel(const float* __restrict__ w, int cout, int cin, int CinP, int CoutP, float* __restrict__ wf) {
  const int total = 9 * CinP * CoutP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % CoutP, ci = (i / CoutP) % CinP, t = i / (CoutP * CinP);
    wf[i] = (co < cout && ci < cin) ? w[((size_t)co * cin + ci) * 9 + t] : 0.f;
  }
}
// input-gradient filter wb[tap][co][ci] = W[co][ci][8 - tap] (co < CoutK rows = the dgrad conv's K, ci < CinB its output channels)
__global__ void repack_w_bwd_kernel(const float* __restrict__ w, int cout, int cin, int CoutK, int CinB, float* __restrict__ wb) {
  const int total = 9 * CoutK * CinB;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % CinB, co = (i / CinB) % CoutK, t = i / (CinB * CoutK);
    wb[i] = (co < cout && ci < cin) ? w[((size_t)co * cin + ci) * 9 + (8 - t)] : 0.f;
  }
}
// like conv_wgrad_reduce_kernel with a padded channel stride of the partials (CoutS >= cout)
__global__ void conv_wgrad_reduce_padded_kernel(const float* __restrict__ part, int nsplit, int CinP, int CoutS, int cin, int cout,
                                                float* __restrict__ dW) {
  const int total = cout * cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % cout, ci = (i / cout) % cin, t = i / (cout * cin);
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += part[(((size_t)k * 9 + t) * CinP + ci) * CoutS + co];
    dW[((size_t)co * cin + ci) * 9 + t] = s;
  }
}

namespace {
struct StreamScratch {  // stream-ordered scratch of one op call
  cudaStream_t s;
  std::vector<void*> ptrs;
  explicit StreamScratch(cudaStream_t st) : s(st) {
    // keep freed scratch in the device's default pool across synchronisations (the default release threshold of 0 hands it
    // back to the driver at every sync -- a training loop synchronises once per step for loss.item())
    thread_local int tuned_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev != tuned_dev) {
      cudaMemPool_t pool = nullptr;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
      tuned_dev = dev;
    }
  }
  float* get(size_t floats) {
    void* p = nullptr;
    RGFM_CUDA(cudaMallocAsync(&p, std::max<size_t>(floats, 4) * sizeof(float), s));
    ptrs.push_back(p);
    return (float*)p;
  }
  ~StreamScratch() {
    for (void* p : ptrs) cudaFreeAsync(p, s);
  }
};
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline unsigned blocks_for(size_t n) { return (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16); }
}  // namespace

void conv3x3_forward(const float* x, int n, int H, int W, int Cin, const float* w, const float* bias, int Cout, int stride, float* y,
                     cudaStream_t s) {
  RGFM_CHECK(x && w && y && n >= 1 && H >= 1 && W >= 1 && Cin >= 1 && Cout >= 1, "rgfm_conv3x3_forward: bad argument");
  RGFM_CHECK(stride == 1 || (stride == 2 && H % 2 == 0 && W % 2 == 0), "rgfm_conv3x3_forward: stride 1, or stride 2 on even sizes");
  StreamScratch sc(s);
  const int CinP = round_up(Cin, 16), CoutP = round_up(Cout, 32);
  const float* xp = x;
  if (CinP != Cin) {
    float* t = sc.get((size_t)n * H * W * CinP);
    pad_channels_kernel<<<blocks_for((size_t)n * H * W * CinP), 256, 0, s>>>(x, Cin, t, CinP, (size_t)n * H * W);
    post_launch_t("pad_channels");
    xp = t;
  }
  float* wf = sc.get((size_t)9 * CinP * CoutP);
  repack_w_fwd_kernel<<<blocks_for((size_t)9 * CinP * CoutP), 256, 0, s>>>(w, Cout, Cin, CinP, CoutP, wf);
  post_launch_t("repack_w_fwd");
  float* bp = sc.get(CoutP);
  if (bias) {
    pad_vec_kernel<<<(CoutP + 255) / 256, 256, 0, s>>>(bias, Cout, CoutP, bp);
    post_launch_t("pad_vec");
  } else {
    RGFM_CUDA(cudaMemsetAsync(bp, 0, CoutP * sizeof(float), s));
  }
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  ConvArgs a{};
  a.in = xp; a.Hin = H; a.Win = W; a.Cin = CinP;
  a.w = wf; a.bias = bp;
  a.CoutP = CoutP; a.Cout = Cout; a.ks = 3; a.stride = stride; a.pad = 1;
  a.g = make_geom(Ho, Wo);
  a.nimg = n; a.out_mode = OUT_RAW; a.out = y;
  launch_conv_simt(a, s);
}

void conv3x3_backward(const float* x, const float* dy, int n, int H, int W, int Cin, const float* w, int Cout, int stride, float* dx,
                      float* dw, float* db, cudaStream_t s) {
  RGFM_CHECK(dy && n >= 1 && H >= 1 && W >= 1 && Cin >= 1 && Cout >= 1, "rgfm_conv3x3_backward: bad argument");
  RGFM_CHECK(stride == 1 || (stride == 2 && H % 2 == 0 && W % 2 == 0), "rgfm_conv3x3_backward: stride 1, or stride 2 on even sizes");
  RGFM_CHECK((!dw || x) && (!dx || w), "rgfm_conv3x3_backward: dw needs x, dx needs w");
  StreamScratch sc(s);
  const int Ho = H / stride, Wo = W / stride;
  const size_t rows_o = (size_t)n * Ho * Wo, rows = (size_t)n * H * W;
  if (db) {  // bias gradient: column sums of dy in a fixed two-stage order
    int splits = (int)std::min<size_t>(64, (rows_o + 63) / 64);
    const int rps = (int)((rows_o + splits - 1) / splits);
    splits = (int)((rows_o + rps - 1) / rps);
    if (splits == 1) {
      colsum_kernel<<<dim3((Cout + 31) / 32, 1), 256, 0, s>>>(dy, Cout, Cout, (long long)rows_o, rps, db);
    } else {
      float* tmp = sc.get((size_t)splits * Cout);
      colsum_kernel<<<dim3((Cout + 31) / 32, splits), 256, 0, s>>>(dy, Cout, Cout, (long long)rows_o, rps, tmp);
      colsum_kernel<<<dim3((Cout + 31) / 32, 1), 256, 0, s>>>(tmp, Cout, Cout, splits, splits, db);
    }
    post_launch_t("colsum");
  }
  if (!dw && !dx) return;
  // the output gradient on the full-resolution grid, channels padded to a multiple of 16 (the dgrad conv's K, float4 loads of wgrad)
  const int CoQ = round_up(Cout, 16);
  const float* g = dy;
  if (stride == 2) {
    float* t = sc.get(rows * CoQ);
    zero_insert_kernel<<<blocks_for(rows * CoQ), 256, 0, s>>>(dy, Ho, Wo, Cout, t, CoQ, n);
    post_launch_t("zero_insert");
    g = t;
  } else if (CoQ != Cout) {
    float* t = sc.get(rows * CoQ);
    pad_channels_kernel<<<blocks_for(rows * CoQ), 256, 0, s>>>(dy, Cout, t, CoQ, rows);
    post_launch_t("pad_channels");
    g = t;
  }
  if (dw) {
    const int CinP = round_up(Cin, 16);
    const float* xp = x;
    if (CinP != Cin) {
      float* t = sc.get(rows * CinP);
      pad_channels_kernel<<<blocks_for(rows * CinP), 256, 0, s>>>(x, Cin, t, CinP, rows);
      post_launch_t("pad_channels");
      xp = t;
    }
    const int tci = CinP % 64 == 0 ? 64 : (CinP % 32 == 0 ? 32 : 16);
    const int tiles = (CinP / tci) * ((CoQ + 63) / 64);
    const int RB = std::max(1, std::min(H, 32 / std::max(1, W)));  // rows per unit: about 32 pixels
    const int units = n * ((H + RB - 1) / RB);
    int splits = std::max(1, std::min(units, (148 * 3 + tiles - 1) / tiles));
    const int ups = (units + splits - 1) / splits;
    splits = (units + ups - 1) / ups;
    const size_t smem = ((size_t)(RB + 2) * (W + 2) * tci + (size_t)RB * W * 64) * sizeof(float);
    RGFM_CHECK(smem <= 48 * 1024, "rgfm_conv3x3_backward: image too wide for the weight-gradient tile");
    float* part = sc.get((size_t)splits * 9 * CinP * CoQ);
    dim3 grid(tiles, splits);
    if (tci == 64) conv_wgrad_kernel<4><<<grid, 256, smem, s>>>(xp, CinP, g, CoQ, H, W, RB, n, ups, part);
    else if (tci == 32) conv_wgrad_kernel<2><<<grid, 256, smem, s>>>(xp, CinP, g, CoQ, H, W, RB, n, ups, part);
    else conv_wgrad_kernel<1><<<grid, 256, smem, s>>>(xp, CinP, g, CoQ, H, W, RB, n, ups, part);
    post_launch_t("conv_wgrad");
    const int total = Cout * Cin * 9;
    conv_wgrad_reduce_padded_kernel<<<std::min((total + 255) / 256, 2048), 256, 0, s>>>(part, splits, CinP, CoQ, Cin, Cout, dw);
    post_launch_t("conv_wgrad_reduce");
  }
  if (dx) {
    const int CinB = round_up(Cin, 32);
    float* wb = sc.get((size_t)9 * CoQ * CinB);
    repack_w_bwd_kernel<<<blocks_for((size_t)9 * CoQ * CinB), 256, 0, s>>>(w, Cout, Cin, CoQ, CinB, wb);
    post_launch_t("repack_w_bwd");
    float* zb = sc.get(CinB);
    RGFM_CUDA(cudaMemsetAsync(zb, 0, CinB * sizeof(float), s));
    ConvArgs a{};
    a.in = g; a.Hin = H; a.Win = W; a.Cin = CoQ;
    a.w = wb; a.bias = zb;
    a.CoutP = CinB; a.Cout = Cin; a.ks = 3; a.stride = 1; a.pad = 1;
    a.g = make_geom(H, W);
    a.nimg = n; a.out_mode = OUT_RAW; a.out = dx;
    launch_conv_simt(a, s);
  }
}


// ------------------------------------------------------------------------------------------------
// GroupNorm on dense fp32 NHWC tensors, forward and backward, for the flow-matching training step (nn.GroupNorm(8, C) of
// unet.py:54,64,209 / unet_flexible.py:51,61,196).  PyTorch's CUDA GroupNorm works in NCHW: on the channels_last tensors the
// conv ops produce it converts its input and returns NCHW, which the next conv converts back (7 % of a step in copies).
// Statistics: per (image, pixel slice, channel) partial sums in fp32, combined per group in double in a fixed order
// (deterministic); stats[n][g] = {mean, rstd}.  Backward (x_hat = (x - mean) rstd, m = pixels x channels per group):
//   dgamma_c = sum dy x_hat,  dbeta_c = sum dy,  dx = rstd (dy gamma - (x_hat A_g + B_g) / m),
//   A_g = sum_{c in g} gamma_c sum_hw dy x_hat,  B_g = sum_{c in g} gamma_c sum_hw dy.
// ------------------------------------------------------------------------------------------------
// partial[n][slice][c][2]: MODE 0: {sum x, sum x^2}; MODE 1: {sum dy x_hat, sum dy}
template <int MODE>
__global__ void __launch_bounds__(256) gn_nhwc_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              const float* __restrict__ stats, int HW, int C, int groups, int slices,
                                                              float* __restrict__ part) {
  __shared__ float4 red[2][256];
  const int n = blockIdx.y, sl = blockIdx.x, tid = threadIdx.x;
  const int Q = C >> 2, lanes = 256 / Q;  // channel quads, pixel lanes
  const int q = tid % Q, pl = tid / Q;
  const int per = (HW + slices - 1) / slices, p0 = sl * per, p1 = min(HW, p0 + per);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (pl < lanes) {
    float mean = 0.f, rstd = 0.f;
    if (MODE == 1) {
      const int g = (q * 4) / (C / groups);
      mean = stats[((size_t)n * groups + g) * 2];
      rstd = stats[((size_t)n * groups + g) * 2 + 1];
    }
    for (int p = p0 + pl; p < p1; p += lanes) {
      const size_t o = ((size_t)n * HW + p) * C + q * 4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + o));
      if (MODE == 0) {
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        b.x = fmaf(v.x, v.x, b.x); b.y = fmaf(v.y, v.y, b.y); b.z = fmaf(v.z, v.z, b.z); b.w = fmaf(v.w, v.w, b.w);
      } else {
        const float4 d = __ldg(reinterpret_cast<const float4*>(dy + o));
        a.x = fmaf(d.x, (v.x - mean) * rstd, a.x); a.y = fmaf(d.y, (v.y - mean) * rstd, a.y);
        a.z = fmaf(d.z, (v.z - mean) * rstd, a.z); a.w = fmaf(d.w, (v.w - mean) * rstd, a.w);
        b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
      }
    }
  }
  red[0][tid] = a;
  red[1][tid] = b;
  __syncthreads();
  if (tid < Q) {  // pixel lanes in increasing order
    float4 sa = red[0][tid], sb = red[1][tid];
    for (int l = 1; l < lanes; ++l) {
      const float4 ua = red[0][l * Q + tid], ub = red[1][l * Q + tid];
      sa.x += ua.x; sa.y += ua.y; sa.z += ua.z; sa.w += ua.w;
      sb.x += ub.x; sb.y += ub.y; sb.z += ub.z; sb.w += ub.w;
    }
    float* o = part + (((size_t)n * slices + sl) * C + tid * 4) * 2;
    o[0] = sa.x; o[1] = sb.x; o[2] = sa.y; o[3] = sb.y; o[4] = sa.z; o[5] = sb.z; o[6] = sa.w; o[7] = sb.w;
  }
}
// one block per image.  MODE 0: stats[n][g] = {mean, rstd}.  MODE 1: chan[n][c] = {sum dy x_hat, sum dy} and gsum[n][g] = {A_g, B_g}.
template <int MODE>
__global__ void __launch_bounds__(256) gn_nhwc_finish_kernel(const float* __restrict__ part, int HW, int C, int groups, int slices,
                                                             const float* __restrict__ gamma, float eps, float* __restrict__ stats,
                                                             float* __restrict__ chan) {
  __shared__ double s0[1024], s1[1024];
  const int n = blockIdx.x, cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += 256) {
    double a = 0.0, b = 0.0;
    for (int sl = 0; sl < slices; ++sl) {
      const float* p = part + (((size_t)n * slices + sl) * C + c) * 2;
      a += (double)p[0];
      b += (double)p[1];
    }
    if (MODE == 1) {
      chan[((size_t)n * C + c) * 2] = (float)a;
      chan[((size_t)n * C + c) * 2 + 1] = (float)b;
      a *= (double)gamma[c];
      b *= (double)gamma[c];
    }
    s0[c] = a;
    s1[c] = b;
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    double a = 0.0, b = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      a += s0[c];
      b += s1[c];
    }
    float* o = stats + ((size_t)n * groups + g) * 2;
    if (MODE == 0) {
      const double m = (double)cpg * HW, mean = a / m;
      double var = b / m - mean * mean;
      if (var < 0.0) var = 0.0;
      o[0] = (float)mean;
      o[1] = (float)(1.0 / sqrt(var + (double)eps));
    } else {
      o[0] = (float)a;
      o[1] = (float)b;
    }
  }
}
__global__ void __launch_bounds__(256) gn_nhwc_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, int HW, int C,
                                                            int groups, size_t total4, float* __restrict__ y) {
  const int Q = C >> 2, cpg = C / groups;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % Q);
    const int n = (int)(i / ((size_t)Q * HW));
    const int g = (q * 4) / cpg;
    const float mean = stats[((size_t)n * groups + g) * 2], rstd = stats[((size_t)n * groups + g) * 2 + 1];
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + q), be = __ldg(reinterpret_cast<const float4*>(beta) + q);
    float4 o;
    o.x = fmaf((v.x - mean) * rstd, ga.x, be.x); o.y = fmaf((v.y - mean) * rstd, ga.y, be.y);
    o.z = fmaf((v.z - mean) * rstd, ga.z, be.z); o.w = fmaf((v.w - mean) * rstd, ga.w, be.w);
    reinterpret_cast<float4*>(y)[i] = o;
  }
}
__global__ void __launch_bounds__(256) gn_nhwc_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                const float* __restrict__ stats, const float* __restrict__ gsum,
                                                                const float* __restrict__ gamma, int HW, int C, int groups, size_t total4,
                                                                float* __restrict__ dx) {
  const int Q = C >> 2, cpg = C / groups;
  const float inv_m = 1.0f / ((float)cpg * (float)HW);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % Q);
    const int n = (int)(i / ((size_t)Q * HW));
    const int g = (q * 4) / cpg;
    const size_t sg = ((size_t)n * groups + g) * 2;
    const float mean = stats[sg], rstd = stats[sg + 1], A = gsum[sg] * inv_m, B = gsum[sg + 1] * inv_m;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i), d = __ldg(reinterpret_cast<const float4*>(dy) + i);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + q);
    float4 o;
    o.x = rstd * (d.x * ga.x - fmaf((v.x - mean) * rstd, A, B)); o.y = rstd * (d.y * ga.y - fmaf((v.y - mean) * rstd, A, B));
    o.z = rstd * (d.z * ga.z - fmaf((v.z - mean) * rstd, A, B)); o.w = rstd * (d.w * ga.w - fmaf((v.w - mean) * rstd, A, B));
    reinterpret_cast<float4*>(dx)[i] = o;
  }
}
// dgamma[c] = sum_n chan[n][c][0], dbeta[c] = sum_n chan[n][c][1] (image order)
__global__ void gn_nhwc_param_grad_kernel(const float* __restrict__ chan, int nimg, int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int n = 0; n < nimg; ++n) {
    a += (double)chan[((size_t)n * C + c) * 2];
    b += (double)chan[((size_t)n * C + c) * 2 + 1];
  }
  if (dgamma) dgamma[c] = (float)a;
  if (dbeta) dbeta[c] = (float)b;
}

namespace {
inline int gn_slices(int n, int HW) { return std::max(1, std::min(std::max(1, HW / 64), (2 * 148 + n - 1) / n)); }
inline void gn_check(int n, int HW, int C, int groups, const char* what) {
  RGFM_CHECK(n >= 1 && HW >= 1 && groups >= 1 && groups <= 256 && C % groups == 0 && (C / groups) % 4 == 0 && C <= 1024,
             std::string(what) + ": channels per group must be a multiple of 4, at most 1024 channels");
}
}  // namespace

void gn_nhwc_forward(const float* x, int n, int HW, int C, int groups, const float* gamma, const float* beta, float eps, float* y,
                     float* stats, cudaStream_t s) {
  gn_check(n, HW, C, groups, "rgfm_groupnorm_nhwc_forward");
  StreamScratch sc(s);
  const int slices = gn_slices(n, HW);
  float* part = sc.get((size_t)n * slices * C * 2);
  gn_nhwc_partial_kernel<0><<<dim3(slices, n), 256, 0, s>>>(x, nullptr, nullptr, HW, C, groups, slices, part);
  post_launch_t("gn_nhwc_partial");
  gn_nhwc_finish_kernel<0><<<n, 256, 0, s>>>(part, HW, C, groups, slices, nullptr, eps, stats, nullptr);
  post_launch_t("gn_nhwc_finish");
  const size_t total4 = (size_t)n * HW * C / 4;
  gn_nhwc_apply_kernel<<<blocks_for(total4), 256, 0, s>>>(x, stats, gamma, beta, HW, C, groups, total4, y);
  post_launch_t("gn_nhwc_apply");
}

void gn_nhwc_backward(const float* x, const float* dy, int n, int HW, int C, int groups, const float* gamma, const float* stats,
                      float* dx, float* dgamma, float* dbeta, cudaStream_t s) {
  gn_check(n, HW, C, groups, "rgfm_groupnorm_nhwc_backward");
  StreamScratch sc(s);
  const int slices = gn_slices(n, HW);
  float* part = sc.get((size_t)n * slices * C * 2);
  float* chan = sc.get((size_t)n * C * 2);
  float* gsum = sc.get((size_t)n * groups * 2);
  gn_nhwc_partial_kernel<1><<<dim3(slices, n), 256, 0, s>>>(x, dy, stats, HW, C, groups, slices, part);
  post_launch_t("gn_nhwc_partial");
  gn_nhwc_finish_kernel<1><<<n, 256, 0, s>>>(part, HW, C, groups, slices, gamma, 0.f, gsum, chan);
  post_launch_t("gn_nhwc_finish");
  if (dgamma || dbeta) {
    gn_nhwc_param_grad_kernel<<<(C + 127) / 128, 128, 0, s>>>(chan, n, C, dgamma, dbeta);
    post_launch_t("gn_nhwc_param_grad");
  }
  if (dx) {
    const size_t total4 = (size_t)n * HW * C / 4;
    gn_nhwc_bwd_apply_kernel<<<blocks_for(total4), 256, 0, s>>>(x, dy, stats, gsum, gamma, HW, C, groups, total4, dx);
    post_launch_t("gn_nhwc_bwd_apply");
  }
}

}  // namespace rgfm

// ---- C ABI (include/rgfm.h) ----
struct rgfm_ratio_trainer {
  std::unique_ptr<rgfm::RatioTrainer> t;
};
namespace rgfm {
rgfm_ratio_trainer* ratio_trainer_create(int kind, const rgfm_weights* params, const rgfm_weights* grads, int max_batch, int device) {
  std::unique_ptr<rgfm_ratio_trainer> h(new rgfm_ratio_trainer());
  h->t.reset(new RatioTrainer(kind, params, grads, max_batch, device));
  return h.release();
}
void ratio_trainer_step(rgfm_ratio_trainer* h, const float* x, const float* y, const float* is_real, int nb, int loss_type, float lambda,
                        const float* const* masks, int n_masks, float* out8, cudaStream_t s) {
  RGFM_CUDA(cudaSetDevice(h->t->device()));
  h->t->step(x, y, is_real, nb, loss_type, lambda, masks, n_masks, out8, s);
}
void ratio_trainer_destroy(rgfm_ratio_trainer* h) {
  if (h) cudaSetDevice(h->t->device());
  delete h;
}
}  // namespace rgfm

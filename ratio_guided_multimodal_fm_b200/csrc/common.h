// Shared declarations for librgfm.so (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdexcept>
#include <string>
#include <vector>
#include <map>
#include <memory>
#include <cstdio>
#include <utility>
#include <cstdlib>

namespace rgfm {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define RGFM_CUDA(expr)                                                                                  \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      throw ::rgfm::Error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ +  \
                          ":" + std::to_string(__LINE__) + ")");                                         \
  } while (0)

#define RGFM_CHECK(cond, msg)                                                                     \
  do {                                                                                            \
    if (!(cond)) throw ::rgfm::Error(std::string(msg) + " [" #cond "] (" + __FILE__ + ":" +       \
                                     std::to_string(__LINE__) + ")");                             \
  } while (0)

constexpr int TM = 128;  // output pixels per conv tile (one tcgen05 M=128 accumulator)

// Tiling of an H x W output image into <=128-pixel tiles of whole rows (or several whole images per tile when
// H*W < 128).  Shared by every conv implementation so per-tile GroupNorm partial sums have one layout.
struct TileGeom {
  int H, W;    // output spatial size
  int th;      // rows per tile
  int imgs;    // images per tile (1 unless H*W*2 <= TM)
  int parts;   // tiles per image (1 when imgs > 1)
  __host__ __device__ int tiles(int nimg) const { return imgs == 1 ? nimg * parts : (nimg + imgs - 1) / imgs; }
};

inline TileGeom make_geom(int H, int W) {
  TileGeom g;
  g.H = H;
  g.W = W;
  if (H * W >= TM) {
    g.th = TM / W;
    g.imgs = 1;
    g.parts = (H + g.th - 1) / g.th;
  } else {
    g.th = H;
    g.imgs = TM / (H * W);
    g.parts = 1;
  }
  return g;
}

// Row r of tile `tile` -> (n, ho, wo, part); returns validity.
__host__ __device__ inline bool tile_row(const TileGeom& g, int tile, int r, int nimg, int& n, int& ho, int& wo,
                                         int& part) {
  if (g.imgs == 1) {
    n = tile / g.parts;
    part = tile - n * g.parts;
    if (r >= g.th * g.W) return false;
    int rr = r / g.W;
    ho = part * g.th + rr;
    wo = r - rr * g.W;
    return ho < g.H && n < nimg;
  }
  int hw = g.H * g.W;
  int il = r / hw;
  if (il >= g.imgs) return false;
  n = tile * g.imgs + il;
  part = 0;
  int q = r - il * hw;
  ho = q / g.W;
  wo = q - ho * g.W;
  return n < nimg;
}

// Raw (pre-norm) activation: fp32 NHWC plus optional per-(image, part, channel) {sum, sumsq} partials.
// Pixel (n, h, w) of a raw activation -> element offset of its channel 0.  Dense NHWC, or the padded-linear layout
// the tcgen05 conv stores with TMA: row index = GEMM row m' = (n*(H+1) + h)*(W+2) + w (pad rows/columns hold garbage).
__host__ __device__ inline size_t raw_offset(int padded, int n, int h, int w, int H, int W, int C) {
  return padded ? ((size_t)((size_t)n * (H + 1) + h) * (W + 2) + w) * C : (((size_t)n * H + h) * W + w) * C;
}

// GroupNorm partial statistics of a raw activation, {sum, sumsq} per channel, in one of two layouts:
//   ST_PARTS: [image][parts][C][2]              (SIMT conv: one partial per output tile of the image)
//   ST_UNITS: [unit][slots][C][2]               (tcgen05 conv: one partial per unit of unit_rows rows of the padded-linear
//             GEMM row space and per image slot inside the unit; image n owns rows [n*R, (n+1)*R), the first image of
//             unit u is (u*unit_rows)/R and slot s is image (u*unit_rows)/R + s)
// Partials are combined by the consumer in a fixed order, so results are deterministic.
enum StatsMode { ST_NONE = 0, ST_PARTS = 1, ST_UNITS = 2 };
struct StatsRef {
  const float* p = nullptr;
  int mode = ST_NONE;
  int C = 0;      // channel stride of a partial row
  int parts = 0;  // ST_PARTS
  int R = 0;      // ST_UNITS: GEMM rows per image
  int slots = 0;  // ST_UNITS: image slots per unit
  int unit_rows = 128;            // ST_UNITS: GEMM rows per unit (128, or 127 / 126 when the producing conv stacks taps)
  unsigned int magic_unit = 0;    // ceil(2^32 / unit_rows): x / unit_rows = umulhi(x, magic) for x * (unit_rows - 1) < 2^32
};

__host__ __device__ inline int stats_slots_for(int R) { return 127 / R + 2; }

// {sum, sumsq} of channel c of image n, combined from the partials in a fixed order.  BATCH partials are in flight at a
// time (a rolled loop would wait out one global-load latency per unit); the sum runs in unit order whatever BATCH is.
template <int BATCH = 4>
__device__ __forceinline__ void stats_channel_sums(const StatsRef& st, int n, int c, float& s, float& ss) {
  s = 0.f;
  ss = 0.f;
  if (st.mode == ST_PARTS) {
    const float* q = st.p + ((size_t)n * st.parts * st.C + c) * 2;
    for (int p = 0; p < st.parts; ++p) {
      s += q[(size_t)p * st.C * 2];
      ss += q[(size_t)p * st.C * 2 + 1];
    }
  } else if (st.mode == ST_UNITS) {
    const int u0 = (int)__umulhi((unsigned)(n * st.R), st.magic_unit), u1 = (int)__umulhi((unsigned)((n + 1) * st.R - 1), st.magic_unit);
    const float inv_R = 1.0f / (float)st.R;
    for (int ub = u0; ub <= u1; ub += BATCH) {
      float2 v[BATCH];
#pragma unroll
      for (int j = 0; j < BATCH; ++j) {
        const int u = ub + j;
        v[j] = make_float2(0.f, 0.f);
        if (u <= u1) {
          // image of the unit's first row: floor(128 u / R) by a float reciprocal + exact correction (an integer
          // division by a run-time divisor was half of gn_coef's instructions; 128 u < 2^24 is exact in fp32)
          const int x = u * st.unit_rows;
          int q = (int)((float)x * inv_R);
          q += ((q + 1) * st.R <= x) ? 1 : 0;
          q -= (q * st.R > x) ? 1 : 0;
          const int slot = n - q;
          v[j] = __ldg(reinterpret_cast<const float2*>(st.p + (((size_t)u * st.slots + slot) * st.C + c) * 2));
        }
      }
#pragma unroll
      for (int j = 0; j < BATCH; ++j) {
        s += v[j].x;
        ss += v[j].y;
      }
    }
  }
}

struct Act {
  float* p = nullptr;      // element type per `fmt` (the pointer is reinterpreted for 16-bit storage)
  int H = 0, W = 0, C = 0;
  int Cs = 0;              // channel stride of a pixel row (>= C; tcgen05 raws pad it to a multiple of 64 with zeros)
  bool padded = false;     // padded-linear layout (tcgen05 path); pad rows / columns hold exact zeros
  int fmt = 0;             // storage: 0 = fp32, 1 = bf16, 2 = fp16 (tensor-core nets keep raw activations in 16 bits)
  float* stats = nullptr;  // partial GroupNorm statistics (layout per `st`)
  StatsRef st;
  int parts = 0;
};

// Normalised/activated conv operand: NHWC, fp32 (SIMT path) or bf16 (tcgen05 path)
struct Packed {
  void* p = nullptr;
  int H = 0, W = 0, C = 0;
  bool bf16 = false;    // 16-bit operand (bf16 or fp16 per the net's precision)
  bool padded = false;  // padded-linear layout [slot][C] (tcgen05 path), see pack_padded_kernel
};

enum NormMode { NORM_NONE = 0, NORM_GN = 1, NORM_AFFINE = 2 };
// SP_D2S: depth-to-space x2 of a [H][W][4*C] source whose channel (a*2+b)*C + c is output pixel (2h+a, 2w+b), channel c --
// a ConvTranspose2d(4, stride 2, padding 1) computed as a 3x3 convolution with 4*C_out outputs (fp32 pack kernel only)
enum SpatialMode { SP_ID = 0, SP_UP2 = 1, SP_POOL2 = 2, SP_D2S = 3 };
// OUT_NCHW_SCALED: out = dt * val (the last dgrad conv of the 16-bit gradient chain removes its loss scale)
enum OutMode { OUT_RAW = 0, OUT_NCHW = 1, OUT_EULER = 2, OUT_NCHW_SCALED = 3 };
enum ActMode { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2 };

// ---- kernel launch descriptors (plain structs passed by value) ----
struct PackSrc {
  const float* p;      // raw NHWC
  StatsRef st;         // GroupNorm partials (mode ST_NONE when norm != GN)
  int C;               // real channels taken from this source
  int Cs;              // channel stride of the source rows
  int padded;          // source layout (see raw_offset)
  int fmt;             // source storage: 0 = fp32, 1 = bf16, 2 = fp16
};

struct PackArgs {
  PackSrc src[2];
  int nsrc;
  int Hs, Ws;           // source spatial
  int Ho, Wo;           // output spatial
  int C;                // total channels
  int norm, act, spatial;
  const float* gamma;   // GN: gamma[C]; AFFINE: scale[C]
  const float* beta;    // GN: beta[C];  AFFINE: shift[C]
  int groups;
  void* out;
  int out_bf16;         // 0 = fp32, 1 = bf16, 2 = fp16
  int nimg;
  int Cpad;             // padded-linear layout only: channels of the operand tensor (multiple of 64, zero filled)
};

struct ConvArgs {
  const void* in;  // packed NHWC (fp32 for SIMT, bf16 for TC)
  int Hin, Win, Cin;
  const float* w;  // SIMT: [ks*ks][Cin][CoutP]
  const float* bias;  // [CoutP]
  const float* temb;  // time-bias table [rows][temb_stride] or null
  int temb_stride, temb_off;
  const int* step;    // device step index (row of temb)
  const float* res;   // raw NHWC [n][Ho][Wo][Cout] or null
  int CoutP;          // padded Cout (multiple of 32)
  int Cout;           // real Cout
  int ks, stride, pad;
  TileGeom g;
  int nimg;
  int out_mode;
  float* out;         // OUT_RAW: NHWC [n][Ho][Wo][Cout]; OUT_NCHW: v[n][Cout][Ho][Wo]; OUT_EULER: x[n][Cout][Ho][Wo] += dt*val
  float* stats;       // [n][parts][Cout][2] or null
  float dt;
  int out_padded;     // OUT_RAW: write the padded-linear layout (see raw_offset)
  int out_fmt;        // OUT_RAW storage: 0 = fp32, 1 = bf16, 2 = fp16
};

// ops_simt.cu
void launch_pack(const PackArgs& a, cudaStream_t s);
void launch_pack_padded(const PackArgs& a, cudaStream_t s);
// coef[nimg][C][2] = {scale, shift}; coef_h (optional) = fp16 {scale / 2, shift / 2} per 8-channel octet (see ops_simt.cu)
void launch_gn_coef(const PackArgs& a, float* coef, void* coef_h, cudaStream_t s);
void launch_pack_nchw_padded(const float* x, int C, int H, int W, void* out, int fmt, int nimg, cudaStream_t s);
void launch_pack_nchw(const float* x, int C, int H, int W, int Cp, void* out, int out_bf16, int nimg, cudaStream_t s);
void launch_conv_simt(const ConvArgs& a, cudaStream_t s);
void launch_gap_fc(const float* in, int HW, int C, const float* w, const float* b, int Cout, float* out, int ld,
                   int nimg, cudaStream_t s);
struct MlpLayer {
  const float* w;   // [out][in]
  const float* b;
  const float* ln_g;  // null => no LayerNorm/SiLU (final layer)
  const float* ln_b;
  int in, out, act;
};
void launch_mlp(const float* in, int ld_in, const MlpLayer* layers, int nlayers, float* out, int ld_out, int nimg,
                cudaStream_t s);
void launch_mlp_bwd(const float* in, int ld_in, const MlpLayer* layers, int nlayers, int loss_type, float* scores, float* log_ratio,
                    float* dfeat, int ld_d, int nimg, cudaStream_t s);
void launch_gap_fc_bwd(const float* dfeat, int ld, const float* w, int Cout, int HW, int C, float gscale, float* dact, int nimg, cudaStream_t s);
// storage of the gradient tensors of one pack backward: dense fp32 NHWC, or the tensor-core kernel's 16-bit padded-linear raw layout
struct PackBwdIO {
  int dcur_fmt = 0, dcur_padded = 0, dcur_Cs = 0;  // incoming gradient (geometry of the pack's output): 0 = fp32, 2 = fp16; channel stride
  int draw_padded = 0, draw_Cs = 0;                // outgoing gradient (geometry of the raw activation): fp32 dense, or fp16 padded with stride Cs
  // training step (train.cu): per-(image, slice, channel) {sum dz * xhat, sum dz} already computed by norm_param_grad_kernel;
  // the GroupNorm group moments are formed from them, so the image is split over `yslices` CTAs with no second pass
  const float* chan_sums = nullptr;
  int nslices = 0;   // slices of chan_sums
  int yslices = 1;   // CTAs per image of the dense fp32 generic kernel
};
void launch_pack_bwd(const PackArgs& a, const PackBwdIO& io, const float* dcur, float* draw, cudaStream_t s);

void launch_ratio_head(const float* scores, int loss_type, float* log_ratio, float* ratios, int n, cudaStream_t s);
void launch_step_inc(int* step, cudaStream_t s);
void launch_euler(float* x, const float* v, float dt, size_t n, cudaStream_t s);

struct ProfRec {  // one timed op launch (CUDA events on the launching stream)
  const char* kind;
  const char* label = nullptr;  // optional per-op description (RGFM_PROFILE_VERBOSE)
  cudaEvent_t e0, e1;
  double flops, bytes;
  int reps = 1;  // the op was launched `reps` times back to back between the events (GPU-bound timing)
};
struct ProfScope {  // brackets a launch with events when profiling is on
  std::vector<ProfRec>* prof;
  cudaStream_t s;
  ProfRec r;
  ProfScope(std::vector<ProfRec>* p, cudaStream_t st, const char* kind, double flops, double bytes) : prof(p), s(st) {
    if (!prof) return;
    r = ProfRec{kind, nullptr, nullptr, nullptr, flops, bytes};
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, s);
  }
  ~ProfScope() {
    if (!prof) return;
    cudaEventRecord(r.e1, s);
    prof->push_back(r);
  }
};

// guidance.cu
struct StepScalars {  // per Euler step, computed on the host in double then rounded like the reference
  float t, sig2, inv_sigma, sigma, one_minus_gamma, gamma, dt, pad;
};
struct McfArgs {
  int B, N, Dx, Dy;
  const float *x, *y, *vx, *vy, *mcx, *mcy, *ratios;
  const StepScalars* sc;  // table
  const int* step;        // index into sc (null => 0)
  float *x_out, *y_out;   // may alias x, y
  float *gx, *gy, *w;     // optional
  float* logits;          // scratch [ksplit][B][N]
  float* wbuf;            // scratch [B][N]
  float* zbuf;            // scratch [B][4]: Z_bar, w_min, w_max; followed by the sparse-weight lists (see guidance.cu)
  int ksplit_x, ksplit_y;
  int g_step;             // write gx / gy only at this step (diagnostics); < 0 = always
};
void launch_euler_grad(float* x, const float* v, const float* g, const StepScalars* sc, const int* step, float dt, size_t n, cudaStream_t s);
size_t mcf_scratch_floats(int B, int N, int Dx, int Dy, int* ksx, int* ksy);
void launch_mcf(const McfArgs& a, cudaStream_t s, std::vector<ProfRec>* prof = nullptr);
// the two halves of launch_mcf: distances + stabilised weights depend only on the state (x_t, y_t) and the bank, so the
// sampler's graph runs them on a third branch beside the two U-Nets; the fused guidance sum + blend + Euler update needs v
void launch_mcf_prep(const McfArgs& a, cudaStream_t s, std::vector<ProfRec>* prof = nullptr);
void launch_mcf_apply(const McfArgs& a, cudaStream_t s, std::vector<ProfRec>* prof = nullptr);
struct DiagArgs {
  int B, Dx, Dy, N;
  const float *vx, *vy, *gx, *gy, *zbuf;
  const StepScalars* sc;
  const int* step;
  int diag_step;
  float* diag;
};
void launch_diag(const DiagArgs& a, cudaStream_t s);
void launch_sat_report(const unsigned int* a, const unsigned int* b, const unsigned int* c, float* out, cudaStream_t s);
// count += #{i : argmax la[i] == argmax lb[i]} over [n][10] logits (first maximum, like torch.argmax)
void launch_argmax_match(const float* la, const float* lb, int n, float* count, cudaStream_t s);
// the inverse of a 28x28 pair transform (0 identity, 1 rotate90, 2 rotate180, 3 rotate270, 4 invert, 5 flip_h, 6 flip_v)
void launch_inverse_transform(const float* in, float* out, int n, int H, int code, cudaStream_t s);
void launch_ratio_stats(const float* ratios, int n, float* diag3, cudaStream_t s);

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream drains; it must execute
// pdl_wait() before touching anything the predecessor wrote (and calls pdl_trigger() early so ITS successor can follow).
// Inside stream capture these launches become programmatic edges of the CUDA graph.  Off by default (no measured gain:
// the persistent kernels leave nothing for the dependent's prologue to overlap with); RGFM_PDL=1 enables it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
  if (e != cudaSuccess) throw Error(std::string("cudaLaunchKernelEx failed: ") + cudaGetErrorString(e));
}

extern thread_local int64_t g_launch_counter;  // incremented by every launch_* (graph build-time kernel count)
// CTA cap of the persistent tcgen05 conv launches issued by this thread (0 = one CTA per SM).  The sampler sets it while it
// captures two INDEPENDENT network chains as parallel graph branches, so that the two chains' kernels own disjoint sets of
// SMs and run side by side instead of alternating: the pipeline fill / drain of one chain's launch (~5..10 us per launch
// and SM) is then covered by the other chain's steady state (sampler.cu, profiles/notes_r02.md section 14).
extern thread_local int g_tc_cta_cap;

}  // namespace rgfm

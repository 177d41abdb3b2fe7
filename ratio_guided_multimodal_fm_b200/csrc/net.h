// Op programs for the networks on the hot path: three U-Net velocity fields and two ratio estimators
// (SURVEY.md section 8a), built once from a reference state_dict and run on chunks of images.
#pragma once
#include "common.h"
#include "rgfm.h"
#include <functional>

namespace rgfm {

// name -> host tensor view of the caller's state_dict
struct WeightMap {
  std::map<std::string, std::pair<const float*, int64_t>> m;
  explicit WeightMap(const rgfm_weights* w);
  const float* get(const std::string& name, int64_t numel) const;
  bool has(const std::string& name) const { return m.count(name) != 0; }
};

// Device memory for a plan: bump allocations that live as long as the plan, plus a free-list for activations.
class DevPool {
 public:
  ~DevPool();
  void* alloc(size_t bytes);     // reuse a released block of sufficient size or cudaMalloc
  void release(void* p);         // return a block to the free list (plan-build time liveness)
  void* persistent(size_t bytes);
  template <typename T>
  T* upload(const std::vector<T>& h) {
    T* d = (T*)persistent(h.size() * sizeof(T));
    RGFM_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
  }
  size_t total_bytes() const { return total_; }

 private:
  struct Block { void* p; size_t bytes; bool free; };
  std::vector<Block> blocks_;
  std::vector<void*> persist_;
  size_t total_ = 0;
};

struct RunCtx {
  std::vector<ProfRec>* prof = nullptr;  // when set, every op is bracketed by events
  cudaStream_t s;
  int nimg;            // images in this chunk
  const float* x;      // chunk input, NCHW fp32 (U-Net: x_t; ratio: x)
  const float* y;      // ratio: y
  float* out;          // U-Net: v (OUT_NCHW) or x to update (OUT_EULER); ratio: scores
  int out_mode;
  float dt;
  const int* step;     // device step index selecting the time-bias row (null = row 0)
  float* gx = nullptr; // ratio estimator in gradient mode: d log r / dx, dy of the chunk (NCHW fp32)
  float* gy = nullptr;
  int loss_type = 0;
};

struct TcConv;  // conv_tc.cu
struct TcDeleter { void operator()(TcConv* p) const; };
using TcPtr = std::unique_ptr<TcConv, TcDeleter>;

class Net {
 public:
  // want_grad (ratio estimators only): build the forward + input-gradient (dgrad-only) program; exact fp32
  Net(int kind, const rgfm_weights* w, int chunk, int precision, int device, bool want_grad = false);
  bool has_grad() const { return want_grad_; }
  // d log r(x, y) / d(x, y) per README.md:159-164; log_ratio [nimg] optional
  void grad_log_ratio(cudaStream_t s, int nimg, const float* x, const float* y, int loss_type, float* gx, float* gy,
                      float* log_ratio, std::vector<ProfRec>* prof = nullptr);
  ~Net();
  int kind() const { return kind_; }
  bool is_unet() const { return kind_ <= RGFM_NET_UNET_SVHN || kind_ == RGFM_NET_FLOW28; }  // velocity fields v(x, t)
  int in_ch() const { return in_ch_; }
  int img() const { return img_; }
  int chunk() const { return chunk_; }
  size_t dim() const { return (size_t)in_ch_ * img_ * img_; }
  // device counter: output rows of the tensor-core convs whose fp16 store hit +-65504 since the caller last cleared it
  unsigned int* sat_counter() const { return d_sat_; }

  // U-Net: precompute the per-step time-bias rows for times ts (fp32 values of step*dt).
  // reference: timestep_embedding + time_embed + every ResBlock.time_mlp (unet.py:228-229,88)
  void set_times(const std::vector<float>& ts);

  // Run on `nimg` images (looping over chunks).  U-Net: out_mode OUT_NCHW writes v, OUT_EULER does x += dt*v
  // in place on `out`.  Ratio: x, y -> scores[nimg].  Classifier: x -> logits[nimg][10].
  void forward(cudaStream_t s, int nimg, const float* x, const float* y, float* out, int out_mode, float dt,
               const int* step, std::vector<ProfRec>* prof = nullptr);

 private:
  friend struct Builder;
  int kind_, chunk_, precision_, device_;
  bool want_grad_ = false;
  float* d_lr_scratch_ = nullptr;
  unsigned int* d_sat_ = nullptr;
  float *g_gx_ = nullptr, *g_gy_ = nullptr;  // set for the duration of grad_log_ratio()
  int g_loss_ = 0;
  bool g_lr_null_ = false;
  int in_ch_ = 0, img_ = 0, in_ch_y_ = 0;
  int out_stride_ = 1;  // floats per image of a non-U-Net output (ratio score: 1, classifier logits: 10)
  DevPool pool_;
  struct Op {
    std::function<void(const RunCtx&)> fn;
    const char* kind;       // kernel family, e.g. "conv_simt"
    double flops_per_img;   // algorithmic FLOPs (2*MACs of the reference op)
    double bytes_per_img;   // algorithmic bytes: operands read once + result written once
    std::string label;      // human-readable shape, for verbose profiles
  };
  std::vector<Op> ops_;
  std::vector<TcPtr> tc_;
  // time embedding (host copies)
  int model_ch_ = 0, temb_dim_ = 0, temb_C_ = 0;
  std::vector<float> te0_w_, te0_b_, te2_w_, te2_b_;
  struct TimeMlp { std::vector<float> w, b; int cout, off; };
  std::vector<TimeMlp> tmlps_;
  std::vector<float> flow_tw_;  // RGFM_NET_FLOW28: time columns of decoder.fc1, rows in the kernel's (pixel, channel) order
  float* d_temb_ = nullptr;
  int temb_rows_ = 0;
};

// conv_tc.cu -- tcgen05 implicit-GEMM convolution (16-bit operands, fp32 accumulate in TMEM)
struct TcSegment {
  const void* in;        // 16-bit padded-linear operand buffer [slots][Cin]
  int Cin;               // operand channels (multiple of 64, zero padded)
  int Cin_real;
  int ks;                // 1 or 3
  const float* w_oihw;   // host weights, torch layout [Cout][w_cin_total][ks][ks]; nullptr = identity (folded residual)
  bool raw;              // the operand is a (zero-padded) raw activation tensor: rows indexed by the GEMM row itself
  int w_cin_total;       // input channels of the weight tensor (0 = Cin_real); this segment uses [w_c0, w_c0 + Cin_real)
  int w_c0;
  // fused GroupNorm + SiLU (raw segments only): the kernel's transform warps rewrite the raw tile in shared memory as
  // silu(x * scale + shift) with coef[image][coef_C][2], channels [coef_c0, coef_c0 + Cin_real) -- no packed operand tensor
  // coef: {scale, shift}[chunk][coef_C] fp32, immediately followed by the fp16 copy {scale / 2 x 8, shift / 2 x 8} per
  // (image, 8-channel octet) that the fp16 mode's half2 transform reads (written by launch_gn_coef)
  const float* coef;
  int coef_C, coef_c0;
};
struct TcConvDesc {
  int Hin, Win;          // operand image geometry (stride-1 output geometry)
  int Cout, CoutP;       // real / padded output channels
  int chunk;             // max images
  int stride;            // 1, or 2 (computed at stride 1, stored subsampled)
  int nseg;              // 1, or 2-3 when a 1x1 skip conv / identity residual is folded in as extra K segments
  TcSegment seg[3];
  void* out;             // raw padded-linear 16-bit output buffer (TMA store) or nullptr
  float* stats;          // GroupNorm partials of the output [unit][slots][CoutP][2] or nullptr
  unsigned int* sat;     // device counter of output rows that hit the fp16 saturation bound (nullptr = not counted)
  bool fp16;             // operand format: fp16 (true) or bf16 (false); fp32 accumulate either way
};
TcPtr make_tc_conv(const TcConvDesc& d, DevPool& pool);
void launch_conv_tc(TcConv& tc, const ConvArgs& a, cudaStream_t s);
bool tc_available();
size_t tc_operand_slots(int H, int W, int chunk);  // rows of a padded-linear operand tensor
size_t tc_raw_rows(int H, int W, int chunk);       // rows of a padded-linear raw activation tensor
size_t tc_stats_floats(int R, int chunk, int unit_rows, int CoutP);  // floats of a conv's GroupNorm partials buffer (R: rows per image)
// output rows per GroupNorm-partials unit of a conv (128; 127 / 126 when a padded-linear conv stacks taps: stride 1, raw output, Cout <= 64)
int tc_unit_rows(int cout, int W, int stride, bool final, int stages3, int stages);  // stages: 64-channel operand blocks (3x3 segments / all)
// GEMM rows per image of the conv's row space: (H + 1) * (W + 2) padded-linear, (H + 1) * Wt row-dense (conv_tc.cu, TcArgs::dense)
int tc_rows_per_image(int cout, int H, int W, bool final, int stages);

}  // namespace rgfm

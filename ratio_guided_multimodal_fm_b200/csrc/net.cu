// Network programs: weight repacking from the reference state_dict and the op sequence of each network.
#include "net.h"
#include <cmath>
#include <cstring>
#include <algorithm>
#include <cstdlib>

namespace rgfm {

// ------------------------------------------------------------------------------------------------
WeightMap::WeightMap(const rgfm_weights* w) {
  RGFM_CHECK(w != nullptr && w->tensors != nullptr && w->count > 0, "weights: empty state_dict");
  for (int i = 0; i < w->count; ++i) {
    const rgfm_tensor& t = w->tensors[i];
    RGFM_CHECK(t.name != nullptr, "weights: null tensor name");
    m[t.name] = {t.data, t.numel};
  }
}

const float* WeightMap::get(const std::string& name, int64_t numel) const {
  auto it = m.find(name);
  if (it == m.end()) throw Error("state_dict is missing key '" + name + "'");
  if (it->second.second != numel)
    throw Error("state_dict key '" + name + "' has " + std::to_string(it->second.second) + " elements, expected " +
                std::to_string(numel));
  RGFM_CHECK(it->second.first != nullptr, "weights: null data pointer");
  return it->second.first;
}

DevPool::~DevPool() {
  for (auto& b : blocks_) cudaFree(b.p);
  for (void* p : persist_) cudaFree(p);
}

void* DevPool::alloc(size_t bytes) {
  bytes = (bytes + 1023) / 1024 * 1024;
  int best = -1;
  for (size_t i = 0; i < blocks_.size(); ++i)
    if (blocks_[i].free && blocks_[i].bytes >= bytes && (best < 0 || blocks_[i].bytes < blocks_[best].bytes)) best = (int)i;
  if (best >= 0 && blocks_[best].bytes <= bytes * 2) {
    blocks_[best].free = false;
    return blocks_[best].p;
  }
  void* p = nullptr;
  RGFM_CUDA(cudaMalloc(&p, bytes));
  blocks_.push_back({p, bytes, false});
  total_ += bytes;
  return p;
}

void DevPool::release(void* p) {
  for (auto& b : blocks_)
    if (b.p == p) {
      b.free = true;
      return;
    }
  throw Error("DevPool::release: unknown block");
}

void* DevPool::persistent(size_t bytes) {
  void* p = nullptr;
  bytes = std::max<size_t>(bytes, 16);
  RGFM_CUDA(cudaMalloc(&p, bytes));
  RGFM_CUDA(cudaMemset(p, 0, bytes));
  persist_.push_back(p);
  total_ += bytes;
  return p;
}

// ------------------------------------------------------------------------------------------------
// loss scale of the 16-bit gradient chain (encoder(): dgrad on the tensor cores)
constexpr float GRAD_SCALE = 1024.f;

struct Builder {
  using Op = Net::Op;
  Net& net;
  const WeightMap& W;
  DevPool& pool;
  int chunk;
  bool tc_mode;  // the net runs in a 16-bit tensor-core precision
  bool fp16;     // 16-bit operand format: fp16 (true) or bf16
  int fmt16;     // storage code of the 16-bit format (1 = bf16, 2 = fp16)
  // GroupNorm + SiLU applied by the conv kernel's transform warps on the raw tiles in shared memory (no packed operand
  // tensor, no pack launch): 8.56 -> 7.78 ms per guided Euler step.  RGFM_FUSE_GN=0 selects the separate pack kernels.
  bool fuse_gn = getenv("RGFM_FUSE_GN") ? atoi(getenv("RGFM_FUSE_GN")) != 0 : true;

  Builder(Net& n, const WeightMap& w)
      : net(n), W(w), pool(n.pool_), chunk(n.chunk_),
        tc_mode(n.precision_ == RGFM_PREC_BF16 || n.precision_ == RGFM_PREC_FP16), fp16(n.precision_ == RGFM_PREC_FP16),
        fmt16(n.precision_ == RGFM_PREC_FP16 ? 2 : 1) {}

  // channel stride of a tensor-core tensor: 32-channel tensors keep 64-byte rows (the conv reads them through a 64-wide
  // TMA box whose upper half is out of bounds = zero), everything else is padded to a multiple of 64 (128-byte rows)
  static int tc_stride(int C) { return C == 32 ? 32 : (C + 63) / 64 * 64; }

  // ---- buffers ----
  // fp32 dense raw activation (SIMT path) with optional per-tile GroupNorm partials
  Act new_act(int H, int Wd, int C, int parts) {
    Act a;
    a.H = H; a.W = Wd; a.C = C; a.Cs = C;
    a.p = (float*)pool.alloc((size_t)chunk * H * Wd * C * sizeof(float));
    if (parts > 0) {
      a.parts = parts;
      a.stats = (float*)pool.alloc((size_t)chunk * parts * C * 2 * sizeof(float));
      a.st.p = a.stats; a.st.mode = ST_PARTS; a.st.C = C; a.st.parts = parts;
    }
    return a;
  }
  // 16-bit zero-padded padded-linear raw activation (tcgen05 path); (Hc, Wc) is the geometry of the producing conv's
  // GEMM row space (differs from (H, W) for stride-2 convs), which is what the GroupNorm partials are indexed by
  Act new_act_tc(int H, int Wd, int C, bool stats, int R, int unit_rows) {  // R: rows per image of the producing conv's row space
    Act a;
    a.H = H; a.W = Wd; a.C = C; a.Cs = tc_stride(C);
    a.padded = true; a.fmt = 2;  // raw activations are fp16 in both 16-bit modes
    a.p = (float*)pool.alloc(tc_raw_rows(H, Wd, chunk) * a.Cs * 2);
    if (stats) {
      a.stats = (float*)pool.alloc(tc_stats_floats(R, chunk, unit_rows, a.Cs) * sizeof(float));
      a.st.p = a.stats; a.st.mode = ST_UNITS; a.st.C = a.Cs; a.st.R = R; a.st.slots = stats_slots_for(a.st.R);
      a.st.unit_rows = unit_rows;  // the producing conv's unit (conv_tc.cu: tap stacking shortens it)
      a.st.magic_unit = (unsigned int)(((1ull << 32) + a.st.unit_rows - 1) / a.st.unit_rows);
    }
    return a;
  }
  void release(Act& a) {
    if (a.p) pool.release(a.p);
    if (a.stats) pool.release(a.stats);
    a.p = nullptr; a.stats = nullptr;
  }
  Packed new_packed(int H, int Wd, int C, bool bf16) {
    Packed p;
    p.H = H; p.W = Wd; p.bf16 = bf16;
    if (bf16) {  // tcgen05 operand: padded-linear layout, channels padded to a multiple of 64
      p.padded = true;
      p.C = tc_stride(C);
      p.p = pool.alloc(tc_operand_slots(H, Wd, chunk) * p.C * 2);
    } else {
      p.C = C;
      p.p = pool.alloc((size_t)chunk * H * Wd * C * 4);
    }
    return p;
  }
  void release(Packed& p) {
    if (p.p) pool.release(p.p);
    p.p = nullptr;
  }

  const float* vec(const std::string& name, int n, int nP = 0) {
    if (nP < n) nP = n;
    const float* h = W.get(name, n);
    std::vector<float> v(nP, 0.f);
    std::memcpy(v.data(), h, n * sizeof(float));
    return pool.upload(v);
  }
  // torch [Cout][Cin][ks][ks] -> [tap][CinP][CoutP]
  const float* conv_w_simt(const std::string& name, int cout, int cin, int ks, int cinP, int coutP) {
    const float* h = W.get(name, (int64_t)cout * cin * ks * ks);
    std::vector<float> v((size_t)ks * ks * cinP * coutP, 0.f);
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int t = 0; t < ks * ks; ++t) v[((size_t)t * cinP + c) * coutP + o] = h[((size_t)o * cin + c) * ks * ks + t];
    return pool.upload(v);
  }

  // ---- ops ----
  Packed pack_input(int C, int H, int Wd, int Cp, bool from_y) {
    Packed p = new_packed(H, Wd, Cp, false);
    void* out = p.p;
    net.ops_.push_back({[=](const RunCtx& c) {
      launch_pack_nchw(from_y ? c.y : c.x, C, H, Wd, Cp, out, 0, c.nimg, c.s);
    }, "pack_nchw", 0.0, (double)H * Wd * (C * 4.0 + Cp * 4.0)});
    return p;
  }
  // tensor-core input operand: NCHW fp32 image -> padded-linear 16-bit [x_hi | x_lo | 0...] (64 channels per slot)
  Packed pack_input_tc(int C, int H, int Wd) {
    Packed p;  // 16-channel (32-byte) rows
    p.H = H; p.W = Wd; p.bf16 = true; p.padded = true; p.C = 16;
    p.p = pool.alloc(tc_operand_slots(H, Wd, chunk) * 16 * 2);
    void* out = p.p;
    const int f = fmt16;
    net.ops_.push_back({[=](const RunCtx& c) { launch_pack_nchw_padded(c.x, C, H, Wd, out, f, c.nimg, c.s); },
                        "pack_nchw", 0.0, (double)H * Wd * C * 4.0 + (double)(H + 1) * (Wd + 2) * 32.0,
                        std::to_string(H) + "x" + std::to_string(Wd) + " C" + std::to_string(C) + " hi/lo"});
    return p;
  }

  Packed pack(const std::vector<Act>& srcs, int norm, const float* gamma, const float* beta, int act, int spatial,
              bool out_bf16, PackArgs* saved = nullptr) {
    PackArgs a{};
    a.nsrc = (int)srcs.size();
    RGFM_CHECK(a.nsrc >= 1 && a.nsrc <= 2, "pack: 1 or 2 sources");
    int C = 0;
    double in_bytes = 0.0;
    for (int i = 0; i < a.nsrc; ++i) {
      PackSrc ps{};
      ps.p = srcs[i].p; ps.st = srcs[i].st; ps.C = srcs[i].C; ps.Cs = srcs[i].Cs;
      ps.padded = srcs[i].padded ? 1 : 0; ps.fmt = srcs[i].fmt;
      if (norm == NORM_GN) RGFM_CHECK(ps.st.mode != ST_NONE, "pack: GroupNorm needs the producer's partial statistics");
      a.src[i] = ps;
      C += srcs[i].C;
      in_bytes += (double)srcs[i].H * srcs[i].W * srcs[i].C * (srcs[i].fmt ? 2.0 : 4.0);
    }
    if (a.nsrc == 1) a.src[1] = a.src[0];
    a.Hs = srcs[0].H; a.Ws = srcs[0].W;
    if (spatial == SP_D2S) {
      RGFM_CHECK(a.nsrc == 1 && C % 4 == 0 && !out_bf16 && srcs[0].fmt == 0 && !srcs[0].padded, "pack: depth-to-space is an fp32, single-source mode");
      C /= 4;
    }
    const bool x2 = spatial == SP_UP2 || spatial == SP_D2S;
    a.Ho = x2 ? a.Hs * 2 : (spatial == SP_POOL2 ? a.Hs / 2 : a.Hs);
    a.Wo = x2 ? a.Ws * 2 : (spatial == SP_POOL2 ? a.Ws / 2 : a.Ws);
    a.C = C; a.norm = norm; a.act = act; a.spatial = spatial;
    a.gamma = gamma; a.beta = beta; a.groups = 8;
    Packed p = new_packed(a.Ho, a.Wo, C, out_bf16);
    a.out = p.p; a.out_bf16 = out_bf16 ? fmt16 : 0;
    a.Cpad = p.C;
    const double bytes = in_bytes + (double)a.Ho * a.Wo * C * (out_bf16 ? 2.0 : 4.0);
    const bool padded = p.padded;
    if (saved) *saved = a;
    net.ops_.push_back({[a, padded](const RunCtx& c) {
      PackArgs b = a;
      b.nimg = c.nimg;
      if (padded) launch_pack_padded(b, c.s);
      else launch_pack(b, c.s);
    }, padded ? "pack_padded" : "pack", 0.0, bytes,
       std::to_string(a.Hs) + "x" + std::to_string(a.Ws) + " C" + std::to_string(C) + (norm == NORM_GN ? " gn" : "") +
           (spatial == SP_UP2 ? " up2" : (spatial == SP_POOL2 ? " pool2" : (spatial == SP_D2S ? " d2s" : "")))});
    return p;
  }

  // GroupNorm partials index at most three images per 128-row unit (>= 64 GEMM rows per image); a conv without statistics
  // (BatchNorm nets, gradient tensors) only needs a tile that holds its 3x3 halo
  static bool tc_geom_ok(int H, int Wd, bool stats = true) {
    static const bool small_off = getenv("RGFM_TC_SMALL") && atoi(getenv("RGFM_TC_SMALL")) == 0;  // A/B switch
    return (H + 1) * (Wd + 2) >= (stats || small_off ? 64 : 16) && Wd + 2 <= 39 && H >= 2;
  }

  // ---- tcgen05 convolution -------------------------------------------------------------------------------------
  // K segment of a tensor-core conv: a packed operand or a (zero-padded) raw activation, 3x3 or 1x1, with weights taken
  // from a torch [Cout][w_cin_total][ks][ks] tensor (channels [w_c0, w_c0 + Creal)) or the identity
  struct Seg {
    const void* p; int Cs; int Creal; bool raw; int ks; const float* w; int w_cin_total; int w_c0;
    const float* coef = nullptr; int coef_C = 0, coef_c0 = 0;  // fused GroupNorm + SiLU (TcSegment::coef)
  };
  // raw segment normalised on the fly: channels [c0, c0 + in.C) of a GroupNorm over coef_C channels
  static Seg seg_gn(const Act& in, int ks, const float* w, int cin_total, int c0, const float* coef) {
    Seg s{in.p, in.Cs, in.C, true, ks, w, cin_total, c0};
    s.coef = coef; s.coef_C = cin_total; s.coef_c0 = c0;
    return s;
  }
  // per-image GroupNorm {scale, shift} of the concatenation of srcs, from the producers' partial statistics
  float* gn_coef(const std::vector<Act>& srcs, const float* gamma, const float* beta) {
    PackArgs a{};
    a.nsrc = (int)srcs.size();
    RGFM_CHECK(a.nsrc >= 1 && a.nsrc <= 2, "gn_coef: 1 or 2 sources");
    int C = 0;
    for (int i = 0; i < a.nsrc; ++i) {
      PackSrc ps{};
      ps.p = srcs[i].p; ps.st = srcs[i].st; ps.C = srcs[i].C; ps.Cs = srcs[i].Cs;
      ps.padded = srcs[i].padded ? 1 : 0; ps.fmt = srcs[i].fmt;
      a.src[i] = ps;
      C += srcs[i].C;
    }
    if (a.nsrc == 1) a.src[1] = a.src[0];
    a.Hs = a.Ho = srcs[0].H; a.Ws = a.Wo = srcs[0].W;
    a.C = C; a.norm = NORM_GN; a.act = ACT_SILU; a.spatial = SP_ID;
    a.gamma = gamma; a.beta = beta; a.groups = 8;
    for (int i = 0; i < a.nsrc; ++i) RGFM_CHECK(a.src[i].st.mode != ST_NONE, "gn_coef: GroupNorm needs the producer's partial statistics");
    // {scale, shift} fp32 pairs, followed by the fp16 copy the half2 transform reads (TcSegment::coef documents the layout)
    RGFM_CHECK(C % 8 == 0, "gn_coef: channels must be a multiple of 8");
    float* coef = (float*)pool.alloc((size_t)chunk * C * 3 * sizeof(float));
    void* coef_h = coef + (size_t)chunk * C * 2;
    net.ops_.push_back({[a, coef, coef_h](const RunCtx& c) {
      PackArgs b = a;
      b.nimg = c.nimg;
      launch_gn_coef(b, coef, coef_h, c.s);
    }, "gn_coef", 0.0, (double)C * 12.0, std::to_string(a.Hs) + "x" + std::to_string(a.Ws) + " C" + std::to_string(C)});
    return coef;
  }
  static Seg seg_of(const Packed& in, int creal, int ks, const float* w) { return Seg{in.p, in.C, creal, false, ks, w, 0, 0}; }
  static Seg seg_of(const Act& in, int ks, const float* w, int w_cin_total = 0, int w_c0 = 0) {
    return Seg{in.p, in.Cs, in.C, true, ks, w, w_cin_total, w_c0};
  }

  // (H, W): geometry of the operands (the stride-1 output geometry).  Returns the raw output (none when `final`).
  // grad_out (final convs only): 1 / 2 = the output is the chunk's d log r / dx / dy (RunCtx::gx / gy), written as
  // grad_scale * value (OUT_NCHW_SCALED) -- the last dgrad conv of the 16-bit gradient chain
  Act conv_tc(int H, int Wd, const std::vector<Seg>& segs, const float* bias, int cout, int stride, int temb_off, bool stats,
              bool final, const std::string& label, int grad_out = 0, float grad_scale = 1.f, const char* kind = "conv_tc") {
    RGFM_CHECK(tc_geom_ok(H, Wd, stats), "conv_tc: image geometry is not supported by the tensor-core kernel");
    const int Ho = stride == 2 ? H / 2 : H, Wo = stride == 2 ? Wd / 2 : Wd;
    ConvArgs a{};
    a.Hin = H; a.Win = Wd;
    a.bias = bias;
    a.out_padded = 1;
    a.out_fmt = 2;
    a.temb_off = temb_off < 0 ? 0 : temb_off;
    a.Cout = cout; a.ks = segs[0].ks; a.stride = stride; a.pad = segs[0].ks / 2;
    a.out_mode = OUT_RAW;
    Act o;
    if (!final) {
      int st3 = 0, stall = 0;
      for (const Seg& S : segs) {
        stall += (S.Creal + 63) / 64;
        if (S.ks == 3) st3 += (S.Creal + 63) / 64;
      }
      o = new_act_tc(Ho, Wo, cout, stats, tc_rows_per_image(cout, H, Wd, false, stall), tc_unit_rows(cout, Wd, stride, false, st3, stall));
      a.out = o.p;
      a.CoutP = o.Cs;
    } else {
      a.CoutP = 16;
    }
    TcConvDesc d{};
    d.Hin = H; d.Win = Wd; d.Cout = cout; d.CoutP = a.CoutP; d.chunk = chunk; d.fp16 = fp16; d.stride = stride;
    d.nseg = (int)segs.size();
    double macs = 0.0, in_bytes = 0.0;
    for (int i = 0; i < d.nseg; ++i) {
      const Seg& S = segs[i];
      d.seg[i] = TcSegment{S.p, S.Cs, S.Creal, S.ks, S.w, S.raw, S.w_cin_total, S.w_c0, S.coef, S.coef_C, S.coef_c0};
      if (S.w) macs += (double)S.Creal * S.ks * S.ks;  // the identity residual is not a reference FLOP
      in_bytes += (double)H * Wd * S.Creal * 2.0;
    }
    d.out = (final || stride != 1) ? nullptr : (void*)o.p;
    d.stats = final ? nullptr : o.stats;
    d.sat = net.d_sat_;
    net.tc_.push_back(make_tc_conv(d, pool));
    TcConv* tcp = net.tc_.back().get();
    const bool has_temb = temb_off >= 0;
    Net* np = &net;
    const double flops = 2.0 * Ho * Wo * (double)cout * macs;
    const double bytes = in_bytes + (double)Ho * Wo * cout * (final ? 4.0 : 2.0);
    net.ops_.push_back({[a, tcp, final, has_temb, np, grad_out, grad_scale](const RunCtx& c) {
      ConvArgs b = a;
      b.nimg = c.nimg;
      if (has_temb) {
        b.temb = np->d_temb_;
        b.temb_stride = np->temb_C_;
        b.step = c.step;
      }
      if (final) {
        b.out = c.out;
        b.out_mode = c.out_mode;
        b.dt = c.dt;
      }
      if (grad_out) {
        b.out = grad_out == 2 ? c.gy : c.gx;
        b.out_mode = OUT_NCHW_SCALED;
        b.dt = grad_scale;
      }
      launch_conv_tc(*tcp, b, c.s);
    }, kind, flops, bytes, label});
    return o;
  }

  // ---- fp32 SIMT convolution; returns the dense fp32 raw activation (or none when `final`) ----
  Act conv(const Packed& in, const std::string& wname, int cin_real, int cout, int ks, int stride, int temb_off,
           const float* res, bool stats, bool final) {
    const int coutP = (cout + 31) / 32 * 32;
    return conv_dev(in, conv_w_simt(wname + ".weight", cout, cin_real, ks, in.C, coutP), vec(wname + ".bias", cout, coutP), cout, ks,
                    stride, temb_off, res, stats, final, 2.0 * cout * cin_real * ks * ks,
                    wname + " " + std::to_string(in.H) + "x" + std::to_string(in.W) + " " + std::to_string(in.C) + "->" +
                        std::to_string(cout) + " k" + std::to_string(ks) + (stride == 2 ? " s2" : ""));
  }
  // the same with device weights already in the kernel's [tap][in.C][coutP] layout and a bias of coutP entries
  Act conv_dev(const Packed& in, const float* w, const float* bias, int cout, int ks, int stride, int temb_off, const float* res,
               bool stats, bool final, double flops_per_out_pixel, const std::string& label) {
    RGFM_CHECK(!in.bf16, "SIMT conv needs an fp32 operand");
    const int pad = ks / 2;
    const int Ho = (in.H + 2 * pad - ks) / stride + 1, Wo = (in.W + 2 * pad - ks) / stride + 1;
    const int stat_parts = make_geom(Ho, Wo).parts;
    const int coutP = (cout + 31) / 32 * 32;
    ConvArgs a{};
    a.in = in.p; a.Hin = in.H; a.Win = in.W; a.Cin = in.C;
    a.bias = bias;
    a.temb_off = temb_off < 0 ? 0 : temb_off;
    a.res = res;
    a.CoutP = coutP; a.Cout = cout; a.ks = ks; a.stride = stride; a.pad = pad;
    a.g = make_geom(Ho, Wo);
    a.out_mode = OUT_RAW;
    Act o;
    if (!final) {
      o = new_act(Ho, Wo, cout, stats ? stat_parts : 0);
      a.out = o.p;
      a.stats = o.stats;
    }
    a.w = w;
    const bool has_temb = temb_off >= 0;
    Net* np = &net;
    const double flops = flops_per_out_pixel * Ho * Wo;
    const double bytes = (double)in.H * in.W * in.C * 4.0 + (double)Ho * Wo * cout * 4.0 * (res ? 2.0 : 1.0);
    net.ops_.push_back({[a, final, has_temb, np](const RunCtx& c) {
      ConvArgs b = a;
      b.nimg = c.nimg;
      if (has_temb) {
        b.temb = np->d_temb_;
        b.temb_stride = np->temb_C_;
        b.step = c.step;
      }
      if (final) {
        b.out = c.out;
        b.out_mode = c.out_mode;
        b.dt = c.dt;
      }
      launch_conv_simt(b, c.s);
    }, "conv_simt", flops, bytes, label});
    return o;
  }

  static std::string shape_label(const std::string& w, int H, int Wd, int cin, int cout, const char* extra = "") {
    return w + " " + std::to_string(H) + "x" + std::to_string(Wd) + " " + std::to_string(cin) + "->" + std::to_string(cout) + extra;
  }

  // GN -> SiLU -> conv3x3 (+time bias) -> GN -> SiLU -> conv3x3 -> + skip.  reference: unet.py:74-96
  Act resblock(const std::vector<Act>& x, const std::string& p, int cout) {
    int cin = 0;
    for (auto& s : x) cin += s.C;
    const int H = x[0].H, Wd = x[0].W;
    // register this block's time MLP
    Net::TimeMlp tm;
    tm.cout = cout;
    tm.off = net.temb_C_;
    const float* tw = W.get(p + "time_mlp.1.weight", (int64_t)cout * net.temb_dim_);
    const float* tbias = W.get(p + "time_mlp.1.bias", cout);
    tm.w.assign(tw, tw + (size_t)cout * net.temb_dim_);
    tm.b.assign(tbias, tbias + cout);
    net.temb_C_ += cout;
    net.tmlps_.push_back(std::move(tm));
    const int toff = net.tmlps_.back().off;
    const bool has_skip = W.has(p + "skip.weight");
    if (!has_skip) RGFM_CHECK(x.size() == 1 && cin == cout, "resblock: identity skip needs matching channels");

    if (tc_mode) {
      // tensor-core path: 2 packs + 2 convs.  conv2 accumulates the block input as extra K segments read straight from
      // the raw 16-bit activations (the 1x1 skip conv, or identity weights for the plain residual): no separate skip
      // kernel, no residual read in the epilogue.
      const int coutP = (cout + 63) / 64 * 64;  // bias array length (>= nmma)
      if (fuse_gn) {
        // no packs at all: both norms become per-image coefficient kernels and the convs read the raw tensors
        float* cf1 = gn_coef(x, vec(p + "norm1.weight", cin), vec(p + "norm1.bias", cin));
        const float* w1 = W.get(p + "conv1.weight", (int64_t)cout * cin * 9);
        std::vector<Seg> s1;
        int c0 = 0;
        for (auto& xi : x) {
          s1.push_back(seg_gn(xi, 3, w1, cin, c0, cf1));
          c0 += xi.C;
        }
        Act h1 = conv_tc(H, Wd, s1, vec(p + "conv1.bias", cout, coutP), cout, 1, toff, true, false,
                         shape_label(p + "conv1", H, Wd, cin, cout, " gn"));
        pool.release(cf1);
        float* cf2 = gn_coef({h1}, vec(p + "norm2.weight", cout), vec(p + "norm2.bias", cout));
        std::vector<Seg> s2{seg_gn(h1, 3, W.get(p + "conv2.weight", (int64_t)cout * cout * 9), cout, 0, cf2)};
        const float* bias;
        if (has_skip) {
          RGFM_CHECK(x.size() <= 2, "resblock: at most two concatenated inputs");
          const float* sw = W.get(p + "skip.weight", (int64_t)cout * cin);
          c0 = 0;
          for (auto& xi : x) {
            s2.push_back(seg_of(xi, 1, sw, cin, c0));
            c0 += xi.C;
          }
          const float* b0 = W.get(p + "conv2.bias", cout);
          const float* b1 = W.get(p + "skip.bias", cout);
          std::vector<float> bs(coutP, 0.f);
          for (int i = 0; i < cout; ++i) bs[i] = b0[i] + b1[i];
          bias = pool.upload(bs);
        } else {
          s2.push_back(seg_of(x[0], 1, nullptr));
          bias = vec(p + "conv2.bias", cout, coutP);
        }
        Act out = conv_tc(H, Wd, s2, bias, cout, 1, -1, true, false,
                          shape_label(p + "conv2", H, Wd, cout, cout, has_skip ? " gn +skip" : " gn +id"));
        pool.release(cf2);
        release(h1);
        return out;
      }
      Packed a1 = pack(x, NORM_GN, vec(p + "norm1.weight", cin), vec(p + "norm1.bias", cin), ACT_SILU, SP_ID, true);
      Act h1 = conv_tc(H, Wd, {seg_of(a1, cin, 3, W.get(p + "conv1.weight", (int64_t)cout * cin * 9))},
                       vec(p + "conv1.bias", cout, coutP), cout, 1, toff, true, false, shape_label(p + "conv1", H, Wd, cin, cout));
      release(a1);
      Packed a2 = pack({h1}, NORM_GN, vec(p + "norm2.weight", cout), vec(p + "norm2.bias", cout), ACT_SILU, SP_ID, true);
      release(h1);
      std::vector<Seg> segs{seg_of(a2, cout, 3, W.get(p + "conv2.weight", (int64_t)cout * cout * 9))};
      const float* bias;
      if (has_skip) {
        const float* sw = W.get(p + "skip.weight", (int64_t)cout * cin);
        int c0 = 0;
        for (auto& xi : x) {
          segs.push_back(seg_of(xi, 1, sw, cin, c0));
          c0 += xi.C;
        }
        const float* b0 = W.get(p + "conv2.bias", cout);
        const float* b1 = W.get(p + "skip.bias", cout);
        std::vector<float> bs(coutP, 0.f);
        for (int i = 0; i < cout; ++i) bs[i] = b0[i] + b1[i];
        bias = pool.upload(bs);
      } else {
        segs.push_back(seg_of(x[0], 1, nullptr));
        bias = vec(p + "conv2.bias", cout, coutP);
      }
      Act out = conv_tc(H, Wd, segs, bias, cout, 1, -1, true, false,
                        shape_label(p + "conv2", H, Wd, cout, cout, has_skip ? " +skip" : " +id"));
      release(a2);
      return out;
    }
    Packed a1 = pack(x, NORM_GN, vec(p + "norm1.weight", cin), vec(p + "norm1.bias", cin), ACT_SILU, SP_ID, false);
    Act h1 = conv(a1, p + "conv1", cin, cout, 3, 1, toff, nullptr, true, false);
    release(a1);
    Packed a2 = pack({h1}, NORM_GN, vec(p + "norm2.weight", cout), vec(p + "norm2.bias", cout), ACT_SILU, SP_ID, false);
    release(h1);
    Act out;
    if (has_skip) {
      Packed ax;
      bool own = true;
      if (x.size() == 1) {  // a single fp32 raw tensor already is a valid operand
        ax.p = x[0].p; ax.H = x[0].H; ax.W = x[0].W; ax.C = x[0].C; ax.bf16 = false;
        own = false;
      } else {
        ax = pack(x, NORM_NONE, nullptr, nullptr, ACT_NONE, SP_ID, false);
      }
      Act sk = conv(ax, p + "skip", cin, cout, 1, 1, -1, nullptr, false, false);
      if (own) release(ax);
      out = conv(a2, p + "conv2", cout, cout, 3, 1, -1, sk.p, true, false);
      release(sk);
    } else {
      out = conv(a2, p + "conv2", cout, cout, 3, 1, -1, x[0].p, true, false);
    }
    release(a2);
    return out;
  }

  // ---- `--model original`: FlowMatchingModel (flow_matching.py:34-173), fp32 kernels ------------------------------
  // encoder: 4 x [conv3x3 (stride 1,2,2,1) -> GroupNorm(8) -> SiLU] -> flatten -> fc 12544->256
  // decoder: fc [features | time] -> 12544 -> view [256,7,7] -> 2 x [ConvTranspose2d(4,2,1) -> GN -> SiLU] -> conv3x3 -> GN
  //          -> SiLU -> conv3x3 -> v.  The Linear layers run as 1x1 convolutions over a 1x1 "image" (the NHWC activation
  //          [7][7][256] IS the 12544-vector once the weight rows are permuted); a transposed conv is a 3x3 convolution
  //          with 4*C_out outputs -- one block per output-pixel parity -- followed by the pack's depth-to-space.
  void build_flow28() {
    net.in_ch_ = 1; net.img_ = 28; net.model_ch_ = 0; net.temb_dim_ = 128;
    const int F = 256, S = 49, FS = F * S;  // feature channels, 7x7 pixels
    auto gn = [&](const Act& h, const std::string& p, int c, int spatial) {
      Act hh = h;
      Packed o = pack({hh}, NORM_GN, vec(p + ".weight", c), vec(p + ".bias", c), ACT_SILU, spatial, false);
      return o;
    };
    // encoder
    Packed cur = pack_input(1, 28, 28, 16, false);
    const int chans[5] = {1, 32, 64, 128, 256};
    const int strides[4] = {1, 2, 2, 1};
    for (int i = 0; i < 4; ++i) {
      const std::string n = std::to_string(i + 1);
      Act h = conv(cur, "encoder.conv" + n, chans[i], chans[i + 1], 3, strides[i], -1, nullptr, true, false);
      release(cur);
      cur = gn(h, "encoder.gn" + n, chans[i + 1], SP_ID);
      release(h);
    }
    // encoder.fc: y[o] = b[o] + sum_{c,hw} W[o][c*49 + hw] x[c][hw]  ->  operand index k = hw*256 + c
    Act feat;
    {
      const float* w = W.get("encoder.fc.weight", (int64_t)F * FS);
      std::vector<float> v((size_t)FS * F);
      for (int o = 0; o < F; ++o)
        for (int c = 0; c < F; ++c)
          for (int hw = 0; hw < S; ++hw) v[((size_t)hw * F + c) * F + o] = w[(size_t)o * FS + c * S + hw];
      Packed flat = cur;
      flat.H = 1; flat.W = 1; flat.C = FS;
      feat = conv_dev(flat, pool.upload(v), vec("encoder.fc.bias", F), F, 1, 1, -1, nullptr, false, false, 2.0 * F * FS,
                      "encoder.fc 12544->256");
      release(cur);
    }
    // decoder.fc1 over [features | time]: feature columns as a 1x1 conv, time columns as the per-step bias row; output
    // row o = c*49 + hw of the reference lands at hw*256 + c, i.e. directly as the NHWC [7][7][256] operand of deconv1
    Packed dec_in;
    {
      const int Kin = F + 128;
      const float* w = W.get("decoder.fc1.weight", (int64_t)FS * Kin);
      const float* b = W.get("decoder.fc1.bias", FS);
      std::vector<float> v((size_t)F * FS), bias(FS);
      net.flow_tw_.assign((size_t)FS * 128, 0.f);
      for (int c = 0; c < F; ++c)
        for (int hw = 0; hw < S; ++hw) {
          const int o = c * S + hw, op = hw * F + c;
          bias[op] = b[o];
          for (int k = 0; k < F; ++k) v[(size_t)k * FS + op] = w[(size_t)o * Kin + k];
          for (int j = 0; j < 128; ++j) net.flow_tw_[(size_t)op * 128 + j] = w[(size_t)o * Kin + F + j];
        }
      net.temb_C_ = FS;
      Packed fin;
      fin.p = feat.p; fin.H = 1; fin.W = 1; fin.C = F; fin.bf16 = false;
      Act h = conv_dev(fin, pool.upload(v), pool.upload(bias), FS, 1, 1, /*temb_off=*/0, nullptr, false, false, 2.0 * FS * Kin,
                       "decoder.fc1 384->12544");
      release(feat);
      dec_in.p = h.p; dec_in.H = 7; dec_in.W = 7; dec_in.C = F; dec_in.bf16 = false;
    }
    // ConvTranspose2d(cin, cout, 4, stride 2, padding 1): out[2i+a] takes in[i+dy] * W[ky] with (a, dy) -> ky:
    // (0,0)->1, (0,-1)->3, (1,+1)->0, (1,0)->2 (same along x); torch weight layout [cin][cout][4][4]
    auto deconv = [&](Packed& in, const std::string& name, int cin, int cout, const std::string& gname) {
      const float* w = W.get(name + ".weight", (int64_t)cin * cout * 16);
      const float* b = W.get(name + ".bias", cout);
      const int co4 = 4 * cout;
      std::vector<float> v((size_t)9 * cin * co4, 0.f), bias(co4);
      auto kof = [](int par, int d) { return par == 0 ? (d == 0 ? 1 : (d == -1 ? 3 : -1)) : (d == 1 ? 0 : (d == 0 ? 2 : -1)); };
      for (int a = 0; a < 2; ++a)
        for (int bb = 0; bb < 2; ++bb)
          for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
              const int ky = kof(a, dy), kx = kof(bb, dx);
              if (ky < 0 || kx < 0) continue;
              const int tap = (dy + 1) * 3 + (dx + 1);
              for (int ci = 0; ci < cin; ++ci)
                for (int co = 0; co < cout; ++co)
                  v[((size_t)tap * cin + ci) * co4 + (a * 2 + bb) * cout + co] = w[(((size_t)ci * cout + co) * 4 + ky) * 4 + kx];
            }
      for (int q = 0; q < 4; ++q)
        for (int co = 0; co < cout; ++co) bias[q * cout + co] = b[co];
      Act h = conv_dev(in, pool.upload(v), pool.upload(bias), co4, 3, 1, -1, nullptr, true, false, 2.0 * 16 * cin * cout,
                       name + " " + std::to_string(in.H) + "x" + std::to_string(in.W) + " " + std::to_string(cin) + "->" +
                           std::to_string(cout) + " k4 s2 transposed (3x3 conv, 4 parities)");
      release(in);
      Packed o = gn(h, gname, cout, SP_D2S);
      release(h);
      return o;
    };
    Packed d1 = deconv(dec_in, "decoder.deconv1", 256, 128, "decoder.gn1");
    Packed d2 = deconv(d1, "decoder.deconv2", 128, 64, "decoder.gn2");
    Act h3 = conv(d2, "decoder.conv3", 64, 32, 3, 1, -1, nullptr, true, false);
    release(d2);
    Packed p3 = gn(h3, "decoder.gn3", 32, SP_ID);
    release(h3);
    conv(p3, "decoder.conv_out", 32, 1, 3, 1, -1, nullptr, false, true);
    release(p3);
  }

  void build_unet(int in_ch, int img, int mc, const std::vector<int>& mult, int nres) {
    net.in_ch_ = in_ch; net.img_ = img; net.model_ch_ = mc; net.temb_dim_ = mc * 4;
    {
      const float* p;
      p = W.get("time_embed.0.weight", (int64_t)net.temb_dim_ * mc); net.te0_w_.assign(p, p + (size_t)net.temb_dim_ * mc);
      p = W.get("time_embed.0.bias", net.temb_dim_); net.te0_b_.assign(p, p + net.temb_dim_);
      p = W.get("time_embed.2.weight", (int64_t)net.temb_dim_ * net.temb_dim_); net.te2_w_.assign(p, p + (size_t)net.temb_dim_ * net.temb_dim_);
      p = W.get("time_embed.2.bias", net.temb_dim_); net.te2_b_.assign(p, p + net.temb_dim_);
    }
    Act h;
    if (tc_mode) {
      // input conv on the tensor cores over the [x_hi | x_lo] split of the fp32 state (weights duplicated over the two
      // halves): the state enters with ~2x the operand mantissa, and K = 2 * in_ch <= 8 real channels is ONE K16 step per tap
      Packed pin = pack_input_tc(in_ch, img, img);
      const float* w = W.get("input_conv.weight", (int64_t)mc * in_ch * 9);
      std::vector<float> w2((size_t)mc * 2 * in_ch * 9);
      for (int o = 0; o < mc; ++o)
        for (int c = 0; c < 2 * in_ch; ++c)
          for (int t = 0; t < 9; ++t) w2[((size_t)o * 2 * in_ch + c) * 9 + t] = w[((size_t)o * in_ch + (c % in_ch)) * 9 + t];
      h = conv_tc(img, img, {seg_of(pin, 2 * in_ch, 3, w2.data())}, vec("input_conv.bias", mc, (mc + 63) / 64 * 64), mc, 1, -1, true,
                  false, shape_label("input_conv", img, img, in_ch, mc, " hi/lo"));
      release(pin);
    } else {
      // exact fp32 on CUDA cores over a 16-channel zero-padded operand
      Packed pin = pack_input(in_ch, img, img, 16, false);
      h = conv(pin, "input_conv", in_ch, mc, 3, 1, -1, nullptr, true, false);
      release(pin);
    }
    std::vector<Act> hs{h};
    int bi = 0, di = 0;
    for (size_t level = 0; level < mult.size(); ++level) {
      for (int r = 0; r < nres; ++r) {
        h = resblock({h}, "encoder_blocks." + std::to_string(bi++) + ".", mc * mult[level]);
        hs.push_back(h);
      }
      if (level + 1 < mult.size()) {
        // Downsample: conv3x3 stride 2 on the raw activation (unet.py:99-107).  The zero-padded 16-bit raw tensor IS a
        // valid tensor-core operand, so no pack is needed; the fp32 path handles the stride in the gather.
        const std::string p = "downsamplers." + std::to_string(di++) + ".conv";
        Act dsrc = h;
        if (tc_mode) {
          h = conv_tc(dsrc.H, dsrc.W, {seg_of(dsrc, 3, W.get(p + ".weight", (int64_t)dsrc.C * dsrc.C * 9))},
                      vec(p + ".bias", dsrc.C, (dsrc.C + 63) / 64 * 64), dsrc.C, 2, -1, true, false,
                      shape_label(p, dsrc.H, dsrc.W, dsrc.C, dsrc.C, " s2"));
        } else {
          Packed op;
          op.p = dsrc.p; op.H = dsrc.H; op.W = dsrc.W; op.C = dsrc.C; op.bf16 = false;
          h = conv(op, p, dsrc.C, dsrc.C, 3, 2, -1, nullptr, true, false);
        }
        hs.push_back(h);
      }
    }
    Act m1 = resblock({h}, "middle_block1.", h.C);
    Act m2 = resblock({m1}, "middle_block2.", m1.C);
    release(m1);
    h = m2;
    bi = 0;
    int ui = 0;
    for (int level = (int)mult.size() - 1; level >= 0; --level) {
      for (int r = 0; r < nres + 1; ++r) {
        Act skip = hs.back();
        hs.pop_back();
        Act o = resblock({h, skip}, "decoder_blocks." + std::to_string(bi++) + ".", mc * mult[level]);
        release(h);
        release(skip);
        h = o;
      }
      if (level > 0) {
        // Upsample: nearest x2 folded into the operand pack, then conv3x3 (unet.py:110-119)
        const std::string p = "upsamplers." + std::to_string(ui++) + ".conv";
        Packed up = pack({h}, NORM_NONE, nullptr, nullptr, ACT_NONE, SP_UP2, tc_mode);
        Act o;
        if (tc_mode)
          o = conv_tc(up.H, up.W, {seg_of(up, h.C, 3, W.get(p + ".weight", (int64_t)h.C * h.C * 9))},
                      vec(p + ".bias", h.C, (h.C + 63) / 64 * 64), h.C, 1, -1, true, false, shape_label(p, up.H, up.W, h.C, h.C));
        else
          o = conv(up, p, h.C, h.C, 3, 1, -1, nullptr, true, false);
        release(up);
        release(h);
        h = o;
      }
    }
    const int out_cin = h.C;
    if (tc_mode && fuse_gn) {
      float* cf = gn_coef({h}, vec("out_norm.weight", h.C), vec("out_norm.bias", h.C));
      conv_tc(h.H, h.W, {seg_gn(h, 3, W.get("out_conv.weight", (int64_t)in_ch * out_cin * 9), out_cin, 0, cf)},
              vec("out_conv.bias", in_ch, 16), in_ch, 1, -1, false, true, shape_label("out_conv", h.H, h.W, out_cin, in_ch, " gn"));
      pool.release(cf);
      release(h);
      return;
    }
    Packed po = pack({h}, NORM_GN, vec("out_norm.weight", h.C), vec("out_norm.bias", h.C), ACT_SILU, SP_ID, tc_mode);
    release(h);
    if (tc_mode)
      conv_tc(po.H, po.W, {seg_of(po, out_cin, 3, W.get("out_conv.weight", (int64_t)in_ch * out_cin * 9))}, vec("out_conv.bias", in_ch, 16),
              in_ch, 1, -1, false, true, shape_label("out_conv", po.H, po.W, out_cin, in_ch));
    else
      conv(po, "out_conv", out_cin, in_ch, 3, 1, -1, nullptr, false, true);
    release(po);
  }

  // ---- ratio estimators ----
  struct EncLayer { std::string conv, norm; int cin, cout; bool pool; };

  void bn_affine(const std::string& p, int c, const float** scale, const float** shift) {
    const float* g = W.get(p + ".weight", c);
    const float* b = W.get(p + ".bias", c);
    const float* m = W.get(p + ".running_mean", c);
    const float* v = W.get(p + ".running_var", c);
    std::vector<float> sc(c), sh(c);
    for (int i = 0; i < c; ++i) {
      const float inv = 1.0f / std::sqrt(v[i] + 1e-5f);
      sc[i] = g[i] * inv;
      sh[i] = b[i] - m[i] * sc[i];
    }
    *scale = pool.upload(sc);
    *shift = pool.upload(sh);
  }

  // conv -> (GN | BN-eval) -> SiLU [-> maxpool2] ... -> GAP -> fc into feat[:, off:off+256]
  // `bwd` (gradient mode): every raw activation is kept, and the dgrad chain of this encoder is appended to *bwd:
  //   dfeat -> GAP/fc backward -> [pack backward (pool routing, SiLU', GN / BN backward) -> conv dgrad] x layers -> d input.
  // Conv dgrad is the same fp32 implicit-GEMM kernel run with the transposed, spatially flipped filter.
  void encoder(const std::string& pre, const std::vector<EncLayer>& L, bool gn, int in_ch, int img, bool from_y,
               float* feat, int feat_ld, int feat_off, int fc_in, std::vector<Net::Op>* bwd = nullptr, const float* dfeat = nullptr) {
    Packed cur = pack_input(in_ch, img, img, 16, from_y);
    std::vector<PackArgs> saved(L.size());
    // spatial size seen by each conv, and whether it runs on the tensor cores: layer 0 (K = 9/27) and the tiny
    // late layers (fewer than 64 GEMM rows per image) stay on the exact fp32 path
    std::vector<int> sp(L.size());
    std::vector<bool> ltc(L.size());
    int hcur = img;
    for (size_t i = 0; i < L.size(); ++i) {
      sp[i] = hcur;
      ltc[i] = tc_mode && i > 0 && tc_geom_ok(hcur, hcur, gn);
      if (L[i].pool) hcur /= 2;
    }
    for (size_t i = 0; i < L.size(); ++i) {
      const EncLayer& l = L[i];
      const bool last = (i + 1 == L.size());
      Act h;
      if (ltc[i]) {
        const int coutP = (l.cout + 63) / 64 * 64;
        h = conv_tc(cur.H, cur.W, {seg_of(cur, l.cin, 3, W.get(pre + l.conv + ".weight", (int64_t)l.cout * l.cin * 9))},
                    vec(pre + l.conv + ".bias", l.cout, coutP), l.cout, 1, -1, gn, false,
                    shape_label(pre + l.conv, cur.H, cur.W, l.cin, l.cout));
      } else {
        h = conv(cur, pre + l.conv, l.cin, l.cout, 3, 1, -1, nullptr, gn, false);
      }
      release(cur);
      const float *g, *b;
      if (gn) {
        g = vec(pre + l.norm + ".weight", l.cout);
        b = vec(pre + l.norm + ".bias", l.cout);
      } else {
        bn_affine(pre + l.norm, l.cout, &g, &b);
      }
      cur = pack({h}, gn ? NORM_GN : NORM_AFFINE, g, b, ACT_SILU, l.pool ? SP_POOL2 : SP_ID, last ? false : (bool)ltc[i + 1],
                 &saved[i]);
      if (!bwd) release(h);
    }
    const float* fw = vec(pre + "fc.weight", 256 * fc_in);
    const float* fb = vec(pre + "fc.bias", 256);
    const float* src = (const float*)cur.p;
    const int HW = cur.H * cur.W, C = cur.C;
    net.ops_.push_back({[=](const RunCtx& c) { launch_gap_fc(src, HW, C, fw, fb, 256, feat + feat_off, feat_ld, c.nimg, c.s); },
                        "gap_fc", 2.0 * 256 * C, (double)HW * C * 4.0});
    if (!bwd) {
      release(cur);
      return;
    }
    // ---- dgrad chain (appended after the score-net backward) ----
    // 16-bit modes: the dgrad convs of the layers the tensor-core kernel takes (every layer with >= 64 GEMM rows per image,
    // the first one included: its 1 / 3 outputs are an N = 16 output conv) run on conv_tc over fp16 gradient tensors in the
    // padded-linear raw layout.  Gradients of log r are ~1e-5, far below fp16's normal range, so the whole chain carries a
    // LOSS SCALE: gap_fc_bwd multiplies by GRAD_SCALE, every later op is linear in the gradient, and the last dgrad conv
    // stores value / GRAD_SCALE (OUT_NCHW_SCALED).  Saturated rows are counted like everywhere else (diag[12]).
    const float gscale = tc_mode ? GRAD_SCALE : 1.f;
    std::vector<Op>& B = *bwd;
    std::vector<bool> btc(L.size());
    for (size_t i = 0; i < L.size(); ++i) btc[i] = tc_mode && tc_geom_ok(sp[i], sp[i], false);
    float* dcur = (float*)pool.alloc((size_t)chunk * HW * C * sizeof(float));
    B.push_back({[=](const RunCtx& c) { launch_gap_fc_bwd(dfeat + feat_off, feat_ld, fw, 256, HW, C, gscale, dcur, c.nimg, c.s); }, "gap_fc_bwd",
                 2.0 * 256 * C, (double)HW * C * 4.0});
    PackBwdIO io{};  // format of the gradient arriving at layer i's pack output (dense fp32 out of the GAP / fc backward)
    io.dcur_fmt = 0; io.dcur_padded = 0; io.dcur_Cs = C;
    for (int i = (int)L.size() - 1; i >= 0; --i) {
      const EncLayer& l = L[i];
      const int H = saved[i].Hs;
      const PackArgs pa = saved[i];
      const float* dc = dcur;
      const bool first = (i == 0);
      // conv dgrad: dIn[h][w][ci] = sum_{tap, co} dRaw[h + ky - 1][w + kx - 1][co] * W[co][ci][2 - ky][2 - kx]
      const float* wt = W.get(pre + l.conv + ".weight", (int64_t)l.cout * l.cin * 9);
      const std::string lab = pre + l.conv + " dgrad " + std::to_string(H) + "x" + std::to_string(H) + " " + std::to_string(l.cout) + "->" + std::to_string(l.cin);
      if (btc[i]) {
        // outgoing gradient of the pack backward: fp16 padded-linear raw tensor = zero-padded operand of the dgrad conv
        Act dr = new_act_tc(H, H, l.cout, false, 0, 128);
        io.draw_padded = 1; io.draw_Cs = dr.Cs;
        float* draw = dr.p;
        const PackBwdIO iob = io;
        B.push_back({[=](const RunCtx& c) {
          PackArgs b = pa;
          b.nimg = c.nimg;
          launch_pack_bwd(b, iob, dc, draw, c.s);
        }, "pack_bwd", 0.0, (double)H * H * l.cout * 8.0});
        std::vector<float> wf((size_t)l.cin * l.cout * 9);  // torch layout [Cout' = cin][Cin' = cout][3][3], taps flipped
        for (int co = 0; co < l.cout; ++co)
          for (int ci = 0; ci < l.cin; ++ci)
            for (int t = 0; t < 9; ++t) wf[((size_t)ci * l.cout + co) * 9 + t] = wt[((size_t)co * l.cin + ci) * 9 + (8 - t)];
        const int coutP = first ? 16 : (l.cin + 63) / 64 * 64;
        const size_t nops = net.ops_.size();
        Act dn = conv_tc(H, H, {seg_of(dr, 3, wf.data())}, pool.upload(std::vector<float>(coutP, 0.f)), l.cin, 1, -1, false, first, lab,
                         first ? (from_y ? 2 : 1) : 0, 1.f / gscale, "conv_dgrad");
        // conv_tc() appended to the forward program: the dgrad chain runs after the score-net backward
        B.push_back(net.ops_.back());
        net.ops_.pop_back();
        RGFM_CHECK(net.ops_.size() == nops, "encoder: dgrad op bookkeeping");
        // NOTE: dr / dcur stay allocated (the backward program reuses them every call)
        io.dcur_fmt = 2; io.dcur_padded = 1; io.dcur_Cs = dn.Cs;
        dcur = dn.p;
      } else {
        float* draw = (float*)pool.alloc((size_t)chunk * H * H * l.cout * sizeof(float));
        io.draw_padded = 0; io.draw_Cs = l.cout;
        const PackBwdIO iob = io;
        B.push_back({[=](const RunCtx& c) {
          PackArgs b = pa;
          b.nimg = c.nimg;
          launch_pack_bwd(b, iob, dc, draw, c.s);
        }, "pack_bwd", 0.0, (double)H * H * l.cout * 12.0});
        const int cinP = (l.cin + 31) / 32 * 32;
        std::vector<float> wf((size_t)9 * l.cout * cinP, 0.f);
        for (int co = 0; co < l.cout; ++co)
          for (int ci = 0; ci < l.cin; ++ci)
            for (int t = 0; t < 9; ++t) wf[((size_t)t * l.cout + co) * cinP + ci] = wt[((size_t)co * l.cin + ci) * 9 + (8 - t)];
        ConvArgs ca{};
        ca.in = draw; ca.Hin = H; ca.Win = H; ca.Cin = l.cout;
        ca.w = pool.upload(wf);
        ca.bias = pool.upload(std::vector<float>(cinP, 0.f));
        ca.CoutP = cinP; ca.Cout = l.cin; ca.ks = 3; ca.stride = 1; ca.pad = 1;
        ca.g = make_geom(H, H);
        float* dnext = first ? nullptr : (float*)pool.alloc((size_t)chunk * H * H * l.cin * sizeof(float));
        ca.out_mode = first ? OUT_NCHW : OUT_RAW;
        ca.out = dnext;
        RGFM_CHECK(!first || gscale == 1.f, "encoder: the first dgrad conv of a 16-bit chain runs on the tensor cores");
        B.push_back({[=](const RunCtx& c) {
          ConvArgs b = ca;
          b.nimg = c.nimg;
          if (first) b.out = from_y ? c.gy : c.gx;
          launch_conv_simt(b, c.s);
        }, "conv_dgrad", 2.0 * H * H * (double)l.cout * l.cin * 9, (double)H * H * (l.cout + l.cin) * 4.0, lab});
        io.dcur_fmt = 0; io.dcur_padded = 0; io.dcur_Cs = l.cin;
        dcur = dnext;
      }
    }
  }

  void score_net(const float* feat, const std::vector<std::pair<int, int>>& lin_ln, int last, const std::vector<int>& dims,
                 float* dfeat = nullptr) {
    std::vector<MlpLayer> layers;
    int in = 512;
    for (size_t i = 0; i < lin_ln.size(); ++i) {
      MlpLayer l{};
      const std::string a = "score_net." + std::to_string(lin_ln[i].first), n = "score_net." + std::to_string(lin_ln[i].second);
      l.in = in; l.out = dims[i];
      l.w = vec(a + ".weight", l.in * l.out); l.b = vec(a + ".bias", l.out);
      l.ln_g = vec(n + ".weight", l.out); l.ln_b = vec(n + ".bias", l.out);
      l.act = ACT_SILU;
      layers.push_back(l);
      in = l.out;
    }
    MlpLayer f{};
    const std::string a = "score_net." + std::to_string(last);
    f.in = in; f.out = 1; f.w = vec(a + ".weight", in); f.b = vec(a + ".bias", 1); f.act = ACT_NONE;
    layers.push_back(f);
    double mf = 0;
    for (auto& l : layers) mf += 2.0 * l.in * l.out;
    if (dfeat) {  // forward + head + backward to the 512 features in one kernel
      net.ops_.push_back({[=](const RunCtx& c) {
        launch_mlp_bwd(feat, 512, layers.data(), (int)layers.size(), c.loss_type, nullptr, c.out, dfeat, 512, c.nimg, c.s);
      }, "mlp_bwd", 3.0 * mf, 512 * 8.0});
      return;
    }
    net.ops_.push_back({[=](const RunCtx& c) { launch_mlp(feat, 512, layers.data(), (int)layers.size(), c.out, 1, c.nimg, c.s); },
                        "mlp", mf, 512 * 4.0});
  }

  // Evaluators' classifiers (classifier.py:9-52, svhn_classifier.py:11-118): conv [-> BN(eval)] -> ReLU [-> maxpool2] ... ->
  // flatten (NCHW order) -> fc1 -> ReLU -> fc2.  Always exact fp32 (they decide an argmax; 8-44 MFLOP per sample).
  struct ClfLayer { std::string conv, bn; int cin, cout; bool pool; };
  void build_classifier(int in_ch, int img, const std::vector<ClfLayer>& L, int fc_hidden) {
    net.in_ch_ = in_ch; net.img_ = img; net.out_stride_ = 10;
    Packed cur = pack_input(in_ch, img, img, 16, false);
    for (const ClfLayer& l : L) {
      Act h = conv(cur, l.conv, l.cin, l.cout, 3, 1, -1, nullptr, false, false);
      release(cur);
      const float *g = nullptr, *b = nullptr;
      if (!l.bn.empty()) bn_affine(l.bn, l.cout, &g, &b);
      cur = pack({h}, l.bn.empty() ? NORM_NONE : NORM_AFFINE, g, b, ACT_RELU, l.pool ? SP_POOL2 : SP_ID, false);
      release(h);
    }
    // fc1 consumes x.view(B, -1) of an NCHW tensor: permute its columns to our NHWC order once
    const int HW = cur.H * cur.W, C = cur.C, in = HW * C;
    const float* w1 = W.get("fc1.weight", (int64_t)fc_hidden * in);
    std::vector<float> w1p((size_t)fc_hidden * in);
    for (int j = 0; j < fc_hidden; ++j)
      for (int c = 0; c < C; ++c)
        for (int q = 0; q < HW; ++q) w1p[(size_t)j * in + (size_t)q * C + c] = w1[(size_t)j * in + (size_t)c * HW + q];
    std::vector<MlpLayer> layers(2);
    layers[0] = MlpLayer{pool.upload(w1p), vec("fc1.bias", fc_hidden), nullptr, nullptr, in, fc_hidden, ACT_RELU};
    layers[1] = MlpLayer{vec("fc2.weight", 10 * fc_hidden), vec("fc2.bias", 10), nullptr, nullptr, fc_hidden, 10, ACT_NONE};
    const float* src = (const float*)cur.p;
    net.ops_.push_back({[=](const RunCtx& c) { launch_mlp(src, in, layers.data(), 2, c.out, 10, c.nimg, c.s); }, "mlp",
                        2.0 * in * fc_hidden + 20.0 * fc_hidden, (double)in * 4.0});
    // `cur` stays allocated: it is the MLP's input
  }

  void build_ratio28(bool grad) {  // reference: ratio_estimator.py:34-158
    net.in_ch_ = 1; net.in_ch_y_ = 1; net.img_ = 28;
    float* feat = (float*)pool.persistent((size_t)chunk * 512 * sizeof(float));
    float* dfeat = grad ? (float*)pool.persistent((size_t)chunk * 512 * sizeof(float)) : nullptr;
    std::vector<Net::Op> bwd;
    std::vector<EncLayer> L{{"conv1", "gn1", 1, 32, true}, {"conv2", "gn2", 32, 64, true}, {"conv3", "gn3", 64, 128, true}, {"conv4", "gn4", 128, 128, false}};
    encoder("encoder_x.", L, true, 1, 28, false, feat, 512, 0, 128, grad ? &bwd : nullptr, dfeat);
    encoder("encoder_y.", L, true, 1, 28, true, feat, 512, 256, 128, grad ? &bwd : nullptr, dfeat);
    score_net(feat, {{0, 1}, {4, 5}}, 8, {512, 256}, dfeat);
    for (auto& op : bwd) net.ops_.push_back(op);
  }

  void build_ratio_mnist_svhn(bool grad) {  // reference: ratio_flexible.py:185-364
    net.in_ch_ = 1; net.in_ch_y_ = 3; net.img_ = 32;
    float* feat = (float*)pool.persistent((size_t)chunk * 512 * sizeof(float));
    float* dfeat = grad ? (float*)pool.persistent((size_t)chunk * 512 * sizeof(float)) : nullptr;
    std::vector<Net::Op> bwd;
    std::vector<EncLayer> M{{"conv1", "bn1", 1, 32, true}, {"conv2", "bn2", 32, 64, true}, {"conv3", "bn3", 64, 128, true}, {"conv4", "bn4", 128, 128, false}};
    encoder("encoder_mnist.", M, false, 1, 32, false, feat, 512, 0, 128, grad ? &bwd : nullptr, dfeat);
    std::vector<EncLayer> S{{"conv1a", "bn1a", 3, 64, false},   {"conv1b", "bn1b", 64, 64, true},   {"conv2a", "bn2a", 64, 128, false},
                            {"conv2b", "bn2b", 128, 128, true}, {"conv3a", "bn3a", 128, 256, false}, {"conv3b", "bn3b", 256, 256, true},
                            {"conv4a", "bn4a", 256, 256, false}, {"conv4b", "bn4b", 256, 256, true}};
    encoder("encoder_svhn.", S, false, 3, 32, true, feat, 512, 256, 256, grad ? &bwd : nullptr, dfeat);
    score_net(feat, {{0, 1}, {4, 5}, {8, 9}}, 11, {512, 512, 256}, dfeat);
    for (auto& op : bwd) net.ops_.push_back(op);
  }
};

// ------------------------------------------------------------------------------------------------
Net::Net(int kind, const rgfm_weights* w, int chunk, int precision, int device, bool want_grad)
    : kind_(kind), chunk_(chunk), precision_(precision), device_(device), want_grad_(want_grad) {
  RGFM_CHECK(!want_grad || kind == RGFM_NET_RATIO28 || kind == RGFM_NET_RATIO_MNIST_SVHN, "net: only ratio estimators have a gradient program");
  RGFM_CHECK(chunk >= 1, "net: chunk");
  RGFM_CHECK(precision == RGFM_PREC_FP32 || precision == RGFM_PREC_BF16 || precision == RGFM_PREC_FP16,
             "net: precision must be fp32, bf16 or fp16");
  RGFM_CUDA(cudaSetDevice(device));
  if (precision != RGFM_PREC_FP32) RGFM_CHECK(tc_available(), "16-bit tensor-core precision needs an sm_100 device");
  WeightMap W(w);
  d_sat_ = (unsigned int*)pool_.persistent(sizeof(unsigned int));
  Builder b(*this, W);
  switch (kind) {
    case RGFM_NET_UNET_MNIST28: b.build_unet(1, 28, 32, {1, 2}, 2); break;
    case RGFM_NET_UNET_MNIST32: b.build_unet(1, 32, 32, {1, 2}, 2); break;
    case RGFM_NET_UNET_SVHN: b.build_unet(3, 32, 64, {1, 2, 2}, 2); break;
    case RGFM_NET_FLOW28:
      b.tc_mode = false;  // fp32 kernels in every precision mode (compatibility path, not the headline)
      b.build_flow28();
      break;
    case RGFM_NET_RATIO28:
      b.build_ratio28(want_grad);
      break;
    case RGFM_NET_RATIO_MNIST_SVHN:
      b.build_ratio_mnist_svhn(want_grad);
      break;
    case RGFM_NET_CLF_MNIST28:
      b.tc_mode = false;
      b.build_classifier(1, 28, {{"conv1", "", 1, 32, true}, {"conv2", "", 32, 64, true}}, 128);
      break;
    case RGFM_NET_CLF_MNIST32:
      b.tc_mode = false;
      b.build_classifier(1, 32, {{"conv1", "", 1, 32, true}, {"conv2", "", 32, 64, true}, {"conv3", "", 64, 64, false}}, 128);
      break;
    case RGFM_NET_CLF_SVHN:
      b.tc_mode = false;
      b.build_classifier(3, 32, {{"conv1", "bn1", 3, 32, true}, {"conv2", "bn2", 32, 64, true}, {"conv3", "bn3", 64, 128, false},
                                 {"conv4", "bn4", 128, 128, false}}, 256);
      break;
    default: throw Error("net kind " + std::to_string(kind) + " is not implemented");
  }
  if (is_unet()) set_times({0.0f});
  if (want_grad_) d_lr_scratch_ = (float*)pool_.persistent((size_t)chunk_ * sizeof(float));  // never allocate inside a capture
}

Net::~Net() {
  cudaSetDevice(device_);
  tc_.clear();
  if (d_temb_) cudaFree(d_temb_);
}

static inline float silu_h(float x) { return x / (1.0f + std::exp(-x)); }

void Net::set_times(const std::vector<float>& ts) {
  RGFM_CHECK(is_unet(), "set_times: not a U-Net");
  const int half = model_ch_ / 2;
  std::vector<float> table((size_t)ts.size() * temb_C_);
  if (kind_ == RGFM_NET_FLOW28) {
    // SinusoidalPositionEmbeddings(128) (flow_matching.py:18-31): f_k = exp(k * -(ln 1e4 / 63)), emb = [sin(t f), cos(t f)],
    // t unscaled; its contribution to decoder.fc1 (the last 128 input columns, flow_matching.py:115-118) is a per-step bias
    const int hd = 64;
    const float step_f = -(float)(std::log(10000.0) / (hd - 1));
    std::vector<float> e(2 * hd);
    for (size_t r = 0; r < ts.size(); ++r) {
      for (int k = 0; k < hd; ++k) {
        const float arg = ts[r] * std::exp((float)k * step_f);
        e[k] = std::sin(arg);
        e[hd + k] = std::cos(arg);
      }
      for (int o = 0; o < temb_C_; ++o) {
        double s = 0.0;
        for (int j = 0; j < 2 * hd; ++j) s += (double)flow_tw_[(size_t)o * 2 * hd + j] * e[j];
        table[r * temb_C_ + o] = (float)s;
      }
    }
  }
  std::vector<float> e(model_ch_), h1(temb_dim_), h2(temb_dim_);
  for (size_t r = 0; r < ts.size() && kind_ != RGFM_NET_FLOW28; ++r) {
    const float t = ts[r];
    for (int k = 0; k < half; ++k) {
      // freqs = exp(-ln(1e4) * k / half) in fp32; emb = [cos(t f), sin(t f)]  (unet.py:25-30)
      const float f = std::exp((-(float)std::log(10000.0) * (float)k) / (float)half);
      const float arg = t * f;
      e[k] = std::cos(arg);
      e[half + k] = std::sin(arg);
    }
    for (int j = 0; j < temb_dim_; ++j) {
      double s = te0_b_[j];
      for (int i = 0; i < model_ch_; ++i) s += (double)te0_w_[(size_t)j * model_ch_ + i] * e[i];
      h1[j] = silu_h((float)s);
    }
    for (int j = 0; j < temb_dim_; ++j) {
      double s = te2_b_[j];
      for (int i = 0; i < temb_dim_; ++i) s += (double)te2_w_[(size_t)j * temb_dim_ + i] * h1[i];
      h2[j] = silu_h((float)s);  // every ResBlock applies SiLU before its Linear (unet.py:58-61)
    }
    for (const TimeMlp& m : tmlps_)
      for (int j = 0; j < m.cout; ++j) {
        double s = m.b[j];
        for (int i = 0; i < temb_dim_; ++i) s += (double)m.w[(size_t)j * temb_dim_ + i] * h2[i];
        table[r * temb_C_ + m.off + j] = (float)s;
      }
  }
  RGFM_CUDA(cudaSetDevice(device_));
  if ((int)ts.size() > temb_rows_) {
    if (d_temb_) RGFM_CUDA(cudaFree(d_temb_));
    RGFM_CUDA(cudaMalloc(&d_temb_, table.size() * sizeof(float)));
    temb_rows_ = (int)ts.size();
  }
  RGFM_CUDA(cudaMemcpy(d_temb_, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice));
}

void Net::grad_log_ratio(cudaStream_t s, int nimg, const float* x, const float* y, int loss_type, float* gx, float* gy,
                         float* log_ratio, std::vector<ProfRec>* prof) {
  RGFM_CHECK(want_grad_, "net: built without the gradient program");
  RGFM_CHECK(loss_type == RGFM_LOSS_DISC || loss_type == RGFM_LOSS_RULSIF, "Unknown loss_type");
  g_gx_ = gx; g_gy_ = gy; g_loss_ = loss_type; g_lr_null_ = (log_ratio == nullptr);
  forward(s, nimg, x, y, log_ratio ? log_ratio : d_lr_scratch_, 0, 0.f, nullptr, prof);
  g_gx_ = g_gy_ = nullptr;
}

void Net::forward(cudaStream_t s, int nimg, const float* x, const float* y, float* out, int out_mode, float dt,
                  const int* step, std::vector<ProfRec>* prof) {
  const size_t dx = (size_t)in_ch_ * img_ * img_;
  const size_t dy = (size_t)in_ch_y_ * img_ * img_;
  RGFM_CHECK(!want_grad_ || g_gx_ != nullptr, "net: a gradient-mode estimator runs through grad_log_ratio()");
  for (int c0 = 0; c0 < nimg; c0 += chunk_) {
    RunCtx c;
    c.gx = g_gx_ ? g_gx_ + c0 * dx : nullptr;
    c.gy = g_gy_ ? g_gy_ + c0 * dy : nullptr;
    c.loss_type = g_loss_;
    c.s = s;
    c.nimg = std::min(chunk_, nimg - c0);
    c.x = x + c0 * dx;
    c.y = y ? y + c0 * dy : nullptr;
    c.out = is_unet() ? out + c0 * dx : (g_lr_null_ ? out : out + (size_t)c0 * out_stride_);
    c.out_mode = out_mode;
    c.dt = dt;
    c.step = step;
    c.prof = prof;
    for (auto& op : ops_) {
      if (prof) {
        ProfRec r{op.kind, op.label.c_str(), nullptr, nullptr, op.flops_per_img * c.nimg, op.bytes_per_img * c.nimg};
        RGFM_CUDA(cudaEventCreate(&r.e0));
        RGFM_CUDA(cudaEventCreate(&r.e1));
        static const int reps = getenv("RGFM_PROFILE_REPS") ? std::max(1, atoi(getenv("RGFM_PROFILE_REPS"))) : 3;
        r.reps = reps;
        RGFM_CUDA(cudaEventRecord(r.e0, s));
        for (int k = 0; k < reps; ++k) op.fn(c);  // back-to-back launches keep the GPU ahead of the CPU
        RGFM_CUDA(cudaEventRecord(r.e1, s));
        prof->push_back(r);
      } else {
        op.fn(c);
      }
    }
  }
}

}  // namespace rgfm

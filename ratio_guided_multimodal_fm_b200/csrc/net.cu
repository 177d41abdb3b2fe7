// Network programs: weight repacking from the reference state_dict and the op sequence of each network.
#include "net.h"
#include <cmath>
#include <cstring>
#include <algorithm>

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
struct Builder {
  Net& net;
  const WeightMap& W;
  DevPool& pool;
  int chunk;
  bool tc;       // route the NEXT conv to the tensor cores (toggled off around the K=9/27 input convs)
  bool tc_mode;  // the net runs in a 16-bit tensor-core precision
  bool fp16;     // 16-bit operand format: fp16 (true) or bf16

  Builder(Net& n, const WeightMap& w) : net(n), W(w), pool(n.pool_), chunk(n.chunk_), tc(n.precision_ == RGFM_PREC_BF16 || n.precision_ == RGFM_PREC_FP16), tc_mode(tc), fp16(n.precision_ == RGFM_PREC_FP16) {}

  Act new_act(int H, int Wd, int C, int parts) {
    Act a;
    a.H = H; a.W = Wd; a.C = C;
    a.padded = tc_mode;  // every raw activation of a tensor-core net lives in the padded-linear layout
    const size_t rows = a.padded ? tc_raw_rows(H, Wd, chunk) : (size_t)chunk * H * Wd;
    a.p = (float*)pool.alloc(rows * C * sizeof(float));
    if (parts > 0) {
      a.parts = parts;
      a.stats = (float*)pool.alloc((size_t)chunk * parts * C * 2 * sizeof(float));
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
      p.C = (C + 63) / 64 * 64;
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

  Packed pack(const std::vector<Act>& srcs, int norm, const float* gamma, const float* beta, int act, int spatial,
              bool out_bf16, Packed* raw16 = nullptr) {
    PackArgs a{};
    a.nsrc = (int)srcs.size();
    RGFM_CHECK(a.nsrc >= 1 && a.nsrc <= 2, "pack: 1 or 2 sources");
    int C = 0;
    std::vector<float*> tmp_stats;
    for (int i = 0; i < a.nsrc; ++i) {
      a.src[i] = PackSrc{srcs[i].p, srcs[i].stats, srcs[i].C, srcs[i].parts, srcs[i].padded ? 1 : 0};
      if (norm == NORM_GN && !out_bf16 && srcs[i].stats == nullptr) {
        // fp32-operand pack after a producer without GroupNorm partials (tcgen05 conv): compute them here
        float* st = (float*)pool.alloc((size_t)chunk * srcs[i].C * 2 * sizeof(float));
        tmp_stats.push_back(st);
        const float* src = srcs[i].p;
        const int Hh = srcs[i].H, Ww = srcs[i].W, Cs = srcs[i].C, pd = srcs[i].padded ? 1 : 0;
        net.ops_.push_back({[=](const RunCtx& c) { launch_gn_stats(src, Hh, Ww, Cs, pd, st, c.nimg, c.s); }, "gn_stats", 0.0,
                            (double)Hh * Ww * Cs * 4.0});
        a.src[i].stats = st;
        a.src[i].parts = 1;
      }
      C += srcs[i].C;
    }
    if (a.nsrc == 1) a.src[1] = a.src[0];
    a.Hs = srcs[0].H; a.Ws = srcs[0].W;
    a.Ho = spatial == SP_UP2 ? a.Hs * 2 : (spatial == SP_POOL2 ? a.Hs / 2 : a.Hs);
    a.Wo = spatial == SP_UP2 ? a.Ws * 2 : (spatial == SP_POOL2 ? a.Ws / 2 : a.Ws);
    a.C = C; a.norm = norm; a.act = act; a.spatial = spatial;
    a.gamma = gamma; a.beta = beta; a.groups = 8;
    Packed p = new_packed(a.Ho, a.Wo, C, out_bf16);
    a.out = p.p; a.out_bf16 = out_bf16 ? (fp16 ? 2 : 1) : 0;
    a.Cpad = p.C;
    a.out2 = nullptr;
    if (raw16) {
      RGFM_CHECK(out_bf16 && spatial == SP_ID, "pack: the raw 16-bit copy needs a padded 16-bit pack without resampling");
      *raw16 = new_packed(a.Ho, a.Wo, C, true);
      a.out2 = raw16->p;
    }
    const double bytes = (double)a.Hs * a.Ws * C * 4.0 + (double)a.Ho * a.Wo * C * (out_bf16 ? 2.0 : 4.0) * (raw16 ? 2.0 : 1.0);
    const bool padded = p.padded;
    net.ops_.push_back({[a, padded](const RunCtx& c) {
      PackArgs b = a;
      b.nimg = c.nimg;
      if (padded) launch_pack_padded(b, c.s);
      else launch_pack(b, c.s);
    }, padded ? "pack_padded" : "pack", 0.0, bytes,
       std::to_string(a.Hs) + "x" + std::to_string(a.Ws) + " C" + std::to_string(C) + (raw16 ? " dual" : "") + (norm == NORM_GN ? " gn" : "")});
    for (float* st : tmp_stats) pool.release(st);
    return p;
  }

  static bool tc_stats_ok(int H, int Wd) { return (H + 1) * (Wd + 2) >= 64; }

  // generic conv; returns raw activation (or none when `final`)
  // `skip` (tensor-core path only): a second K segment accumulated into the same tile -- the ResBlock's 1x1 skip conv
  // over the raw 16-bit copy of the block input (weights `skip_w`), or the identity residual when skip_w is empty.
  Act conv(const Packed& in, const std::string& wname, int cin_real, int cout, int ks, int stride, int temb_off,
           const float* res, bool stats, bool final, const Packed* skip = nullptr, const std::string& skip_w = "",
           int skip_cin_real = 0) {
    const int pad = ks / 2;
    const int Ho = (in.H + 2 * pad - ks) / stride + 1, Wo = (in.W + 2 * pad - ks) / stride + 1;
    const bool use_tc = tc && in.bf16;
    // the tcgen05 kernel walks the operand (stride-1) space and stores stride-2 outputs subsampled
    if (use_tc) stats = false;  // consumers compute GroupNorm statistics themselves (pack_padded / gn_stats)
    const int stat_parts = make_geom(Ho, Wo).parts;
    const int coutP = (use_tc && final) ? 16 : (cout + 31) / 32 * 32;
    ConvArgs a{};
    a.in = in.p; a.Hin = in.H; a.Win = in.W; a.Cin = in.C;
    if (skip && !skip_w.empty()) {  // conv bias + folded skip-conv bias
      const float* b0 = W.get(wname + ".bias", cout);
      const float* b1 = W.get(skip_w + ".bias", cout);
      std::vector<float> bs(coutP, 0.f);
      for (int i = 0; i < cout; ++i) bs[i] = b0[i] + b1[i];
      a.bias = pool.upload(bs);
    } else {
      a.bias = vec(wname + ".bias", cout, coutP);
    }
    a.out_padded = tc_mode ? 1 : 0;
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
    TcConv* tcp = nullptr;
    if (use_tc) {
      TcConvDesc d{};
      d.Hin = in.H; d.Win = in.W; d.Cout = cout; d.CoutP = coutP; d.chunk = chunk; d.fp16 = fp16;
      d.nseg = skip ? 2 : 1;
      d.seg[0] = TcSegment{in.p, in.C, cin_real, ks, W.get(wname + ".weight", (int64_t)cout * cin_real * ks * ks)};
      if (skip) {
        RGFM_CHECK(skip->padded && skip->H == in.H && skip->W == in.W, "conv: skip operand geometry");
        const float* sw = skip_w.empty() ? nullptr : W.get(skip_w + ".weight", (int64_t)cout * skip_cin_real);
        d.seg[1] = TcSegment{skip->p, skip->C, skip_cin_real, 1, sw};
      }
      d.out = (final || stride != 1) ? nullptr : o.p;
      net.tc_.push_back(make_tc_conv(d, pool));
      tcp = net.tc_.back().get();
    } else {
      RGFM_CHECK(!in.bf16 && !skip, "SIMT conv needs an fp32 operand and has no folded skip segment");
      a.w = conv_w_simt(wname + ".weight", cout, cin_real, ks, in.C, coutP);
    }
    const bool has_temb = temb_off >= 0;
    Net* np = &net;
    const double flops = 2.0 * Ho * Wo * (double)cout * (cin_real * ks * ks + (skip && !skip_w.empty() ? skip_cin_real : 0));
    const double bytes = (double)in.H * in.W * in.C * (in.bf16 ? 2.0 : 4.0) + (double)Ho * Wo * cout * 4.0 * (res ? 2.0 : 1.0);
    net.ops_.push_back({[a, tcp, final, has_temb, np](const RunCtx& c) {
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
      if (tcp)
        launch_conv_tc(*tcp, b, c.s);
      else
        launch_conv_simt(b, c.s);
    }, use_tc ? "conv_tc" : "conv_simt", flops, bytes,
       wname + " " + std::to_string(in.H) + "x" + std::to_string(in.W) + " " + std::to_string(in.C) + "->" + std::to_string(cout) +
           " k" + std::to_string(ks) + (stride == 2 ? " s2" : "") + (skip ? " +skip" + std::to_string(skip->C) : "")});
    return o;
  }

  // GN -> SiLU -> conv3x3 (+time bias) -> GN -> SiLU -> conv3x3 -> + skip.  reference: unet.py:74-96
  Act resblock(const std::vector<Act>& x, const std::string& p, int cout) {
    int cin = 0;
    for (auto& s : x) cin += s.C;
    const int H = x[0].H;
    (void)H;
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

    if (tc) {
      // tensor-core path: 2 packs + 2 convs.  The first pack also emits the raw 16-bit copy of x, which conv2 consumes
      // as a second K segment (1x1 skip conv, or identity weights for the plain residual) -- no separate skip
      // kernel, no residual read in the epilogue.
      Packed x16;
      Packed a1 = pack(x, NORM_GN, vec(p + "norm1.weight", cin), vec(p + "norm1.bias", cin), ACT_SILU, SP_ID, true, &x16);
      Act h1 = conv(a1, p + "conv1", cin, cout, 3, 1, toff, nullptr, false, false);
      release(a1);
      Packed a2 = pack({h1}, NORM_GN, vec(p + "norm2.weight", cout), vec(p + "norm2.bias", cout), ACT_SILU, SP_ID, true);
      release(h1);
      const bool has_skip = W.has(p + "skip.weight");
      if (!has_skip) RGFM_CHECK(cin == cout, "resblock: identity skip needs matching channels");
      Act out = conv(a2, p + "conv2", cout, cout, 3, 1, -1, nullptr, false, false, &x16, has_skip ? p + "skip" : std::string(), cin);
      release(a2);
      release(x16);
      return out;
    }
    Packed a1 = pack(x, NORM_GN, vec(p + "norm1.weight", cin), vec(p + "norm1.bias", cin), ACT_SILU, SP_ID, tc);
    Act h1 = conv(a1, p + "conv1", cin, cout, 3, 1, toff, nullptr, true, false);
    release(a1);
    Packed a2 = pack({h1}, NORM_GN, vec(p + "norm2.weight", cout), vec(p + "norm2.bias", cout), ACT_SILU, SP_ID, tc);
    release(h1);
    Act out;
    if (W.has(p + "skip.weight")) {
      Packed ax;
      bool own = true;
      if (!tc && x.size() == 1) {  // a single fp32 raw tensor already is a valid operand
        ax.p = x[0].p; ax.H = x[0].H; ax.W = x[0].W; ax.C = x[0].C; ax.bf16 = false;
        own = false;
      } else {
        ax = pack(x, NORM_NONE, nullptr, nullptr, ACT_NONE, SP_ID, tc);
      }
      Act sk = conv(ax, p + "skip", cin, cout, 1, 1, -1, nullptr, false, false);
      if (own) release(ax);
      out = conv(a2, p + "conv2", cout, cout, 3, 1, -1, sk.p, true, false);
      release(sk);
    } else {
      RGFM_CHECK(x.size() == 1 && cin == cout, "resblock: identity skip needs matching channels");
      out = conv(a2, p + "conv2", cout, cout, 3, 1, -1, x[0].p, true, false);
    }
    release(a2);
    return out;
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
    // input conv: tiny K (9 or 27) -> exact fp32 on CUDA cores over a 16-channel zero-padded operand
    Packed pin = pack_input(in_ch, img, img, 16, false);
    const bool tc_saved = tc;
    tc = false;
    Act h = conv(pin, "input_conv", in_ch, mc, 3, 1, -1, nullptr, !tc_saved, false);
    tc = tc_saved;
    release(pin);
    std::vector<Act> hs{h};
    int bi = 0, di = 0;
    for (size_t level = 0; level < mult.size(); ++level) {
      for (int r = 0; r < nres; ++r) {
        h = resblock({h}, "encoder_blocks." + std::to_string(bi++) + ".", mc * mult[level]);
        hs.push_back(h);
      }
      if (level + 1 < mult.size()) {
        // Downsample: conv3x3 stride 2 on the raw activation (unet.py:99-107); exact fp32 path handles the stride
        const std::string p = "downsamplers." + std::to_string(di++) + ".conv";
        Act dsrc = h;
        if (tc) {
          Packed op = pack({dsrc}, NORM_NONE, nullptr, nullptr, ACT_NONE, SP_ID, true);
          h = conv(op, p, dsrc.C, dsrc.C, 3, 2, -1, nullptr, true, false);
          release(op);
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
        Packed up = pack({h}, NORM_NONE, nullptr, nullptr, ACT_NONE, SP_UP2, tc);
        Act o = conv(up, "upsamplers." + std::to_string(ui++) + ".conv", h.C, h.C, 3, 1, -1, nullptr, true, false);
        release(up);
        release(h);
        h = o;
      }
    }
    const int out_cin = h.C;
    Packed po = pack({h}, NORM_GN, vec("out_norm.weight", h.C), vec("out_norm.bias", h.C), ACT_SILU, SP_ID, tc);
    release(h);
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
  void encoder(const std::string& pre, const std::vector<EncLayer>& L, bool gn, int in_ch, int img, bool from_y,
               float* feat, int feat_ld, int feat_off, int fc_in) {
    Packed cur = pack_input(in_ch, img, img, 16, from_y);
    const bool tc_saved = tc;
    // spatial size seen by each conv, and whether it can run on the tensor cores
    std::vector<int> sp(L.size());
    std::vector<bool> ltc(L.size());
    int hcur = img;
    for (size_t i = 0; i < L.size(); ++i) {
      sp[i] = hcur;
      ltc[i] = tc_saved && i > 0;
      if (L[i].pool) hcur /= 2;
    }
    for (size_t i = 0; i < L.size(); ++i) {
      const EncLayer& l = L[i];
      const bool last = (i + 1 == L.size());
      tc = ltc[i];  // layer 0 (K = 9/27) and odd GroupNorm geometries stay on the exact fp32 path
      // GroupNorm partials are only needed when the following pack writes an fp32 operand from a SIMT producer
      const bool next_padded = !last && ltc[i + 1];
      Act h = conv(cur, pre + l.conv, l.cin, l.cout, 3, 1, -1, nullptr, gn && !next_padded, false);
      tc = tc_saved;
      release(cur);
      const float *g, *b;
      if (gn) {
        g = vec(pre + l.norm + ".weight", l.cout);
        b = vec(pre + l.norm + ".bias", l.cout);
      } else {
        bn_affine(pre + l.norm, l.cout, &g, &b);
      }
      cur = pack({h}, gn ? NORM_GN : NORM_AFFINE, g, b, ACT_SILU, l.pool ? SP_POOL2 : SP_ID, last ? false : (bool)ltc[i + 1]);
      release(h);
    }
    const float* fw = vec(pre + "fc.weight", 256 * fc_in);
    const float* fb = vec(pre + "fc.bias", 256);
    const float* src = (const float*)cur.p;
    const int HW = cur.H * cur.W, C = cur.C;
    net.ops_.push_back({[=](const RunCtx& c) { launch_gap_fc(src, HW, C, fw, fb, 256, feat + feat_off, feat_ld, c.nimg, c.s); },
                        "gap_fc", 2.0 * 256 * C, (double)HW * C * 4.0});
    release(cur);
  }

  void score_net(const float* feat, const std::vector<std::pair<int, int>>& lin_ln, int last, const std::vector<int>& dims) {
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
    net.ops_.push_back({[=](const RunCtx& c) { launch_mlp(feat, 512, layers.data(), (int)layers.size(), c.out, 1, c.nimg, c.s); },
                        "mlp", mf, 512 * 4.0});
  }

  void build_ratio28() {  // reference: ratio_estimator.py:34-158
    net.in_ch_ = 1; net.in_ch_y_ = 1; net.img_ = 28;
    float* feat = (float*)pool.persistent((size_t)chunk * 512 * sizeof(float));
    std::vector<EncLayer> L{{"conv1", "gn1", 1, 32, true}, {"conv2", "gn2", 32, 64, true}, {"conv3", "gn3", 64, 128, true}, {"conv4", "gn4", 128, 128, false}};
    encoder("encoder_x.", L, true, 1, 28, false, feat, 512, 0, 128);
    encoder("encoder_y.", L, true, 1, 28, true, feat, 512, 256, 128);
    score_net(feat, {{0, 1}, {4, 5}}, 8, {512, 256});
  }

  void build_ratio_mnist_svhn() {  // reference: ratio_flexible.py:185-364
    net.in_ch_ = 1; net.in_ch_y_ = 3; net.img_ = 32;
    float* feat = (float*)pool.persistent((size_t)chunk * 512 * sizeof(float));
    std::vector<EncLayer> M{{"conv1", "bn1", 1, 32, true}, {"conv2", "bn2", 32, 64, true}, {"conv3", "bn3", 64, 128, true}, {"conv4", "bn4", 128, 128, false}};
    encoder("encoder_mnist.", M, false, 1, 32, false, feat, 512, 0, 128);
    std::vector<EncLayer> S{{"conv1a", "bn1a", 3, 64, false},   {"conv1b", "bn1b", 64, 64, true},   {"conv2a", "bn2a", 64, 128, false},
                            {"conv2b", "bn2b", 128, 128, true}, {"conv3a", "bn3a", 128, 256, false}, {"conv3b", "bn3b", 256, 256, true},
                            {"conv4a", "bn4a", 256, 256, false}, {"conv4b", "bn4b", 256, 256, true}};
    encoder("encoder_svhn.", S, false, 3, 32, true, feat, 512, 256, 256);
    score_net(feat, {{0, 1}, {4, 5}, {8, 9}}, 11, {512, 512, 256});
  }
};

// ------------------------------------------------------------------------------------------------
Net::Net(int kind, const rgfm_weights* w, int chunk, int precision, int device)
    : kind_(kind), chunk_(chunk), precision_(precision), device_(device) {
  RGFM_CHECK(chunk >= 1, "net: chunk");
  RGFM_CHECK(precision == RGFM_PREC_FP32 || precision == RGFM_PREC_BF16 || precision == RGFM_PREC_FP16,
             "net: precision must be fp32, bf16 or fp16");
  RGFM_CUDA(cudaSetDevice(device));
  if (precision != RGFM_PREC_FP32) RGFM_CHECK(tc_available(), "16-bit tensor-core precision needs an sm_100 device");
  WeightMap W(w);
  Builder b(*this, W);
  switch (kind) {
    case RGFM_NET_UNET_MNIST28: b.build_unet(1, 28, 32, {1, 2}, 2); break;
    case RGFM_NET_UNET_MNIST32: b.build_unet(1, 32, 32, {1, 2}, 2); break;
    case RGFM_NET_UNET_SVHN: b.build_unet(3, 32, 64, {1, 2, 2}, 2); break;
    case RGFM_NET_RATIO28: b.build_ratio28(); break;
    case RGFM_NET_RATIO_MNIST_SVHN: b.build_ratio_mnist_svhn(); break;
    default: throw Error("net kind " + std::to_string(kind) + " is not implemented");
  }
  if (is_unet()) set_times({0.0f});
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
  std::vector<float> e(model_ch_), h1(temb_dim_), h2(temb_dim_);
  for (size_t r = 0; r < ts.size(); ++r) {
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

void Net::forward(cudaStream_t s, int nimg, const float* x, const float* y, float* out, int out_mode, float dt,
                  const int* step, std::vector<ProfRec>* prof) {
  const size_t dx = (size_t)in_ch_ * img_ * img_;
  const size_t dy = (size_t)in_ch_y_ * img_ * img_;
  for (int c0 = 0; c0 < nimg; c0 += chunk_) {
    RunCtx c;
    c.s = s;
    c.nimg = std::min(chunk_, nimg - c0);
    c.x = x + c0 * dx;
    c.y = y ? y + c0 * dy : nullptr;
    c.out = is_unet() ? out + c0 * dx : out + c0;
    c.out_mode = out_mode;
    c.dt = dt;
    c.step = step;
    c.prof = prof;
    for (auto& op : ops_) {
      if (prof) {
        ProfRec r{op.kind, op.label.c_str(), nullptr, nullptr, op.flops_per_img * c.nimg, op.bytes_per_img * c.nimg};
        RGFM_CUDA(cudaEventCreate(&r.e0));
        RGFM_CUDA(cudaEventCreate(&r.e1));
        RGFM_CUDA(cudaEventRecord(r.e0, s));
        op.fn(c);
        RGFM_CUDA(cudaEventRecord(r.e1, s));
        prof->push_back(r);
      } else {
        op.fn(c);
      }
    }
  }
}

}  // namespace rgfm

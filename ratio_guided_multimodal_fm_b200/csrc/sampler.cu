// The coupled guided Euler sampler as a plan: MC pre-pass, ratio forward, N-step main loop, all captured as
// per-step CUDA graphs driven by a device-side step counter (no host synchronisation inside the loop).
// reference: flow_utils.py:178-375 (28x28 pair), sample_mnist_svhn.py:39-177 (MNIST32 <-> SVHN).
#include "net.h"
#include <cmath>
#include <cstdlib>
#include <algorithm>

namespace rgfm {

struct Graph {
  cudaGraph_t g = nullptr;
  cudaGraphExec_t e = nullptr;
  int64_t kernels = 0;
  ~Graph() {
    if (e) cudaGraphExecDestroy(e);
    if (g) cudaGraphDestroy(g);
  }
};

struct Sampler {
  rgfm_sampler_desc d;
  std::unique_ptr<Net> fx, fy, fr;
  int Dx = 0, Dy = 0, B = 0, N = 0, steps = 0;
  bool guided = false;      // mc_feng with a ratio estimator
  bool grad_guided = false; // grad_log_ratio with a ratio estimator
  float dtf = 0.f;
  DevPool pool;
  float *x = nullptr, *y = nullptr, *vx = nullptr, *vy = nullptr, *gx = nullptr, *gy = nullptr;
  float *mcx = nullptr, *mcy = nullptr, *scores = nullptr, *ratios = nullptr, *scratch = nullptr;
  float* diag = nullptr;
  int* step = nullptr;
  StepScalars* sc = nullptr;
  std::vector<StepScalars> hsc;  // host copy of the per-step scalars (set_strength rewrites gamma)
  bool bank_valid = false;       // a guided run has left its MC bank (mc_x, mc_y, mc_ratios) in the plan
  int first_guided = 0;  // first step with t > eps
  int ksx = 1, ksy = 1;
  cudaStream_t cap = nullptr, cap2 = nullptr;   // capture streams: the x- and y-network forwards are independent branches
  cudaStream_t cap3 = nullptr;                  // third branch of a guided step: mc_feng distances + weights (need x_t, y_t only)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join3 = nullptr;
  Graph g_pre, g_ratio, g_plain, g_guided, g_grad;
  int64_t launches = 0;
  float *hx = nullptr, *hy = nullptr, *hmx = nullptr, *hmy = nullptr;  // pinned staging for run_host
  // evaluators attached to the plan (rgfm_sampler_attach_evaluators): classifier forwards + argmax match on the finished
  // pairs at the end of every run, count in diag[15].  The nets are the caller's (they must outlive the plan's runs).
  Net *clf_x = nullptr, *clf_y = nullptr;
  int clf_transform = 0;
  float *clf_lx = nullptr, *clf_ly = nullptr, *clf_ty = nullptr;
  std::vector<float*> stage;  // device staging for run_host

  ~Sampler() {
    cudaSetDevice(d.device);
    if (cap) cudaStreamDestroy(cap);
    if (cap2) cudaStreamDestroy(cap2);
    if (cap3) cudaStreamDestroy(cap3);
    if (ev_join3) cudaEventDestroy(ev_join3);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (hx) cudaFreeHost(hx);
    if (hy) cudaFreeHost(hy);
    if (hmx) cudaFreeHost(hmx);
    if (hmy) cudaFreeHost(hmy);
  }

  McfArgs mcf_args() const {
    McfArgs a{};
    a.B = B; a.N = N; a.Dx = Dx; a.Dy = Dy;
    a.x = x; a.y = y; a.vx = vx; a.vy = vy; a.mcx = mcx; a.mcy = mcy; a.ratios = ratios;
    a.sc = sc; a.step = step; a.x_out = x; a.y_out = y; a.gx = gx; a.gy = gy; a.w = nullptr;
    a.logits = scratch;
    a.wbuf = scratch + (size_t)(ksx + ksy) * B * N;
    a.zbuf = a.wbuf + (size_t)B * N;
    a.ksplit_x = ksx; a.ksplit_y = ksy;
    a.g_step = (int)(0.3 * steps);  // g_x / g_y are only read by the diagnostics of that step (flow_utils.py:349-363)
    return a;
  }

  // One main-loop step run eagerly with every kernel bracketed by CUDA events on the launching stream.
  // Destroys the plan's transient state (the next run() re-initialises it); call after a run().
  std::vector<ProfRec> profile_step() {
    RGFM_CUDA(cudaSetDevice(d.device));
    std::vector<ProfRec> recs;
    const int st = std::min(steps - 1, std::max(first_guided, steps / 2));
    RGFM_CUDA(cudaMemcpyAsync(step, &st, sizeof(int), cudaMemcpyHostToDevice, cap));
    RGFM_CUDA(cudaStreamSynchronize(cap));
    if (guided && st >= first_guided) {
      fx->forward(cap, B, x, nullptr, vx, OUT_NCHW, dtf, step, &recs);
      fy->forward(cap, B, y, nullptr, vy, OUT_NCHW, dtf, step, &recs);
      launch_mcf(mcf_args(), cap, &recs);
    } else {
      fx->forward(cap, B, x, nullptr, x, OUT_EULER, dtf, step, &recs);
      fy->forward(cap, B, y, nullptr, y, OUT_EULER, dtf, step, &recs);
    }
    RGFM_CUDA(cudaStreamSynchronize(cap));
    // the same step as ONE CUDA-graph launch (what run() executes): no per-kernel CPU launch gaps
    {
      const int reps = 5;
      cudaGraphExec_t ge = (guided && st >= first_guided) ? g_guided.e : g_plain.e;
      ProfRec r{"graph_step", nullptr, nullptr, nullptr, 0.0, 0.0};
      RGFM_CUDA(cudaEventCreate(&r.e0));
      RGFM_CUDA(cudaEventCreate(&r.e1));
      RGFM_CUDA(cudaMemcpyAsync(step, &st, sizeof(int), cudaMemcpyHostToDevice, cap));
      RGFM_CUDA(cudaGraphLaunch(ge, cap));  // warm
      RGFM_CUDA(cudaEventRecord(r.e0, cap));
      for (int i = 0; i < reps; ++i) {
        RGFM_CUDA(cudaMemcpyAsync(step, &st, sizeof(int), cudaMemcpyHostToDevice, cap));  // replay the same step
        RGFM_CUDA(cudaGraphLaunch(ge, cap));
      }
      RGFM_CUDA(cudaEventRecord(r.e1, cap));
      RGFM_CUDA(cudaStreamSynchronize(cap));
      graph_reps = reps;
      recs.push_back(r);
    }
    return recs;
  }
  int graph_reps = 1;

  template <typename F>
  void capture(Graph& G, F&& body) {
    const int64_t before = g_launch_counter;
    RGFM_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    try {
      body(cap);
    } catch (...) {
      cudaGraph_t junk = nullptr;
      cudaStreamEndCapture(cap, &junk);
      if (junk) cudaGraphDestroy(junk);
      throw;
    }
    RGFM_CUDA(cudaStreamEndCapture(cap, &G.g));
    RGFM_CUDA(cudaGraphInstantiate(&G.e, G.g, 0));
    G.kernels = g_launch_counter - before;
  }

  void build(const rgfm_weights* wx, const rgfm_weights* wy, const rgfm_weights* wr) {
    RGFM_CHECK(d.abi_version == RGFM_ABI_VERSION, "rgfm_sampler_desc.abi_version mismatch");
    RGFM_CHECK(d.num_samples >= 1 && d.num_steps >= 1, "num_samples and num_steps must be positive");
    RGFM_CUDA(cudaSetDevice(d.device));
    B = d.num_samples;
    N = d.mc_batch_size;
    steps = d.num_steps;
    // Unknown methods, or mc_feng without a ratio estimator, silently degrade to unguided exactly like the
    // reference (flow_utils.py:231,273).
    guided = (d.guidance == RGFM_GUIDE_MC_FENG) && wr != nullptr && N >= 1;
    grad_guided = (d.guidance == RGFM_GUIDE_GRAD_LOG_RATIO) && wr != nullptr;
    int kx, ky, kr;
    if (d.pair == RGFM_PAIR_28) {
      // src/sample.py --model {unet, original}: the same sampler takes either flow; the state_dict says which one it is
      auto kind28 = [](const rgfm_weights* w) {
        return WeightMap(w).has("encoder.conv1.weight") ? (int)RGFM_NET_FLOW28 : (int)RGFM_NET_UNET_MNIST28;
      };
      kx = kind28(wx);
      ky = kind28(wy);
      kr = RGFM_NET_RATIO28;
    } else if (d.pair == RGFM_PAIR_MNIST_SVHN) {
      kx = RGFM_NET_UNET_MNIST32;
      ky = RGFM_NET_UNET_SVHN;
      kr = RGFM_NET_RATIO_MNIST_SVHN;
    } else {
      throw Error("unknown pair kind");
    }
    // images per conv launch: 1024, or 2048 for large batches (every launch carries ~10 us that does not scale with its
    // size: B = 2048 / 4096 main loops 1425 -> 1390 / 2844 -> 2778 ms, bit-identical pairs; 4096 would overflow the fast row division)
    int chunk = d.chunk > 0 ? d.chunk : (B >= 2048 ? 2048 : 1024);
    chunk = std::min(chunk, std::max(B, guided ? N : 1));
    fx.reset(new Net(kx, wx, chunk, d.precision, d.device));
    fy.reset(new Net(ky, wy, chunk, d.precision, d.device));
    if (guided) fr.reset(new Net(kr, wr, std::min(chunk, N), d.precision, d.device));
    // gradient guidance: forward + dgrad of the estimator at the sampler's batch, every step, in the plan's precision
    // (16-bit modes: both conv chains on the tensor cores, net.cu encoder(); fp32: exact, 256-image chunks)
    if (grad_guided)
      fr.reset(new Net(kr, wr, d.precision == RGFM_PREC_FP32 ? std::min(256, B) : std::min(chunk, B), d.precision, d.device, /*want_grad=*/true));
    Dx = (int)fx->dim();
    Dy = (int)fy->dim();

    // per-step scalars, computed in double like the reference's Python floats and rounded once to fp32
    const double dt = 1.0 / steps, eps = 1e-3;
    dtf = (float)dt;
    std::vector<float> ts(steps);
    std::vector<StepScalars> hs(steps);
    first_guided = steps;
    for (int s = 0; s < steps; ++s) {
      const double t = s * dt;
      ts[s] = (float)t;  // torch.full((B,), t) -> fp32
      const double sigma = 1.0 - t + eps;
      hs[s].t = (float)t;
      hs[s].sig2 = (float)(sigma * sigma);
      hs[s].sigma = (float)sigma;
      hs[s].inv_sigma = (float)(1.0 / sigma);
      hs[s].one_minus_gamma = (float)(1.0 - d.guidance_strength);
      hs[s].gamma = (float)d.guidance_strength;
      hs[s].dt = (float)dt;
      hs[s].pad = 0.f;
      if (t > eps && first_guided == steps) first_guided = s;
    }
    fx->set_times(ts);
    fy->set_times(ts);
    sc = pool.upload(hs);
    hsc = hs;
    step = (int*)pool.persistent(sizeof(int));
    diag = (float*)pool.persistent(RGFM_DIAG_FLOATS * sizeof(float));
    x = (float*)pool.persistent((size_t)B * Dx * sizeof(float));
    y = (float*)pool.persistent((size_t)B * Dy * sizeof(float));
    if (grad_guided) {
      vx = (float*)pool.persistent((size_t)B * Dx * sizeof(float));
      vy = (float*)pool.persistent((size_t)B * Dy * sizeof(float));
      gx = (float*)pool.persistent((size_t)B * Dx * sizeof(float));
      gy = (float*)pool.persistent((size_t)B * Dy * sizeof(float));
    }
    if (guided) {
      vx = (float*)pool.persistent((size_t)B * Dx * sizeof(float));
      vy = (float*)pool.persistent((size_t)B * Dy * sizeof(float));
      gx = (float*)pool.persistent((size_t)B * Dx * sizeof(float));
      gy = (float*)pool.persistent((size_t)B * Dy * sizeof(float));
      mcx = (float*)pool.persistent((size_t)N * Dx * sizeof(float));
      mcy = (float*)pool.persistent((size_t)N * Dy * sizeof(float));
      scores = (float*)pool.persistent((size_t)N * sizeof(float));
      ratios = (float*)pool.persistent((size_t)N * sizeof(float));
      const size_t nf = mcf_scratch_floats(B, N, Dx, Dy, &ksx, &ksy);
      scratch = (float*)pool.persistent(nf * sizeof(float));
    }
    RGFM_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    RGFM_CUDA(cudaStreamCreateWithFlags(&cap2, cudaStreamNonBlocking));
    RGFM_CUDA(cudaStreamCreateWithFlags(&cap3, cudaStreamNonBlocking));
    RGFM_CUDA(cudaEventCreateWithFlags(&ev_join3, cudaEventDisableTiming));
    RGFM_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    RGFM_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    // fork / join inside a capture: the second stream becomes a parallel branch of the graph, so the tail of one
    // network's persistent kernels overlaps the head of the other's (and small-batch pre-pass kernels run side by side)
    auto fork = [&](cudaStream_t s) {
      RGFM_CUDA(cudaEventRecord(ev_fork, s));
      RGFM_CUDA(cudaStreamWaitEvent(cap2, ev_fork, 0));
    };
    auto join = [&](cudaStream_t s) {
      RGFM_CUDA(cudaEventRecord(ev_join, cap2));
      RGFM_CUDA(cudaStreamWaitEvent(s, ev_join, 0));
    };

    // SM split between the two U-Net chains (common.h g_tc_cta_cap): the x-network's conv launches take `sx` CTAs, the
    // y-network's the remaining SMs, so the two branches of a step run side by side instead of alternating and the
    // pipeline fill / drain of every launch (~10 us per launch of a step that has 68 of them) is covered by the other chain.
    // It pays where that fixed cost is a large share of a launch, i.e. at small batches -- the MC pre-pass (N_mc = 256:
    // 205 -> 190 ms per call, plateau at 40..46 of 148 SMs for the MNIST32 chain = its share of the step's conv time; 36 or
    // 56 lose it again) -- and costs 0..4 % at B = 1024, where the launches are long (profiles/notes_r02.md section 14).
    // Rule: identical networks (28x28 pair) are always split in half (a balanced split still gains 1.5..2 % at B = 1024 /
    // 2048); the MNIST32-SVHN pair by the chains' share of the conv time when the batch is at most 768 images (B = 768:
    // 548 -> 540 ms; B = 1024: 709 vs 714 ms).  RGFM_SPLIT_PRE / RGFM_SPLIT_MAIN override the x chain's CTA count per graph
    // family (0 = no split).
    int num_sms = 148;
    RGFM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, d.device));
    const bool tc_mode = d.precision != RGFM_PREC_FP32;
    const bool twins = d.pair != RGFM_PAIR_MNIST_SVHN;
    const int sx_rule = twins ? num_sms / 2 : (num_sms * 2 + 3) / 7;
    const int split_pre = getenv("RGFM_SPLIT_PRE") ? atoi(getenv("RGFM_SPLIT_PRE")) : ((tc_mode && (twins || N <= 768)) ? sx_rule : 0);
    const int split_main = getenv("RGFM_SPLIT_MAIN") ? atoi(getenv("RGFM_SPLIT_MAIN")) : ((tc_mode && (twins || B <= 768)) ? sx_rule : 0);
    auto both = [&](int sx, cudaStream_t s, int n, const float* xi, float* xo, const float* yi, float* yo, int mode) {
      const bool on = sx > 0 && sx < num_sms;
      g_tc_cta_cap = on ? sx : 0;
      fx->forward(s, n, xi, nullptr, xo, mode, dtf, step);
      g_tc_cta_cap = on ? num_sms - sx : 0;
      fy->forward(cap2, n, yi, nullptr, yo, mode, dtf, step);
      g_tc_cta_cap = 0;
    };

    // ---- graphs ----
    if (guided) {
      // mc_x += fm_x(mc_x, t) * dt and mc_y += fm_y(mc_y, t) * dt (flow_utils.py:236-250): the reference runs the two
      // chains one after the other; they are independent, so step s of both runs concurrently here
      capture(g_pre, [&](cudaStream_t s) {
        fork(s);
        both(split_pre, s, N, mcx, mcx, mcy, mcy, OUT_EULER);
        join(s);
        launch_step_inc(step, s);
      });
      capture(g_ratio, [&](cudaStream_t s) {  // mc_ratios = exp(log_ratio(mc_x, mc_y))  (flow_utils.py:255-258)
        fr->forward(s, N, mcx, mcy, scores, 0, 0.f, nullptr);
        launch_ratio_head(scores, d.loss_type, nullptr, ratios, N, s);
        launch_ratio_stats(ratios, N, diag + 9, s);
      });
      // RGFM_MCF_BRANCH=0: distances + weights after the U-Nets on the main branch (A/B timing; same results)
      static const bool mcf_branch = !(getenv("RGFM_MCF_BRANCH") && atoi(getenv("RGFM_MCF_BRANCH")) == 0);
      capture(g_guided, [&](cudaStream_t s) {
        McfArgs a = mcf_args();
        fork(s);
        if (mcf_branch) {  // third branch: the importance weights depend on (x_t, y_t) only -- beside the U-Nets
          RGFM_CUDA(cudaStreamWaitEvent(cap3, ev_fork, 0));
          launch_mcf_prep(a, cap3);
          RGFM_CUDA(cudaEventRecord(ev_join3, cap3));
        }
        both(split_main, s, B, x, vx, y, vy, OUT_NCHW);
        join(s);
        if (mcf_branch) RGFM_CUDA(cudaStreamWaitEvent(s, ev_join3, 0));
        else launch_mcf_prep(a, s);
        launch_mcf_apply(a, s);
        DiagArgs da{B, Dx, Dy, N, vx, vy, gx, gy, a.zbuf, sc, step, (int)(0.3 * steps), diag};
        launch_diag(da, s);
        launch_step_inc(step, s);
      });
    }
    if (grad_guided) {
      // v = v_ind + gamma * grad log r(x_t, y_t) (README.md:163), then x += v * dt; every step, no MC bank
      const int lt = d.loss_type;
      capture(g_grad, [&](cudaStream_t s) {
        fork(s);
        both(split_main, s, B, x, vx, y, vy, OUT_NCHW);
        join(s);
        fr->grad_log_ratio(s, B, x, y, lt, gx, gy, nullptr);
        launch_euler_grad(x, vx, gx, sc, step, dtf, (size_t)B * Dx, s);
        launch_euler_grad(y, vy, gy, sc, step, dtf, (size_t)B * Dy, s);
        launch_step_inc(step, s);
      });
    }
    capture(g_plain, [&](cudaStream_t s) {  // x += fm_x(x, t) * dt ; y += fm_y(y, t) * dt
      fork(s);
      both(split_main, s, B, x, x, y, y, OUT_EULER);
      join(s);
      launch_step_inc(step, s);
    });
    RGFM_CUDA(cudaStreamSynchronize(cap));
    launches = 0;
    if (guided) launches += (int64_t)steps * g_pre.kernels + g_ratio.kernels;
    for (int s = 0; s < steps; ++s)
      launches += grad_guided ? g_grad.kernels : ((guided && s >= first_guided) ? g_guided.kernels : g_plain.kernels);
  }

  // gamma lives in the per-step scalar table the kernels read through a pointer: a sweep over guidance strengths -- or a
  // time-dependent schedule gamma(t_k), one value per Euler step -- needs no new plan and no graph re-capture.
  void set_schedule(const double* gammas, int n, cudaStream_t s) {
    RGFM_CHECK(gammas != nullptr && n == (int)hsc.size(), "rgfm_sampler_set_strength_schedule: one strength per Euler step");
    RGFM_CUDA(cudaSetDevice(d.device));
    for (int k = 0; k < n; ++k) {
      hsc[k].one_minus_gamma = (float)(1.0 - gammas[k]);
      hsc[k].gamma = (float)gammas[k];
    }
    RGFM_CUDA(cudaMemcpyAsync(sc, hsc.data(), hsc.size() * sizeof(StepScalars), cudaMemcpyHostToDevice, s));
    RGFM_CUDA(cudaStreamSynchronize(s));  // hsc may change again before an asynchronous copy has read it
  }
  void set_strength(double gamma, cudaStream_t s) {
    d.guidance_strength = gamma;
    std::vector<double> g(hsc.size(), gamma);
    set_schedule(g.data(), (int)g.size(), s);
  }

  // reuse_bank: skip the MC pre-pass and the ratio pass and guide with the bank of the previous run on this plan
  // (the evaluators' gamma sweep recomputes an identical-in-distribution bank per gamma)
  void run(const float* x0, const float* y0, const float* mx0, const float* my0, float* x1, float* y1,
           float* diag_out, cudaStream_t s, bool reuse_bank = false) {
    RGFM_CHECK(x0 && y0 && x1 && y1, "rgfm_sampler_run: null state pointer");
    RGFM_CHECK(!reuse_bank || (guided && bank_valid), "rgfm_sampler_run_reuse_bank: needs an mc_feng plan that has completed a full run");
    RGFM_CUDA(cudaSetDevice(d.device));
    RGFM_CUDA(cudaMemcpyAsync(x, x0, (size_t)B * Dx * sizeof(float), cudaMemcpyDeviceToDevice, s));
    RGFM_CUDA(cudaMemcpyAsync(y, y0, (size_t)B * Dy * sizeof(float), cudaMemcpyDeviceToDevice, s));
    // diag[9..] are the bank's ratio statistics (written by the ratio pass): kept when the bank is reused
    RGFM_CUDA(cudaMemsetAsync(diag, 0, (reuse_bank ? 9 : RGFM_DIAG_FLOATS) * sizeof(float), s));
    if (reuse_bank) RGFM_CUDA(cudaMemsetAsync(diag + 12, 0, 4 * sizeof(float), s));  // per-run counters
    for (Net* n : {fx.get(), fy.get(), fr.get()})
      if (n) RGFM_CUDA(cudaMemsetAsync(n->sat_counter(), 0, sizeof(unsigned int), s));
    if (guided && !reuse_bank) {
      RGFM_CHECK(mx0 && my0, "rgfm_sampler_run: mc_feng needs mc_x0 and mc_y0");
      RGFM_CUDA(cudaMemcpyAsync(mcx, mx0, (size_t)N * Dx * sizeof(float), cudaMemcpyDeviceToDevice, s));
      RGFM_CUDA(cudaMemcpyAsync(mcy, my0, (size_t)N * Dy * sizeof(float), cudaMemcpyDeviceToDevice, s));
      RGFM_CUDA(cudaMemsetAsync(step, 0, sizeof(int), s));
      for (int i = 0; i < steps; ++i) RGFM_CUDA(cudaGraphLaunch(g_pre.e, s));
      RGFM_CUDA(cudaGraphLaunch(g_ratio.e, s));
      bank_valid = true;
    }
    RGFM_CUDA(cudaMemsetAsync(step, 0, sizeof(int), s));
    for (int i = 0; i < steps; ++i)
      RGFM_CUDA(cudaGraphLaunch(grad_guided ? g_grad.e : ((guided && i >= first_guided) ? g_guided.e : g_plain.e), s));
    RGFM_CUDA(cudaMemcpyAsync(x1, x, (size_t)B * Dx * sizeof(float), cudaMemcpyDeviceToDevice, s));
    RGFM_CUDA(cudaMemcpyAsync(y1, y, (size_t)B * Dy * sizeof(float), cudaMemcpyDeviceToDevice, s));
    launch_sat_report(fx->sat_counter(), fy->sat_counter(), fr ? fr->sat_counter() : nullptr, diag + 12, s);
    if (clf_x) {  // fused evaluator tail: the pairs are still in HBM / L2, nothing goes back to Python but the count
      const float* ys = y;
      if (clf_transform != 0) {
        launch_inverse_transform(y, clf_ty, B, fy->img(), clf_transform, s);
        ys = clf_ty;
      }
      clf_x->forward(s, B, x, nullptr, clf_lx, 0, 0.f, nullptr);
      clf_y->forward(s, B, ys, nullptr, clf_ly, 0, 0.f, nullptr);
      launch_argmax_match(clf_lx, clf_ly, B, diag + 15, s);
    }
    if (diag_out) RGFM_CUDA(cudaMemcpyAsync(diag_out, diag, RGFM_DIAG_FLOATS * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
};

// ---- C-ABI glue (api.cu calls these) ----
Sampler* sampler_create(const rgfm_sampler_desc* d, const rgfm_weights* wx, const rgfm_weights* wy, const rgfm_weights* wr) {
  RGFM_CHECK(d && wx && wy, "rgfm_sampler_create: null argument");
  std::unique_ptr<Sampler> s(new Sampler());
  s->d = *d;
  s->build(wx, wy, wr);
  return s.release();
}
void sampler_destroy(Sampler* s) { delete s; }
void sampler_run(Sampler* s, const float* x0, const float* y0, const float* mx0, const float* my0, float* x1, float* y1,
                 float* diag, cudaStream_t st) {
  s->run(x0, y0, mx0, my0, x1, y1, diag, st);
}
void sampler_run_reuse_bank(Sampler* s, const float* x0, const float* y0, float* x1, float* y1, float* diag, cudaStream_t st) {
  s->run(x0, y0, nullptr, nullptr, x1, y1, diag, st, /*reuse_bank=*/true);
}
void sampler_set_strength(Sampler* s, double gamma, cudaStream_t st) { s->set_strength(gamma, st); }
void sampler_set_schedule(Sampler* s, const double* g, int n, cudaStream_t st) { s->set_schedule(g, n, st); }
void sampler_attach_evaluators(Sampler* s, Net* cx, Net* cy, int inverse_transform) {
  RGFM_CHECK((cx == nullptr) == (cy == nullptr), "rgfm_sampler_attach_evaluators: both classifiers or none");
  if (cx) {
    RGFM_CHECK(cx->kind() >= RGFM_NET_CLF_MNIST28 && cx->kind() <= RGFM_NET_CLF_SVHN && cy->kind() >= RGFM_NET_CLF_MNIST28 &&
                   cy->kind() <= RGFM_NET_CLF_SVHN, "rgfm_sampler_attach_evaluators: not classifiers");
    RGFM_CHECK(cx->dim() == (size_t)s->Dx && cy->dim() == (size_t)s->Dy, "rgfm_sampler_attach_evaluators: classifier input shapes do not match the pair");
    RGFM_CHECK(inverse_transform >= 0 && inverse_transform <= 6, "rgfm_sampler_attach_evaluators: unknown transform");
    RGFM_CHECK(inverse_transform == 0 || s->fy->in_ch() == 1, "rgfm_sampler_attach_evaluators: transforms apply to single-channel pairs");
    RGFM_CUDA(cudaSetDevice(s->d.device));
    if (!s->clf_lx) {
      s->clf_lx = (float*)s->pool.persistent((size_t)s->B * 10 * sizeof(float));
      s->clf_ly = (float*)s->pool.persistent((size_t)s->B * 10 * sizeof(float));
      s->clf_ty = (float*)s->pool.persistent((size_t)s->B * s->Dy * sizeof(float));
    }
  }
  s->clf_x = cx;
  s->clf_y = cy;
  s->clf_transform = inverse_transform;
}
int64_t sampler_launches(const Sampler* s) { return s->launches; }

void sampler_run_host(Sampler* s, const float* x0, const float* y0, const float* mx0, const float* my0, float* x1,
                      float* y1, float* diag) {
  RGFM_CUDA(cudaSetDevice(s->d.device));
  const size_t bx = (size_t)s->B * s->Dx * sizeof(float), by = (size_t)s->B * s->Dy * sizeof(float);
  const size_t mx = (size_t)s->N * s->Dx * sizeof(float), my = (size_t)s->N * s->Dy * sizeof(float);
  if (!s->hx) {
    RGFM_CUDA(cudaMallocHost(&s->hx, bx));
    RGFM_CUDA(cudaMallocHost(&s->hy, by));
    if (s->guided) {
      RGFM_CUDA(cudaMallocHost(&s->hmx, mx));
      RGFM_CUDA(cudaMallocHost(&s->hmy, my));
    }
  }
  // staging in device memory: reuse v buffers' worth of fresh allocations (kept for the plan's lifetime)
  auto& st = s->stage;
  if (st.empty()) {
    st.resize(5, nullptr);
    st[0] = (float*)s->pool.persistent(bx);
    st[1] = (float*)s->pool.persistent(by);
    if (s->guided) {
      st[2] = (float*)s->pool.persistent(mx);
      st[3] = (float*)s->pool.persistent(my);
    }
    st[4] = (float*)s->pool.persistent(RGFM_DIAG_FLOATS * sizeof(float));
  }
  cudaStream_t cs = s->cap;
  memcpy(s->hx, x0, bx);
  memcpy(s->hy, y0, by);
  RGFM_CUDA(cudaMemcpyAsync(st[0], s->hx, bx, cudaMemcpyHostToDevice, cs));
  RGFM_CUDA(cudaMemcpyAsync(st[1], s->hy, by, cudaMemcpyHostToDevice, cs));
  if (s->guided) {
    RGFM_CHECK(mx0 && my0, "rgfm_sampler_run_host: mc_feng needs mc_x0 and mc_y0");
    memcpy(s->hmx, mx0, mx);
    memcpy(s->hmy, my0, my);
    RGFM_CUDA(cudaMemcpyAsync(st[2], s->hmx, mx, cudaMemcpyHostToDevice, cs));
    RGFM_CUDA(cudaMemcpyAsync(st[3], s->hmy, my, cudaMemcpyHostToDevice, cs));
  }
  s->run(st[0], st[1], st[2], st[3], st[0], st[1], st[4], cs);
  RGFM_CUDA(cudaMemcpyAsync(s->hx, st[0], bx, cudaMemcpyDeviceToHost, cs));
  RGFM_CUDA(cudaMemcpyAsync(s->hy, st[1], by, cudaMemcpyDeviceToHost, cs));
  float hd[RGFM_DIAG_FLOATS];
  RGFM_CUDA(cudaMemcpyAsync(hd, st[4], sizeof(hd), cudaMemcpyDeviceToHost, cs));
  RGFM_CUDA(cudaStreamSynchronize(cs));
  memcpy(x1, s->hx, bx);
  memcpy(y1, s->hy, by);
  if (diag) memcpy(diag, hd, sizeof(hd));
}

int sampler_profile(Sampler* s, rgfm_prof_entry* out, int max_entries) {
  std::vector<ProfRec> recs = s->profile_step();
  std::vector<rgfm_prof_entry> agg;
  const bool verbose = getenv("RGFM_PROFILE_VERBOSE") != nullptr;
  for (auto& r : recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    ms /= (float)r.reps;
    if (verbose) fprintf(stderr, "PROF %-12s %8.2f us  %7.1f TFLOP/s  %s\n", r.kind, ms * 1e3, r.flops / (ms * 1e-3 + 1e-12) / 1e12, r.label ? r.label : "");
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
    size_t k = 0;
    for (; k < agg.size(); ++k)
      if (std::string(agg[k].name) == r.kind) break;
    if (k == agg.size()) {
      rgfm_prof_entry e{};
      snprintf(e.name, sizeof(e.name), "%s", r.kind);
      agg.push_back(e);
    }
    if (std::string(r.kind) == "graph_step") ms /= (float)s->graph_reps;
    agg[k].ms += ms;
    agg[k].flops += r.flops;
    agg[k].bytes += r.bytes;
    agg[k].launches += 1;
  }
  const int n = std::min<int>((int)agg.size(), max_entries);
  for (int i = 0; i < n; ++i) out[i] = agg[i];
  return n;
}

void sampler_get_bank(Sampler* s, float* mx, float* my, float* r, cudaStream_t st) {
  RGFM_CHECK(s->guided, "rgfm_sampler_get_bank: plan is unguided");
  if (mx) RGFM_CUDA(cudaMemcpyAsync(mx, s->mcx, (size_t)s->N * s->Dx * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (my) RGFM_CUDA(cudaMemcpyAsync(my, s->mcy, (size_t)s->N * s->Dy * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (r) RGFM_CUDA(cudaMemcpyAsync(r, s->ratios, (size_t)s->N * sizeof(float), cudaMemcpyDeviceToDevice, st));
}

}  // namespace rgfm

// extern "C" boundary of librgfm.so (include/rgfm.h).  Exceptions stop here: every entry point returns an int
// status and stores the message for rgfm_last_error().
#include "net.h"
#include <cstring>

namespace rgfm {
struct Sampler;
Sampler* sampler_create(const rgfm_sampler_desc*, const rgfm_weights*, const rgfm_weights*, const rgfm_weights*);
void sampler_destroy(Sampler*);
void sampler_run(Sampler*, const float*, const float*, const float*, const float*, float*, float*, float*, cudaStream_t);
void sampler_run_host(Sampler*, const float*, const float*, const float*, const float*, float*, float*, float*);
void sampler_get_bank(Sampler*, float*, float*, float*, cudaStream_t);
int64_t sampler_launches(const Sampler*);
void sampler_run_reuse_bank(Sampler*, const float*, const float*, float*, float*, float*, cudaStream_t);
void sampler_set_strength(Sampler*, double, cudaStream_t);
void sampler_set_schedule(Sampler*, const double*, int, cudaStream_t);
void sampler_attach_evaluators(Sampler*, Net*, Net*, int);
void launch_argmax_match(const float*, const float*, int, float*, cudaStream_t);
int sampler_profile(Sampler*, rgfm_prof_entry*, int);
}  // namespace rgfm

struct rgfm_ratio_trainer;
namespace rgfm {
rgfm_ratio_trainer* ratio_trainer_create(int, const rgfm_weights*, const rgfm_weights*, int, int);
void ratio_trainer_step(rgfm_ratio_trainer*, const float*, const float*, const float*, int, int, float, const float* const*, int, float*,
                        cudaStream_t);
void ratio_trainer_destroy(rgfm_ratio_trainer*);
void launch_ratio_loss(const float*, const float*, int, int, float, float*, float*, cudaStream_t);
void launch_cfm_interp(const float*, const float*, const float*, int, size_t, float*, float*, cudaStream_t);
void launch_fm_mse(const float*, const float*, size_t, float*, float*, float*, cudaStream_t);
void conv3x3_forward(const float*, int, int, int, int, const float*, const float*, int, int, float*, cudaStream_t);
void conv3x3_backward(const float*, const float*, int, int, int, int, const float*, int, int, float*, float*, float*, cudaStream_t);
void gn_nhwc_forward(const float*, int, int, int, int, const float*, const float*, float, float*, float*, cudaStream_t);
void gn_nhwc_backward(const float*, const float*, int, int, int, int, const float*, const float*, float*, float*, float*, cudaStream_t);
}  // namespace rgfm

static thread_local std::string g_err;

// Every entry point works on the plan's / net's device (cudaSetDevice inside) and puts the caller's current device back
// on the way out: PyTorch reads the current device from the runtime, so a library call that left it switched would send
// the caller's later `device="cuda"` allocations and streams to the wrong GPU.
struct DeviceRestore {
  int prev = -1;
  DeviceRestore() {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  }
  ~DeviceRestore() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
// device that owns a device pointer (cudaSetDevice to it); pointers the runtime does not know are left alone
static void set_device_of(const void* p) {
  cudaPointerAttributes at{};
  if (p && cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice) RGFM_CUDA(cudaSetDevice(at.device));
  else cudaGetLastError();
}

#define RGFM_API_BEGIN   \
  DeviceRestore _restore; \
  try {
#define RGFM_API_END                          \
  }                                           \
  catch (const std::exception& e) {           \
    g_err = e.what();                         \
    return 1;                                 \
  }                                           \
  catch (...) {                               \
    g_err = "unknown error";                  \
    return 1;                                 \
  }                                           \
  return 0;

struct rgfm_net {
  std::unique_ptr<rgfm::Net> net;
  float* scores = nullptr;  // ratio scratch
  int* step = nullptr;      // device step counter of rgfm_unet_sample
  int device = 0;
};

extern "C" {

const char* rgfm_last_error(void) { return g_err.c_str(); }
int rgfm_abi_version(void) { return RGFM_ABI_VERSION; }

int rgfm_sampler_create(const rgfm_sampler_desc* desc, const rgfm_weights* fm_x, const rgfm_weights* fm_y,
                        const rgfm_weights* ratio, rgfm_sampler** out) {
  RGFM_API_BEGIN
  RGFM_CHECK(out != nullptr, "rgfm_sampler_create: null out");
  *out = reinterpret_cast<rgfm_sampler*>(rgfm::sampler_create(desc, fm_x, fm_y, ratio));
  RGFM_API_END
}

int rgfm_sampler_run(rgfm_sampler* s, const float* x0, const float* y0, const float* mc_x0, const float* mc_y0,
                     float* x1, float* y1, float* diag, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(s != nullptr, "rgfm_sampler_run: null plan");
  rgfm::sampler_run(reinterpret_cast<rgfm::Sampler*>(s), x0, y0, mc_x0, mc_y0, x1, y1, diag, (cudaStream_t)stream);
  RGFM_API_END
}

int rgfm_sampler_run_host(rgfm_sampler* s, const float* x0, const float* y0, const float* mc_x0, const float* mc_y0,
                          float* x1, float* y1, float* diag) {
  RGFM_API_BEGIN
  RGFM_CHECK(s != nullptr, "rgfm_sampler_run_host: null plan");
  rgfm::sampler_run_host(reinterpret_cast<rgfm::Sampler*>(s), x0, y0, mc_x0, mc_y0, x1, y1, diag);
  RGFM_API_END
}

int rgfm_sampler_set_strength_schedule(rgfm_sampler* s, const double* guidance_strengths, int32_t num_steps, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(s != nullptr, "rgfm_sampler_set_strength_schedule: null plan");
  rgfm::sampler_set_schedule(reinterpret_cast<rgfm::Sampler*>(s), guidance_strengths, num_steps, (cudaStream_t)stream);
  RGFM_API_END
}

int rgfm_sampler_set_strength(rgfm_sampler* s, double guidance_strength, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(s != nullptr, "rgfm_sampler_set_strength: null plan");
  rgfm::sampler_set_strength(reinterpret_cast<rgfm::Sampler*>(s), guidance_strength, (cudaStream_t)stream);
  RGFM_API_END
}

int rgfm_sampler_attach_evaluators(rgfm_sampler* s, rgfm_net* clf_x, rgfm_net* clf_y, int32_t inverse_transform) {
  RGFM_API_BEGIN
  RGFM_CHECK(s != nullptr, "rgfm_sampler_attach_evaluators: null plan");
  rgfm::sampler_attach_evaluators(reinterpret_cast<rgfm::Sampler*>(s), clf_x ? clf_x->net.get() : nullptr, clf_y ? clf_y->net.get() : nullptr,
                                  inverse_transform);
  RGFM_API_END
}

int rgfm_coherence_matches(rgfm_net* clf_a, rgfm_net* clf_b, const float* xa, const float* xb, int32_t batch, float* logits_scratch,
                           float* matches, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(clf_a && clf_b && xa && xb && logits_scratch && matches && batch >= 1, "rgfm_coherence_matches: null argument");
  for (rgfm_net* n : {clf_a, clf_b})
    RGFM_CHECK(n->net->kind() >= RGFM_NET_CLF_MNIST28 && n->net->kind() <= RGFM_NET_CLF_SVHN, "rgfm_coherence_matches: not a classifier");
  RGFM_CUDA(cudaSetDevice(clf_a->device));
  cudaStream_t st = (cudaStream_t)stream;
  float* la = logits_scratch;
  float* lb = logits_scratch + (size_t)batch * 10;
  clf_a->net->forward(st, batch, xa, nullptr, la, 0, 0.f, nullptr);
  clf_b->net->forward(st, batch, xb, nullptr, lb, 0, 0.f, nullptr);
  rgfm::launch_argmax_match(la, lb, batch, matches, st);
  RGFM_API_END
}

int rgfm_sampler_run_reuse_bank(rgfm_sampler* s, const float* x0, const float* y0, float* x1, float* y1, float* diag,
                                void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(s != nullptr, "rgfm_sampler_run_reuse_bank: null plan");
  rgfm::sampler_run_reuse_bank(reinterpret_cast<rgfm::Sampler*>(s), x0, y0, x1, y1, diag, (cudaStream_t)stream);
  RGFM_API_END
}

int rgfm_sampler_get_bank(rgfm_sampler* s, float* mc_x, float* mc_y, float* mc_ratios, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(s != nullptr, "rgfm_sampler_get_bank: null plan");
  rgfm::sampler_get_bank(reinterpret_cast<rgfm::Sampler*>(s), mc_x, mc_y, mc_ratios, (cudaStream_t)stream);
  RGFM_API_END
}

int rgfm_sampler_profile_step(rgfm_sampler* s, rgfm_prof_entry* out, int32_t max_entries, int32_t* n_entries) {
  RGFM_API_BEGIN
  RGFM_CHECK(s && out && n_entries, "rgfm_sampler_profile_step: null argument");
  *n_entries = rgfm::sampler_profile(reinterpret_cast<rgfm::Sampler*>(s), out, max_entries);
  RGFM_API_END
}

int64_t rgfm_sampler_launch_count(const rgfm_sampler* s) {
  return s ? rgfm::sampler_launches(reinterpret_cast<const rgfm::Sampler*>(s)) : 0;
}

void rgfm_sampler_destroy(rgfm_sampler* s) {
  DeviceRestore _restore;
  if (s) rgfm::sampler_destroy(reinterpret_cast<rgfm::Sampler*>(s));
}

int rgfm_net_create(int32_t net_kind, const rgfm_weights* w, int32_t max_batch, int32_t precision, int32_t device,
                    rgfm_net** out) {
  RGFM_API_BEGIN
  RGFM_CHECK(out != nullptr, "rgfm_net_create: null out");
  std::unique_ptr<rgfm_net> n(new rgfm_net());
  n->device = device;
  n->net.reset(new rgfm::Net(net_kind, w, max_batch, precision, device));
  *out = n.release();
  RGFM_API_END
}

int rgfm_unet_forward(rgfm_net* net, const float* x, float t, int32_t batch, float* v, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(net && net->net->is_unet(), "rgfm_unet_forward: not a U-Net");
  RGFM_CUDA(cudaSetDevice(net->device));
  RGFM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));  // the time-bias row is rewritten synchronously
  net->net->set_times({t});
  net->net->forward((cudaStream_t)stream, batch, x, nullptr, v, rgfm::OUT_NCHW, 0.f, nullptr);
  RGFM_API_END
}

int rgfm_unet_sample(rgfm_net* net, float* x, int32_t batch, int32_t num_steps, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(net && net->net->is_unet(), "rgfm_unet_sample: not a U-Net");
  RGFM_CHECK(x != nullptr && batch >= 0 && num_steps >= 1, "rgfm_unet_sample: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  RGFM_CUDA(cudaSetDevice(net->device));
  RGFM_CUDA(cudaStreamSynchronize(s));  // the time-bias table is rewritten synchronously
  // dt and t = step * dt as Python doubles, rounded once to fp32 like torch.full((n,), t) (flow_utils.py:85-90)
  const double dt = 1.0 / num_steps;
  std::vector<float> ts(num_steps);
  for (int i = 0; i < num_steps; ++i) ts[i] = (float)(i * dt);
  net->net->set_times(ts);
  if (!net->step) RGFM_CUDA(cudaMalloc(&net->step, sizeof(int)));
  RGFM_CUDA(cudaMemsetAsync(net->step, 0, sizeof(int), s));
  for (int i = 0; i < num_steps && batch > 0; ++i) {  // x_t = x_t + v_t * dt, fused into the output conv (flow_utils.py:93-96)
    net->net->forward(s, batch, x, nullptr, x, rgfm::OUT_EULER, (float)dt, net->step);
    rgfm::launch_step_inc(net->step, s);
  }
  RGFM_API_END
}

int rgfm_ratio_log_ratio(rgfm_net* net, const float* x, const float* y, int32_t batch, int32_t loss_type,
                         float* log_ratio, float* scores, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(net && (net->net->kind() == RGFM_NET_RATIO28 || net->net->kind() == RGFM_NET_RATIO_MNIST_SVHN),
             "rgfm_ratio_log_ratio: not a ratio estimator");
  // unknown loss_type -> error, as ratio_estimator.py:190-191 raises ValueError
  RGFM_CHECK(loss_type == RGFM_LOSS_DISC || loss_type == RGFM_LOSS_RULSIF, "Unknown loss_type");
  RGFM_CUDA(cudaSetDevice(net->device));
  float* sc = scores;
  if (!sc) {
    RGFM_CUDA(cudaMalloc(&sc, (size_t)batch * sizeof(float)));
  }
  net->net->forward((cudaStream_t)stream, batch, x, y, sc, 0, 0.f, nullptr);
  rgfm::launch_ratio_head(sc, loss_type, log_ratio, nullptr, batch, (cudaStream_t)stream);
  if (!scores) {
    RGFM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    RGFM_CUDA(cudaFree(sc));
  }
  RGFM_API_END
}

int rgfm_ratio_grad_create_prec(int32_t net_kind, const rgfm_weights* w, int32_t max_batch, int32_t precision, int32_t device,
                                rgfm_net** out) {
  RGFM_API_BEGIN
  RGFM_CHECK(out != nullptr, "rgfm_ratio_grad_create: null out");
  RGFM_CHECK(precision == RGFM_PREC_FP32 || precision == RGFM_PREC_FP16 || precision == RGFM_PREC_BF16, "rgfm_ratio_grad_create: unknown precision");
  std::unique_ptr<rgfm_net> n(new rgfm_net());
  n->device = device;
  n->net.reset(new rgfm::Net(net_kind, w, max_batch, precision, device, /*want_grad=*/true));
  *out = n.release();
  RGFM_API_END
}

int rgfm_ratio_grad_create(int32_t net_kind, const rgfm_weights* w, int32_t max_batch, int32_t device, rgfm_net** out) {
  return rgfm_ratio_grad_create_prec(net_kind, w, max_batch, RGFM_PREC_FP32, device, out);
}

int rgfm_ratio_grad_log_ratio(rgfm_net* net, const float* x, const float* y, int32_t batch, int32_t loss_type, float* grad_x,
                              float* grad_y, float* log_ratio, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(net && net->net->has_grad(), "rgfm_ratio_grad_log_ratio: not a gradient-mode ratio estimator");
  RGFM_CHECK(x && y && grad_x && grad_y && batch >= 1, "rgfm_ratio_grad_log_ratio: null argument");
  RGFM_CHECK(loss_type == RGFM_LOSS_DISC || loss_type == RGFM_LOSS_RULSIF, "Unknown loss_type");
  RGFM_CUDA(cudaSetDevice(net->device));
  net->net->grad_log_ratio((cudaStream_t)stream, batch, x, y, loss_type, grad_x, grad_y, log_ratio);
  RGFM_API_END
}

int rgfm_classifier_logits(rgfm_net* net, const float* x, int32_t batch, float* logits, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(net && net->net->kind() >= RGFM_NET_CLF_MNIST28 && net->net->kind() <= RGFM_NET_CLF_SVHN,
             "rgfm_classifier_logits: not a classifier");
  RGFM_CHECK(x && logits && batch >= 1, "rgfm_classifier_logits: null argument");
  RGFM_CUDA(cudaSetDevice(net->device));
  net->net->forward((cudaStream_t)stream, batch, x, nullptr, logits, 0, 0.f, nullptr);
  RGFM_API_END
}

void rgfm_net_destroy(rgfm_net* net) {
  DeviceRestore _restore;
  if (net) cudaSetDevice(net->device);
  if (net && net->step) cudaFree(net->step);
  delete net;
}

int rgfm_mc_feng_step(int32_t batch, int32_t n_mc, int32_t dx, int32_t dy, const float* x, const float* y,
                      const float* v_x, const float* v_y, const float* mc_x, const float* mc_y,
                      const float* mc_ratios, double t, double dt, double gamma, float* x_out, float* y_out,
                      float* g_x, float* g_y, float* weights, void* stream) {
  RGFM_API_BEGIN
  using namespace rgfm;
  RGFM_CHECK(batch >= 1 && n_mc >= 1, "rgfm_mc_feng_step: sizes");
  RGFM_CHECK(x && y && v_x && v_y && mc_x && mc_y && mc_ratios && x_out && y_out, "rgfm_mc_feng_step: null argument");
  set_device_of(x);  // the scratch below must live on the GPU that holds the state
  const double eps = 1e-3, sigma = 1.0 - t + eps;
  StepScalars h{};
  h.t = (float)t; h.sig2 = (float)(sigma * sigma); h.sigma = (float)sigma; h.inv_sigma = (float)(1.0 / sigma);
  h.one_minus_gamma = (float)(1.0 - gamma); h.gamma = (float)gamma; h.dt = (float)dt;
  McfArgs a{};
  a.B = batch; a.N = n_mc; a.Dx = dx; a.Dy = dy;
  a.x = x; a.y = y; a.vx = v_x; a.vy = v_y; a.mcx = mc_x; a.mcy = mc_y; a.ratios = mc_ratios;
  a.x_out = x_out; a.y_out = y_out; a.gx = g_x; a.gy = g_y; a.w = weights;
  a.g_step = -1;  // always write the optional g outputs
  const size_t nf = mcf_scratch_floats(batch, n_mc, dx, dy, &a.ksplit_x, &a.ksplit_y);
  float* scratch = nullptr;
  StepScalars* dsc = nullptr;
  RGFM_CUDA(cudaMalloc(&scratch, nf * sizeof(float)));
  RGFM_CUDA(cudaMalloc(&dsc, sizeof(StepScalars)));
  RGFM_CUDA(cudaMemcpy(dsc, &h, sizeof(h), cudaMemcpyHostToDevice));
  a.sc = dsc; a.step = nullptr;
  a.logits = scratch;
  a.wbuf = scratch + (size_t)(a.ksplit_x + a.ksplit_y) * batch * n_mc;
  a.zbuf = a.wbuf + (size_t)batch * n_mc;
  try {
    launch_mcf(a, (cudaStream_t)stream);
    RGFM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  } catch (...) {
    cudaFree(scratch);
    cudaFree(dsc);
    throw;
  }
  cudaFree(scratch);
  cudaFree(dsc);
  RGFM_API_END
}

int rgfm_ratio_trainer_create(int32_t net_kind, const rgfm_weights* params, const rgfm_weights* grads, int32_t max_batch,
                              int32_t device, rgfm_ratio_trainer** out) {
  RGFM_API_BEGIN
  RGFM_CHECK(out != nullptr && params != nullptr && grads != nullptr, "rgfm_ratio_trainer_create: null argument");
  *out = rgfm::ratio_trainer_create(net_kind, params, grads, max_batch, device);
  RGFM_API_END
}

int rgfm_ratio_trainer_step(rgfm_ratio_trainer* t, const float* x, const float* y, const float* is_real, int32_t batch,
                            int32_t loss_type, float lambda_penalty, const float* const* dropout_masks, int32_t n_masks,
                            float* out8, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(t && x && y && is_real && out8 && batch >= 1, "rgfm_ratio_trainer_step: null argument");
  RGFM_CHECK(loss_type == RGFM_LOSS_DISC || loss_type == RGFM_LOSS_RULSIF, "Unknown loss_type");
  RGFM_CHECK(n_masks == 0 || dropout_masks != nullptr, "rgfm_ratio_trainer_step: null dropout_masks");
  for (int i = 0; i < n_masks; ++i) RGFM_CHECK(dropout_masks[i] != nullptr, "rgfm_ratio_trainer_step: null dropout mask");
  rgfm::ratio_trainer_step(t, x, y, is_real, batch, loss_type, lambda_penalty, dropout_masks, n_masks, out8, (cudaStream_t)stream);
  RGFM_API_END
}

void rgfm_ratio_trainer_destroy(rgfm_ratio_trainer* t) {
  DeviceRestore _restore;
  rgfm::ratio_trainer_destroy(t);
}

int rgfm_ratio_loss(const float* scores, const float* is_real, int32_t n, int32_t loss_type, float lambda_penalty, float* out8,
                    float* dscore, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(scores && is_real && out8 && dscore && n >= 1, "rgfm_ratio_loss: null argument");
  RGFM_CHECK(loss_type == RGFM_LOSS_DISC || loss_type == RGFM_LOSS_RULSIF, "Unknown loss_type");
  set_device_of(scores);
  rgfm::launch_ratio_loss(scores, is_real, n, loss_type, lambda_penalty, out8, dscore, (cudaStream_t)stream);
  RGFM_API_END
}

int rgfm_cfm_add_noise(const float* x1, const float* x0, const float* t, int32_t batch, int64_t dim, float* x_t, float* u_t,
                       void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(x1 && x0 && t && x_t && u_t && batch >= 1 && dim >= 1, "rgfm_cfm_add_noise: null argument");
  set_device_of(x1);
  rgfm::launch_cfm_interp(x1, x0, t, batch, (size_t)dim, x_t, u_t, (cudaStream_t)stream);
  RGFM_API_END
}

int rgfm_fm_mse_loss(const float* v, const float* u, int64_t n, float* loss, float* dv, float* scratch, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(v && u && loss && scratch && n >= 1, "rgfm_fm_mse_loss: null argument");
  set_device_of(v);
  rgfm::launch_fm_mse(v, u, (size_t)n, loss, dv, scratch, (cudaStream_t)stream);
  RGFM_API_END
}

int rgfm_conv3x3_forward(const float* x, int32_t n, int32_t H, int32_t W, int32_t Cin, const float* w, const float* bias, int32_t Cout,
                         int32_t stride, float* y, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(x && w && y, "rgfm_conv3x3_forward: null argument");
  set_device_of(x);
  rgfm::conv3x3_forward(x, n, H, W, Cin, w, bias, Cout, stride, y, (cudaStream_t)stream);
  RGFM_API_END
}

int rgfm_conv3x3_backward(const float* x, const float* dy, int32_t n, int32_t H, int32_t W, int32_t Cin, const float* w, int32_t Cout,
                          int32_t stride, float* dx, float* dw, float* db, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(dy, "rgfm_conv3x3_backward: null argument");
  set_device_of(dy);
  rgfm::conv3x3_backward(x, dy, n, H, W, Cin, w, Cout, stride, dx, dw, db, (cudaStream_t)stream);
  RGFM_API_END
}

int rgfm_groupnorm_nhwc_forward(const float* x, int32_t n, int32_t hw, int32_t C, int32_t groups, const float* gamma, const float* beta,
                                float eps, float* y, float* stats, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(x && gamma && beta && y && stats, "rgfm_groupnorm_nhwc_forward: null argument");
  set_device_of(x);
  rgfm::gn_nhwc_forward(x, n, hw, C, groups, gamma, beta, eps, y, stats, (cudaStream_t)stream);
  RGFM_API_END
}

int rgfm_groupnorm_nhwc_backward(const float* x, const float* dy, int32_t n, int32_t hw, int32_t C, int32_t groups, const float* gamma,
                                 const float* stats, float* dx, float* dgamma, float* dbeta, void* stream) {
  RGFM_API_BEGIN
  RGFM_CHECK(x && dy && gamma && stats, "rgfm_groupnorm_nhwc_backward: null argument");
  set_device_of(x);
  rgfm::gn_nhwc_backward(x, dy, n, hw, C, groups, gamma, stats, dx, dgamma, dbeta, (cudaStream_t)stream);
  RGFM_API_END
}

}  // extern "C"

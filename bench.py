#!/usr/bin/env python
"""bench.py -- guided coherent pairs/sec of the ratio-guided coupled Euler sampler on B200.

One "step" = one complete sampler call (MC pre-pass + ratio forward + 100 guided Euler steps) on
BASELINE.json configs[1]: MNIST32<->SVHN, mc_feng gamma=0.5, batch 1024 per GPU, N_mc=256,
RatioEstimatorMNISTSVHN(disc), synthetic seeded weights and noise.

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path (librgfm.so)
  python bench.py --impl reference ...                   # the UNMODIFIED reference on the host CPU (oracle/_ref)

N>1 is launched by torchrun (one rank per GPU); the batch is sharded (1024 pairs per rank, weak scaling), finished
pairs are all-gathered with NCCL inside the timed region, time is the max over ranks.  After the timed region the line
also carries the other BASELINE configs (`configs`): [0] 28x28 B=64, [2] the B=4096 guidance sweep with classifier
coherence, and at N>1 the strong-scaling legs [3] (65536 pairs / N) and [4] (16384 rulsif pairs / N).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

PAIR = "mnist_svhn"
B_PER_GPU = 1024
N_MC = 256
EULER_STEPS = 100
GAMMA = 0.5
SEEDS = (12, 13, 22)
# exact forward FLOPs per sample (hooked from the reference modules)
FLOPS_UNET_X, FLOPS_UNET_Y, FLOPS_RATIO = 675_086_336, 2_828_861_440, 368_968_192
METRIC = "guided coherent pairs/sec (100 Euler steps)"


def workload(prec, world):
    """The `config` both arms report (the reference arm runs the same workload on the host CPU)."""
    return dict(workload=f"BASELINE configs[1]: MNIST32-SVHN mc_feng gamma={GAMMA}, B={B_PER_GPU}/GPU, N_mc={N_MC}, "
                         f"{EULER_STEPS} Euler steps, RatioEstimatorMNISTSVHN(disc), incl. MC pre-pass",
                precision=prec, l2="256 MiB flush between timed iterations", sharding=f"batch x{world}, all_gather of pairs")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=5)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(self.rows))


def make_inputs(B_total, shape_x=(1, 32, 32), shape_y=(3, 32, 32), n_mc=N_MC, seed=42):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(B_total, *shape_x, generator=g)
    y0 = torch.randn(B_total, *shape_y, generator=g)
    mx0 = torch.randn(n_mc, *shape_x, generator=g)
    my0 = torch.randn(n_mc, *shape_y, generator=g)
    return x0, y0, mx0, my0


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules and sampler (oracle/_ref, a verbatim copy of /root/reference/src made by
# oracle/build_ref.py) on the host cores; the oracle port only when no reference tree travelled with the snapshot
# ------------------------------------------------------------------------------------------------
class CpuArm:
    """Times the reference on a BOUNDED sample of configs[1] and composes the full configuration from it.

    The sampler's cost is a sum of independent per-image pieces (flow_utils.py:231-373 / sample_mnist_svhn.py:85-175):
      a = one Euler step of both U-Nets per image     (the reference's own nn.Modules, fm(x, t), torch.no_grad)
      b = one ratio-estimator log_ratio per image     (RatioEstimatorMNISTSVHN.log_ratio)
      g = one mc_feng guidance block per pair against the full N_mc bank (the verbatim broadcast arithmetic)
      T(B, N_mc, S) = S * N_mc * a + N_mc * b + S * B * a + (S - 1) * B * g
    validate() runs the UNMODIFIED sample_bimodal_guided_mnist_svhn end to end on a small configuration and reports the
    composition's error on it, so the extrapolation is checked against a real full call in the same run."""

    def __init__(self):
        from oracle import rgfm_oracle as O  # checker-side code: only the CPU legs of bench.py execute it
        from oracle import ref_loader
        from ratio_guided_multimodal_fm_b200 import synthetic as S
        self.O, self.S = O, S
        torch.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread
        self.cores = torch.get_num_threads()
        self.sd = [S.state_dict("unet_mnist32", SEEDS[0]), S.state_dict("unet_svhn", SEEDS[1]), S.state_dict("ratio_mnist_svhn", SEEDS[2])]
        self.ns = ref_loader.load()
        self.kind = "reference" if self.ns is not None else "port"
        if self.ns is not None:
            self.fx = ref_loader.module(self.ns, "unet_mnist32", self.sd[0])
            self.fy = ref_loader.module(self.ns, "unet_svhn", self.sd[1])
            self.fr = ref_loader.module(self.ns, "ratio_mnist_svhn", self.sd[2], "disc")
        g = torch.Generator().manual_seed(1)
        self.nb, self.bs = 32, 16  # sample sizes: images per U-Net / ratio timing, pairs per guidance timing
        self.mx = torch.randn(self.nb, 1, 32, 32, generator=g)
        self.my = torch.randn(self.nb, 3, 32, 32, generator=g)
        self.bank = (torch.randn(N_MC, 1, 32, 32, generator=g), torch.randn(N_MC, 3, 32, 32, generator=g), torch.rand(N_MC, generator=g) + 0.5)
        self.x = torch.randn(self.bs, 1, 32, 32, generator=g)
        self.y = torch.randn(self.bs, 3, 32, 32, generator=g)
        self.pieces()  # warm the thread pool and the allocator

    @torch.no_grad()
    def unets(self, x, y, t):
        tb = torch.full((x.shape[0],), t)
        if self.ns is not None:
            return self.fx(x, tb), self.fy(y, tb)
        return self.O.unet_forward(self.sd[0], "unet_mnist32", x, tb), self.O.unet_forward(self.sd[1], "unet_svhn", y, tb)

    @torch.no_grad()
    def pieces(self, n_mc=N_MC):
        t0 = time.perf_counter()
        self.unets(self.mx, self.my, 0.5)
        a = (time.perf_counter() - t0) / self.nb
        t0 = time.perf_counter()
        if self.ns is not None:
            self.fr.log_ratio(self.mx, self.my)
        else:
            self.O.ratio_log_ratio(self.sd[2], "ratio_mnist_svhn", self.mx, self.my, "disc")
        b = (time.perf_counter() - t0) / self.nb
        t0 = time.perf_counter()
        self.O.mc_feng_guidance(self.x, self.y, self.bank[0][:n_mc], self.bank[1][:n_mc], self.bank[2][:n_mc], 0.5)
        g = (time.perf_counter() - t0) / self.bs
        return a, b, g

    @staticmethod
    def compose(a, b, g, B, n_mc, steps):
        return steps * n_mc * a + n_mc * b + steps * B * a + (steps - 1) * B * g

    def sample(self):
        """One bounded sample -> (pairs/s composed to configs[1], wall seconds spent, pieces)."""
        t0 = time.perf_counter()
        reps = [self.pieces() for _ in range(3)]
        a, b, g = (float(np.median([r[i] for r in reps])) for i in range(3))
        wall = time.perf_counter() - t0
        return B_PER_GPU / self.compose(a, b, g, B_PER_GPU, N_MC, EULER_STEPS), wall, dict(unets_img_step_s=a, ratio_img_s=b, mc_feng_pair_step_s=g)

    def describe(self):
        who = "the reference's own nn.Modules (oracle/_ref)" if self.ns is not None else "the oracle port"
        return (f"median of 3: both U-Nets on {self.nb} images and the ratio estimator on {self.nb} images with {who}, the mc_feng block on "
                f"{self.bs} pairs vs the full N_mc={N_MC} bank; composed to B={B_PER_GPU}, N_mc={N_MC}, {EULER_STEPS} steps")

    @torch.no_grad()
    def validate(self):
        """The UNMODIFIED reference sampler end to end on a small configuration vs the composition's prediction for it."""
        if self.ns is None:
            return None
        B, n_mc, steps = 16, 32, 3
        a, b, g = self.pieces(n_mc)
        pred = self.compose(a, b, g, B, n_mc, steps)
        self.ns.set_seed(0)
        import contextlib
        import io
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            x1, _ = self.ns.sample_bimodal_guided_mnist_svhn(self.fx, self.fy, self.fr, guidance_method="mc_feng", guidance_strength=GAMMA,
                                                             num_samples=B, num_steps=steps, device="cpu", mc_batch_size=n_mc)
        real = time.perf_counter() - t0
        return dict(config=f"sample_bimodal_guided_mnist_svhn(B={B}, N_mc={n_mc}, steps={steps}) on the host CPU, unmodified reference",
                    measured_s=real, composed_s=pred, composition_error=(pred - real) / real, finite=bool(torch.isfinite(x1).all()))

    def train_step(self, dev=None):
        """The UNMODIFIED reference trainer (src/utils/trainer.py RatioTrainer + src/utils/losses.py) for the training-step
        numbers in `configs.train_step`: same batch and optimiser, on the host cores and -- as the fair comparator of the
        GPU path -- with stock PyTorch on the same GPU."""
        if self.ns is None:
            return None
        from src.utils.trainer import RatioTrainer as RefTrainer
        from src.utils.losses import get_ratio_loss
        res = {}
        for kind, seed, loss in (("ratio28", 21, "disc"), ("ratio_mnist_svhn", 22, "rulsif")):
            res[kind] = {}
            for where in (["cpu"] + ([str(dev)] if dev is not None else [])):
                m = self.ns.classes[kind](loss)
                m.load_state_dict(self.S.state_dict(kind, seed), strict=True)
                m = m.to(where).train()
                tr = RefTrainer(m, get_ratio_loss(loss), torch.optim.Adam(m.parameters(), lr=1e-4), where)
                b = train_batch(kind)
                warm, reps = (1, 2) if where == "cpu" else (3, 10)
                for _ in range(warm):
                    tr.train_step(b)
                t0 = time.perf_counter()
                for _ in range(reps):
                    tr.train_step(b)  # its metrics' .item() calls synchronise
                res[kind]["cpu_ms" if where == "cpu" else "same_gpu_stock_pytorch_ms"] = 1e3 * (time.perf_counter() - t0) / reps
        res["note"] = ("unmodified reference RatioTrainer.train_step, batch 128, Adam; BatchNorm in TRAINING mode (batch statistics) "
                       "as the reference trains it -- the library step freezes them; cuDNN may use TF32 for the GPU comparator")
        return res

    @torch.no_grad()
    def config0_full(self):
        """BASELINE configs[0] (28x28, mc_feng, B=64, N_mc=128, 100 steps, disc) executed ONCE in full by the unmodified
        reference sampler: a genuinely measured CPU pairs/s next to the composed configs[1] estimate."""
        if self.ns is None:
            return None
        from oracle import ref_loader
        S = self.S
        fx = ref_loader.module(self.ns, "unet_mnist28", S.state_dict("unet_mnist28", 11))
        fy = ref_loader.module(self.ns, "unet_mnist28", S.state_dict("unet_mnist28", 14))
        fr = ref_loader.module(self.ns, "ratio28", S.state_dict("ratio28", 21), "disc")
        self.ns.set_seed(42)
        import contextlib
        import io
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            x1, _ = self.ns.sample_bimodal_guided(fx, fy, fr, guidance_method="mc_feng", guidance_strength=0.5, num_samples=64,
                                                  num_steps=100, device="cpu", mc_batch_size=128)
        s = time.perf_counter() - t0
        return dict(config="BASELINE configs[0]: sample_bimodal_guided 28x28 mc_feng gamma=0.5 B=64 N_mc=128 100 steps, unmodified reference",
                    seconds=s, pairs_per_s=64.0 / s, finite=bool(torch.isfinite(x1).all()))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.manual_seed(0)
    arm = CpuArm()
    vals, walls, parts = [], [], None
    for i in range(args.warmup + args.steps):
        v, wall, parts = arm.sample()
        if i >= args.warmup:
            vals.append(v)
            walls.append(wall)
    value = float(np.median(vals))
    line = dict(impl="reference", metric=METRIC, value=value, unit="pairs/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * float(np.mean(walls)), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", config=workload(args.precision, max(1, args.gpus)),
                cpu_baseline=dict(value=value, unit="pairs/s", cores=arm.cores, kind=arm.kind, sample=arm.describe(),
                                  spread=dict(min=float(np.min(vals)), max=float(np.max(vals)), n=len(vals)), pieces=parts,
                                  full_config_seconds_composed=B_PER_GPU / value),
                e2e=dict(value=value, unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="ms_per_step is the wall time of one bounded sample (what was executed); value is the composed full-configuration throughput",
                host=dict(os_cpu_count=os.cpu_count(), torch_threads=arm.cores))
    line["validation"] = arm.validate()
    if not args.skip_config0:
        line["config0_full"] = arm.config0_full()
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def _timed(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps, out


TRAIN_B = 128  # --batch_size default of src/train_ratio.py:32 and src/train_ratio_mnist_svhn.py:85


def train_batch(kind, B=TRAIN_B, seed=5):
    g = torch.Generator().manual_seed(seed)
    sy = (1, 28, 28) if kind == "ratio28" else (3, 32, 32)
    hw = 28 if kind == "ratio28" else 32
    return dict(x=torch.randn(B, 1, hw, hw, generator=g), y=torch.randn(B, *sy, generator=g), is_real=(torch.arange(B) % 2 == 0).float())


def train_step_numbers(dev):
    """Training-side reuse: RatioTrainer.train_step (src/utils/trainer.py:26-56) on the library's fp32
    kernels, batch 128, Adam, gradient clipping, dropout on.  `e2e_ms`: through the drop-in trainer with a HOST batch (H2D
    of x, y, is_real, the step, clipping + Adam in torch, metrics D2H); `device_ms`: the C-ABI step alone on resident data."""
    from ratio_guided_multimodal_fm_b200 import RatioTrainer, TrainableModule, synthetic as S
    res = dict(workload=f"RatioTrainer.train_step, batch {TRAIN_B}, Adam lr 1e-4, clip_grad_norm 1.0, dropout 0.1 (masks drawn on the device), "
                        "fp32 CUDA-core kernels, frozen BatchNorm statistics")
    for kind, seed, loss in (("ratio28", 21, "disc"), ("ratio_mnist_svhn", 22, "rulsif")):
        model = TrainableModule(kind, S.state_dict(kind, seed), loss).to(dev)
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)

        class _L:
            pass
        lf = _L()
        if loss == "rulsif":
            lf.lambda_penalty = 0.1
        tr = RatioTrainer(model, lf, opt, device=dev, max_batch=TRAIN_B)
        hb = {k: v.pin_memory() for k, v in train_batch(kind).items()}
        for _ in range(3):
            m = tr.train_step(hb)
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            m = tr.train_step(hb)  # ends with the metrics' D2H (a synchronisation)
        e2e = (time.perf_counter() - t0) / reps
        db = {k: v.to(dev) for k, v in hb.items()}
        plan = tr._plan
        masks = [torch.ones(TRAIN_B, w, device=dev) for w in ((512, 256) if kind == "ratio28" else (512, 512))]
        s, _ = _timed(lambda: plan.step(db["x"], db["y"], db["is_real"], loss, 0.1, masks), reps=10)
        res[kind] = dict(loss_type=loss, e2e_ms=1e3 * e2e, device_ms=1e3 * s, pairs_per_s=TRAIN_B / e2e, loss=m["loss"],
                         finite=bool(np.isfinite(m["loss"])))
        plan.close()
    res["flow_matching"] = fm_step_numbers(dev)
    return res


def fm_step_numbers(dev):
    """Flow-matching step of the U-Nets (flow_utils.py:127-153: t, add_noise, model, MSE, backward, Adam), batch 128, dropout
    on: every 3x3 convolution (forward, input / weight / bias gradient) on the library's exact-fp32 kernels
    (fm_train.fm_step), beside the same step in stock PyTorch on the same GPU (cuDNN, TF32 allowed = its default, and off)."""
    import copy
    import torch.nn.functional as F
    from oracle import ref_loader  # the subject of both arms is the reference's OWN module when its tree travelled (oracle/_ref)
    from ratio_guided_multimodal_fm_b200 import CFMSchedule, synthetic as S
    from ratio_guided_multimodal_fm_b200.fm_train import fm_step
    ns = ref_loader.load()

    def make(kind):
        if ns is not None:
            return ns.classes[kind](), "reference class (oracle/_ref)"
        from oracle.unet_module import UNet  # the same module tree restated (test infrastructure)
        return UNet(kind), "oracle.unet_module restatement"
    sched = CFMSchedule()
    out = dict(workload=f"flow-matching step, batch {TRAIN_B}, Adam lr 1e-4, dropout 0.1; 3x3 convs on the library (exact fp32 CUDA-core "
                        "kernels), GroupNorm on the library in NHWC, SiLU / 1x1 skips / time MLP stock PyTorch")
    for kind in ("unet_mnist32", "unet_svhn"):
        c = 3 if kind == "unet_svhn" else 1
        base, what = make(kind)
        base.load_state_dict(S.state_dict(kind, 12), strict=True)
        x_1 = torch.randn(TRAIN_B, c, 32, 32, device=dev)
        r = {"module": what}
        tf32 = torch.backends.cudnn.allow_tf32
        try:
            for name in ("library_ms", "same_gpu_stock_pytorch_ms", "same_gpu_stock_pytorch_tf32_off_ms"):
                m = copy.deepcopy(base).to(dev).train()
                opt = torch.optim.Adam(m.parameters(), lr=1e-4)
                torch.backends.cudnn.allow_tf32 = name != "same_gpu_stock_pytorch_tf32_off_ms"

                def step():
                    opt.zero_grad()
                    if name == "library_ms":
                        loss = fm_step(m, x_1, sched)
                    else:
                        t = torch.rand(TRAIN_B, device=dev)
                        x_t, u = sched.add_noise(x_1, t)
                        loss = F.mse_loss(m(x_t, t), u)
                        loss.backward()
                    opt.step()
                    return loss
                for _ in range(3):
                    loss = step()
                sec, _ = _timed(step, reps=5)
                r[name] = 1e3 * sec
                if name == "library_ms":
                    r["loss"] = float(loss.item())
                    r["finite"] = bool(np.isfinite(r["loss"]))
                    r["images_per_s"] = TRAIN_B / sec
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
        out[kind] = r
    return out


def other_configs(dev, world, rank, prec, dist, gather_pairs):
    """The BASELINE configs bench.py's timed region does not cover; every number is one warm call, CUDA-event timed, on
    this rank's shard (max over ranks for the multi-GPU legs)."""
    from ratio_guided_multimodal_fm_b200 import _lib, synthetic as S, evaluate_coherence_mnist_svhn
    out = {}

    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sd28 = [S.state_dict("unet_mnist28", 11), S.state_dict("unet_mnist28", 14), S.state_dict("ratio28", 21)]
    sdms = [S.state_dict("unet_mnist32", SEEDS[0]), S.state_dict("unet_svhn", SEEDS[1]), S.state_dict("ratio_mnist_svhn", SEEDS[2])]
    if rank == 0:
        # configs[0]: the reference's CPU-runnable case, here on one GPU
        B, n_mc = 64, 128
        x0, y0, m0, n0 = (t.to(dev) for t in make_inputs(B, (1, 28, 28), (1, 28, 28), n_mc, seed=7))
        t0 = time.perf_counter()
        plan = _lib.SamplerPlan("pair28", *sd28, guidance_method="mc_feng", guidance_strength=0.5, num_samples=B, num_steps=EULER_STEPS,
                                mc_batch_size=n_mc, loss_type="disc", precision=prec, device=dev)
        build_s = time.perf_counter() - t0
        plan.run(x0, y0, m0, n0)
        s, _ = _timed(lambda: plan.run(x0, y0, m0, n0), reps=3)
        plan.close()
        out["configs[0]"] = dict(workload="28x28 pair, mc_feng gamma=0.5, B=64, N_mc=128, 100 steps, disc", pairs_per_s=B / s, ms_per_call=1e3 * s,
                                 plan_build_s=build_s)
        # configs[2]: guidance-strength sweep at B=4096 with classifier pair coherence (one MC bank for the whole sweep)
        B, gammas = 4096, [0.0, 0.5, 1.0, 2.0, 5.0]
        cm = S.StateDictModule("clf_mnist32", S.state_dict("clf_mnist32", 31))
        cs = S.StateDictModule("clf_svhn", S.state_dict("clf_svhn", 32))
        x0, y0, m0, n0 = (t.to(dev) for t in make_inputs(B, seed=8))
        plan = _lib.SamplerPlan(PAIR, *sdms, guidance_method="mc_feng", guidance_strength=gammas[0], num_samples=B, num_steps=EULER_STEPS,
                                mc_batch_size=N_MC, loss_type="disc", precision=prec, device=dev)
        plan.run(x0, y0, m0, n0)  # warm
        evaluate_coherence_mnist_svhn(x0[:64], y0[:64], cm, cs, dev)  # builds the classifier nets
        # the evaluator is fused into the sampler's tail: classifier forwards + argmax match on the device, count in diag[15]
        from ratio_guided_multimodal_fm_b200 import evaluate as _ev
        plan.attach_evaluators(_ev._net(cm, x0.shape, dev), _ev._net(cs, y0.shape, dev))
        coh, dg = {}, {}

        def sweep():
            for i, gm in enumerate(gammas):
                plan.set_strength(gm)
                r = plan.run(x0, y0, m0, n0, want_diag=True) if i == 0 else plan.run(x0, y0, reuse_bank=True, want_diag=True)
                dg[str(gm)] = r[2]
            for k, v in dg.items():  # one scalar per gamma crosses to the host
                coh[k] = float(v[15]) / B
        s, _ = _timed(sweep)
        plan.close()
        out["configs[2]"] = dict(workload=f"MNIST32-SVHN gamma sweep {gammas}, B=4096 per gamma, N_mc=256, 100 steps, shared MC bank, "
                                          "MNISTClassifier32 / SVHNClassifier pair coherence (synthetic untrained classifiers)",
                                 pairs_per_s=B * len(gammas) / s, s_per_sweep=s, coherence_acc=coh)
    if rank == 0:
        out["train_step"] = train_step_numbers(dev)
    if world > 1:
        # configs[3]: 65536 pairs sharded over the ranks (strong scaling), NCCL all-gather of the pairs inside the timed call
        total = 65536
        nl = total // world
        x0, y0, m0, n0 = make_inputs(nl, seed=100 + rank)
        m0, n0 = make_inputs(1, seed=9)[2:]  # every rank integrates the same MC bank noise
        x0, y0, m0, n0 = (t.to(dev) for t in (x0, y0, m0, n0))
        plan = _lib.SamplerPlan(PAIR, *sdms, guidance_method="mc_feng", guidance_strength=GAMMA, num_samples=nl, num_steps=EULER_STEPS,
                                mc_batch_size=N_MC, loss_type="disc", precision=prec, device=dev)

        def call3():
            x1, y1 = plan.run(x0, y0, m0, n0)
            return gather_pairs(x1, y1, total)
        dist.barrier()
        s, _ = _timed(call3)
        s = maxr(s)
        plan.close()
        if rank == 0:
            out["configs[3]"] = dict(workload=f"MNIST32-SVHN mc_feng gamma={GAMMA}, 65536 pairs / {world} GPUs (strong scaling), all_gather of pairs",
                                     pairs_per_s=total / s, s_per_call=s, pairs_per_gpu=nl)
        # configs[4]: 28x28, RuLSIF-trained estimator head, 16384 pairs over the ranks
        total = 16384
        nl = total // world
        x0, y0, _, _ = make_inputs(nl, (1, 28, 28), (1, 28, 28), 128, seed=200 + rank)
        m0, n0 = make_inputs(1, (1, 28, 28), (1, 28, 28), 128, seed=10)[2:]
        x0, y0, m0, n0 = (t.to(dev) for t in (x0, y0, m0, n0))
        plan = _lib.SamplerPlan("pair28", *sd28, guidance_method="mc_feng", guidance_strength=0.5, num_samples=nl, num_steps=EULER_STEPS,
                                mc_batch_size=128, loss_type="rulsif", precision=prec, device=dev)
        plan.run(x0, y0, m0, n0)

        def call4():
            x1, y1 = plan.run(x0, y0, m0, n0)
            return gather_pairs(x1, y1, total)
        dist.barrier()
        s, _ = _timed(call4)
        s = maxr(s)
        plan.close()
        if rank == 0:
            out["configs[4]"] = dict(workload=f"28x28 pair, mc_feng gamma=0.5, RuLSIF estimator, 16384 pairs / {world} GPUs (strong scaling), 100 steps",
                                     pairs_per_s=total / s, s_per_call=s, pairs_per_gpu=nl)
    return out


def run_gpu(args):
    import torch.distributed as dist
    from ratio_guided_multimodal_fm_b200 import _lib, synthetic as S
    from ratio_guided_multimodal_fm_b200.sampler import gather_pairs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    prec = args.precision
    B_total = B_PER_GPU * world
    sdx, sdy, sdr = S.state_dict("unet_mnist32", SEEDS[0]), S.state_dict("unet_svhn", SEEDS[1]), S.state_dict("ratio_mnist_svhn", SEEDS[2])
    t0 = time.perf_counter()
    plan = _lib.SamplerPlan(PAIR, sdx, sdy, sdr, guidance_method="mc_feng", guidance_strength=GAMMA,
                            num_samples=B_PER_GPU, num_steps=EULER_STEPS, mc_batch_size=N_MC, loss_type="disc",
                            precision=prec, device=dev, chunk=args.chunk)
    plan_build_s = time.perf_counter() - t0
    # every rank draws the full batch with the same seed and takes its rows
    x0, y0, mx0, my0 = make_inputs(B_total)
    lo = rank * B_PER_GPU
    hx0, hy0 = x0[lo:lo + B_PER_GPU].contiguous().numpy(), y0[lo:lo + B_PER_GPU].contiguous().numpy()
    hmx0, hmy0 = mx0.numpy(), my0.numpy()
    dx0, dy0, dmx0, dmy0 = (torch.from_numpy(a).to(dev) for a in (hx0, hy0, hmx0, hmy0))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        x1, y1 = plan.run(dx0, dy0, dmx0, dmy0)
        if world > 1:
            x1, y1 = gather_pairs(x1, y1, B_total)
        return x1, y1

    for _ in range(args.warmup):
        one_step()
        flush.zero_()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        x1, y1 = one_step()
        flush.zero_()  # L2 flush between timed iterations
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.finish() if clocks else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = B_total / (ms_per_step * 1e-3)
    finite = bool(torch.isfinite(x1).all() and torch.isfinite(y1).all())

    # end to end through the C-ABI host-buffer call: H2D of the noise + run + D2H of the pairs inside the timed region
    plan.run_host(hx0, hy0, hmx0, hmy0)  # warm-up: the first host call allocates the plan's pinned staging buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        _, _, hdiag = plan.run_host(hx0, hy0, hmx0, hmy0)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / max(1, args.steps)
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = 4 * (hx0.size + hy0.size + hmx0.size + hmy0.size)
    d2h = 4 * (hx0.size + hy0.size) + 4 * _lib.DIAG_FLOATS

    # main loop only (the MC bank of the previous call is reused): the pre-pass / main-loop split of one call
    main_s, _ = _timed(lambda: plan.run(dx0, dy0, reuse_bank=True))
    full_s, _ = _timed(lambda: plan.run(dx0, dy0, dmx0, dmy0))

    extra = {} if args.skip_configs else other_configs(dev, world, rank, prec, dist, gather_pairs)

    if rank == 0:
        pk = peaks()
        # Per-kernel CUDA-event profile of one eager Euler step, taken three times (each after a warm main loop) and the
        # least-disturbed one kept (smallest conv total): a single pass right after the other legs of this script has read
        # 7 % high on one box in five (clock / power state left by the preceding work), the graph step and the throughput not.
        profs = []
        for _ in range(3):
            plan.run(dx0, dy0, reuse_bank=True)
            torch.cuda.synchronize()
            profs.append(plan.profile_step())
        conv_total = lambda pr: sum(v["ms"] for k, v in pr.items() if k.startswith("conv"))
        prof = min(profs, key=conv_total)
        profile_runs = [round(conv_total(pr), 4) for pr in profs]
        graph_step_ms = prof.pop("graph_step", {"ms": None})["ms"]
        step_ms = sum(v["ms"] for v in prof.values())
        conv_kinds = [k for k in prof if k.startswith("conv")]
        top = max(conv_kinds, key=lambda k: prof[k]["ms"])
        cv = prof[top]
        ach = cv["flops"] / (cv["ms"] * 1e-3) / 1e12
        # the fused guidance-sum + blend + Euler update has a sparse (saturated weights) and a dense path selected on the
        # device; exactly one does the work of a launch -- the other returns at once -- so take the one that took longer
        upd_name = max(("mcf_apply", "mcf_apply_dense"), key=lambda k: prof.get(k, {"ms": 0.0})["ms"])
        upd = prof.get(upd_name)
        # DRAM traffic of the dominant kernel: per launch, from the committed ncu --set full capture of one launch
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = {k: tj[k] for k in tj if k != "dram_bytes_per_launch"}
        roof = dict(bound="tensor", kernel=top, achieved=ach, peak=pk["tc_sustained"], unit="TFLOP/s",
                    frac=ach / pk["tc_sustained"], traffic=traffic, traffic_source=traffic_note,
                    peak_source=pk["src"] + " bf16 sustained (kernel timed inside a long step)",
                    launches_per_euler_step=cv["launches"], avg_launch_ms=cv["ms"] / cv["launches"],
                    share_of_step=cv["ms"] / step_ms)
        if upd:
            ua = upd["bytes"] / (upd["ms"] * 1e-3) / 1e9
            roof["update_kernel"] = dict(bound="hbm", kernel=upd_name + ("_sparse" if upd_name == "mcf_apply" else ""), achieved=ua, peak=pk["hbm"], unit="GB/s",
                                         frac=ua / pk["hbm"], share_of_step=upd["ms"] / step_ms)
        roof["per_kernel_ms_one_euler_step"] = {k: round(v["ms"], 4) for k, v in prof.items()}
        roof["graph_ms_one_euler_step"] = graph_step_ms
        roof["conv_ms_of_the_three_profile_passes"] = profile_runs  # the smallest one is reported above
        total_flops = EULER_STEPS * (N_MC + B_PER_GPU) * (FLOPS_UNET_X + FLOPS_UNET_Y) + N_MC * FLOPS_RATIO
        # the CPU baseline is an N=1 number (rank 0, all host threads); multi-GPU lines carry null
        cpu_base = None
        if world == 1 and not args.skip_cpu:
            arm = CpuArm()
            cpu_v, _, parts = arm.sample()
            cpu_base = dict(value=cpu_v, unit="pairs/s", cores=arm.cores, kind=arm.kind, sample=arm.describe(), pieces=parts,
                            validation=arm.validate())
            if not args.skip_configs:
                cpu_base["train_step_reference"] = arm.train_step(dev)
        line = dict(metric=METRIC, value=value, unit="pairs/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype={"fp32": "f32", "bf16": "bf16", "fp16": "f16"}[prec], data="synthetic",
                    config=workload(prec, world),
                    e2e=dict(value=B_total / e2e_s, unit="pairs/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=int(plan.launch_count) * args.steps, clocks=clk, roofline=roof,
                    cpu_baseline=cpu_base,
                    call_split=dict(plan_build_s=plan_build_s, full_call_ms=1e3 * full_s, main_loop_ms=1e3 * main_s,
                                    prepass_and_ratio_ms=1e3 * (full_s - main_s),
                                    main_loop_only_pairs_per_s=B_PER_GPU / main_s),
                    model_flops_utilisation=dict(algorithmic_tflop_per_call=total_flops / 1e12,
                                                 achieved_tflops=total_flops / (ms_per_step * 1e-3) / 1e12,
                                                 frac_of_peak=total_flops / (ms_per_step * 1e-3) / 1e12 / pk["tc_sustained"]),
                    configs=extra, saturated_rows=float(hdiag[12]), outputs_finite=finite)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rgfm", choices=["rgfm", "reference"])
    ap.add_argument("--precision", default=os.environ.get("RGFM_PRECISION", "fp16"))
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--skip-configs", action="store_true", help="only the timed configs[1] region (no configs[0], [2], [3], [4] legs)")
    ap.add_argument("--skip-cpu", action="store_true", help="no cpu_baseline leg")
    ap.add_argument("--skip-config0", action="store_true", help="reference arm: skip the full configs[0] call (~1 min of CPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

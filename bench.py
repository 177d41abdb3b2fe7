#!/usr/bin/env python
"""bench.py -- guided coherent pairs/sec of the ratio-guided coupled Euler sampler on B200.

One "step" = one complete sampler call (MC pre-pass + ratio forward + 100 guided Euler steps) on
BASELINE.json configs[1]: MNIST32<->SVHN, mc_feng gamma=0.5, batch 1024 per GPU, N_mc=256,
RatioEstimatorMNISTSVHN(disc), synthetic seeded weights and noise.

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path (librgfm.so)
  python bench.py --impl reference ...                   # the reference algorithm on the host CPU (oracle port)

N>1 is launched by torchrun (one rank per GPU); the batch is sharded (1024 pairs per rank, weak scaling), finished
pairs are all-gathered with NCCL inside the timed region, time is the max over ranks.
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


def make_inputs(B_total, device=None, pinned=False):
    g = torch.Generator().manual_seed(42)
    x0 = torch.randn(B_total, 1, 32, 32, generator=g)
    y0 = torch.randn(B_total, 3, 32, 32, generator=g)
    mx0 = torch.randn(N_MC, 1, 32, 32, generator=g)
    my0 = torch.randn(N_MC, 3, 32, 32, generator=g)
    return x0, y0, mx0, my0


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port of flow_utils.py / sample_mnist_svhn.py) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample_throughput(budget_s=20.0):
    """Bounded sample of the SAME workload, composed to the full configuration:
    time (a) one MC pre-pass Euler step per image for each U-Net, (b) one ratio forward per image, (c) one guided
    main-loop step per pair (both U-Nets + the verbatim broadcast mc_feng block against the full N_mc=256 bank);
    pairs/s = B / (steps*(N_mc*(a) + B*(c)) + N_mc*(b)) with B=1024, N_mc=256, steps=100."""
    from oracle import rgfm_oracle as O
    from ratio_guided_multimodal_fm_b200 import synthetic as S
    sdx, sdy, sdr = S.state_dict("unet_mnist32", SEEDS[0]), S.state_dict("unet_svhn", SEEDS[1]), S.state_dict("ratio_mnist_svhn", SEEDS[2])
    g = torch.Generator().manual_seed(1)
    nb, bs = 32, 16  # prepass sample images, main-loop sample pairs
    with torch.no_grad():
        mx = torch.randn(nb, 1, 32, 32, generator=g); my = torch.randn(nb, 3, 32, 32, generator=g)
        tb = torch.full((nb,), 0.5)
        O.unet_forward(sdx, "unet_mnist32", mx[:2], tb[:2])  # warm the thread pool
        t0 = time.perf_counter()
        O.unet_forward(sdx, "unet_mnist32", mx, tb); O.unet_forward(sdy, "unet_svhn", my, tb)
        t_pre = (time.perf_counter() - t0) / nb
        t0 = time.perf_counter()
        O.ratio_log_ratio(sdr, "ratio_mnist_svhn", mx, my, "disc")
        t_ratio = (time.perf_counter() - t0) / nb
        bank_x = torch.randn(N_MC, 1, 32, 32, generator=g); bank_y = torch.randn(N_MC, 3, 32, 32, generator=g)
        r = torch.rand(N_MC, generator=g) + 0.5
        x = torch.randn(bs, 1, 32, 32, generator=g); y = torch.randn(bs, 3, 32, 32, generator=g)
        reps, t_main = 0, 0.0
        while t_main < budget_s * 0.6 and reps < 4:
            t0 = time.perf_counter()
            vx = O.unet_forward(sdx, "unet_mnist32", x, torch.full((bs,), 0.5))
            vy = O.unet_forward(sdy, "unet_svhn", y, torch.full((bs,), 0.5))
            gx, gy, _, _ = O.mc_feng_guidance(x, y, bank_x, bank_y, r, 0.5)
            x = x + ((1 - GAMMA) * vx + GAMMA * gx) * 0.01
            y = y + ((1 - GAMMA) * vy + GAMMA * gy) * 0.01
            t_main += time.perf_counter() - t0
            reps += 1
        t_main /= reps * bs
    total = EULER_STEPS * (N_MC * t_pre + B_PER_GPU * t_main) + N_MC * t_ratio
    sample = (f"U-Net pre-pass step on {nb} images, ratio forward on {nb} images, {reps} guided main-loop step(s) on {bs} "
              f"pairs vs the full N_mc={N_MC} bank; composed to B={B_PER_GPU}, {EULER_STEPS} steps")
    return B_PER_GPU / total, sample, dict(t_prepass_img_step=t_pre, t_ratio_img=t_ratio, t_main_pair_step=t_main)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread
    cores = torch.get_num_threads()
    vals = []
    for i in range(args.warmup + args.steps):
        v, sample, parts = cpu_sample_throughput(budget_s=12.0)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = dict(impl="reference", metric=METRIC, value=value, unit="pairs/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * B_PER_GPU / value, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=f"MNIST32-SVHN mc_feng gamma={GAMMA} B={B_PER_GPU} N_mc={N_MC} {EULER_STEPS} Euler steps "
                                     "(reference algorithm on host CPU, oracle port; bounded sample composed to the full config)"),
                cpu_baseline=dict(value=value, unit="pairs/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                host=dict(os_cpu_count=os.cpu_count(), torch_threads=cores))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
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
    plan = _lib.SamplerPlan(PAIR, sdx, sdy, sdr, guidance_method="mc_feng", guidance_strength=GAMMA,
                            num_samples=B_PER_GPU, num_steps=EULER_STEPS, mc_batch_size=N_MC, loss_type="disc",
                            precision=prec, device=dev, chunk=args.chunk)
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
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        plan.run_host(hx0, hy0, hmx0, hmy0)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / max(1, args.steps)
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = 4 * (hx0.size + hy0.size + hmx0.size + hmy0.size)
    d2h = 4 * (hx0.size + hy0.size) + 4 * _lib.DIAG_FLOATS

    line = None
    if rank == 0:
        pk = peaks()
        prof = plan.profile_step()
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
        roof = dict(bound="tensor", kernel=top, achieved=ach, peak=pk["tc_sustained"], unit="TFLOP/s",
                    frac=ach / pk["tc_sustained"], traffic=None, peak_source=pk["src"] + " bf16 sustained (kernel timed inside a long step)",
                    launches_per_euler_step=cv["launches"], avg_launch_ms=cv["ms"] / cv["launches"],
                    share_of_step=cv["ms"] / step_ms)
        if upd:
            ua = upd["bytes"] / (upd["ms"] * 1e-3) / 1e9
            roof["update_kernel"] = dict(bound="hbm", kernel=upd_name + ("_sparse" if upd_name == "mcf_apply" else ""), achieved=ua, peak=pk["hbm"], unit="GB/s",
                                         frac=ua / pk["hbm"], share_of_step=upd["ms"] / step_ms)
        roof["per_kernel_ms_one_euler_step"] = {k: round(v["ms"], 4) for k, v in prof.items()}
        roof["graph_ms_one_euler_step"] = graph_step_ms
        total_flops = EULER_STEPS * (N_MC + B_PER_GPU) * (FLOPS_UNET_X + FLOPS_UNET_Y) + N_MC * FLOPS_RATIO
        # the CPU baseline is an N=1 number (rank 0, all host threads); multi-GPU lines carry null
        cpu_base = None
        if world == 1:
            torch.set_num_threads(os.cpu_count() or 1)
            cpu_v, cpu_sample, _ = cpu_sample_throughput(budget_s=15.0)
            cpu_base = dict(value=cpu_v, unit="pairs/s", cores=torch.get_num_threads(), kind="port", sample=cpu_sample)
        line = dict(metric=METRIC, value=value, unit="pairs/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype={"fp32": "f32", "bf16": "bf16", "fp16": "f16"}[prec], data="synthetic",
                    config=dict(workload=f"BASELINE configs[1]: MNIST32-SVHN mc_feng gamma={GAMMA}, B={B_PER_GPU}/GPU, N_mc={N_MC}, "
                                         f"{EULER_STEPS} Euler steps, RatioEstimatorMNISTSVHN(disc), incl. MC pre-pass",
                                precision=prec, l2="256 MiB flush between timed iterations", sharding=f"batch x{world}, all_gather of pairs"),
                    e2e=dict(value=B_total / e2e_s, unit="pairs/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=int(plan.launch_count) * args.steps, clocks=clk, roofline=roof,
                    cpu_baseline=cpu_base,
                    model_flops_utilisation=dict(algorithmic_tflop_per_call=total_flops / 1e12,
                                                 achieved_tflops=total_flops / (ms_per_step * 1e-3) / 1e12 ,
                                                 frac_of_peak=total_flops / (ms_per_step * 1e-3) / 1e12 / pk["tc_sustained"]),
                    outputs_finite=finite)
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

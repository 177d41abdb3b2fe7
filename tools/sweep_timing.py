"""Time a guidance-strength sweep at the bench configuration: one full call per gamma (what the reference's evaluators do)
against one full call + bank-reusing calls (sample_guidance_sweep's mechanism).  python tools/sweep_timing.py"""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from ratio_guided_multimodal_fm_b200 import _lib, synthetic as S

dev = torch.device("cuda", 0)
B, N, steps, gammas = 1024, 256, 100, [0.0, 0.25, 0.5, 1.0, 2.0]
sdx, sdy, sdr = S.state_dict("unet_mnist32", 12), S.state_dict("unet_svhn", 13), S.state_dict("ratio_mnist_svhn", 22)
plan = _lib.SamplerPlan("mnist_svhn", sdx, sdy, sdr, guidance_method="mc_feng", guidance_strength=gammas[0], num_samples=B,
                        num_steps=steps, mc_batch_size=N, precision="fp16", device=dev)
g = torch.Generator(device=dev).manual_seed(0)
x0 = torch.randn(B, 1, 32, 32, device=dev, generator=g); y0 = torch.randn(B, 3, 32, 32, device=dev, generator=g)
m0 = torch.randn(N, 1, 32, 32, device=dev, generator=g); n0 = torch.randn(N, 3, 32, 32, device=dev, generator=g)
plan.run(x0, y0, m0, n0); torch.cuda.synchronize()


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def per_gamma_full():
    for gm in gammas:
        plan.set_strength(gm)
        plan.run(x0, y0, m0, n0)


def shared_bank():
    for i, gm in enumerate(gammas):
        plan.set_strength(gm)
        plan.run(x0, y0, m0, n0) if i == 0 else plan.run(x0, y0, reuse_bank=True)


a, b = timed(per_gamma_full), timed(shared_bank)
print(f"{len(gammas)} gammas x {B} pairs: one bank per gamma {a:.0f} ms ({len(gammas) * B / a * 1e3:.0f} pairs/s), "
      f"shared bank {b:.0f} ms ({len(gammas) * B / b * 1e3:.0f} pairs/s), x{a / b:.2f}")

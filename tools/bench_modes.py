"""Throughput of the other sampler configurations (not the headline bench): python tools/bench_modes.py"""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from ratio_guided_multimodal_fm_b200 import _lib, synthetic as S

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)


def run(pair, kinds, shapes, method, loss, B, N, steps, prec, reps=2):
    sdx, sdy, sdr = (S.state_dict(k, s) for k, s in zip(kinds, (12, 13, 22)))
    plan = _lib.SamplerPlan(pair, sdx, sdy, sdr, guidance_method=method, guidance_strength=0.5, num_samples=B, num_steps=steps,
                            mc_batch_size=N, loss_type=loss, precision=prec, device=dev)
    x0, y0 = torch.randn(B, *shapes[0], generator=g).to(dev), torch.randn(B, *shapes[1], generator=g).to(dev)
    m0, n0 = torch.randn(N, *shapes[0], generator=g).to(dev), torch.randn(N, *shapes[1], generator=g).to(dev)
    plan.run(x0, y0, m0, n0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        x1, y1 = plan.run(x0, y0, m0, n0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{pair:10s} {method:14s} {loss:6s} {prec:4s} B={B:5d} N_mc={N:3d} steps={steps:3d}: {ms:9.1f} ms/call  {B / ms * 1e3:9.1f} pairs/s"
          f"  finite={bool(torch.isfinite(x1).all() and torch.isfinite(y1).all())}")
    plan.close()


MS = ("mnist_svhn", ("unet_mnist32", "unet_svhn", "ratio_mnist_svhn"), ((1, 32, 32), (3, 32, 32)))
P28 = ("pair28", ("unet_mnist28", "unet_mnist28", "ratio28"), ((1, 28, 28), (1, 28, 28)))
run(*P28, "mc_feng", "disc", 64, 128, 100, "fp16")           # BASELINE configs[0]
run(*P28, "mc_feng", "rulsif", 2048, 128, 100, "fp16")       # configs[4] per-GPU shard (16384 / 8)
run(*MS, "mc_feng", "disc", 4096, 256, 100, "fp16", reps=1)  # configs[2] batch
run(*MS, "none", "disc", 1024, 256, 100, "fp16")
run(*MS, "grad_log_ratio", "disc", 1024, 256, 10, "fp16", reps=1)
run(*MS, "mc_feng", "disc", 1024, 256, 100, "bf16")
run(*MS, "mc_feng", "disc", 256, 256, 20, "fp32", reps=1)
# `--model original` (FlowMatchingModel, fp32 kernels) in the 28x28 sampler
P28O = ("pair28", ("flow28", "flow28", "ratio28"), ((1, 28, 28), (1, 28, 28)))
run(*P28O, "mc_feng", "disc", 256, 128, 100, "fp16", reps=1)

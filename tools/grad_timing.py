"""grad_log_ratio guidance at the bench batch: per-kernel profile of one guided Euler step (U-Nets + estimator forward + dgrad chain).
  python tools/grad_timing.py [--precision fp16] [--batch 1024]"""
import argparse, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from ratio_guided_multimodal_fm_b200 import _lib, synthetic as S
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--precision", default="fp16")
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda", 0)
sdx, sdy, sdr = S.state_dict("unet_mnist32", 12), S.state_dict("unet_svhn", 13), S.state_dict("ratio_mnist_svhn", 22)
plan = _lib.SamplerPlan("mnist_svhn", sdx, sdy, sdr, guidance_method="grad_log_ratio", guidance_strength=0.3, num_samples=a.batch,
                        num_steps=a.steps, mc_batch_size=4, loss_type="disc", precision=a.precision, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
x0 = torch.randn(a.batch, 1, 32, 32, device=dev, generator=g); y0 = torch.randn(a.batch, 3, 32, 32, device=dev, generator=g)
plan.run(x0, y0); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); x1, y1, diag = plan.run(x0, y0, want_diag=True); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"grad_log_ratio [{a.precision}] B={a.batch}: {ms / a.steps:.2f} ms per guided Euler step ({a.batch * 1e3 / ms * a.steps / 100:.1f} pairs/s at 100 steps), "
      f"finite={bool(torch.isfinite(x1).all() and torch.isfinite(y1).all())}, saturated rows={float(diag[12])}")

"""Tiny end-to-end run for compute-sanitizer: every kernel family once at ragged sizes (B=5, N_mc=6, chunk 4, 3 steps)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from ratio_guided_multimodal_fm_b200 import _lib, synthetic as S

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
for pair, kx, ky, kr, sx, sy in (("mnist_svhn", "unet_mnist32", "unet_svhn", "ratio_mnist_svhn", (1, 32, 32), (3, 32, 32)),
                                 ("pair28", "unet_mnist28", "unet_mnist28", "ratio28", (1, 28, 28), (1, 28, 28))):
    sdx, sdy, sdr = S.state_dict(kx, 12), S.state_dict(ky, 13), S.state_dict(kr, 22)
    B, N = 5, 6
    x0, y0 = torch.randn(B, *sx, generator=g).to(dev), torch.randn(B, *sy, generator=g).to(dev)
    m0, n0 = torch.randn(N, *sx, generator=g).to(dev), torch.randn(N, *sy, generator=g).to(dev)
    for prec, method in (("fp16", "mc_feng"), ("fp32", "mc_feng"), ("fp32", "grad_log_ratio")):
        plan = _lib.SamplerPlan(pair, sdx, sdy, sdr, guidance_method=method, guidance_strength=0.5, num_samples=B, num_steps=3,
                                mc_batch_size=N, precision=prec, device=dev, chunk=4)
        x1, y1 = plan.run(x0, y0, m0, n0)
        torch.cuda.synchronize()
        print(pair, prec, method, float(x1.abs().mean()), float(y1.abs().mean()))
        plan.close()
c = _lib.Net("clf_svhn", S.state_dict("clf_svhn", 3), max_batch=4, device=dev)
print(c.classifier_logits(torch.randn(3, 3, 32, 32, generator=g).to(dev)).shape)

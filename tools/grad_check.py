"""Gradient accuracy of the ratio estimators' dgrad chain vs autograd through the oracle, per precision (developer check)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from ratio_guided_multimodal_fm_b200 import _lib, synthetic as S
from oracle import rgfm_oracle as O
dev = "cuda:0"
for kind in ("ratio28", "ratio_mnist_svhn"):
    sd = S.state_dict(kind, 21 if kind == "ratio28" else 22)
    hw, cy = (28, 1) if kind == "ratio28" else (32, 3)
    gen = torch.Generator().manual_seed(17)
    x, y = torch.randn(7, 1, hw, hw, generator=gen), torch.randn(7, cy, hw, hw, generator=gen)
    ogx, ogy, olr = O.grad_log_ratio(sd, kind, x, y, "disc")
    for prec in ("fp32", "fp16", "bf16"):
        net = _lib.Net(kind, sd, max_batch=4, precision=prec, device=dev, grad=True)
        gx, gy, lr = net.grad_log_ratio(x.to(dev), y.to(dev), "disc")
        torch.cuda.synchronize()
        out = []
        for ours, ref in ((gx, ogx), (gy, ogy)):
            e = ours.cpu() - ref
            # share of the error carried by the 1 % largest entries: pool-routing flips are sparse and large
            k = max(1, e.numel() // 100)
            top = e.abs().flatten().topk(k).values.pow(2).sum().sqrt() / e.norm()
            out.append(f"rel_l2 {float(e.norm() / ref.norm()):.4f} max {float(e.abs().max() / ref.abs().max()):.3f} top1%share {float(top):.2f}")
        print(f"{kind:18s} {prec}: lr err {float((lr.cpu() - olr).abs().max()):.2e} | gx {out[0]} | gy {out[1]}")
        net.close()

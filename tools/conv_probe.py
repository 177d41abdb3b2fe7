"""One U-Net forward at a given batch through the C ABI (ncu target: conv launches in network order).
  python tools/conv_probe.py [--kind unet_mnist32] [--batch 512] [--precision fp16]"""
import argparse, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from ratio_guided_multimodal_fm_b200 import _lib, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="unet_mnist32")
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--precision", default="fp16")
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
sd = S.state_dict(a.kind, 12)
net = _lib.Net(a.kind, sd, max_batch=a.batch, precision=a.precision, device=dev)
c = 3 if a.kind == "unet_svhn" else 1
hw = 28 if a.kind == "unet_mnist28" else 32
x = torch.randn(a.batch, c, hw, hw, device=dev)
for _ in range(a.reps):
    v = net.unet_forward(x, 0.5)
torch.cuda.synchronize()
print("ok", float(v.abs().mean()))

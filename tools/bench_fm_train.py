"""Flow-matching training step (fm_train.fm_step + Adam) timed beside stock PyTorch on the same module and GPU.
  python tools/bench_fm_train.py [--batch 128] [--reps 10]
Columns: library 3x3 convs (exact fp32 CUDA-core kernels) | stock PyTorch, cuDNN TF32 allowed (its default) | stock, TF32 off."""
import argparse, copy, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from oracle.unet_module import UNet  # test infrastructure: the reference's module tree
from ratio_guided_multimodal_fm_b200 import CFMSchedule, synthetic as S
from ratio_guided_multimodal_fm_b200.fm_train import fm_step

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
dev = torch.device("cuda", 0)
sched = CFMSchedule()


def timed(fn):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(args.reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.reps


for kind in ("unet_mnist28", "unet_mnist32", "unet_svhn"):
    c, hw = (3, 32) if kind == "unet_svhn" else (1, 28 if kind == "unet_mnist28" else 32)
    base = UNet(kind)
    base.load_state_dict(S.state_dict(kind, 12), strict=True)
    x_1 = torch.randn(args.batch, c, hw, hw, device=dev)
    res = {}
    for name in ("library", "stock_tf32", "stock_fp32"):
        m = copy.deepcopy(base).to(dev).train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-4)
        torch.backends.cudnn.allow_tf32 = name != "stock_fp32"

        def step():
            opt.zero_grad()
            if name == "library":
                loss = fm_step(m, x_1, sched)
            else:
                t = torch.rand(args.batch, device=dev)
                x_t, u = sched.add_noise(x_1, t)
                loss = F.mse_loss(m(x_t, t), u)
                loss.backward()
            opt.step()
            return loss
        res[name] = timed(step)
    torch.backends.cudnn.allow_tf32 = True
    print(f"FMTRAIN {kind} batch {args.batch}: library {res['library']:.2f} ms | stock PyTorch (cuDNN, TF32 allowed) {res['stock_tf32']:.2f} ms | "
          f"stock PyTorch (TF32 off) {res['stock_fp32']:.2f} ms per step (forward + backward + Adam)")

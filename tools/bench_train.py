"""Time one ratio-estimator training step (rgfm_ratio_trainer_step) on the device: CUDA events, parameters resident.

    python tools/bench_train.py [--batch 128] [--reps 20] [--once KIND]   (--once: a single step, for an ncu launch list)
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from ratio_guided_multimodal_fm_b200 import RatioTrainPlan, synthetic  # noqa: E402

# forward MACs * 2 per sample (a9 / a10); a training step is ~3x (forward, dgrad, wgrad)
FWD_FLOP = {"ratio28": 36_030_976, "ratio_mnist_svhn": 368_968_192}
SHAPES = {"ratio28": ((1, 28, 28), (1, 28, 28), (512, 256)), "ratio_mnist_svhn": ((1, 32, 32), (3, 32, 32), (512, 512))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--once", default=None)
    args = ap.parse_args()
    dev = "cuda:0"
    for kind in ([args.once] if args.once else ["ratio28", "ratio_mnist_svhn"]):
        sd = synthetic.state_dict(kind, 21)
        params = {k: v.to(dev).contiguous() for k, v in sd.items() if v.dtype.is_floating_point}
        B = args.batch
        plan = RatioTrainPlan(kind, params, max_batch=B, device=dev)
        sx, sy, widths = SHAPES[kind]
        gen = torch.Generator(device=dev).manual_seed(1)
        x = torch.randn(B, *sx, device=dev, generator=gen)
        y = torch.randn(B, *sy, device=dev, generator=gen)
        r = (torch.arange(B, device=dev) % 2 == 0).float()
        masks = [torch.bernoulli(torch.full((B, w), 0.9, device=dev), generator=gen) / 0.9 for w in widths]
        for loss in ("disc", "rulsif"):
            if args.once:
                plan.step(x, y, r, loss, 0.1, masks)
                torch.cuda.synchronize()
                continue
            for _ in range(3):
                plan.step(x, y, r, loss, 0.1, masks)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                out8, _ = plan.step(x, y, r, loss, 0.1, masks)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            tf = 3.0 * FWD_FLOP[kind] * B / (ms * 1e-3) / 1e12
            print(f"{kind:18s} {loss:7s} B={B}: {ms:8.3f} ms/step  {B / ms * 1e3:10.0f} pairs/s  ~{tf:6.2f} TFLOP/s (3x forward)  loss={out8[0].item():.5f}")
        plan.close()


if __name__ == "__main__":
    main()

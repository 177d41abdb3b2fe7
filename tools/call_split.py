"""Times one configs[1] sampler call and its main loop alone (bank reused) with CUDA events; used for same-box A/B of
environment switches (RGFM_PDL, RGFM_SPLIT_PRE / _MAIN, ...) and of the plan options.
  python tools/call_split.py [--pair mnist_svhn|pair28] [--batch 1024] [--nmc 256] [--chunk 0] [--reps 3]"""
import argparse, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from ratio_guided_multimodal_fm_b200 import _lib, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--nmc", type=int, default=256)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--precision", default="fp16")
ap.add_argument("--pair", default="mnist_svhn", choices=["mnist_svhn", "pair28"])
ap.add_argument("--chunk", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda", 0)
if args.pair == "pair28":
    sdx, sdy, sdr = S.state_dict("unet_mnist28", 11), S.state_dict("unet_mnist28", 14), S.state_dict("ratio28", 21)
    shx, shy = (1, 28, 28), (1, 28, 28)
else:
    sdx, sdy, sdr = S.state_dict("unet_mnist32", 12), S.state_dict("unet_svhn", 13), S.state_dict("ratio_mnist_svhn", 22)
    shx, shy = (1, 32, 32), (3, 32, 32)
plan = _lib.SamplerPlan(args.pair, sdx, sdy, sdr, guidance_method="mc_feng", guidance_strength=0.5,
                        num_samples=args.batch, num_steps=args.steps, mc_batch_size=args.nmc, precision=args.precision, device=dev, chunk=args.chunk)
g = torch.Generator(device=dev).manual_seed(0)
x0 = torch.randn(args.batch, *shx, device=dev, generator=g); y0 = torch.randn(args.batch, *shy, device=dev, generator=g)
m0 = torch.randn(args.nmc, *shx, device=dev, generator=g); n0 = torch.randn(args.nmc, *shy, device=dev, generator=g)


def timed(fn):
    best = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1))
    return sorted(best)[len(best) // 2]


plan.run(x0, y0, m0, n0)
full = timed(lambda: plan.run(x0, y0, m0, n0))
main = timed(lambda: plan.run(x0, y0, reuse_bank=True))
x1, y1 = plan.run(x0, y0, m0, n0)
print(f"SPLIT {args.pair} B={args.batch} N={args.nmc} chunk={args.chunk} env={ {k: v for k, v in os.environ.items() if k.startswith('RGFM_')} } full {full:.1f} ms  main {main:.1f} ms  "
      f"prepass+ratio {full - main:.1f} ms  pairs/s {1e3 * args.batch / full:.1f}  checksum {x1.double().sum().item():.6f} {y1.double().sum().item():.6f}")

"""Small driver for ncu: one eager main-loop step (every kernel of the hot path once) at the bench batch.
  python tools/profile_step.py [--batch 1024] [--precision fp16]"""
import argparse, os, sys, json
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from ratio_guided_multimodal_fm_b200 import _lib, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--nmc", type=int, default=256)
ap.add_argument("--precision", default="fp16")
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--no-warm-run", action="store_true", help="skip the full sampler call before the profiled step (ncu captures: 27k fewer launches)")
args = ap.parse_args()
dev = torch.device("cuda", 0)
sdx, sdy, sdr = S.state_dict("unet_mnist32", 12), S.state_dict("unet_svhn", 13), S.state_dict("ratio_mnist_svhn", 22)
plan = _lib.SamplerPlan("mnist_svhn", sdx, sdy, sdr, guidance_method="mc_feng", guidance_strength=0.5,
                        num_samples=args.batch, num_steps=2, mc_batch_size=args.nmc, precision=args.precision, device=dev,
                        chunk=args.chunk)
g = torch.Generator(device=dev).manual_seed(0)
x0 = torch.randn(args.batch, 1, 32, 32, device=dev, generator=g); y0 = torch.randn(args.batch, 3, 32, 32, device=dev, generator=g)
m0 = torch.randn(args.nmc, 1, 32, 32, device=dev, generator=g); n0 = torch.randn(args.nmc, 3, 32, 32, device=dev, generator=g)
if not args.no_warm_run:
    plan.run(x0, y0, m0, n0)
torch.cuda.synchronize()
prof = plan.profile_step()
graph_ms = prof.pop("graph_step", {"ms": float("nan")})["ms"]
tot = sum(v["ms"] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:14s} {v['ms']:8.3f} ms  {100*v['ms']/tot:5.1f}%  launches {v['launches']:4d}  "
          f"{v['flops']/max(v['ms'],1e-9)/1e9:8.1f} TFLOP/s  {v['bytes']/max(v['ms'],1e-9)/1e6:8.1f} GB/s(alg)")
print("total", round(tot, 3), "ms (eager, per-kernel events);  one CUDA-graph step:", round(graph_ms, 3), "ms")

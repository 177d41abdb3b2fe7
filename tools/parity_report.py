"""Print max-abs errors of the CUDA path (through the C ABI) against every sampler golden fixture, per precision.
Run on a GPU box:  python tools/parity_report.py > profiles/parity_rNN.md"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import load_golden, T
from oracle import rgfm_oracle as O
from ratio_guided_multimodal_fm_b200 import _lib, synthetic as S

DEV = "cuda:0"
names = ["sampler_pair28_small", "sampler_pair28_rulsif", "sampler_pair28_none", "sampler_mnist_svhn_small",
         "sampler_mnist_svhn_g2", "sampler_pair28_cfg0", "sampler_mnist_svhn_100"]
print("| fixture | B | N_mc | steps | gamma | loss | precision | max abs dx1 | max abs dy1 | bank dx | bank dy | ratios |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for name in names:
    g = load_golden(name)
    pair = str(g["pair"]); kx, ky, kr, _, _ = O.PAIR_KINDS[pair]
    sx, sy, sr = (int(s) for s in g["seeds"]); method = str(g["method"]); guided = method == "mc_feng"
    for prec in ("fp32", "fp16", "bf16"):
        plan = _lib.SamplerPlan(pair, S.state_dict(kx, sx), S.state_dict(ky, sy), S.state_dict(kr, sr) if guided else None,
                                guidance_method=method, guidance_strength=float(g["gamma"]), num_samples=int(g["B"]),
                                num_steps=int(g["steps"]), mc_batch_size=int(g["N_mc"]), loss_type=str(g["loss_type"]),
                                precision=prec, device=DEV)
        a = [T(g["x0"]).to(DEV), T(g["y0"]).to(DEV)] + ([T(g["mc_x0"]).to(DEV), T(g["mc_y0"]).to(DEV)] if guided else [None, None])
        x1, y1 = plan.run(*a); torch.cuda.synchronize()
        e = lambda u, v: (u.cpu() - T(v)).abs().max().item()
        bank = ("", "", "")
        if guided:
            mx, my, r = plan.bank(); bank = (f"{e(mx, g['mc_x']):.2e}", f"{e(my, g['mc_y']):.2e}", f"{e(r, g['mc_ratios']):.2e}")
        print(f"| {name} | {int(g['B'])} | {int(g['N_mc'])} | {int(g['steps'])} | {float(g['gamma'])} | {g['loss_type']} | {prec} | "
              f"{e(x1, g['x1']):.2e} | {e(y1, g['y1']):.2e} | {bank[0]} | {bank[1]} | {bank[2]} |")
        plan.close()

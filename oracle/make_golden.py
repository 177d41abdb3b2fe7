"""Pin the oracle against the live reference and write tests/golden/*.npz.

Runs ONLY in the authoring container (needs /root/reference, which does not travel
to the GPU box).  For every case it
  1. draws the synthetic state_dicts (ratio_guided_multimodal_fm_b200.synthetic),
  2. loads them with strict=True into the UNMODIFIED reference modules,
  3. runs the reference function (the reference samplers draw their own noise from
     the global RNG after set_seed; the same four torch.randn draws are replayed to
     obtain x0, y0, mc_x0, mc_y0 -- flow_utils.py:220-244, sample_mnist_svhn.py:74-98),
  4. runs oracle/rgfm_oracle.py on the same weights + noise and asserts max-abs == 0.0,
  5. stores inputs + reference outputs as a fixture.

Usage:  python oracle/make_golden.py [--only NAME] [--skip-long]
"""
from __future__ import annotations

import argparse
import hashlib
import io
import os
import sys
import time
import types
import contextlib

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
# matplotlib is absent in this image; the reference only uses it for PNG output
for m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(m, types.ModuleType(m))

from oracle import rgfm_oracle as O  # noqa: E402
from ratio_guided_multimodal_fm_b200 import synthetic as S  # noqa: E402

from src.models.unet import FlowMatchingUNet  # noqa: E402
from src.models.flow_matching import FlowMatchingModel  # noqa: E402
from src.models.unet_flexible import FlowMatchingUNetMNIST, FlowMatchingUNetSVHN  # noqa: E402
from src.models.ratio_estimator import RatioEstimator  # noqa: E402
from src.models.ratio_flexible import RatioEstimatorMNISTSVHN  # noqa: E402
from src.models.classifier import MNISTClassifier  # noqa: E402
from src.models.svhn_classifier import SVHNClassifier, MNISTClassifier32  # noqa: E402
from src.utils import set_seed  # noqa: E402
from src.utils.flow_utils import sample_bimodal_guided, CFMSchedule  # noqa: E402
from src.sample_mnist_svhn import sample_bimodal_guided_mnist_svhn  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

REF_CLASSES = {
    "unet_mnist28": lambda lt: FlowMatchingUNet(),
    "flow28": lambda lt: FlowMatchingModel(),
    "unet_mnist32": lambda lt: FlowMatchingUNetMNIST(img_size=32),
    "unet_svhn": lambda lt: FlowMatchingUNetSVHN(),
    "ratio28": lambda lt: RatioEstimator(loss_type=lt),
    "ratio_mnist_svhn": lambda lt: RatioEstimatorMNISTSVHN(loss_type=lt),
    "clf_mnist28": lambda lt: MNISTClassifier(),
    "clf_mnist32": lambda lt: MNISTClassifier32(),
    "clf_svhn": lambda lt: SVHNClassifier(),
}


def ref_module(kind, seed, loss_type="disc"):
    sd = S.state_dict(kind, seed)
    m = REF_CLASSES[kind](loss_type)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


def sd_digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"  wrote tests/golden/{name}.npz")


def exact(a, b, what):
    d = (a - b).abs().max().item()
    assert d == 0.0, f"oracle != reference for {what}: max-abs {d}"


def case_unet(kind, seed, B):
    m, sd = ref_module(kind, seed)
    a = S.UNET_ARCH[kind]
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(B, a["in_channels"], a["img_size"], a["img_size"], generator=g)
    t = torch.rand(B, generator=g)
    t[0] = 0.0
    with torch.no_grad():
        v_ref = m(x, t)
        v_orc = O.unet_forward(sd, kind, x, t)
    exact(v_ref, v_orc, kind)
    save(f"{kind}_forward", seed=seed, x=x, t=t, v=v_ref, digest=sd_digest(sd))


def case_ratio(kind, seed, B):
    shp = {"ratio28": ((1, 28, 28), (1, 28, 28)), "ratio_mnist_svhn": ((1, 32, 32), (3, 32, 32))}[kind]
    g = torch.Generator().manual_seed(2000 + seed)
    x = torch.randn(B, *shp[0], generator=g)
    y = torch.randn(B, *shp[1], generator=g)
    out = {}
    for lt in ("disc", "rulsif"):
        m, sd = ref_module(kind, seed, lt)
        with torch.no_grad():
            s_ref = m(x, y)
            lr_ref = m.log_ratio(x, y)
            exact(s_ref, O.ratio_scores(sd, kind, x, y), kind + " scores")
            exact(lr_ref, O.ratio_log_ratio(sd, kind, x, y, lt), kind + " log_ratio " + lt)
        out["scores"] = s_ref
        out["log_ratio_" + lt] = lr_ref
    # gradient extension: autograd through the reference module itself
    m, sd = ref_module(kind, seed, "disc")
    xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    gx, gy = torch.autograd.grad(m.log_ratio(xg, yg).sum(), (xg, yg))
    ogx, ogy, _ = O.grad_log_ratio(sd, kind, x, y, "disc")
    exact(gx, ogx, kind + " grad x")
    exact(gy, ogy, kind + " grad y")
    save(f"{kind}_forward", seed=seed, x=x, y=y, grad_x_disc=gx, grad_y_disc=gy, digest=sd_digest(sd), **out)


def case_ratio_train(kind, seed, B, loss_type):
    """One RatioTrainer.train_step's loss + param.grad through the UNMODIFIED reference module and loss classes
    (trainer.py:36-52, losses.py:27-133), with the two training-mode effects pinned: fixed dropout masks are replayed by
    forward hooks on the nn.Dropout modules of a module in eval mode (= training mode with those masks and frozen BatchNorm)."""
    from src.utils.losses import get_ratio_loss
    shp = {"ratio28": ((1, 28, 28), (1, 28, 28)), "ratio_mnist_svhn": ((1, 32, 32), (3, 32, 32))}[kind]
    widths = {"ratio28": (512, 256), "ratio_mnist_svhn": (512, 512)}[kind]
    g = torch.Generator().manual_seed(3000 + seed)
    x = torch.randn(B, *shp[0], generator=g)
    y = torch.randn(B, *shp[1], generator=g)
    is_real = (torch.arange(B) % 3 != 1).float()          # uneven split: ~2/3 real
    masks = [torch.bernoulli(torch.full((B, w), 0.9), generator=g) / 0.9 for w in widths]
    m, sd = ref_module(kind, seed, loss_type)
    drops = [mod for mod in m.score_net if isinstance(mod, torch.nn.Dropout)]
    assert len(drops) == len(masks)
    hooks = [d.register_forward_hook(lambda mod, inp, out, mk=mk: out * mk) for d, mk in zip(drops, masks)]
    loss_fn = get_ratio_loss(loss_type)
    m.zero_grad()
    scores = m(x, y)
    loss, metrics = loss_fn(scores[is_real == 1], scores[is_real == 0])
    loss.backward()
    for h in hooks:
        h.remove()
    lam = getattr(loss_fn, "lambda_penalty", 0.0)
    oloss, ometrics, ograds = O.ratio_train_grads(sd, kind, x, y, is_real, loss_type, lam, masks)
    exact(loss.detach(), oloss, kind + " train loss " + loss_type)
    for k in metrics:
        assert metrics[k] == ometrics[k], (k, metrics[k], ometrics[k])
    out = {}
    for name, p in m.named_parameters():
        exact(p.grad, ograds[name], f"{kind} grad {name} ({loss_type})")
        flat = p.grad.flatten()
        idx = torch.linspace(0, flat.numel() - 1, min(flat.numel(), 64)).long()
        out["g:" + name] = torch.cat([flat.norm().reshape(1), flat.sum().reshape(1), flat[idx]])
    assert set(ograds) == {n for n, _ in m.named_parameters()}
    save(f"{kind}_train_{loss_type}", seed=seed, x=x, y=y, is_real=is_real, loss=loss.detach(), lambda_penalty=lam,
         metric_names=np.array(sorted(metrics)), metric_values=np.array([metrics[k] for k in sorted(metrics)]),
         digest=sd_digest(sd), **{f"mask{i}": mk for i, mk in enumerate(masks)}, **out)


def case_classifier(kind, seed, B):
    m, sd = ref_module(kind, seed)
    shp = {"clf_mnist28": (1, 28, 28), "clf_mnist32": (1, 32, 32), "clf_svhn": (3, 32, 32)}[kind]
    x = torch.randn(B, *shp, generator=torch.Generator().manual_seed(3000 + seed))
    with torch.no_grad():
        l_ref = m(x)
    exact(l_ref, O.classifier_logits(sd, kind, x), kind)
    save(f"{kind}_forward", seed=seed, x=x, logits=l_ref, digest=sd_digest(sd))


def case_sampler(name, pair, seeds, B, N_mc, steps, method, gamma, loss_type, noise_seed=42, kinds=None):
    kx, ky, kr, sx, sy = O.PAIR_KINDS[pair]
    if kinds:
        kx, ky = kinds
    fx, sdx = ref_module(kx, seeds[0])
    fy, sdy = ref_module(ky, seeds[1])
    fr, sdr = ref_module(kr, seeds[2], loss_type)
    fn = sample_bimodal_guided if pair == "pair28" else sample_bimodal_guided_mnist_svhn
    t0 = time.time()
    set_seed(noise_seed)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
        x1, y1 = fn(fx, fy, fr if method != "none" else None, guidance_method=method, guidance_strength=gamma,
                    num_samples=B, num_steps=steps, device="cpu", mc_batch_size=N_mc)
    t_ref = time.time() - t0
    # replay the reference's RNG draws
    set_seed(noise_seed)
    x0 = torch.randn(B, *sx)
    y0 = torch.randn(B, *sy)
    mcx0 = mcy0 = None
    if method == "mc_feng":
        mcx0 = torch.randn(N_mc, *sx)
        mcy0 = torch.randn(N_mc, *sy)
    ox, oy, aux = O.sample_pair(sdx, sdy, sdr if method != "none" else None, pair, method, gamma, steps,
                                x0, y0, mcx0, mcy0, loss_type, return_aux=True)
    exact(x1, ox, name + " x1")
    exact(y1, oy, name + " y1")
    extra = {}
    if method == "mc_feng":
        extra = dict(mc_x0=mcx0, mc_y0=mcy0, mc_x=aux["mc_x"], mc_y=aux["mc_y"], mc_ratios=aux["mc_ratios"])
    print(f"  {name}: reference {t_ref:.1f}s, oracle == reference (max-abs 0.0); stdout:\n    " +
          "\n    ".join(l for l in buf.getvalue().splitlines() if l.strip()))
    save(name, pair=pair, seeds=np.array(seeds), B=B, N_mc=N_mc, steps=steps, method=method, gamma=gamma,
         loss_type=loss_type, x0=x0, y0=y0, x1=x1, y1=y1, ref_stdout=buf.getvalue(),
         digests=np.array([sd_digest(sdx), sd_digest(sdy), sd_digest(sdr)]), **extra)


def case_cfm_sample(seed, B, steps, noise_seed=42):
    """CFMSchedule.sample (flow_utils.py:69-100): the unguided single-modality sampler, 28x28 by construction."""
    fm, sd = ref_module("unet_mnist28", seed)
    set_seed(noise_seed)
    x1 = CFMSchedule().sample(fm, B, num_steps=steps, device="cpu")
    set_seed(noise_seed)
    x0 = torch.randn(B, 1, 28, 28)
    exact(x1, O.flow_sample(sd, "unet_mnist28", x0, steps), "cfm_sample")
    save("cfm_sample_mnist28", seed=seed, B=B, steps=steps, x0=x0, x1=x1, digest=sd_digest(sd))


def case_mc_feng_block():
    """One guidance block on random state, several t (incl. the saturated regime)."""
    g = torch.Generator().manual_seed(77)
    B, N = 5, 12
    mcx = torch.randn(N, 1, 32, 32, generator=g) * 0.7
    mcy = torch.randn(N, 3, 32, 32, generator=g) * 0.7
    r = torch.rand(N, generator=g) * 3 + 0.01
    out = {}
    for i, t in enumerate((0.01, 0.2, 0.5, 0.9, 0.99)):
        x = t * mcx[torch.randint(0, N, (B,), generator=g)] + (1 - t) * torch.randn(B, 1, 32, 32, generator=g)
        y = t * mcy[torch.randint(0, N, (B,), generator=g)] + (1 - t) * torch.randn(B, 3, 32, 32, generator=g)
        gx, gy, w, Z = O.mc_feng_guidance(x, y, mcx, mcy, r, t)
        out.update({f"t{i}": t, f"x{i}": x, f"y{i}": y, f"gx{i}": gx, f"gy{i}": gy, f"w{i}": w, f"Z{i}": Z})
    save("mc_feng_block", mc_x=mcx, mc_y=mcy, mc_ratios=r, **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--skip-long", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(8)
    cases = [
        ("unet_mnist28", lambda: case_unet("unet_mnist28", 11, 3)),
        ("unet_mnist32", lambda: case_unet("unet_mnist32", 12, 3)),
        ("unet_svhn", lambda: case_unet("unet_svhn", 13, 3)),
        ("flow28", lambda: case_unet("flow28", 15, 3)),
        ("ratio28", lambda: case_ratio("ratio28", 21, 4)),
        ("ratio_mnist_svhn", lambda: case_ratio("ratio_mnist_svhn", 22, 4)),
        ("ratio28_train_disc", lambda: case_ratio_train("ratio28", 21, 6, "disc")),
        ("ratio28_train_rulsif", lambda: case_ratio_train("ratio28", 21, 6, "rulsif")),
        ("ratio_mnist_svhn_train_disc", lambda: case_ratio_train("ratio_mnist_svhn", 22, 6, "disc")),
        ("ratio_mnist_svhn_train_rulsif", lambda: case_ratio_train("ratio_mnist_svhn", 22, 6, "rulsif")),
        ("clf_mnist28", lambda: case_classifier("clf_mnist28", 31, 4)),
        ("clf_mnist32", lambda: case_classifier("clf_mnist32", 32, 4)),
        ("clf_svhn", lambda: case_classifier("clf_svhn", 33, 4)),
        ("mc_feng_block", case_mc_feng_block),
        ("cfm_sample_mnist28", lambda: case_cfm_sample(11, 4, 7)),
        ("sampler_pair28_small", lambda: case_sampler("sampler_pair28_small", "pair28", (11, 14, 21), 5, 8, 6, "mc_feng", 0.5, "disc")),
        ("sampler_pair28_rulsif", lambda: case_sampler("sampler_pair28_rulsif", "pair28", (11, 14, 21), 4, 8, 5, "mc_feng", 1.0, "rulsif")),
        ("sampler_pair28_none", lambda: case_sampler("sampler_pair28_none", "pair28", (11, 14, 21), 4, 8, 5, "none", 0.0, "disc")),
        # `--model original` (src/sample.py --model original): both flows are FlowMatchingModel
        ("sampler_pair28_original", lambda: case_sampler("sampler_pair28_original", "pair28", (15, 16, 21), 4, 8, 5, "mc_feng", 0.5, "disc",
                                                         kinds=("flow28", "flow28"))),
        ("sampler_mnist_svhn_small", lambda: case_sampler("sampler_mnist_svhn_small", "mnist_svhn", (12, 13, 22), 3, 6, 5, "mc_feng", 0.5, "disc")),
        ("sampler_mnist_svhn_g2", lambda: case_sampler("sampler_mnist_svhn_g2", "mnist_svhn", (12, 13, 22), 3, 6, 5, "mc_feng", 2.0, "rulsif")),
    ]
    long_cases = [
        # BASELINE.json configs[0]: the reference's own CPU-runnable case, full 100 steps
        ("sampler_pair28_cfg0", lambda: case_sampler("sampler_pair28_cfg0", "pair28", (11, 14, 21), 64, 128, 100, "mc_feng", 0.5, "disc")),
        # configs[1] shape at reduced batch, full 100 steps
        ("sampler_mnist_svhn_100", lambda: case_sampler("sampler_mnist_svhn_100", "mnist_svhn", (12, 13, 22), 16, 32, 100, "mc_feng", 0.5, "disc")),
    ]
    if not args.skip_long:
        cases += long_cases
    for name, fn in cases:
        if args.only and args.only != name:
            continue
        print(f"[{name}]")
        fn()


if __name__ == "__main__":
    main()

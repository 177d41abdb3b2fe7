"""CPU oracle for the ratio-guided coupled Euler sampler.

TEST INFRASTRUCTURE ONLY.  This file is a functional, state-dict driven CPU
restatement (torch fp32) of the reference's hot path.  It is imported by
`tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s CPU-baseline /
`--impl reference` legs, and by nothing else: the product path
(`ratio_guided_multimodal_fm_b200/`) never imports it and fails loudly when the
CUDA library is missing.

Parity status: PINNED against the live reference.  The reference ships no tests
or golden vectors, so the pin is
`oracle/make_golden.py`, which imports the unmodified reference modules from
`/root/reference` in the authoring container, loads the same synthetic
state_dicts into them, runs the reference samplers on identical noise and checks
this oracle bit-for-bit (max-abs 0.0) before writing `tests/golden/*.npz`.

Every function cites the reference file:line it restates.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

# ----------------------------------------------------------------------------
# Architectures on the hot path
# ----------------------------------------------------------------------------
UNET_CONFIGS = {
    # reference: src/models/unet.py:282-305 (FlowMatchingUNet defaults)
    "unet_mnist28": dict(in_channels=1, img_size=28, model_channels=32, channel_mult=(1, 2), num_res_blocks=2),
    # reference: src/models/unet_flexible.py:266-277 (FlowMatchingUNetMNIST(img_size=32))
    "unet_mnist32": dict(in_channels=1, img_size=32, model_channels=32, channel_mult=(1, 2), num_res_blocks=2),
    # reference: src/models/unet_flexible.py:280-291 (FlowMatchingUNetSVHN)
    "unet_svhn": dict(in_channels=3, img_size=32, model_channels=64, channel_mult=(1, 2, 2), num_res_blocks=2),
}

GN_GROUPS = 8  # unet.py:54,64,209 ; unet_flexible.py:51,61,196 (min(8, C) == 8 for every C used)


# ----------------------------------------------------------------------------
# U-Net velocity field
# ----------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    """Sinusoidal embedding, cos first then sin.  reference: unet.py:13-33, unet_flexible.py:16-36."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, device=t.device) / half)
    args = t[:, None] * freqs[None, :]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _resblock(sd: StateDict, p: str, x: Tensor, t_emb: Tensor) -> Tensor:
    """GN -> SiLU -> conv3x3 -> + time bias -> GN -> SiLU -> conv3x3 -> + skip.
    reference: unet.py:74-96, unet_flexible.py:71-85 (dropout is identity in eval)."""
    h = F.group_norm(x, GN_GROUPS, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    h = F.silu(h)
    h = F.conv2d(h, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    tb = F.linear(F.silu(t_emb), sd[p + "time_mlp.1.weight"], sd[p + "time_mlp.1.bias"])
    h = h + tb[:, :, None, None]
    h = F.group_norm(h, GN_GROUPS, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    h = F.silu(h)
    h = F.conv2d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if (p + "skip.weight") in sd:
        sk = F.conv2d(x, sd[p + "skip.weight"], sd[p + "skip.bias"])
    else:
        sk = x
    return h + sk


def flow28_forward(sd: StateDict, x: Tensor, t: Tensor) -> Tensor:
    """`--model original` velocity field.  reference: flow_matching.py:10-31 (sinusoidal embedding: sin then cos,
    log(1e4)/(half-1) spacing, t unscaled), :34-73 (encoder), :76-126 (decoder), :129-173 (model)."""
    half = 64  # time_emb_dim 128
    f = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    e = t[:, None] * f[None, :]
    t_emb = torch.cat((e.sin(), e.cos()), dim=-1)
    h = x
    for i, stride in ((1, 1), (2, 2), (3, 2), (4, 1)):
        h = F.conv2d(h, sd[f"encoder.conv{i}.weight"], sd[f"encoder.conv{i}.bias"], stride=stride, padding=1)
        h = F.silu(F.group_norm(h, GN_GROUPS, sd[f"encoder.gn{i}.weight"], sd[f"encoder.gn{i}.bias"]))
    feat = F.linear(h.reshape(h.shape[0], -1), sd["encoder.fc.weight"], sd["encoder.fc.bias"])
    h = F.linear(torch.cat([feat, t_emb], dim=1), sd["decoder.fc1.weight"], sd["decoder.fc1.bias"])
    h = h.view(h.shape[0], 256, 7, 7)
    for i in (1, 2):
        h = F.conv_transpose2d(h, sd[f"decoder.deconv{i}.weight"], sd[f"decoder.deconv{i}.bias"], stride=2, padding=1)
        h = F.silu(F.group_norm(h, GN_GROUPS, sd[f"decoder.gn{i}.weight"], sd[f"decoder.gn{i}.bias"]))
    h = F.conv2d(h, sd["decoder.conv3.weight"], sd["decoder.conv3.bias"], padding=1)
    h = F.silu(F.group_norm(h, GN_GROUPS, sd["decoder.gn3.weight"], sd["decoder.gn3.bias"]))
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


def unet_forward(sd: StateDict, kind: str, x: Tensor, t: Tensor) -> Tensor:
    """Velocity field v = fm(x, t).  reference: unet.py:216-278, unet_flexible.py:203-261.
    A state_dict of the `--model original` network (flow_matching.py) is recognised by its keys, like the CUDA plan does."""
    if kind == "flow28" or "encoder.conv1.weight" in sd:
        return flow28_forward(sd, x, t)
    cfg = UNET_CONFIGS[kind]
    mc, mult, nres = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    t_emb = timestep_embedding(t, mc)
    t_emb = F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    t_emb = F.linear(F.silu(t_emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])

    h = F.conv2d(x, sd["input_conv.weight"], sd["input_conv.bias"], padding=1)
    hs = [h]
    bi = di = 0
    for level in range(len(mult)):
        for _ in range(nres):
            h = _resblock(sd, f"encoder_blocks.{bi}.", h, t_emb)
            hs.append(h)
            bi += 1
        if level < len(mult) - 1:
            h = F.conv2d(h, sd[f"downsamplers.{di}.conv.weight"], sd[f"downsamplers.{di}.conv.bias"],
                         stride=2, padding=1)
            hs.append(h)
            di += 1
    h = _resblock(sd, "middle_block1.", h, t_emb)
    h = _resblock(sd, "middle_block2.", h, t_emb)
    bi = ui = 0
    for level in reversed(range(len(mult))):
        for _ in range(nres + 1):
            h = torch.cat([h, hs.pop()], dim=1)
            h = _resblock(sd, f"decoder_blocks.{bi}.", h, t_emb)
            bi += 1
        if level > 0:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[f"upsamplers.{ui}.conv.weight"], sd[f"upsamplers.{ui}.conv.bias"], padding=1)
            ui += 1
    h = F.group_norm(h, GN_GROUPS, sd["out_norm.weight"], sd["out_norm.bias"])
    h = F.silu(h)
    return F.conv2d(h, sd["out_conv.weight"], sd["out_conv.bias"], padding=1)


# ----------------------------------------------------------------------------
# Ratio estimators
# ----------------------------------------------------------------------------
def _head(scores: Tensor, loss_type: str) -> Tensor:
    """reference: ratio_estimator.py:180-191, ratio_flexible.py:377-385."""
    if loss_type == "disc":
        return F.logsigmoid(scores) - F.logsigmoid(-scores)
    if loss_type == "rulsif":
        return torch.log(F.softplus(scores) + 1e-8)
    raise ValueError(f"Unknown loss_type: {loss_type}")


def _encoder28(sd: StateDict, p: str, img: Tensor) -> Tensor:
    """(conv -> GN(8) -> SiLU -> maxpool2) x3, conv -> GN -> SiLU, GAP, fc.
    reference: ratio_estimator.py:67-93."""
    x = img
    for i in (1, 2, 3, 4):
        x = F.conv2d(x, sd[f"{p}conv{i}.weight"], sd[f"{p}conv{i}.bias"], padding=1)
        x = F.silu(F.group_norm(x, 8, sd[f"{p}gn{i}.weight"], sd[f"{p}gn{i}.bias"]))
        if i < 4:
            x = F.max_pool2d(x, 2)
    x = x.mean(dim=(2, 3))
    return F.linear(x, sd[p + "fc.weight"], sd[p + "fc.bias"])


def _bn_eval(sd: StateDict, p: str, x: Tensor) -> Tensor:
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        training=False, eps=1e-5)


def _encoder_mnist32(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """reference: ratio_flexible.py:210-232 (BatchNorm in eval mode = running statistics)."""
    h = x
    for i in (1, 2, 3, 4):
        h = F.conv2d(h, sd[f"{p}conv{i}.weight"], sd[f"{p}conv{i}.bias"], padding=1)
        h = F.silu(_bn_eval(sd, f"{p}bn{i}.", h))
        if i < 4:
            h = F.max_pool2d(h, 2)
    h = h.mean(dim=(2, 3))
    return F.linear(h, sd[p + "fc.weight"], sd[p + "fc.bias"])


def _encoder_svhn(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """reference: ratio_flexible.py:271-302."""
    h = x
    for blk in (1, 2, 3, 4):
        for ab in ("a", "b"):
            h = F.conv2d(h, sd[f"{p}conv{blk}{ab}.weight"], sd[f"{p}conv{blk}{ab}.bias"], padding=1)
            h = F.silu(_bn_eval(sd, f"{p}bn{blk}{ab}.", h))
        h = F.max_pool2d(h, 2)
    h = h.mean(dim=(2, 3))
    return F.linear(h, sd[p + "fc.weight"], sd[p + "fc.bias"])


def _score_net(sd: StateDict, feats: Tensor, linear_ln: Tuple[Tuple[int, int], ...], last: int, masks=None) -> Tensor:
    """Linear -> LayerNorm -> SiLU -> Dropout per hidden layer (ratio_estimator.py:125-135).  Dropout is the identity in eval
    mode; `masks` (training step, one [B, width] tensor of 0 | 1/(1-p) per hidden layer) multiplies where nn.Dropout would."""
    h = feats
    for i, (li, ni) in enumerate(linear_ln):
        h = F.linear(h, sd[f"score_net.{li}.weight"], sd[f"score_net.{li}.bias"])
        h = F.layer_norm(h, (h.shape[-1],), sd[f"score_net.{ni}.weight"], sd[f"score_net.{ni}.bias"])
        h = F.silu(h)
        if masks is not None and i < len(masks):  # ratio_flexible.py:332-345 has no Dropout after the third hidden layer
            h = h * masks[i]
    return F.linear(h, sd[f"score_net.{last}.weight"], sd[f"score_net.{last}.bias"]).squeeze(-1)


def ratio_scores(sd: StateDict, kind: str, x: Tensor, y: Tensor, masks=None) -> Tensor:
    """T_theta(x, y).  reference: ratio_estimator.py:137-158, ratio_flexible.py:347-364."""
    if kind == "ratio28":
        fx = _encoder28(sd, "encoder_x.", x)
        fy = _encoder28(sd, "encoder_y.", y)
        return _score_net(sd, torch.cat([fx, fy], dim=1), ((0, 1), (4, 5)), 8, masks)
    if kind == "ratio_mnist_svhn":
        fx = _encoder_mnist32(sd, "encoder_mnist.", x)
        fy = _encoder_svhn(sd, "encoder_svhn.", y)
        return _score_net(sd, torch.cat([fx, fy], dim=1), ((0, 1), (4, 5), (8, 9)), 11, masks)
    raise ValueError(kind)


# ----------------------------------------------------------------------------
# Training side: the ratio losses and one step's parameter gradients
# ----------------------------------------------------------------------------
def ratio_loss(scores_real: Tensor, scores_fake: Tensor, loss_type: str, lambda_penalty: float = 0.1):
    """(loss, metrics) of DiscriminatorLoss.forward (losses.py:41-65) / RuLSIFLoss.forward (losses.py:93-130; `alpha` is
    stored by the reference but not used by its forward)."""
    if loss_type == "disc":
        loss = (F.binary_cross_entropy_with_logits(scores_real, torch.ones_like(scores_real))
                + F.binary_cross_entropy_with_logits(scores_fake, torch.zeros_like(scores_fake)))
        with torch.no_grad():
            m = {"loss": loss.item(), "acc_real": (scores_real > 0).float().mean().item(),
                 "acc_fake": (scores_fake < 0).float().mean().item()}
        return loss, m
    if loss_type == "rulsif":
        w_real, w_fake = F.softplus(scores_real), F.softplus(scores_fake)
        w_mix = torch.cat([w_real, w_fake])
        constraint = lambda_penalty * (w_mix.mean() - 1.0).pow(2)
        loss = 0.5 * w_mix.pow(2).mean() - w_real.mean() + constraint
        with torch.no_grad():
            m = {"loss": loss.item(), "mean_w_real": w_real.mean().item(), "mean_w_fake": w_fake.mean().item(),
                 "constraint_term": constraint.item()}
        return loss, m
    raise ValueError(f"Unknown loss_type: {loss_type}")


TRAIN_BUFFERS = ("running_mean", "running_var", "num_batches_tracked")


def ratio_train_grads(sd: StateDict, kind: str, x: Tensor, y: Tensor, is_real: Tensor, loss_type: str,
                      lambda_penalty: float = 0.1, masks=None):
    """What RatioTrainer.train_step (trainer.py:36-52) leaves in param.grad: scores = model(x, y); split by is_real;
    loss_fn; backward.  BatchNorm uses its running statistics (frozen) and dropout is `masks` -- the two training-mode
    effects that are not a pure function of (weights, batch).  Returns (loss, metrics, {name: grad})."""
    with torch.enable_grad():
        p = {k: (v.detach().clone().requires_grad_(True) if v.dtype.is_floating_point and not k.endswith(TRAIN_BUFFERS) else v)
             for k, v in sd.items()}
        scores = ratio_scores(p, kind, x, y, masks)
        loss, metrics = ratio_loss(scores[is_real == 1], scores[is_real == 0], loss_type, lambda_penalty)
        names = [k for k, v in p.items() if v.requires_grad]
        grads = torch.autograd.grad(loss, [p[k] for k in names])
    return loss.detach(), metrics, dict(zip(names, grads))


def ratio_log_ratio(sd: StateDict, kind: str, x: Tensor, y: Tensor, loss_type: str) -> Tensor:
    return _head(ratio_scores(sd, kind, x, y), loss_type)


# ----------------------------------------------------------------------------
# mc_feng guidance step and the coupled Euler loop
# ----------------------------------------------------------------------------
def mc_feng_guidance(x_t: Tensor, y_t: Tensor, mc_x: Tensor, mc_y: Tensor, mc_ratios: Tensor,
                     t: float, eps: float = 1e-3):
    """Importance-weighted conditional velocity, verbatim operation order.
    reference: flow_utils.py:275-341 == sample_mnist_svhn.py:125-167.
    Returns (g_x, g_y, weights, Z_bar)."""
    B, N = x_t.shape[0], mc_x.shape[0]
    sigma_t = 1 - t + eps
    mu_x = t * mc_x.unsqueeze(0)
    mu_y = t * mc_y.unsqueeze(0)
    diff_x = (x_t.unsqueeze(1) - mu_x).view(B, N, -1)
    diff_y = (y_t.unsqueeze(1) - mu_y).view(B, N, -1)
    log_p = (-0.5 * (diff_x ** 2).sum(dim=-1) / (sigma_t ** 2)) + (-0.5 * (diff_y ** 2).sum(dim=-1) / (sigma_t ** 2))
    p = (log_p - log_p.max(dim=1, keepdim=True)[0]).exp()
    p_bar = p.mean(dim=1, keepdim=True) + 1e-10
    r = mc_ratios.unsqueeze(0)
    Z_bar = (r * p).mean(dim=1, keepdim=True) + 1e-10
    w = (r / Z_bar) * (p / p_bar)
    w = w / (w.sum(dim=1, keepdim=True) + 1e-10)
    vc_x = (mc_x.unsqueeze(0) - x_t.unsqueeze(1)) / (1 - t + eps)
    vc_y = (mc_y.unsqueeze(0) - y_t.unsqueeze(1)) / (1 - t + eps)
    wv = w.view(B, N, 1, 1, 1)
    return (wv * vc_x).sum(dim=1), (wv * vc_y).sum(dim=1), w, Z_bar


PAIR_KINDS = {
    # pair kind -> (unet_x kind, unet_y kind, ratio kind, x shape, y shape)
    "pair28": ("unet_mnist28", "unet_mnist28", "ratio28", (1, 28, 28), (1, 28, 28)),        # flow_utils.py:220-221
    "mnist_svhn": ("unet_mnist32", "unet_svhn", "ratio_mnist_svhn", (1, 32, 32), (3, 32, 32)),  # sample_mnist_svhn.py:74-75
}


@torch.no_grad()
def mc_prepass(sd_x: StateDict, sd_y: StateDict, sd_r: StateDict, pair: str, loss_type: str,
               mc_x0: Tensor, mc_y0: Tensor, num_steps: int):
    """Euler-integrate the MC bank through both flows, then one ratio forward.
    reference: flow_utils.py:231-258, sample_mnist_svhn.py:85-112."""
    kx, ky, kr, _, _ = PAIR_KINDS[pair]
    dt = 1.0 / num_steps
    n = mc_x0.shape[0]
    mc_x = mc_x0
    for s in range(num_steps):
        mc_x = mc_x + unet_forward(sd_x, kx, mc_x, torch.full((n,), s * dt)) * dt
    mc_y = mc_y0
    for s in range(num_steps):
        mc_y = mc_y + unet_forward(sd_y, ky, mc_y, torch.full((n,), s * dt)) * dt
    mc_ratios = ratio_log_ratio(sd_r, kr, mc_x, mc_y, loss_type).exp()
    return mc_x, mc_y, mc_ratios


@torch.no_grad()
def flow_sample(sd: StateDict, kind: str, x0: Tensor, num_steps: int) -> Tensor:
    """Unguided single-modality Euler sampler on explicit prior noise.
    reference: CFMSchedule.sample, flow_utils.py:69-100 (which draws x0 = randn(num_samples, 1, 28, 28) itself)."""
    dt = 1.0 / num_steps
    x_t = x0
    n = x0.shape[0]
    for step in range(num_steps):
        v_t = unet_forward(sd, kind, x_t, torch.full((n,), step * dt))
        x_t = x_t + v_t * dt
    return x_t


@torch.no_grad()
def sample_pair(sd_x: StateDict, sd_y: StateDict, sd_r: Optional[StateDict], pair: str,
                guidance_method: str, guidance_strength: float, num_steps: int,
                x0: Tensor, y0: Tensor, mc_x0: Optional[Tensor] = None, mc_y0: Optional[Tensor] = None,
                loss_type: str = "disc", return_aux: bool = False, bank=None, gamma_schedule=None):
    """Coupled guided Euler sampler on explicit noise.
    gamma_schedule: optional per-step strengths (time-dependent guidance, SURVEY.md 8f-2: an extension, the reference's
    gamma is constant); step k uses gamma_schedule[k] where the reference uses guidance_strength.
    bank: optional precomputed (mc_x, mc_y, mc_ratios) -- the full-size parity tests check a row subset of a large
    batch against the bank the CUDA path computed, instead of paying the 256-sample CPU pre-pass.
    reference: flow_utils.py:178-375 and sample_mnist_svhn.py:39-177 (same algorithm, different shapes).
    The reference draws x0, y0, mc_x0, mc_y0 from the global RNG in this order; here they are arguments so the
    CUDA path can be fed bit-identical noise."""
    kx, ky, _, _, _ = PAIR_KINDS[pair]
    dt = 1.0 / num_steps
    eps = 1e-3
    x_t, y_t = x0, y0
    B = x0.shape[0]
    if not (guidance_method == "mc_feng" and sd_r is not None):
        bank = None
    elif bank is None:
        bank = mc_prepass(sd_x, sd_y, sd_r, pair, loss_type, mc_x0, mc_y0, num_steps)
    aux = {}
    for step in range(num_steps):
        t = step * dt
        if gamma_schedule is not None:
            guidance_strength = float(gamma_schedule[step])
        tb = torch.full((B,), t)
        v_x = unet_forward(sd_x, kx, x_t, tb)
        v_y = unet_forward(sd_y, ky, y_t, tb)
        if bank is not None and t > eps:
            g_x, g_y, w, Z = mc_feng_guidance(x_t, y_t, bank[0], bank[1], bank[2], t, eps)
            if return_aux and step == int(0.3 * num_steps):
                # diagnostics the 28x28 reference prints at this step (flow_utils.py:349-363)
                aux["diag"] = dict(
                    t=t, sigma_t=1 - t + eps,
                    v_x_norm=v_x.reshape(B, -1).norm(dim=1).mean().item(),
                    v_y_norm=v_y.reshape(B, -1).norm(dim=1).mean().item(),
                    g_x_norm=g_x.reshape(B, -1).norm(dim=1).mean().item(),
                    g_y_norm=g_y.reshape(B, -1).norm(dim=1).mean().item(),
                    w_min=w.min().item(), w_max=w.max().item(), z_bar=Z.mean().item())
            v_x = (1 - guidance_strength) * v_x + guidance_strength * g_x
            v_y = (1 - guidance_strength) * v_y + guidance_strength * g_y
        elif guidance_method == "grad_log_ratio" and sd_r is not None:
            # README.md:159-164: v_guided = v_ind + gamma * grad log r(x_t, y_t) (no reference implementation: UNPINNED)
            _, _, kr, _, _ = PAIR_KINDS[pair]
            g_x, g_y, _ = grad_log_ratio(sd_r, kr, x_t, y_t, loss_type)
            v_x = v_x + guidance_strength * g_x
            v_y = v_y + guidance_strength * g_y
        x_t = x_t + v_x * dt
        y_t = y_t + v_y * dt
    if return_aux:
        if bank is not None:
            aux["mc_x"], aux["mc_y"], aux["mc_ratios"] = bank
        return x_t, y_t, aux
    return x_t, y_t


@torch.no_grad()
def grad_log_ratio(sd_r: StateDict, kind: str, x: Tensor, y: Tensor, loss_type: str):
    """d/dx, d/dy of sum_b log r(x_b, y_b): the README-only guidance (README.md:159-164).  The reference has no
    implementation, so parity for this extension is UNPINNED; the oracle is autograd through the
    restated estimator above."""
    with torch.enable_grad():
        xg = x.detach().clone().requires_grad_(True)
        yg = y.detach().clone().requires_grad_(True)
        lr = ratio_log_ratio(sd_r, kind, xg, yg, loss_type)
        gx, gy = torch.autograd.grad(lr.sum(), (xg, yg))
    return gx, gy, lr.detach()


# ----------------------------------------------------------------------------
# Coherence evaluators' classifiers -- stock CNNs, restated functionally
# ----------------------------------------------------------------------------
def classifier_logits(sd: StateDict, kind: str, x: Tensor) -> Tensor:
    """reference: classifier.py:29-52 (mnist28), svhn_classifier.py:41-74 (svhn), :101-118 (mnist32)."""
    if kind == "clf_mnist28":
        h = F.max_pool2d(F.relu(F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"], padding=1)), 2)
        h = F.max_pool2d(F.relu(F.conv2d(h, sd["conv2.weight"], sd["conv2.bias"], padding=1)), 2)
    elif kind == "clf_mnist32":
        h = F.max_pool2d(F.relu(F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"], padding=1)), 2)
        h = F.max_pool2d(F.relu(F.conv2d(h, sd["conv2.weight"], sd["conv2.bias"], padding=1)), 2)
        h = F.relu(F.conv2d(h, sd["conv3.weight"], sd["conv3.bias"], padding=1))
    elif kind == "clf_svhn":
        h = x
        for i in (1, 2, 3, 4):
            h = F.conv2d(h, sd[f"conv{i}.weight"], sd[f"conv{i}.bias"], padding=1)
            h = F.relu(_bn_eval(sd, f"bn{i}.", h))
            if i <= 2:
                h = F.max_pool2d(h, 2)
    else:
        raise ValueError(kind)
    h = h.reshape(h.shape[0], -1)
    h = F.relu(F.linear(h, sd["fc1.weight"], sd["fc1.bias"]))
    return F.linear(h, sd["fc2.weight"], sd["fc2.bias"])

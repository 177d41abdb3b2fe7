"""nn.Module restatement of the reference's flow-matching U-Nets, for the training-step tests.

TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's training leg, never by the product package).  The product's
flow-matching step (ratio_guided_multimodal_fm_b200/fm_train.py) operates on the reference's OWN module objects: it
reroutes the 3x3 `nn.Conv2d` children of whatever module it is given.  The GPU box has no reference tree, so the tests need
a module with the same tree (child names = the reference's state_dict keys, so the synthetic state_dicts load strictly)
and the same forward.  reference: src/models/unet.py:36-278, src/models/unet_flexible.py:39-261 (identical structure;
the configurations are oracle.rgfm_oracle.UNET_CONFIGS).

Parity status: PINNED.  tests/test_oracle_golden.py checks this module against oracle.rgfm_oracle.unet_forward (itself
pinned bit-for-bit to the live reference by oracle/make_golden.py) at max-abs 0.0, forward and parameter gradients, and
oracle/make_golden.py checks it against the unmodified reference classes when /root/reference is present.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .rgfm_oracle import GN_GROUPS, UNET_CONFIGS, timestep_embedding


class ResBlock(nn.Module):  # unet.py:36-96, unet_flexible.py:39-85
    def __init__(self, cin, cout, tdim, dropout):
        super().__init__()
        self.norm1 = nn.GroupNorm(min(GN_GROUPS, cin), cin)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(tdim, cout))
        self.norm2 = nn.GroupNorm(min(GN_GROUPS, cout), cout)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.skip = nn.Conv2d(cin, cout, 1) if cin != cout else nn.Identity()

    def forward(self, x, t_emb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_mlp(t_emb)[:, :, None, None]
        h = self.conv2(self.dropout(F.silu(self.norm2(h))))
        return h + self.skip(x)


class Downsample(nn.Module):  # unet.py:99-107
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample(nn.Module):  # unet.py:110-119
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))


class UNet(nn.Module):  # unet.py:122-278, unet_flexible.py:111-261
    def __init__(self, kind: str, dropout: float = 0.1):
        super().__init__()
        cfg = UNET_CONFIGS[kind]
        self.kind = kind
        self.model_channels, self.channel_mult, self.num_res_blocks = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
        mc, mult, nres, cin = self.model_channels, self.channel_mult, self.num_res_blocks, cfg["in_channels"]
        tdim = mc * 4
        self.time_embed = nn.Sequential(nn.Linear(mc, tdim), nn.SiLU(), nn.Linear(tdim, tdim))
        self.input_conv = nn.Conv2d(cin, mc, 3, padding=1)
        self.encoder_blocks, self.downsamplers = nn.ModuleList(), nn.ModuleList()
        ch, enc = mc, [mc]
        for level, m in enumerate(mult):
            for _ in range(nres):
                self.encoder_blocks.append(ResBlock(ch, mc * m, tdim, dropout))
                ch = mc * m
                enc.append(ch)
            if level < len(mult) - 1:
                self.downsamplers.append(Downsample(ch))
                enc.append(ch)
        self.middle_block1 = ResBlock(ch, ch, tdim, dropout)
        self.middle_block2 = ResBlock(ch, ch, tdim, dropout)
        self.decoder_blocks, self.upsamplers = nn.ModuleList(), nn.ModuleList()
        for level, m in reversed(list(enumerate(mult))):
            for _ in range(nres + 1):
                self.decoder_blocks.append(ResBlock(ch + enc.pop(), mc * m, tdim, dropout))
                ch = mc * m
            if level > 0:
                self.upsamplers.append(Upsample(ch))
        self.out_norm = nn.GroupNorm(min(GN_GROUPS, ch), ch)
        self.out_conv = nn.Conv2d(ch, cin, 3, padding=1)

    def forward(self, x, t):
        t_emb = self.time_embed(timestep_embedding(t, self.model_channels))
        h = self.input_conv(x)
        hs = [h]
        bi = di = 0
        for level in range(len(self.channel_mult)):
            for _ in range(self.num_res_blocks):
                h = self.encoder_blocks[bi](h, t_emb)
                hs.append(h)
                bi += 1
            if level < len(self.channel_mult) - 1:
                h = self.downsamplers[di](h)
                hs.append(h)
                di += 1
        h = self.middle_block2(self.middle_block1(h, t_emb), t_emb)
        bi = ui = 0
        for level in reversed(range(len(self.channel_mult))):
            for _ in range(self.num_res_blocks + 1):
                h = self.decoder_blocks[bi](torch.cat([h, hs.pop()], dim=1), t_emb)
                bi += 1
            if level > 0:
                h = self.upsamplers[ui](h)
                ui += 1
        return self.out_conv(F.silu(self.out_norm(h)))


def fm_loss_and_grads(sd, kind: str, x_t, t, u, dtype=torch.float32):
    """loss = F.mse_loss(model(x_t, t), u) (flow_utils.py:139-142) and d loss / d parameter for every state_dict entry, by
    autograd through oracle.rgfm_oracle.unet_forward (functional, dropout = identity).  dtype=torch.float64 gives the
    yardstick that says how far ANY fp32 evaluation (the reference's own included) is from the exact gradient."""
    from .rgfm_oracle import unet_forward
    p = {k: v.detach().to(dtype).requires_grad_(True) for k, v in sd.items()}
    loss = F.mse_loss(unet_forward(p, kind, x_t.to(dtype), t.to(dtype), ), u.to(dtype))
    names = list(p)
    grads = torch.autograd.grad(loss, [p[k] for k in names])
    return loss.detach(), dict(zip(names, grads))

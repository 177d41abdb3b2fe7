"""Seeded synthetic weights with the reference's state_dict names and shapes.

There are no checkpoints and no network, so parity and throughput
runs use weights drawn here.  Keys and shapes follow the reference modules exactly
(`load_state_dict(strict=True)` into the reference classes succeeds -- checked by
`oracle/make_golden.py`), so the same dict can drive the reference, the oracle
and the CUDA path.  Fresh reference U-Nets output exactly zero (zero-initialised
`out_conv`); here `out_conv` is random so parity is not vacuous.

`StateDictModule` is a minimal stand-in for a loaded reference `nn.Module`
(`.state_dict()`, `.eval()`, `.loss_type`), used where the reference classes are
not importable (the GPU box).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict

import torch

UNET_ARCH = {
    # reference: unet.py:282-305, unet_flexible.py:266-291
    "unet_mnist28": dict(in_channels=1, img_size=28, model_channels=32, channel_mult=(1, 2), num_res_blocks=2),
    "unet_mnist32": dict(in_channels=1, img_size=32, model_channels=32, channel_mult=(1, 2), num_res_blocks=2),
    "unet_svhn": dict(in_channels=3, img_size=32, model_channels=64, channel_mult=(1, 2, 2), num_res_blocks=2),
    # `--model original` (flow_matching.py:129-173): not a U-Net; listed for its input geometry
    "flow28": dict(in_channels=1, img_size=28),
}


class _Gen:
    def __init__(self, seed: int):
        self.g = torch.Generator(device="cpu")
        self.g.manual_seed(seed)
        self.sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def uniform(self, name, shape, bound):
        self.sd[name] = (torch.rand(shape, generator=self.g) * 2 - 1) * bound

    def normal(self, name, shape, mean, std):
        self.sd[name] = torch.randn(shape, generator=self.g) * std + mean

    def conv(self, p, cout, cin, k, w_std=None):
        fan_in = cin * k * k
        b = 1.0 / math.sqrt(fan_in)
        if w_std is None:
            self.uniform(p + ".weight", (cout, cin, k, k), b)
        else:
            self.normal(p + ".weight", (cout, cin, k, k), 0.0, w_std)
        self.uniform(p + ".bias", (cout,), b)

    def linear(self, p, cout, cin):
        b = 1.0 / math.sqrt(cin)
        self.uniform(p + ".weight", (cout, cin), b)
        self.uniform(p + ".bias", (cout,), b)

    def affine(self, p, c):  # GroupNorm / LayerNorm affine ~ 1 + 0.1 N, bias ~ 0.1 N
        self.normal(p + ".weight", (c,), 1.0, 0.1)
        self.normal(p + ".bias", (c,), 0.0, 0.1)

    def bn(self, p, c):
        self.affine(p, c)
        self.normal(p + ".running_mean", (c,), 0.0, 0.1)
        self.sd[p + ".running_var"] = torch.rand((c,), generator=self.g) + 0.5
        self.sd[p + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.int64)


def _resblock(g: _Gen, p: str, cin: int, cout: int, temb: int):
    g.affine(p + ".norm1", cin)
    g.conv(p + ".conv1", cout, cin, 3)
    g.linear(p + ".time_mlp.1", cout, temb)
    g.affine(p + ".norm2", cout)
    g.conv(p + ".conv2", cout, cout, 3)
    if cin != cout:
        g.conv(p + ".skip", cout, cin, 1)


def unet_state_dict(kind: str, seed: int) -> Dict[str, torch.Tensor]:
    a = UNET_ARCH[kind]
    mc, mult, nres, cin0 = a["model_channels"], a["channel_mult"], a["num_res_blocks"], a["in_channels"]
    temb = mc * 4
    g = _Gen(seed)
    g.linear("time_embed.0", temb, mc)
    g.linear("time_embed.2", temb, temb)
    g.conv("input_conv", mc, cin0, 3)
    ch = mc
    enc_ch = [ch]
    bi = di = 0
    downs = []
    for level, m in enumerate(mult):
        for _ in range(nres):
            _resblock(g, f"encoder_blocks.{bi}", ch, mc * m, temb)
            ch = mc * m
            enc_ch.append(ch)
            bi += 1
        if level < len(mult) - 1:
            downs.append((di, ch))
            enc_ch.append(ch)
            di += 1
    for i, c in downs:
        g.conv(f"downsamplers.{i}.conv", c, c, 3)
    _resblock(g, "middle_block1", ch, ch, temb)
    _resblock(g, "middle_block2", ch, ch, temb)
    bi = ui = 0
    ups = []
    for level, m in reversed(list(enumerate(mult))):
        for _ in range(nres + 1):
            skip = enc_ch.pop()
            _resblock(g, f"decoder_blocks.{bi}", ch + skip, mc * m, temb)
            ch = mc * m
            bi += 1
        if level > 0:
            ups.append((ui, ch))
            ui += 1
    for i, c in ups:
        g.conv(f"upsamplers.{i}.conv", c, c, 3)
    g.affine("out_norm", ch)
    g.conv("out_conv", cin0, ch, 3, w_std=0.05)
    return g.sd


def ratio_state_dict(kind: str, seed: int) -> Dict[str, torch.Tensor]:
    g = _Gen(seed)
    if kind == "ratio28":  # reference: ratio_estimator.py:43-65,113-135
        for enc in ("encoder_x", "encoder_y"):
            for i, (ci, co) in enumerate(((1, 32), (32, 64), (64, 128), (128, 128)), start=1):
                g.conv(f"{enc}.conv{i}", co, ci, 3)
                g.affine(f"{enc}.gn{i}", co)
            g.linear(f"{enc}.fc", 256, 128)
        g.linear("score_net.0", 512, 512); g.affine("score_net.1", 512)
        g.linear("score_net.4", 256, 512); g.affine("score_net.5", 256)
        g.linear("score_net.8", 1, 256)
    elif kind == "ratio_mnist_svhn":  # reference: ratio_flexible.py:191-208,241-269,320-345
        for i, (ci, co) in enumerate(((1, 32), (32, 64), (64, 128), (128, 128)), start=1):
            g.conv(f"encoder_mnist.conv{i}", co, ci, 3)
            g.bn(f"encoder_mnist.bn{i}", co)
        g.linear("encoder_mnist.fc", 256, 128)
        chans = ((3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256), (256, 256))
        names = ("1a", "1b", "2a", "2b", "3a", "3b", "4a", "4b")
        for n, (ci, co) in zip(names, chans):
            g.conv(f"encoder_svhn.conv{n}", co, ci, 3)
            g.bn(f"encoder_svhn.bn{n}", co)
        g.linear("encoder_svhn.fc", 256, 256)
        g.linear("score_net.0", 512, 512); g.affine("score_net.1", 512)
        g.linear("score_net.4", 512, 512); g.affine("score_net.5", 512)
        g.linear("score_net.8", 256, 512); g.affine("score_net.9", 256)
        g.linear("score_net.11", 1, 256)
    else:
        raise ValueError(kind)
    return g.sd


def classifier_state_dict(kind: str, seed: int) -> Dict[str, torch.Tensor]:
    """reference: classifier.py:18-27, svhn_classifier.py:19-39,84-99."""
    g = _Gen(seed)
    if kind == "clf_mnist28":
        g.conv("conv1", 32, 1, 3); g.conv("conv2", 64, 32, 3)
        g.linear("fc1", 128, 64 * 7 * 7); g.linear("fc2", 10, 128)
    elif kind == "clf_mnist32":
        g.conv("conv1", 32, 1, 3); g.conv("conv2", 64, 32, 3); g.conv("conv3", 64, 64, 3)
        g.linear("fc1", 128, 64 * 8 * 8); g.linear("fc2", 10, 128)
    elif kind == "clf_svhn":
        for i, (ci, co) in enumerate(((3, 32), (32, 64), (64, 128), (128, 128)), start=1):
            g.conv(f"conv{i}", co, ci, 3)
            g.bn(f"bn{i}", co)
        g.linear("fc1", 256, 128 * 8 * 8); g.linear("fc2", 10, 256)
    else:
        raise ValueError(kind)
    return g.sd


def flow28_state_dict(seed: int) -> Dict[str, torch.Tensor]:
    """`--model original`: FlowMatchingModel (src/models/flow_matching.py:34-173): conv encoder -> fc -> [features | time]
    -> fc -> two ConvTranspose2d -> conv -> conv.  ConvTranspose2d weights are [Cin, Cout, 4, 4]."""
    g = _Gen(seed)
    g.conv("encoder.conv1", 32, 1, 3); g.affine("encoder.gn1", 32)
    g.conv("encoder.conv2", 64, 32, 3); g.affine("encoder.gn2", 64)
    g.conv("encoder.conv3", 128, 64, 3); g.affine("encoder.gn3", 128)
    g.conv("encoder.conv4", 256, 128, 3); g.affine("encoder.gn4", 256)
    g.linear("encoder.fc", 256, 256 * 7 * 7)
    g.linear("decoder.fc1", 256 * 7 * 7, 256 + 128)
    for name, cin, cout in (("decoder.deconv1", 256, 128), ("decoder.deconv2", 128, 64)):
        b = 1.0 / math.sqrt(cout * 16)  # torch: fan_in of a transposed conv is weight.size(1) * k * k
        g.uniform(name + ".weight", (cin, cout, 4, 4), b)
        g.uniform(name + ".bias", (cout,), b)
    g.affine("decoder.gn1", 128); g.affine("decoder.gn2", 64)
    g.conv("decoder.conv3", 32, 64, 3); g.affine("decoder.gn3", 32)
    g.conv("decoder.conv_out", 1, 32, 3)
    return g.sd


def state_dict(kind: str, seed: int) -> Dict[str, torch.Tensor]:
    if kind == "flow28":
        return flow28_state_dict(seed)
    if kind.startswith("unet"):
        return unet_state_dict(kind, seed)
    if kind.startswith("ratio"):
        return ratio_state_dict(kind, seed)
    return classifier_state_dict(kind, seed)


class StateDictModule:
    """Duck-typed stand-in for a loaded reference module: what the samplers consume is
    `.eval()`, `.state_dict()` and (ratio estimators) `.loss_type`."""

    def __init__(self, kind: str, sd: Dict[str, torch.Tensor], loss_type: str = "disc"):
        self.rgfm_kind = kind
        self._sd = sd
        self.loss_type = loss_type
        self.training = True

    def eval(self):
        self.training = False
        return self

    def state_dict(self):
        return self._sd


class TrainableModule(torch.nn.Module):
    """A real nn.Module whose parameters / buffers carry the reference's state_dict key names (nested containers), for the
    training-side entry points: `.parameters()` feed a stock optimiser, `.state_dict(keep_vars=True)` hands the live
    tensors to the trainer plan.  It has no forward of its own -- the forward / backward run in librgfm.so."""

    def __init__(self, kind: str, sd: Dict[str, torch.Tensor], loss_type: str = "disc"):
        super().__init__()
        self.rgfm_kind = kind
        self.loss_type = loss_type
        for k, v in sd.items():
            parts = k.split(".")
            m = self
            for part in parts[:-1]:
                if part not in m._modules:
                    m.add_module(part, torch.nn.Module())
                m = m._modules[part]
            if parts[-1].startswith("running_") or parts[-1] == "num_batches_tracked" or not v.dtype.is_floating_point:
                m.register_buffer(parts[-1], v.clone())
            else:
                m.register_parameter(parts[-1], torch.nn.Parameter(v.clone()))


def synthetic_module(kind: str, seed: int, loss_type: str = "disc") -> StateDictModule:
    return StateDictModule(kind, state_dict(kind, seed), loss_type)

"""Drop-in for the reference's `CFMSchedule` (src/utils/flow_utils.py:10-100): the rectified-flow path helpers and the
unguided single-modality Euler sampler.  `sample` runs on the CUDA library (rgfm_unet_sample); the path helpers are the
training-side formulas and stay plain torch."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .sampler import default_precision


class CFMSchedule:
    """x_t = (1 - t) x_0 + t x_1 with x_0 ~ N(0, I); conditional velocity u_t = x_1 - x_0 (flow_utils.py:10-22)."""

    def __init__(self, sigma: float = 0.0):
        self.sigma = sigma

    def compute_mu_t(self, x_0, x_1, t):
        t = t.view(-1, 1, 1, 1)  # flow_utils.py:31-34
        return (1 - t) * x_0 + t * x_1

    def compute_sigma_t(self, t):
        return self.sigma  # flow_utils.py:36-38

    def add_noise(self, x_1, t):
        """(x_t, u_t_target) for data x_1 at times t; draws the prior with randn_like (flow_utils.py:40-66)."""
        x_0 = torch.randn_like(x_1)
        tb = t.view(-1, 1, 1, 1)
        return (1 - tb) * x_0 + tb * x_1, x_1 - x_0

    def sample(self, model, num_samples, num_steps=100, device='cuda', *, precision: Optional[str] = None,
               chunk: int = 0):
        """Euler integration of dx/dt = model(x, t) from the prior, t = step / num_steps (flow_utils.py:69-100).
        Like the reference: puts the model in eval(), draws `torch.randn(num_samples, 1, 28, 28, device=device)` from
        the device's global generator, returns `[num_samples, 1, 28, 28]` on `device`.  `model` is the reference's
        28x28 `FlowMatchingUNet` or `FlowMatchingModel` (anything with .eval() and .state_dict() in their key layouts)."""
        model.eval()
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.RgfmError("rgfm samplers run on CUDA devices only; there is no CPU fallback")
        x_t = torch.randn(num_samples, 1, 28, 28, device=device)
        if num_samples == 0:
            return x_t
        sd = model.state_dict()
        # `--model original` (FlowMatchingModel, flow_matching.py) or the 28x28 U-Net: the keys say which
        kind = "flow28" if "encoder.conv1.weight" in sd else "unet_mnist28"
        net = _lib.Net(kind, sd,
                       max_batch=chunk or min(int(num_samples), 1024), precision=precision or default_precision(),
                       device=device)
        try:
            out = net.unet_sample(x_t, int(num_steps))
            torch.cuda.current_stream(device).synchronize()
        finally:
            net.close()
        return out

"""Drop-in for the reference's `CFMSchedule` (src/utils/flow_utils.py:10-100): the rectified-flow path helpers and the
unguided single-modality Euler sampler.  `sample` runs on the CUDA library (rgfm_unet_sample).  The training-side path
helpers run on the library for CUDA tensors -- `add_noise` is one fused streaming kernel
(rgfm_cfm_add_noise) and `fm_mse_loss` gives the flow-matching loss with its gradient (rgfm_fm_mse_loss) -- and stay plain
torch for CPU tensors (host-side data preparation, as in the reference)."""
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
        if x_1.is_cuda and x_1.dtype == torch.float32 and x_1[0].numel() % 4 == 0:
            x1c, tc = x_1.contiguous(), t.to(device=x_1.device, dtype=torch.float32).contiguous()
            x_t, u = torch.empty_like(x1c), torch.empty_like(x1c)
            _lib.check(_lib.lib().rgfm_cfm_add_noise(_lib._ptr(x1c), _lib._ptr(x_0.contiguous()), _lib._ptr(tc), x1c.shape[0],
                                                     x1c[0].numel(), _lib._ptr(x_t), _lib._ptr(u), _lib._stream_ptr(x_1.device)))
            return x_t, u
        tb = t.view(-1, 1, 1, 1)
        return (1 - tb) * x_0 + tb * x_1, x_1 - x_0

    @staticmethod
    def fm_mse_loss(v_t: torch.Tensor, u_t_target: torch.Tensor):
        """F.mse_loss(v_t, u_t_target) (flow_utils.py:142) and its gradient wrt v_t, in one pass on the device:
        returns (loss [scalar tensor], dv [like v_t]) -- what `loss.backward()` would seed the model's backward with."""
        dev = v_t.device
        _lib._dev_index(dev)
        v, u = _lib._f32c(v_t, dev), _lib._f32c(u_t_target, dev)
        if v.numel() % 4 != 0 or v.shape != u.shape:
            raise _lib.RgfmError("fm_mse_loss: equal shapes with a multiple of 4 elements")
        loss = torch.empty((), device=dev, dtype=torch.float32)
        dv = torch.empty_like(v)
        scratch = torch.empty(1184, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().rgfm_fm_mse_loss(_lib._ptr(v), _lib._ptr(u), v.numel(), _lib._ptr(loss), _lib._ptr(dv),
                                               _lib._ptr(scratch), _lib._stream_ptr(dev)))
        return loss, dv

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

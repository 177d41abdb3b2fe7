"""Training-side reuse: the ratio estimator's optimisation step on the B200 kernels.

`RatioTrainer` mirrors the reference trainer (src/utils/trainer.py:9-81: same constructor, `train_step(batch)` and
`train_epoch(dataloader)`, same metrics dict as src/utils/losses.py:27-133).  What changes is the middle of the step:
instead of `scores = model(x, y); loss = loss_fn(...); loss.backward()` through autograd, ONE C-ABI call
(`rgfm_ratio_trainer_step`) runs the estimator forward, the loss head and the whole backward -- input-gradient chain,
conv / linear / norm-affine weight gradients -- on the library's fp32 kernels, reading the module's live parameter tensors
in place and writing `param.grad`.  Gradient clipping and the optimiser stay stock PyTorch (host-side plumbing over the
same device tensors), exactly as the reference has them (trainer.py:53-54).

Scope, stated plainly:
* dropout (p = 0.1 in the score net, ratio_estimator.py:129,133) is applied from masks drawn HERE with
  `torch.bernoulli` on the device when the module is in training mode -- the same distribution as `nn.Dropout`, not the
  same random stream; `dropout_masks=` overrides them (the parity tests replay fixed masks into the reference);
* BatchNorm layers (RatioEstimatorMNISTSVHN) run with their running statistics (frozen), whatever `model.training` says;
  the GroupNorm estimator (RatioEstimator) has no such difference;
* there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import LOSS, NET_KINDS, RgfmError

# widths of the score-net layers that are followed by nn.Dropout (ratio_estimator.py:129,133; ratio_flexible.py:336,340)
_HIDDEN = {"ratio28": (512, 256), "ratio_mnist_svhn": (512, 512)}


def estimator_kind(model) -> str:
    """Which of the two estimators a module is, from its state_dict keys (ratio_estimator.py:96 / ratio_flexible.py:305)."""
    keys = model.state_dict().keys()
    if any(k.startswith("encoder_svhn.") for k in keys):
        return "ratio_mnist_svhn"
    if any(k.startswith("encoder_x.") for k in keys):
        return "ratio28"
    raise RgfmError("not a RatioEstimator / RatioEstimatorMNISTSVHN module")


def loss_kind(loss_fn) -> tuple:
    """(loss_type, lambda_penalty) of a losses.py loss object (DiscriminatorLoss | RuLSIFLoss) or of a name."""
    if isinstance(loss_fn, str):
        if loss_fn not in LOSS:
            raise ValueError(f"Unknown loss_type: {loss_fn}")
        return loss_fn, 0.1
    if hasattr(loss_fn, "lambda_penalty"):
        return "rulsif", float(loss_fn.lambda_penalty)
    return "disc", 0.0


def ratio_loss(scores: torch.Tensor, is_real: torch.Tensor, loss_type: str, lambda_penalty: float = 0.1):
    """losses.py:27-133 on the device: returns (out8, dscore) with out8 = {loss, m1, m2, m3, ...} (see rgfm.h)."""
    if loss_type not in LOSS:
        raise ValueError(f"Unknown loss_type: {loss_type}")
    dev = scores.device
    _lib._dev_index(dev)
    s = _lib._f32c(scores, dev)
    r = _lib._f32c(is_real, dev)
    out8 = torch.zeros(8, device=dev, dtype=torch.float32)
    ds = torch.empty_like(s)
    _lib.check(_lib.lib().rgfm_ratio_loss(_lib._ptr(s), _lib._ptr(r), s.numel(), LOSS[loss_type], float(lambda_penalty),
                                          _lib._ptr(out8), _lib._ptr(ds), _lib._stream_ptr(dev)))
    return out8, ds


class RatioTrainPlan:
    """Device plan of one estimator's training step: live parameters in, gradients out (rgfm_ratio_trainer_*)."""

    def __init__(self, kind: str, params: Dict[str, torch.Tensor], max_batch: int, device):
        self.kind, self.device, self.max_batch = kind, torch.device(device), int(max_batch)
        dev_index = _lib._dev_index(device)
        self.params = {k: v.detach() for k, v in params.items() if torch.is_tensor(v) and v.dtype.is_floating_point}
        trainable = [k for k in self.params if not (k.endswith("running_mean") or k.endswith("running_var"))]
        # one flat gradient arena; `grads[name]` are views of it (kept across steps: zero_grad(set_to_none) cannot free them)
        total = sum(self.params[k].numel() for k in trainable)
        self.flat_grad = torch.zeros(total, device=self.device, dtype=torch.float32)
        self.grads, off = {}, 0
        for k in trainable:
            n = self.params[k].numel()
            self.grads[k] = self.flat_grad[off:off + n].view(self.params[k].shape)
            off += n
        self._pw = _lib.DeviceWeights(self.params, self.device)
        self._gw = _lib.DeviceWeights(self.grads, self.device)
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().rgfm_ratio_trainer_create(NET_KINDS[kind], self._pw.ref(), self._gw.ref(), self.max_batch, dev_index,
                                                        C.byref(self.handle)))
        self.out8 = torch.zeros(8, device=self.device, dtype=torch.float32)

    def param_ptrs(self):
        return self._pw.ptrs

    def step(self, x, y, is_real, loss_type: str, lambda_penalty: float = 0.1, dropout_masks: Optional[Sequence[torch.Tensor]] = None):
        if loss_type not in LOSS:
            raise ValueError(f"Unknown loss_type: {loss_type}")
        B = x.shape[0]
        if B > self.max_batch:
            raise RgfmError(f"batch {B} exceeds the plan's max_batch {self.max_batch}")
        x, y = _lib._f32c(x, self.device), _lib._f32c(y, self.device)
        r = _lib._f32c(is_real, self.device)
        if y.shape[0] != B or r.numel() != B:
            raise RgfmError("x, y and is_real must have the same batch size")
        masks = [] if dropout_masks is None else [_lib._f32c(m, self.device) for m in dropout_masks]
        if masks:
            widths = _HIDDEN[self.kind]
            if len(masks) != len(widths) or any(tuple(m.shape) != (B, w) for m, w in zip(masks, widths)):
                raise RgfmError(f"dropout_masks must be {[(B, w) for w in widths]}")
        arr = (C.c_void_p * max(1, len(masks)))(*[m.data_ptr() for m in masks])
        _lib.check(_lib.lib().rgfm_ratio_trainer_step(self.handle, _lib._ptr(x), _lib._ptr(y), _lib._ptr(r), B, LOSS[loss_type],
                                                      float(lambda_penalty), arr, len(masks), _lib._ptr(self.out8),
                                                      _lib._stream_ptr(self.device)))
        return self.out8, self.grads

    def close(self):
        if self.handle:
            _lib.lib().rgfm_ratio_trainer_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def average_gradients(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """Data-parallel training's one exchange step: every rank runs the step on its own shard of the batch, then the
    gradients are averaged with ONE all-reduce over the plan's flat gradient arena (NCCL on GPUs: 3.8 / 13 MB for the two
    estimators), in place -- `param.grad` are views of it.  The reference trains on one device; this is the multi-GPU
    extension of its step (a collective only where the path has a real exchange)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return flat_grad
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.div_(world)
    return flat_grad


class RatioTrainer:
    """Drop-in for src/utils/trainer.py:9-81 (`RatioTrainer(model, loss_fn, optimizer, device)`).
    data_parallel=True: every rank calls train_step on its own shard; gradients are averaged over `group` (average_gradients)."""

    def __init__(self, model, loss_fn, optimizer, device="cuda", max_batch: int = 256, dropout_p: float = 0.1,
                 data_parallel: bool = False, group=None):
        self.model, self.loss_fn, self.optimizer, self.device = model, loss_fn, optimizer, device
        self.data_parallel, self.group = bool(data_parallel), group
        self.kind = estimator_kind(model)
        self.loss_type, self.lambda_penalty = loss_kind(loss_fn)
        self.max_batch, self.dropout_p = int(max_batch), float(dropout_p)
        self._plan: Optional[RatioTrainPlan] = None

    def _named(self) -> Dict[str, torch.Tensor]:
        return {k: v for k, v in self.model.state_dict(keep_vars=True).items() if torch.is_tensor(v) and v.dtype.is_floating_point}

    def _get_plan(self, B: int) -> RatioTrainPlan:
        named = self._named()
        ptrs = tuple(v.data_ptr() for v in named.values())
        if self._plan is None or self._plan.param_ptrs() != ptrs or B > self._plan.max_batch:
            if self._plan is not None:
                self._plan.close()
            self._plan = RatioTrainPlan(self.kind, named, max(B, self.max_batch), self.device)
        return self._plan

    def train_step(self, batch, dropout_masks: Optional[Sequence[torch.Tensor]] = None):
        """trainer.py:26-56.  Returns the metrics dict of the loss (losses.py:61-65 / :125-130)."""
        x, y, is_real = batch["x"].to(self.device), batch["y"].to(self.device), batch["is_real"].to(self.device)
        B = x.shape[0]
        plan = self._get_plan(B)
        if dropout_masks is None and self.model.training and self.dropout_p > 0:
            keep = 1.0 - self.dropout_p
            dropout_masks = [torch.bernoulli(torch.full((B, w), keep, device=plan.device)) / keep for w in _HIDDEN[self.kind]]
        self.optimizer.zero_grad()
        out8, grads = plan.step(x, y, is_real.float(), self.loss_type, self.lambda_penalty, dropout_masks)
        if self.data_parallel:
            average_gradients(plan.flat_grad, self.group)
        named = self._named()
        for k, g in grads.items():
            named[k].grad = g
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=1.0)
        self.optimizer.step()
        o = out8.tolist()
        if self.loss_type == "disc":
            return {"loss": o[0], "acc_real": o[1], "acc_fake": o[2]}
        return {"loss": o[0], "mean_w_real": o[1], "mean_w_fake": o[2], "constraint_term": o[3]}

    def train_epoch(self, dataloader):
        """trainer.py:58-81 (the progress bar is left to the caller)."""
        self.model.train()
        metrics_list: List[dict] = [self.train_step(batch) for batch in dataloader]
        return {k: float(np.mean([m[k] for m in metrics_list])) for k in metrics_list[0].keys()}

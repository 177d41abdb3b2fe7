"""Flow-matching training step on the library: mirror of `train_flow_matching_epoch`
(src/utils/flow_utils.py:103-156).

What runs where.  The step is `x_t, u = schedule.add_noise(x_1, t)`, `v = model(x_t, t)`, `loss = F.mse_loss(v, u)`,
`loss.backward()`, `optimizer.step()`.  The data side is two fused streaming kernels (flow.py: rgfm_cfm_add_noise,
rgfm_fm_mse_loss).  On the model side the U-Nets (unet.py, unet_flexible.py) spend > 95 % of their FLOPs in 3x3
convolutions -- forward, input gradient and weight gradient: those run on the library's fp32 kernels
(rgfm_conv3x3_forward / rgfm_conv3x3_backward: the implicit-GEMM conv_simt kernel, the same kernel on the flipped /
transposed filter, the tiled weight-gradient kernel with a fixed-order split-K reduction, exact fp32, deterministic)
through ONE torch.autograd.Function that takes the module's own parameter tensors in torch layout, and its GroupNorms run on
the library as well, in the convs' NHWC layout (rgfm_groupnorm_nhwc_forward / _backward; PyTorch's CUDA GroupNorm would
convert every channels_last tensor to NCHW and back).  The rest (SiLU, dropout, the time MLP's Linears, the 1x1 skip
convs, concat, nearest upsampling) stays stock PyTorch autograd on the same tensors, and the optimiser is the caller's.  The model object is
the reference's UNMODIFIED module: `library_convs(model)` reroutes the `forward` of its 3x3 `nn.Conv2d` and `nn.GroupNorm` children for
the duration of a `with` block, so parameters, state_dict keys, `.grad` fields and checkpoints are the reference's own.

There is no CPU fallback: CUDA tensors only."""
from __future__ import annotations

import contextlib
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .flow import CFMSchedule


def _cl(t: torch.Tensor) -> torch.Tensor:
    """fp32, NHWC memory (channels_last) -- a no-op for tensors this module produced itself."""
    return t.to(torch.float32).contiguous(memory_format=torch.channels_last)


class _Conv3x3(torch.autograd.Function):
    """nn.Conv2d(k=3, padding=1, stride 1 | 2) on the library; x [n,Cin,H,W], w [Cout,Cin,3,3], b [Cout] or None."""

    @staticmethod
    def forward(ctx, x, w, b, stride):
        dev = x.device
        _lib._dev_index(dev)
        xc, wc = _cl(x), w.to(torch.float32).contiguous()
        bc = None if b is None else b.to(torch.float32).contiguous()
        n, cin, H, W = xc.shape
        cout = wc.shape[0]
        y = torch.empty((n, cout, H // stride, W // stride), device=dev, dtype=torch.float32, memory_format=torch.channels_last)
        _lib.check(_lib.lib().rgfm_conv3x3_forward(_lib._ptr(xc), n, H, W, cin, _lib._ptr(wc), _lib._ptr(bc), cout, stride,
                                                   _lib._ptr(y), _lib._stream_ptr(dev)))
        ctx.save_for_backward(xc, wc)
        ctx.stride, ctx.has_bias = stride, b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, wc = ctx.saved_tensors
        dev = dy.device
        dyc = _cl(dy)
        n, cin, H, W = xc.shape
        cout = wc.shape[0]
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dx = torch.empty_like(xc) if need_x else None  # preserve_format: channels_last
        dw = torch.empty_like(wc) if need_w else None
        db = torch.empty(cout, device=dev, dtype=torch.float32) if need_b else None
        _lib.check(_lib.lib().rgfm_conv3x3_backward(_lib._ptr(xc), _lib._ptr(dyc), n, H, W, cin, _lib._ptr(wc), cout, ctx.stride,
                                                    _lib._ptr(dx), _lib._ptr(dw), _lib._ptr(db), _lib._stream_ptr(dev)))
        return dx, dw, db, None


def conv3x3(x, weight, bias=None, stride: int = 1):
    """Differentiable F.conv2d(x, weight, bias, stride=stride, padding=1) for 3x3 kernels on the library's kernels."""
    if not x.is_cuda:
        raise _lib.RgfmError("rgfm conv ops run on CUDA devices only; there is no CPU fallback")
    if tuple(weight.shape[2:]) != (3, 3) or weight.shape[1] != x.shape[1] or stride not in (1, 2):
        raise _lib.RgfmError("conv3x3: weight [Cout, Cin, 3, 3] matching x's channels, stride 1 or 2")
    if stride == 2 and (x.shape[2] % 2 or x.shape[3] % 2):
        raise _lib.RgfmError("conv3x3: stride 2 needs even image sizes")
    return _Conv3x3.apply(x, weight, bias, stride)


class _GroupNormNHWC(torch.autograd.Function):
    """nn.GroupNorm on channels_last tensors without leaving NHWC (PyTorch's CUDA GroupNorm converts to NCHW and back)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps):
        dev = x.device
        _lib._dev_index(dev)
        xc = _cl(x)
        ga, be = gamma.to(torch.float32).contiguous(), beta.to(torch.float32).contiguous()
        n, C, H, W = xc.shape
        y = torch.empty_like(xc)  # preserve_format: channels_last
        stats = torch.empty((n, groups, 2), device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().rgfm_groupnorm_nhwc_forward(_lib._ptr(xc), n, H * W, C, groups, _lib._ptr(ga), _lib._ptr(be), float(eps),
                                                          _lib._ptr(y), _lib._ptr(stats), _lib._stream_ptr(dev)))
        ctx.save_for_backward(xc, ga, stats)
        ctx.groups = groups
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, ga, stats = ctx.saved_tensors
        dev = dy.device
        dyc = _cl(dy)
        n, C, H, W = xc.shape
        dx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        dg = torch.empty_like(ga) if ctx.needs_input_grad[1] else None
        db = torch.empty_like(ga) if ctx.needs_input_grad[2] else None
        _lib.check(_lib.lib().rgfm_groupnorm_nhwc_backward(_lib._ptr(xc), _lib._ptr(dyc), n, H * W, C, ctx.groups, _lib._ptr(ga),
                                                           _lib._ptr(stats), _lib._ptr(dx), _lib._ptr(dg), _lib._ptr(db),
                                                           _lib._stream_ptr(dev)))
        return dx, dg, db, None, None


def group_norm_nhwc(x, num_groups: int, weight, bias, eps: float = 1e-5):
    """Differentiable F.group_norm(x, num_groups, weight, bias, eps) on the library, NHWC in and out."""
    if not x.is_cuda:
        raise _lib.RgfmError("rgfm norm ops run on CUDA devices only; there is no CPU fallback")
    return _GroupNormNHWC.apply(x, weight, bias, int(num_groups), float(eps))


def _is_lib_norm(m: nn.Module) -> bool:
    return (isinstance(m, nn.GroupNorm) and m.affine and m.num_channels % m.num_groups == 0
            and (m.num_channels // m.num_groups) % 4 == 0 and m.num_channels <= 1024)


def _is_lib_conv(m: nn.Module) -> bool:
    return (isinstance(m, nn.Conv2d) and m.kernel_size == (3, 3) and m.padding == (1, 1) and m.stride in ((1, 1), (2, 2))
            and m.dilation == (1, 1) and m.groups == 1 and m.padding_mode == "zeros")


@contextlib.contextmanager
def library_convs(model: nn.Module, norms: bool = True):
    """Inside the block every 3x3 nn.Conv2d of `model` (padding 1, stride 1 or 2) computes on the library, forward and
    backward; the module tree, its parameters and its own `forward` code (unet.py:216-278, unet_flexible.py:203-261)
    are untouched; with `norms` its nn.GroupNorm children compute on the library too, in the convs' NHWC layout.
    Yields the number of rerouted convolutions."""
    patched, nconv = [], 0
    try:
        for m in model.modules():
            if "forward" in m.__dict__:  # already rerouted (a nested block, or the caller's own instance-level forward): leave it
                nconv += 1 if _is_lib_conv(m) and getattr(m.forward, "__func__", None) is not None and m.forward.__func__.__name__ == "fwd" else 0
                continue
            if _is_lib_conv(m):
                def fwd(self, x):
                    return conv3x3(x, self.weight, self.bias, self.stride[0])
                m.forward = types.MethodType(fwd, m)
                patched.append(m)
                nconv += 1
            elif norms and _is_lib_norm(m):  # GroupNorm without leaving the convs' NHWC layout (4-D inputs only)
                def nfwd(self, x):
                    if x.dim() != 4:
                        return F.group_norm(x, self.num_groups, self.weight, self.bias, self.eps)
                    return group_norm_nhwc(x, self.num_groups, self.weight, self.bias, self.eps)
                m.forward = types.MethodType(nfwd, m)
                patched.append(m)
        yield nconv
    finally:
        for m in patched:
            m.__dict__.pop("forward", None)  # back to the class's own forward


class _FmMse(torch.autograd.Function):
    """F.mse_loss(v, u) with its gradient from the same pass (rgfm_fm_mse_loss)."""

    @staticmethod
    def forward(ctx, v, u):
        loss, dv = CFMSchedule.fm_mse_loss(v, u)
        ctx.save_for_backward(dv)
        ctx.shape = v.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (dv,) = ctx.saved_tensors
        return (dv * g).view(ctx.shape), None


def fm_mse(v, u):
    """loss = mean((v - u)^2) as an autograd node (both tensors are brought to NCHW order first: autograd carries the
    gradient back through that layout change)."""
    return _FmMse.apply(v.contiguous(), u.contiguous())


def average_param_grads(params, group=None):
    """Data-parallel exchange of a training step: the `.grad` tensors of `params` are averaged over the ranks of `group` with
    ONE all-reduce over a flat buffer (NCCL on GPUs; 3.6 MB for the MNIST U-Nets, 24 MB for the SVHN one) and copied back in
    place.  No process group, or a single rank: nothing happens.  The reference trains on one device; this is the multi-GPU
    extension of its step (a collective only where the path has a real exchange)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    if world > 1 and grads:
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        o = 0
        for g in grads:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()
    return world


def fm_step(model, x_1, schedule: CFMSchedule, t=None, data_parallel: bool = False, group=None, library_norms: bool = True):
    """One flow-matching loss evaluation with gradients left in `param.grad` (flow_utils.py:131-147 without the
    optimiser): returns the loss tensor.  `t` defaults to `torch.rand(B)` drawn like the reference.
    data_parallel=True: every rank calls this on its own shard of the batch; the gradients are averaged over `group`.
    library_norms: GroupNorm on the library in NHWC (faster: no layout conversions).  nn.Dropout then sees channels_last
    tensors and assigns the SAME random draws to elements in NHWC memory order -- the same distribution, not the same
    mask as stock PyTorch; library_norms=False keeps GroupNorm stock and reproduces the stock masks element for element."""
    B = x_1.shape[0]
    if t is None:
        t = torch.rand(B, device=x_1.device)
    x_t, u = schedule.add_noise(x_1, t)
    with library_convs(model, norms=library_norms):
        v = model(x_t, t)
        loss = fm_mse(v, u)
        loss.backward()
    if data_parallel:
        average_param_grads(list(model.parameters()), group)
    return loss


def train_flow_matching_epoch(model, dataloader, optimizer, schedule, device, modality='x', data_parallel: bool = False, group=None,
                              library_norms: bool = True):
    """Drop-in for flow_utils.py:103-156: same arguments, RNG draws (t, then the prior inside add_noise), loss and return
    value; progress goes through tqdm when it is installed.  data_parallel / group / library_norms: see fm_step (extensions)."""
    try:
        from tqdm import tqdm
    except ImportError:  # pragma: no cover
        def tqdm(it, **kw):
            return it
    model.train()
    total_loss, num_batches = 0.0, 0
    for batch in tqdm(dataloader, desc=f"Training FM_{modality}"):
        x_1 = batch[modality].to(device)
        optimizer.zero_grad()
        loss = fm_step(model, x_1, schedule, data_parallel=data_parallel, group=group, library_norms=library_norms)
        optimizer.step()
        total_loss += loss.item()
        num_batches += 1
    return total_loss / num_batches

"""ctypes binding of librgfm.so (include/rgfm.h).

The product path has no CPU fallback: if the CUDA library is missing or a call
fails, this module raises.  PyTorch is used only for device memory, streams and
torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# RGFM_LIB_PATH: developer override for the trace / experiment build (csrc/Makefile target `trace`)
LIB_PATH = os.environ.get("RGFM_LIB_PATH") or os.path.join(_HERE, "csrc", "librgfm.so")

ABI_VERSION = 1
DIAG_FLOATS = 16

NET_KINDS = {"unet_mnist28": 0, "unet_mnist32": 1, "unet_svhn": 2, "ratio28": 3, "ratio_mnist_svhn": 4,
             "clf_mnist28": 5, "clf_mnist32": 6, "clf_svhn": 7, "flow28": 8}
PAIRS = {"pair28": 0, "mnist_svhn": 1}
GUIDANCE = {"none": 0, "mc_feng": 1, "grad_log_ratio": 2}
LOSS = {"disc": 0, "rulsif": 1}
PRECISION = {"fp32": 0, "bf16": 1, "tf32": 2, "fp16": 3}


class RgfmError(RuntimeError):
    pass


class _Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("numel", C.c_int64)]


class _Weights(C.Structure):
    _fields_ = [("tensors", C.POINTER(_Tensor)), ("count", C.c_int32)]


class _SamplerDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("pair", C.c_int32), ("guidance", C.c_int32), ("loss_type", C.c_int32),
                ("precision", C.c_int32), ("num_samples", C.c_int32), ("mc_batch_size", C.c_int32),
                ("num_steps", C.c_int32), ("guidance_strength", C.c_double), ("device", C.c_int32),
                ("chunk", C.c_int32)]


class _ProfEntry(C.Structure):
    _fields_ = [("name", C.c_char * 24), ("ms", C.c_double), ("flops", C.c_double), ("bytes", C.c_double),
                ("launches", C.c_int64)]


_lib = None

# every symbol include/rgfm.h declares (checked by tests/test_abi.py without a GPU)
EXPORTS = [
    "rgfm_last_error", "rgfm_abi_version", "rgfm_sampler_create", "rgfm_sampler_run", "rgfm_sampler_run_host",
    "rgfm_sampler_set_strength", "rgfm_sampler_set_strength_schedule", "rgfm_sampler_run_reuse_bank", "rgfm_sampler_get_bank", "rgfm_sampler_profile_step", "rgfm_sampler_launch_count", "rgfm_sampler_destroy", "rgfm_net_create",
    "rgfm_unet_forward", "rgfm_unet_sample", "rgfm_ratio_log_ratio", "rgfm_ratio_grad_create", "rgfm_ratio_grad_create_prec", "rgfm_ratio_grad_log_ratio", "rgfm_coherence_matches", "rgfm_sampler_attach_evaluators", "rgfm_classifier_logits", "rgfm_net_destroy", "rgfm_mc_feng_step",
    "rgfm_ratio_trainer_create", "rgfm_ratio_trainer_step", "rgfm_ratio_trainer_destroy", "rgfm_ratio_loss", "rgfm_cfm_add_noise", "rgfm_fm_mse_loss", "rgfm_conv3x3_forward", "rgfm_conv3x3_backward", "rgfm_groupnorm_nhwc_forward", "rgfm_groupnorm_nhwc_backward",
]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RgfmError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    fp, vp = C.POINTER(C.c_float), C.c_void_p
    L.rgfm_last_error.restype = C.c_char_p
    L.rgfm_abi_version.restype = C.c_int
    L.rgfm_sampler_create.argtypes = [C.POINTER(_SamplerDesc), C.POINTER(_Weights), C.POINTER(_Weights),
                                      C.POINTER(_Weights), C.POINTER(vp)]
    L.rgfm_sampler_run.argtypes = [vp] + [vp] * 7 + [vp]
    L.rgfm_sampler_run_host.argtypes = [vp] + [vp] * 7
    L.rgfm_sampler_set_strength.argtypes = [vp, C.c_double, vp]
    L.rgfm_sampler_set_strength_schedule.argtypes = [vp, C.POINTER(C.c_double), C.c_int32, vp]
    L.rgfm_sampler_run_reuse_bank.argtypes = [vp] + [vp] * 5 + [vp]
    L.rgfm_sampler_get_bank.argtypes = [vp, vp, vp, vp, vp]
    L.rgfm_sampler_profile_step.argtypes = [vp, C.POINTER(_ProfEntry), C.c_int32, C.POINTER(C.c_int32)]
    L.rgfm_sampler_launch_count.argtypes = [vp]
    L.rgfm_sampler_launch_count.restype = C.c_int64
    L.rgfm_sampler_destroy.argtypes = [vp]
    L.rgfm_sampler_destroy.restype = None
    L.rgfm_net_create.argtypes = [C.c_int32, C.POINTER(_Weights), C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.rgfm_unet_forward.argtypes = [vp, vp, C.c_float, C.c_int32, vp, vp]
    L.rgfm_unet_sample.argtypes = [vp, vp, C.c_int32, C.c_int32, vp]
    L.rgfm_ratio_log_ratio.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, vp, vp, vp]
    L.rgfm_classifier_logits.argtypes = [vp, vp, C.c_int32, vp, vp]
    L.rgfm_ratio_grad_create.argtypes = [C.c_int32, C.POINTER(_Weights), C.c_int32, C.c_int32, C.POINTER(vp)]
    L.rgfm_ratio_grad_create_prec.argtypes = [C.c_int32, C.POINTER(_Weights), C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.rgfm_coherence_matches.argtypes = [vp, vp, vp, vp, C.c_int32, vp, vp, vp]
    L.rgfm_sampler_attach_evaluators.argtypes = [vp, vp, vp, C.c_int32]
    L.rgfm_ratio_grad_log_ratio.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, vp, vp, vp, vp]
    L.rgfm_net_destroy.argtypes = [vp]
    L.rgfm_net_destroy.restype = None
    L.rgfm_mc_feng_step.argtypes = [C.c_int32] * 4 + [vp] * 7 + [C.c_double] * 3 + [vp] * 5 + [vp]
    L.rgfm_ratio_trainer_create.argtypes = [C.c_int32, C.POINTER(_Weights), C.POINTER(_Weights), C.c_int32, C.c_int32, C.POINTER(vp)]
    L.rgfm_ratio_trainer_step.argtypes = [vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_float, C.POINTER(vp), C.c_int32, vp, vp]
    L.rgfm_ratio_trainer_destroy.argtypes = [vp]
    L.rgfm_ratio_trainer_destroy.restype = None
    L.rgfm_ratio_loss.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_float, vp, vp, vp]
    L.rgfm_cfm_add_noise.argtypes = [vp, vp, vp, C.c_int32, C.c_int64, vp, vp, vp]
    L.rgfm_fm_mse_loss.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp]
    L.rgfm_conv3x3_forward.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, C.c_int32, C.c_int32, vp, vp]
    L.rgfm_conv3x3_backward.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int32, C.c_int32, vp, vp, vp, vp]
    L.rgfm_groupnorm_nhwc_forward.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, C.c_float, vp, vp, vp]
    L.rgfm_groupnorm_nhwc_backward.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp]
    if L.rgfm_abi_version() != ABI_VERSION:
        raise RgfmError("librgfm.so ABI version mismatch; rebuild")
    _lib = L
    return L


def check(status: int):
    if status != 0:
        raise RgfmError(lib().rgfm_last_error().decode("utf-8", "replace"))


class Weights:
    """state_dict -> rgfm_weights (host fp32 views kept alive for the duration of the create call)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor]):
        self._keep = []
        items = []
        for k, v in state_dict.items():
            if not torch.is_tensor(v) or not v.dtype.is_floating_point:
                continue  # num_batches_tracked etc.
            a = np.ascontiguousarray(v.detach().to("cpu", torch.float32).numpy())
            self._keep.append(a)
            items.append((k.encode(), a))
        self.arr = (_Tensor * len(items))()
        for i, (k, a) in enumerate(items):
            self._keep.append(k)
            self.arr[i].name = k
            self.arr[i].data = a.ctypes.data_as(C.POINTER(C.c_float))
            self.arr[i].numel = a.size
        self.w = _Weights(self.arr, len(items))

    def ref(self):
        return C.byref(self.w)


class DeviceWeights:
    """name -> DEVICE fp32 tensor table as rgfm_weights (the training entry points read parameters and write gradients in
    place).  The tensors must stay alive and must not move while the handle built from this table exists."""

    def __init__(self, tensors: Dict[str, torch.Tensor], device):
        self._keep = []
        items = []
        for k, v in tensors.items():
            if not torch.is_tensor(v) or not v.dtype.is_floating_point:
                continue
            if v.dtype != torch.float32 or not v.is_contiguous() or v.device.type != "cuda" or v.device.index != _dev_index(device):
                raise RgfmError(f"training tensors must be contiguous fp32 on {device}: {k} is {v.dtype} on {v.device}")
            items.append((k.encode(), v))
        self.arr = (_Tensor * len(items))()
        for i, (k, v) in enumerate(items):
            self._keep.append(k)
            self._keep.append(v)
            self.arr[i].name = k
            self.arr[i].data = C.cast(C.c_void_p(v.data_ptr()), C.POINTER(C.c_float))
            self.arr[i].numel = v.numel()
        self.w = _Weights(self.arr, len(items))
        self.ptrs = tuple(v.data_ptr() for _, v in items)

    def ref(self):
        return C.byref(self.w)


def content_fingerprint(state_dict) -> tuple:
    """Cache key of a state_dict by CONTENT: (name, shape) of every floating tensor plus its L1 and L2 norms, computed
    with two fused foreach kernels and ONE device-to-host copy.  Plans hold a repacked COPY of the weights, so a key made
    of addresses / version counters is not enough: a re-used module (`load_state_dict` of the next checkpoint into
    recycled storage, `param.data.copy_()` EMA swaps) keeps both while the values change."""
    if state_dict is None:
        return ()
    items = [(k, v) for k, v in state_dict.items() if torch.is_tensor(v) and v.dtype.is_floating_point and v.numel() > 0]
    if not items:
        return ()
    with torch.no_grad():
        ts = [v.detach() for _, v in items]
        sig = torch.stack([n.double() for n in torch._foreach_norm(ts, 1)] + [n.double() for n in torch._foreach_norm(ts, 2)])
        sig = tuple(sig.cpu().tolist())
    return tuple((k, tuple(v.shape)) for k, v in items) + sig


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dev_index(device) -> int:
    d = torch.device(device)
    if d.type != "cuda":
        raise RgfmError(f"rgfm runs on CUDA devices only (got {d}); there is no CPU fallback")
    return d.index if d.index is not None else torch.cuda.current_device()


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


TRANSFORMS = {"identity": 0, "none": 0, "rotate90": 1, "rotate180": 2, "rotate270": 3, "invert": 4, "flip_h": 5, "flip_v": 6}


def coherence_matches(clf_a: "Net", clf_b: "Net", xa: torch.Tensor, xb: torch.Tensor, matches: torch.Tensor = None) -> torch.Tensor:
    """Number of pairs whose classifier predictions agree, as ONE device float that is added to (rgfm_coherence_matches):
    both classifier forwards, argmax and the comparison stay on the device."""
    dev = clf_a.device
    xa, xb = _f32c(xa, dev), _f32c(xb, dev)
    assert xa.shape[0] == xb.shape[0] and xa.shape[0] >= 1
    if matches is None:
        matches = torch.zeros(1, device=dev, dtype=torch.float32)
    scratch = torch.empty(20 * xa.shape[0], device=dev, dtype=torch.float32)
    check(lib().rgfm_coherence_matches(clf_a.h, clf_b.h, _ptr(xa), _ptr(xb), xa.shape[0], _ptr(scratch), _ptr(matches), _stream_ptr(dev)))
    return matches


class Net:
    """One network of the path behind the C ABI (component entry points)."""

    def __init__(self, kind: str, state_dict, max_batch: int, precision: str = "fp32", device="cuda", grad: bool = False):
        self.kind = kind
        self.max_batch = int(max_batch)
        self.device = torch.device("cuda", _dev_index(device))
        self.h = C.c_void_p()
        w = Weights(state_dict)
        if grad:  # forward + input-gradient program of a ratio estimator (fp32: exact; 16-bit: both conv chains on the tensor cores)
            check(lib().rgfm_ratio_grad_create_prec(NET_KINDS[kind], w.ref(), int(max_batch), PRECISION[precision], self.device.index,
                                                    C.byref(self.h)))
        else:
            check(lib().rgfm_net_create(NET_KINDS[kind], w.ref(), int(max_batch), PRECISION[precision],
                                        self.device.index, C.byref(self.h)))

    def grad_log_ratio(self, x: torch.Tensor, y: torch.Tensor, loss_type: str = "disc"):
        """(d log r/dx, d log r/dy, log r) -- README.md:159-164; oracle: autograd through the restated estimator."""
        if loss_type not in LOSS:
            raise ValueError(f"Unknown loss_type: {loss_type}")
        self._check_pair(x, y, "grad_log_ratio")
        x, y = _f32c(x, self.device), _f32c(y, self.device)
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        lr = torch.empty(x.shape[0], device=self.device, dtype=torch.float32)
        check(lib().rgfm_ratio_grad_log_ratio(self.h, _ptr(x), _ptr(y), x.shape[0], LOSS[loss_type], _ptr(gx), _ptr(gy), _ptr(lr),
                                              _stream_ptr(self.device)))
        return gx, gy, lr

    _SHAPES = {"unet_mnist28": (1, 28, 28), "unet_mnist32": (1, 32, 32), "unet_svhn": (3, 32, 32), "flow28": (1, 28, 28),
               "clf_mnist28": (1, 28, 28), "clf_mnist32": (1, 32, 32), "clf_svhn": (3, 32, 32)}
    _RATIO_SHAPES = {"ratio28": ((1, 28, 28), (1, 28, 28)), "ratio_mnist_svhn": ((1, 32, 32), (3, 32, 32))}

    def _check_images(self, x: torch.Tensor, want, what: str):
        """The kernels take raw pointers: a wrong shape would read out of bounds instead of raising."""
        if x.dim() != 4 or tuple(x.shape[1:]) != tuple(want):
            raise RgfmError(f"{what}: expected [B, {want[0]}, {want[1]}, {want[2]}] for {self.kind}, got {tuple(x.shape)}")

    def unet_forward(self, x: torch.Tensor, t: float) -> torch.Tensor:
        self._check_images(x, self._SHAPES[self.kind], "unet_forward")
        x = _f32c(x, self.device)
        v = torch.empty_like(x)
        check(lib().rgfm_unet_forward(self.h, _ptr(x), float(t), x.shape[0], _ptr(v), _stream_ptr(self.device)))
        return v

    def unet_sample(self, x0: torch.Tensor, num_steps: int = 100) -> torch.Tensor:
        """CFMSchedule.sample's loop (flow_utils.py:69-100) from the prior noise x0; returns a fresh tensor."""
        self._check_images(x0, self._SHAPES[self.kind], "unet_sample")
        x = _f32c(x0, self.device).clone()
        check(lib().rgfm_unet_sample(self.h, _ptr(x), x.shape[0], int(num_steps), _stream_ptr(self.device)))
        return x

    def log_ratio(self, x: torch.Tensor, y: torch.Tensor, loss_type: str = "disc", return_scores: bool = False):
        if loss_type not in LOSS:
            raise ValueError(f"Unknown loss_type: {loss_type}")  # ratio_estimator.py:190-191
        self._check_pair(x, y, "log_ratio")
        x, y = _f32c(x, self.device), _f32c(y, self.device)
        out = torch.empty(x.shape[0], device=self.device, dtype=torch.float32)
        sc = torch.empty_like(out)
        check(lib().rgfm_ratio_log_ratio(self.h, _ptr(x), _ptr(y), x.shape[0], LOSS[loss_type], _ptr(out), _ptr(sc),
                                         _stream_ptr(self.device)))
        return (out, sc) if return_scores else out

    def _check_pair(self, x, y, what):
        sx, sy = self._RATIO_SHAPES[self.kind]
        self._check_images(x, sx, what)
        self._check_images(y, sy, what)
        if x.shape[0] != y.shape[0]:
            raise RgfmError(f"{what}: x and y batch sizes differ ({x.shape[0]} vs {y.shape[0]})")

    def classifier_logits(self, x: torch.Tensor) -> torch.Tensor:
        """logits [B, 10] of an evaluator classifier (evaluate_mnist_svhn.py:45-47, evaluate.py:77-80); exact fp32."""
        self._check_images(x, self._SHAPES[self.kind], "classifier_logits")
        x = _f32c(x, self.device)
        if x.shape[0] > self.max_batch:
            raise RgfmError(f"batch {x.shape[0]} exceeds the net's max_batch {self.max_batch}")
        out = torch.empty(x.shape[0], 10, device=self.device, dtype=torch.float32)
        check(lib().rgfm_classifier_logits(self.h, _ptr(x), x.shape[0], _ptr(out), _stream_ptr(self.device)))
        return out

    def close(self):
        if self.h:
            lib().rgfm_net_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SamplerPlan:
    """rgfm_sampler: repacked weights + activation arenas + CUDA graphs for one (shapes, device) configuration."""

    def __init__(self, pair: str, sd_x, sd_y, sd_ratio, *, guidance_method: str, guidance_strength: float,
                 num_samples: int, num_steps: int, mc_batch_size: int, loss_type: str = "disc",
                 precision: str = "fp32", device="cuda", chunk: int = 0):
        self.device = torch.device("cuda", _dev_index(device))
        self.pair = pair
        self.B, self.N, self.steps = int(num_samples), int(mc_batch_size), int(num_steps)
        # unknown guidance methods silently degrade to unguided like the reference (flow_utils.py:231,273)
        g = GUIDANCE.get(guidance_method, 0)
        self.guided = (g == 1 and sd_ratio is not None)          # needs the MC noise
        self.grad_guided = (g == 2 and sd_ratio is not None)
        if sd_ratio is not None and g in (1, 2) and loss_type not in LOSS:
            raise ValueError(f"Unknown loss_type: {loss_type}")
        d = _SamplerDesc(ABI_VERSION, PAIRS[pair], g, LOSS.get(loss_type, 0), PRECISION[precision], self.B, self.N,
                         self.steps, float(guidance_strength), self.device.index, int(chunk))
        wx, wy = Weights(sd_x), Weights(sd_y)
        wr = Weights(sd_ratio) if sd_ratio is not None else None
        self.h = C.c_void_p()
        check(lib().rgfm_sampler_create(C.byref(d), wx.ref(), wy.ref(), wr.ref() if wr else None, C.byref(self.h)))
        self.guidance_strength = float(guidance_strength)
        self.schedule = None
        self.shape_x, self.shape_y = {"pair28": ((1, 28, 28), (1, 28, 28)),
                                      "mnist_svhn": ((1, 32, 32), (3, 32, 32))}[pair]

    @property
    def launch_count(self) -> int:
        return int(lib().rgfm_sampler_launch_count(self.h))

    def attach_evaluators(self, clf_x: "Net" = None, clf_y: "Net" = None, transform_type: str = None):
        """Fuse the coherence evaluator into the plan's tail (evaluate_mnist_svhn.py:28-57 / evaluate.py:56-90): every run
        classifies its finished pairs on the device and reports the number of coherent pairs in diag[15].
        transform_type: the 28x28 pair's transform, undone on y before its classifier.  None, None detaches."""
        code = TRANSFORMS.get(transform_type or "identity")
        if code is None:
            raise ValueError(f"Unknown transform_type: {transform_type}")
        check(lib().rgfm_sampler_attach_evaluators(self.h, clf_x.h if clf_x else None, clf_y.h if clf_y else None, code))
        self._evaluators = (clf_x, clf_y)  # keep the nets alive as long as the plan uses them

    def set_strength(self, guidance_strength: float):
        """Change gamma in place: the graphs read it from a device table."""
        check(lib().rgfm_sampler_set_strength(self.h, float(guidance_strength), _stream_ptr(self.device)))
        self.guidance_strength = float(guidance_strength)
        self.schedule = None

    def set_strength_schedule(self, gammas):
        """Time-dependent guidance: one strength per Euler step (gamma_k at t_k = k / num_steps)."""
        g = [float(v) for v in gammas]
        arr = (C.c_double * len(g))(*g)
        check(lib().rgfm_sampler_set_strength_schedule(self.h, arr, len(g), _stream_ptr(self.device)))
        self.schedule = tuple(g)

    def run(self, x0, y0, mc_x0=None, mc_y0=None, want_diag: bool = False, reuse_bank: bool = False):
        """Device-resident run on the current stream; returns fresh tensors (inputs are not mutated).
        reuse_bank: main loop only, guided by the MC bank of the previous full run on this plan."""
        dev = self.device
        x0, y0 = _f32c(x0, dev), _f32c(y0, dev)
        assert tuple(x0.shape) == (self.B,) + self.shape_x and tuple(y0.shape) == (self.B,) + self.shape_y
        if reuse_bank:
            x1, y1 = torch.empty_like(x0), torch.empty_like(y0)
            diag = torch.zeros(DIAG_FLOATS, device=dev, dtype=torch.float32)
            check(lib().rgfm_sampler_run_reuse_bank(self.h, _ptr(x0), _ptr(y0), _ptr(x1), _ptr(y1), _ptr(diag),
                                                    _stream_ptr(dev)))
            return (x1, y1, diag) if want_diag else (x1, y1)
        if self.guided:
            mc_x0, mc_y0 = _f32c(mc_x0, dev), _f32c(mc_y0, dev)
            assert mc_x0.shape[0] == self.N and mc_y0.shape[0] == self.N
        else:
            mc_x0 = mc_y0 = None
        x1, y1 = torch.empty_like(x0), torch.empty_like(y0)
        diag = torch.zeros(DIAG_FLOATS, device=dev, dtype=torch.float32)
        check(lib().rgfm_sampler_run(self.h, _ptr(x0), _ptr(y0), _ptr(mc_x0), _ptr(mc_y0), _ptr(x1), _ptr(y1),
                                     _ptr(diag), _stream_ptr(dev)))
        return (x1, y1, diag) if want_diag else (x1, y1)

    def run_host(self, x0: np.ndarray, y0: np.ndarray, mc_x0=None, mc_y0=None):
        """Host buffers in, host buffers out (H2D + run + D2H inside the call); synchronous."""
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
        x0, y0, mc_x0, mc_y0 = f(x0), f(y0), f(mc_x0), f(mc_y0)
        x1, y1 = np.empty_like(x0), np.empty_like(y0)
        diag = np.zeros(DIAG_FLOATS, np.float32)
        p = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
        check(lib().rgfm_sampler_run_host(self.h, p(x0), p(y0), p(mc_x0), p(mc_y0), p(x1), p(y1), p(diag)))
        return x1, y1, diag

    def profile_step(self):
        """{kernel family: dict(ms, flops, bytes, launches)} for one main-loop step, CUDA-event timed."""
        buf = (_ProfEntry * 32)()
        n = C.c_int32(0)
        check(lib().rgfm_sampler_profile_step(self.h, buf, 32, C.byref(n)))
        return {buf[i].name.decode(): dict(ms=buf[i].ms, flops=buf[i].flops, bytes=buf[i].bytes,
                                           launches=int(buf[i].launches)) for i in range(n.value)}

    def bank(self):
        mx = torch.empty((self.N,) + self.shape_x, device=self.device)
        my = torch.empty((self.N,) + self.shape_y, device=self.device)
        r = torch.empty(self.N, device=self.device)
        check(lib().rgfm_sampler_get_bank(self.h, _ptr(mx), _ptr(my), _ptr(r), _stream_ptr(self.device)))
        return mx, my, r

    def close(self):
        if self.h:
            lib().rgfm_sampler_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mc_feng_step(x, y, v_x, v_y, mc_x, mc_y, mc_ratios, t: float, dt: float, gamma: float):
    """One guidance block + blend + Euler update.  Returns x_new, y_new, g_x, g_y, weights."""
    dev = x.device
    B, N = x.shape[0], mc_x.shape[0]
    args = [_f32c(a, dev) for a in (x, y, v_x, v_y, mc_x, mc_y, mc_ratios)]
    dx, dy = x[0].numel(), y[0].numel()
    xo, yo = torch.empty_like(args[0]), torch.empty_like(args[1])
    gx, gy = torch.empty_like(args[0]), torch.empty_like(args[1])
    w = torch.empty(B, N, device=dev)
    check(lib().rgfm_mc_feng_step(B, N, dx, dy, *[_ptr(a) for a in args], float(t), float(dt), float(gamma),
                                  _ptr(xo), _ptr(yo), _ptr(gx), _ptr(gy), _ptr(w), _stream_ptr(dev)))
    return xo, yo, gx, gy, w

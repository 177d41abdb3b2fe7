"""CPU: the C-ABI library loads, exports every symbol include/rgfm.h declares, and fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT
from ratio_guided_multimodal_fm_b200 import _lib, synthetic


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rgfm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rgfm_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_header_symbols():
    L = _lib.lib()
    syms = header_symbols()
    assert sorted(_lib.EXPORTS) == syms
    for s in syms:
        assert hasattr(L, s), f"librgfm.so does not export {s}"
    assert L.rgfm_abi_version() == _lib.ABI_VERSION


def test_desc_struct_layout_matches_header():
    # int32 x 8, double, int32 x 2 with natural alignment = 48 bytes
    assert C.sizeof(_lib._SamplerDesc) == 48
    assert C.sizeof(_lib._Tensor) == 24


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    sd = synthetic.state_dict("unet_mnist28", 1)
    with pytest.raises(_lib.RgfmError):
        _lib.Net("unet_mnist28", sd, 2, "fp32", device="cpu")
    w = _lib.Weights(sd)
    h = C.c_void_p()
    rc = _lib.lib().rgfm_net_create(0, w.ref(), 2, 0, 0, C.byref(h))
    assert rc != 0 and len(_lib.lib().rgfm_last_error()) > 0


def test_cpu_device_is_rejected():
    from ratio_guided_multimodal_fm_b200 import sample_bimodal_guided, synthetic_module
    fx, fy = synthetic_module("unet_mnist28", 1), synthetic_module("unet_mnist28", 2)
    with pytest.raises(_lib.RgfmError):
        sample_bimodal_guided(fx, fy, num_samples=2, num_steps=2, device="cpu")
    assert fx.training is False and fy.training is False  # modules are put in eval mode like the reference


def test_weights_marshalling_skips_integer_buffers():
    sd = synthetic.state_dict("ratio_mnist_svhn", 5)
    w = _lib.Weights(sd)
    names = {w.arr[i].name.decode() for i in range(w.w.count)}
    assert "encoder_svhn.bn1a.running_var" in names
    assert not any(n.endswith("num_batches_tracked") for n in names)
    assert w.w.count == 102 - 12

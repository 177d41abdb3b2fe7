"""Import the UNMODIFIED reference (foubari/ratio_guided_Multimodal_FM) for benchmarking / validation.

TEST / BENCH INFRASTRUCTURE ONLY (used by bench.py's `--impl reference` arm and cpu_baseline leg, and by
oracle/make_golden.py's pattern): looks for the reference package in oracle/_ref (built by oracle/build_ref.py, travels to
the GPU box) and then in /root/reference (authoring container).  matplotlib is absent in this image and only used by the
reference for PNG output, so empty stand-in modules are registered before the import.
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
CANDIDATES = (os.path.join(ROOT, "oracle", "_ref"), "/root/reference")


def reference_root():
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "src", "sample_mnist_svhn.py")):
            return c
    return None


def load():
    """Returns a namespace with the reference's own classes / functions, or None when no reference tree is available."""
    root = reference_root()
    if root is None:
        return None
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    if root not in sys.path:
        sys.path.insert(0, root)
    ns = types.SimpleNamespace(root=root)
    from src.models.unet import FlowMatchingUNet
    from src.models.unet_flexible import FlowMatchingUNetMNIST, FlowMatchingUNetSVHN
    from src.models.ratio_estimator import RatioEstimator
    from src.models.ratio_flexible import RatioEstimatorMNISTSVHN
    from src.utils import set_seed
    from src.utils.flow_utils import sample_bimodal_guided
    from src.sample_mnist_svhn import sample_bimodal_guided_mnist_svhn
    ns.classes = {
        "unet_mnist28": lambda lt="disc": FlowMatchingUNet(),
        "unet_mnist32": lambda lt="disc": FlowMatchingUNetMNIST(img_size=32),
        "unet_svhn": lambda lt="disc": FlowMatchingUNetSVHN(),
        "ratio28": lambda lt="disc": RatioEstimator(loss_type=lt),
        "ratio_mnist_svhn": lambda lt="disc": RatioEstimatorMNISTSVHN(loss_type=lt),
    }
    ns.set_seed = set_seed
    ns.sample_bimodal_guided = sample_bimodal_guided
    ns.sample_bimodal_guided_mnist_svhn = sample_bimodal_guided_mnist_svhn
    return ns


def module(ns, kind: str, state_dict, loss_type: str = "disc"):
    """The reference's own nn.Module of `kind` with `state_dict` loaded strictly, in eval mode."""
    m = ns.classes[kind](loss_type)
    m.load_state_dict(state_dict, strict=True)
    return m.eval()

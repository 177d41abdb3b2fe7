"""Drop-in coherence evaluators.

Same names, arguments and return dicts as the reference:
  evaluate_coherence             /root/reference/src/evaluate.py:56-90           (MNIST, transformed MNIST)
  evaluate_coherence_mnist_svhn  /root/reference/src/evaluate_mnist_svhn.py:28-57 (named evaluate_coherence there)
The classifier forwards run on librgfm.so (exact fp32 CUDA kernels, rgfm_classifier_logits); argmax and the match
count stay on the device and ONE scalar crosses to the host.  With `shard=True` every rank evaluates its contiguous
slice of the pairs and the match counts are summed with an all_reduce (the only collective besides the pair gather).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from .sampler import shard_bounds

_CLF_CACHE: Dict[tuple, "_lib.Net"] = {}
_CHUNK = 1024


def _kind_of(classifier, shape) -> str:
    kind = getattr(classifier, "rgfm_kind", None)
    if kind is not None:
        return kind
    name = type(classifier).__name__
    if name == "SVHNClassifier" or shape[1] == 3:
        return "clf_svhn"
    return "clf_mnist28" if shape[-1] == 28 else "clf_mnist32"


def _net(classifier, shape, device) -> "_lib.Net":
    sd = classifier.state_dict()
    key = (_kind_of(classifier, shape), str(device), _lib.content_fingerprint(sd))  # by content: see _lib.content_fingerprint
    net = _CLF_CACHE.get(key)
    if net is None:
        if len(_CLF_CACHE) >= 4:
            _CLF_CACHE.pop(next(iter(_CLF_CACHE))).close()
        net = _lib.Net(key[0], sd, max_batch=_CHUNK, precision="fp32", device=device)
        _CLF_CACHE[key] = net
    return net


def classifier_predictions(classifier, samples: torch.Tensor, device) -> torch.Tensor:
    """argmax of the classifier logits, [N] int64 on `device`."""
    classifier.eval()
    device = torch.device(device)
    net = _net(classifier, samples.shape, device)
    preds = []
    for i in range(0, samples.shape[0], _CHUNK):
        preds.append(net.classifier_logits(samples[i:i + _CHUNK].to(device)).argmax(dim=1))
    return torch.cat(preds) if preds else torch.empty(0, dtype=torch.int64, device=device)


def _matches(clf_a, clf_b, xa: torch.Tensor, xb: torch.Tensor, device) -> torch.Tensor:
    """#{i : argmax clf_a(xa_i) == argmax clf_b(xb_i)} as one device scalar: classifier forwards, argmax and the comparison in
    one C-ABI call per chunk (rgfm_coherence_matches), nothing but the final count crosses to the host."""
    clf_a.eval()
    clf_b.eval()
    device = torch.device(device)
    cnt = torch.zeros(1, device=device, dtype=torch.float32)
    if xa.shape[0] == 0:
        return cnt
    na, nb = _net(clf_a, xa.shape, device), _net(clf_b, xb.shape, device)
    for i in range(0, xa.shape[0], _CHUNK):
        _lib.coherence_matches(na, nb, xa[i:i + _CHUNK].to(device), xb[i:i + _CHUNK].to(device), cnt)
    return cnt


def get_inverse_transform(transform_type):
    """evaluate.py:30-53.  The reference uses torchvision's rotate / flips; for these exact quarter turns they are
    pixel permutations (TF.rotate(img, 90) is a counter-clockwise quarter turn = torch.rot90(k=1))."""
    if transform_type == "rotate90":
        return lambda img: torch.rot90(img, 1, dims=(-2, -1))
    if transform_type == "rotate180":
        return lambda img: torch.rot90(img, 2, dims=(-2, -1))
    if transform_type == "rotate270":
        return lambda img: torch.rot90(img, -1, dims=(-2, -1))
    if transform_type == "invert":
        return lambda img: -img
    if transform_type == "flip_h":
        return lambda img: torch.flip(img, dims=(-1,))
    if transform_type == "flip_v":
        return lambda img: torch.flip(img, dims=(-2,))
    return lambda img: img


def _coherence(cnt: torch.Tensor, total: int, shard: bool, group) -> Dict[str, float]:
    cnt = cnt.to(torch.int64).reshape(1)
    if shard:
        import torch.distributed as dist
        dist.all_reduce(cnt, group=group)
    return {"coherence_acc": float(cnt.item()) / float(max(total, 1)), "num_samples": int(total)}


def evaluate_coherence(samples_x, samples_y, classifier, device, transform_type="rotate90", *, shard: bool = False,
                       group=None):
    """reference: evaluate.py:56-90.  Returns {'coherence_acc', 'num_samples'}."""
    n = samples_x.shape[0]
    lo, hi = 0, n
    if shard:
        import torch.distributed as dist
        lo, hi = shard_bounds(n, dist.get_world_size(group), dist.get_rank(group))
    inv = get_inverse_transform(transform_type)
    return _coherence(_matches(classifier, classifier, samples_x[lo:hi], inv(samples_y[lo:hi]).contiguous(), device), n, shard, group)


def evaluate_coherence_mnist_svhn(samples_mnist, samples_svhn, mnist_classifier, svhn_classifier, device, *,
                                  shard: bool = False, group=None):
    """reference: evaluate_mnist_svhn.py:28-57 (called evaluate_coherence there)."""
    n = samples_mnist.shape[0]
    lo, hi = 0, n
    if shard:
        import torch.distributed as dist
        lo, hi = shard_bounds(n, dist.get_world_size(group), dist.get_rank(group))
    return _coherence(_matches(mnist_classifier, svhn_classifier, samples_mnist[lo:hi], samples_svhn[lo:hi], device), n, shard, group)

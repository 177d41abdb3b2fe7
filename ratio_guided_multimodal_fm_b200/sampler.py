"""Drop-in Python entry points of the ratio-guided coupled Euler sampler.

Same names, arguments, defaults, RNG consumption and return values as the reference:
  sample_bimodal_guided             /root/reference/src/utils/flow_utils.py:178-375
  sample_bimodal_guided_mnist_svhn  /root/reference/src/sample_mnist_svhn.py:39-177
The body is one call into librgfm.so (C ABI, include/rgfm.h).  Optional keyword-only
extras (precision, shard, plan cache) never have to be passed.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch

from . import _lib

_PLAN_CACHE: Dict[tuple, "_lib.SamplerPlan"] = {}
_PLAN_CACHE_MAX = 4


def default_precision() -> str:
    return os.environ.get("RGFM_PRECISION", "fp16")


def _fingerprint(module) -> tuple:
    return () if module is None else _lib.content_fingerprint(module.state_dict())


def _get_plan(pair, fm_x, fm_y, ratio, method, gamma, B, steps, N, loss_type, precision, device, chunk, schedule=None):
    """Plans are cached by weight CONTENT (see _lib.content_fingerprint) and shape.  gamma is not part of the key: every
    guidance kernel reads it from a device table, so a hit only rewrites that table (rgfm_sampler_set_strength, or
    rgfm_sampler_set_strength_schedule for a time-dependent gamma(t))."""
    key = (pair, _fingerprint(fm_x), _fingerprint(fm_y), _fingerprint(ratio), method, B, steps, N,
           loss_type, precision, str(device), chunk)
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE))).close()
        plan = _lib.SamplerPlan(pair, fm_x.state_dict(), fm_y.state_dict(),
                                ratio.state_dict() if ratio is not None else None,
                                guidance_method=method, guidance_strength=gamma, num_samples=B, num_steps=steps,
                                mc_batch_size=N, loss_type=loss_type, precision=precision, device=device, chunk=chunk)
        _PLAN_CACHE[key] = plan
    if schedule is not None:
        if plan.schedule != tuple(schedule):
            plan.set_strength_schedule(schedule)
    elif plan.schedule is not None or plan.guidance_strength != float(gamma):
        plan.set_strength(gamma)
    return plan


def _check_saturation(diag, precision):
    """diag[12] counts output rows of the tensor-core convs whose fp16 store hit +-65504 (include/rgfm.h).  Activations
    are GroupNorm-bounded, so this stays 0; with weights that do drive a layer past the fp16 range the clamp would
    corrupt samples silently -- raise instead (RGFM_ALLOW_SATURATION=1 downgrades to a warning)."""
    if diag[12] > 0:
        msg = (f"{int(diag[12])} activation rows saturated at the fp16 bound in precision={precision!r}; "
               "samples are not trustworthy -- use precision='fp32'")
        if os.environ.get("RGFM_ALLOW_SATURATION") == "1":
            import warnings
            warnings.warn(msg)
        else:
            raise _lib.RgfmError(msg)


SCHEDULES = ("constant", "linear_decay", "linear_ramp", "cosine_decay")


def strength_schedule(guidance_schedule, guidance_strength: float, num_steps: int):
    """Time-dependent guidance strength: gamma_k used by Euler step k at t_k = k / num_steps.
    None / 'constant' -> None (the reference's behaviour); 'linear_decay' gamma (1 - t); 'linear_ramp' gamma t;
    'cosine_decay' gamma (1 + cos(pi t)) / 2; a callable t -> gamma(t); or a sequence of num_steps values."""
    if guidance_schedule is None or guidance_schedule == "constant":
        return None
    ts = [k / num_steps for k in range(num_steps)]
    g = float(guidance_strength)
    if callable(guidance_schedule):
        return tuple(float(guidance_schedule(t)) for t in ts)
    if isinstance(guidance_schedule, str):
        if guidance_schedule == "linear_decay":
            return tuple(g * (1.0 - t) for t in ts)
        if guidance_schedule == "linear_ramp":
            return tuple(g * t for t in ts)
        if guidance_schedule == "cosine_decay":
            import math
            return tuple(g * 0.5 * (1.0 + math.cos(math.pi * t)) for t in ts)
        raise ValueError(f"Unknown guidance_schedule: {guidance_schedule}")
    vals = tuple(float(v) for v in guidance_schedule)
    if len(vals) != num_steps:
        raise ValueError(f"guidance_schedule must have num_steps = {num_steps} entries (got {len(vals)})")
    return vals


def clear_plan_cache():
    for p in _PLAN_CACHE.values():
        p.close()
    _PLAN_CACHE.clear()


def shard_bounds(num_samples: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous partition of the batch: the first `rem` ranks get one extra sample."""
    base, rem = divmod(num_samples, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_pairs(x_local: torch.Tensor, y_local: torch.Tensor, num_samples: int, group=None):
    """all_gather of finished pairs in rank-major order (the only data-path collective; NCCL on GPUs)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bounds = [shard_bounds(num_samples, world, r) for r in range(world)]
    nmax = max(hi - lo for lo, hi in bounds)

    def gather(t):
        pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        out = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(out, pad, group=group)
        return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, bounds)], dim=0)

    return gather(x_local), gather(y_local)


def _sample(pair, shape_x, shape_y, fm_x, fm_y, ratio_estimator, guidance_method, guidance_strength, num_samples,
            num_steps, device, mc_batch_size, precision, shard, group, chunk, verbose, noise=None, guidance_schedule=None):
    fm_x.eval()
    fm_y.eval()
    if ratio_estimator is not None:
        ratio_estimator.eval()
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.RgfmError("rgfm samplers run on CUDA devices only; there is no CPU fallback")
    precision = precision or default_precision()
    guided = guidance_method == "mc_feng" and ratio_estimator is not None
    # README.md:159-164 extension (the reference's argparse only offers none / mc_feng): no MC bank, no extra RNG draws
    grad_guided = guidance_method == "grad_log_ratio" and ratio_estimator is not None
    loss_type = getattr(ratio_estimator, "loss_type", "disc") if ratio_estimator is not None else "disc"
    if (guided or grad_guided) and loss_type not in _lib.LOSS:
        raise ValueError(f"Unknown loss_type: {loss_type}")

    # The reference consumes the device's global RNG with exactly these draws, in this order
    # (flow_utils.py:220-221,235,244).  Every rank of a sharded run draws the FULL batch and slices its rows, so
    # the gathered result equals the single-device one.
    if noise is None:
        x0 = torch.randn(num_samples, *shape_x, device=device)
        y0 = torch.randn(num_samples, *shape_y, device=device)
        mc_x0 = mc_y0 = None
        if guided:
            if verbose:
                print(f"  Generating {mc_batch_size} independent MC samples from flows...")
            mc_x0 = torch.randn(mc_batch_size, *shape_x, device=device)
            mc_y0 = torch.randn(mc_batch_size, *shape_y, device=device)
    else:
        x0, y0, mc_x0, mc_y0 = noise

    lo, hi = 0, num_samples
    if shard:
        import torch.distributed as dist
        lo, hi = shard_bounds(num_samples, dist.get_world_size(group), dist.get_rank(group))
    n_local = hi - lo
    if n_local > 0:
        use_ratio = guided or grad_guided
        plan = _get_plan(pair, fm_x, fm_y, ratio_estimator if use_ratio else None, guidance_method if use_ratio else "none",
                         guidance_strength, n_local, num_steps, mc_batch_size, loss_type, precision, device, chunk,
                         strength_schedule(guidance_schedule, guidance_strength, num_steps) if use_ratio else None)
        x1, y1, diag = plan.run(x0[lo:hi], y0[lo:hi], mc_x0, mc_y0, want_diag=True)
    else:
        x1 = torch.empty((0,) + shape_x, device=device)
        y1 = torch.empty((0,) + shape_y, device=device)
        diag = torch.zeros(_lib.DIAG_FLOATS, device=device)

    d = diag.tolist()  # one D2H after the whole loop (the reference syncs mid-loop)
    _check_saturation(d, precision)
    if verbose and guided:
        print(f"  Generated MC samples: x shape={torch.Size((mc_batch_size,) + shape_x)}, "
              f"y shape={torch.Size((mc_batch_size,) + shape_y)}")
        print(f"  MC ratios: min={d[9]:.4f}, max={d[10]:.4f}, mean={d[11]:.4f}")
        if pair == "pair28" and d[1] != 0.0:
            print(f"\n[MC Guidance Diagnostics at t={d[0]:.2f}]")
            print(f"  sigma_t={d[1]:.4f}")
            print(f"  ||v_x||={d[2]:.4f}, ||v_y||={d[3]:.4f}")
            print(f"  ||g_x||={d[4]:.4f}, ||g_y||={d[5]:.4f}")
            print(f"  weights: min={d[6]:.6f}, max={d[7]:.6f}")
            print(f"  Z_bar: {d[8]:.4f}")
    if shard:
        x1, y1 = gather_pairs(x1, y1, num_samples, group)
    return x1, y1


def sample_bimodal_guided(fm_x, fm_y, ratio_estimator=None, guidance_method='none', guidance_strength=0.0,
                          num_samples=16, num_steps=100, device='cuda', mc_batch_size=64, *,
                          precision: Optional[str] = None, shard: bool = False, group=None, chunk: int = 0,
                          verbose: bool = True, _noise=None, guidance_schedule=None):
    """(MNIST, transformed-MNIST) 1x28x28 pairs.  Returns (samples_x [N,1,28,28], samples_y [N,1,28,28]).
    guidance_schedule (extension, see strength_schedule): time-dependent gamma(t) for mc_feng / grad_log_ratio."""
    return _sample("pair28", (1, 28, 28), (1, 28, 28), fm_x, fm_y, ratio_estimator, guidance_method,
                   guidance_strength, num_samples, num_steps, device, mc_batch_size, precision, shard, group, chunk,
                   verbose, _noise, guidance_schedule)


def sample_bimodal_guided_mnist_svhn(fm_mnist, fm_svhn, ratio_estimator=None, guidance_method='none',
                                     guidance_strength=0.0, num_samples=16, num_steps=100, device='cuda',
                                     mc_batch_size=64, *, precision: Optional[str] = None, shard: bool = False,
                                     group=None, chunk: int = 0, verbose: bool = True, _noise=None, guidance_schedule=None):
    """(MNIST 1x32x32, SVHN 3x32x32) pairs.  Returns (samples_mnist [N,1,32,32], samples_svhn [N,3,32,32])."""
    return _sample("mnist_svhn", (1, 32, 32), (3, 32, 32), fm_mnist, fm_svhn, ratio_estimator, guidance_method,
                   guidance_strength, num_samples, num_steps, device, mc_batch_size, precision, shard, group, chunk,
                   verbose, _noise, guidance_schedule)


def sample_guidance_sweep(pair, fm_x, fm_y, ratio_estimator, guidance_strengths, num_samples=16, num_steps=100,
                          device='cuda', mc_batch_size=64, *, precision: Optional[str] = None, same_prior: bool = False,
                          chunk: int = 0, evaluators=None, transform_type: Optional[str] = None, coherence_out: Optional[dict] = None):
    """mc_feng pairs for SEVERAL guidance strengths from ONE plan and ONE MC bank.

    The reference's evaluators call the sampler once per gamma (src/evaluate.py:189-199,
    src/evaluate_mnist_svhn.py:155-165) and every call rebuilds the bank: N_mc samples through 100 steps of both flows
    plus the ratio pass, a quarter of a call at B=1024 / N_mc=256.  Here the MC noise is drawn once, the first gamma
    runs the full call, and every further gamma only rewrites the per-step scalar table and runs the main loop.
    pair: "pair28" (1x28x28, 1x28x28) or "mnist_svhn" (1x32x32, 3x32x32).  A fresh prior (x0, y0) is drawn per gamma,
    in call order like the reference would, unless same_prior=True (paired comparison across gammas).
    evaluators=(classifier_x, classifier_y) fuses the coherence evaluation (evaluate_mnist_svhn.py:28-57, evaluate.py:56-90)
    into the sampler's tail: the classifiers run on the finished pairs while they are still on the device and
    coherence_out[gamma] = {'coherence_acc', 'num_samples'} is filled from ONE device scalar per gamma (diag[15]);
    transform_type is the 28x28 pair's transform (undone on y before its classifier).
    Returns {gamma: (samples_x, samples_y)}."""
    shapes = {"pair28": ((1, 28, 28), (1, 28, 28)), "mnist_svhn": ((1, 32, 32), (3, 32, 32))}
    if pair not in shapes:
        raise ValueError(f"Unknown pair: {pair}")
    if ratio_estimator is None:
        raise ValueError("sample_guidance_sweep needs a ratio estimator (mc_feng guidance)")
    shape_x, shape_y = shapes[pair]
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.RgfmError("rgfm samplers run on CUDA devices only; there is no CPU fallback")
    for m in (fm_x, fm_y, ratio_estimator):
        m.eval()
    loss_type = getattr(ratio_estimator, "loss_type", "disc")
    if loss_type not in _lib.LOSS:
        raise ValueError(f"Unknown loss_type: {loss_type}")
    gammas = [float(g) for g in guidance_strengths]
    out: Dict[float, Tuple[torch.Tensor, torch.Tensor]] = {}
    if not gammas:
        return out
    plan = _lib.SamplerPlan(pair, fm_x.state_dict(), fm_y.state_dict(), ratio_estimator.state_dict(),
                            guidance_method="mc_feng", guidance_strength=gammas[0], num_samples=num_samples,
                            num_steps=num_steps, mc_batch_size=mc_batch_size, loss_type=loss_type,
                            precision=precision or default_precision(), device=device, chunk=chunk)
    try:
        x0 = y0 = None
        diags = {}
        if evaluators is not None:
            from . import evaluate as _ev  # classifier nets are cached by weight content there
            cx, cy = evaluators
            cx.eval()
            cy.eval()
            plan.attach_evaluators(_ev._net(cx, (num_samples,) + shape_x, device), _ev._net(cy, (num_samples,) + shape_y, device),
                                   transform_type)
        for i, g in enumerate(gammas):
            if x0 is None or not same_prior:
                x0 = torch.randn(num_samples, *shape_x, device=device)
                y0 = torch.randn(num_samples, *shape_y, device=device)
            if i == 0:
                mc_x0 = torch.randn(mc_batch_size, *shape_x, device=device)
                mc_y0 = torch.randn(mc_batch_size, *shape_y, device=device)
                r = plan.run(x0, y0, mc_x0, mc_y0, want_diag=True)
            else:
                plan.set_strength(g)
                r = plan.run(x0, y0, reuse_bank=True, want_diag=True)
            out[g] = (r[0], r[1])
            diags[g] = r[2]
        torch.cuda.current_stream(device).synchronize()
        if evaluators is not None and coherence_out is not None:
            for g in gammas:
                coherence_out[g] = {"coherence_acc": float(diags[g][15]) / float(max(num_samples, 1)), "num_samples": int(num_samples)}
    finally:
        plan.close()
    return out

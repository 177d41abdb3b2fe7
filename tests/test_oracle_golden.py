"""CPU: the oracle (oracle/rgfm_oracle.py) against the fixtures generated from the live reference
(oracle/make_golden.py asserted max-abs 0.0 at generation time).  Tolerance 2e-5 only absorbs a different host
CPU's oneDNN/MKL kernel selection; on the authoring container the match is exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden, sd_digest, T
from oracle import rgfm_oracle as O

TOL = 2e-5


def close(a, b, tol=TOL):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    d = (a - b).abs().max().item()
    assert d <= tol, f"max-abs {d} > {tol}"


@pytest.mark.parametrize("kind", ["unet_mnist28", "unet_mnist32", "unet_svhn", "flow28"])
def test_unet_forward_golden(kind, synth):
    g = load_golden(f"{kind}_forward")
    sd = synth(kind, g["seed"])
    assert sd_digest(sd) == str(g["digest"]), "synthetic weight generator drifted from the fixture"
    with torch.no_grad():
        v = O.unet_forward(sd, kind, T(g["x"]), T(g["t"]))
    close(v, g["v"])
    assert v.abs().max() > 0.1  # not the vacuous zero-initialised out_conv


@pytest.mark.parametrize("kind", ["ratio28", "ratio_mnist_svhn"])
def test_ratio_golden(kind, synth):
    g = load_golden(f"{kind}_forward")
    sd = synth(kind, g["seed"])
    assert sd_digest(sd) == str(g["digest"])
    x, y = T(g["x"]), T(g["y"])
    with torch.no_grad():
        close(O.ratio_scores(sd, kind, x, y), g["scores"])
        close(O.ratio_log_ratio(sd, kind, x, y, "disc"), g["log_ratio_disc"])
        close(O.ratio_log_ratio(sd, kind, x, y, "rulsif"), g["log_ratio_rulsif"])
        with pytest.raises(ValueError):
            O.ratio_log_ratio(sd, kind, x, y, "nope")
    gx, gy, _ = O.grad_log_ratio(sd, kind, x, y, "disc")
    close(gx, g["grad_x_disc"], 1e-4)
    close(gy, g["grad_y_disc"], 1e-4)


@pytest.mark.parametrize("loss_type", ["disc", "rulsif"])
@pytest.mark.parametrize("kind", ["ratio28", "ratio_mnist_svhn"])
def test_ratio_train_golden(kind, loss_type, synth):
    """Training side: loss, metrics and every parameter gradient of one RatioTrainer.train_step vs the
    fixture written from the UNMODIFIED reference module + loss classes (digest per parameter: norm, sum, 64 samples)."""
    g = load_golden(f"{kind}_train_{loss_type}")
    sd = synth(kind, g["seed"])
    assert sd_digest(sd) == str(g["digest"])
    masks = [T(g[k]) for k in sorted(k for k in g if k.startswith("mask"))]
    loss, metrics, grads = O.ratio_train_grads(sd, kind, T(g["x"]), T(g["y"]), T(g["is_real"]), loss_type,
                                               float(g["lambda_penalty"]), masks)
    close(loss, g["loss"], 1e-5)
    for name, val in zip(g["metric_names"], g["metric_values"]):
        assert abs(metrics[str(name)] - float(val)) <= 1e-5
    gkeys = [k for k in g if k.startswith("g:")]
    assert {k[2:] for k in gkeys} == set(grads)
    for k in gkeys:
        flat = grads[k[2:]].flatten()
        idx = torch.linspace(0, flat.numel() - 1, min(flat.numel(), 64)).long()
        mine = torch.cat([flat.norm().reshape(1), flat.sum().reshape(1), flat[idx]])
        ref = T(g[k])
        assert (mine - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()), k
    with pytest.raises(ValueError):
        O.ratio_loss(torch.zeros(2), torch.zeros(2), "nope")


@pytest.mark.parametrize("kind", ["clf_mnist28", "clf_mnist32", "clf_svhn"])
def test_classifier_golden(kind, synth):
    g = load_golden(f"{kind}_forward")
    sd = synth(kind, g["seed"])
    with torch.no_grad():
        close(O.classifier_logits(sd, kind, T(g["x"])), g["logits"], 1e-4)


def test_mc_feng_block_golden():
    g = load_golden("mc_feng_block")
    mcx, mcy, r = T(g["mc_x"]), T(g["mc_y"]), T(g["mc_ratios"])
    for i in range(5):
        gx, gy, w, Z = O.mc_feng_guidance(T(g[f"x{i}"]), T(g[f"y{i}"]), mcx, mcy, r, float(g[f"t{i}"]))
        close(gx, g[f"gx{i}"], 1e-4)
        close(gy, g[f"gy{i}"], 1e-4)
        close(w, g[f"w{i}"], 1e-5)
        # weights are a normalised distribution over the MC bank
        assert torch.allclose(w.sum(dim=1), torch.ones(w.shape[0]), atol=1e-5)


def test_cfm_sample_golden(synth):
    """CFMSchedule.sample (flow_utils.py:69-100): oracle against the live reference's output on the replayed prior draw."""
    g = load_golden("cfm_sample_mnist28")
    sd = synth("unet_mnist28", int(g["seed"]))
    close(O.flow_sample(sd, "unet_mnist28", T(g["x0"]), int(g["steps"])), g["x1"], 1e-5)


@pytest.mark.parametrize("name", ["sampler_pair28_small", "sampler_pair28_rulsif", "sampler_pair28_none",
                                  "sampler_mnist_svhn_small", "sampler_mnist_svhn_g2", "sampler_pair28_original"])
def test_sampler_golden(name, synth):
    g = load_golden(name)
    pair = str(g["pair"])
    kx, ky, kr, _, _ = O.PAIR_KINDS[pair]
    if name.endswith("_original"):  # src/sample.py --model original: both flows are FlowMatchingModel
        kx = ky = "flow28"
    sx, sy, sr = (int(s) for s in g["seeds"])
    sdx, sdy, sdr = synth(kx, sx), synth(ky, sy), synth(kr, sr)
    assert [sd_digest(sdx), sd_digest(sdy), sd_digest(sdr)] == [str(d) for d in g["digests"]]
    method = str(g["method"])
    mcx0 = T(g["mc_x0"]) if method == "mc_feng" else None
    mcy0 = T(g["mc_y0"]) if method == "mc_feng" else None
    x1, y1 = O.sample_pair(sdx, sdy, sdr if method != "none" else None, pair, method, float(g["gamma"]),
                           int(g["steps"]), T(g["x0"]), T(g["y0"]), mcx0, mcy0, str(g["loss_type"]))
    close(x1, g["x1"], 1e-4)
    close(y1, g["y1"], 1e-4)


def test_guidance_is_convex_blend_and_skipped_at_t0(synth):
    """gamma is a convex blend; step 0 (t=0 <= eps) is unguided; unknown methods degrade to unguided."""
    sdx, sdy, sdr = synth("unet_mnist28", 11), synth("unet_mnist28", 14), synth("ratio28", 21)
    g = torch.Generator().manual_seed(3)
    x0, y0 = torch.randn(2, 1, 28, 28, generator=g), torch.randn(2, 1, 28, 28, generator=g)
    m0, n0 = torch.randn(4, 1, 28, 28, generator=g), torch.randn(4, 1, 28, 28, generator=g)
    a = O.sample_pair(sdx, sdy, sdr, "pair28", "mc_feng", 0.0, 3, x0, y0, m0, n0)
    b = O.sample_pair(sdx, sdy, None, "pair28", "none", 0.0, 3, x0, y0)
    c = O.sample_pair(sdx, sdy, sdr, "pair28", "bogus", 0.7, 3, x0, y0, m0, n0)
    one = O.sample_pair(sdx, sdy, sdr, "pair28", "mc_feng", 0.9, 1, x0, y0, m0, n0)
    one_plain = O.sample_pair(sdx, sdy, None, "pair28", "none", 0.0, 1, x0, y0)
    assert torch.allclose(a[0], b[0], atol=1e-5) and torch.allclose(a[1], b[1], atol=1e-5)
    assert torch.equal(c[0], b[0]) and torch.equal(c[1], b[1])
    assert torch.equal(one[0], one_plain[0])


@pytest.mark.parametrize("kind", ["unet_mnist28", "unet_svhn"])
def test_unet_module_restatement_matches_functional_oracle(kind, synth):
    """oracle/unet_module.py (the module tree the flow-matching training tests reroute) against O.unet_forward, which
    oracle/make_golden.py pins to the live reference: forward and every parameter gradient of the FM loss."""
    from oracle.unet_module import UNet, fm_loss_and_grads
    sd = synth(kind, 12)
    m = UNet(kind).eval()
    m.load_state_dict(sd, strict=True)
    c, hw = (3, 32) if kind == "unet_svhn" else (1, 28)
    g = torch.Generator().manual_seed(4)
    x, t, u = torch.randn(2, c, hw, hw, generator=g), torch.rand(2, generator=g), torch.randn(2, c, hw, hw, generator=g)
    v = m(x, t)
    close(v.detach(), O.unet_forward(sd, kind, x, t))
    loss = torch.nn.functional.mse_loss(v, u)
    loss.backward()
    oloss, ograds = fm_loss_and_grads(sd, kind, x, t, u)
    close(loss.detach(), oloss)
    assert set(ograds) == {k for k, _ in m.named_parameters()}
    for k, p in m.named_parameters():
        close(p.grad, ograds[k], tol=TOL * max(1.0, ograds[k].abs().max().item()))

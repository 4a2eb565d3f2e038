"""DLinear latent predictors: oracle vs golden (CPU), oracle vs the unmodified reference
classes (build container), CUDA kernel vs oracle / golden through the C ABI (GPU)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, has_reference

VARIANTS = ["shared", "individual", "indc_indp"]


@pytest.fixture(scope="module")
def golden_extra():
    return dict(np.load(os.path.join(GOLDEN, "extra_golden.npz")))


def _oracle(variant):
    from oracle import predictor_oracle as PO
    from weatherforecastingtoolkit_b200.synthetic import make_dlinear_case
    cfg, params, lat = make_dlinear_case(variant)
    with torch.no_grad():
        return PO.dlinear_rollout(lat, *params, cfg.kernel_size, interleave_channels=variant == "indc_indp")


@pytest.mark.parametrize("variant", VARIANTS)
def test_dlinear_oracle_matches_golden(golden_extra, variant):
    pred, tgt, loss = _oracle(variant)
    np.testing.assert_allclose(pred.numpy(), golden_extra[f"dlinear_{variant}_pred"], rtol=0, atol=2e-6)
    np.testing.assert_array_equal(tgt.numpy(), golden_extra[f"dlinear_{variant}_tgt"])
    assert abs(loss.item() - float(golden_extra[f"dlinear_{variant}_loss"])) < 1e-6


@pytest.mark.skipif(not has_reference(), reason="/root/reference not present")
@pytest.mark.parametrize("variant", VARIANTS)
def test_dlinear_oracle_vs_reference_classes(variant):
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden_extra as G
    from weatherforecastingtoolkit_b200.synthetic import make_dlinear_case
    cfg, params, lat = make_dlinear_case(variant)
    m = G.ref_dlinear(variant, cfg, params)
    with torch.no_grad():
        want = G.ref_dlinear_step(m, lat, variant)
    got = _oracle(variant)
    torch.testing.assert_close(got[0], want[0], rtol=0, atol=2e-6)
    assert torch.equal(got[1], want[1])
    # drop-in state_dict surface: same parameter names and shapes as the reference module
    from weatherforecastingtoolkit_b200 import predictors as P
    mine = P.DLinearIndcIndp(cfg) if variant == "indc_indp" else P.DLinear(cfg)
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in m.state_dict().items()}


def test_dlinear_rejects_cpu_tensors():
    from weatherforecastingtoolkit_b200 import predictors as P
    from weatherforecastingtoolkit_b200.synthetic import make_dlinear_case
    cfg, _, lat = make_dlinear_case("shared")
    with pytest.raises(RuntimeError):
        P.DLinear(cfg).rollout(lat)
    with pytest.raises(ValueError):
        P.DLinear(P.dlinear_config(kernel_size=4))


def _load(mod, cfg, params):
    ws, bs, wt, bt = params
    with torch.no_grad():
        if cfg.individual:
            for i in range(cfg.enc_in):
                mod.Linear_Seasonal[i].weight.copy_(ws[i]); mod.Linear_Seasonal[i].bias.copy_(bs[i])
                mod.Linear_Trend[i].weight.copy_(wt[i]); mod.Linear_Trend[i].bias.copy_(bt[i])
        else:
            mod.Linear_Seasonal.weight.copy_(ws); mod.Linear_Seasonal.bias.copy_(bs)
            mod.Linear_Trend.weight.copy_(wt); mod.Linear_Trend.bias.copy_(bt)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
def test_dlinear_kernel_matches_oracle_and_golden(golden_extra, variant):
    from weatherforecastingtoolkit_b200 import predictors as P
    from weatherforecastingtoolkit_b200.synthetic import make_dlinear_case
    cfg, params, lat = make_dlinear_case(variant)
    mod = P.DLinearIndcIndp(cfg) if variant == "indc_indp" else P.DLinear(cfg)
    _load(mod, cfg, params)
    pred, tgt, loss = mod.rollout(lat.cuda())
    want = _oracle(variant)
    # fp32 both sides; only the summation order differs
    torch.testing.assert_close(pred.cpu(), want[0], rtol=0, atol=5e-6)
    assert torch.equal(tgt.cpu(), want[1])
    assert abs(loss.item() - want[2].item()) < 1e-6 * max(1.0, want[2].item())
    np.testing.assert_allclose(pred.cpu().numpy(), golden_extra[f"dlinear_{variant}_pred"], rtol=0, atol=5e-6)
    # plain DLinear.forward (no residual framing) on the same weights
    from oracle import predictor_oracle as PO
    b, t, c, h, w = lat.shape
    x = lat[:, :13].reshape(b, 13 * c, h * w) if variant == "indc_indp" else lat[:, :13].reshape(b, 13, c * h * w)
    torch.testing.assert_close(mod(x.cuda()).cpu(), PO.dlinear_forward(x, *params, cfg.kernel_size), rtol=0, atol=5e-6)


@pytest.mark.gpu
def test_dlinear_full_size_properties():
    """BASELINE-size latents (B=32, 25x4x48x48): with the reference's 1/L constant init and zero bias the
    prediction equals last + mean over the residual series (seasonal + trend = x), for every series."""
    from weatherforecastingtoolkit_b200 import predictors as P
    torch.manual_seed(3)
    lat = torch.randn(32, 25, 4, 48, 48, device="cuda")
    mod = P.DLinear(P.dlinear_config(individual=False, enc_in=9216))
    with torch.no_grad():
        mod.Linear_Seasonal.bias.zero_(); mod.Linear_Trend.bias.zero_()
    pred, tgt, loss = mod.rollout(lat)
    last = lat[:, 12:13]
    want = last + (lat[:, :13] - last).mean(dim=1, keepdim=True).expand(-1, 12, -1, -1, -1)
    torch.testing.assert_close(pred, want, rtol=0, atol=2e-5)
    assert torch.equal(tgt, (lat[:, 13:] - last) + last)
    ref_loss = ((pred - last) - (lat[:, 13:] - last)).double().pow(2).mean().item()
    assert abs(loss.item() - ref_loss) < 1e-5 * ref_loss

"""ConvModel latent compressor: oracle vs golden / the unmodified reference classes on CPU; the
single fused sm_100a kernel vs oracle and golden on the GPU (fp32 both sides)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, has_reference

sys.path.insert(0, GOLDEN)


@pytest.fixture(scope="module")
def golden_extra():
    return dict(np.load(os.path.join(GOLDEN, "extra_golden.npz")))


@pytest.fixture(scope="module")
def case():
    import make_golden_extra as G
    return G.convmodel_case()


def test_convmodel_oracle_matches_golden(golden_extra, case):
    from oracle import predictor_oracle as PO
    sd, x = case
    with torch.no_grad():
        z, rec = PO.convmodel_forward(x, sd)
    np.testing.assert_allclose(z.numpy(), golden_extra["convmodel_z"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(rec.numpy(), golden_extra["convmodel_recon"], rtol=0, atol=1e-5)


@pytest.mark.skipif(not has_reference(), reason="/root/reference not present")
def test_convmodel_oracle_vs_reference_classes(case):
    import make_golden_extra as G
    from oracle import predictor_oracle as PO
    from weatherforecastingtoolkit_b200.predictors import ConvModel as Mine
    sd, x = case
    ns = G.ref_script_classes(G.CONVMODEL_SCRIPT, ["ConvEncoder", "ConvDecoder", "ConvModel"])
    m = ns["ConvModel"](latent_dim=512).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        z, rec = m(x)
        z2, rec2 = PO.convmodel_forward(x, sd)
    assert torch.equal(z, z2) and torch.equal(rec, rec2)
    assert {k: tuple(v.shape) for k, v in Mine().state_dict().items()} == {k: tuple(v.shape) for k, v in m.state_dict().items()}


def test_convmodel_refuses_cpu(case):
    from weatherforecastingtoolkit_b200.predictors import ConvModel
    with pytest.raises(RuntimeError):
        ConvModel()(case[1])


@pytest.mark.gpu
def test_convmodel_kernel_matches_oracle_and_golden(golden_extra, case):
    from oracle import predictor_oracle as PO
    from weatherforecastingtoolkit_b200.predictors import ConvModel
    sd, x = case
    m = ConvModel(latent_dim=512)
    m.load_state_dict(sd, strict=True)
    z, rec, loss = m(x.cuda(), return_loss=True)
    with torch.no_grad():
        wz, wrec = PO.convmodel_forward(x, sd)
    # fp32 on both sides; only reduction orders differ
    torch.testing.assert_close(z.cpu(), wz, rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(rec.cpu(), wrec, rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(rec.cpu().numpy(), golden_extra["convmodel_recon"], rtol=1e-4, atol=2e-4)
    assert abs(loss.item() - float(golden_extra["convmodel_huber"])) < 1e-4
    z2, rec2 = m(x.cuda())
    assert torch.equal(z, z2) and torch.equal(rec, rec2)           # deterministic
    with pytest.raises(ValueError):
        m(torch.zeros(1, 1, 4, 32, 32, device="cuda"))

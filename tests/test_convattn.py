"""ConvAttnModel latent compressor: oracle vs golden / the unmodified reference class on CPU; the
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
    return G.convattn_case()


def test_convattn_oracle_matches_golden(golden_extra, case):
    from oracle import predictor_oracle as PO
    sd, x = case
    with torch.no_grad():
        z, rec = PO.convattn_forward(x, sd)
    np.testing.assert_allclose(z.numpy(), golden_extra["convattn_z"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(rec.numpy(), golden_extra["convattn_recon"], rtol=0, atol=5e-5)


@pytest.mark.skipif(not has_reference(), reason="/root/reference not present")
def test_convattn_oracle_vs_reference_class(case):
    import make_golden_extra as G
    from oracle import predictor_oracle as PO
    from weatherforecastingtoolkit_b200.predictors import ConvAttnModel as Mine
    sd, x = case
    ns = G.ref_script_classes(G.CONVATTN_SCRIPT, ["ConvAttnModel"])
    m = ns["ConvAttnModel"]().eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        z, rec = m(x)
        z2, rec2 = PO.convattn_forward(x, sd)
    torch.testing.assert_close(z, z2, rtol=0, atol=2e-5)
    torch.testing.assert_close(rec, rec2, rtol=0, atol=5e-5)
    mine = Mine()
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert len(mine._weight_pointers()) == 29 + 30 * 4 == len(list(mine.parameters())) + 1   # pool in_proj twice


def test_convattn_host_errors(case):
    from weatherforecastingtoolkit_b200.predictors import ConvAttnModel
    with pytest.raises(RuntimeError):
        ConvAttnModel()(case[1])
    with pytest.raises(ValueError):
        ConvAttnModel(transformer_embed_dim=256)
    with pytest.raises(ValueError):
        ConvAttnModel(num_tf_layers=9)


@pytest.mark.gpu
def test_convattn_kernel_matches_oracle_and_golden(golden_extra, case):
    from oracle import predictor_oracle as PO
    from weatherforecastingtoolkit_b200.predictors import ConvAttnModel
    sd, x = case
    m = ConvAttnModel()
    m.load_state_dict(sd, strict=True)
    z, rec, loss = m(x.cuda(), return_loss=True)
    with torch.no_grad():
        wz, wrec = PO.convattn_forward(x, sd)
    # fp32 on both sides; reduction orders and the online softmax differ
    torch.testing.assert_close(z.cpu(), wz, rtol=1e-4, atol=3e-4)
    torch.testing.assert_close(rec.cpu(), wrec, rtol=1e-4, atol=5e-4)
    np.testing.assert_allclose(z.cpu().numpy(), golden_extra["convattn_z"], rtol=1e-4, atol=3e-4)
    np.testing.assert_allclose(rec.cpu().numpy(), golden_extra["convattn_recon"], rtol=1e-4, atol=5e-4)
    assert abs(loss.item() - float(golden_extra["convattn_huber"])) < 1e-4
    # encode / decode separately == forward; deterministic
    z1 = m.encode(x.cuda())
    rec1 = m.decode(z1)
    assert torch.equal(z1, z) and torch.equal(rec1, rec)
    z2, rec2 = m(x.cuda())
    assert torch.equal(z, z2) and torch.equal(rec, rec2)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 4, 32, 32, device="cuda"))
    with pytest.raises(ValueError):
        m.decode(torch.zeros(1, 64, device="cuda"))


@pytest.mark.gpu
@pytest.mark.parametrize("layers,latent_dim,b", [(1, 64, 1), (2, 512, 150)])
def test_convattn_other_configs(layers, latent_dim, b):
    """Fewer layers / smaller latent / more frames than SMs, weights regenerated from the seed (oracle only, no golden)."""
    import make_golden_extra as G
    from oracle import predictor_oracle as PO
    from weatherforecastingtoolkit_b200.predictors import ConvAttnModel
    sd, x = G.convattn_case(seed=3, b=b, layers=layers, latent_dim=latent_dim)
    m = ConvAttnModel(num_tf_layers=layers, latent_dim=latent_dim)
    m.load_state_dict(sd, strict=True)
    z, rec = m(x.cuda())
    with torch.no_grad():
        wz, wrec = PO.convattn_forward(x, sd)
    torch.testing.assert_close(z.cpu(), wz, rtol=1e-4, atol=3e-4)
    torch.testing.assert_close(rec.cpu(), wrec, rtol=1e-4, atol=5e-4)

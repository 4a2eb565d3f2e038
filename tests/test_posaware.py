"""PosAwareAE_TF (BASELINE config 1 at the model's native 128x128): oracle vs
golden / reference on CPU; the sm_100a kernel program vs oracle and golden on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE, has_reference

sys.path.insert(0, GOLDEN)

REL_L2 = 1e-2   # fp16 operands / fp32 accumulate vs the fp32 reference


@pytest.fixture(scope="module")
def golden_extra():
    return dict(np.load(os.path.join(GOLDEN, "extra_golden.npz")))


@pytest.fixture(scope="module")
def case():
    import make_golden_extra as G
    return G.posaware_state_dict(), G.posaware_inputs()


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm()).item()


def test_posaware_oracle_matches_golden(golden_extra, case):
    from oracle import aux_oracle as AO
    sd, x = case
    with torch.no_grad():
        z = AO.posaware_encode(x, sd)
        y = AO.posaware_decode(z, sd)
    np.testing.assert_allclose(z.numpy(), golden_extra["posaware_latent"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(y.numpy(), golden_extra["posaware_recon"], rtol=0, atol=2e-5)


@pytest.mark.skipif(not has_reference(), reason="/root/reference not present")
def test_posaware_oracle_bitexact_vs_reference_and_state_dict_surface(case):
    sys.dont_write_bytecode = True
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    from pipeline.models.ae_64x8x8_lin import PosAwareAE_TF as Ref
    from oracle import aux_oracle as AO
    from weatherforecastingtoolkit_b200.models.ae_64x8x8_lin import PosAwareAE_TF as Mine
    sd, x = case
    m = Ref().eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        y, z = m(x[:1])
        z2 = AO.posaware_encode(x[:1], sd)
        assert torch.equal(z, z2) and torch.equal(y, AO.posaware_decode(z2, sd))
    mine = Mine()
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert mine.dec[-1].weight.shape == m.dec[-1].weight.shape   # `last_layer` of experiments/ae_v2_2/train.py:123-124


def test_posaware_refuses_cpu_training_and_other_sizes():
    from weatherforecastingtoolkit_b200.models.ae_64x8x8_lin import PosAwareAE_TF
    m = PosAwareAE_TF().eval()
    with pytest.raises(RuntimeError):
        m.encode(torch.rand(1, 1, 128, 128))
    with pytest.raises(RuntimeError):
        m.decode(torch.rand(1, 2048))
    with pytest.raises(ValueError):
        PosAwareAE_TF(in_channels=3)


@pytest.mark.gpu
def test_posaware_cuda_matches_oracle_and_golden(golden_extra, case):
    from oracle import aux_oracle as AO
    from weatherforecastingtoolkit_b200.models.ae_64x8x8_lin import PosAwareAE_TF
    sd, x = case
    m = PosAwareAE_TF().eval()
    m.load_state_dict(sd, strict=True)
    with pytest.raises(ValueError):
        m.encode(torch.rand(1, 1, 384, 384, device="cuda"))   # the reference raises at 384x384 too
    y, z = m(x.cuda())
    assert z.shape == (2, 2048) and y.shape == (2, 1, 128, 128)
    gz, gy = golden_extra["posaware_latent"], golden_extra["posaware_recon"]
    assert _rel(z, gz) < REL_L2, _rel(z, gz)
    assert _rel(y, gy) < REL_L2, _rel(y, gy)
    # decode alone from the reference latent, and the error measured on the centred signal (the sigmoid output has
    # a large mean that would flatter a plain relative L2)
    yd = m.decode(torch.from_numpy(gz).cuda())
    assert _rel(yd, gy) < REL_L2
    gyc = torch.from_numpy(gy).double()
    gyc = gyc - gyc.mean()
    assert ((yd.double().cpu() - torch.from_numpy(gy).double()).norm() / gyc.norm()).item() < 2e-2
    with torch.no_grad():
        assert _rel(z, AO.posaware_encode(x, sd)) < REL_L2
    # reproducible, and frames are independent (eval-mode BatchNorm): batch == per-frame
    y2, z2 = m(x.cuda())
    assert torch.equal(y, y2) and torch.equal(z, z2)
    z0 = m.encode(x[:1].cuda())
    assert _rel(z0, z[:1]) < 1e-5

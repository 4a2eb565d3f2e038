"""NLayerDiscriminator forward scoring (BASELINE config 5): oracle vs golden / reference on
CPU; CUDA path (stem kernel, tcgen05 4x4 convs with folded BatchNorm + LeakyReLU epilogue, logit head, hinge
reductions) vs oracle and golden on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE, has_reference

sys.path.insert(0, GOLDEN)

REL_L2 = 1e-2   # fp16 operands / fp32 accumulate vs the fp32 reference (same gate as the autoencoder forecasts)


@pytest.fixture(scope="module")
def golden_extra():
    return dict(np.load(os.path.join(GOLDEN, "extra_golden.npz")))


def _inputs(n, hw, seed):
    import make_golden_extra as G
    return G.disc_inputs(n, hw, seed)


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("tag,n,hw,seed", [("disc64", 2, 64, 41), ("disc384", 1, 384, 42)])
def test_disc_oracle_matches_golden(golden_extra, tag, n, hw, seed):
    from oracle import aux_oracle as AO
    from weatherforecastingtoolkit_b200.synthetic import make_discriminator_state_dict
    sd = make_discriminator_state_dict()
    real, fake = _inputs(n, hw, seed)
    with torch.no_grad():
        lr, lf = AO.discriminator_forward(real, sd), AO.discriminator_forward(fake, sd)
    np.testing.assert_allclose(lr.numpy(), golden_extra[f"{tag}_logits_real"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(lf.numpy(), golden_extra[f"{tag}_logits_fake"], rtol=0, atol=1e-4)
    assert abs(AO.hinge_d_loss(lr, lf).item() - float(golden_extra[f"{tag}_hinge"])) < 1e-5


@pytest.mark.skipif(not has_reference(), reason="/root/reference not present")
def test_disc_oracle_bitexact_vs_reference_and_state_dict_surface():
    sys.dont_write_bytecode = True
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    from pipeline.models.autoencoderkl.losses.model import NLayerDiscriminator as Ref
    from oracle import aux_oracle as AO
    from weatherforecastingtoolkit_b200.models.autoencoderkl.losses import NLayerDiscriminator as Mine
    from weatherforecastingtoolkit_b200.synthetic import make_discriminator_state_dict
    sd = make_discriminator_state_dict()
    m = Ref(input_nc=1).eval()
    m.load_state_dict(sd, strict=True)
    real, _ = _inputs(2, 64, 41)
    with torch.no_grad():
        assert torch.equal(m(real), AO.discriminator_forward(real, sd))
    mine = Mine(input_nc=1)
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == {k: tuple(v.shape) for k, v in m.state_dict().items()}
    mine.load_state_dict(sd, strict=True)


def test_disc_refuses_cpu_and_training_mode():
    from weatherforecastingtoolkit_b200.models.autoencoderkl.losses import NLayerDiscriminator, hinge_d_loss
    d = NLayerDiscriminator(input_nc=1)
    with pytest.raises(RuntimeError):
        d(torch.rand(1, 1, 64, 64))          # training mode
    with pytest.raises(RuntimeError):
        d.eval()(torch.rand(1, 1, 64, 64))   # CPU tensor
    with pytest.raises(RuntimeError):
        hinge_d_loss(torch.rand(4), torch.rand(4))
    with pytest.raises(ValueError):
        NLayerDiscriminator(input_nc=3)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,n,hw,seed", [("disc64", 2, 64, 41), ("disc384", 1, 384, 42)])
def test_disc_cuda_matches_oracle_and_golden(golden_extra, tag, n, hw, seed):
    from oracle import aux_oracle as AO
    from weatherforecastingtoolkit_b200.models.autoencoderkl.losses import NLayerDiscriminator, hinge_d_loss
    from weatherforecastingtoolkit_b200.synthetic import make_discriminator_state_dict
    sd = make_discriminator_state_dict()
    d = NLayerDiscriminator(input_nc=1).eval()
    d.load_state_dict(sd, strict=True)
    real, fake = _inputs(n, hw, seed)
    lr, lf = d(real.cuda()), d(fake.cuda())
    assert lr.shape == golden_extra[f"{tag}_logits_real"].shape
    with torch.no_grad():
        wr = AO.discriminator_forward(real, sd)
    assert _rel(lr, wr) < REL_L2, _rel(lr, wr)
    assert _rel(lr, golden_extra[f"{tag}_logits_real"]) < REL_L2
    assert _rel(lf, golden_extra[f"{tag}_logits_fake"]) < REL_L2
    # the padded 1x1 head: the border ring is the bare bias, exactly
    bias = sd["main.11.bias"].item()
    ring = torch.cat([lr[:, :, 0].flatten(), lr[:, :, -1].flatten(), lr[:, :, :, 0].flatten(), lr[:, :, :, -1].flatten()])
    assert torch.all(ring.cpu() == torch.tensor(bias))
    # hinge loss: exact reduction of the CUDA logits, and close to the reference value
    h = hinge_d_loss(lr, lf).item()
    assert abs(h - AO.hinge_d_loss(lr.cpu(), lf.cpu()).item()) < 1e-5
    assert abs(h - float(golden_extra[f"{tag}_hinge"])) < 2e-2 * max(1.0, float(golden_extra[f"{tag}_hinge"]))


@pytest.mark.gpu
def test_disc_batch_independence_and_rerun():
    """Frames are scored independently (eval-mode BatchNorm): a batch equals the per-frame results bit for bit,
    and a second call reproduces the first."""
    from weatherforecastingtoolkit_b200.models.autoencoderkl.losses import NLayerDiscriminator
    from weatherforecastingtoolkit_b200.synthetic import make_discriminator_state_dict
    d = NLayerDiscriminator(input_nc=1).eval()
    d.load_state_dict(make_discriminator_state_dict(), strict=True)
    real, fake = _inputs(3, 128, 43)
    x = torch.cat([real, fake]).cuda()
    full = d(x)
    assert torch.equal(full, d(x))
    for i in range(x.shape[0]):
        assert torch.equal(full[i:i + 1], d(x[i:i + 1]))

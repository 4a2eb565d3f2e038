"""AE_ViT_2048 (BASELINE config 4): oracle vs golden / reference on CPU; the sm_100a kernel
program (tcgen05 GEMMs + small attention / LayerNorm kernels) vs oracle and golden on the GPU."""
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
    return G.vit_state_dict(), G.vit_inputs()


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm()).item()


def test_vit_oracle_matches_golden(golden_extra, case):
    from oracle import aux_oracle as AO
    sd, x = case
    with torch.no_grad():
        tokens = AO.vit_encode_tokens(x, sd)
        y, lat = AO.vit_forward(x, sd)
    # torch's TransformerEncoderLayer takes fused inference paths in the reference; the restatement is the plain
    # op sequence, so it agrees to fp32 round-off rather than bit for bit
    np.testing.assert_allclose(tokens.numpy(), golden_extra["vit_tokens"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(lat.numpy(), golden_extra["vit_latent"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(y.numpy(), golden_extra["vit_recon"], rtol=0, atol=1e-4)


@pytest.mark.skipif(not has_reference(), reason="/root/reference not present")
def test_vit_oracle_vs_reference_and_state_dict_surface(case):
    import contextlib
    import io
    sys.dont_write_bytecode = True
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    with contextlib.redirect_stdout(io.StringIO()):
        from pipeline.models.ae_vit import AE_ViT_2048 as Ref
    from oracle import aux_oracle as AO
    from weatherforecastingtoolkit_b200.models.ae_vit import AE_ViT_2048 as Mine
    sd, x = case
    m = Ref().eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        y, lat = m(x[:1])
        y2, lat2 = AO.vit_forward(x[:1], sd)
    torch.testing.assert_close(lat2, lat, rtol=0, atol=2e-5)
    torch.testing.assert_close(y2, y, rtol=0, atol=1e-4)
    assert {k: tuple(v.shape) for k, v in Mine().state_dict().items()} == {k: tuple(v.shape) for k, v in m.state_dict().items()}


def test_vit_refuses_cpu_and_training():
    from weatherforecastingtoolkit_b200.models.ae_vit import AE_ViT_2048
    m = AE_ViT_2048()
    with pytest.raises(RuntimeError):
        m.eval()(torch.rand(1, 1, 128, 128))
    with pytest.raises(RuntimeError):
        m.eval().decode_tokens(torch.rand(1, 64, 512))


@pytest.mark.gpu
def test_vit_cuda_matches_oracle_and_golden(golden_extra, case):
    from oracle import aux_oracle as AO
    from weatherforecastingtoolkit_b200.models.ae_vit import AE_ViT_2048
    sd, x = case
    m = AE_ViT_2048().eval()
    m.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError):
        AE_ViT_2048()(x.cuda())                       # training mode (dropout) is refused
    y, lat = m(x.cuda())
    assert y.shape == (2, 1, 128, 128) and lat.shape == (2, 2048)
    assert _rel(lat, golden_extra["vit_latent"]) < REL_L2, _rel(lat, golden_extra["vit_latent"])
    assert _rel(y, golden_extra["vit_recon"]) < REL_L2, _rel(y, golden_extra["vit_recon"])
    tokens = m.encode_tokens(x.cuda())
    assert tokens.shape == (2, 64, 512)
    assert _rel(tokens, golden_extra["vit_tokens"]) < REL_L2
    # decoder stack alone, from the reference's own decoder input (from_latent(latent) + pos_embed)
    with torch.no_grad():
        zdec = AO.vit_from_latent(torch.from_numpy(golden_extra["vit_latent"]), sd)
        want = AO.vit_decode_tokens(zdec, sd)
    assert _rel(m.decode_tokens(zdec.cuda()), want) < REL_L2
    # reproducible; images are independent
    y2, lat2 = m(x.cuda())
    assert torch.equal(y, y2) and torch.equal(lat, lat2)
    y0, lat0 = m(x[:1].cuda())
    assert _rel(lat0, lat[:1]) < 1e-5 and _rel(y0, y[:1]) < 1e-5

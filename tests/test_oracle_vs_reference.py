"""Pin the oracle against the UNMODIFIED reference modules (build container only: skipped where
/root/reference does not exist, e.g. on the GPU box)."""
import sys
import types

import pytest
import torch

from conftest import REFERENCE, has_reference, metric_case_inputs

pytestmark = pytest.mark.skipif(not has_reference(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    sys.dont_write_bytecode = True
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    from oracle import metrics_oracle as MO
    tm, tmi = types.ModuleType("torchmetrics"), types.ModuleType("torchmetrics.image")

    class _M:
        def to(self, d):
            return self

    class SSIM(_M):
        def __init__(self, data_range=None):
            pass

        def __call__(self, p, g):
            return MO._tm_ssim(p, g)

    class PSNR(_M):
        def __call__(self, p, g):
            return MO._tm_psnr(p, g)

    tmi.StructuralSimilarityIndexMeasure, tmi.PeakSignalNoiseRatio = SSIM, PSNR
    tm.image = tmi
    sys.modules.setdefault("torchmetrics", tm)
    sys.modules.setdefault("torchmetrics.image", tmi)
    import pipeline.metrics as RM
    from pipeline.models.autoencoderkl.autoencoder_kl import AutoencoderKL
    return RM, AutoencoderKL


def test_akl_oracle_bitexact_vs_reference(ref, akl_weights):
    from oracle import akl_oracle as O
    from weatherforecastingtoolkit_b200.synthetic import make_vil_sequences
    _, AutoencoderKL = ref
    cfg, sd = akl_weights
    m = AutoencoderKL(**cfg).eval()
    m.load_state_dict(sd, strict=True)
    x = O.stage_vil(make_vil_sequences(1, 64, 64, 2, seed=7)).permute(0, 3, 1, 2)[:, :1].contiguous()
    with torch.no_grad():
        post = m.encode(x)
        assert torch.equal(post.parameters, O.akl_encode_moments(x, sd, cfg))
        z = post.mode().contiguous()
        assert torch.equal(m.decode(z), O.akl_decode(z, sd, cfg))
        mean, logvar, std, var = O.posterior_from_moments(post.parameters)
        assert torch.equal(post.std, std) and torch.equal(post.var, var) and torch.equal(post.logvar, logvar)


def test_drop_in_state_dict_keys_match_reference(ref):
    from weatherforecastingtoolkit_b200.synthetic import PATHB_AKL_CONFIG, akl_param_shapes
    _, AutoencoderKL = ref
    want = {k: tuple(v.shape) for k, v in AutoencoderKL(**PATHB_AKL_CONFIG).state_dict().items()}
    assert dict(akl_param_shapes(PATHB_AKL_CONFIG)) == want
    from weatherforecastingtoolkit_b200.models.autoencoderkl import AutoencoderKL as Mine
    got = {k: tuple(v.shape) for k, v in Mine(**PATHB_AKL_CONFIG).state_dict().items()}
    assert got == want


@pytest.mark.parametrize("name", ["rand_2x10x64", "unclamped_1x3x50x70"])
def test_metrics_oracle_bitexact_vs_reference(ref, name):
    from oracle import metrics_oracle as MO
    RM, _ = ref
    p, t = metric_case_inputs(name)
    want, got = RM.calc_metrics(p, t), MO.calc_metrics(p, t)
    assert list(want) == list(got)
    for k in want:
        assert want[k] == got[k] or (want[k] != want[k] and got[k] != got[k]), k
    for th in MO.THRESHOLDS:
        a = [float(v) for v in RM._hit_miss_fa_cn(p, t, th)]
        b = [float(v) for v in MO._hit_miss_fa_cn(p, t, th)]
        assert a == b

"""The CPU oracle reproduces the golden vectors that tests/golden/make_golden.py produced from
the unmodified reference (runs anywhere; no GPU, no /root/reference)."""
import numpy as np
import pytest
import torch

from conftest import METRIC_CASES, metric_case_inputs
from oracle import akl_oracle as O
from oracle import metrics_oracle as MO
from weatherforecastingtoolkit_b200.synthetic import make_predictor_params, make_vil_sequences


def _frames(n, hw, seed):
    u8 = make_vil_sequences(n, hw, hw, 1, seed=seed)
    return O.stage_vil(u8).permute(0, 3, 1, 2).contiguous()


def test_akl_64_bitexact(golden_akl, akl_weights):
    cfg, sd = akl_weights
    x = _frames(2, 64, 11)
    with torch.no_grad():
        m = O.akl_encode_moments(x, sd, cfg)
        d = O.akl_decode(m[:, :4].contiguous(), sd, cfg)
    np.testing.assert_allclose(m.numpy(), golden_akl["akl64_moments"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(d.numpy(), golden_akl["akl64_decoded"], rtol=0, atol=1e-5)


def test_posterior_matches_reference_semantics():
    m = torch.randn(2, 8, 4, 4) * 20
    mean, logvar, std, var = O.posterior_from_moments(m)
    assert logvar.max() <= 20 and logvar.min() >= -30
    assert torch.equal(mean, m[:, :4])
    assert torch.allclose(std * std, var, rtol=1e-5)


def test_rollout64_latent_algebra(golden_akl):
    """Predictor step (train.py:100-113) on the golden latents."""
    lat = torch.from_numpy(golden_akl["rollout64_latents"])
    w, b = make_predictor_params(seed=0)
    pred, tgt, loss = O.predictor_rollout(lat, w, b)
    np.testing.assert_allclose(pred.numpy(), golden_akl["rollout64_pred_latents"], atol=1e-6)
    assert abs(loss.item() - float(golden_akl["rollout64_val_loss"])) < 1e-6
    np.testing.assert_allclose(tgt.numpy(), lat[:, 13:].numpy(), atol=1e-6)


def test_rollout384_oracle_vs_reference_golden(akl_weights):
    """HEADLINE-size golden (tests/golden/make_golden_rollout384.py, unmodified reference modules, 384 x 384): the
    oracle reproduces one encoded frame, the predictor algebra and one decoded forecast frame of it (a bounded
    subset: the full 25 + 24 frame pass is ~2.5 minutes of CPU and is what the GPU test compares against)."""
    import json
    import os

    from conftest import GOLDEN
    gold = dict(np.load(os.path.join(GOLDEN, "rollout384_golden.npz")))
    with open(os.path.join(GOLDEN, "rollout384_metrics.json")) as f:
        meta = json.load(f)
    cfg, sd = akl_weights
    u8 = make_vil_sequences(1, 384, 384, 25, seed=meta["seed"])
    v = O.stage_vil(u8).permute(0, 3, 1, 2).unsqueeze(2)
    lat = torch.from_numpy(gold["latents"])
    w, b = make_predictor_params(seed=0)
    with torch.no_grad():
        z7 = O.akl_encode_moments(v[:, 7], sd, cfg)[:, :4]
        pred, tgt, loss = O.predictor_rollout(lat, w, b)
        d3 = O.akl_decode(torch.from_numpy(gold["pred_latents"][:, 3]), sd, cfg)
    np.testing.assert_allclose(z7.numpy(), gold["latents"][:, 7], atol=2e-5)
    np.testing.assert_allclose(pred.numpy(), gold["pred_latents"], atol=2e-6)
    np.testing.assert_allclose(tgt.numpy(), gold["tgt_latents"], atol=2e-6)
    assert loss.item() == pytest.approx(float(gold["val_loss"]), rel=1e-5)
    st, ph = meta["stride"], meta["phase"]
    np.testing.assert_allclose(d3[..., ph::st, ph::st].numpy(), gold["decoded_pred_sub"][:, 3], atol=5e-5)


@pytest.mark.parametrize("name", METRIC_CASES)
def test_metrics_oracle_vs_golden(golden_metrics, name):
    p, t = metric_case_inputs(name)
    got = MO.calc_metrics(p, t)
    want = golden_metrics[name]["metrics"]
    assert set(got) == set(want) and len(got) == 56
    for k in want:
        if np.isfinite(want[k]):
            assert got[k] == pytest.approx(want[k], rel=1e-6, abs=1e-7), k
        else:
            assert not np.isfinite(got[k])
    counts = MO.integer_counts(p, t)
    assert counts.tolist() == golden_metrics[name]["counts"]


def test_integer_counts_agree_with_float32_sums(golden_metrics):
    """Below 2**24 the reference's float32 counts are exact and must equal the integer counts (H1)."""
    for name in ("rand_2x10x64", "rand_2x12x384"):
        c = golden_metrics[name]["counts"][0][1]
        f = golden_metrics[name]["float32_counts_th1"]
        assert [float(v) for v in c] == f


def test_partials_consistency():
    p, t = metric_case_inputs("rand_2x10x64")
    pr = MO.partials(p, t)
    m = MO.calc_metrics(p, t)
    assert pr["abs_sum"][0] / pr["n_elems"][0] == pytest.approx(m["CRPS"], abs=1e-6)
    assert pr["ssim_sum"] / pr["n_frames"] == pytest.approx(m["SSIM"], abs=1e-6)
    assert pr["psnr_sum"] / pr["n_frames"] == pytest.approx(m["PSNR"], rel=1e-5)
    assert pr["counts"][0].sum(axis=1).tolist() == [pr["n_elems"][0]] * 6


def test_stage_vil_formula():
    u8 = torch.arange(256, dtype=torch.uint8).reshape(1, 16, 16, 1)
    x = O.stage_vil(u8)
    scale = np.float32(1 / 255)
    want = (np.arange(256, dtype=np.float32) * scale).reshape(1, 16, 16, 1)
    assert np.array_equal(x.numpy(), want)
    assert scale.view(np.uint32) == 0x3B808081


def test_metrics_oracle_vs_reference_extra_golden():
    """pool_type='max', arbitrary scale and ensemble forecasts (tests/golden/make_golden_metrics_extra.py: values from
    the unmodified reference module)."""
    import json
    import os
    import sys

    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    from make_golden_metrics_extra import THS, inputs
    with open(os.path.join(GOLDEN, "metrics_extra_golden.json")) as f:
        gold = json.load(f)
    p, t, ens, gt = inputs()
    for row in gold["pooled"]:
        pool, scale = row["pool_type"], row["scale"]
        assert MO.crps(p, t, pool, scale) == pytest.approx(row["crps"], rel=1e-6)
        for th in THS:
            assert MO.csi(p, t, th, pool, scale) == row[f"csi_{th:.6f}"]
            assert MO.hss(p, t, th, pool, scale) == row[f"hss_{th:.6f}"]
    for row in gold["crps_ensemble"]:
        assert MO.crps(ens, gt, row["pool_type"], row["scale"]) == pytest.approx(row["crps"], rel=1e-6)
    got = MO.calc_metrics(ens, gt)
    for k, v in gold["ensemble_calc_metrics"].items():
        assert got[k] == pytest.approx(v, rel=1e-6, abs=1e-7), k

"""The other Path-B experiments behind the same step: DLinear predictors inside
``PathBNowcast`` and the ConvModel / ConvAttnModel latent compressors inside ``LatentReconstruction``, against the CPU
oracle composition (fp32 AutoencoderKL oracle + predictor oracle)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
DEV = "cuda:0"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_wrappers_reject_bad_predictors(akl_weights):
    from weatherforecastingtoolkit_b200.rollout import PathBNowcast
    with pytest.raises(TypeError):
        PathBNowcast(akl_weights[0], predictor=torch.nn.Linear(52, 48))


@pytest.mark.gpu
def test_pathb_with_dlinear_predictor(akl_weights):
    from oracle import akl_oracle as O
    from oracle import predictor_oracle as PO
    from weatherforecastingtoolkit_b200.predictors import DLinear, dlinear_config
    from weatherforecastingtoolkit_b200.rollout import PathBNowcast
    from weatherforecastingtoolkit_b200.synthetic import make_vil_sequences
    cfg, sd = akl_weights
    dcfg = dlinear_config(seq_len=13, pred_len=12, individual=False, enc_in=4 * 8 * 8, kernel_size=3)
    torch.manual_seed(11)
    pred = DLinear(dcfg)
    net = PathBNowcast(cfg, posterior="mode", frames_per_call=16, predictor=pred)
    net.autoencoder.autoencoder.load_state_dict(sd, strict=True)
    net = net.to(DEV)
    u8 = make_vil_sequences(2, 64, 64, 25, seed=8)
    dp, dt, loss = net.validation_step(u8.to(DEV))
    with torch.no_grad():
        lat = O.wrapper_encode(O.stage_vil(u8).permute(0, 3, 1, 2).unsqueeze(2), sd, cfg)
        p = {k: v.detach().cpu() for k, v in pred.state_dict().items()}
        wp, wt, wl = PO.dlinear_rollout(lat, p["Linear_Seasonal.weight"], p["Linear_Seasonal.bias"],
                                        p["Linear_Trend.weight"], p["Linear_Trend.bias"], 3)
        odp, odt = O.wrapper_decode(wp, sd, cfg), O.wrapper_decode(wt, sd, cfg)
    assert rel_l2(dp, odp) < 1e-2 and rel_l2(dt, odt) < 1e-2
    assert loss.item() == pytest.approx(wl.item(), rel=2e-2)
    res = net.evaluate(u8.to(DEV))
    assert "CSI_0" in res and "val_loss" in res


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["convmodel", "convattn"])
def test_latent_reconstruction_step(akl_weights, kind):
    import make_golden_extra as G
    from oracle import akl_oracle as O
    from oracle import predictor_oracle as PO
    from weatherforecastingtoolkit_b200.predictors import ConvAttnModel, ConvModel
    from weatherforecastingtoolkit_b200.rollout import LatentReconstruction
    from weatherforecastingtoolkit_b200.synthetic import make_vil_sequences
    cfg, sd = akl_weights
    if kind == "convmodel":
        psd, _ = G.convmodel_case()
        predictor, b, t = ConvModel(latent_dim=512), 1, 2
    else:
        psd, _ = G.convattn_case()
        predictor, b, t = ConvAttnModel(), 2, 1
    predictor.load_state_dict(psd, strict=True)
    net = LatentReconstruction(cfg, predictor, posterior="mode", frames_per_call=4)
    net.autoencoder.autoencoder.load_state_dict(sd, strict=True)
    net = net.to(DEV)
    u8 = make_vil_sequences(b, 384, 384, t, seed=4)
    dp, inp, loss = net.validation_step(u8.to(DEV))
    assert dp.shape == (b, t, 1, 384, 384) and inp.shape == dp.shape
    with torch.no_grad():
        x = O.stage_vil(u8).permute(0, 3, 1, 2).unsqueeze(2).contiguous()
        assert torch.equal(inp.cpu(), x)
        lat = O.wrapper_encode(x, sd, cfg)
        if kind == "convmodel":
            _, rec = PO.convmodel_forward(lat, psd)
        else:
            _, r = PO.convattn_forward(lat[:, 0], psd)
            rec = r.unsqueeze(1)
        wl = torch.nn.HuberLoss()(rec, lat)
        odp = O.wrapper_decode(rec, sd, cfg)
    assert rel_l2(dp, odp) < 1e-2
    assert loss.item() == pytest.approx(wl.item(), rel=2e-2)
    res = net.evaluate(u8.to(DEV), extended=True)
    assert res["val_loss"] == pytest.approx(loss.item(), rel=1e-6) and 0.0 <= res["SSIM"] <= 1.0


@pytest.mark.gpu
def test_validation_epoch_end_to_end(akl_weights):
    """Loader -> rollout -> accumulator -> panels (scripts/validate_epoch.py) on small synthetic events: the epoch scores
    equal calc_metrics of the concatenated batches; the loader's float batches equal the reference formula."""
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "scripts"))
    import validate_epoch as V
    from weatherforecastingtoolkit_b200 import metrics as M
    from weatherforecastingtoolkit_b200.datastage import DeviceSEVIRLoader
    from weatherforecastingtoolkit_b200.rollout import PathBNowcast
    from weatherforecastingtoolkit_b200.synthetic import make_predictor_params, make_vil_sequences
    cfg, sd = akl_weights
    net = PathBNowcast(cfg, posterior="mode", frames_per_call=16)
    net.autoencoder.autoencoder.load_state_dict(sd, strict=True)
    w, b = make_predictor_params(seed=0)
    net.predictor.weight.data.copy_(w)
    net.predictor.bias.data.copy_(b)
    net = net.to(DEV)
    events = make_vil_sequences(4, 64, 64, 49, seed=9).numpy()
    scores, mosaics, loader = V.run_epoch(events, net, batch_size=5, panels=2)
    assert len(loader) == 2 and loader.h2d_bytes <= 2 * 3 * events[0].size       # 12 windows // 5; <= 3 events per batch
    assert len(mosaics) == 2 and mosaics[0].shape == (3 * 64, 12 * 64, 4)
    # the same epoch by hand: windows of the sequential sampler, one concatenated calc_metrics
    ld = DeviceSEVIRLoader(events, batch_size=5, layout="NHWT", split_mode="floor")
    dps, dts = [], []
    for batch in ld:
        x = batch["vil"]
        assert x.shape == (5, 64, 64, 25) and float(x.max()) <= 1.0
        dp, dt, _ = net.validation_step(x)
        dps.append(dp), dts.append(dt)
    want = M.calc_metrics(torch.cat(dps), torch.cat(dts), extended=True)
    for k, v in want.items():
        if np.isnan(v):          # a frame of the 64x64 synthetic data with a constant target has no PSNR
            assert np.isnan(scores[k]), k
        else:
            assert scores[k] == pytest.approx(v, rel=1e-6, abs=1e-7), k

"""Generate the golden fixtures in this directory from the UNMODIFIED reference
(/root/reference, importable only in the build container). Run:

    python tests/golden/make_golden.py

Inputs are regenerated from seeds at test time (``weatherforecastingtoolkit_b200.synthetic``);
only reference OUTPUTS are stored. ``pipeline.metrics`` needs ``torchmetrics`` which is not
installed: a stub module backed by ``oracle.metrics_oracle._tm_ssim/_tm_psnr`` (the restated
torchmetrics algorithm) is injected, so SSIM / PSNR goldens are NOT independent of the oracle
("parity unpinned" for those two keys); every other key comes from reference code alone.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from oracle import metrics_oracle as MO  # noqa: E402
from weatherforecastingtoolkit_b200.synthetic import (PATHB_AKL_CONFIG, make_akl_state_dict,  # noqa: E402
                                                      make_predictor_params, make_vil_sequences)


def _stub_torchmetrics():
    tm = types.ModuleType("torchmetrics")
    tmi = types.ModuleType("torchmetrics.image")

    class _M:
        def to(self, d):
            return self

    class SSIM(_M):
        def __init__(self, data_range=None):
            assert data_range == 1.0

        def __call__(self, p, g):
            return MO._tm_ssim(p, g)

    class PSNR(_M):
        def __call__(self, p, g):
            return MO._tm_psnr(p, g)

    tmi.StructuralSimilarityIndexMeasure = SSIM
    tmi.PeakSignalNoiseRatio = PSNR
    tm.image = tmi
    sys.modules["torchmetrics"] = tm
    sys.modules["torchmetrics.image"] = tmi


def main():
    _stub_torchmetrics()
    import pipeline.metrics as RM
    from pipeline.models.autoencoderkl.autoencoder_kl import AutoencoderKL

    torch.set_num_threads(os.cpu_count())
    cfg = PATHB_AKL_CONFIG
    sd = make_akl_state_dict(cfg, seed=0, affine_jitter=0.1)
    model = AutoencoderKL(**cfg).eval()
    model.load_state_dict(sd, strict=True)
    out = {}
    with torch.no_grad():
        # --- autoencoder, 64x64 (2 frames) and 384x384 (1 frame)
        for tag, n, hw, seed in (("akl64", 2, 64, 11), ("akl384", 1, 384, 12)):
            u8 = make_vil_sequences(n, hw, hw, 1, seed=seed)
            x = ((1 / 255) * (u8.float() + 0)).permute(0, 3, 1, 2).contiguous()
            post = model.encode(x)
            z = post.mode().contiguous()
            dec = model.decode(z)
            out[f"{tag}_moments"] = post.parameters.numpy()
            out[f"{tag}_decoded"] = dec.numpy()
        # --- Path-B validation_step algebra (train.py:100-116), B=1, 64x64, .mode()
        w, b = make_predictor_params(seed=0)
        u8 = make_vil_sequences(1, 64, 64, 25, seed=21)
        batch = (1 / 255) * (u8.float() + 0)
        v = batch.permute(0, 3, 1, 2).unsqueeze(2)
        lat = torch.cat([model.encode(v[:, i]).mode().unsqueeze(1) for i in range(25)], dim=1)
        bb, t, c, h, ww = lat.shape
        inp, tgt = lat[:, :13], lat[:, 13:]
        inp_t = inp[:, -1].unsqueeze(1)
        inp = inp - inp_t
        tgt = tgt - inp_t
        pred = torch.nn.functional.linear(inp.permute(0, 3, 4, 1, 2).reshape(bb, h, ww, 13 * c), w, b)
        pred = pred.permute(0, 3, 1, 2).reshape(bb, 12, c, h, ww)
        loss = torch.nn.functional.mse_loss(pred, tgt)
        pred = pred + inp_t
        tgt = tgt + inp_t
        dpred = torch.cat([model.decode(pred[:, i]).unsqueeze(1) for i in range(12)], dim=1)
        dtgt = torch.cat([model.decode(tgt[:, i]).unsqueeze(1) for i in range(12)], dim=1)
        out["rollout64_latents"] = lat.numpy()
        out["rollout64_pred_latents"] = pred.numpy()
        out["rollout64_decoded_pred"] = dpred.numpy()
        out["rollout64_decoded_tgt"] = dtgt.numpy()
        out["rollout64_val_loss"] = np.array(loss.item())
        roll_metrics = RM.calc_metrics(dpred, dtgt)
        roll_counts = MO.integer_counts(dpred, dtgt)
    np.savez_compressed(os.path.join(HERE, "akl_golden.npz"), **{k: v.astype(np.float32) for k, v in out.items()})

    # --- metrics goldens (inputs from seeds)
    met = {"rollout64": {"metrics": roll_metrics, "counts": roll_counts.tolist()}}
    cases = {
        "rand_2x10x64": ("rand", (2, 10, 1, 64, 64), 0),        # the reference's own __main__ smoke (metrics.py:135-141)
        "rand_2x12x384": ("rand", (2, 12, 1, 384, 384), 0),      # SURVEY 8c anchor
        "unclamped_1x3x50x70": ("randn", (1, 3, 1, 50, 70), 5),  # ragged size, values outside [0,1]
    }
    for name, (kind, shape, seed) in cases.items():
        torch.manual_seed(seed)
        if kind == "rand":
            p, t = torch.rand(*shape), torch.rand(*shape)
        else:
            p, t = torch.randn(*shape) * 0.6 + 0.4, torch.randn(*shape) * 0.6 + 0.4
        met[name] = {"metrics": RM.calc_metrics(p, t), "counts": MO.integer_counts(p, t).tolist(),
                     "float32_counts_th1": [float(v) for v in RM._hit_miss_fa_cn(p.clamp(0, 1), t.clamp(0, 1), 74 / 255)]}
    # smooth VIL-like pair
    u8 = make_vil_sequences(2, 384, 384, 13, seed=31)
    x = ((1 / 255) * u8.float()).permute(0, 3, 1, 2).unsqueeze(2)
    p, t = x[:, :12], x[:, 1:13]
    met["vil_2x12x384"] = {"metrics": RM.calc_metrics(p, t), "counts": MO.integer_counts(p, t).tolist()}
    with open(os.path.join(HERE, "metrics_golden.json"), "w") as f:
        json.dump(met, f, indent=1)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()

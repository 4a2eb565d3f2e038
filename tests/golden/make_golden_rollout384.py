"""Golden fixture for the HEADLINE configuration: one full Path-B ``validation_step`` (B = 1, 25 frames of 384 x 384,
encode 25 -> Linear(52 -> 48) -> decode 12 + 12, ``.mode()`` posterior) and ``calc_metrics`` on its output, computed
with the UNMODIFIED reference modules (/root/reference, build container only; ~3 minutes of CPU):

    python tests/golden/make_golden_rollout384.py

Follows ``experiments/v1_experiments/pretrained_ae_linear_sevir/train.py:100-116`` line by line (the script itself
cannot be imported: pytorch_lightning / omegaconf / wandb are absent). Inputs are regenerated from seeds at test time;
stored are the reference OUTPUTS: all latents (small) and, to keep the fixture at ~3 MB instead of 14 MB, every third
pixel (rows and columns 1, 4, 7, ...) of the decoded forecast / target frames together with their full-tensor norms.
SSIM / PSNR come from the restated torchmetrics algorithm (see make_golden.py)."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import _stub_torchmetrics  # noqa: E402
from oracle import metrics_oracle as MO  # noqa: E402
from weatherforecastingtoolkit_b200.synthetic import (PATHB_AKL_CONFIG, make_akl_state_dict,  # noqa: E402
                                                      make_predictor_params, make_vil_sequences)

SEED_DATA = 41
STRIDE, PHASE = 3, 1


def main():
    _stub_torchmetrics()
    import pipeline.metrics as RM
    from pipeline.models.autoencoderkl.autoencoder_kl import AutoencoderKL

    torch.set_num_threads(os.cpu_count())
    cfg = PATHB_AKL_CONFIG
    sd = make_akl_state_dict(cfg, seed=0, affine_jitter=0.1)
    model = AutoencoderKL(**cfg).eval()
    model.load_state_dict(sd, strict=True)
    w, b = make_predictor_params(seed=0)
    u8 = make_vil_sequences(1, 384, 384, 25, seed=SEED_DATA)
    with torch.no_grad():
        batch = (1 / 255) * (u8.float() + 0)                      # preprocess_data_dict, rescale '01'
        v = batch.permute(0, 3, 1, 2).unsqueeze(2)                # train.py:101
        lat = torch.cat([model.encode(v[:, i]).mode().unsqueeze(1) for i in range(25)], dim=1)   # train.py:32-43
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
        dpred = torch.cat([model.decode(pred[:, i]).unsqueeze(1) for i in range(12)], dim=1)      # train.py:45-56
        dtgt = torch.cat([model.decode(tgt[:, i]).unsqueeze(1) for i in range(12)], dim=1)
        metrics = RM.calc_metrics(dpred, dtgt)
        counts = MO.integer_counts(dpred, dtgt)
    sub = (slice(None), slice(None), slice(None), slice(PHASE, None, STRIDE), slice(PHASE, None, STRIDE))
    out = {
        "latents": lat.numpy(), "pred_latents": pred.numpy(), "tgt_latents": tgt.numpy(),
        "decoded_pred_sub": dpred[sub].numpy(), "decoded_tgt_sub": dtgt[sub].numpy(),
        "decoded_pred_norm": np.array(dpred.double().norm().item()), "decoded_tgt_norm": np.array(dtgt.double().norm().item()),
        "decoded_pred_frame_mean": dpred.double().mean(dim=(0, 2, 3, 4)).numpy(),
        "decoded_tgt_frame_mean": dtgt.double().mean(dim=(0, 2, 3, 4)).numpy(),
        "val_loss": np.array(loss.item()),
    }
    np.savez_compressed(os.path.join(HERE, "rollout384_golden.npz"),
                        **{k: (v.astype(np.float32) if v.dtype == np.float32 else v) for k, v in out.items()})
    with open(os.path.join(HERE, "rollout384_metrics.json"), "w") as f:
        json.dump({"seed": SEED_DATA, "stride": STRIDE, "phase": PHASE, "metrics": metrics, "counts": counts.tolist(),
                   "decoded_range": [float(dpred.min()), float(dpred.max())]}, f, indent=1)
    print("val_loss", loss.item(), "SSIM", metrics["SSIM"], "CSI_0", metrics["CSI_0"])


if __name__ == "__main__":
    main()

"""CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product package) for the DLinear latent
predictors of the reference experiments. Plain torch-CPU fp32 restatement on explicit weight tensors; every
function cites the reference lines it follows (paths relative to the reference repo root).

Pinned: ``tests/test_oracle_vs_reference.py`` executes the UNMODIFIED reference class definitions
(``moving_avg``, ``series_decomp``, ``DLinear``; extracted from the train scripts with ``ast`` because the
scripts themselves import pytorch_lightning / wandb, which are absent) and checks this restatement against
them; ``tests/golden/extra_golden.npz`` holds outputs generated the same way
(``tests/golden/make_golden_extra.py``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def moving_avg(x: torch.Tensor, kernel_size: int) -> torch.Tensor:
    """``moving_avg.forward`` (experiments/v1_experiments/pretrained_ae_dlinear_sevir/train.py:31-38):
    x [B, L, C]; both ends padded with (k-1)//2 copies of the end points, AvgPool1d(k, stride 1)."""
    front = x[:, 0:1, :].repeat(1, (kernel_size - 1) // 2, 1)
    end = x[:, -1:, :].repeat(1, (kernel_size - 1) // 2, 1)
    xp = torch.cat([front, x, end], dim=1)
    return F.avg_pool1d(xp.permute(0, 2, 1), kernel_size=kernel_size, stride=1, padding=0).permute(0, 2, 1)


def dlinear_forward(x, w_seasonal, b_seasonal, w_trend, b_trend, kernel_size: int) -> torch.Tensor:
    """``DLinear.forward`` (train.py:81-97). x [B, L, C]. Shared weights: w [P, L], b [P]; individual
    (train.py:69-77, 86-91): w [C, P, L], b [C, P] = the stacked ``nn.ModuleList``. Returns [B, P, C]."""
    trend = moving_avg(x, kernel_size)                      # series_decomp (train.py:48-51)
    seasonal = (x - trend).permute(0, 2, 1)                 # [B, C, L]
    trend = trend.permute(0, 2, 1)
    if w_seasonal.ndim == 3:
        so = torch.einsum("bcl,cpl->bcp", seasonal, w_seasonal) + b_seasonal.unsqueeze(0)
        to = torch.einsum("bcl,cpl->bcp", trend, w_trend) + b_trend.unsqueeze(0)
    else:
        so = F.linear(seasonal, w_seasonal, b_seasonal)
        to = F.linear(trend, w_trend, b_trend)
    return (so + to).permute(0, 2, 1)


def dlinear_rollout(v, w_seasonal, b_seasonal, w_trend, b_trend, kernel_size: int, in_frames: int = 13,
                    interleave_channels: bool = False):
    """Predictor part of ``Model.validation_step`` (pretrained_ae_dlinear_sevir/train.py:179-192; with
    ``interleave_channels`` the reshape of pretrained_ae_dlinear_indc_indp/train.py:185-186).
    v [B, T, C, h, w] latents -> (pred, tgt, val_loss), last input frame added back."""
    b, t, c, h, w = v.shape
    inp, tgt = v[:, :in_frames], v[:, in_frames:]
    inp_t = inp[:, -1].unsqueeze(1)
    inp = inp - inp_t
    tgt = tgt - inp_t
    if interleave_channels:
        x = inp.reshape(b, in_frames * c, h * w)
    else:
        x = inp.reshape(b, in_frames, c * h * w)
    pred = dlinear_forward(x, w_seasonal, b_seasonal, w_trend, b_trend, kernel_size).reshape(b, t - in_frames, c, h, w)
    loss = F.mse_loss(pred, tgt)
    return pred + inp_t, tgt + inp_t, loss


def convmodel_forward(x, sd):
    """``ConvModel.forward`` (experiments/v1_experiments/pretrained_ae_convae_sevir/train.py:133-143) with
    ``ConvEncoder`` (:58-89) and ``ConvDecoder`` (:92-117) restated on a state_dict. x [B, T, 4, 48, 48] -> (z, recon)."""
    b, t, c, h, w = x.shape
    y = x.reshape(b * t, c, h, w)

    def ln_act(v, p):
        return F.leaky_relu(F.layer_norm(v, tuple(v.shape[1:]), sd[p + ".weight"], sd[p + ".bias"], 1e-5))
    y = ln_act(F.conv2d(y, sd["encoder.conv0.0.weight"], sd["encoder.conv0.0.bias"], padding=1), "encoder.conv0.1")
    for name in ("down1", "down2", "down3"):
        y = ln_act(F.conv2d(y, sd[f"encoder.{name}.0.weight"], sd[f"encoder.{name}.0.bias"], stride=2, padding=1),
                   f"encoder.{name}.1")
    z = F.linear(y.reshape(b * t, -1), sd["to_latent.weight"], sd["to_latent.bias"])
    y = F.linear(z, sd["to_reconstruction.weight"], sd["to_reconstruction.bias"]).reshape(b * t, 8, 6, 6)
    for name in ("up1", "up2", "up3"):
        y = ln_act(F.conv_transpose2d(y, sd[f"decoder.{name}.0.weight"], sd[f"decoder.{name}.0.bias"], stride=2, padding=1),
                   f"decoder.{name}.1")
    y = F.conv2d(y, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)
    return z, y.reshape(b, t, c, h, w)

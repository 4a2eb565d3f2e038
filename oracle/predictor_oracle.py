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


def _mha(q_in, kv_in, sd, p, nhead):
    """nn.MultiheadAttention(batch_first=True) forward without masks: packed in_proj, heads of E/nhead, q scaled by
    1/sqrt(dh), softmax, out_proj."""
    e = q_in.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(q_in, w[:e], b[:e])
    k = F.linear(kv_in, w[e:2 * e], b[e:2 * e])
    v = F.linear(kv_in, w[2 * e:], b[2 * e:])
    bsz, tq, tk, dh = q.shape[0], q.shape[1], k.shape[1], e // nhead
    q = q.reshape(bsz, tq, nhead, dh).transpose(1, 2) * (dh ** -0.5)
    k = k.reshape(bsz, tk, nhead, dh).transpose(1, 2)
    v = v.reshape(bsz, tk, nhead, dh).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v
    att = att.transpose(1, 2).reshape(bsz, tq, e)
    return F.linear(att, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def convattn_forward(x, sd, nhead: int = 8):
    """``ConvAttnModel.forward`` (experiments/v1_experiments/pretrained_ae_convattn_ae_sevir/train.py:131-163) restated on
    a state_dict: encoder_cnn (:69-76), pre-norm nn.TransformerEncoderLayer stack (:79-86), attention pooling (:89-91,
    137-138), encoder_head (:93-96), decoder_head + learned queries (:98-101, 147-150), pre-norm
    nn.TransformerDecoderLayer stack (:103-110), decoder_cnn (:112-117). x [B, 4, 48, 48] -> (z [B, latent], recon)."""
    def ln(v, p):
        return F.layer_norm(v, (v.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)

    def ff(v, p):
        return F.linear(F.gelu(F.linear(v, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                        sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    b = x.shape[0]
    y = F.conv2d(x, sd["encoder_cnn.0.weight"], sd["encoder_cnn.0.bias"], stride=2, padding=1)
    y = F.gelu(F.group_norm(y, 8, sd["encoder_cnn.1.weight"], sd["encoder_cnn.1.bias"], 1e-5))
    y = F.conv2d(y, sd["encoder_cnn.3.weight"], sd["encoder_cnn.3.bias"], stride=2, padding=1)
    y = F.gelu(F.group_norm(y, 8, sd["encoder_cnn.4.weight"], sd["encoder_cnn.4.bias"], 1e-5))
    t = y.flatten(2).transpose(1, 2) + sd["encoder_pos_embedding"]
    layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder_tf.layers."))
    for l in range(layers):
        p = f"encoder_tf.layers.{l}"
        h = ln(t, p + ".norm1")
        t = t + _mha(h, h, sd, p + ".self_attn", nhead)
        t = t + ff(ln(t, p + ".norm2"), p)
    pooled = _mha(sd["pooling_query"].expand(b, -1, -1), t, sd, "attention_pool", nhead)
    z = F.linear(ln(pooled, "encoder_head.0"), sd["encoder_head.1.weight"], sd["encoder_head.1.bias"]).squeeze(1)
    ctx = F.linear(z, sd["decoder_head.weight"], sd["decoder_head.bias"]).unsqueeze(1)
    t = (sd["decoder_queries"] + sd["decoder_pos_embedding"]).expand(b, -1, -1)
    for l in range(layers):
        p = f"decoder_tf.layers.{l}"
        h = ln(t, p + ".norm1")
        t = t + _mha(h, h, sd, p + ".self_attn", nhead)
        t = t + _mha(ln(t, p + ".norm2"), ctx, sd, p + ".multihead_attn", nhead)
        t = t + ff(ln(t, p + ".norm3"), p)
    e = t.shape[-1]
    y = t.transpose(1, 2).reshape(b, e, 12, 12)
    y = F.conv_transpose2d(y, sd["decoder_cnn.0.weight"], sd["decoder_cnn.0.bias"], stride=2, padding=1)
    y = F.gelu(F.group_norm(y, 8, sd["decoder_cnn.1.weight"], sd["decoder_cnn.1.bias"], 1e-5))
    y = F.conv_transpose2d(y, sd["decoder_cnn.3.weight"], sd["decoder_cnn.3.bias"], stride=2, padding=1)
    return z, y

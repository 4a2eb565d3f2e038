"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the reference
Path-B autoencoder ``AutoencoderKL`` forward, written against a plain state_dict.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import this file. The product path (``weatherforecastingtoolkit_b200``)
never does.

Pinning: validated in the build container against the UNMODIFIED reference module imported
from ``/root/reference`` (``tests/test_oracle_vs_reference.py``; bit-for-bit on CPU because the
same ATen ops run in the same order) and against the committed golden fixtures under
``tests/golden/`` that ``tests/golden/make_golden.py`` produced from the reference itself.

Each function cites the reference lines it follows (paths relative to ``/root/reference``).

``emulate_bf16=True`` additionally rounds every tensor-core operand (conv / linear inputs and
weights, attention probabilities) to bf16 while accumulating in fp32. It predicts the numerics
of the sm_100a kernels so that the 1e-2 relative-L2 budget can be checked without a GPU; it is
not part of the parity definition.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


EMU_DTYPE = torch.bfloat16  # tensor-core operand type emulated by ``emulate`` mode


def _q(x: torch.Tensor, emu: bool) -> torch.Tensor:
    return x.to(EMU_DTYPE).to(torch.float32) if emu else x


def _conv(x, sd: SD, name: str, stride=1, padding=1, emu=False):
    return F.conv2d(_q(x, emu), _q(sd[f"{name}.weight"], emu), sd[f"{name}.bias"],
                    stride=stride, padding=padding)


def _linear(x, sd: SD, name: str, emu=False):
    return F.linear(_q(x, emu), _q(sd[f"{name}.weight"], emu), sd[f"{name}.bias"])


def _gn(x, sd: SD, name: str, groups: int, eps: float = 1e-6):
    return F.group_norm(x, groups, sd[f"{name}.weight"], sd[f"{name}.bias"], eps)


def resnet_block(x, sd: SD, p: str, groups: int, emu=False, stream_bf16=False):
    """``ResnetBlock2D.forward`` with temb=None, dropout p=0, output_scale_factor=1
    (``pipeline/models/autoencoderkl/resnet.py:454-495``)."""
    h = F.silu(_gn(x, sd, f"{p}.norm1", groups))
    h = _conv(h, sd, f"{p}.conv1", emu=emu)
    h = F.silu(_gn(_q(h, emu), sd, f"{p}.norm2", groups))
    h = _conv(h, sd, f"{p}.conv2", emu=emu)
    if f"{p}.conv_shortcut.weight" in sd:
        x = _conv(x, sd, f"{p}.conv_shortcut", padding=0, emu=emu)
    out = (x + h) / 1.0
    return _q(out, emu and stream_bf16)


def attention_block(x, sd: SD, p: str, groups: int, emu=False, stream_bf16=False):
    """``AttentionBlock.forward`` with one head, rescale_output_factor=1, non-xformers path
    (``pipeline/models/autoencoderkl/attention.py:136-189``)."""
    b, c, hh, ww = x.shape
    h = _gn(x, sd, f"{p}.group_norm", groups)
    h = h.view(b, c, hh * ww).transpose(1, 2)
    q = _linear(h, sd, f"{p}.query", emu)
    k = _linear(h, sd, f"{p}.key", emu)
    v = _linear(h, sd, f"{p}.value", emu)
    scale = 1 / math.sqrt(c / 1)
    scores = torch.baddbmm(
        torch.empty(b, q.shape[1], k.shape[1], dtype=q.dtype),
        _q(q, emu), _q(k, emu).transpose(-1, -2), beta=0, alpha=scale)
    probs = torch.softmax(scores.float(), dim=-1).type(scores.dtype)
    h = torch.bmm(_q(probs, emu), _q(v, emu))
    h = _linear(h, sd, f"{p}.proj_attn", emu)
    h = h.transpose(-1, -2).reshape(b, c, hh, ww)
    return _q((h + x) / 1.0, emu and stream_bf16)


def mid_block(x, sd: SD, p: str, groups: int, emu=False, stream_bf16=False):
    """``UNetMidBlock2D.forward`` (``unet_2d_blocks.py:162-169``)."""
    x = resnet_block(x, sd, f"{p}.resnets.0", groups, emu, stream_bf16)
    x = attention_block(x, sd, f"{p}.attentions.0", groups, emu, stream_bf16)
    x = resnet_block(x, sd, f"{p}.resnets.1", groups, emu, stream_bf16)
    return x


def encoder_forward(x, sd: SD, cfg: dict, emu=False, stream_bf16=False):
    """``Encoder.forward`` (``vae.py:70-86``); ``DownEncoderBlock2D.forward``
    (``unet_2d_blocks.py:229-238``); ``Downsample2D.forward`` pad (0,1,0,1) + conv s2 p0
    (``resnet.py:181-190``)."""
    boc = list(cfg["block_out_channels"])
    lpb = int(cfg.get("layers_per_block", 1))
    g = int(cfg.get("norm_num_groups", 32))
    h = _conv(x, sd, "encoder.conv_in", emu=False)
    h = _q(h, emu and stream_bf16)
    for i in range(len(boc)):
        for j in range(lpb):
            h = resnet_block(h, sd, f"encoder.down_blocks.{i}.resnets.{j}", g, emu, stream_bf16)
        if i != len(boc) - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)
            h = _conv(h, sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", stride=2, padding=0, emu=emu)
            h = _q(h, emu and stream_bf16)
    h = mid_block(h, sd, "encoder.mid_block", g, emu, stream_bf16)
    h = F.silu(_gn(h, sd, "encoder.conv_norm_out", g))
    h = _conv(h, sd, "encoder.conv_out", emu=emu)
    return h


def decoder_forward(z, sd: SD, cfg: dict, emu=False, stream_bf16=False):
    """``Decoder.forward`` (``vae.py:150-166``); ``UpDecoderBlock2D.forward``
    (``unet_2d_blocks.py:270-279``); ``Upsample2D.forward`` nearest x2 + conv
    (``resnet.py:108-143``)."""
    boc = list(cfg["block_out_channels"])
    lpb = int(cfg.get("layers_per_block", 1))
    g = int(cfg.get("norm_num_groups", 32))
    h = _conv(z, sd, "decoder.conv_in", emu=False)
    h = _q(h, emu and stream_bf16)
    h = mid_block(h, sd, "decoder.mid_block", g, emu, stream_bf16)
    for i in range(len(boc)):
        for j in range(lpb + 1):
            h = resnet_block(h, sd, f"decoder.up_blocks.{i}.resnets.{j}", g, emu, stream_bf16)
        if i != len(boc) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", emu=emu)
            h = _q(h, emu and stream_bf16)
    h = F.silu(_gn(h, sd, "decoder.conv_norm_out", g))
    h = _conv(h, sd, "decoder.conv_out", emu=emu)
    return h


def akl_encode_moments(x, sd: SD, cfg: dict, emu=False, stream_bf16=False):
    """``AutoencoderKL.encode`` up to the posterior parameters
    (``autoencoder_kl.py:80-84``): encoder then ``quant_conv`` 1x1."""
    h = encoder_forward(x, sd, cfg, emu, stream_bf16)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def posterior_from_moments(moments):
    """``DiagonalGaussianDistribution.__init__`` (``distributions.py:27-35``):
    returns (mean, logvar clamped to [-30, 20], std, var)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean, logvar, torch.exp(0.5 * logvar), torch.exp(logvar)


def akl_decode(z, sd: SD, cfg: dict, emu=False, stream_bf16=False):
    """``AutoencoderKL._decode`` (``autoencoder_kl.py:86-89``)."""
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    return decoder_forward(z, sd, cfg, emu, stream_bf16)


# ----------------------------------------------------------------------------------------------
# Path-B wrapper + validation_step tensor algebra


def wrapper_encode(x, sd: SD, cfg: dict, noise: Optional[torch.Tensor] = None, **kw):
    """``Autoencoder.encode`` (``experiments/v1_experiments/pretrained_ae_linear_sevir/train.py:32-41``):
    per-frame loop; ``.mode()`` unless ``noise`` [B,T,LC,h,w] is injected (then mean+std*noise,
    the deterministic form of ``.sample()``, SURVEY hazard H3)."""
    b, t = x.shape[:2]
    out = []
    for i in range(t):
        mean, _, std, _ = posterior_from_moments(akl_encode_moments(x[:, i], sd, cfg, **kw))
        z = mean if noise is None else mean + std * noise[:, i]
        out.append(z.unsqueeze(1))
    return torch.cat(out, dim=1)


def wrapper_decode(z, sd: SD, cfg: dict, **kw):
    """``Autoencoder.decode`` (``train.py:45-54``)."""
    return torch.cat([akl_decode(z[:, i], sd, cfg, **kw).unsqueeze(1) for i in range(z.shape[1])], dim=1)


def predictor_rollout(v, weight, bias, in_frames: int = 13):
    """Latent predictor step of ``Model.validation_step`` (``train.py:100-113``):
    residual w.r.t. the last input frame, one ``Linear(13*C -> 12*C)`` per latent pixel.
    Returns (pred, tgt, val_loss) with the last frame added back."""
    b, t, c, h, w = v.shape
    inp, tgt = v[:, :in_frames], v[:, in_frames:]
    inp_t = inp[:, -1].unsqueeze(1)
    inp = inp - inp_t
    tgt = tgt - inp_t
    x = inp.permute(0, 3, 4, 1, 2).reshape(b, h, w, in_frames * c)
    pred = F.linear(x, weight, bias).permute(0, 3, 1, 2).reshape(b, t - in_frames, c, h, w)
    loss = F.mse_loss(pred, tgt)
    return pred + inp_t, tgt + inp_t, loss


def predictor_autoregressive(inp, weight, bias, blocks: int):
    """The same ``Linear(13*C -> 12*C)`` predictor stepped autoregressively (north_star: "a linear latent predictor
    stepped autoregressively"; the reference experiment itself applies it once, ``train.py:100-113``): each block of 12
    predicted frames is appended and the next block is predicted from the LAST 13 frames of the sequence so far, with
    the residual framing of ``train.py:104-112`` re-anchored on the newest frame. inp [B, 13, C, h, w] ->
    [B, 12 * blocks, C, h, w]."""
    b, t_in, c, h, w = inp.shape
    t_out = weight.shape[0] // c
    seq = inp
    outs = []
    for _ in range(blocks):
        win = seq[:, -t_in:]
        last = win[:, -1].unsqueeze(1)
        x = (win - last).permute(0, 3, 4, 1, 2).reshape(b, h, w, t_in * c)
        pred = F.linear(x, weight, bias).permute(0, 3, 1, 2).reshape(b, t_out, c, h, w) + last
        outs.append(pred)
        seq = torch.cat([seq, pred], dim=1)
    return torch.cat(outs, dim=1)


def validation_step(batch_nhwt, sd: SD, cfg: dict, weight, bias, in_frames: int = 13, **kw):
    """``Model.validation_step`` (``train.py:100-120``) up to the tensors handed to
    ``log_metrics``: returns (decoded_pred, decoded_tgt, val_loss)."""
    v = batch_nhwt.permute(0, 3, 1, 2).unsqueeze(2)
    lat = wrapper_encode(v, sd, cfg, **kw)
    pred, tgt, loss = predictor_rollout(lat, weight, bias, in_frames)
    return wrapper_decode(pred, sd, cfg, **kw), wrapper_decode(tgt, sd, cfg, **kw), loss


def stage_vil(u8_nhwt: torch.Tensor) -> torch.Tensor:
    """uint8 VIL -> float32 in [0,1], NHWT kept (``pipeline/datasets/sevir/sevir.py:54-63,
    587-592,656-665``): ``fl32(1/255) * (float(x) + 0)``."""
    return (1 / 255) * (u8_nhwt.float() + 0)

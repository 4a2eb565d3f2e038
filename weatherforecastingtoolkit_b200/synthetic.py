"""Deterministic synthetic weights and VIL sequences for tests and the benchmark.

There is no network on the GPU box and the reference's pretrained checkpoint
(``/home/vatsal/NWM/PreDiff/scripts/vae/sevirlr/autoencoder_ckpt.pth``,
reference ``experiments/v1_experiments/pretrained_ae_linear_sevir/train.py:27``) is not
shipped, so weights are random-init with PyTorch's default distributions
(uniform(+-1/sqrt(fan_in)) for conv/linear weights and biases, ones/zeros for GroupNorm),
generated per-parameter from a name-keyed seed so that the SAME state_dict can be loaded
into the unmodified reference ``AutoencoderKL`` (to make golden fixtures) and into the
B200 drop-in, independent of module construction order.

Key names and shapes are the reference's state_dict surface
(``pipeline/models/autoencoderkl/autoencoder_kl.py:37-77``, ``vae.py:9-166``,
``unet_2d_blocks.py:89-279``, ``resnet.py:367-453``, ``attention.py:48-91``).
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

PATHB_AKL_CONFIG = dict(
    in_channels=1,
    out_channels=1,
    down_block_types=["DownEncoderBlock2D"] * 4,
    up_block_types=["UpDecoderBlock2D"] * 4,
    block_out_channels=[128, 256, 512, 512],
    layers_per_block=2,
    latent_channels=4,
    norm_num_groups=32,
)
"""``experiments/v1_experiments/pretrained_ae_linear_sevir/config.yaml:5-13``."""

INPUT_FRAMES = 13   # config.yaml:36
PRED_FRAMES = 12    # config.yaml:37


def _resnet_shapes(prefix: str, cin: int, cout: int, out: "OrderedDict[str, Tuple[int, ...]]"):
    out[f"{prefix}.norm1.weight"] = (cin,)
    out[f"{prefix}.norm1.bias"] = (cin,)
    out[f"{prefix}.conv1.weight"] = (cout, cin, 3, 3)
    out[f"{prefix}.conv1.bias"] = (cout,)
    out[f"{prefix}.norm2.weight"] = (cout,)
    out[f"{prefix}.norm2.bias"] = (cout,)
    out[f"{prefix}.conv2.weight"] = (cout, cout, 3, 3)
    out[f"{prefix}.conv2.bias"] = (cout,)
    if cin != cout:
        out[f"{prefix}.conv_shortcut.weight"] = (cout, cin, 1, 1)
        out[f"{prefix}.conv_shortcut.bias"] = (cout,)


def _mid_shapes(prefix: str, c: int, out):
    for n in ("group_norm",):
        out[f"{prefix}.attentions.0.{n}.weight"] = (c,)
        out[f"{prefix}.attentions.0.{n}.bias"] = (c,)
    for n in ("query", "key", "value", "proj_attn"):
        out[f"{prefix}.attentions.0.{n}.weight"] = (c, c)
        out[f"{prefix}.attentions.0.{n}.bias"] = (c,)
    _resnet_shapes(f"{prefix}.resnets.0", c, c, out)
    _resnet_shapes(f"{prefix}.resnets.1", c, c, out)


def akl_param_shapes(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    """Every parameter of the reference ``AutoencoderKL(**cfg)`` (name -> shape)."""
    boc = list(cfg["block_out_channels"])
    lpb = int(cfg.get("layers_per_block", 1))
    lc = int(cfg.get("latent_channels", 4))
    cin_img = int(cfg.get("in_channels", 3))
    cout_img = int(cfg.get("out_channels", 3))
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    # ---- encoder (vae.py:9-86)
    out["encoder.conv_in.weight"] = (boc[0], cin_img, 3, 3)
    out["encoder.conv_in.bias"] = (boc[0],)
    ch = boc[0]
    for i, co in enumerate(boc):
        for j in range(lpb):
            _resnet_shapes(f"encoder.down_blocks.{i}.resnets.{j}", ch if j == 0 else co, co, out)
        ch = co
        if i != len(boc) - 1:
            out[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = (co, co, 3, 3)
            out[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = (co,)
    _mid_shapes("encoder.mid_block", boc[-1], out)
    out["encoder.conv_norm_out.weight"] = (boc[-1],)
    out["encoder.conv_norm_out.bias"] = (boc[-1],)
    out["encoder.conv_out.weight"] = (2 * lc, boc[-1], 3, 3)
    out["encoder.conv_out.bias"] = (2 * lc,)
    # ---- decoder (vae.py:89-166)
    rev = list(reversed(boc))
    out["decoder.conv_in.weight"] = (rev[0], lc, 3, 3)
    out["decoder.conv_in.bias"] = (rev[0],)
    ch = rev[0]
    for i, co in enumerate(rev):
        for j in range(lpb + 1):
            _resnet_shapes(f"decoder.up_blocks.{i}.resnets.{j}", ch if j == 0 else co, co, out)
        ch = co
        if i != len(boc) - 1:
            out[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (co, co, 3, 3)
            out[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (co,)
    _mid_shapes("decoder.mid_block", rev[0], out)
    out["decoder.conv_norm_out.weight"] = (boc[0],)
    out["decoder.conv_norm_out.bias"] = (boc[0],)
    out["decoder.conv_out.weight"] = (cout_img, boc[0], 3, 3)
    out["decoder.conv_out.bias"] = (cout_img,)
    # ---- quant convs (autoencoder_kl.py:76-77)
    out["quant_conv.weight"] = (2 * lc, 2 * lc, 1, 1)
    out["quant_conv.bias"] = (2 * lc,)
    out["post_quant_conv.weight"] = (lc, lc, 1, 1)
    out["post_quant_conv.bias"] = (lc,)
    return out


def _name_seed(name: str, seed: int) -> int:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF


def _is_norm(name: str) -> bool:
    leaf = name.rsplit(".", 2)[-2]
    return leaf.startswith("norm") or leaf in ("group_norm", "conv_norm_out")


def make_akl_state_dict(cfg: dict = PATHB_AKL_CONFIG, seed: int = 0,
                        affine_jitter: float = 0.0) -> "OrderedDict[str, torch.Tensor]":
    """Random-init state_dict with the reference's key names.

    ``affine_jitter`` > 0 perturbs GroupNorm gamma/beta away from (1, 0) so parity tests
    exercise the affine path (the default init would hide a swapped gamma/beta).
    """
    shapes = akl_param_shapes(cfg)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    fan_in_of: Dict[str, int] = {}
    for name, shp in shapes.items():
        if name.endswith(".weight") and len(shp) >= 2:
            fan_in_of[name[: -len(".weight")]] = int(np.prod(shp[1:]))
    for name, shp in shapes.items():
        g = torch.Generator().manual_seed(_name_seed(name, seed))
        mod = name.rsplit(".", 1)[0]
        if _is_norm(name):
            if name.endswith(".weight"):
                t = torch.ones(shp)
                if affine_jitter:
                    t = t + affine_jitter * torch.randn(shp, generator=g)
            else:
                t = torch.zeros(shp)
                if affine_jitter:
                    t = t + affine_jitter * torch.randn(shp, generator=g)
        else:
            bound = 1.0 / math.sqrt(fan_in_of[mod])
            t = (torch.rand(shp, generator=g) * 2.0 - 1.0) * bound
        sd[name] = t.to(torch.float32)
    return sd


def make_predictor_params(in_frames: int = INPUT_FRAMES, pred_frames: int = PRED_FRAMES,
                          latent_channels: int = 4, seed: int = 0):
    """``nn.Linear(13*4, 12*4)`` default init (``.../pretrained_ae_linear_sevir/train.py:67``)."""
    k, n = in_frames * latent_channels, pred_frames * latent_channels
    bound = 1.0 / math.sqrt(k)
    gw = torch.Generator().manual_seed(_name_seed("predictor.weight", seed))
    gb = torch.Generator().manual_seed(_name_seed("predictor.bias", seed))
    w = (torch.rand((n, k), generator=gw) * 2 - 1) * bound
    b = (torch.rand((n,), generator=gb) * 2 - 1) * bound
    return w.float(), b.float()


def make_loader_events(e: int, h: int, w: int, t_raw: int = 49, seed: int = 7) -> torch.Tensor:
    """uint8 stand-in for a SEVIR HDF5 'vil' dataset [E, H, W, raw_seq_len] (sevir.py:562-566): iid bytes, so every
    window of every event is distinguishable in the loader tests."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (e, h, w, t_raw), generator=g, dtype=torch.uint8)


def make_vil_sequences(n: int, h: int = 384, w: int = 384, t: int = 25, seed: int = 1,
                       kind: str = "smooth") -> torch.Tensor:
    """Synthetic SEVIR-VIL-like uint8 tensor in the on-disk NHWT layout
    (``pipeline/datasets/sevir/sevir.py:403``).

    ``smooth``: advected low-pass fields, roughly half zeros, with hit and miss mass at every
    one of the six thresholds 16/74/133/160/181/219 (``pipeline/metrics.py:107``).
    ``uniform``: iid ``randint(0,256)`` -- worst case for the contingency counts.
    """
    g = torch.Generator().manual_seed(seed)
    if kind == "uniform":
        return torch.randint(0, 256, (n, h, w, t), generator=g, dtype=torch.uint8)
    if kind != "smooth":
        raise ValueError(f"unknown kind {kind!r}")
    gh, gw_ = max(4, h // 24), max(4, w // 24)
    pad = 6
    coarse = torch.randn((n, 1, gh + 2 * pad, gw_ + 2 * pad), generator=g)
    fine = torch.randn((n, 1, 4 * gh + 8 * pad, 4 * gw_ + 8 * pad), generator=g)
    vel = (torch.rand((n, 2), generator=g) - 0.5) * 0.02  # fraction of frame per step
    out = torch.empty((n, h, w, t), dtype=torch.uint8)
    ys = torch.linspace(-1, 1, h)
    xs = torch.linspace(-1, 1, w)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    base = torch.stack([gx, gy], dim=-1)[None]  # [1,h,w,2]
    inner = gh / (gh + 2 * pad)
    for ti in range(t):
        shift = (vel * ti)[:, None, None, :]  # [n,1,1,2]
        grid = (base * inner + shift).to(torch.float32)
        f1 = torch.nn.functional.grid_sample(coarse, grid.expand(n, -1, -1, -1), mode="bicubic",
                                             padding_mode="border", align_corners=False)
        f2 = torch.nn.functional.grid_sample(fine, grid.expand(n, -1, -1, -1), mode="bilinear",
                                             padding_mode="border", align_corners=False)
        f = f1[:, 0] + 0.35 * f2[:, 0] + 0.01 * ti
        v = torch.clamp(f * 110.0, 0.0, 255.0)
        out[..., ti] = v.round().to(torch.uint8)
    return out


# ------------------------------------------------------------------------------------------------
# DLinear predictors (reference: experiments/v1_experiments/pretrained_ae_dlinear_*/train.py)
def _uniform(shape, name: str, seed: int, bound: float) -> torch.Tensor:
    g = torch.Generator().manual_seed(_name_seed(name, seed))
    return ((torch.rand(shape, generator=g) * 2.0 - 1.0) * bound).float()


def make_dlinear_case(variant: str, seed: int = 0, batch: int = 2, c: int = 4, h: int = 6, w: int = 5):
    """Small DLinear parity case: (configs, (w_seasonal, b_seasonal, w_trend, b_trend), latents [B,25,C,h,w]).

    ``shared`` / ``individual``: series over t (13 -> 12), enc_in = C*h*w, kernel 3
    (pretrained_ae_dlinear_sevir / _ind config.yaml:4-9); ``indc_indp``: series over the interleaved
    (t, c) axis (52 -> 48), enc_in = h*w, kernel 5, individual (pretrained_ae_dlinear_indc_indp/config.yaml).
    Weights are random (the reference's 1/L constant init would hide index bugs)."""
    from types import SimpleNamespace
    if variant == "indc_indp":
        cfg = SimpleNamespace(seq_len=INPUT_FRAMES, pred_len=PRED_FRAMES, individual=True, enc_in=h * w, kernel_size=5)
        L, P = INPUT_FRAMES * c, PRED_FRAMES * c
    elif variant in ("shared", "individual"):
        cfg = SimpleNamespace(seq_len=INPUT_FRAMES, pred_len=PRED_FRAMES, individual=variant == "individual",
                              enc_in=c * h * w, kernel_size=3)
        L, P = INPUT_FRAMES, PRED_FRAMES
    else:
        raise ValueError(variant)
    bound = 1.0 / math.sqrt(L)
    lead = (cfg.enc_in,) if cfg.individual else ()
    params = tuple(_uniform(lead + shp, f"dlinear.{variant}.{nm}", seed, bound)
                   for nm, shp in (("ws", (P, L)), ("bs", (P,)), ("wt", (P, L)), ("bt", (P,))))
    g = torch.Generator().manual_seed(_name_seed(f"dlinear.{variant}.lat", seed))
    lat = torch.randn((batch, INPUT_FRAMES + PRED_FRAMES, c, h, w), generator=g).float()
    return cfg, params, lat


# ------------------------------------------------------------------------------------------------
# NLayerDiscriminator (reference: pipeline/models/autoencoderkl/losses/model.py:100-150)
def _normal(shape, name: str, seed: int, mean: float, std: float) -> torch.Tensor:
    g = torch.Generator().manual_seed(_name_seed(name, seed))
    return (torch.randn(shape, generator=g) * std + mean).float()


def make_discriminator_state_dict(input_nc: int = 1, ndf: int = 64, n_layers: int = 3, seed: int = 0,
                                  weight_std: float = 0.02) -> "OrderedDict[str, torch.Tensor]":
    """state_dict of ``NLayerDiscriminator(input_nc, ndf, n_layers)`` after ``.apply(weights_init)`` (conv
    N(0, 0.02), BatchNorm weight N(1, 0.02), bias 0; losses/model.py:6-12) with NON-trivial running statistics,
    so that eval-mode BatchNorm folding is exercised. ``weight_std`` can be raised for better-conditioned
    parity cases (0.02-std weights shrink the activations layer by layer)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["main.0.weight"] = _normal((ndf, input_nc, 4, 4), "disc.main.0.weight", seed, 0.0, weight_std * 4)
    sd["main.0.bias"] = _uniform((ndf,), "disc.main.0.bias", seed, 0.25)
    idx, mult = 2, 1
    for n in range(1, n_layers + 1):
        prev, mult = mult, min(2 ** n, 8)
        cin, cout = ndf * prev, ndf * mult
        sd[f"main.{idx}.weight"] = _normal((cout, cin, 4, 4), f"disc.main.{idx}.weight", seed, 0.0, weight_std)
        sd[f"main.{idx + 1}.weight"] = _normal((cout,), f"disc.main.{idx + 1}.weight", seed, 1.0, 0.02)
        sd[f"main.{idx + 1}.bias"] = _normal((cout,), f"disc.main.{idx + 1}.bias", seed, 0.0, 0.05)
        sd[f"main.{idx + 1}.running_mean"] = _normal((cout,), f"disc.main.{idx + 1}.running_mean", seed, 0.0, 0.05)
        sd[f"main.{idx + 1}.running_var"] = _uniform((cout,), f"disc.main.{idx + 1}.running_var", seed, 0.02) + 0.05
        sd[f"main.{idx + 1}.num_batches_tracked"] = torch.tensor(10, dtype=torch.long)
        idx += 3
    sd[f"main.{idx}.weight"] = _normal((1, ndf * mult, 1, 1), f"disc.main.{idx}.weight", seed, 0.0, 0.05)
    sd[f"main.{idx}.bias"] = _uniform((1,), f"disc.main.{idx}.bias", seed, 0.1)
    return sd


# ------------------------------------------------------------------------------------------------
# PosAwareAE_TF / AE_ViT_2048: name-keyed random parameters over a module's own state_dict surface
def fill_state_dict(module, prefix: str, seed: int = 0, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic, well-conditioned random parameters for every entry of ``module.state_dict()`` (keys and
    shapes come from the module, which mirrors the reference's): conv / linear weights uniform with bound
    gain*sqrt(3/fan_in) (unit-gain, so activations keep their scale through ~100 layers), biases small,
    BatchNorm / LayerNorm weights 1 +- 0.1, biases +- 0.1, running_mean +- 0.1, running_var in [0.8, 1.2],
    embeddings N(0, 1) * 0.5. The same dict loads into the unmodified reference module (golden fixtures)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, t in module.state_dict().items():
        shp = tuple(t.shape)
        leaf = name.rsplit(".", 1)[-1]
        tag = f"{prefix}.{name}"
        if leaf == "num_batches_tracked":
            sd[name] = torch.tensor(10, dtype=torch.long)
        elif leaf == "running_mean":
            sd[name] = _uniform(shp, tag, seed, 0.1)
        elif leaf == "running_var":
            sd[name] = _uniform(shp, tag, seed, 0.2) + 1.0
        elif leaf in ("weight", "in_proj_weight") and len(shp) >= 2:
            fan_in = int(np.prod(shp[1:]))
            if "up.0" in name:            # ConvTranspose2d [cin, cout, 4, 4], stride 2: 4 taps per output pixel
                fan_in = shp[0] * 4
            sd[name] = _uniform(shp, tag, seed, gain * math.sqrt(3.0 / fan_in))
        elif leaf == "weight":            # norm weights
            sd[name] = _uniform(shp, tag, seed, 0.1) + 1.0
        elif leaf in ("bias", "in_proj_bias"):
            sd[name] = _uniform(shp, tag, seed, 0.1)
        else:                             # pos_emb, pos_embed, query_vec, dec_queries
            sd[name] = _normal(shp, tag, seed, 0.0, 0.5)
    return sd

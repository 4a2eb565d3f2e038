"""Drop-in for the reference ``pipeline.models.autoencoderkl.autoencoder_kl.AutoencoderKL``.

Same constructor arguments, the same ``state_dict`` key names and shapes (so a reference
checkpoint loads with ``strict=True``), and the same methods: ``encode(x) -> posterior``,
``decode(z)``, ``forward(sample, sample_posterior, return_posterior, generator)``,
``enable_slicing()/disable_slicing()`` (reference autoencoder_kl.py:37-140). The parameters are
held in ordinary ``nn`` containers whose own ``forward`` is never used: ``encode`` / ``decode``
run the sm_100a kernel sequence of ``engine.AKLEngine``. Inference only (the reference freezes
this module: experiments/v1_experiments/pretrained_ae_linear_sevir/train.py:25-30).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from ...engine import AKLEngine
from .distributions import DiagonalGaussianDistribution


def _resnet(cin: int, cout: int, groups: int) -> nn.Module:
    m = nn.Module()
    m.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
    m.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
    m.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
    m.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
    if cin != cout:
        m.conv_shortcut = nn.Conv2d(cin, cout, 1)
    return m


def _mid(c: int, groups: int) -> nn.Module:
    m = nn.Module()
    attn = nn.Module()
    attn.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
    attn.query = nn.Linear(c, c)
    attn.key = nn.Linear(c, c)
    attn.value = nn.Linear(c, c)
    attn.proj_attn = nn.Linear(c, c)
    m.attentions = nn.ModuleList([attn])
    m.resnets = nn.ModuleList([_resnet(c, c, groups), _resnet(c, c, groups)])
    return m


def _sampler(c: int, stride: int) -> nn.Module:
    m = nn.Module()
    m.conv = nn.Conv2d(c, c, 3, stride=stride, padding=1 if stride == 1 else 0)
    return m


class _Encoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, groups):
        super().__init__()
        boc = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        ch = boc[0]
        for i, co in enumerate(boc):
            blk = nn.Module()
            blk.resnets = nn.ModuleList([_resnet(ch if j == 0 else co, co, groups) for j in range(layers_per_block)])
            if i != len(boc) - 1:
                blk.downsamplers = nn.ModuleList([_sampler(co, 2)])
            self.down_blocks.append(blk)
            ch = co
        self.mid_block = _mid(boc[-1], groups)
        self.conv_norm_out = nn.GroupNorm(groups, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * out_channels, 3, padding=1)


class _Decoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, groups):
        super().__init__()
        rev = list(reversed(block_out_channels))
        self.conv_in = nn.Conv2d(in_channels, rev[0], 3, padding=1)
        self.up_blocks = nn.ModuleList()
        ch = rev[0]
        for i, co in enumerate(rev):
            blk = nn.Module()
            blk.resnets = nn.ModuleList([_resnet(ch if j == 0 else co, co, groups) for j in range(layers_per_block + 1)])
            if i != len(rev) - 1:
                blk.upsamplers = nn.ModuleList([_sampler(co, 1)])
            self.up_blocks.append(blk)
            ch = co
        self.mid_block = _mid(rev[0], groups)
        self.conv_norm_out = nn.GroupNorm(groups, block_out_channels[0], eps=1e-6)
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, 3, padding=1)


class AutoencoderKL(nn.Module):
    def __init__(
        self,
        in_channels: int = 3,
        out_channels: int = 3,
        down_block_types: Tuple[str] = ("DownEncoderBlock2D",),
        up_block_types: Tuple[str] = ("UpDecoderBlock2D",),
        block_out_channels: Tuple[int] = (64,),
        layers_per_block: int = 1,
        act_fn: str = "silu",
        latent_channels: int = 4,
        norm_num_groups: int = 32,
        sample_size: int = 32,
        scaling_factor: float = 0.18215,
    ):
        super().__init__()
        for bt in down_block_types:
            if bt != "DownEncoderBlock2D":
                raise ValueError(f"{bt} does not exist.")  # unet_2d_blocks.py:52
        for bt in up_block_types:
            if bt != "UpDecoderBlock2D":
                raise ValueError(f"{bt} does not exist.")  # unet_2d_blocks.py:86
        if act_fn not in ("silu", "swish"):
            raise ValueError(f"act_fn={act_fn!r}: only silu/swish has a B200 kernel")
        if len(down_block_types) != len(block_out_channels) or len(up_block_types) != len(block_out_channels):
            raise ValueError("block type lists and block_out_channels must have the same length")
        self._cfg = dict(in_channels=in_channels, out_channels=out_channels,
                         down_block_types=list(down_block_types), up_block_types=list(up_block_types),
                         block_out_channels=list(block_out_channels), layers_per_block=layers_per_block,
                         latent_channels=latent_channels, norm_num_groups=norm_num_groups)
        self.encoder = _Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.decoder = _Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.use_slicing = False
        self._engine: Optional[AKLEngine] = None
        self._engine_key = None
        self.requires_grad_(False)
        self.eval()

    # ---- engine management: (re)pack weights whenever the parameters may have changed
    def _get_engine(self, device: torch.device) -> AKLEngine:
        key = (str(device), tuple(p._version for p in self.parameters()), tuple(p.data_ptr() for p in self.parameters()))
        if self._engine is None or self._engine_key != key:
            sd = {k: v for k, v in self.state_dict().items()}
            self._engine = AKLEngine(self._cfg, sd, device=device)
            self._engine_key = key
        return self._engine

    @staticmethod
    def _check_input(x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("weatherforecastingtoolkit_b200.AutoencoderKL runs on a B200 only: move the input to "
                               "cuda (there is no CPU fallback)")
        return x.detach().to(torch.float32).contiguous()

    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> DiagonalGaussianDistribution:
        x = self._check_input(x)
        moments = self._get_engine(x.device).encode_moments(x)
        return DiagonalGaussianDistribution(moments)

    def check_finite(self, sync: bool = True) -> None:
        """Raise RuntimeError if any encode / decode of this model on its device produced inf / NaN (an fp16 activation
        beyond 65504). ``sync=True`` synchronises the current stream first, so every call made so far is covered;
        every encode / decode also reports, for free, what earlier completed calls found."""
        if self._engine is not None:
            self._engine.raise_if_nonfinite(sync=sync)

    def repack(self) -> None:
        """Drop the packed fp16 weights: call after writing parameters through ``.data`` (EMA swaps, manual init), which
        does not bump ``Parameter._version`` and is therefore invisible to the automatic check in ``_get_engine``."""
        self._engine = None
        self._engine_key = None

    @torch.no_grad()
    def _decode(self, z: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        z = self._check_input(z)
        return self._get_engine(z.device).decode(z, out=out)

    def decode_into(self, z: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """``decode`` writing into a caller-owned contiguous fp32 tensor (no reference counterpart: saves the
        clone + ``torch.cat`` of the per-chunk loop in ``rollout.Autoencoder.decode``)."""
        return self._decode(z, out=out)

    def enable_slicing(self):
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        if self.use_slicing and z.shape[0] > 1:
            return torch.cat([self._decode(z_slice) for z_slice in z.split(1)])
        return self._decode(z)

    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, return_posterior: bool = False,
                generator: Optional[torch.Generator] = None):
        posterior = self.encode(sample)
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        dec = self.decode(z)
        if return_posterior:
            return dec, posterior
        return dec

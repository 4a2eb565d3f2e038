"""PatchGAN discriminator forward scoring on libwfk_b200.so (BASELINE config 5).

Drop-in for ``pipeline/models/autoencoderkl/losses/model.py:100-150`` (``NLayerDiscriminator``) and
``losses/contperceptual.py:19-23`` (``hinge_d_loss``) of the reference: same constructor, the same ``main``
``nn.Sequential`` layout (hence ``state_dict`` keys), ``weights_init``. ``forward`` is inference scoring:
eval-mode BatchNorm (running statistics) / ActNorm are folded into the convolution weights at pack time, the
stem and the logit head are direct kernels, the 4x4 convolutions run on the tcgen05 conv-GEMM with LeakyReLU
in the epilogue. Training-mode BatchNorm (batch statistics) belongs to the training loop, which is out of
scope: ``forward`` raises in training mode. No CPU path.
"""
from __future__ import annotations

import functools
from typing import Dict, Tuple

import torch
import torch.nn as nn

from .... import _cabi
from ....netprog import NetProgram, fold_bn, pack_conv

LRELU_SLOPE = 0.2


def weights_init(m):
    """losses/model.py:6-12."""
    classname = m.__class__.__name__
    if classname.find('Conv') != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find('BatchNorm') != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


class ActNorm(nn.Module):
    """Parameter container of the reference ActNorm (losses/model.py:15-97): h = scale * (x + loc). Data-dependent
    initialisation is a training-time step and is not reproduced; load trained parameters instead."""

    def __init__(self, num_features, logdet=False, affine=True, allow_reverse_init=False):
        assert affine
        super().__init__()
        self.logdet = logdet
        self.loc = nn.Parameter(torch.zeros(1, num_features, 1, 1))
        self.scale = nn.Parameter(torch.ones(1, num_features, 1, 1))
        self.allow_reverse_init = allow_reverse_init
        self.register_buffer('initialized', torch.tensor(0, dtype=torch.uint8))


class NLayerDiscriminator(nn.Module):
    """``NLayerDiscriminator(input_nc=3, ndf=64, n_layers=3, use_actnorm=False)`` (losses/model.py:100-150)."""

    def __init__(self, input_nc=3, ndf=64, n_layers=3, use_actnorm=False):
        super().__init__()
        norm_layer = nn.BatchNorm2d if not use_actnorm else ActNorm
        if type(norm_layer) == functools.partial:
            use_bias = norm_layer.func != nn.BatchNorm2d
        else:
            use_bias = norm_layer != nn.BatchNorm2d
        kw, padw = 4, 1
        sequence = [nn.Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw), nn.LeakyReLU(LRELU_SLOPE, True)]
        nf_mult = 1
        for n in range(1, n_layers):
            nf_mult_prev = nf_mult
            nf_mult = min(2 ** n, 8)
            sequence += [nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=2, padding=padw, bias=use_bias),
                         norm_layer(ndf * nf_mult), nn.LeakyReLU(LRELU_SLOPE, True)]
        nf_mult_prev = nf_mult
        nf_mult = min(2 ** n_layers, 8)
        sequence += [nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=1, padding=padw, bias=use_bias),
                     norm_layer(ndf * nf_mult), nn.LeakyReLU(LRELU_SLOPE, True)]
        sequence += [nn.Conv2d(ndf * nf_mult, 1, kernel_size=1, stride=1, padding=padw)]
        self.main = nn.Sequential(*sequence)
        self.input_nc, self.ndf, self.n_layers = input_nc, ndf, n_layers
        if input_nc != 1:
            raise ValueError("the B200 path scores single-channel VIL frames (disc_in_channels=1 in every reference "
                             "config that builds a discriminator); input_nc != 1 is unsupported")
        if ndf % 64:
            raise ValueError("ndf must be a multiple of 64 (one 64-channel K block per 4x4 tap)")
        self._packed = None
        self._programs: Dict[Tuple, "_DiscProgram"] = {}

    # ------------------------------------------------------------------ packing
    def _pack(self, device):
        key = (str(device), sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers()))
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1]
        mods = list(self.main)
        f32 = dict(device=device, dtype=torch.float32)
        with torch.no_grad():
            stem = mods[0]
            pk = {"stem_w": stem.weight.detach().to(**f32).reshape(stem.out_channels, 16).t().contiguous(),
                  "stem_b": stem.bias.detach().to(**f32).contiguous(), "mid": []}
            i = 2
            while i + 2 < len(mods):
                conv, norm = mods[i], mods[i + 1]
                w = conv.weight.detach().to(**f32)
                b = conv.bias.detach().to(**f32) if conv.bias is not None else None
                if isinstance(norm, nn.BatchNorm2d):
                    wf, bf = fold_bn(w, b, norm.weight.detach().to(**f32), norm.bias.detach().to(**f32),
                                     norm.running_mean.to(**f32), norm.running_var.to(**f32), norm.eps)
                else:  # ActNorm: scale * (conv + bias + loc)
                    sc, loc = norm.scale.detach().to(**f32).reshape(-1), norm.loc.detach().to(**f32).reshape(-1)
                    wf = w * sc.view(-1, 1, 1, 1)
                    bf = ((b if b is not None else 0) + loc) * sc
                pk["mid"].append((pack_conv(wf), bf.contiguous(), conv.stride[0]))
                i += 3
            head = mods[-1]
            pk["head_w"] = head.weight.detach().to(**f32).reshape(-1).contiguous()
            pk["head_b"] = float(head.bias.detach().item())
            pk["head_pad"] = int(head.padding[0])
        self._packed = (key, pk)
        self._programs.clear()
        return pk

    @torch.no_grad()
    def forward(self, input: torch.Tensor) -> torch.Tensor:
        """input [N, 1, H, W] float32 CUDA -> logits [N, 1, h', w'] float32 (losses/model.py:152-154)."""
        if self.training:
            raise RuntimeError("NLayerDiscriminator on the B200 path is inference scoring: call .eval() first "
                               "(training-mode BatchNorm is not part of the rebuilt hot path)")
        if not input.is_cuda:
            raise RuntimeError("this path runs on a B200 only (no CPU fallback): pass CUDA tensors")
        if input.ndim != 4 or input.shape[1] != 1:
            raise ValueError(f"expected [N, 1, H, W], got {tuple(input.shape)}")
        pk = self._pack(input.device)
        key = (str(input.device), tuple(input.shape))
        prog = self._programs.get(key)
        if prog is None:
            prog = _DiscProgram(pk, tuple(input.shape), input.device)
            self._programs[key] = prog
        return prog(input)


class _DiscProgram(NetProgram):
    def __init__(self, pk, shape, device):
        super().__init__(device)
        n, _, h, w = shape
        lib = self.lib
        self.input = torch.empty(shape, dtype=torch.float32, device=self.dev)
        c0 = pk["stem_w"].shape[1]
        x = self.buf((n, h // 2, w // 2, c0))
        self.add(lib.wfk_conv4x4s2_c1in,
                 (self.input.data_ptr(), n, h, w, pk["stem_w"].data_ptr(), pk["stem_b"].data_ptr(), c0,
                  _cabi.ACT_LEAKY_RELU, LRELU_SLOPE, x.data_ptr(), 0, None, None, None), "disc.stem")
        for li, (wt, bias, stride) in enumerate(pk["mid"]):
            if stride == 2:
                y, _ = self.conv4x4_s2(x, wt, bias=bias, act=_cabi.ACT_LEAKY_RELU, slope=LRELU_SLOPE, what=f"disc.conv{li + 1}")
            else:
                y, _ = self.conv_s1(x, wt, 4, 1, bias=bias, act=_cabi.ACT_LEAKY_RELU, slope=LRELU_SLOPE,
                                    what=f"disc.conv{li + 1}")
            self.free(x)
            x = y
        _, fh, fw, fc = x.shape
        pad = pk["head_pad"]
        self.output = torch.empty((n, 1, fh + 2 * pad, fw + 2 * pad), dtype=torch.float32, device=self.dev)
        self.add(lib.wfk_conv1x1_cout1, (x.data_ptr(), n, fh, fw, fc, pk["head_w"].data_ptr(), pk["head_b"], pad,
                                         self.output.data_ptr()), "disc.head")
        self.keep.append((x, pk))

    def __call__(self, inp: torch.Tensor) -> torch.Tensor:
        self.input.copy_(inp)
        self.run()
        return self.output.clone()


def _logit_sums(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("this path runs on a B200 only (no CPU fallback): pass CUDA tensors")
    lib = _cabi.init(x.device.index if x.device.index is not None else 0)
    x = x.detach().to(torch.float32).contiguous()
    sums = torch.zeros(3, dtype=torch.float64, device=x.device)
    _cabi.check(lib.wfk_logit_sums(x.data_ptr(), x.numel(), sums.data_ptr(),
                                   torch.cuda.current_stream(x.device).cuda_stream), "wfk_logit_sums")
    return sums


def hinge_d_loss(logits_real: torch.Tensor, logits_fake: torch.Tensor) -> torch.Tensor:
    """losses/contperceptual.py:19-23: 0.5 * (mean(relu(1 - real)) + mean(relu(1 + fake)))."""
    sr, sf = _logit_sums(logits_real), _logit_sums(logits_fake)
    return (0.5 * (sr[1] / logits_real.numel() + sf[2] / logits_fake.numel())).to(torch.float32)


def generator_adv_loss(logits_fake: torch.Tensor) -> torch.Tensor:
    """g_loss = -mean(logits_fake) (experiments/v1_experiments/ae_gan_kl/train.py:84-85)."""
    return (-_logit_sums(logits_fake)[0] / logits_fake.numel()).to(torch.float32)

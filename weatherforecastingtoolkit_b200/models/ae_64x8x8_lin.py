"""PosAwareAE_TF (2048-d bottleneck conv autoencoder, BASELINE config 1 / SURVEY row a16) on libwfk_b200.so.

Drop-in for ``pipeline/models/ae_64x8x8_lin.py:52-106`` of the reference: same constructor, same module tree
(``enc``, ``pos_emb``, ``to_latent``, ``from_latent``, ``dec``, ``act``; ``Bottleneck.f``, ``EncBlock.down/res``,
``DecBlock.up/res``), hence the same ``state_dict`` keys and the attributes the experiments touch
(``dec[-1].weight``, experiments/ae_v2_2/train.py:123-124). The ``nn`` modules only hold parameters; ``encode`` /
``decode`` / ``forward`` run a static program of sm_100a kernels (inference, eval-mode BatchNorm):

* every BatchNorm that FOLLOWS a convolution is folded into its weights; every pre-activation ``BN -> GELU`` in
  front of a convolution is produced by the PREVIOUS convolution's epilogue as a second output
  (``out2 = gelu(scale*v + shift)``), so a bottleneck is three conv-GEMM launches and no elementwise pass;
* Conv 4x4 stride 2 = 16 taps on a parity view, ConvTranspose 4x4 stride 2 = four 2x2 sub-pixel convolutions,
  grouped 3x3 = block-diagonal dense 3x3 (the groups are 4-32 channels wide, below a tensor-core tile);
* ``pos_emb`` is folded into the ``to_latent`` bias, the NCHW ``flatten`` order into the linear weights.

The model is hard-wired to 128x128 inputs like the reference (``pos_emb`` is (1, 64, 8, 8)).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

from .. import _cabi
from ..netprog import NetProgram, bn_affine, fold_bn, pack_conv, pack_convT4x4, pack_grouped3x3, pack_linear

GELU = _cabi.ACT_GELU


class Bottleneck(nn.Module):
    """Pre-activation bottleneck (ae_64x8x8_lin.py:7-23): x + f(x)."""

    def __init__(self, channels: int, groups: int = 8):
        super().__init__()
        mid = channels // 4
        g = min(groups, mid)
        assert mid % g == 0, f"groups ({g}) must divide mid channels ({mid})"
        self.f = nn.Sequential(
            nn.BatchNorm2d(channels), nn.GELU(), nn.Conv2d(channels, mid, 1, bias=False),
            nn.BatchNorm2d(mid), nn.GELU(), nn.Conv2d(mid, mid, 3, padding=1, groups=g, bias=False),
            nn.BatchNorm2d(mid), nn.GELU(), nn.Conv2d(mid, channels, 1, bias=False))


class EncBlock(nn.Module):
    """ae_64x8x8_lin.py:28-37."""

    def __init__(self, in_ch: int, out_ch: int, num_blocks: int = 2, groups: int = 8):
        super().__init__()
        self.down = nn.Sequential(nn.Conv2d(in_ch, out_ch, 4, stride=2, padding=1, bias=False),
                                  nn.BatchNorm2d(out_ch), nn.GELU())
        self.res = nn.Sequential(*[Bottleneck(out_ch, groups) for _ in range(num_blocks)])


class DecBlock(nn.Module):
    """ae_64x8x8_lin.py:39-48."""

    def __init__(self, in_ch: int, out_ch: int, num_blocks: int = 2, groups: int = 8):
        super().__init__()
        self.up = nn.Sequential(nn.ConvTranspose2d(in_ch, out_ch, 4, stride=2, padding=1, bias=False),
                                nn.BatchNorm2d(out_ch), nn.GELU())
        self.res = nn.Sequential(*[Bottleneck(out_ch, groups) for _ in range(num_blocks)])


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32)


def _bn_args(bn: nn.BatchNorm2d, device):
    return (_f32(bn.weight, device), _f32(bn.bias, device), _f32(bn.running_mean, device),
            _f32(bn.running_var, device), bn.eps)


class PosAwareAE_TF(nn.Module):
    """``PosAwareAE_TF(in_channels=1, latent_channels=64, groups=8, latent_dim=2048)`` (ae_64x8x8_lin.py:53-86)."""

    def __init__(self, in_channels: int = 1, latent_channels: int = 64, groups: int = 8, latent_dim: int = 2048):
        super().__init__()
        if in_channels != 1:
            raise ValueError("the B200 path encodes single-channel VIL frames (in_channels=1)")
        if latent_channels % 8 or latent_dim % 8:
            raise ValueError("latent_channels and latent_dim must be multiples of 8")
        self.latent_channels = latent_channels
        self.latent_dim = latent_dim
        self.enc = nn.Sequential(
            EncBlock(in_channels, 256, num_blocks=4, groups=groups),
            EncBlock(256, 512, num_blocks=4, groups=groups),
            EncBlock(512, 1024, num_blocks=4, groups=groups),
            EncBlock(1024, 1024, num_blocks=4, groups=groups),
            nn.Conv2d(1024, latent_channels, 1))
        self.pos_emb = nn.Parameter(torch.randn(1, latent_channels, 8, 8))
        self.to_latent = nn.Linear(8 * 8 * latent_channels, latent_dim)
        self.from_latent = nn.Linear(latent_dim, 8 * 8 * latent_channels)
        self.dec = nn.Sequential(
            nn.Conv2d(latent_channels, 1024, 1),
            DecBlock(1024, 1024, num_blocks=4, groups=groups),
            DecBlock(1024, 512, num_blocks=4, groups=groups),
            DecBlock(512, 256, num_blocks=4, groups=groups),
            DecBlock(256, 128, num_blocks=4, groups=groups),
            nn.Conv2d(128, in_channels, 3, padding=1))
        self.act = nn.Sigmoid()
        self._packed = None
        self._programs: Dict[Tuple, "_PosAwareProgram"] = {}

    # ------------------------------------------------------------------ packing
    def _pack_bottleneck(self, b: Bottleneck, device):
        f = b.f
        w1, b1 = fold_bn(_f32(f[2].weight, device), None, *_bn_args(f[3], device))      # conv1 followed by BN2
        w2, b2 = fold_bn(_f32(f[5].weight, device), None, *_bn_args(f[6], device))      # conv2 followed by BN3
        return {"pre": bn_affine(*_bn_args(f[0], device)),                                # BN1 (pre-activation affine)
                "w1": pack_linear(w1.reshape(w1.shape[0], w1.shape[1])), "b1": b1.contiguous(),
                "w2": pack_grouped3x3(w2, f[5].groups), "b2": b2.contiguous(),
                "w3": pack_linear(_f32(f[8].weight, device).reshape(f[8].out_channels, f[8].in_channels))}

    def _pack(self, device):
        key = (str(device), sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers()))
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1]
        lc = self.latent_channels
        with torch.no_grad():
            pk = {"enc": [], "dec": []}
            for i in range(4):
                blk = self.enc[i]
                w, b = fold_bn(_f32(blk.down[0].weight, device), None, *_bn_args(blk.down[1], device))
                if i == 0:
                    e = {"stem_w": w.reshape(w.shape[0], 16).t().contiguous(), "b": b.contiguous()}
                else:
                    e = {"w": pack_conv(w), "b": b.contiguous()}
                e["res"] = [self._pack_bottleneck(bt, device) for bt in blk.res]
                pk["enc"].append(e)
            head = self.enc[4]
            pk["enc_head_w"] = pack_linear(_f32(head.weight, device).reshape(lc, -1))
            pk["enc_head_b"] = _f32(head.bias, device).contiguous()
            # to_latent on the NHWC flatten (y, x, c) instead of the reference's (c, y, x); pos_emb folded into the bias
            wl = _f32(self.to_latent.weight, device)                                      # [latent_dim, lc*64]
            pe = _f32(self.pos_emb, device).reshape(-1)                                   # (c, y, x)
            pk["to_latent_b"] = (_f32(self.to_latent.bias, device) + wl @ pe).contiguous()
            pk["to_latent_w"] = pack_linear(wl.reshape(-1, lc, 64).permute(0, 2, 1).reshape(-1, 64 * lc))
            wf = _f32(self.from_latent.weight, device)                                    # [lc*64, latent_dim]
            pk["from_latent_w"] = pack_linear(wf.reshape(lc, 64, -1).permute(1, 0, 2).reshape(64 * lc, -1))
            pk["from_latent_b"] = _f32(self.from_latent.bias, device).reshape(lc, 64).t().reshape(-1).contiguous()
            stem = self.dec[0]
            pk["dec_stem_w"] = pack_linear(_f32(stem.weight, device).reshape(stem.out_channels, lc))
            pk["dec_stem_b"] = _f32(stem.bias, device).contiguous()
            for i in range(1, 5):
                blk = self.dec[i]
                w, b = fold_bn(_f32(blk.up[0].weight, device), None, *_bn_args(blk.up[1], device), out_dim=1)
                pk["dec"].append({"w": pack_convT4x4(w), "b": b.contiguous(),
                                  "res": [self._pack_bottleneck(bt, device) for bt in blk.res]})
            tail = self.dec[5]
            pk["tail_w"] = _f32(tail.weight, device).permute(0, 2, 3, 1).reshape(1, 9, -1).contiguous().to(torch.float16)
            pk["tail_b"] = _f32(tail.bias, device).contiguous()
        self._packed = (key, pk)
        self._programs.clear()
        return pk

    def _program(self, kind: str, shape, device) -> "_PosAwareProgram":
        if self.training:
            raise RuntimeError("PosAwareAE_TF on the B200 path is inference only: call .eval() first "
                               "(training-mode BatchNorm is not part of the rebuilt hot path)")
        pk = self._pack(device)
        key = (kind, str(device), tuple(shape))
        prog = self._programs.get(key)
        if prog is None:
            prog = _PosAwareProgram(self, pk, kind, tuple(shape), device)
            self._programs[key] = prog
        return prog

    # ------------------------------------------------------------------ reference interface
    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, 1, 128, 128] float32 CUDA -> z [B, latent_dim] float32 (ae_64x8x8_lin.py:88-94)."""
        if not x.is_cuda:
            raise RuntimeError("this path runs on a B200 only (no CPU fallback): pass CUDA tensors")
        if x.ndim != 4 or tuple(x.shape[1:]) != (1, 128, 128):
            raise ValueError(f"PosAwareAE_TF is hard-wired to [B, 1, 128, 128] inputs (pos_emb is 8x8), got {tuple(x.shape)}")
        return self._program("enc", x.shape, x.device)(x.to(torch.float32))

    @torch.no_grad()
    def decode(self, z_flat: torch.Tensor) -> torch.Tensor:
        """z [B, latent_dim] float32 CUDA -> [B, 1, 128, 128] float32 in (0, 1) (ae_64x8x8_lin.py:96-103)."""
        if not z_flat.is_cuda:
            raise RuntimeError("this path runs on a B200 only (no CPU fallback): pass CUDA tensors")
        if z_flat.ndim != 2 or z_flat.shape[1] != self.latent_dim:
            raise ValueError(f"expected [B, {self.latent_dim}], got {tuple(z_flat.shape)}")
        return self._program("dec", z_flat.shape, z_flat.device)(z_flat.to(torch.float32))

    def forward(self, x):
        z = self.encode(x)
        return self.decode(z), z


class _PosAwareProgram(NetProgram):
    def __init__(self, model: PosAwareAE_TF, pk, kind: str, shape, device):
        super().__init__(device)
        self.pk = pk
        self.input = torch.empty(shape, dtype=torch.float32, device=self.dev)
        self.output = self._build_encoder(model, shape) if kind == "enc" else self._build_decoder(model, shape)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        self.input.copy_(x)
        self.run()
        return self.output.clone()

    def _bottlenecks(self, x, a, blocks, next_pre, tag):
        """x: raw stream [n, h, w, C]; a = gelu(bn1(x)) of the first block. Returns (x_out, a_out) where a_out is
        gelu(next_pre(x_out)) (None when ``next_pre`` is None)."""
        n, h, w, c = x.shape
        for bi, bk in enumerate(blocks):
            mid = bk["b1"].numel()
            h1 = self.gemm(a.view(-1, c), bk["w1"], bias=bk["b1"], act=GELU, what=f"{tag}.res{bi}.conv1")
            self.free(a)
            h2, _ = self.conv_s1(h1.view(n, h, w, mid), bk["w2"], 3, 1, bias=bk["b2"], act=GELU, what=f"{tag}.res{bi}.conv2")
            self.free(h1)
            pre = blocks[bi + 1]["pre"] if bi + 1 < len(blocks) else next_pre
            if pre is not None:
                xn, a = self.gemm(h2.view(-1, mid), bk["w3"], residual=x, out2=True, scale2=pre[0], shift2=pre[1],
                                  act2=GELU, what=f"{tag}.res{bi}.conv3+res")
                a = a.view(n, h, w, c)
            else:
                xn, a = self.gemm(h2.view(-1, mid), bk["w3"], residual=x, what=f"{tag}.res{bi}.conv3+res"), None
            self.free(h2, x)
            x = xn.view(n, h, w, c)
        return x, a

    def _build_encoder(self, model, shape):
        pk, lib = self.pk, self.lib
        n, _, hh, ww = shape
        x = a = None
        for i, e in enumerate(pk["enc"]):
            pre = e["res"][0]["pre"]
            if i == 0:
                c0 = e["b"].numel()
                x, a = self.buf((n, hh // 2, ww // 2, c0)), self.buf((n, hh // 2, ww // 2, c0))
                self.add(lib.wfk_conv4x4s2_c1in,
                         (self.input.data_ptr(), n, hh, ww, e["stem_w"].data_ptr(), e["b"].data_ptr(), c0, GELU, 0.0,
                          x.data_ptr(), GELU, pre[0].data_ptr(), pre[1].data_ptr(), a.data_ptr()), "enc0.down")
            else:
                xin = x
                x, a = self.conv4x4_s2(xin, e["w"], bias=e["b"], act=GELU, out2=True, scale2=pre[0], shift2=pre[1],
                                       act2=GELU, what=f"enc{i}.down")
                self.free(xin)
            x, _ = self._bottlenecks(x, a, e["res"], None, f"enc{i}")
        nb, fh, fw, fc = x.shape
        lc = model.latent_channels
        z = self.gemm(x.view(-1, fc), pk["enc_head_w"], bias=pk["enc_head_b"], what="enc.head1x1")
        self.free(x)
        out = torch.empty((n, model.latent_dim), dtype=torch.float32, device=self.dev)
        self.gemm(z.view(n, fh * fw * lc), pk["to_latent_w"], bias=pk["to_latent_b"], out_f32=True, out=out,
                  what="to_latent(+pos_emb)")
        self.keep.append(z)
        return out

    def _build_decoder(self, model, shape):
        pk, lib = self.pk, self.lib
        n, ld = shape
        lc = model.latent_channels
        zh = self.buf((n, ld))
        self.add(lib.wfk_f32_to_f16, (self.input.data_ptr(), n * ld, zh.data_ptr()), "latent->fp16")
        z = self.gemm(zh, pk["from_latent_w"], bias=pk["from_latent_b"], what="from_latent")   # [n, (y, x, c)]
        self.free(zh)
        x = self.gemm(z.view(n * 64, lc), pk["dec_stem_w"], bias=pk["dec_stem_b"], what="dec.stem1x1").view(n, 8, 8, -1)
        self.free(z)
        for i, e in enumerate(pk["dec"]):
            pre = e["res"][0]["pre"]
            xin = x
            x, a = self.convT4x4_s2(xin, e["w"], bias=e["b"], act=GELU, out2=True, scale2=pre[0], shift2=pre[1], act2=GELU,
                                    what=f"dec{i + 1}.up")
            self.free(xin)
            x, _ = self._bottlenecks(x, a, e["res"], None, f"dec{i + 1}")
        nb, H, W, c = x.shape
        out = torch.empty((n, 1, H, W), dtype=torch.float32, device=self.dev)
        self.add(lib.wfk_conv3x3_small_cout_act,
                 (x.data_ptr(), n, H, W, c, pk["tail_w"].data_ptr(), pk["tail_b"].data_ptr(), 1, None, None,
                  _cabi.ACT_SIGMOID, out.data_ptr()), "dec.tail3x3+sigmoid")
        self.keep.append(x)
        return out

"""AE_ViT_2048 (ViT autoencoder with a [64, 512] token sequence, BASELINE config 4 / SURVEY row a18) on
libwfk_b200.so.

Drop-in for ``pipeline/models/ae_vit.py:84-162`` of the reference: same constructor, same module tree
(``patch_embed, pos_embed, encoder, query_vec, to_latent, dec_queries, from_latent, decoder, unpatch``), hence the
same ``state_dict`` keys. The ``nn`` modules only hold parameters; ``forward`` runs a static program of sm_100a
kernels (inference: dropout is the identity in eval mode):

* patch embedding / unpatch: 16x16 patches <-> rows, then a tcgen05 GEMM (K = 256 / N = 256);
* each post-norm ``TransformerEncoderLayer``: in_proj GEMM -> ``wfk_mha_small`` (64 tokens x 8 heads of 64, in
  shared memory) -> out_proj GEMM with the residual added in the epilogue (fp32 out) -> LayerNorm -> linear1 GEMM
  with GELU in the epilogue -> linear2 GEMM + residual (fp32 out) -> LayerNorm;
* ``GlobalCrossEncode``: ``q_proj(query_vec)`` is input independent and is precomputed at pack time;
  ``GlobalCrossDecode`` attends over ONE key/value token, so its softmax is exactly 1 and the block reduces to
  ``out(v_proj(latent))`` broadcast over the 64 query positions (``dec_queries`` / ``q_proj`` cannot influence the
  result); both are reproduced exactly, not approximated.

Extras next to the reference interface: ``encode_tokens`` / ``decode_tokens`` expose the token-sequence latent
[B, 64, 512] that BASELINE config 4 names.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

from .. import _cabi
from ..netprog import NetProgram, pack_linear

LN_EPS = 1e-5


class GlobalCrossEncode(nn.Module):
    """Parameter container of ae_vit.py:4-42 (sequence of d_token tokens -> one d_latent vector)."""

    def __init__(self, d_token, d_latent, n_heads=8):
        super().__init__()
        assert d_latent % n_heads == 0 and d_token % n_heads == 0
        self.nh, self.dh_q, self.dh_kv = n_heads, d_latent // n_heads, d_token // n_heads
        self.scale = self.dh_q ** -0.5
        self.q_proj = nn.Linear(d_latent, d_latent)
        self.kv_proj = nn.Linear(d_token, 2 * d_latent)
        self.out = nn.Linear(d_latent, d_latent)


class GlobalCrossDecode(nn.Module):
    """Parameter container of ae_vit.py:44-82 (one d_latent vector -> sequence of d_token tokens)."""

    def __init__(self, d_token, d_latent, n_heads=8):
        super().__init__()
        assert d_latent % n_heads == 0 and d_token % n_heads == 0
        self.nh, self.dh_q, self.dh_kv = n_heads, d_token // n_heads, d_latent // n_heads
        self.scale = self.dh_kv ** -0.5
        self.q_proj = nn.Linear(d_token, d_token)
        self.kv_proj = nn.Linear(d_latent, 2 * d_token)
        self.out = nn.Linear(d_token, d_token)


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class AE_ViT_2048(nn.Module):
    """``AE_ViT_2048()`` (ae_vit.py:84-143): 128x128 single-channel frames, 16x16 patches."""

    def __init__(self):
        super().__init__()
        img, patch, ch = 128, 16, 1
        seq = img // patch
        n_patches = seq * seq
        d_token, d_latent = 512, 2048
        depth_enc, depth_dec, heads = 6, 6, 8
        self.seq, self.d_token, self.d_latent, self.heads, self.img = seq, d_token, d_latent, heads, img
        self.patch_embed = nn.Conv2d(ch, d_token, patch, patch)
        self.pos_embed = nn.Parameter(torch.randn(1, n_patches, d_token))

        def layer():
            return nn.TransformerEncoderLayer(d_model=d_token, nhead=heads, dim_feedforward=4 * d_token, dropout=0.1,
                                              activation='gelu', batch_first=True)
        self.encoder = nn.TransformerEncoder(layer(), depth_enc)
        self.query_vec = nn.Parameter(torch.randn(1, 1, d_latent))
        self.to_latent = GlobalCrossEncode(d_token, d_latent, n_heads=heads)
        self.dec_queries = nn.Parameter(torch.randn(1, n_patches, d_token))
        self.from_latent = GlobalCrossDecode(d_token, d_latent, n_heads=heads)
        self.decoder = nn.TransformerEncoder(layer(), depth_dec)
        self.unpatch = nn.ConvTranspose2d(d_token, ch, patch, patch)
        self._packed = None
        self._programs: Dict[Tuple, "_ViTProgram"] = {}

    # ------------------------------------------------------------------ packing
    @staticmethod
    def _pack_layer(l: nn.TransformerEncoderLayer, device):
        return {"in_w": pack_linear(_f32(l.self_attn.in_proj_weight, device)), "in_b": _f32(l.self_attn.in_proj_bias, device),
                "out_w": pack_linear(_f32(l.self_attn.out_proj.weight, device)), "out_b": _f32(l.self_attn.out_proj.bias, device),
                "w1": pack_linear(_f32(l.linear1.weight, device)), "b1": _f32(l.linear1.bias, device),
                "w2": pack_linear(_f32(l.linear2.weight, device)), "b2": _f32(l.linear2.bias, device),
                "n1": (_f32(l.norm1.weight, device), _f32(l.norm1.bias, device), float(l.norm1.eps)),
                "n2": (_f32(l.norm2.weight, device), _f32(l.norm2.bias, device), float(l.norm2.eps))}

    def _pack(self, device):
        key = (str(device), sum(p._version for p in self.parameters()))
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1]
        d, dl = self.d_token, self.d_latent
        with torch.no_grad():
            pk = {"enc": [self._pack_layer(l, device) for l in self.encoder.layers],
                  "dec": [self._pack_layer(l, device) for l in self.decoder.layers]}
            pk["patch_w"] = pack_linear(_f32(self.patch_embed.weight, device).reshape(d, 256))
            pk["patch_b"] = _f32(self.patch_embed.bias, device)
            pk["pos"] = _f32(self.pos_embed, device).reshape(-1, d)
            tl = self.to_latent
            q = torch.nn.functional.linear(_f32(self.query_vec, device).reshape(1, dl), _f32(tl.q_proj.weight, device),
                                           _f32(tl.q_proj.bias, device))
            pk["tl_q"] = (q * tl.scale).reshape(tl.nh, tl.dh_q).contiguous()     # ae_vit.py:26-28, 35: input independent
            pk["tl_kv_w"] = pack_linear(_f32(tl.kv_proj.weight, device))
            pk["tl_kv_b"] = _f32(tl.kv_proj.bias, device)
            pk["tl_out_w"] = pack_linear(_f32(tl.out.weight, device))
            pk["tl_out_b"] = _f32(tl.out.bias, device)
            fl = self.from_latent
            pk["fl_v_w"] = pack_linear(_f32(fl.kv_proj.weight, device)[d:2 * d])  # the value half (ae_vit.py:69-71)
            pk["fl_v_b"] = _f32(fl.kv_proj.bias, device)[d:2 * d].contiguous()
            pk["fl_out_w"] = pack_linear(_f32(fl.out.weight, device))
            pk["fl_out_b"] = _f32(fl.out.bias, device)
            pk["unpatch_w"] = pack_linear(_f32(self.unpatch.weight, device).reshape(d, 256).t())
            pk["unpatch_b"] = _f32(self.unpatch.bias, device).expand(256).contiguous()
        self._packed = (key, pk)
        self._programs.clear()
        return pk

    def _program(self, kind, shape, device) -> "_ViTProgram":
        if self.training:
            raise RuntimeError("AE_ViT_2048 on the B200 path is inference only: call .eval() first (dropout)")
        pk = self._pack(device)
        key = (kind, str(device), tuple(shape))
        prog = self._programs.get(key)
        if prog is None:
            prog = _ViTProgram(self, pk, kind, tuple(shape), device)
            self._programs[key] = prog
        return prog

    def _check_img(self, x):
        if not x.is_cuda:
            raise RuntimeError("this path runs on a B200 only (no CPU fallback): pass CUDA tensors")
        if x.ndim != 4 or tuple(x.shape[1:]) != (1, self.img, self.img):
            raise ValueError(f"AE_ViT_2048 takes [B, 1, {self.img}, {self.img}] frames, got {tuple(x.shape)}")

    # ------------------------------------------------------------------ reference interface
    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """x [B, 1, 128, 128] float32 CUDA -> (out [B, 1, 128, 128], latent [B, 2048]) (ae_vit.py:135-160)."""
        self._check_img(x)
        prog = self._program("full", x.shape, x.device)
        out = prog(x.to(torch.float32))
        return out, prog.latent.clone()

    # ------------------------------------------------------------------ extras (token-sequence latent)
    @torch.no_grad()
    def encode_tokens(self, x: torch.Tensor) -> torch.Tensor:
        """x -> encoder token sequence [B, 64, 512] float32 (ae_vit.py:141-145)."""
        self._check_img(x)
        return self._program("enc_tokens", x.shape, x.device)(x.to(torch.float32))

    @torch.no_grad()
    def decode_tokens(self, z: torch.Tensor) -> torch.Tensor:
        """Decoder stack + unpatch on a token sequence [B, 64, 512] (ae_vit.py:155-159) -> [B, 1, 128, 128]."""
        if not z.is_cuda:
            raise RuntimeError("this path runs on a B200 only (no CPU fallback): pass CUDA tensors")
        if z.ndim != 3 or tuple(z.shape[1:]) != (self.seq * self.seq, self.d_token):
            raise ValueError(f"expected [B, {self.seq * self.seq}, {self.d_token}], got {tuple(z.shape)}")
        return self._program("dec_tokens", z.shape, z.device)(z.to(torch.float32))


class _ViTProgram(NetProgram):
    def __init__(self, model: AE_ViT_2048, pk, kind: str, shape, device):
        super().__init__(device)
        self.pk, self.m = pk, model
        self.input = torch.empty(shape, dtype=torch.float32, device=self.dev)
        self.latent = None
        n = shape[0]
        T, d = model.seq * model.seq, model.d_token
        if kind in ("full", "enc_tokens"):
            x = self._embed(n)
            xf = None
            for i, l in enumerate(pk["enc"]):
                x, xf = self._layer(x, l, n, T, f"enc{i}", want_f32=(kind == "enc_tokens" and i == len(pk["enc"]) - 1))
            if kind == "enc_tokens":
                self.output = xf.view(n, T, d)
                return
            x = self._collapse_expand(x, n, T)
        else:
            x = self.buf((n * T, d))
            self.add(self.lib.wfk_f32_to_f16, (self.input.data_ptr(), n * T * d, x.data_ptr()), "tokens->fp16")
        for i, l in enumerate(pk["dec"]):
            x, _ = self._layer(x, l, n, T, f"dec{i}")
        rows = self.gemm(x, pk["unpatch_w"], bias=pk["unpatch_b"], out_f32=True, what="unpatch")
        self.output = torch.empty((n, 1, model.img, model.img), dtype=torch.float32, device=self.dev)
        self.add(self.lib.wfk_unpatchify16, (rows.data_ptr(), n, model.img, model.img, self.output.data_ptr()), "unpatchify")
        self.keep.append((x, rows))

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        self.input.copy_(x)
        self.run()
        return self.output.clone()

    def _embed(self, n):
        pk, m = self.pk, self.m
        T, d = m.seq * m.seq, m.d_token
        rows = self.buf((n * T, 256))
        self.add(self.lib.wfk_patchify16, (self.input.data_ptr(), n, m.img, m.img, rows.data_ptr()), "patchify")
        # + pos_embed (ae_vit.py:143) rides in as the epilogue's residual operand, expanded over the batch once
        pos = pk["pos"].to(torch.float16).unsqueeze(0).expand(n, T, d).contiguous().view(n * T, d)
        self.keep.append(pos)
        x = self.gemm(rows, pk["patch_w"], bias=pk["patch_b"], residual=pos, what="patch_embed+pos")
        self.free(rows)
        return x

    def _ln(self, s, norm, rows, d, want_f32=False):
        y = self.buf((rows, d))
        yf = self.buf((rows, d), torch.float32) if want_f32 else None
        self.add(self.lib.wfk_layernorm_rows, (s.data_ptr(), rows, d, norm[0].data_ptr(), norm[1].data_ptr(), norm[2],
                                               y.data_ptr(), None if yf is None else yf.data_ptr()), "layernorm")
        return y, yf

    def _layer(self, x, l, n, T, tag, want_f32=False):
        """Post-norm TransformerEncoderLayer (torch.nn, norm_first=False): x = LN1(x + SA(x)); x = LN2(x + FF(x))."""
        d, heads = self.m.d_token, self.m.heads
        qkv = self.gemm(x, l["in_w"], bias=l["in_b"], what=f"{tag}.in_proj")
        att = self.buf((n * T, d))
        self.add(self.lib.wfk_mha_small, (qkv.data_ptr(), n, T, d, heads, att.data_ptr()), f"{tag}.attention")
        self.free(qkv)
        s1 = self.gemm(att, l["out_w"], bias=l["out_b"], residual=x, out_f32=True, what=f"{tag}.out_proj+res")
        self.free(att, x)
        x1, _ = self._ln(s1, l["n1"], n * T, d)
        self.free(s1)
        f = self.gemm(x1, l["w1"], bias=l["b1"], act=_cabi.ACT_GELU, what=f"{tag}.linear1+gelu")
        s2 = self.gemm(f, l["w2"], bias=l["b2"], residual=x1, out_f32=True, what=f"{tag}.linear2+res")
        self.free(f, x1)
        x2, x2f = self._ln(s2, l["n2"], n * T, d, want_f32)
        self.free(s2)
        return x2, x2f

    def _collapse_expand(self, z, n, T):
        """to_latent (GlobalCrossEncode) -> latent [n, 2048]; from_latent (GlobalCrossDecode) + pos_embed."""
        pk, m = self.pk, self.m
        d, dl, heads = m.d_token, m.d_latent, m.heads
        kv = self.gemm(z, pk["tl_kv_w"], bias=pk["tl_kv_b"], what="to_latent.kv_proj")
        self.free(z)
        att = self.buf((n, dl))
        self.add(self.lib.wfk_cross_encode_attn, (pk["tl_q"].data_ptr(), kv.data_ptr(), n, T, dl, heads, att.data_ptr()),
                 "to_latent.attention")
        self.free(kv)
        self.latent = torch.empty((n, dl), dtype=torch.float32, device=self.dev)
        self.gemm(att, pk["tl_out_w"], bias=pk["tl_out_b"], out_f32=True, out=self.latent, what="to_latent.out")
        self.free(att)
        lat_h = self.buf((n, dl))
        self.add(self.lib.wfk_f32_to_f16, (self.latent.data_ptr(), n * dl, lat_h.data_ptr()), "latent->fp16")
        v = self.gemm(lat_h, pk["fl_v_w"], bias=pk["fl_v_b"], what="from_latent.v_proj")
        self.free(lat_h)
        t = self.gemm(v, pk["fl_out_w"], bias=pk["fl_out_b"], out_f32=True, what="from_latent.out")
        self.free(v)
        x = self.buf((n * T, d))
        self.add(self.lib.wfk_bcast_add_rows, (t.data_ptr(), pk["pos"].data_ptr(), n, T, d, x.data_ptr()),
                 "from_latent.broadcast+pos")
        self.keep.append(t)
        return x

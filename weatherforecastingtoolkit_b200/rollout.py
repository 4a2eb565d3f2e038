"""Path-B latent nowcast rollout behind the reference experiment's interfaces.

Mirrors ``experiments/v1_experiments/pretrained_ae_linear_sevir/train.py`` (reference repo):
``Autoencoder`` (the per-frame encode / decode wrapper, train.py:21-56), the ``nn.Linear(13*4,
12*4)`` predictor with its residual framing (train.py:67, 101-113) and ``Model.validation_step``
(train.py:100-120) up to ``log_metrics`` (pipeline/helpers.py:142-153). Host code only sequences
kernels of libwfk_b200.so; PyTorch supplies device buffers, streams and ``torch.distributed``.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _cabi, engine
from . import metrics as wf_metrics
from .models.autoencoderkl import AutoencoderKL
from .synthetic import INPUT_FRAMES, PRED_FRAMES


def _lib_for(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError("this path runs on a B200 only (no CPU fallback): pass CUDA tensors")
    return _cabi.init(t.device.index if t.device.index is not None else 0)


def stage_vil(batch_u8_nhwt: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """uint8 VIL [N, H, W, T] (SEVIR on-disk layout) -> normalised [N, T, 1, H, W].

    Replaces ``SEVIRDataLoader.preprocess_data_dict`` + ``change_layout`` + the
    ``permute(0,3,1,2).unsqueeze(2)`` of the train script (pipeline/datasets/sevir/sevir.py:626-666,
    88-101; train.py:101): ``fl32(1/255) * float(x)``, bit-exact."""
    if batch_u8_nhwt.dtype != torch.uint8 or batch_u8_nhwt.ndim != 4:
        raise TypeError("expected a uint8 [N, H, W, T] tensor")
    lib = _lib_for(batch_u8_nhwt)
    x = batch_u8_nhwt.contiguous()
    n, h, w, t = x.shape
    if dtype not in (torch.float32, torch.float16):
        raise ValueError("dtype must be float32 or float16")
    out = torch.empty((n, t, 1, h, w), dtype=dtype, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with engine.timed_pass("stage_vil", float(x.numel()) * (1 + out.element_size())):
        _cabi.check(lib.wfk_stage_vil_u8(x.data_ptr(), n, h, w, t, out.data_ptr(), 0 if dtype == torch.float32 else 1,
                                         stream), "wfk_stage_vil_u8")
    return out


def sequent_windows(num_events: int, raw_seq_len: int = 49, seq_len: int = 25, stride: int = 12, start: int = 0,
                    count: Optional[int] = None):
    """(event, first frame) pairs in the order of the reference's sequential sampler
    (``SEVIRDataLoader._idx_sample``, pipeline/datasets/sevir/sevir.py:864-877): ``num_seq_per_event =
    1 + (raw_seq_len - seq_len) // stride`` (:327-328) windows per event, events in order; ``start`` / ``count``
    select a batch (index * batch_size, batch_size)."""
    per_event = 1 + (raw_seq_len - seq_len) // stride
    total = per_event * num_events
    stop = total if count is None else min(total, start + count)
    return [(i // per_event, (i % per_event) * stride) for i in range(start, stop)]


def stage_vil_windows(events_u8: torch.Tensor, windows, seq_len: int = 25, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """uint8 SEVIR events [E, H, W, T_raw] resident on the GPU + (event, t0) windows -> normalised [N, seq_len, 1, H, W]:
    the slicing of ``_idx_sample`` (sevir.py:879-889) fused with ``preprocess_data_dict`` + ``change_layout`` (no
    intermediate uint8 batch, no 4x fp32 host-to-device inflation)."""
    if events_u8.dtype != torch.uint8 or events_u8.ndim != 4:
        raise TypeError("expected a uint8 [E, H, W, T_raw] tensor")
    lib = _lib_for(events_u8)
    ev = events_u8.contiguous()
    e, h, w, t_raw = ev.shape
    win = torch.as_tensor(windows, dtype=torch.int32).reshape(-1, 2)
    if win.numel() == 0:
        raise ValueError("no windows")
    if int(win[:, 0].min()) < 0 or int(win[:, 0].max()) >= e or int(win[:, 1].min()) < 0 or int(win[:, 1].max()) + seq_len > t_raw:
        raise ValueError("window out of range")
    if dtype not in (torch.float32, torch.float16):
        raise ValueError("dtype must be float32 or float16")
    n = win.shape[0]
    win_d = win.to(ev.device)
    out = torch.empty((n, seq_len, 1, h, w), dtype=dtype, device=ev.device)
    stream = torch.cuda.current_stream(ev.device).cuda_stream
    _cabi.check(lib.wfk_stage_vil_windows(ev.data_ptr(), e, h, w, t_raw, win_d.data_ptr(), n, seq_len, out.data_ptr(),
                                          0 if dtype == torch.float32 else 1, stream), "wfk_stage_vil_windows")
    return out


class LatentLinearPredictor(nn.Linear):
    """``self.predictor = nn.Linear(input_frames * 4, pred_frames * 4)`` (train.py:67). ``forward`` is
    inherited (training stays PyTorch, out of scope); ``rollout`` is the fused inference kernel."""

    def __init__(self, input_frames: int = INPUT_FRAMES, pred_frames: int = PRED_FRAMES, latent_channels: int = 4):
        super().__init__(input_frames * latent_channels, pred_frames * latent_channels)
        self.input_frames, self.pred_frames, self.latent_channels = input_frames, pred_frames, latent_channels

    @torch.no_grad()
    def rollout(self, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """v [B, t_in + t_out, C, h, w] fp32 latents -> (pred, tgt, val_loss): train.py:102-113 in one
        pass (subtract last input frame, Linear per latent pixel, add it back; F.mse_loss of the
        residual-space prediction)."""
        lib = _lib_for(v)
        b, t, c, h, w = v.shape
        if t != self.input_frames + self.pred_frames or c != self.latent_channels:
            raise ValueError(f"latents {tuple(v.shape)} do not match predictor "
                             f"({self.input_frames}+{self.pred_frames} frames, {self.latent_channels} channels)")
        v = v.detach().to(torch.float32).contiguous()
        wt = self.weight.detach().to(device=v.device, dtype=torch.float32).contiguous()
        bs = self.bias.detach().to(device=v.device, dtype=torch.float32).contiguous()
        pred = torch.empty((b, self.pred_frames, c, h, w), dtype=torch.float32, device=v.device)
        tgt = torch.empty_like(pred)
        loss = torch.zeros(2, dtype=torch.float64, device=v.device)
        stream = torch.cuda.current_stream(v.device).cuda_stream
        # algorithmic traffic: 13 input frames read + 12 predicted frames written = 25 latent frames per
        # sequence (the kernel also reads the 12 target frames and writes them back framed: not counted)
        with engine.timed_pass("predict_linear", 4.0 * v.numel()):
            _cabi.check(lib.wfk_predict_linear(v.data_ptr(), wt.data_ptr(), bs.data_ptr(), b, self.input_frames,
                                               self.pred_frames, c, h * w, pred.data_ptr(), tgt.data_ptr(),
                                               loss.data_ptr(), stream), "wfk_predict_linear")
        return pred, tgt, (loss[0] / loss[1]).to(torch.float32)


    @torch.no_grad()
    def rollout_autoregressive(self, inp: torch.Tensor, blocks: int = 2) -> torch.Tensor:
        """The predictor stepped autoregressively (north_star; the reference experiment applies it once): every call of
        the fused kernel yields ``pred_frames`` frames from the last ``input_frames`` frames of the sequence so far,
        re-anchoring the residual framing of train.py:104-112 on the newest frame. inp [B, input_frames, C, h, w] ->
        [B, pred_frames * blocks, C, h, w]. One kernel launch per block; the window lives in one reused device buffer."""
        lib = _lib_for(inp)
        b, t, c, h, w = inp.shape
        if t != self.input_frames or c != self.latent_channels:
            raise ValueError(f"expected [B, {self.input_frames}, {self.latent_channels}, h, w] input latents")
        if blocks < 1:
            raise ValueError("blocks must be >= 1")
        ti, to = self.input_frames, self.pred_frames
        wt = self.weight.detach().to(device=inp.device, dtype=torch.float32).contiguous()
        bs = self.bias.detach().to(device=inp.device, dtype=torch.float32).contiguous()
        win = torch.zeros((b, ti + to, c, h, w), dtype=torch.float32, device=inp.device)   # kernel layout: inputs | targets
        win[:, :ti] = inp.detach().to(torch.float32)
        out = torch.empty((b, to * blocks, c, h, w), dtype=torch.float32, device=inp.device)
        pred = torch.empty((b, to, c, h, w), dtype=torch.float32, device=inp.device)
        stream = torch.cuda.current_stream(inp.device).cuda_stream
        for k in range(blocks):
            _cabi.check(lib.wfk_predict_linear(win.data_ptr(), wt.data_ptr(), bs.data_ptr(), b, ti, to, c, h * w,
                                               pred.data_ptr(), None, None, stream), "wfk_predict_linear")
            out[:, k * to:(k + 1) * to] = pred
            if k + 1 < blocks:      # next window = last `ti` frames of (window inputs, prediction)
                seq = torch.cat([win[:, :ti], pred], dim=1)
                win[:, :ti] = seq[:, -ti:]
        return out


class Autoencoder(nn.Module):
    """The train script's frozen-AE wrapper (train.py:21-56): ``encode([B,T,C,H,W]) -> [B,T,LC,h,w]``,
    ``decode([B,T,LC,h,w]) -> [B,T,1,H,W]``. The reference loops over T with batch B; frames are
    independent (GroupNorm is per sample), so they are processed ``frames_per_call`` at a time (default 37: every
    Path-B layer then has a tile count that is a multiple of the B200's 74 CTA pairs, i.e. whole waves; +0.9 %).

    ``posterior``: "sample" reproduces the reference's ``.sample()`` (train.py:39, device RNG, not
    reproducible CPU<->GPU); "mode" is the deterministic alternative the reference keeps
    commented (train.py:42) and what parity tests use; a ``noise`` tensor can be injected instead."""

    def __init__(self, config: dict, posterior: str = "sample", frames_per_call: int = 37):
        super().__init__()
        self.autoencoder = AutoencoderKL(**config)
        self.autoencoder.eval()
        self.scaling_factor = 0.18125  # train.py:26 (unused there too)
        self.posterior = posterior
        self.frames_per_call = int(frames_per_call)
        self.autoencoder.requires_grad_(False)

    @torch.no_grad()
    def encode(self, x: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        b, t, c, h, w = x.shape
        flat = x.reshape(b * t, c, h, w)
        outs = []
        for i in range(0, b * t, self.frames_per_call):
            post = self.autoencoder.encode(flat[i:i + self.frames_per_call])
            if noise is not None:
                nz = noise.reshape(b * t, *noise.shape[2:])[i:i + self.frames_per_call]
                outs.append(post.sample_with_noise(nz))
            elif self.posterior == "sample":
                outs.append(post.sample())
            else:
                outs.append(post.mode())
        z = torch.cat(outs, dim=0)
        return z.reshape(b, t, *z.shape[1:])

    @torch.no_grad()
    def decode(self, x: torch.Tensor) -> torch.Tensor:
        b, t, c, h, w = x.shape
        flat = x.reshape(b * t, c, h, w)
        ae = self.autoencoder
        nb = len(ae._cfg["block_out_channels"]) - 1
        y = torch.empty((b * t, ae._cfg["out_channels"], h << nb, w << nb), dtype=torch.float32, device=x.device)
        for i in range(0, b * t, self.frames_per_call):   # each chunk lands in its slice of the result
            ae.decode_into(flat[i:i + self.frames_per_call], y[i:i + self.frames_per_call])
        return y.reshape(b, t, *y.shape[1:])


class PathBNowcast(nn.Module):
    """``Model`` of the reference experiment, inference side (train.py:58-125)."""

    def __init__(self, autoencoder_cfg: dict, input_frames: int = INPUT_FRAMES, pred_frames: int = PRED_FRAMES,
                 posterior: str = "mode", frames_per_call: int = 37, predictor: Optional[nn.Module] = None):
        """``predictor``: any module with ``rollout(latents[B, t_in + t_out, C, h, w]) -> (pred, tgt, val_loss)`` --
        the default ``nn.Linear(13*4, 12*4)`` of ``pretrained_ae_linear_sevir`` or ``predictors.DLinear`` /
        ``predictors.DLinearIndcIndp`` of the ``pretrained_ae_dlinear_*`` experiments (same ``validation_step`` around
        ``self.predictor``, pretrained_ae_dlinear_sevir/train.py:179-192)."""
        super().__init__()
        self.autoencoder = Autoencoder(autoencoder_cfg, posterior=posterior, frames_per_call=frames_per_call)
        self.input_frames, self.pred_frames = input_frames, pred_frames
        if predictor is None:
            predictor = LatentLinearPredictor(input_frames, pred_frames, autoencoder_cfg.get("latent_channels", 4))
        elif not hasattr(predictor, "rollout"):
            raise TypeError("predictor must provide rollout(latents) -> (pred, tgt, val_loss)")
        self.predictor = predictor

    def forward(self, x):
        return self.predictor(x)

    @torch.no_grad()
    def validation_step(self, batch: torch.Tensor, noise: Optional[torch.Tensor] = None):
        """batch: [B, H, W, T] float32 in [0,1] (the reference loader's output) or uint8 (staged here).
        Returns (decoded_pred, decoded_tgt, val_loss) -- the tensors train.py:115-120 hands to log_metrics."""
        if batch.dtype == torch.uint8:
            v = stage_vil(batch)
        else:
            v = batch.permute(0, 3, 1, 2).unsqueeze(2).contiguous()
        lat = self.autoencoder.encode(v, noise=noise)
        pred, tgt, loss = self.predictor.rollout(lat)
        decoded_pred = self.autoencoder.decode(pred)
        decoded_tgt = self.autoencoder.decode(tgt)
        return decoded_pred, decoded_tgt, loss

    @torch.no_grad()
    def evaluate(self, batch: torch.Tensor, process_group=None, extended: bool = False) -> Dict[str, float]:
        """validation_step + log_metrics' calc_metrics (pipeline/helpers.py:142-153), optionally summed
        over the ranks of ``process_group`` with one all-reduce."""
        dp, dt, loss = self.validation_step(batch)
        res = wf_metrics.calc_metrics(dp, dt, extended=extended, process_group=process_group)
        res["val_loss"] = float(loss.item())
        return res


class LatentReconstruction(nn.Module):
    """``Model`` of the latent-compressor experiments, inference side: ``pretrained_ae_convae_sevir/train.py:145-198``
    (``predictors.ConvModel``, all frames of a sequence) and ``pretrained_ae_convattn_ae_sevir/train.py:167-222``
    (``predictors.ConvAttnModel``, one frame per sample). ``validation_step``: encode with the frozen AutoencoderKL,
    run the compressor on the latents, ``nn.HuberLoss()(pred, latents)``, decode the reconstruction; the caller scores
    it against the INPUT frames (``log_metrics(decoded_pred, inp, ...)``)."""

    def __init__(self, autoencoder_cfg: dict, predictor: nn.Module, posterior: str = "mode", frames_per_call: int = 37):
        super().__init__()
        self.autoencoder = Autoencoder(autoencoder_cfg, posterior=posterior, frames_per_call=frames_per_call)
        self.predictor = predictor

    @torch.no_grad()
    def forward(self, latents: torch.Tensor, return_loss: bool = False):
        """latents [B, T, C, 48, 48] -> reconstruction of the same shape (``Model.forward`` of both scripts)."""
        b, t = latents.shape[:2]
        from .predictors import ConvAttnModel
        if isinstance(self.predictor, ConvAttnModel):
            assert t == 1, "input should be 1 frame"       # pretrained_ae_convattn_ae_sevir/train.py:179
            out = self.predictor(latents[:, 0], return_loss=return_loss)
            rec = out[1].unsqueeze(1)
        else:
            out = self.predictor(latents, return_loss=return_loss)
            rec = out[1]
        return (rec, out[2]) if return_loss else rec

    @torch.no_grad()
    def validation_step(self, batch: torch.Tensor):
        """batch [B, H, W, T] uint8 or float in [0, 1] -> (decoded_pred, inp [B, T, 1, H, W], val_loss)."""
        inp = stage_vil(batch) if batch.dtype == torch.uint8 else batch.permute(0, 3, 1, 2).unsqueeze(2).contiguous()
        lat = self.autoencoder.encode(inp)
        rec, loss = self.forward(lat, return_loss=True)
        return self.autoencoder.decode(rec), inp, loss

    @torch.no_grad()
    def evaluate(self, batch: torch.Tensor, process_group=None, extended: bool = False) -> Dict[str, float]:
        dp, inp, loss = self.validation_step(batch)
        res = wf_metrics.calc_metrics(dp, inp, extended=extended, process_group=process_group)
        res["val_loss"] = float(loss.item())
        return res


def log_metrics(predictions, targets, tag, pl_module, process_group=None):
    """``pipeline.helpers.log_metrics`` (helpers.py:142-153), same signature plus ``process_group``."""
    if isinstance(predictions, torch.Tensor):
        predictions = predictions.detach()
    if isinstance(targets, torch.Tensor):
        targets = targets.detach()
    m = wf_metrics.calc_metrics(predictions, targets, process_group=process_group)
    m = {f"{tag}_{k}": v for k, v in m.items()}
    # counts were already summed over ranks, so the logged scalar needs no further reduction
    pl_module.log_dict(m, on_step=True, on_epoch=True, sync_dist=process_group is None)

"""GPU rendering of the validation panels behind ``pipeline.helpers.log_wandb_images`` (reference
pipeline/helpers.py:155-225) and ``vil_cmap`` (reference pipeline/datasets/sevir/sevir.py:1237-1268).

The reference moves both tensors to the host, quantises them with numpy and lets matplotlib colour-map every
frame. Here one kernel (``wfk_render_panels``) reads prediction and target once and writes the uint8 frames,
their absolute difference and the three RGBA images; only the finished uint8 mosaics of the first
``batch_idxs`` samples cross PCIe. Figure layout (axes, titles, colour bars) is matplotlib's business and is
not reproduced: a panel is the bare 3 x T mosaic (rows: original, reconstruction, abs diff).
"""
from __future__ import annotations

from copy import deepcopy
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _cabi

# reference pipeline/datasets/sevir/sevir.py:1237-1250
VIL_COLORS = [[0, 0, 0],
              [0.30196078431372547, 0.30196078431372547, 0.30196078431372547],
              [0.1568627450980392, 0.7450980392156863, 0.1568627450980392],
              [0.09803921568627451, 0.5882352941176471, 0.09803921568627451],
              [0.0392156862745098, 0.4117647058823529, 0.0392156862745098],
              [0.0392156862745098, 0.29411764705882354, 0.0392156862745098],
              [0.9607843137254902, 0.9607843137254902, 0.0],
              [0.9294117647058824, 0.6745098039215687, 0.0],
              [0.9411764705882353, 0.43137254901960786, 0.0],
              [0.6274509803921569, 0.0, 0.0],
              [0.9058823529411765, 0.0, 1.0]]
VIL_LEVELS = [0.0, 16.0, 31.0, 59.0, 74.0, 100.0, 133.0, 160.0, 181.0, 219.0, 255.0]

# ColorBrewer 9-class Reds as matplotlib spells the anchors of 'Reds' (helpers.py:207 `cmap='Reds'`); the literals
# matter: 0.05098039215686274 * 255 truncates to 12, 13 / 255 * 255 does not
_REDS = [(1.0, 0.96078431372549022, 0.94117647058823528),
         (0.99607843137254903, 0.8784313725490196, 0.82352941176470584),
         (0.9882352941176471, 0.73333333333333328, 0.63137254901960782),
         (0.9882352941176471, 0.5725490196078431, 0.44705882352941179),
         (0.98431372549019602, 0.41568627450980394, 0.29019607843137257),
         (0.93725490196078431, 0.23137254901960785, 0.17254901960784313),
         (0.79607843137254897, 0.094117647058823528, 0.11372549019607843),
         (0.6470588235294118, 0.058823529411764705, 0.08235294117647058),
         (0.40392156862745099, 0.0, 0.05098039215686274)]


class ByteColormap:
    """A colour map over uint8 data as the 256 x RGBA byte table ``cmap(norm(v), bytes=True)`` evaluates to."""

    def __init__(self, table_u8: np.ndarray, name: str):
        assert table_u8.shape == (256, 4) and table_u8.dtype == np.uint8
        self.table, self.name = table_u8, name
        self._dev: Dict[int, torch.Tensor] = {}

    def __call__(self, x_u8):
        return self.table[np.asarray(x_u8, dtype=np.uint8)]

    def device_table(self, device: torch.device) -> torch.Tensor:
        key = device.index or 0
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(self.table.copy()).to(device)
        return self._dev[key]


def _listed_boundary_table(colors, under, over, levels) -> np.ndarray:
    """ListedColormap(colors) with set_under / set_over + BoundaryNorm(levels, len(colors)) on values 0..255:
    bin = (number of boundaries <= v) - 1, v >= levels[-1] -> over, v < levels[0] -> under; bytes = trunc(c * 255)."""
    n = len(colors)
    lut = np.ones((n + 2, 4), dtype=np.float64)
    lut[:n, :3] = np.asarray(colors, dtype=np.float64)
    lut[n, :3], lut[n + 1, :3] = under, over
    lut_b = (lut * 255).astype(np.uint8)
    v = np.arange(256, dtype=np.float32)
    lev = np.asarray(levels)
    idx = np.searchsorted(lev, v, side="right") - 1
    idx = np.where(v < lev[0], n, np.where(v >= lev[-1], n + 1, idx))
    return lut_b[idx]


def _segmented_table(anchors, vmin: float, vmax: float, n: int = 256) -> np.ndarray:
    """LinearSegmentedColormap.from_list(anchors) (n-entry table, linear interpolation) behind Normalize(vmin, vmax)
    on values 0..255 (float32 arithmetic as matplotlib uses for uint8 input)."""
    anchors = np.asarray(anchors, dtype=np.float64)
    x = np.linspace(0, 1, len(anchors))
    xi = np.linspace(0, 1, n)
    seg = np.searchsorted(x, xi)[1:-1]
    frac = (xi[1:-1] - x[seg - 1]) / (x[seg] - x[seg - 1])
    lut = np.ones((n, 4), dtype=np.float64)
    for c in range(3):
        y = anchors[:, c]
        lut[1:-1, c] = frac * (y[seg] - y[seg - 1]) + y[seg - 1]
        lut[0, c], lut[-1, c] = y[0], y[-1]
    lut = np.clip(lut, 0, 1)
    lut_b = (lut * 255).astype(np.uint8)
    xa = (np.arange(256, dtype=np.float32) - np.float32(vmin)) / np.float32(vmax - vmin) * np.float32(n)
    xa[xa == n] = n - 1
    idx = np.clip(xa.astype(int), 0, n - 1)
    return lut_b[idx]


_CMAPS: Dict[str, ByteColormap] = {}


def vil_cmap(encoded: bool = True):
    """Same return shape as the reference's ``vil_cmap`` (sevir.py:1252-1268): ``(cmap, norm, vmin, vmax)``. ``cmap`` is
    the byte table over uint8 VIL values with the boundary norm already folded in, so ``norm`` is None."""
    if "vil" not in _CMAPS:
        cols = deepcopy(VIL_COLORS)
        cols.pop(0)  # `nil` (masked values): never produced by uint8 frames
        _CMAPS["vil"] = ByteColormap(_listed_boundary_table(cols, cols[0], cols[-1], VIL_LEVELS), "vil")
    return _CMAPS["vil"], None, None, None


def diff_cmap() -> ByteColormap:
    """``imshow(diff, cmap='Reds', vmin=0, vmax=255)`` (helpers.py:207)."""
    if "reds" not in _CMAPS:
        _CMAPS["reds"] = ByteColormap(_segmented_table(_REDS, 0.0, 255.0), "Reds")
    return _CMAPS["reds"]


@torch.no_grad()
def render_panels(predicted: torch.Tensor, target: torch.Tensor, rgba: bool = True) -> Dict[str, torch.Tensor]:
    """predicted / target: CUDA fp32 tensors of one shape (any rank). Returns uint8 tensors ``target_u8``, ``pred_u8``,
    ``diff_u8`` (input shape) and, with ``rgba``, ``target_rgba``, ``pred_rgba``, ``diff_rgba`` (input shape + [4])."""
    if not (predicted.is_cuda and target.is_cuda):
        raise RuntimeError("render_panels runs on a B200 only (no CPU fallback): pass CUDA tensors")
    if predicted.shape != target.shape:
        raise ValueError(f"shape mismatch: {tuple(predicted.shape)} vs {tuple(target.shape)}")
    lib = _cabi.init(predicted.device.index or 0)
    p = predicted.detach().to(torch.float32).contiguous()
    t = target.detach().to(torch.float32).contiguous()
    dev, shape = p.device, tuple(p.shape)
    out = {k: torch.empty(shape, dtype=torch.uint8, device=dev) for k in ("target_u8", "pred_u8", "diff_u8")}
    if rgba:
        out.update({k: torch.empty(shape + (4,), dtype=torch.uint8, device=dev)
                    for k in ("target_rgba", "pred_rgba", "diff_rgba")})
    ptr = lambda k: out[k].data_ptr() if k in out else None  # noqa: E731
    stream = torch.cuda.current_stream(dev).cuda_stream
    _cabi.check(lib.wfk_render_panels(p.data_ptr(), t.data_ptr(), p.numel(), vil_cmap()[0].device_table(dev).data_ptr(),
                                      diff_cmap().device_table(dev).data_ptr(), ptr("target_u8"), ptr("pred_u8"),
                                      ptr("diff_u8"), ptr("target_rgba"), ptr("pred_rgba"), ptr("diff_rgba"), stream),
                "wfk_render_panels")
    return out


@torch.no_grad()
def panel_mosaics(predicted: torch.Tensor, target: torch.Tensor, batch_idxs: int = 4) -> List[np.ndarray]:
    """The image content of ``log_wandb_images``: for each of the first ``batch_idxs`` samples one uint8
    [3*H, T*W, 4] RGBA mosaic (rows: original, reconstruction, abs diff; columns: time)."""
    if predicted.ndim == 5:
        assert predicted.shape[2] == 1, "Predicted must be (B,T,1,H,W)"
        predicted = predicted.squeeze(2)
    if target.ndim == 5:
        assert target.shape[2] == 1, "Target must be (B,T,1,H,W)"
        target = target.squeeze(2)
    b = min(batch_idxs, predicted.shape[0])
    r = render_panels(predicted[:b], target[:b])
    rows = torch.stack([r["target_rgba"], r["pred_rgba"], r["diff_rgba"]], dim=1)  # [b, 3, T, H, W, 4]
    _, _, t, h, w, _ = rows.shape
    mos = rows.permute(0, 1, 3, 2, 4, 5).reshape(b, 3 * h, t * w, 4)
    host = mos.cpu().numpy()
    return [host[i] for i in range(b)]


def log_wandb_images(predicted, target, label, pl_module, batch_idxs: int = 4) -> Optional[List[np.ndarray]]:
    """``pipeline.helpers.log_wandb_images`` (helpers.py:155-225), same signature. Logs one image per sample to a
    W&B logger when ``wandb`` is importable and the module's logger exposes ``experiment.log``; always returns the
    mosaics so other loggers can take them."""
    in_range = int(((target >= 0) & (target <= 1)).sum().item())
    ratio = in_range / max(1, target.numel())
    if ratio < 0.9:
        print(f"\033[91mtarget data not in [0,1] range: {ratio:.2%}\033[0m")
    mosaics = panel_mosaics(predicted, target, batch_idxs)
    logger = getattr(pl_module, "logger", None)
    experiment = getattr(logger, "experiment", None)
    if experiment is not None and hasattr(experiment, "log"):
        try:
            import wandb  # type: ignore
        except ImportError:
            wandb = None
        for b, m in enumerate(mosaics):
            img = wandb.Image(m, caption=f"{label} (batch {b})") if wandb is not None else m
            experiment.log({f"{label}": img, "global_step": getattr(pl_module, "global_step", 0)})
    return mosaics

"""TEST INFRASTRUCTURE ONLY (imported by tests/ and __graft_entry__.smoke(), never by the product path).

CPU restatement (numpy) of the numeric part of ``log_wandb_images`` (reference pipeline/helpers.py:155-225):
uint8 quantisation of target / prediction, their absolute difference, and the two colour mappings ``imshow``
applies -- ``vil_cmap()`` (reference pipeline/datasets/sevir/sevir.py:1237-1268: VIL_COLORS, VIL_LEVELS,
ListedColormap + BoundaryNorm) and ``'Reds'`` with ``vmin=0, vmax=255`` (helpers.py:207).

PARITY UNPINNED for the colour mapping: matplotlib is a third-party dependency of the reference (version not
pinned: no requirements file, setup.py:3-9 lists no deps) and is not installed in this image, so the maps
below restate matplotlib 3.x's published algorithm --
  * ``BoundaryNorm.__call__``: ``iret = np.digitize(x, boundaries) - 1``; ``x < vmin -> -1``;
    ``x >= vmax -> ncolors`` (no ``clip``); no bin rescaling because ncolors == number of bins (10);
  * ``Colormap.__call__`` on integer indices: ``< 0 -> under``, ``> N-1 -> over``; ``bytes=True`` takes
    ``(lut * 255).astype(np.uint8)`` (float64, truncation);
  * ``Normalize(0, 255)`` on uint8 data computes in float32 (``process_value`` promotes uint8 to float32);
    ``Colormap.__call__`` on floats: ``xa *= N; xa[xa == N] = N - 1; xa.astype(int)``;
  * ``'Reds'`` = ``LinearSegmentedColormap.from_list`` over the nine ColorBrewer anchors, 256-entry table built by
    ``_create_lookup_table`` (linear interpolation at ``linspace(0, 1, 256)``).
The quantisation / difference part is plain numpy in the reference and is restated exactly.
"""
from copy import deepcopy

import numpy as np

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

# matplotlib _cm.py `_Reds_data` (ColorBrewer 9-class Reds)
REDS_ANCHORS = [(1.0, 0.96078431372549022, 0.94117647058823528),
                (0.99607843137254903, 0.8784313725490196, 0.82352941176470584),
                (0.9882352941176471, 0.73333333333333328, 0.63137254901960782),
                (0.9882352941176471, 0.5725490196078431, 0.44705882352941179),
                (0.98431372549019602, 0.41568627450980394, 0.29019607843137257),
                (0.93725490196078431, 0.23137254901960785, 0.17254901960784313),
                (0.79607843137254897, 0.094117647058823528, 0.11372549019607843),
                (0.6470588235294118, 0.058823529411764705, 0.08235294117647058),
                (0.40392156862745099, 0.0, 0.05098039215686274)]


def quantise(x: np.ndarray) -> np.ndarray:
    """helpers.py:181-182: ``(x.clamp(0,1) * 255).numpy().astype('uint8')`` (float32 product, truncation; NaN -> 0)."""
    x = np.asarray(x, dtype=np.float32)
    v = np.clip(x, np.float32(0), np.float32(1)) * np.float32(255)
    v = np.where(np.isnan(v), np.float32(0), v)
    return v.astype(np.uint8)


def abs_diff(t_u8: np.ndarray, p_u8: np.ndarray) -> np.ndarray:
    """helpers.py:183."""
    return np.abs(t_u8.astype(float) - p_u8.astype(float)).clip(0, 255).astype(np.uint8)


def vil_rgba(x_u8: np.ndarray) -> np.ndarray:
    """``cmap(norm(x), bytes=True)`` with ``cmap, norm, _, _ = vil_cmap()`` (sevir.py:1252-1268), value by value."""
    cols = deepcopy(VIL_COLORS)
    lev = np.asarray(deepcopy(VIL_LEVELS))
    cols.pop(0)                      # `nil`, used for masked values only
    under, over = cols[0], cols[-1]  # sevir.py:1259-1261
    n = len(cols)                    # ListedColormap.N == 10 == number of bins
    xx = np.asarray(x_u8).astype(np.float32)
    iret = np.digitize(xx, lev) - 1
    iret[xx < lev[0]] = -1
    iret[xx >= lev[-1]] = n
    lut = np.ones((n + 2, 4), dtype=np.float64)
    lut[:n, :3] = np.asarray(cols, dtype=np.float64)
    lut[n, :3] = under
    lut[n + 1, :3] = over
    lut_b = (lut * 255).astype(np.uint8)
    idx = np.where(iret < 0, n, np.where(iret > n - 1, n + 1, iret))
    return lut_b[idx]


def reds_lut(n: int = 256) -> np.ndarray:
    """float64 [n, 4] lookup table of matplotlib's 'Reds' (from_list + _create_lookup_table)."""
    anchors = np.asarray(REDS_ANCHORS, dtype=np.float64)
    x = np.linspace(0, 1, len(anchors))
    xind = np.linspace(0, 1, n)
    ind = np.searchsorted(x, xind)[1:-1]
    dist = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
    lut = np.ones((n, 4), dtype=np.float64)
    for c in range(3):
        y = anchors[:, c]
        lut[:, c] = np.clip(np.concatenate([[y[0]], dist * (y[ind] - y[ind - 1]) + y[ind - 1], [y[-1]]]), 0, 1)
    return lut


def reds_rgba(d_u8: np.ndarray) -> np.ndarray:
    """``cm.Reds(Normalize(0, 255)(d), bytes=True)``."""
    n = 256
    xa = np.asarray(d_u8).astype(np.float32)
    xa = xa / np.float32(255)
    xa = xa * np.float32(n)
    xa[xa == n] = n - 1
    idx = xa.astype(int)
    lut_b = (reds_lut(n) * 255).astype(np.uint8)
    return lut_b[idx]


def render_panels(pred: np.ndarray, tgt: np.ndarray) -> dict:
    t8, p8 = quantise(tgt), quantise(pred)
    d8 = abs_diff(t8, p8)
    return {"target_u8": t8, "pred_u8": p8, "diff_u8": d8, "target_rgba": vil_rgba(t8), "pred_rgba": vil_rgba(p8),
            "diff_rgba": reds_rgba(d8)}

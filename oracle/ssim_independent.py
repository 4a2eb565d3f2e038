"""ORACLE (test infrastructure, not product code): a SECOND, independent statement of the SSIM / PSNR the reference
obtains from ``torchmetrics`` (``pipeline/metrics.py:71-84`` in /root/reference), used to pin ``metrics_oracle._tm_ssim``
/ ``_tm_psnr`` and the CUDA kernel against something that shares no code with either.

``metrics_oracle`` restates torchmetrics' implementation strategy (float32 ``F.conv2d`` with the outer-product 11 x 11
kernel on reflect-padded images, then cropping the padding away). This file instead follows the DEFINITION (Wang et
al. 2004, the form torchmetrics documents): float64 throughout, two separable ``scipy.ndimage.correlate1d`` passes with
the sigma = 1.5, 11-tap Gaussian, no padding at all -- only window centres whose 11 x 11 support lies inside the image
are evaluated -- and local moments E[x], E[y], E[xx], E[yy], E[xy].

torchmetrics itself is not installed here (un-vendored, un-pinned dependency, no network), so neither statement can be
run against the library: what this pins is that two independent derivations of the published algorithm agree.
``variance_clamp``: "each" = torchmetrics >= 1.x (each variance clamped at 0), "none" = older releases, "sum" = the
CUDA kernel (sigma_p^2 + sigma_t^2 is only ever used as a sum); the three differ by rounding-level amounts."""
from __future__ import annotations

import numpy as np
from scipy.ndimage import correlate1d


def gaussian_window(size: int = 11, sigma: float = 1.5) -> np.ndarray:
    x = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    g = np.exp(-0.5 * (x / sigma) ** 2)
    return g / g.sum()


def _local_mean_valid(img: np.ndarray, g: np.ndarray) -> np.ndarray:
    """Gaussian-weighted local mean at the window centres whose support is inside the image."""
    r = (len(g) - 1) // 2
    out = correlate1d(img, g, axis=-1, mode="constant", cval=np.nan)
    out = correlate1d(out, g, axis=-2, mode="constant", cval=np.nan)
    return out[..., r:-r, r:-r]


def ssim_per_image(pred, target, data_range: float = 1.0, k1: float = 0.01, k2: float = 0.03,
                   variance_clamp: str = "each") -> np.ndarray:
    """pred, target: [..., H, W] array-likes -> SSIM per leading index (mean over the valid centres), float64."""
    p = np.asarray(pred, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    g = gaussian_window()
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    mu_p, mu_t = _local_mean_valid(p, g), _local_mean_valid(t, g)
    e_pp, e_tt, e_pt = _local_mean_valid(p * p, g), _local_mean_valid(t * t, g), _local_mean_valid(p * t, g)
    var_p, var_t, cov = e_pp - mu_p ** 2, e_tt - mu_t ** 2, e_pt - mu_p * mu_t
    if variance_clamp == "each":
        var_sum = np.maximum(var_p, 0.0) + np.maximum(var_t, 0.0)
    elif variance_clamp == "sum":
        var_sum = np.maximum(var_p + var_t, 0.0)
    elif variance_clamp == "none":
        var_sum = var_p + var_t
    else:
        raise ValueError(variance_clamp)
    s = ((2 * mu_p * mu_t + c1) * (2 * cov + c2)) / ((mu_p ** 2 + mu_t ** 2 + c1) * (var_sum + c2))
    return s.reshape(*s.shape[:-2], -1).mean(-1)


def ssim(pred, target, **kw) -> float:
    """StructuralSimilarityIndexMeasure(data_range=1.0): mean over the images of a [..., 1, H, W] batch."""
    p = np.asarray(pred, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    per = ssim_per_image(p.reshape(-1, *p.shape[-2:]), t.reshape(-1, *t.shape[-2:]), **kw)
    return float(per.mean())


def psnr_per_frame_mean(pred, target) -> float:
    """``pipeline/metrics.py:77-84``: a fresh PeakSignalNoiseRatio() per frame (data_range=None: the metric tracks
    min(target.min(), 0) and max(target.max(), 0), so range = max(t.max(), 0) - min(t.min(), 0)), base-10 log of
    range^2 / MSE, averaged over the b*t frames."""
    p = np.asarray(pred, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    p = p.reshape(-1, p.shape[-2] * p.shape[-1])
    t = t.reshape(-1, t.shape[-2] * t.shape[-1])
    mse = ((p - t) ** 2).mean(-1)
    rng = np.maximum(t.max(-1), 0.0) - np.minimum(t.min(-1), 0.0)
    with np.errstate(divide="ignore"):
        return float((10.0 * np.log10(rng ** 2 / mse)).mean())

"""ORACLE (test infrastructure, not product code): CPU restatement of the reference skill scores
``pipeline/metrics.py`` (paths relative to ``/root/reference``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this file.

Two layers:

* ``calc_metrics`` & friends -- the reference's own torch-CPU arithmetic, function by function
  (``_hit_miss_fa_cn`` :9-16, ``crps`` :18-41, ``csi`` :43-54, ``hss`` :56-69, ``ssim`` :71-75,
  ``psnr`` :77-84, ``calc_metrics`` :86-133). Pinned in the build container against the UNMODIFIED
  ``pipeline.metrics`` module (``tests/test_oracle_vs_reference.py``).
* ``integer_counts`` / ``partials`` -- the exact-integer statement of the same counts (SURVEY hazard
  H1: the reference sums 0/1 float32 tensors, which is only exact below 2**24), the quantity the
  CUDA kernel must reproduce bit-for-bit.

PARITY UNPINNED for SSIM and PSNR: the reference delegates them to ``torchmetrics``
(``StructuralSimilarityIndexMeasure(data_range=1.0)``, ``PeakSignalNoiseRatio()``), which is an
un-vendored, un-pinned dependency that is not installed here and cannot be fetched. ``_tm_ssim`` /
``_tm_psnr`` restate the published torchmetrics (>=0.11 / 1.x) functional algorithm
(``torchmetrics/functional/image/ssim.py::_ssim_update`` with gaussian_kernel=True, sigma=1.5,
kernel_size=11, k1=0.01, k2=0.03; ``psnr.py::_psnr_compute`` with data_range=None -> tracked
target min/max against 0). The reference holds no golden value for either.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch
import torch.nn.functional as F

_eps = 1e-8
THRESHOLDS = [16 / 255, 74 / 255, 133 / 255, 160 / 255, 181 / 255, 219 / 255]  # metrics.py:107
POOLS = (1, 4, 16)


# ------------------------------------------------------------------------------ torchmetrics restated
def _gaussian(kernel_size: int, sigma: float, dtype) -> torch.Tensor:
    dist = torch.arange(start=(1 - kernel_size) / 2, end=(1 + kernel_size) / 2, step=1, dtype=dtype)
    gauss = torch.exp(-torch.pow(dist / sigma, 2) / 2)
    return (gauss / gauss.sum()).unsqueeze(dim=0)


def _tm_ssim_per_image(preds: torch.Tensor, target: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """torchmetrics ``_ssim_update`` -> one SSIM value per image of a [N, C, H, W] batch."""
    sigma, k1, k2 = 1.5, 0.01, 0.03
    ks = int(3.5 * sigma + 0.5) * 2 + 1
    pad = (ks - 1) // 2
    c = preds.size(1)
    dtype = preds.dtype
    target = target.to(dtype)
    c1 = pow(k1 * data_range, 2)
    c2 = pow(k2 * data_range, 2)
    g = _gaussian(ks, sigma, dtype)
    kernel = torch.matmul(g.t(), g).expand(c, 1, ks, ks)
    preds = F.pad(preds, (pad, pad, pad, pad), mode="reflect")
    target = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    inp = torch.cat((preds, target, preds * preds, target * target, preds * target))
    out = F.conv2d(inp, kernel, groups=c)
    o = out.split(preds.shape[0])
    mu_p_sq, mu_t_sq, mu_pt = o[0].pow(2), o[1].pow(2), o[0] * o[1]
    sig_p = torch.clamp(o[2] - mu_p_sq, min=0.0)
    sig_t = torch.clamp(o[3] - mu_t_sq, min=0.0)
    sig_pt = o[4] - mu_pt
    upper = 2 * sig_pt + c2
    lower = sig_p + sig_t + c2
    full = ((2 * mu_pt + c1) * upper) / ((mu_p_sq + mu_t_sq + c1) * lower)
    full = full[..., pad:-pad, pad:-pad]
    return full.reshape(full.shape[0], -1).mean(-1)


def _tm_ssim(preds, target) -> torch.Tensor:
    """``StructuralSimilarityIndexMeasure(data_range=1.0)(p, g)``: mean over images."""
    per = _tm_ssim_per_image(preds, target, 1.0)
    return per.sum() / per.shape[0]


def _tm_psnr(preds, target) -> torch.Tensor:
    """``PeakSignalNoiseRatio()(p, g)`` on a fresh metric: data_range = max(target.max(), 0) -
    min(target.min(), 0); 10*log10(range^2 / mse) evaluated as torchmetrics does (natural logs)."""
    sse = torch.sum(torch.pow(preds - target, 2))
    n = torch.tensor(target.numel())
    mx = torch.maximum(target.max(), torch.tensor(0.0))
    mn = torch.minimum(target.min(), torch.tensor(0.0))
    data_range = mx - mn
    base_e = 2 * torch.log(data_range) - torch.log(sse / n)
    return base_e * (10 / torch.log(torch.tensor(10.0)))


# ------------------------------------------------------------------------------ pipeline/metrics.py
def _hit_miss_fa_cn(pred, target, threshold):
    """metrics.py:9-16."""
    p = (pred >= threshold).float()
    t = (target >= threshold).float()
    tp = torch.sum(p * t)
    fn = torch.sum((1 - p) * t)
    fp = torch.sum(p * (1 - t))
    tn = torch.sum((1 - p) * (1 - t))
    return tp, fn, fp, tn


def _pool(x5, pool_type, scale):
    b = x5.shape[0]
    fn = F.avg_pool2d if pool_type == "avg" else F.max_pool2d
    x = x5.reshape(-1, *x5.shape[2:])
    x = fn(x, scale, stride=scale)
    return x.reshape(b, -1, *x.shape[1:])


def crps(pred, target, pool_type="none", scale=1):
    """metrics.py:18-41 (one member: pred.ndim == 5; ensemble: (b, n, t, c, h, w))."""
    normal = torch.distributions.Normal(0, 1)
    frac_sqrt_pi = 1 / np.sqrt(np.pi)
    eps = 1e-10
    if pred.ndim == 5:
        pred = pred.unsqueeze(1)
    b, n, t, c, h, w = pred.shape
    gt = target.reshape(b * t, c, h, w)
    pr = pred.reshape(b * n * t, c, h, w)
    if pool_type == "avg":
        pr = F.avg_pool2d(pr, scale, stride=scale)
        gt = F.avg_pool2d(gt, scale, stride=scale)
    elif pool_type == "max":
        pr = F.max_pool2d(pr, scale, stride=scale)
        gt = F.max_pool2d(gt, scale, stride=scale)
    gt = gt.reshape(b, t, *gt.shape[1:])
    pr = pr.reshape(b, n, t, *pr.shape[1:])
    mean = torch.mean(pr, dim=1)
    std = torch.std(pr, dim=1) if n > 1 else torch.zeros_like(mean)
    normed = (mean - gt + eps) / (std + eps)
    cdf = normal.cdf(normed)
    pdf = normal.log_prob(normed).exp()
    val = (std + eps) * (normed * (2 * cdf - 1) + 2 * pdf - frac_sqrt_pi)
    return float(torch.mean(val).item())


def csi(pred, target, threshold, pool_type="none", scale=1):
    """metrics.py:43-54."""
    if pool_type in ("avg", "max"):
        pred, target = _pool(pred, pool_type, scale), _pool(target, pool_type, scale)
    tp, fn, fp, _ = _hit_miss_fa_cn(pred, target, threshold)
    return float((tp / (tp + fn + fp + _eps)).item())


def hss(pred, target, threshold, pool_type="none", scale=1):
    """metrics.py:56-69."""
    if pool_type in ("avg", "max"):
        pred, target = _pool(pred, pool_type, scale), _pool(target, pool_type, scale)
    tp, fn_, fp, tn = _hit_miss_fa_cn(pred, target, threshold)
    num = 2 * (tp * tn - fn_ * fp)
    den = (tp + fn_) * (fn_ + tn) + (tp + fp) * (fp + tn) + _eps
    return float((num / den).item())


def ssim(pred, target):
    """metrics.py:71-75."""
    p = pred.reshape(-1, *pred.shape[2:])
    g = target.reshape(-1, *target.shape[2:])
    return float(_tm_ssim(p, g).item())


def psnr(pred, target):
    """metrics.py:77-84 (fresh-metric semantics per frame)."""
    p = pred.reshape(-1, *pred.shape[2:])
    g = target.reshape(-1, *target.shape[2:])
    total = 0.0
    for i in range(p.shape[0]):
        total += _tm_psnr(p[i:i + 1], g[i:i + 1]).item()
    return float(total / p.shape[0])


def calc_metrics(pred, target) -> Dict[str, float]:
    """metrics.py:86-133."""
    pred = pred.detach().clamp(0, 1)
    target = target.detach().clamp(0, 1)
    single = pred.mean(dim=1) if pred.ndim == 6 else pred
    results = {}
    results["CRPS"] = crps(pred, target, "none", 1)
    results["CRPS_4"] = crps(pred, target, "avg", 4)
    results["CRPS_16"] = crps(pred, target, "avg", 16)
    results["SSIM"] = ssim(single, target)
    results["PSNR"] = psnr(single, target)
    for i, th in enumerate(THRESHOLDS):
        results[f"CSI_{i}"] = csi(single, target, th, "none", 1)
        results[f"CSI_{i}_4"] = csi(single, target, th, "avg", 4)
        results[f"CSI_{i}_16"] = csi(single, target, th, "avg", 16)
        results[f"HSS_{i}"] = hss(single, target, th, "none", 1)
        results[f"HSS_{i}_4"] = hss(single, target, th, "avg", 4)
        results[f"HSS_{i}_16"] = hss(single, target, th, "avg", 16)
    results["paper_SSIM"] = results["SSIM"]
    results["paper_PSNR"] = results["PSNR"]
    results["paper_CRPS"] = results["CRPS"]
    for pool_name, suffix in [("POOL1", ""), ("POOL4", "_4"), ("POOL16", "_16")]:
        csi_vals = [results[f"CSI_{i}{suffix}"] for i in range(6)]
        hss_vals = [results[f"HSS_{i}{suffix}"] for i in range(6)]
        results[f"paper_CSI_M_{pool_name}"] = float(np.mean(csi_vals))
        results[f"paper_CSI_181_{pool_name}"] = results[f"CSI_4{suffix}"]
        results[f"paper_CSI_219_{pool_name}"] = results[f"CSI_5{suffix}"]
        results[f"paper_HSS_{pool_name}"] = float(np.mean(hss_vals))
    return results


# ------------------------------------------------------------------------------ exact statement
def integer_counts(pred, target, thresholds: Sequence[float] = THRESHOLDS, clamp: bool = True) -> np.ndarray:
    """Exact int64 [pool][threshold][tp, fn, fp, tn] of ``_hit_miss_fa_cn`` (metrics.py:9-16):
    compares in float32 against the float32-rounded threshold (hazard H2), pools with
    ``F.avg_pool2d`` exactly as ``csi`` does (metrics.py:46-50), sums as integers (hazard H1)."""
    if clamp:
        pred, target = pred.clamp(0, 1), target.clamp(0, 1)
    p4 = pred.reshape(-1, 1, *pred.shape[-2:]).float()
    t4 = target.reshape(-1, 1, *target.shape[-2:]).float()
    out = np.zeros((len(POOLS), len(thresholds), 4), dtype=np.int64)
    for pi, k in enumerate(POOLS):
        pp = p4 if k == 1 else F.avg_pool2d(p4, k, stride=k)
        tt = t4 if k == 1 else F.avg_pool2d(t4, k, stride=k)
        for ti, th in enumerate(thresholds):
            pb = pp >= th
            tb = tt >= th
            out[pi, ti, 0] = int((pb & tb).sum())
            out[pi, ti, 1] = int((~pb & tb).sum())
            out[pi, ti, 2] = int((pb & ~tb).sum())
            out[pi, ti, 3] = int((~pb & ~tb).sum())
    return out


def partials(pred, target, thresholds: Sequence[float] = THRESHOLDS, clamp: bool = True) -> dict:
    """Everything the fused kernel accumulates, computed the slow way in float64 / int64."""
    if clamp:
        pred, target = pred.clamp(0, 1), target.clamp(0, 1)
    p4 = pred.reshape(-1, 1, *pred.shape[-2:]).float()
    t4 = target.reshape(-1, 1, *target.shape[-2:]).float()
    res = {"counts": integer_counts(pred, target, thresholds, clamp=False), "n_frames": p4.shape[0]}
    abs_sum, n_elems = [], []
    for k in POOLS:
        pp = p4 if k == 1 else F.avg_pool2d(p4, k, stride=k)
        tt = t4 if k == 1 else F.avg_pool2d(t4, k, stride=k)
        abs_sum.append(float((pp.double() - tt.double()).abs().sum()))
        n_elems.append(pp.numel())
    res["abs_sum"] = abs_sum
    res["n_elems"] = n_elems
    res["sq_sum"] = float(((p4.double() - t4.double()) ** 2).sum())
    res["ssim_sum"] = float(_tm_ssim_per_image(p4.double(), t4.double()).sum())
    ps = 0.0
    for i in range(p4.shape[0]):
        ps += float(_tm_psnr(p4[i:i + 1].double(), t4[i:i + 1].double()))
    res["psnr_sum"] = ps
    return res

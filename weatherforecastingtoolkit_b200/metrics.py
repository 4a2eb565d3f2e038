"""Drop-in for the reference ``pipeline/metrics.py`` on top of the fused sm_100a skill-score kernel.

Same names, arities and return types as the reference module (paths relative to the reference repo
root): ``_hit_miss_fa_cn`` (pipeline/metrics.py:9-16), ``crps`` (:18-41), ``csi`` (:43-54), ``hss``
(:56-69), ``ssim`` (:71-75), ``psnr`` (:77-84), ``calc_metrics`` (:86-133, the same 56 keys).
Every function makes ONE pass over (pred, target) on the GPU (``wfk_metrics``) and one D2H copy of
an 864-byte partials struct, instead of the reference's 41 passes and 41 + B*T ``.item()`` syncs.

Additions: ``metric_partials`` / ``scores_from_partials`` expose the exact int64
contingency counts and float64 partial sums; ``calc_metrics(..., extended=True)`` adds POD / FAR /
MSE / MAE; ``process_group`` sums the partials over ranks with a single all-reduce so every rank
reports the score of the GLOBAL batch (sum of counts, not the reference's mean of per-rank ratios).

Numerics: counts are exact integers (the reference's float32 sums round above 2**24, hazard H1);
ratios are then formed with the reference's float32 operation order, so CSI / HSS are bit-identical
to the reference whenever its own counts are exact. There is no CPU path: inputs must be CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from . import engine as _engine

_eps = 1e-8
THRESHOLDS = [16 / 255, 74 / 255, 133 / 255, 160 / 255, 181 / 255, 219 / 255]  # metrics.py:107
_POOL_INDEX = {1: 0, 4: 1, 16: 2}
_N_INT = 3 * 8 * 4 + 3 + 1   # int64 words of wfk_metric_partials
_N_F64 = 3 + 1 + 1 + 1 + 2   # float64 words


class MetricPartials:
    """Host copy of ``wfk_metric_partials`` (include/wfk_b200.h): additive over frames and ranks."""

    def __init__(self, ints: np.ndarray, floats: np.ndarray, n_thresholds: int):
        self.ints = ints.astype(np.int64, copy=True)
        self.floats = floats.astype(np.float64, copy=True)
        self.n_thresholds = int(n_thresholds)

    @property
    def counts(self) -> np.ndarray:  # [pool][threshold][tp, fn, fp, tn]
        return self.ints[:96].reshape(3, 8, 4)[:, : self.n_thresholds]

    @property
    def n_elems(self) -> np.ndarray:
        return self.ints[96:99]

    @property
    def n_frames(self) -> int:
        return int(self.ints[99])

    @property
    def abs_sum(self) -> np.ndarray:
        return self.floats[0:3]

    @property
    def sq_sum(self) -> float:
        return float(self.floats[3])

    @property
    def ssim_sum(self) -> float:
        return float(self.floats[4])

    @property
    def psnr_sum(self) -> float:
        return float(self.floats[5])

    def __add__(self, other: "MetricPartials") -> "MetricPartials":
        assert self.n_thresholds == other.n_thresholds
        return MetricPartials(self.ints + other.ints, self.floats + other.floats, self.n_thresholds)

    def as_f64_vector(self) -> np.ndarray:
        """One float64 vector (counts < 2**53 are exact) -- the payload of the single all-reduce."""
        return np.concatenate([self.ints.astype(np.float64), self.floats])

    @staticmethod
    def from_f64_vector(v: np.ndarray, n_thresholds: int) -> "MetricPartials":
        return MetricPartials(np.rint(v[:_N_INT]).astype(np.int64), v[_N_INT:_N_INT + _N_F64], n_thresholds)


_workspaces: Dict = {}


def _as_frames(x: torch.Tensor) -> torch.Tensor:
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError("weatherforecastingtoolkit_b200.metrics needs CUDA tensors (no CPU fallback)")
    if x.ndim == 6:
        raise NotImplementedError("ensemble inputs (ndim 6) are outside the Path-B scoring path")
    if x.ndim not in (4, 5) or x.shape[-3] != 1:
        raise ValueError(f"expected (b, t, 1, h, w) or (n, 1, h, w), got {tuple(x.shape)}")
    return x.detach().to(torch.float32).contiguous().view(-1, x.shape[-2], x.shape[-1])


def metric_partials_device(pred: torch.Tensor, target: torch.Tensor, thresholds: Sequence[float] = THRESHOLDS,
                           clamp: bool = True) -> torch.Tensor:
    """Launch the fused pass; returns the device-resident struct as an int64[108] tensor (no sync)."""
    p, t = _as_frames(pred), _as_frames(target)
    if p.shape != t.shape:
        raise ValueError("pred and target shapes differ")
    dev = p.device
    lib = _cabi.init(dev.index if dev.index is not None else 0)
    frames, h, w = p.shape
    thr = (C.c_float * len(thresholds))(*[float(np.float32(th)) for th in thresholds])
    out = torch.empty(_N_INT + _N_F64, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    done = 0
    acc = None
    while done < frames:  # 65535 frames per launch (grid.z limit)
        nf = min(frames - done, 65535)
        ws_bytes = lib.wfk_metrics_workspace_bytes(nf, h, w)
        key = (str(dev), ws_bytes)
        ws = _workspaces.get(key)
        if ws is None:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _workspaces.clear()
            _workspaces[key] = ws
        part = out if done == 0 else torch.empty_like(out)
        with _engine.timed_pass("metrics", 8.0 * nf * h * w):   # algorithmic traffic: pred + target read once
            _cabi.check(lib.wfk_metrics(p[done:].data_ptr(), t[done:].data_ptr(), nf, h, w, thr, len(thresholds),
                                        1 if clamp else 0, part.data_ptr(), ws.data_ptr(), ws_bytes, stream), "wfk_metrics")
        if done:
            acc = _add_device(out, part)
            out = acc
        done += nf
    return out


def _add_device(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    r = a.clone()
    r[:_N_INT] += b[:_N_INT]
    r[_N_INT:] = (a[_N_INT:].view(torch.float64) + b[_N_INT:].view(torch.float64)).view(torch.int64)
    return r


def _to_host(dev_struct: torch.Tensor, n_thresholds: int) -> MetricPartials:
    host = dev_struct.cpu().numpy()
    return MetricPartials(host[:_N_INT], host[_N_INT:].view(np.float64), n_thresholds)


def all_reduce_partials(dev_struct: torch.Tensor, process_group=None) -> torch.Tensor:
    """Sum the partials over the ranks of ``process_group`` with ONE all-reduce of a float64[108]
    vector (int64 counts below 2**53 survive the round trip exactly). Works with NCCL (device
    tensors) and gloo (host tensors)."""
    import torch.distributed as dist

    v = torch.cat([dev_struct[:_N_INT].to(torch.float64), dev_struct[_N_INT:].view(torch.float64)])
    dist.all_reduce(v, op=dist.ReduceOp.SUM, group=process_group)
    return torch.cat([v[:_N_INT].round().to(torch.int64), v[_N_INT:].view(torch.int64)])


def metric_partials(pred, target, thresholds: Sequence[float] = THRESHOLDS, clamp: bool = True,
                    process_group=None) -> MetricPartials:
    dev_struct = metric_partials_device(pred, target, thresholds, clamp)
    if process_group is not None:
        dev_struct = all_reduce_partials(dev_struct, None if process_group is True else process_group)
    return _to_host(dev_struct, len(thresholds))


class MetricAccumulator:
    """Epoch-level running scores: the device-resident partials of every ``update`` are added on the
    GPU without a host sync; ``compute`` does one all-reduce (optional) and one D2H copy and returns the scores of
    everything seen since the last ``reset`` -- a ratio of sums over the epoch, where the reference's
    ``log_dict(on_epoch=True)`` (pipeline/helpers.py:151-153) averages per-batch ratios."""

    def __init__(self, thresholds: Sequence[float] = THRESHOLDS):
        self.thresholds = list(thresholds)
        self._acc: Optional[torch.Tensor] = None
        self.batches = 0

    def reset(self) -> None:
        self._acc, self.batches = None, 0

    def update(self, pred: torch.Tensor, target: torch.Tensor) -> None:
        part = metric_partials_device(pred, target, self.thresholds, clamp=True)
        self._acc = part if self._acc is None else _add_device(self._acc, part)
        self.batches += 1

    def partials(self, process_group=None) -> MetricPartials:
        if self._acc is None:
            raise RuntimeError("MetricAccumulator.compute() before any update()")
        dev_struct = self._acc
        if process_group is not None:
            dev_struct = all_reduce_partials(dev_struct, None if process_group is True else process_group)
        return _to_host(dev_struct, len(self.thresholds))

    def compute(self, extended: bool = False, process_group=None) -> Dict[str, float]:
        return scores_from_partials(self.partials(process_group), extended=extended)


# ------------------------------------------------------------------ scores from counts / sums
def _f32(x) -> np.float32:
    return np.float32(x)


def _csi_from_counts(c) -> float:
    """tp / (tp + fn + fp + 1e-8) in the reference's float32 operation order (metrics.py:53-54)."""
    tp, fn, fp = _f32(c[0]), _f32(c[1]), _f32(c[2])
    return float(tp / (tp + fn + fp + _f32(_eps)))


def _hss_from_counts(c) -> float:
    """metrics.py:66-69 in float32."""
    tp, fn, fp, tn = _f32(c[0]), _f32(c[1]), _f32(c[2]), _f32(c[3])
    with np.errstate(over="ignore", invalid="ignore"):
        num = _f32(2) * (tp * tn - fn * fp)
        den = (tp + fn) * (fn + tn) + (tp + fp) * (fp + tn) + _f32(_eps)
        return float(num / den)


def scores_from_partials(mp: MetricPartials, extended: bool = False) -> Dict[str, float]:
    """The reference's 56-key dict (metrics.py:96-131) from additive partials."""
    if mp.n_thresholds != 6:
        raise ValueError("calc_metrics uses the six reference thresholds")
    res: Dict[str, float] = {}
    n_frames = mp.n_frames
    res["CRPS"] = float(mp.abs_sum[0] / mp.n_elems[0])
    res["CRPS_4"] = float(mp.abs_sum[1] / mp.n_elems[1])
    res["CRPS_16"] = float(mp.abs_sum[2] / mp.n_elems[2])
    res["SSIM"] = float(mp.ssim_sum / n_frames)
    res["PSNR"] = float(mp.psnr_sum / n_frames)
    cnt = mp.counts
    for i in range(6):
        for pool, suffix in ((0, ""), (1, "_4"), (2, "_16")):
            res[f"CSI_{i}{suffix}"] = _csi_from_counts(cnt[pool, i])
        for pool, suffix in ((0, ""), (1, "_4"), (2, "_16")):
            res[f"HSS_{i}{suffix}"] = _hss_from_counts(cnt[pool, i])
    res["paper_SSIM"] = res["SSIM"]
    res["paper_PSNR"] = res["PSNR"]
    res["paper_CRPS"] = res["CRPS"]
    for pool_name, suffix in [("POOL1", ""), ("POOL4", "_4"), ("POOL16", "_16")]:
        csi_vals = [res[f"CSI_{i}{suffix}"] for i in range(6)]
        hss_vals = [res[f"HSS_{i}{suffix}"] for i in range(6)]
        res[f"paper_CSI_M_{pool_name}"] = float(np.mean(csi_vals))
        res[f"paper_CSI_181_{pool_name}"] = res[f"CSI_4{suffix}"]
        res[f"paper_CSI_219_{pool_name}"] = res[f"CSI_5{suffix}"]
        res[f"paper_HSS_{pool_name}"] = float(np.mean(hss_vals))
    if extended:
        res["MAE"] = res["CRPS"]
        res["MSE"] = float(mp.sq_sum / mp.n_elems[0])
        for i in range(6):
            for pool, suffix in ((0, ""), (1, "_4"), (2, "_16")):
                tp, fn, fp, _ = (float(v) for v in cnt[pool, i])
                res[f"POD_{i}{suffix}"] = tp / (tp + fn) if (tp + fn) > 0 else float("nan")
                res[f"FAR_{i}{suffix}"] = fp / (tp + fp) if (tp + fp) > 0 else float("nan")
    return res


# ------------------------------------------------------------------ reference-named entry points
def calc_metrics(pred, target, extended: bool = False, process_group=None) -> Dict[str, float]:
    """pred and target shape == (b, t, c, h, w); clamps to [0, 1] like the reference (metrics.py:86-133)."""
    return scores_from_partials(metric_partials(pred, target, THRESHOLDS, clamp=True, process_group=process_group),
                                extended=extended)


def _pool_idx(pool_type: str, scale: int) -> int:
    if pool_type == "none" or scale == 1:
        return 0
    if pool_type == "avg" and scale in _POOL_INDEX:
        return _POOL_INDEX[scale]
    raise NotImplementedError(f"pool_type={pool_type!r} scale={scale}: the fused kernel implements the pools "
                              "calc_metrics uses (none, avg 4, avg 16)")


def _hit_miss_fa_cn(pred, target, threshold):
    """Exact integer (tp, fn, fp, tn) as float64 scalars (reference returns float32 sums, :9-16)."""
    mp = metric_partials(pred, target, [threshold], clamp=False)
    c = mp.counts[0, 0]
    return float(c[0]), float(c[1]), float(c[2]), float(c[3])


def csi(pred, target, threshold, pool_type="none", scale=1):
    mp = metric_partials(pred, target, [threshold], clamp=False)
    return _csi_from_counts(mp.counts[_pool_idx(pool_type, scale), 0])


def hss(pred, target, threshold, pool_type="none", scale=1):
    mp = metric_partials(pred, target, [threshold], clamp=False)
    return _hss_from_counts(mp.counts[_pool_idx(pool_type, scale), 0])


def crps(pred, target, pool_type="none", scale=1):
    """One ensemble member: CRPS == mean |pred - target| (metrics.py:18-41 with n == 1)."""
    mp = metric_partials(pred, target, [0.5], clamp=False)
    i = _pool_idx(pool_type, scale)
    return float(mp.abs_sum[i] / mp.n_elems[i])


def ssim(pred, target):
    mp = metric_partials(pred, target, [0.5], clamp=False)
    return float(mp.ssim_sum / mp.n_frames)


def psnr(pred, target):
    mp = metric_partials(pred, target, [0.5], clamp=False)
    return float(mp.psnr_sum / mp.n_frames)

"""The three HBM-bound passes of the Path-B step in isolation at the BASELINE batch (32 sequences/GPU): uint8 staging,
the latent predictor, the fused skill scores. CUDA-event timing per launch, L2 flushed between iterations (a 256 MB
write), achieved GB/s = algorithmic bytes / time against the measured copy peak (MEASURED_PEAKS.json).
Also the command the `ncu --set full` captures of these kernels are taken from (`--iters 1 --warmup 1`).

    python scripts/hbm_bench.py [--batch 32] [--iters 10] [--warmup 3] [--only stage|predict|metrics]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from weatherforecastingtoolkit_b200 import _cabi  # noqa: E402
from weatherforecastingtoolkit_b200 import metrics as M  # noqa: E402
from weatherforecastingtoolkit_b200.rollout import LatentLinearPredictor, stage_vil  # noqa: E402
from weatherforecastingtoolkit_b200.synthetic import make_predictor_params, make_vil_sequences  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    _cabi.init(0)
    peak = 6463.7
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    B, H, W = args.batch, 384, 384
    # L2 flush: WRITE a 256 MB buffer, then READ another one -- the write alone would leave 126 MB of dirty lines whose
    # write-back the timed kernel then pays for (it doubled the time of the 30 MB predictor pass)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_r = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    res = {}

    def timeit(name, fn, nbytes):
        for _ in range(args.warmup):
            fn()
        ts = []
        for _ in range(args.iters):
            flush.fill_(1)
            flush_r.sum()
            torch.cuda._sleep(200000)   # ~0.1 ms of GPU spin: the host enqueues e0 / launch / e1 behind it (no launch gap in the timing)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        gbs = nbytes / us / 1e3
        res[name] = {"us_median": us, "us_min": ts[0], "algorithmic_bytes": nbytes, "gbs": gbs, "frac_of_measured_peak": gbs / peak}
        print(f"{name:28s} {us:9.1f} us  {gbs:8.1f} GB/s  = {gbs / peak:6.3f} of {peak:.0f} GB/s (measured copy peak)", flush=True)

    if args.only in (None, "stage"):
        u8 = make_vil_sequences(min(B, 4), H, W, 25, seed=1).repeat((B + 3) // 4, 1, 1, 1)[:B].contiguous().to(dev)
        timeit("stage_vil u8->f32", lambda: stage_vil(u8), u8.numel() * 5.0)
        timeit("stage_vil u8->f16", lambda: stage_vil(u8, dtype=torch.float16), u8.numel() * 3.0)
        del u8
    if args.only in (None, "predict"):
        w, b = make_predictor_params(seed=0)
        pred = LatentLinearPredictor().to(dev)
        pred.weight.data.copy_(w)
        pred.bias.data.copy_(b)
        lat = torch.randn(B, 25, 4, 48, 48, device=dev)
        lib = _cabi.load()
        po, to = torch.empty(B, 12, 4, 48, 48, device=dev), torch.empty(B, 12, 4, 48, 48, device=dev)
        ls = torch.zeros(2, dtype=torch.float64, device=dev)
        wt, bs = pred.weight.detach().contiguous(), pred.bias.detach().contiguous()
        st = torch.cuda.current_stream().cuda_stream
        # the C call alone (rollout() adds a zero-fill and two scalar kernels around it)
        timeit("predict_linear 52->48", lambda: _cabi.check(lib.wfk_predict_linear(
            lat.data_ptr(), wt.data_ptr(), bs.data_ptr(), B, 13, 12, 4, 48 * 48, po.data_ptr(), to.data_ptr(), ls.data_ptr(), st)),
            921600.0 * B)
        timeit("predictor.rollout (API)", lambda: pred.rollout(lat), 921600.0 * B)
    if args.only in (None, "metrics"):
        torch.manual_seed(0)
        u8 = make_vil_sequences(2, H, W, 13, seed=31).to(dev)
        x = stage_vil(u8)                                    # [2, 13, 1, H, W]
        reps = (B * 12 + 23) // 24
        pr = x[:, :12].repeat(reps, 1, 1, 1, 1)[: B * 12 // 12].contiguous() if False else x[:, :12].repeat(reps, 1, 1, 1, 1).reshape(-1, 1, H, W)[: B * 12].contiguous()
        tg = x[:, 1:13].repeat(reps, 1, 1, 1, 1).reshape(-1, 1, H, W)[: B * 12].contiguous()
        pr = (pr + 0.02 * torch.randn_like(pr)).contiguous()
        import ctypes as C
        lib = _cabi.load()
        nf = pr.shape[0]
        ws_bytes = lib.wfk_metrics_workspace_bytes(nf, H, W)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        outp = torch.empty(108, dtype=torch.int64, device=dev)
        thr = (C.c_float * 6)(*[float(torch.tensor(v, dtype=torch.float32)) for v in M.THRESHOLDS])
        st = torch.cuda.current_stream().cuda_stream
        timeit(f"wfk_metrics {nf} frame pairs", lambda: _cabi.check(lib.wfk_metrics(
            pr.data_ptr(), tg.data_ptr(), nf, H, W, thr, 6, 1, outp.data_ptr(), ws.data_ptr(), ws_bytes, st)), 8.0 * pr.numel())
        timeit("metric_partials_device (API)", lambda: M.metric_partials_device(pr, tg), 8.0 * pr.numel())
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

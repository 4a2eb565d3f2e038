#!/usr/bin/env python
"""Benchmark of the Path-B hot path: forecast frames/sec for the reference-faithful validation step
(stage 25 uint8 frames -> encode 25 -> Linear 52->48 -> decode 12 pred + 12 target -> fused
CSI/HSS/CRPS/SSIM/PSNR) at BASELINE.json configs[1]: batch 32 sequences of 384x384 per GPU.

    python bench.py --gpus N --steps K --warmup W            # B200 arm (one JSON line on rank 0)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

A "step" is one pass of the hot path over one batch of 32 synthetic sequences per GPU. `value` is
whole-job forecast frames/s with the uint8 batch already resident in HBM; `e2e` is the same metric
through the public API (PathBNowcast.evaluate) with a pinned HOST uint8 batch: H2D copy and the D2H
read of the score dict are inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# nominal 2*MAC work (FlopCounterMode on the reference modules)
GF_ENCODE = 618.985
GF_DECODE = 1405.282
GF_PRED = 0.0115
GF_PER_SEQ = 25 * GF_ENCODE + 24 * GF_DECODE + GF_PRED      # 49 201.4
GF_PER_FORECAST_FRAME = GF_PER_SEQ / 12                    # 4 100.1
H = W = 384
T_IN, T_OUT = 13, 12


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1410.9))), "hbm": float(d["hbm_gbs"]),
                "which": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "hbm": 6650.0, "which": "fallback (B200_PROFILING.md)"}


def bench_config(args, world):
    """The `config` object of BOTH arms (the driver compares them): the workload and how it is cut."""
    return {"workload": WORKLOAD, "batch_per_gpu": args.batch, "frames_per_call": args.frames_per_call,
            "posterior": "mode", "l2": "inputs (118 MB uint8 + GB-scale activations) larger than the 126 MB L2",
            "parallelism": f"dp{world} (sequence shards, one all-reduce of the 864-byte partials)"}


def read_ncu_raw(path):
    """Rows of an `ncu --page raw --csv` export as dicts {metric: float | str} (the unit row is dropped; Mbyte /
    Kbyte / Gbyte and ms / us / ns are normalised to bytes and microseconds)."""
    import csv
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
             "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
    out = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            try:
                d[h] = float(v.replace(",", "")) * scale.get(u, 1.0)
            except ValueError:
                d[h] = v
        out.append(d)
    return out


def ncu_traffic(patterns):
    """DRAM bytes of ONE launch of the dominant kernel, read from the newest committed `ncu --set full` raw export
    under profiles/ whose name matches one of `patterns` (never typed in by hand)."""
    import glob
    for pat in patterns:
        for path in sorted(glob.glob(os.path.join(ROOT, "profiles", pat)), reverse=True):
            rows = [r for r in read_ncu_raw(path) if "dram__bytes_read.sum" in r]
            if rows:
                r = rows[0]
                return {"source": os.path.relpath(path, ROOT), "kernel": r.get("Kernel Name"),
                        "dram_bytes": r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"],
                        "duration_us": r.get("gpu__time_duration.sum"),
                        "tensor_pipe_pct_elapsed": r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")}
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        pw.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w": pw[len(pw) // 2] if pw else None}


# ------------------------------------------------------------------------------------ CPU arms
def _cpu_sample(threads: int):
    """One bounded sample of the reference's CPU path (oracle port = the reference's own torch ops on a
    state_dict): encode 1 frame + decode 1 frame at 384x384, predictor on 1 sequence of latents,
    calc_metrics on 12 frame pairs; extrapolated to the 25 encodes + 24 decodes of one sequence."""
    import torch
    from oracle import akl_oracle as O
    from oracle import metrics_oracle as MO
    from weatherforecastingtoolkit_b200.synthetic import (PATHB_AKL_CONFIG, make_akl_state_dict, make_predictor_params,
                                                          make_vil_sequences)
    torch.set_num_threads(threads)
    cfg = PATHB_AKL_CONFIG
    if not hasattr(_cpu_sample, "state"):
        sd = make_akl_state_dict(cfg, 0)
        w, b = make_predictor_params(seed=0)
        u8 = make_vil_sequences(1, H, W, 25, seed=1)
        _cpu_sample.state = (sd, w, b, u8)
    sd, w, b, u8 = _cpu_sample.state
    with torch.no_grad():
        t0 = time.perf_counter()
        v = O.stage_vil(u8).permute(0, 3, 1, 2).unsqueeze(2)
        t_stage = time.perf_counter() - t0
        t0 = time.perf_counter()
        mom = O.akl_encode_moments(v[:, 0], sd, cfg)
        t_enc = time.perf_counter() - t0
        z = mom[:, :4].contiguous()
        t0 = time.perf_counter()
        dec = O.akl_decode(z, sd, cfg)
        t_dec = time.perf_counter() - t0
        lat = z.unsqueeze(1).repeat(1, 25, 1, 1, 1) + 0.01 * torch.randn(1, 25, 4, 48, 48)
        t0 = time.perf_counter()
        O.predictor_rollout(lat, w, b)
        t_pred = time.perf_counter() - t0
        p = dec.unsqueeze(1).repeat(1, 12, 1, 1, 1)
        tg = v[:, 13:25].contiguous()
        t0 = time.perf_counter()
        MO.calc_metrics(p, tg)
        t_met = time.perf_counter() - t0
    per_seq = t_stage + 25 * t_enc + 24 * t_dec + t_pred + t_met
    return {"per_seq_s": per_seq, "t_enc": t_enc, "t_dec": t_dec, "t_metrics": t_met, "t_pred": t_pred,
            "measured_s": t_stage + t_enc + t_dec + t_pred + t_met}


CPU_SAMPLE_DESC = ("per step: reference torch-CPU ops (oracle port, fp32) on 1 sequence of 384x384: 1 encode + 1 decode "
                   "+ predictor + calc_metrics on 12 frame pairs, extrapolated to 25 encodes + 24 decodes")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        _cpu_sample(threads)
    t_total, per_seq = 0.0, []
    for _ in range(args.steps):
        r = _cpu_sample(threads)
        per_seq.append(r["per_seq_s"])
        t_total += r["measured_s"]
    ps = sum(per_seq) / len(per_seq)
    fps = T_OUT / ps
    line = {
        "impl": "reference", "metric": "forecast_frames_per_sec", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, int(os.environ.get("WORLD_SIZE", str(args.gpus)))),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": CPU_SAMPLE_DESC,
                         "note": "CPU path; value extrapolated from the bounded sample to the 25 encodes + 24 decodes of a sequence"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


WORKLOAD = ("Path-B validation_step, BASELINE configs[1]: 32 sequences/GPU x (13 in + 12 out) uint8 384x384 VIL frames; "
            "stage -> AutoencoderKL encode x25 -> Linear(52->48) -> decode 12 pred + 12 target -> fused "
            "CSI/HSS/CRPS/SSIM/PSNR (6 thresholds x 3 pools)")


# ------------------------------------------------------------------------------------ B200 arm
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from weatherforecastingtoolkit_b200 import _cabi, engine
    from weatherforecastingtoolkit_b200 import metrics as M
    from weatherforecastingtoolkit_b200.rollout import PathBNowcast
    from weatherforecastingtoolkit_b200.synthetic import (PATHB_AKL_CONFIG, make_akl_state_dict, make_predictor_params,
                                                          make_vil_sequences)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # NCCL prints its version banner (and anything NCCL_DEBUG asks for) on stdout: point fd 1 at stderr while the job
    # runs so that the ONE JSON line is the only thing rank 0 writes to the real stdout
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.init(local_rank)

    cfg = PATHB_AKL_CONFIG
    net = PathBNowcast(cfg, posterior="mode", frames_per_call=args.frames_per_call)
    net.autoencoder.autoencoder.load_state_dict(make_akl_state_dict(cfg, 0), strict=True)
    w, b = make_predictor_params(seed=0)
    net.predictor.weight.data.copy_(w)
    net.predictor.bias.data.copy_(b)
    net = net.to(dev)

    B = args.batch
    host = make_vil_sequences(B, H, W, T_IN + T_OUT, seed=1 + rank).pin_memory()
    dev_batch = host.to(dev)
    group = True if world > 1 else None

    def step_device():
        dp, dt, loss = net.validation_step(dev_batch)
        return M.metric_partials_device(dp, dt), loss

    def finish(dev_struct):
        if world > 1:
            dev_struct = M.all_reduce_partials(dev_struct)
        return dev_struct

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (builds plans, allocates buffers)
    for _ in range(max(args.warmup, 1)):
        finish(step_device()[0])
    barrier()

    # ---- timed region 1: inputs resident in HBM
    timer = engine.KernelTimer(all_ops=bool(args.breakdown))
    engine.TIMER = timer
    sampler = ClockSampler(local_rank)
    launches0 = _cabi.launch_count()
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for _ in range(args.steps):
        last = finish(step_device()[0])
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    engine.TIMER = None
    launches = _cabi.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    tsum = timer.summary()
    sec = timer.secondary()
    # ---- on-device check of the count all-reduce: the NCCL-reduced struct must equal the int64 sum of the
    # ranks' own partials, and cover world x local elements
    parity_check = None
    if world > 1:
        local = step_device()[0]
        reduced = M.all_reduce_partials(local)
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        ni = M._N_INT
        want_i = torch.stack([g[:ni] for g in gathered]).sum(0)
        want_f = torch.stack([g[ni:].view(torch.float64) for g in gathered]).sum(0)
        ok = bool(torch.equal(reduced[:ni], want_i)) and bool(torch.allclose(reduced[ni:].view(torch.float64), want_f, rtol=1e-12))
        ok = ok and int(reduced[96].item()) == world * int(local[96].item()) and int(reduced[99].item()) == world * B * T_OUT
        okt = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        parity_check = "ok" if int(okt.item()) == 1 else "FAILED"
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    scores = M.scores_from_partials(M._to_host(last, 6), extended=True)
    if args.breakdown and rank == 0:
        agg = {}
        for what, fl, s0, s1 in timer.records:
            a = agg.setdefault(what, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += s0.elapsed_time(s1)
            a[2] += fl
        rows = [{"layer": k, "launches": v[0], "ms": v[1], "nominal_tflops": v[2] / (v[1] / 1e3) / 1e12 if v[1] > 0 else 0.0,
                 "gflop_per_launch": v[2] / v[0] / 1e9} for k, v in agg.items()]
        rows.sort(key=lambda r: -r["ms"])
        with open(args.breakdown, "w") as f:
            json.dump(rows, f, indent=1)

    # ---- timed region 2: end to end through the public API, HOST input, score dict read back
    barrier()
    e0.record()
    res = None
    for _ in range(args.steps):
        res = net.evaluate(host.to(dev, non_blocking=True), process_group=group)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())

    if rank == 0:
        peaks = _peaks()
        frames_total = world * B * T_OUT * args.steps
        fps = frames_total / (ms / 1e3)
        fps_e2e = frames_total / (ms_e2e / 1e3)
        achieved = tsum["nominal_flops"] / (tsum["ms"] / 1e3) / 1e12 if tsum["ms"] > 0 else 0.0
        executed = tsum["executed_flops"] / (tsum["ms"] / 1e3) / 1e12 if tsum["ms"] > 0 else 0.0
        traffic = ncu_traffic(["r2_*conv_gemm*n256*raw.csv", "r1_prof_final4_n256_raw.csv"])
        traffic128 = ncu_traffic(["r2_*conv_gemm*n128*raw.csv", "r1_prof_final4_n128_raw.csv"])
        secondary = {}
        for name, d in sec.items():
            gbs = d["bytes"] / (d["ms"] / 1e3) / 1e9 if d["ms"] > 0 else 0.0
            secondary[name] = {"bound": "hbm", "gbs": gbs, "frac": gbs / peaks["hbm"], "launches": d["launches"],
                               "us_per_launch": 1e3 * d["ms"] / max(d["launches"], 1),
                               "algorithmic_bytes_per_launch": d["bytes"] / max(d["launches"], 1)}
        line = {
            "metric": "forecast_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": ("bf16" if os.environ.get("WFK_OPERANDS", "fp16").lower() == "bf16" else "fp16") +
                     " operands, fp32 accumulate (tcgen05 kind::f16)", "data": "synthetic",
            "config": bench_config(args, world),
            "clocks": clocks,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(host.numel()),
                    "d2h_bytes_per_step": 108 * 8 + 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM)",
                "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                # `achieved` / `frac` use the NOMINAL 2*M*N*K of the reference layers; `executed_tflops` /
                # `frac_executed` count the MMA work really issued (sub-pixel upsample: 2.25x fewer MACs; identity-tap
                # residuals: a few more). `traffic`: DRAM bytes of ONE launch of the dominant variant, read from the
                # committed `ncu --set full` raw export named in traffic_ncu[*].source.
                "executed_tflops": executed, "frac_executed": executed / peaks["tflops"],
                "traffic": traffic["dram_bytes"] if traffic else None,
                "traffic_ncu": {"n256": traffic, "n128": traffic128},
                "secondary": secondary,
                "peak_source": peaks["which"], "hbm_peak_gbs": peaks["hbm"],
                "launches_timed": tsum["launches"], "kernel_ms_per_step": tsum["ms"] / args.steps,
                "kernel_share_of_step": tsum["ms"] / ms if ms > 0 else None,
                "step_tflops_nominal": fps * GF_PER_FORECAST_FRAME / 1e3 / world,
                "step_frac_of_peak": fps * GF_PER_FORECAST_FRAME / 1e3 / world / peaks["tflops"],
                "step_frac_of_peak_executed": (fps * GF_PER_FORECAST_FRAME / 1e3 / world / peaks["tflops"]
                                               * (tsum["executed_flops"] / tsum["nominal_flops"]) if tsum["nominal_flops"] else None),
            },
            "parity_check": parity_check,
            "scores": {k: scores[k] for k in ("CSI_0", "CSI_3", "SSIM", "CRPS", "POD_0", "FAR_0", "MSE")},
        }
        if args.cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            _cpu_sample(threads)
            r = _cpu_sample(threads)
            line["cpu_baseline"] = {"value": T_OUT / r["per_seq_s"], "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": CPU_SAMPLE_DESC, "sample_seconds": r["measured_s"]}
        else:
            line["cpu_baseline"] = None
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------ BASELINE configs 0 / 3 / 4
AUX = {
    # name: (BASELINE.json config index, description, nominal GFLOP per frame, frame size, default frames per GPU)
    "posaware": (0, "PosAwareAE_TF (2048-d bottleneck) encode + decode of 128x128 frames (its only valid input size)", 34.77, 128, 128),
    "vit": (3, "AE_ViT_2048 forward, token-sequence latent [64, 512], 128x128 frames, batch 64", 4.97, 128, 64),
    "disc": (4, "NLayerDiscriminator forward + hinge sums on 12-frame forecasts of 384x384, 16 sequences per GPU "
                "(= batch 128 across 8 GPUs)", 14.175, 384, 192),
}


def run_aux_arm(args):
    """The other model families of BASELINE.json `configs` behind the same contract: one step = one forward pass of
    `--batch` frames per GPU (inputs resident in HBM for `value`; pinned host input + host read of the result for
    `e2e`). Parity for these lives in tests/test_posaware.py, test_vit.py, test_discriminator.py."""
    import torch
    import torch.distributed as dist

    from weatherforecastingtoolkit_b200 import _cabi, engine
    from weatherforecastingtoolkit_b200 import synthetic as S

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.init(local_rank)
    idx, desc, gflop, hw, default_b = AUX[args.config]
    B = args.batch if args.batch != 32 else default_b
    if args.config == "posaware":
        from weatherforecastingtoolkit_b200.models.ae_64x8x8_lin import PosAwareAE_TF
        m = PosAwareAE_TF().eval()
        m.load_state_dict(S.fill_state_dict(m, "posaware", 0, gain=1.1))
        fwd = lambda x: m(x)[0].mean()                                   # noqa: E731
    elif args.config == "vit":
        from weatherforecastingtoolkit_b200.models.ae_vit import AE_ViT_2048
        m = AE_ViT_2048().eval()
        m.load_state_dict(S.fill_state_dict(m, "vit", 0))
        fwd = lambda x: m(x)[0].mean()                                   # noqa: E731
    else:
        from weatherforecastingtoolkit_b200.models.autoencoderkl.losses import NLayerDiscriminator
        m = NLayerDiscriminator(input_nc=1).eval()
        m.load_state_dict(S.make_discriminator_state_dict())
        fwd = lambda x: m(x).mean()                                      # noqa: E731
    m = m.to(dev)
    g = torch.Generator().manual_seed(1 + rank)
    host = torch.rand(B, 1, hw, hw, generator=g).pin_memory()
    x_dev = host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            fwd(x_dev)
        barrier()
        # ---- timed region 1: inputs resident in HBM (the programs replay their CUDA graphs)
        sampler = ClockSampler(local_rank)
        l0 = _cabi.launch_count()
        if rank == 0:
            sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = fwd(x_dev)
        e1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = e0.elapsed_time(e1)
        # ---- kernel-time accounting for the roofline leg: the same steps launched eagerly with CUDA events around every
        # tensor-core launch (not part of `value`)
        timer = engine.KernelTimer()
        engine.TIMER = timer
        l1 = _cabi.launch_count()
        for _ in range(args.steps):
            fwd(x_dev)
        barrier()
        engine.TIMER = None
        # kernels of the timed region: graph replays do not pass through the C ABI counter, the eager pass launches the
        # same op list
        launches = max(_cabi.launch_count() - l1, _cabi.launch_count() - l0 - (_cabi.launch_count() - l1))
        tsum = timer.summary()
        # ---- timed region 2: end to end, pinned host input, host read of the result
        barrier()
        e0.record()
        for _ in range(args.steps):
            res = float(fwd(host.to(dev, non_blocking=True)).item())
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = (float(v) for v in t.tolist())
    if rank == 0:
        peaks = _peaks()
        frames = world * B * args.steps
        fps, fps_e2e = frames / (ms / 1e3), frames / (ms_e2e / 1e3)
        ach = tsum["nominal_flops"] / (tsum["ms"] / 1e3) / 1e12 if tsum["ms"] > 0 else 0.0
        line = {
            "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 operands, fp32 accumulate (tcgen05 kind::f16)", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[{idx}]: {desc}", "batch_per_gpu": B, "frame": f"{hw}x{hw}",
                       "l2": "activations of a step exceed the 126 MB L2" if B * hw * hw * 256 > (126 << 20) else "inputs smaller than L2; back-to-back steps",
                       "parallelism": f"dp{world} (frame shards, no collective)"},
            "clocks": clocks,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(host.numel() * 4), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM / GEMM)",
                         "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": ach / peaks["tflops"],
                         "traffic": None, "peak_source": peaks["which"], "launches_timed": tsum["launches"],
                         "kernel_ms_per_step": tsum["ms"] / args.steps, "kernel_share_of_step": tsum["ms"] / ms if ms else None,
                         "step_tflops_nominal": fps * gflop / 1e3 / world,
                         "step_frac_of_peak": fps * gflop / 1e3 / world / peaks["tflops"],
                         "nominal_gflop_per_frame": gflop},
            "cpu_baseline": None, "result": res,
        }
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="sequences per GPU (BASELINE configs[1]: 32)")
    ap.add_argument("--frames-per-call", type=int, default=37,
                    help="frames per AutoencoderKL call; 37 makes every layer's tile count a multiple of the 74 CTA pairs")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--breakdown", default=None, help="write a per-layer conv-GEMM timing table (JSON) to this path")
    ap.add_argument("--config", default="pathb", choices=["pathb"] + sorted(AUX),
                    help="pathb = BASELINE configs[1]/[2] (the headline metric); posaware / vit / disc = configs[0] / [3] / [4]")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.config != "pathb":
        return run_aux_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())

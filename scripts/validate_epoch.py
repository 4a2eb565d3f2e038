"""A whole validation epoch on the GPU path, end to end (f.2 -> hot path -> f.3):

    uint8 SEVIR events (host) --DeviceSEVIRLoader--> [B, H, W, 25] windows on the GPU
        --PathBNowcast.validation_step--> decoded forecast / target frames
        --MetricAccumulator--> epoch scores (ratio of sums, one D2H at the end)
        --render.panel_mosaics--> uint8 RGBA panels of the first batch

With no arguments it runs on synthetic events and random-init weights; pass --events file.npy (uint8 [E, H, W, 49]) and
--akl / --predictor state_dict files to score real data.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from weatherforecastingtoolkit_b200 import metrics as M
from weatherforecastingtoolkit_b200 import render
from weatherforecastingtoolkit_b200.datastage import DeviceSEVIRLoader
from weatherforecastingtoolkit_b200.rollout import PathBNowcast
from weatherforecastingtoolkit_b200.synthetic import PATHB_AKL_CONFIG, make_akl_state_dict, make_predictor_params, make_vil_sequences


def run_epoch(events, net, batch_size=8, panels=2):
    loader = DeviceSEVIRLoader(events, seq_len=25, raw_seq_len=events.shape[3], stride=12, batch_size=batch_size,
                               layout="NHWT", split_mode="floor")
    acc = M.MetricAccumulator()
    mosaics, losses = None, []
    for batch in loader:
        dp, dt, loss = net.validation_step(batch["vil"])      # float [B, H, W, T] in [0, 1], already on the GPU
        acc.update(dp, dt)
        losses.append(loss)
        if mosaics is None and panels:
            mosaics = render.panel_mosaics(dp, dt, batch_idxs=panels)
    scores = acc.compute(extended=True)
    scores["val_loss"] = float(torch.stack(losses).mean().item())
    return scores, mosaics, loader


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--events", default=None, help=".npy with uint8 [E, H, W, 49] events (default: synthetic)")
    ap.add_argument("--akl", default=None, help="AutoencoderKL state_dict (torch.save); default: random init")
    ap.add_argument("--predictor", default=None, help="nn.Linear(52, 48) state_dict; default: random init")
    ap.add_argument("--num-events", type=int, default=8)
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--batch-size", type=int, default=8)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    if args.events:
        events = np.load(args.events, mmap_mode="r")
    else:
        events = make_vil_sequences(args.num_events, args.size, args.size, 49, seed=3).numpy()
    net = PathBNowcast(PATHB_AKL_CONFIG, posterior="mode")
    net.autoencoder.autoencoder.load_state_dict(torch.load(args.akl) if args.akl else make_akl_state_dict(PATHB_AKL_CONFIG, 0))
    if args.predictor:
        net.predictor.load_state_dict(torch.load(args.predictor))
    else:
        w, b = make_predictor_params(seed=0)
        net.predictor.weight.data.copy_(w)
        net.predictor.bias.data.copy_(b)
    net = net.to(dev)
    run_epoch(events[:2], net, batch_size=min(args.batch_size, 6), panels=0)   # warm-up: pack weights, build programs
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scores, mosaics, loader = run_epoch(events, net, batch_size=args.batch_size)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nseq = len(loader) * args.batch_size
    print(json.dumps({"sequences": nseq, "seconds": dt, "forecast_frames_per_s": nseq * 12 / dt,
                      "h2d_MB": loader.h2d_bytes / 1e6, "panel_shape": list(mosaics[0].shape),
                      "scores": {k: scores[k] for k in ("CSI_0", "CSI_3", "SSIM", "CRPS", "MSE", "val_loss")}}))


if __name__ == "__main__":
    main()

"""Goldens for the parts of ``pipeline/metrics.py`` outside calc_metrics' fixed sweep (pool_type='max', arbitrary
``scale``, ensemble forecasts), computed with the UNMODIFIED reference module (torchmetrics stubbed as in
make_golden.py; only SSIM / PSNR of the ensemble calc_metrics case pass through the stub):

    python tests/golden/make_golden_metrics_extra.py
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import _stub_torchmetrics  # noqa: E402

CASES = [("max", 4), ("max", 3), ("avg", 3), ("avg", 8), ("avg", 2), ("max", 16)]
THS = [16 / 255, 133 / 255, 0.7]


def inputs():
    torch.manual_seed(123)
    p = torch.rand(2, 3, 1, 50, 70) * 1.1 - 0.05
    t = torch.rand(2, 3, 1, 50, 70) * 1.1 - 0.05
    ens = torch.rand(2, 3, 4, 1, 48, 64) * 1.2 - 0.1          # (b, n, t, c, h, w)
    gt = torch.rand(2, 4, 1, 48, 64)
    return p, t, ens, gt


def main():
    _stub_torchmetrics()
    import pipeline.metrics as RM
    p, t, ens, gt = inputs()
    out = {"pooled": [], "crps_ensemble": [], "ensemble_calc_metrics": RM.calc_metrics(ens, gt)}
    for pool, scale in CASES:
        row = {"pool_type": pool, "scale": scale, "crps": RM.crps(p, t, pool, scale)}
        for th in THS:
            row[f"csi_{th:.6f}"] = RM.csi(p, t, th, pool, scale)
            row[f"hss_{th:.6f}"] = RM.hss(p, t, th, pool, scale)
        out["pooled"].append(row)
    for pool, scale in [("none", 1), ("avg", 4), ("max", 2), ("avg", 16), ("avg", 3)]:
        out["crps_ensemble"].append({"pool_type": pool, "scale": scale, "crps": RM.crps(ens, gt, pool, scale)})
    out["hit_miss_fa_cn"] = [float(v) for v in RM._hit_miss_fa_cn(p, t, THS[1])]
    with open(os.path.join(HERE, "metrics_extra_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("ok", out["crps_ensemble"][0], out["pooled"][0]["crps"])


if __name__ == "__main__":
    main()

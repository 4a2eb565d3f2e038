"""Throughput of the auxiliary configs of BASELINE.json (parity-test cases, NOT the bench.py line): config 1
PosAwareAE_TF encode+decode at its native 128x128, config 4 AE_ViT_2048 forward (batch 64), config 5
NLayerDiscriminator forward scoring of 384x384 frames. CUDA-event timing, inputs resident in HBM. Nominal FLOPs per
frame from SURVEY.md 8(d): 34.77 / 4.97 / 14.175 GFLOP."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from weatherforecastingtoolkit_b200 import synthetic as S
from weatherforecastingtoolkit_b200.models.ae_64x8x8_lin import PosAwareAE_TF
from weatherforecastingtoolkit_b200.models.ae_vit import AE_ViT_2048
from weatherforecastingtoolkit_b200.models.autoencoderkl.losses import NLayerDiscriminator


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    out = []
    dev = "cuda:0"
    m = PosAwareAE_TF().eval()
    m.load_state_dict(S.fill_state_dict(m, "posaware", 0, gain=1.1))
    for b in (4, 32, 128):
        x = torch.rand(b, 1, 128, 128, device=dev)
        ms = timeit(lambda: m(x))
        out.append({"config": "1: PosAwareAE_TF encode+decode 128x128", "batch": b, "ms": ms, "frames_per_s": b / ms * 1e3,
                    "nominal_tflops": b * 34.77e9 / (ms * 1e-3) / 1e12})
    v = AE_ViT_2048().eval()
    v.load_state_dict(S.fill_state_dict(v, "vit", 0))
    for b in (64, 256):
        x = torch.rand(b, 1, 128, 128, device=dev)
        ms = timeit(lambda: v(x))
        out.append({"config": "4: AE_ViT_2048 forward 128x128", "batch": b, "ms": ms, "frames_per_s": b / ms * 1e3,
                    "nominal_tflops": b * 4.97e9 / (ms * 1e-3) / 1e12})
    d = NLayerDiscriminator(input_nc=1).eval()
    d.load_state_dict(S.make_discriminator_state_dict())
    for b in (12, 192):
        x = torch.rand(b, 1, 384, 384, device=dev)
        ms = timeit(lambda: d(x))
        out.append({"config": "5: NLayerDiscriminator forward 384x384", "batch": b, "ms": ms, "frames_per_s": b / ms * 1e3,
                    "nominal_tflops": b * 14.175e9 / (ms * 1e-3) / 1e12})
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()

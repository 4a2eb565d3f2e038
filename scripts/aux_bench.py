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
    # ---- SURVEY 8f components: latent compressors on 48x48x4 latents, panel rendering, SEVIR device staging
    from weatherforecastingtoolkit_b200.predictors import ConvAttnModel, ConvModel
    from weatherforecastingtoolkit_b200 import render
    from weatherforecastingtoolkit_b200.datastage import DeviceSEVIRLoader
    ca = ConvAttnModel().eval()
    cm = ConvModel().eval()
    for b in (148, 32 * 25):
        x = torch.randn(b, 4, 48, 48, device=dev)
        ms = timeit(lambda: ca(x))
        out.append({"config": "f.4: ConvAttnModel forward (one CTA per frame, fp32)", "batch": b, "ms": ms,
                    "frames_per_s": b / ms * 1e3, "nominal_tflops": b * 0.60e9 / (ms * 1e-3) / 1e12})
        x5 = x.unsqueeze(0)
        ms = timeit(lambda: cm(x5))
        out.append({"config": "f.4: ConvModel forward (one CTA per frame, fp32)", "batch": b, "ms": ms,
                    "frames_per_s": b / ms * 1e3})
    p = torch.rand(32, 12, 1, 384, 384, device=dev)
    t = torch.rand(32, 12, 1, 384, 384, device=dev)
    ms = timeit(lambda: render.render_panels(p, t))
    px = p.numel()
    out.append({"config": "f.3: render_panels 32x12 frames 384x384 (8 B read + 15 B written per pixel)", "ms": ms,
                "frames_per_s": 384 / ms * 1e3, "GB_per_s": px * 23 / (ms * 1e-3) / 1e9})
    ev = S.make_loader_events(24, 384, 384, 49, seed=3).numpy()
    ld = DeviceSEVIRLoader(ev, batch_size=32, layout="NTCHW")

    def one_pass():
        ld.reset()
        for _ in ld:
            pass
    ms = timeit(one_pass, reps=3, warm=1)
    nseq = len(ld) * 32
    out.append({"config": "f.2: DeviceSEVIRLoader pass, 24 events -> 72 sequences (uint8 H2D + window kernel)", "ms": ms,
                "sequences_per_s": nseq / ms * 1e3, "h2d_MB_per_pass": 24 * 384 * 384 * 49 / 1e6,
                "reference_fp32_h2d_MB_per_pass": nseq * 384 * 384 * 25 * 4 / 1e6})
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()

"""Per-layer micro-benchmark of the tcgen05 conv-GEMM kernel (32 frames, Path-B layer shapes)."""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from weatherforecastingtoolkit_b200 import _cabi
from weatherforecastingtoolkit_b200.engine import AKLEngine, PackedAKL, _Act, _Pool, _Program
from weatherforecastingtoolkit_b200.synthetic import PATHB_AKL_CONFIG, make_akl_state_dict

dev = torch.device("cuda:0")


def harness(eng, n):
    self = object.__new__(_Program)
    self.eng, self.lib, self.dev = eng, eng.lib, eng.device
    self.pool = _Pool(self.dev, eng.adt)
    self.bf = 1 if eng.bf16 else 0
    self.ops, self.plans, self.keep = [], [], []
    self.n = n
    self.stats_arena = torch.zeros(8, n, eng.groups, 2, dtype=torch.float64, device=self.dev)
    self._stats_used = 0
    return self


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    reps = 5
    eng = AKLEngine(PATHB_AKL_CONFIG, make_akl_state_dict(PATHB_AKL_CONFIG, 0), device=dev)
    cases = [
        ("plain", 384, 128, 128), ("residual", 384, 128, 128), ("plain", 384, 256, 128), ("shortcut", 384, 128, 128),
        ("plain", 192, 256, 256), ("residual", 192, 256, 256), ("plain", 192, 512, 256),
        ("plain", 96, 512, 512), ("residual", 96, 512, 512), ("residual", 48, 512, 512),
        ("up", 192, 256, 256), ("up", 96, 512, 512), ("down", 384, 128, 128),
        ("plain", 384, 128, 256), ("plain", 384, 128, 512),
        ("gn", 384, 128, 128), ("gn+res", 384, 128, 128), ("gn", 384, 256, 128), ("gn+sc", 384, 128, 128),
        ("gn", 192, 256, 256), ("gn+res", 96, 512, 512), ("gn+sc", 192, 256, 256),
    ]
    stream = torch.cuda.current_stream().cuda_stream
    if os.environ.get("CASE"):
        cases = [cases[int(os.environ["CASE"])]]
    for mode, hw, cin, cout in cases:
        hs = harness(eng, n)
        x = torch.randn(n, hw, hw, cin, device=dev).half()
        wt = torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(9 * cin)
        bias = torch.randn(cout, device=dev)
        t = eng.w.t
        t["tmp.w"] = wt.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous().half()
        if mode == "plain":
            hs.conv3x3(_Act(x, None), "tmp.w", bias, cout, want_stats=not os.environ.get("NOSTATS"))
        elif mode == "residual":
            res = torch.randn(n, hw, hw, cout, device=dev).half()
            hs.conv3x3(_Act(x, None), "tmp.w", bias, cout, residual=res)
        elif mode == "shortcut":
            xs = torch.randn(n, hw, hw, 2 * cin, device=dev).half()
            t["tmp.sc"] = (torch.randn(1, cout, 2 * cin, device=dev) / math.sqrt(2 * cin)).half()
            hs.conv3x3(_Act(x, None), "tmp.w", bias, cout, shortcut=(xs, "tmp.sc"))
        elif mode.startswith("gn"):
            tab = torch.randn(n, cin, 2, device=dev) * 0.3 + 0.5
            res = torch.randn(n, hw, hw, cout, device=dev).half() if mode == "gn+res" else None
            sc = None
            if mode == "gn+sc":
                xs = torch.randn(n, hw, hw, 2 * cin, device=dev).half()
                t["tmp.sc"] = (torch.randn(1, cout, 2 * cin, device=dev) / math.sqrt(2 * cin)).half()
                sc = (xs, "tmp.sc")
            hs.conv3x3(_Act(x, None), "tmp.w", bias, cout, residual=res, shortcut=sc, gn_tab=tab)
        elif mode == "up":
            t["tmp.w"] = PackedAKL._phase_weights(wt)
            hs.upsample(_Act(x, None), "tmp.w", bias)
        elif mode == "down":
            hs.downsample(_Act(x, None), "tmp.w", bias)
        fn, args, what, flops, *_ = hs.ops[0]
        for _ in range(2):
            _cabi.check(fn(*args, stream), what)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            _cabi.check(fn(*args, stream), what)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{mode:9s} {hw:4d}^2 {cin:4d}->{cout:4d}  {ms:8.3f} ms  nominal {flops / ms / 1e9:8.0f} TFLOP/s", flush=True)
        del hs, x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

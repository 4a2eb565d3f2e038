"""GPU bring-up: per-kernel checks against torch fp32 on the same fp16-rounded operands, then the
whole encoder/decoder against the CPU oracle. Prints one line per check; exits non-zero on failure."""
import ctypes as C
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from weatherforecastingtoolkit_b200 import _cabi
from weatherforecastingtoolkit_b200._cabi import ConvDesc, Tap
from weatherforecastingtoolkit_b200.engine import AKLEngine, _Program, _Act
from weatherforecastingtoolkit_b200.synthetic import PATHB_AKL_CONFIG, make_akl_state_dict, make_vil_sequences

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
lib = _cabi.init(0)
fails = 0


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def report(name, err, tol):
    global fails
    ok = err <= tol and err == err
    fails += 0 if ok else 1
    print(f"[{'OK' if ok else 'FAIL'}] {name}: rel_l2={err:.3e} (tol {tol:.1e})", flush=True)


class Harness(_Program):
    """A _Program shell that lets single layers be built and run."""

    def __init__(self, eng, n):
        self.eng = eng
        self.lib = eng.lib
        self.dev = eng.device
        from weatherforecastingtoolkit_b200.engine import _Pool
        self.pool = _Pool(self.dev)
        self.ops = []
        self.plans = []
        self.keep = []
        self.n = n
        self.stats_arena = torch.zeros(16, n, eng.groups, 2, dtype=torch.float64, device=self.dev)
        self._stats_used = 0

    def go(self):
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        self.stats_arena.zero_()
        for fn, args, what, *_ in self.ops:
            _cabi.check(fn(*args, stream), what)
        torch.cuda.synchronize()


def ref_stats(y_nhwc, groups):
    n, h, w, c = y_nhwc.shape
    g = y_nhwc.double().reshape(n, h * w, groups, c // groups)
    return torch.stack([g.sum(dim=(1, 3)), (g * g).sum(dim=(1, 3))], dim=-1)


def test_conv(eng, n, h, w, cin, cout, mode):
    torch.manual_seed(h * 1000 + cin + cout)
    x = torch.randn(n, h, w, cin, device=dev).half()
    wt = (torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(9 * cin))
    bias = torch.randn(cout, device=dev)
    hs = Harness(eng, n)
    xr = x.float().permute(0, 3, 1, 2)
    if mode == "plain" or mode == "residual":
        eng.w.t["tmp.w"] = wt.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous().half()
        res = torch.randn(n, h, w, cout, device=dev).half() if mode == "residual" else None
        out = hs.conv3x3(_Act(x, None), "tmp.w", bias, cout, residual=res)
        hs.go()
        ref = F.conv2d(xr, wt.half().float(), bias, padding=1).permute(0, 2, 3, 1)
        if res is not None:
            ref = ref + res.float()
    elif mode == "shortcut":
        cs = cin * 2
        xs = torch.randn(n, h, w, cs, device=dev).half()
        ws = torch.randn(cout, cs, 1, 1, device=dev) / math.sqrt(cs)
        eng.w.t["tmp.w"] = wt.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous().half()
        eng.w.t["tmp.sc"] = ws.reshape(1, cout, cs).contiguous().half()
        out = hs.conv3x3(_Act(x, None), "tmp.w", bias, cout, shortcut=(xs, "tmp.sc"))
        hs.go()
        ref = (F.conv2d(xr, wt.half().float(), bias, padding=1)
               + F.conv2d(xs.float().permute(0, 3, 1, 2), ws.half().float())).permute(0, 2, 3, 1)
    elif mode == "down":
        assert cin == cout
        eng.w.t["tmp.w"] = wt.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous().half()
        out = hs.downsample(_Act(x, None), "tmp.w", bias)
        hs.go()
        ref = F.conv2d(F.pad(xr, (0, 1, 0, 1)), wt.half().float(), bias, stride=2).permute(0, 2, 3, 1)
    elif mode == "up":
        assert cin == cout
        from weatherforecastingtoolkit_b200.engine import PackedAKL
        eng.w.t["tmp.w"] = PackedAKL._phase_weights(wt)
        out = hs.upsample(_Act(x, None), "tmp.w", bias)
        hs.go()
        ref = F.conv2d(F.interpolate(xr, scale_factor=2.0, mode="nearest"), wt, bias, padding=1).permute(0, 2, 3, 1)
    e = rel(out.t.float(), ref)
    report(f"conv[{mode}] n={n} {h}x{w} {cin}->{cout}", e, 3e-3)
    st = ref_stats(ref, eng.groups)
    es = rel(out.stats, st)
    report(f"  stats[{mode}]", es, 1e-3)


def main():
    cfg = PATHB_AKL_CONFIG
    sd = make_akl_state_dict(cfg, 0, affine_jitter=0.1)
    t0 = time.time()
    eng = AKLEngine(cfg, sd, device=dev)
    print(f"engine packed in {time.time() - t0:.1f}s", flush=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "conv"):
        test_conv(eng, 2, 16, 16, 128, 128, "plain")
        test_conv(eng, 1, 48, 48, 128, 256, "plain")
        test_conv(eng, 2, 24, 40, 256, 512, "residual")
        test_conv(eng, 1, 32, 32, 128, 256, "shortcut")
        test_conv(eng, 2, 32, 48, 128, 128, "down")
        test_conv(eng, 1, 24, 24, 256, 256, "up")
        test_conv(eng, 3, 96, 96, 512, 512, "residual")
    if which in ("all", "net"):
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
        from oracle import akl_oracle as O
        for H in (64, 96):
            u8 = make_vil_sequences(2, H, H, 2, seed=3)
            x = O.stage_vil(u8).permute(0, 3, 1, 2)[:, :1].contiguous()
            with torch.no_grad():
                mref = O.akl_encode_moments(x, sd, cfg)
                z = mref[:, :4].contiguous()
                dref = O.akl_decode(z, sd, cfg)
            m = eng.encode_moments(x.to(dev))
            torch.cuda.synchronize()
            report(f"encode {H}x{H} moments", rel(m.cpu(), mref), 1e-2)
            d = eng.decode(z.to(dev))
            torch.cuda.synchronize()
            report(f"decode {H}x{H}", rel(d.cpu(), dref), 1e-2)
    print("FAILS", fails)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()

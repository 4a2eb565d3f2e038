"""Direct launches of the two bandwidth-bound edge kernels at Path-B shapes (for ncu): encoder.conv_in (1 -> 128 @
384^2, tensor-core stem) and the decoder tail (GroupNorm + SiLU + conv 128 -> 1 @ 384^2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from weatherforecastingtoolkit_b200 import _cabi

n, hw, c = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 384, 128
lib = _cabi.init(0)
dev = "cuda:0"
st = torch.cuda.current_stream().cuda_stream
x = torch.rand(n, 1, hw, hw, device=dev)
w = (torch.randn(16, c, device=dev) / 3).half()
w[9:] = 0
b = torch.randn(c, device=dev)
out = torch.empty(n, hw, hw, c, dtype=torch.float16, device=dev)
stats = torch.zeros(n, 32, 2, dtype=torch.float64, device=dev)
tail_w = torch.randn(9, c, device=dev) / 30
gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
img = torch.empty(n, 1, hw, hw, device=dev)
for rep in range(3):
    stats.zero_()
    _cabi.check(lib.wfk_conv3x3_stem_tc(x.data_ptr(), n, 1, hw, hw, 0, w.data_ptr(), b.data_ptr(), c, out.data_ptr(),
                                        stats.data_ptr(), 4, 0, st), "stem")
    _cabi.check(lib.wfk_gn_silu_conv3x3_c1(out.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), n, hw, hw,
                                           c, 32, 1e-6, tail_w.data_ptr(), 0.1, img.data_ptr(), 0, st), "tail")
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
_cabi.check(lib.wfk_conv3x3_stem_tc(x.data_ptr(), n, 1, hw, hw, 0, w.data_ptr(), b.data_ptr(), c, out.data_ptr(),
                                    stats.data_ptr(), 4, 0, st), "stem")
e1.record()
_cabi.check(lib.wfk_gn_silu_conv3x3_c1(out.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), n, hw, hw, c,
                                       32, 1e-6, tail_w.data_ptr(), 0.1, img.data_ptr(), 0, st), "tail")
e2.record()
torch.cuda.synchronize()
print(f"n={n}: stem {e0.elapsed_time(e1) * 1e3 / n:.1f} us/frame, tail {e1.elapsed_time(e2) * 1e3 / n:.1f} us/frame")

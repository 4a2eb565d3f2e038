"""Probe: which C-ABI launches can be captured into a CUDA graph (one graph per op of an AE_ViT_2048 program)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from weatherforecastingtoolkit_b200 import _cabi
from weatherforecastingtoolkit_b200 import synthetic as S
from weatherforecastingtoolkit_b200.models.ae_vit import AE_ViT_2048

os.environ["WFK_CUDA_GRAPH"] = "0"
dev = torch.device("cuda:0")
m = AE_ViT_2048().eval()
m.load_state_dict(S.fill_state_dict(m, "vit", 0))
m = m.to(dev)
x = torch.rand(8, 1, 128, 128, device=dev)
with torch.no_grad():
    m(x)
prog = next(iter(m._programs.values())) if hasattr(m, "_programs") else None
if prog is None:
    for k, v in vars(m).items():
        if isinstance(v, dict) and v and hasattr(next(iter(v.values())), "ops"):
            prog = next(iter(v.values()))
print("ops:", len(prog.ops))
seen = set()
for fn, args, what, *_ in prog.ops:
    name = fn.__name__ if hasattr(fn, "__name__") else str(fn)
    if name in seen:
        continue
    seen.add(name)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            st = torch.cuda.current_stream().cuda_stream
            rc = fn(*args, st)
        g.replay()
        torch.cuda.synchronize()
        print("ok  ", name, what, "rc", rc)
    except Exception as e:
        print("FAIL", name, what, str(e).splitlines()[0][:100])
        try:
            torch.cuda.synchronize()
        except Exception as e2:
            print("   sync:", str(e2)[:80])

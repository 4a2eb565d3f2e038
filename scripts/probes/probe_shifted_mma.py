"""Probe: shifted-window UMMA A descriptors on a TMA-written halo tile (scripts/probes/debug_mma.cu).
Not part of libwfk_b200.so: the probe kernel is built here into its own library together with csrc/core.cu."""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import torch

from weatherforecastingtoolkit_b200 import _cabi

CSRC = os.path.join(ROOT, "weatherforecastingtoolkit_b200", "csrc")
SO = os.path.join(HERE, "libwfk_probe.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                       "-cudart", "static", "-I", CSRC, "-I", os.path.join(ROOT, "include"), "-o", SO,
                       os.path.join(HERE, "debug_mma.cu"), os.path.join(CSRC, "core.cu")])
lib = C.CDLL(SO)
lib.wfk_init.argtypes = [C.c_int]
assert lib.wfk_init(0) == 0
fn = lib.wfk_debug_shifted_mma
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
torch.manual_seed(0)
dev = "cuda:0"
for pitch in (16, 10):
    rows = 18
    a = torch.randn(rows, pitch, 64, device=dev).half()
    b = torch.randn(128, 64, device=dev).half()
    for (r, s) in [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 2), (1, 2)]:
        # A row m = 8*g + i  ->  halo pixel (g + r, i + s)
        idx_y = torch.arange(128, device=dev) // 8 + r
        idx_x = torch.arange(128, device=dev) % 8 + s
        a_rows = a[idx_y, idx_x].float()
        ref = a_rows @ b.float().t()
        res = []
        for bo_name, bo in (("0", 0), ("s", s), ("row&7", (r * pitch + s) & 7)):
            d = torch.zeros(128, 128, device=dev)
            _cabi.check(fn(a.data_ptr(), pitch, rows, b.data_ptr(), r, s, bo, d.data_ptr(), None), "probe")
            torch.cuda.synchronize()
            err = ((d - ref).norm() / ref.norm()).item()
            res.append(f"base_offset={bo_name}:{err:.2e}")
        print(f"pitch={pitch} tap=({r},{s})  " + "  ".join(res), flush=True)

"""Build libwfk_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Run as ``python -m weatherforecastingtoolkit_b200.build`` or through ``__graft_entry__.build()``.
The shared library has no PyTorch dependency: it is the C ABI declared in ``include/wfk_b200.h``.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libwfk_b200.so"
STAMP = PKG / "csrc" / ".build_stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [PKG.parent / "include" / "wfk_b200.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build_experiments(verbose: bool = False) -> Path:
    """libwfk_b200_exp.so: the same sources with -DWFK_EXPERIMENTS (timing experiments that produce WRONG results by
    design: operand traffic removed, stores skipped, fewer CTA pairs). Never loaded by the package; select it for a
    scripts/ run with WFK_LIB_PATH."""
    nvcc = _nvcc()
    objdir = PKG / "build_exp"
    objdir.mkdir(exist_ok=True)
    out = PKG / "libwfk_b200_exp.so"
    procs, objs = [], []
    for src in _sources():
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-DWFK_EXPERIMENTS", "-I", str(PKG.parent / "include"), "-c", str(src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, pr in procs:
        o, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{o}")
    r = subprocess.run([nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                        "-o", str(out), *objs], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return out


def build(force: bool = False, verbose: bool = False) -> Path:
    dig = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == dig:
        return LIB
    nvcc = _nvcc()
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in _sources():
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(PKG.parent / "include"), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose and out:
            print(out)
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{out}")
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-o", str(LIB), *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    STAMP.write_text(dig)
    return LIB


if __name__ == "__main__":
    if "--experiments" in sys.argv:
        print(build_experiments(verbose="-v" in sys.argv))
    else:
        p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
        print(p)

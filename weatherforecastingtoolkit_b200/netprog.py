"""Static kernel programs for the auxiliary model families (PosAwareAE_TF, NLayerDiscriminator, AE_ViT_2048).

A ``NetProgram`` is a list of (C function, args) bound to preallocated device buffers for one input shape,
like ``engine._Program`` for the AutoencoderKL. Every arithmetic op is a kernel of libwfk_b200.so reached
through the C ABI; PyTorch only owns the buffers and the stream. Activations are NHWC fp16 (rows of a
[M, K] matrix for the GEMM-shaped layers).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _cabi, engine
from ._cabi import ConvDesc, Tap

F16 = torch.float16


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def pad_k(w: torch.Tensor, mult: int = 8) -> torch.Tensor:
    """Zero-pad the last (K) dimension to a multiple of ``mult`` elements (TMA needs 16-byte row pitches)."""
    k = w.shape[-1]
    kp = (k + mult - 1) // mult * mult
    if kp == k:
        return w.contiguous()
    out = torch.zeros(*w.shape[:-1], kp, dtype=w.dtype, device=w.device)
    out[..., :k] = w
    return out


class NetProgram:
    def __init__(self, device):
        self.dev = torch.device(device)
        self.lib = _cabi.init(self.dev.index if self.dev.index is not None else 0)
        self.pool = engine._Pool(self.dev)
        self.ops: List[Tuple[Callable, tuple, str, float]] = []
        self.plans: List = []
        self.keep: List = []
        # These networks are ~100-250 short launches per forward (3 ms for a 64-frame ViT batch): launch-bound. After
        # one eager run (which also sets the kernels' shared-memory attributes) the whole op list is captured into a
        # CUDA graph and replayed; WFK_CUDA_GRAPH=0 keeps the eager path (A/B switch, identical results).
        import os
        self.use_graph = os.environ.get("WFK_CUDA_GRAPH", "1") != "0"
        self._graph = None
        self._eager_runs = 0

    def __del__(self):
        try:
            for p in self.plans:
                self.lib.wfk_conv_plan_destroy(p)
        except Exception:
            pass

    # ------------------------------------------------------------------ running
    def add(self, fn, args, what, flops=0.0):
        self.ops.append((fn, tuple(args), what, float(flops)))

    def run(self):
        if engine.TIMER is None and self.use_graph:
            if self._graph is not None:
                self._graph.replay()
                return
            if self._eager_runs >= 1:
                try:
                    torch.cuda.synchronize(self.dev)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._run_eager()
                    self._graph = g
                    g.replay()
                    return
                except Exception as e:   # capture is an optimisation: fall back to the eager launches
                    import warnings
                    warnings.warn(f"CUDA graph capture failed ({e}); running eagerly")
                    self.use_graph = False
                    torch.cuda.synchronize(self.dev)
        self._eager_runs += 1
        self._run_eager()

    def _run_eager(self):
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        timer = engine.TIMER
        for fn, args, what, flops in self.ops:
            if timer is not None and flops:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                _cabi.check(fn(*args, stream), what)
                ev1.record()
                timer.record(what, flops, ev0, ev1)
            else:
                _cabi.check(fn(*args, stream), what)

    def buf(self, shape: Sequence[int], dtype=F16) -> torch.Tensor:
        return self.pool.get(shape, dtype)

    def free(self, *ts: Optional[torch.Tensor]):
        for t in ts:
            if t is not None:
                self.pool.put(t)

    # ------------------------------------------------------------------ conv-GEMM plans
    @staticmethod
    def _view_nhwc(v, t: torch.Tensor, n, h, w, c, pitch_c=None):
        pitch = c if pitch_c is None else pitch_c
        v.ptr = t.data_ptr()
        dims = (c, w, 1, h, n)
        strides = (2, pitch * 2, w * pitch * 2, w * pitch * 2, h * w * pitch * 2)
        for i in range(5):
            v.dim[i] = dims[i]
            v.stride[i] = strides[i]

    @staticmethod
    def _view_w(v, t: torch.Tensor):
        """t: [slabs, nrows, k] fp16 contiguous."""
        slabs, nrows, k = t.shape
        v.ptr = t.data_ptr()
        dims = (k, nrows, slabs)
        strides = (2, k * 2, nrows * k * 2)
        for i in range(3):
            v.dim[i] = dims[i]
            v.stride[i] = strides[i]

    def _finish(self, d: ConvDesc, *, bias, residual, out, out_f, out2, scale2, shift2, act, act2, slope, rows, cols,
                ldc, sy=1, sx=1, what="conv", flops=0.0):
        d.bias = _ptr(bias)
        d.residual = _ptr(residual)
        d.out_h = _ptr(out)
        d.out_f = _ptr(out_f)
        d.stats = None
        d.out_rows, d.out_cols, d.out_sy, d.out_sx, d.ldc = rows, cols, sy, sx, ldc
        d.cpg = 0
        d.operand_bf16 = 0
        d.gn_table = None
        d.act, d.act2, d.act_slope = act, act2, slope
        d.out2_h = _ptr(out2)
        d.scale2 = _ptr(scale2)
        d.shift2 = _ptr(shift2)
        h = C.c_void_p()
        _cabi.check(self.lib.wfk_conv_plan_create(C.byref(d), C.byref(h)), f"conv_plan_create[{what}]")
        self.plans.append(h)
        self.add(self.lib.wfk_conv_plan_run, (h,), what, flops)

    def conv_s1(self, x: torch.Tensor, w: torch.Tensor, k: int, pad: int, *, bias=None, residual=None, want_out=True,
                out2=False, scale2=None, shift2=None, act=0, act2=0, slope=0.0, out_f32=False, what="conv"):
        """k x k stride-1 convolution (zero padding ``pad``) of x [n, h, w, cin] with w [k*k, cout, cin_padded]
        (slab = r*k + s). Returns (out, out2): fp16 [n, oh, ow, cout] (or fp32 when ``out_f32``) / None."""
        n, h, wd, cin = x.shape
        cout = w.shape[1]
        oh, ow = h + 2 * pad - k + 1, wd + 2 * pad - k + 1
        d = ConvDesc()
        self._view_nhwc(d.a[0], x, n, h, wd, cin)
        self._view_w(d.b[0], w)
        d.n_frames, d.tile_h, d.tile_w, d.n_total = n, oh, ow, cout
        d.num_phases, d.taps_per_phase = 1, k * k
        kb = (cin + 63) // 64
        for r in range(k):
            for s in range(k):
                d.taps[r * k + s] = Tap(s - pad, r - pad, 0, 0, 0, r * k + s, kb, 0)
        d.a_frame_mul, d.b_frame_mul = 1, 0
        o = (self.buf((n, oh, ow, cout), torch.float32 if out_f32 else F16)) if want_out else None
        o2 = self.buf((n, oh, ow, cout)) if out2 else None
        self._finish(d, bias=bias, residual=residual, out=None if out_f32 else o, out_f=o if out_f32 else None, out2=o2,
                     scale2=scale2, shift2=shift2, act=act, act2=act2, slope=slope, rows=oh, cols=ow, ldc=cout,
                     what=what, flops=2.0 * n * oh * ow * cout * k * k * cin)
        return o, o2

    def conv4x4_s2(self, x: torch.Tensor, w: torch.Tensor, *, bias=None, want_out=True, out2=False, scale2=None,
                   shift2=None, act=0, act2=0, slope=0.0, what="conv4x4s2"):
        """Conv2d(cin, cout, 4, stride 2, padding 1) through the parity view (2C, W/2, 2, H/2, N) of x: input pixel
        (2y-1+r, 2x-1+s) is channel block (s+1)%2, column x + (s-1)//2... of row parity (r+1)%2. w [16, cout, cin]."""
        n, h, wd, c = x.shape
        if h % 2 or wd % 2:
            raise ValueError("4x4 stride-2 convolution needs even H and W")
        if c % 8:
            raise ValueError("channel count must be a multiple of 8")
        cout = w.shape[1]
        oh, ow = h // 2, wd // 2
        d = ConvDesc()
        v = d.a[0]
        v.ptr = x.data_ptr()
        dims = (2 * c, ow, 2, oh, n)
        strides = (2, 2 * c * 2, wd * c * 2, 2 * wd * c * 2, h * wd * c * 2)
        for i in range(5):
            v.dim[i] = dims[i]
            v.stride[i] = strides[i]
        self._view_w(d.b[0], w)
        d.n_frames, d.tile_h, d.tile_w, d.n_total = n, oh, ow, cout
        d.num_phases, d.taps_per_phase = 1, 16
        kb = (c + 63) // 64
        if c % 64 and kb * 64 > c:
            # a K block would run into the neighbouring pixel's channels of the parity view
            raise ValueError("4x4 stride-2 convolution needs a multiple of 64 input channels")
        for r in range(4):
            for s in range(4):
                yy, xx = r - 1, s - 1            # offset from (2y, 2x)
                d.taps[r * 4 + s] = Tap(xx // 2, yy // 2, yy % 2, 0, (xx % 2) * c, r * 4 + s, kb, 0)
        d.a_frame_mul, d.b_frame_mul = 1, 0
        o = self.buf((n, oh, ow, cout)) if want_out else None
        o2 = self.buf((n, oh, ow, cout)) if out2 else None
        self._finish(d, bias=bias, residual=None, out=o, out_f=None, out2=o2, scale2=scale2, shift2=shift2, act=act,
                     act2=act2, slope=slope, rows=oh, cols=ow, ldc=cout, what=what,
                     flops=2.0 * n * oh * ow * cout * 16 * c)
        return o, o2

    def convT4x4_s2(self, x: torch.Tensor, w_phase: torch.Tensor, *, bias=None, want_out=True, out2=False, scale2=None,
                    shift2=None, act=0, act2=0, slope=0.0, what="convT4x4s2"):
        """ConvTranspose2d(cin, cout, 4, stride 2, padding 1) as four 2x2 sub-pixel convolutions writing interleaved
        output phases. w_phase [16 = (a, b, i, j), cout, cin] from ``pack_convT4x4``."""
        n, h, wd, c = x.shape
        cout = w_phase.shape[1]
        d = ConvDesc()
        self._view_nhwc(d.a[0], x, n, h, wd, c)
        self._view_w(d.b[0], w_phase)
        d.n_frames, d.tile_h, d.tile_w, d.n_total = n, h, wd, cout
        d.num_phases, d.taps_per_phase = 4, 4
        kb = (c + 63) // 64
        offs = {0: (-1, 0), 1: (0, 1)}
        for a in (0, 1):
            for b in (0, 1):
                ph = a * 2 + b
                for i in (0, 1):
                    for j in (0, 1):
                        d.taps[ph * 4 + i * 2 + j] = Tap(offs[b][j], offs[a][i], 0, 0, 0, ph * 4 + i * 2 + j, kb, 0)
        d.a_frame_mul, d.b_frame_mul = 1, 0
        o = self.buf((n, 2 * h, 2 * wd, cout)) if want_out else None
        o2 = self.buf((n, 2 * h, 2 * wd, cout)) if out2 else None
        self._finish(d, bias=bias, residual=None, out=o, out_f=None, out2=o2, scale2=scale2, shift2=shift2, act=act,
                     act2=act2, slope=slope, rows=2 * h, cols=2 * wd, ldc=cout, sy=2, sx=2, what=what,
                     flops=2.0 * n * h * wd * cout * 16 * c)
        return o, o2

    def gemm(self, x: torch.Tensor, w: torch.Tensor, *, bias=None, residual=None, out_f32=False, act=0, slope=0.0,
             out: Optional[torch.Tensor] = None, ldc: Optional[int] = None, x_pitch: Optional[int] = None,
             k: Optional[int] = None, want_out=True, out2=False, scale2=None, shift2=None, act2=0, what="gemm"):
        """out[M, N] = act(x[M, K] @ w[N, K]^T + bias (+ residual)). x fp16 rows (pitch ``x_pitch`` elements),
        w [1, N, Kp] fp16. ``out`` / ``ldc`` let the result land inside a wider row-major tensor."""
        m = x.shape[0]
        kk = int(k if k is not None else x.shape[1])
        nn_ = w.shape[1]
        d = ConvDesc()
        self._view_nhwc(d.a[0], x, 1, 1, m, kk, pitch_c=x_pitch if x_pitch is not None else x.shape[1])
        self._view_w(d.b[0], w)
        d.n_frames, d.tile_h, d.tile_w, d.n_total = 1, 1, m, nn_
        d.num_phases, d.taps_per_phase = 1, 1
        d.taps[0] = Tap(0, 0, 0, 0, 0, 0, (kk + 63) // 64, 0)
        d.a_frame_mul, d.b_frame_mul = 1, 0
        if out is None and want_out:
            out = self.buf((m, nn_), torch.float32 if out_f32 else F16)
        o2 = self.buf((m, nn_)) if out2 else None
        self._finish(d, bias=bias, residual=residual, out=None if out_f32 else out, out_f=out if out_f32 else None,
                     out2=o2, scale2=scale2, shift2=shift2, act=act, act2=act2, slope=slope, rows=1, cols=m,
                     ldc=ldc if ldc is not None else nn_, what=what, flops=2.0 * m * nn_ * kk)
        return (out, o2) if out2 else out


# ---------------------------------------------------------------------------------------------- weight packing
def fold_bn(w: torch.Tensor, bias: Optional[torch.Tensor], bn_w, bn_b, bn_mean, bn_var, eps: float, out_dim: int = 0):
    """Fold an eval-mode BatchNorm that FOLLOWS a convolution into its weight / bias (fp32):
    y = gamma * (conv(x) + b - mean) / sqrt(var + eps) + beta."""
    scale = bn_w / torch.sqrt(bn_var + eps)
    shape = [1] * w.ndim
    shape[out_dim] = -1
    wf = w * scale.view(shape)
    b0 = bias if bias is not None else torch.zeros_like(bn_mean)
    bf = (b0 - bn_mean) * scale + bn_b
    return wf, bf


def bn_affine(bn_w, bn_b, bn_mean, bn_var, eps: float):
    """(scale, shift) of an eval-mode BatchNorm as a per-channel affine map."""
    scale = bn_w / torch.sqrt(bn_var + eps)
    return scale.contiguous(), (bn_b - bn_mean * scale).contiguous()


def pack_conv(w: torch.Tensor) -> torch.Tensor:
    """Conv2d weight [cout, cin, k, k] -> [k*k, cout, cin_padded] fp16 (slab = r*k + s)."""
    cout, cin, k, _ = w.shape
    return pad_k(w.permute(2, 3, 0, 1).reshape(k * k, cout, cin)).to(F16).contiguous()


def pack_convT4x4(w: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d weight [cin, cout, 4, 4] (stride 2, padding 1) -> [16 = (a, b, i, j), cout, cin] fp16.
    Output row 2m+a receives input rows (m-1, m) with kernel rows (3, 1) for a = 0 and rows (m, m+1) with
    kernel rows (2, 0) for a = 1; columns likewise."""
    cin, cout = w.shape[0], w.shape[1]
    ksel = {0: (3, 1), 1: (2, 0)}
    out = torch.empty(2, 2, 2, 2, cout, cin, dtype=torch.float32, device=w.device)
    for a in (0, 1):
        for b in (0, 1):
            for i in (0, 1):
                for j in (0, 1):
                    out[a, b, i, j] = w[:, :, ksel[a][i], ksel[b][j]].t()
    return pad_k(out.reshape(16, cout, cin)).to(F16).contiguous()


def pack_grouped3x3(w: torch.Tensor, groups: int) -> torch.Tensor:
    """Grouped Conv2d weight [cout, cin/groups, 3, 3] -> dense block-diagonal [9, cout, cin_padded] fp16."""
    cout, cpg, k, _ = w.shape
    cin = cpg * groups
    opg = cout // groups
    dense = torch.zeros(cout, cin, k, k, dtype=w.dtype, device=w.device)
    for g in range(groups):
        dense[g * opg:(g + 1) * opg, g * cpg:(g + 1) * cpg] = w[g * opg:(g + 1) * opg]
    return pack_conv(dense)


def pack_linear(w: torch.Tensor) -> torch.Tensor:
    """[N, K] -> [1, N, K_padded] fp16."""
    return pad_k(w).to(F16).unsqueeze(0).contiguous()

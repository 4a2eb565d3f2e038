"""Host-side executor of the Path-B autoencoder on libwfk_b200.so.

PyTorch is used for device memory and streams only; every arithmetic op below is one of the
hand-written sm_100a kernels reached through the C ABI (``_cabi``). Activations are NHWC fp16 on
device; weights are packed once (fp16, ``[slab][cout][cin]``) from a reference ``state_dict``.

The layer sequence follows the reference modules (paths relative to the reference repo):
``Encoder.forward`` (pipeline/models/autoencoderkl/vae.py:70-86), ``Decoder.forward``
(vae.py:150-166), ``ResnetBlock2D.forward`` (resnet.py:454-495), ``Downsample2D`` /
``Upsample2D`` (resnet.py:181-190, 108-143), ``AttentionBlock.forward`` (attention.py:136-189),
``AutoencoderKL.encode/_decode`` (autoencoder_kl.py:80-89).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _cabi
from ._cabi import ConvDesc, Tap

F16 = torch.float16
GN_EPS = 1e-6  # resnet_eps passed by Encoder/Decoder (vae.py:41,57,128)


class KernelTimer:
    """Optional per-launch CUDA-event timing (bench.py's roofline legs). Events are recorded on the launching stream
    around every conv-GEMM launch and around the HBM-bound passes (staging, predictor, metrics); ``summary()`` /
    ``secondary()`` must be called after a synchronize. ``flops`` are NOMINAL (2*M*N*K of the reference layer);
    ``executed`` is the MMA work the kernel really issues (the 4-phase sub-pixel upsample does 2.25x fewer MACs than
    the nominal 3x3 convolution on the upsampled tensor, the identity-tap residual adds some)."""

    def __init__(self, all_ops: bool = False):
        self.records = []  # (what, nominal_flops, start_event, end_event)
        self.executed = []  # executed flops, parallel to records
        self.all_ops = all_ops  # also time the non-GEMM kernels (bench.py --breakdown)
        self.passes = []   # (name, algorithmic_bytes, start_event, end_event): HBM-bound passes

    def record(self, what, flops, start, end, executed=None):
        self.records.append((what, flops, start, end))
        self.executed.append(flops if executed is None else executed)

    def record_pass(self, name, nbytes, start, end):
        self.passes.append((name, nbytes, start, end))

    def summary(self):
        tot_ms, tot_flops, tot_exec, n = 0.0, 0.0, 0.0, 0
        for (_, fl, s, e), ex in zip(self.records, self.executed):
            if not fl:
                continue
            tot_ms += s.elapsed_time(e)
            tot_flops += fl
            tot_exec += ex
            n += 1
        return {"launches": n, "ms": tot_ms, "nominal_flops": tot_flops, "executed_flops": tot_exec}

    def secondary(self):
        """{pass name: {launches, ms, bytes}} of the HBM-bound passes."""
        out = {}
        for name, nbytes, s, e in self.passes:
            d = out.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += s.elapsed_time(e)
            d["bytes"] += nbytes
        return out


class timed_pass:
    """``with timed_pass(name, algorithmic_bytes):`` around an HBM-bound launch; a no-op unless bench.py set TIMER."""

    def __init__(self, name: str, nbytes: float):
        self.name, self.nbytes, self.timer = name, nbytes, TIMER

    def __enter__(self):
        if self.timer is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.timer is not None:
            self.e1.record()
            self.timer.record_pass(self.name, self.nbytes, self.e0, self.e1)
        return False


TIMER: Optional[KernelTimer] = None  # set by bench.py


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class _Pool:
    """Free-list of device buffers keyed by (numel, dtype): plans bind raw pointers, so buffers
    are handed out at plan-build time following the static liveness of the layer sequence."""

    def __init__(self, device, adt: torch.dtype = F16):
        self.device = device
        self.adt = adt   # default element type of activations
        self.free: Dict[Tuple[int, torch.dtype], List[torch.Tensor]] = {}
        self.all: List[torch.Tensor] = []

    def get(self, shape: Sequence[int], dtype=None) -> torch.Tensor:
        dtype = self.adt if dtype is None else dtype
        n = 1
        for s in shape:
            n *= int(s)
        lst = self.free.get((n, dtype))
        if lst:
            return lst.pop().view(*shape)
        t = torch.empty(n, dtype=dtype, device=self.device)
        self.all.append(t)
        return t.view(*shape)

    def put(self, t: torch.Tensor) -> None:
        self.free.setdefault((t.numel(), t.dtype), []).append(t.reshape(-1))


class _Act:
    """An NHWC fp16 activation plus the GroupNorm statistics its producer accumulated."""

    def __init__(self, t: torch.Tensor, stats: Optional[torch.Tensor]):
        self.t = t          # [N, H, W, C]
        self.stats = stats  # [N, 32, 2] float64 or None

    @property
    def shape(self):
        return self.t.shape


class PackedAKL:
    """fp16 / fp32 device copies of the AutoencoderKL weights in kernel layouts."""

    def __init__(self, cfg: dict, sd: Dict[str, torch.Tensor], device, adt: torch.dtype = torch.float16):
        self.cfg = dict(cfg)
        self.device = device
        self.adt = adt      # 16-bit operand / activation format: torch.float16 (default) or torch.bfloat16
        F16 = adt           # noqa: N806 -- every ".to(F16)" below packs into the selected format
        self.t: Dict[str, torch.Tensor] = {}
        boc = list(cfg["block_out_channels"])
        self.groups = int(cfg.get("norm_num_groups", 32))
        for name, w in sd.items():
            w = w.detach().to(device=device, dtype=torch.float32)
            if name.endswith(".weight") and w.ndim == 4 and w.shape[2] == 3:
                mod = name[: -len(".weight")]
                cout, cin = w.shape[0], w.shape[1]
                if mod == "encoder.conv_in" or mod == "decoder.conv_in":
                    # direct kernel: [cin*9][cout] fp32
                    self.t[mod + ".w_direct"] = w.permute(1, 2, 3, 0).reshape(cin * 9, cout).contiguous()
                    # tensor-core stem kernel: [K padded to 16][cout] fp16, row k = ci*9 + tap. For the decoder the
                    # 1x1 post_quant_conv in front (autoencoder_kl.py:87) is folded in (fp32): W_eff[:, cj] = sum_ci
                    # W[:, ci] pq_w[ci, cj]; its bias rides on a constant-one plane (present only where a tap is inside
                    # the image, exactly like the reference's zero padding AFTER post_quant_conv)
                    wk = w
                    if mod == "decoder.conv_in" and "post_quant_conv.weight" in sd:
                        pw = sd["post_quant_conv.weight"].detach().to(device=device, dtype=torch.float32)
                        pw = pw.reshape(pw.shape[0], pw.shape[1])
                        pb = sd["post_quant_conv.bias"].detach().to(device=device, dtype=torch.float32)
                        wk = torch.cat([torch.einsum("oirs,ij->ojrs", w, pw), torch.einsum("oirs,i->ors", w, pb).unsqueeze(1)], 1)
                    kk = wk.shape[1] * 9
                    if kk <= 48:
                        wp = torch.zeros((kk + 15) // 16 * 16, cout, dtype=torch.float32, device=device)
                        wp[:kk] = wk.permute(1, 2, 3, 0).reshape(kk, cout)
                        self.t[mod + ".w_tc"] = wp.to(F16).contiguous()
                elif mod == "encoder.conv_out" or mod == "decoder.conv_out":
                    # direct kernel: [cout][9][cin] fp16
                    self.t[mod + ".w_direct"] = w.permute(0, 2, 3, 1).reshape(cout, 9, cin).contiguous().to(F16)
                    if mod == "encoder.conv_out" and "quant_conv.weight" in sd:
                        # quant_conv (1x1, autoencoder_kl.py:82) folded behind conv_out in fp32:
                        # Wq (Wc * x + bc) + bq = (Wq Wc) * x + (Wq bc + bq); rows padded to a multiple of 8
                        wq = sd["quant_conv.weight"].detach().to(device=device, dtype=torch.float32)
                        wq = wq.reshape(wq.shape[0], wq.shape[1])
                        bq = sd["quant_conv.bias"].detach().to(device=device, dtype=torch.float32)
                        bc = sd["encoder.conv_out.bias"].detach().to(device=device, dtype=torch.float32)
                        wf = torch.einsum("oc,cirs->oirs", wq, w)
                        npad = (cout + 7) // 8 * 8
                        wp = torch.zeros(9, npad, cin, dtype=torch.float32, device=device)
                        wp[:, :cout] = wf.permute(2, 3, 0, 1).reshape(9, cout, cin)
                        bp = torch.zeros(npad, dtype=torch.float32, device=device)
                        bp[:cout] = wq @ bc + bq
                        self.t[mod + ".w_folded"] = wp.to(F16).contiguous()
                        self.t[mod + ".bias_folded"] = bp.contiguous()
                    if cout == 1:  # fused GroupNorm+SiLU+conv tail: [9][cin] fp32
                        self.t[mod + ".w_tap"] = w.permute(0, 2, 3, 1).reshape(9, cin).contiguous()
                elif ".upsamplers." in mod:
                    self.t[mod + ".w_phase"] = self._phase_weights(w, adt)
                else:
                    # [9][cout][cin] fp16, slab = r*3+s
                    self.t[mod + ".w"] = w.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous().to(F16)
            elif name.endswith(".weight") and w.ndim == 4 and w.shape[2] == 1:
                mod = name[: -len(".weight")]
                if mod in ("quant_conv", "post_quant_conv"):
                    self.t[mod + ".w"] = w.reshape(w.shape[0], w.shape[1]).contiguous()
                else:  # conv_shortcut
                    self.t[mod + ".w"] = w.reshape(1, w.shape[0], w.shape[1]).contiguous().to(F16)
            elif name.endswith(".weight") and w.ndim == 2:  # attention linears
                self.t[name[: -len(".weight")] + ".w"] = w.contiguous().to(F16)
            else:
                self.t[name] = w.contiguous()
        # fused parameters
        for p in ("encoder.mid_block.attentions.0", "decoder.mid_block.attentions.0"):
            if p + ".query.w" in self.t:
                self.t[p + ".qk.w"] = torch.cat([self.t[p + ".query.w"], self.t[p + ".key.w"]], 0).unsqueeze(0).contiguous()
                self.t[p + ".qk.bias"] = torch.cat([self.t[p + ".query.bias"], self.t[p + ".key.bias"]], 0).contiguous()
                self.t[p + ".proj_attn.w3"] = self.t[p + ".proj_attn.w"].unsqueeze(0).contiguous()
        for k in list(self.t):
            if k.endswith(".conv_shortcut.w"):
                r = k[: -len(".conv_shortcut.w")]
                self.t[r + ".conv2.bias_sc"] = (self.t[r + ".conv2.bias"] + self.t[r + ".conv_shortcut.bias"]).contiguous()

    @staticmethod
    def _phase_weights(w: torch.Tensor, adt: torch.dtype = F16) -> torch.Tensor:
        """nearest-x2 upsample + conv3x3 == four 2x2 convolutions on the low-res input, one per
        output parity (a, b): rows {2y+a-1, 2y+a, 2y+a+1} // 2 collapse onto two source rows, so the
        3x3 weights are pre-summed. Returns [16 slabs = (a,b,i,j)][cout][cin] fp16."""
        cout, cin = w.shape[0], w.shape[1]
        sets = {0: [[0], [1, 2]], 1: [[0, 1], [2]]}
        out = torch.zeros(2, 2, 2, 2, cout, cin, dtype=torch.float32, device=w.device)
        for a in (0, 1):
            for b in (0, 1):
                for i, rs in enumerate(sets[a]):
                    for j, ss in enumerate(sets[b]):
                        acc = torch.zeros(cout, cin, dtype=torch.float32, device=w.device)
                        for r in rs:
                            for s in ss:
                                acc = acc + w[:, :, r, s]
                        out[a, b, i, j] = acc
        return out.reshape(16, cout, cin).contiguous().to(adt)


# rows touched by upsample phase a: a=0 -> (y-1, y); a=1 -> (y, y+1)
_PHASE_OFFS = {0: (-1, 0), 1: (0, 1)}


class AKLEngine:
    """Builds (once per input shape) and runs the kernel sequence of encode / decode."""

    def __init__(self, cfg: dict, sd: Dict[str, torch.Tensor], device="cuda:0", operand_bf16: bool = False):
        self.device = torch.device(device)
        dev_index = self.device.index if self.device.index is not None else 0
        self.lib = _cabi.init(dev_index)
        import os
        # Operand / activation format. fp16 is the default (it meets the 1e-2 relative-L2 parity gate with ~3x margin);
        # bf16 -- the north-star format: same tensor-core rate and bytes, fp32 exponent range, 3 fewer mantissa bits --
        # is selected with operand_bf16=True or WFK_OPERANDS=bf16 (measured parity: DESIGN.md section 2).
        self.bf16 = bool(operand_bf16) or os.environ.get("WFK_OPERANDS", "fp16").lower() == "bf16"
        self.adt = torch.bfloat16 if self.bf16 else torch.float16
        self.dev_index = dev_index
        self.cfg = dict(cfg)
        self.boc = list(cfg["block_out_channels"])
        self.lpb = int(cfg.get("layers_per_block", 1))
        self.lc = int(cfg.get("latent_channels", 4))
        self.groups = int(cfg.get("norm_num_groups", 32))
        self.in_ch = int(cfg.get("in_channels", 3))
        self.out_ch = int(cfg.get("out_channels", 3))
        for c in self.boc:
            if c % 64 != 0 or (c // self.groups) not in (4, 8, 16):
                raise ValueError(f"block_out_channels={self.boc} unsupported: channels must be multiples of 64 with "
                                 f"4, 8 or 16 channels per GroupNorm group")
        if self.in_ch > 4 or self.lc > 4 or self.out_ch not in (1, 2, 4, 8) or 2 * self.lc not in (2, 4, 8):
            raise ValueError("unsupported in/out/latent channel counts for the direct edge-conv kernels")
        self.w = PackedAKL(cfg, sd, self.device, self.adt)
        # fused GroupNorm needs the HALO conv path (CTA pairs); both can be switched off for A/B runs
        self.fuse_gn = (os.environ.get("WFK_FUSE_GN", "1") != "0" and os.environ.get("WFK_CONV_HALO", "1") != "0"
                        and os.environ.get("WFK_CONV_PAIR", "1") != "0")
        # (scale, shift) derived inside the conv kernel instead of a wfk_gn_table launch. Off by default: once the MMA
        # issue loop stopped being the bottleneck the halo transform became co-critical, and the per-K-block fp64
        # mean / rstd arithmetic in its warps cost 17 % of the step (233 vs 280 frames/s)
        self.gn_inline = os.environ.get("WFK_GN_INLINE", "0") != "0"
        self.res_as_mma = os.environ.get("WFK_RES_AS_MMA", "1") != "0"
        self.stem_tc = os.environ.get("WFK_STEM_TC", "1") != "0"   # tensor-core stem kernels (A/B switch)
        # 1: three-pass softmax attention without the fp32 score matrix (see _Program.attention); measured equal to the
        # default chain within noise on B200, so it stays an opt-in
        self.attn_fused = os.environ.get("WFK_ATTN_FUSED", "0") == "1"
        self.attn_group = int(os.environ.get("WFK_ATTN_GROUP", "0"))   # frames per attention group (0: all frames at once)
        if self.bf16 and not (self.fuse_gn and self.stem_tc):
            raise ValueError("bf16 operands need the default kernel set (fused GroupNorm, tensor-core stems)")
        self._plans: Dict[Tuple, "_Program"] = {}
        self._keep: List = []

    def identity_weight(self, c: int) -> str:
        """Name of a [1][c][c] fp16 identity matrix in the packed-weight table (residual add as a 1x1 MMA tap)."""
        name = f"identity{c}.w"
        if name not in self.w.t:
            self.w.t[name] = torch.eye(c, dtype=self.adt, device=self.device).unsqueeze(0).contiguous()
        return name

    # ------------------------------------------------------------------ non-finite guard
    def raise_if_nonfinite(self, sync: bool = False) -> None:
        """Raise if a kernel of this device met inf / NaN GroupNorm statistics or model outputs (fp16 activations
        beyond 65504 turn into that one layer later). Without ``sync`` it reports what already-completed work found
        (free: a host read); with ``sync`` the current stream is synchronised first, so the answer covers every call
        made so far."""
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        rc = self.lib.wfk_nonfinite_status(self.dev_index, 1)
        if rc != 0:
            _cabi.check(rc, "non-finite guard")

    # ------------------------------------------------------------------ public
    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        """x [N, in_ch, H, W] fp32 cuda -> moments [N, 2*lc, H/8, W/8] fp32."""
        prog = self._program("enc", tuple(x.shape))
        return prog.run(x)

    def decode(self, z: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """z [N, lc, h, w] fp32 cuda -> [N, out_ch, 8h, 8w] fp32 (written into ``out`` when given: a contiguous
        fp32 CUDA tensor of that shape, e.g. a slice of the caller's [B*T, ...] result)."""
        prog = self._program("dec", tuple(z.shape))
        return prog.run(z, out=out)

    MAX_PLANS = 4   # programs own their activation pools (GBs at 384 x 384): keep the most recent shapes only

    def _program(self, kind: str, shape: Tuple[int, ...]) -> "_Program":
        key = (kind, shape)
        prog = self._plans.pop(key, None)
        if prog is None:
            while len(self._plans) >= self.MAX_PLANS:
                self._plans.pop(next(iter(self._plans)))   # least recently used (dicts keep insertion order)
            prog = _Program(self, kind, shape)
        self._plans[key] = prog
        return prog


class _Program:
    """A static list of (C function, args) bound to preallocated buffers for one input shape."""

    def __init__(self, eng: AKLEngine, kind: str, shape: Tuple[int, ...]):
        self.eng = eng
        self.lib = eng.lib
        self.dev = eng.device
        self.pool = _Pool(self.dev, eng.adt)
        self.bf = 1 if eng.bf16 else 0
        self.ops: List[Tuple[Callable, tuple, str]] = []
        self.plans: List[int] = []
        self.keep: List = []
        self.stats_bufs: List[torch.Tensor] = []
        self.n = int(shape[0])
        n_stats = 80
        self.stats_arena = torch.zeros(n_stats, self.n, eng.groups, 2, dtype=torch.float64, device=self.dev)
        self._stats_used = 0
        self.input = torch.empty(shape, dtype=torch.float32, device=self.dev)
        if kind == "enc":
            self.output = self._build_encoder(shape)
        else:
            self.output = self._build_decoder(shape)

    def __del__(self):
        try:
            for p in self.plans:
                self.lib.wfk_conv_plan_destroy(p)
        except Exception:
            pass

    # ------------------------------------------------------------------ running
    def run(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Single-stream use: the bound buffers are shared by every call of this program."""
        if x.dtype != torch.float32 or not x.is_cuda:
            raise TypeError("expected a float32 CUDA tensor")
        if out is not None and (out.shape != self.output.shape or out.dtype != torch.float32 or not out.is_contiguous()
                                or out.device != self.output.device):
            raise ValueError(f"out must be a contiguous float32 tensor of shape {tuple(self.output.shape)} on {self.dev}")
        self.eng.raise_if_nonfinite()      # what earlier, already completed calls found (no sync)
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        self.input.copy_(x)
        self.stats_arena.zero_()
        timer = TIMER
        for fn, args, what, flops, executed in self.ops:
            if timer is not None and (flops or timer.all_ops):
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                _cabi.check(fn(*args, stream), what)
                ev1.record()
                timer.record(what, flops, ev0, ev1, executed)
            else:
                _cabi.check(fn(*args, stream), what)
        if out is not None:
            out.copy_(self.output)
            return out
        return self.output.clone()

    # ------------------------------------------------------------------ helpers
    def _new_stats(self) -> torch.Tensor:
        s = self.stats_arena[self._stats_used]
        self._stats_used += 1
        if self._stats_used > self.stats_arena.shape[0]:
            raise RuntimeError("stats arena exhausted")
        return s

    def _add(self, fn, args, what, flops=0.0, executed=None):
        self.ops.append((fn, tuple(args), what, float(flops), float(flops if executed is None else executed)))

    def _conv_plan(self, desc: ConvDesc, what: str, nominal_flops: float, executed_flops: Optional[float] = None):
        """``executed_flops``: MMA work really issued when it differs from the nominal 2*M*N*K of the reference layer
        (defaults to: pixels x n_total x 64 x sum of the taps' K blocks of one phase, times the phases)."""
        h = C.c_void_p()
        _cabi.check(self.lib.wfk_conv_plan_create(C.byref(desc), C.byref(h)), f"conv_plan_create[{what}]")
        self.plans.append(h)
        if executed_flops is None:
            kb = sum(desc.taps[i].kblocks for i in range(desc.taps_per_phase))
            executed_flops = 2.0 * desc.n_frames * desc.tile_h * desc.tile_w * desc.n_total * 64 * kb * desc.num_phases
        self._add(self.lib.wfk_conv_plan_run, (h,), what, nominal_flops, executed_flops)

    @staticmethod
    def _view_nhwc(desc_view, t: torch.Tensor, n, h, w, c, pitch_c=None):
        pitch = c if pitch_c is None else pitch_c
        desc_view.ptr = t.data_ptr()
        dims = (c, w, 1, h, n)
        strides = (2, pitch * 2, w * pitch * 2, w * pitch * 2, h * w * pitch * 2)
        for i in range(5):
            desc_view.dim[i] = dims[i]
            desc_view.stride[i] = strides[i]

    @staticmethod
    def _view_w(desc_view, t: torch.Tensor, k, nrows, slabs, pitch_k=None, slab_stride=None):
        pitch = k if pitch_k is None else pitch_k
        desc_view.ptr = t.data_ptr()
        dims = (k, nrows, slabs)
        strides = (2, pitch * 2, (nrows * pitch * 2) if slab_stride is None else slab_stride)
        for i in range(3):
            desc_view.dim[i] = dims[i]
            desc_view.stride[i] = strides[i]

    def _epilogue(self, d: ConvDesc, bias, residual, out_h, out_f, stats, out_rows, out_cols, ldc, sy=1, sx=1):
        d.bias = _ptr(bias)
        d.residual = _ptr(residual)
        d.out_h = _ptr(out_h)
        d.out_f = _ptr(out_f)
        d.stats = _ptr(stats)
        d.out_rows, d.out_cols, d.out_sy, d.out_sx, d.ldc = out_rows, out_cols, sy, sx, ldc
        d.cpg = (d.n_total // self.eng.groups) if stats is not None else 0
        d.operand_bf16 = self.bf

    # ------------------------------------------------------------------ layer builders
    def gn_table(self, x: _Act, pname: str, what="gn_table") -> torch.Tensor:
        """(scale, shift) per (frame, channel) of GroupNorm(x): the apply + SiLU is fused into the consuming
        3x3 convolution's operand staging instead of a separate read+write pass over the tensor."""
        n, h, w, c = x.shape
        tab = self.pool.get((n, c, 2), torch.float32)
        t = self.eng.w.t
        self._add(self.lib.wfk_gn_table,
                  (x.stats.data_ptr(), t[pname + ".weight"].data_ptr(), t[pname + ".bias"].data_ptr(), n, h * w, c,
                   self.eng.groups, GN_EPS, tab.data_ptr()), what)
        return tab

    def _gn_inline(self, d: ConvDesc, x: _Act, pname: str):
        """GroupNorm(pname) + SiLU of the conv's input computed from x's raw statistics inside the kernel."""
        t = self.eng.w.t
        d.gn_stats = x.stats.data_ptr()
        d.gn_gamma = t[pname + ".weight"].data_ptr()
        d.gn_beta = t[pname + ".bias"].data_ptr()
        d.gn_eps = GN_EPS
        d.gn_groups = self.eng.groups

    def conv3x3(self, x: _Act, wname: str, bias: torch.Tensor, cout: int, residual: Optional[torch.Tensor] = None,
                shortcut: Optional[Tuple[torch.Tensor, str]] = None, want_stats=True, what="conv3x3",
                gn_tab: Optional[torch.Tensor] = None, gn_from: Optional[str] = None,
                count_shortcut_flops: bool = True) -> _Act:
        """3x3 stride-1 pad-1 conv (+bias, +residual | fused 1x1 shortcut) -> new activation. With ``gn_tab`` the
        input is the RAW tensor and GroupNorm+SiLU is applied while it is staged in shared memory."""
        n, h, w, cin = x.shape
        wt = self.eng.w.t[wname]
        out = self.pool.get((n, h, w, cout))
        stats = self._new_stats() if want_stats else None
        d = ConvDesc()
        self._view_nhwc(d.a[0], x.t, n, h, w, cin)
        self._view_w(d.b[0], wt, cin, cout, 9)
        d.n_frames, d.tile_h, d.tile_w, d.n_total = n, h, w, cout
        d.num_phases, d.taps_per_phase = 1, 9
        k = 0
        for r in range(3):
            for s in range(3):
                d.taps[k] = Tap(s - 1, r - 1, 0, 0, 0, r * 3 + s, cin // 64, 0)
                k += 1
        if shortcut is not None:
            xs, sname = shortcut
            cs = xs.shape[3]
            self._view_nhwc(d.a[1], xs, n, h, w, cs)
            self._view_w(d.b[1], self.eng.w.t[sname], cs, cout, 1)
            d.taps[9] = Tap(0, 0, 0, 1, 0, 0, cs // 64, 0)
            d.taps_per_phase = 10
        d.a_frame_mul, d.b_frame_mul = 1, 0
        self._epilogue(d, bias, residual, out, None, stats, h, w, cout)
        d.gn_table = _ptr(gn_tab)
        if gn_from is not None:
            self._gn_inline(d, x, gn_from)
        # nominal work of the REFERENCE layer: an identity "shortcut" that carries the residual add is not counted
        k_total = 9 * cin + (shortcut[0].shape[3] if (shortcut is not None and count_shortcut_flops) else 0)
        self._conv_plan(d, what, 2.0 * n * h * w * cout * k_total)
        return _Act(out, stats)

    def downsample(self, x: _Act, wname: str, bias: torch.Tensor, what="downsample") -> _Act:
        """F.pad(0,1,0,1) + conv3x3 stride 2 pad 0 (resnet.py:183-188) through a parity view of the input:
        dims (2C, W/2, 2, H/2, N) so that input pixel (2y+r, 2x+s) is (c + (s%2)C, x + s//2, r%2, y + r//2)."""
        n, h, w, c = x.shape
        if h % 2 or w % 2:
            raise ValueError("downsample needs even H and W")
        wt = self.eng.w.t[wname]
        oh, ow = h // 2, w // 2
        out = self.pool.get((n, oh, ow, c))
        stats = self._new_stats()
        d = ConvDesc()
        v = d.a[0]
        v.ptr = x.t.data_ptr()
        dims = (2 * c, ow, 2, oh, n)
        strides = (2, 2 * c * 2, w * c * 2, 2 * w * c * 2, h * w * c * 2)
        for i in range(5):
            v.dim[i] = dims[i]
            v.stride[i] = strides[i]
        self._view_w(d.b[0], wt, c, c, 9)
        d.n_frames, d.tile_h, d.tile_w, d.n_total = n, oh, ow, c
        d.num_phases, d.taps_per_phase = 1, 9
        k = 0
        for r in range(3):
            for s in range(3):
                d.taps[k] = Tap(s // 2, r // 2, r % 2, 0, (s % 2) * c, r * 3 + s, c // 64, 0)
                k += 1
        d.a_frame_mul, d.b_frame_mul = 1, 0
        self._epilogue(d, bias, None, out, None, stats, oh, ow, c)
        self._conv_plan(d, what, 2.0 * n * oh * ow * c * 9 * c)
        return _Act(out, stats)

    def upsample(self, x: _Act, wname: str, bias: torch.Tensor, what="upsample") -> _Act:
        """F.interpolate(x2, nearest) + conv3x3 (resnet.py:128,137-139) as four 2x2 sub-pixel convolutions
        on the low-res input (2.25x fewer MACs, no upsampled tensor materialised)."""
        n, h, w, c = x.shape
        wt = self.eng.w.t[wname]  # [16][c][c]
        out = self.pool.get((n, 2 * h, 2 * w, c))
        stats = self._new_stats()
        d = ConvDesc()
        self._view_nhwc(d.a[0], x.t, n, h, w, c)
        self._view_w(d.b[0], wt, c, c, 16)
        d.n_frames, d.tile_h, d.tile_w, d.n_total = n, h, w, c
        d.num_phases, d.taps_per_phase = 4, 4
        for a in (0, 1):
            for b in (0, 1):
                ph = a * 2 + b
                for i in (0, 1):
                    for j in (0, 1):
                        d.taps[ph * 4 + i * 2 + j] = Tap(_PHASE_OFFS[b][j], _PHASE_OFFS[a][i], 0, 0, 0,
                                                         ph * 4 + i * 2 + j, c // 64, 0)
        d.a_frame_mul, d.b_frame_mul = 1, 0
        self._epilogue(d, bias, None, out, None, stats, 2 * h, 2 * w, c, sy=2, sx=2)
        # nominal work of the reference layer: 3x3 conv on the upsampled (2h x 2w) tensor
        self._conv_plan(d, what, 2.0 * n * (2 * h) * (2 * w) * c * 9 * c)
        return _Act(out, stats)

    def gn(self, x: _Act, pname: str, silu=True, what="groupnorm") -> torch.Tensor:
        n, h, w, c = x.shape
        out = self.pool.get((n, h, w, c))
        t = self.eng.w.t
        self._add(self.lib.wfk_groupnorm_apply,
                  (x.t.data_ptr(), x.stats.data_ptr(), t[pname + ".weight"].data_ptr(), t[pname + ".bias"].data_ptr(),
                   n, h * w, c, self.eng.groups, GN_EPS, 1 if silu else 0, out.data_ptr(), self.bf), what)
        return out

    def resnet(self, x: _Act, p: str) -> _Act:
        """ResnetBlock2D.forward (resnet.py:454-495)."""
        t = self.eng.w.t
        cin = x.shape[3]
        cout = t[p + ".conv1.bias"].numel()
        if self.eng.fuse_gn and x.shape[1] >= 2:
            # GroupNorm apply + SiLU fused into the convs' halo staging: no normalised copies in HBM
            if self.eng.gn_inline:
                h1 = self.conv3x3(x, p + ".conv1.w", t[p + ".conv1.bias"], cout, what=p + ".gn1+conv1", gn_from=p + ".norm1")
                if (p + ".conv_shortcut.w") in t:
                    out = self.conv3x3(h1, p + ".conv2.w", t[p + ".conv2.bias_sc"], cout, shortcut=(x.t, p + ".conv_shortcut.w"),
                                       what=p + ".gn2+conv2+shortcut", gn_from=p + ".norm2")
                else:
                    out = self.conv3x3(h1, p + ".conv2.w", t[p + ".conv2.bias"], cout, residual=x.t,
                                       what=p + ".gn2+conv2+res", gn_from=p + ".norm2")
                self.pool.put(h1.t)
                self.pool.put(x.t)
                return out
            t1 = self.gn_table(x, p + ".norm1", what=p + ".norm1(table)")
            h1 = self.conv3x3(x, p + ".conv1.w", t[p + ".conv1.bias"], cout, what=p + ".gn1+conv1", gn_tab=t1)
            t2 = self.gn_table(h1, p + ".norm2", what=p + ".norm2(table)")
            if (p + ".conv_shortcut.w") in t:
                out = self.conv3x3(h1, p + ".conv2.w", t[p + ".conv2.bias_sc"], cout,
                                   shortcut=(x.t, p + ".conv_shortcut.w"), what=p + ".gn2+conv2+shortcut", gn_tab=t2)
            elif self.eng.res_as_mma and cout == 128 and cin == cout:
                # Cout = 128 layers are epilogue-bound with a residual (two chunks of loads, converts and adds per
                # accumulator chunk): feed the skip connection through the tensor cores instead, as a fused 1x1
                # "shortcut" with identity weights (exact: 1.0 * x accumulates in fp32), +11 % MMA work
                out = self.conv3x3(h1, p + ".conv2.w", t[p + ".conv2.bias"], cout,
                                   shortcut=(x.t, self.eng.identity_weight(cout)), what=p + ".gn2+conv2+res(mma)", gn_tab=t2,
                                   count_shortcut_flops=False)
            else:
                out = self.conv3x3(h1, p + ".conv2.w", t[p + ".conv2.bias"], cout, residual=x.t,
                                   what=p + ".gn2+conv2+res", gn_tab=t2)
            self.pool.put(t1)
            self.pool.put(t2)
            self.pool.put(h1.t)
            self.pool.put(x.t)
            return out
        a1 = self.gn(x, p + ".norm1", what=p + ".norm1")
        h1 = self.conv3x3(_Act(a1, None), p + ".conv1.w", t[p + ".conv1.bias"], cout, what=p + ".conv1")
        self.pool.put(a1)
        a2 = self.gn(h1, p + ".norm2", what=p + ".norm2")
        self.pool.put(h1.t)
        if (p + ".conv_shortcut.w") in t:
            out = self.conv3x3(_Act(a2, None), p + ".conv2.w", t[p + ".conv2.bias_sc"], cout,
                               shortcut=(x.t, p + ".conv_shortcut.w"), what=p + ".conv2+shortcut")
        else:
            out = self.conv3x3(_Act(a2, None), p + ".conv2.w", t[p + ".conv2.bias"], cout, residual=x.t,
                               what=p + ".conv2+res")
        self.pool.put(a2)
        self.pool.put(x.t)
        return out

    def attention(self, x: _Act, p: str) -> _Act:
        """AttentionBlock.forward, one head (attention.py:136-189)."""
        t = self.eng.w.t
        n, h, w, c = x.shape
        T = h * w
        if T % 8:
            raise ValueError("attention needs h*w to be a multiple of 8")
        a = self.gn(x, p + ".group_norm", silu=False, what=p + ".group_norm")
        # q | k projection: [n, T, 2c]
        qk = self.pool.get((n, T, 2 * c))
        d = ConvDesc()
        self._view_nhwc(d.a[0], a, n, 1, T, c)
        self._view_w(d.b[0], t[p + ".qk.w"], c, 2 * c, 1)
        d.n_frames, d.tile_h, d.tile_w, d.n_total = n, 1, T, 2 * c
        d.num_phases, d.taps_per_phase = 1, 1
        d.taps[0] = Tap(0, 0, 0, 0, 0, 0, c // 64, 0)
        d.a_frame_mul, d.b_frame_mul = 1, 0
        self._epilogue(d, t[p + ".qk.bias"], None, qk, None, None, 1, T, 2 * c)
        self._conv_plan(d, p + ".qk", 2.0 * n * T * (2 * c) * c)
        # V^T = Wv . X^T : [n, c, T]   (value bias is added after P.V: softmax rows sum to 1)
        vt = self.pool.get((n, c, T))
        d = ConvDesc()
        self._view_nhwc(d.a[0], t[p + ".value.w"], 1, 1, c, c)
        self._view_w(d.b[0], a, c, T, n)
        d.n_frames, d.tile_h, d.tile_w, d.n_total = n, 1, c, T
        d.num_phases, d.taps_per_phase = 1, 1
        d.taps[0] = Tap(0, 0, 0, 0, 0, 0, c // 64, 0)
        d.a_frame_mul, d.b_frame_mul = 0, 1
        self._epilogue(d, None, None, vt, None, None, 1, c, T)
        self._conv_plan(d, p + ".value^T", 2.0 * n * T * c * c)
        self.pool.put(a)
        # scores = Q K^T (fp32) -> row softmax -> O = P V (+ value bias): three launches, 63 MB of score / probability
        # traffic per frame at 2304 tokens.
        # Opt-in alternative (WFK_ATTN_FUSED=1, fp16 operands) that never materialises the fp32 scores: pass 1 computes
        # Q K^T and keeps only the row maxima, pass 2 computes it AGAIN and writes exp(s - max) as 16-bit values plus
        # fp32 row sums, pass 3 is P V with the 1 / sum scaling and the value bias in its epilogue (WFK_ACT_ROW_* in
        # include/wfk_b200.h). Measured on one B200, same box, bench.py (frames/s, alternating): fused 324.1 / 323.4,
        # default 324.6 / 324.0. Per 37 frames: default 322 + 276 + 172 us (scores, softmax, P V; since then 270 + 179 +
        # 172 us); fused 249 + 563 + 172 us -- a K = 512 tile gives the tensor pipe only ~2700 cycles of work, while 128 x 256 exponentials cost
        # the SM's 16 MUFU lanes 2048 cycles before any packing or storing, so the exp pass is epilogue-bound (and a
        # single-kernel flash attention has the same exp-per-MMA ratio plus O = 128 x 512 fp32 filling all of TMEM).
        # Both chains can run in groups of `ag` frames (WFK_ATTN_GROUP); measured: all 37 frames at once 314.7
        # frames/s, groups of 4 / 3 / 2 / 1 frames 312.8 / 311.0 / 309.5 / 303.2 -- the extra launches and partial
        # last waves cost more than the L2 residency saves, so the default is one group.
        o = self.pool.get((n, T, c))
        ag = max(1, min(n, self.eng.attn_group if self.eng.attn_group > 0 else n))
        fused = not self.bf and self.eng.attn_fused
        probs = self.pool.get((ag, T, T))
        esz = 2
        if fused:
            slots = 2 * ((T + 255) // 256) if (T % 256 == 0 or T > 256) else 2 * ((T + 127) // 128)   # 2 per N tile
            rmax = self.pool.get((ag, T, slots), torch.float32)
            rsum = self.pool.get((ag, T, slots), torch.float32)
        else:
            scores = self.pool.get((ag, T, T), torch.float32)

        def qk_desc(g0, ng):
            d = ConvDesc()
            self._view_nhwc(d.a[0], qk[g0:], ng, 1, T, c, pitch_c=2 * c)
            kview = qk[g0:].reshape(-1)[c:]
            self._view_w(d.b[0], kview, c, T, ng, pitch_k=2 * c, slab_stride=T * 2 * c * esz)
            d.n_frames, d.tile_h, d.tile_w, d.n_total = ng, 1, T, T
            d.num_phases, d.taps_per_phase = 1, 1
            d.taps[0] = Tap(0, 0, 0, 0, 0, 0, c // 64, 0)
            d.a_frame_mul, d.b_frame_mul = 1, 1
            return d

        for g0 in range(0, n, ag):
            ng = min(ag, n - g0)
            fl = 2.0 * ng * T * T * c
            if fused:
                d = qk_desc(g0, ng)
                self._epilogue(d, None, None, None, None, None, 1, T, T)
                d.act, d.row_out, d.row_ld = _cabi.ACT_ROW_MAX, rmax.data_ptr(), slots
                self._conv_plan(d, p + ".scores(max)", 0.0, fl)
                d = qk_desc(g0, ng)
                self._epilogue(d, None, None, probs, None, None, 1, T, T)
                d.act, d.row_in, d.row_out, d.row_ld = _cabi.ACT_ROW_EXP, rmax.data_ptr(), rsum.data_ptr(), slots
                d.row_scale = math.log2(math.e) / math.sqrt(c)
                self._conv_plan(d, p + ".scores(exp)", fl, fl)
            else:
                d = qk_desc(g0, ng)
                self._epilogue(d, None, None, None, scores, None, 1, T, T)
                self._conv_plan(d, p + ".scores", fl)
                self._add(self.lib.wfk_softmax_rows,
                          (scores.data_ptr(), ng * T, T, 1.0 / math.sqrt(c), probs.data_ptr(), self.bf), p + ".softmax")
            d = ConvDesc()
            self._view_nhwc(d.a[0], probs, ng, 1, T, T)
            self._view_w(d.b[0], vt[g0:], T, c, ng)
            d.n_frames, d.tile_h, d.tile_w, d.n_total = ng, 1, T, c
            d.num_phases, d.taps_per_phase = 1, 1
            d.taps[0] = Tap(0, 0, 0, 0, 0, 0, (T + 63) // 64, 0)
            d.a_frame_mul, d.b_frame_mul = 1, 1
            if fused:
                self._epilogue(d, None, None, o[g0:], None, None, 1, T, c)
                d.act, d.row_in, d.row_ld = _cabi.ACT_ROW_NORM, rsum.data_ptr(), slots
                d.shift2 = t[p + ".value.bias"].data_ptr()
            else:
                self._epilogue(d, t[p + ".value.bias"], None, o[g0:], None, None, 1, T, c)
            self._conv_plan(d, p + ".pv", fl)
        self.pool.put(qk)
        if fused:
            self.pool.put(rmax)
            self.pool.put(rsum)
        else:
            self.pool.put(scores)
        self.pool.put(probs)
        self.pool.put(vt)
        # proj + residual
        out = self.pool.get((n, h, w, c))
        stats = self._new_stats()
        d = ConvDesc()
        self._view_nhwc(d.a[0], o, n, 1, T, c)
        self._view_w(d.b[0], t[p + ".proj_attn.w3"], c, c, 1)
        d.n_frames, d.tile_h, d.tile_w, d.n_total = n, 1, T, c
        d.num_phases, d.taps_per_phase = 1, 1
        d.taps[0] = Tap(0, 0, 0, 0, 0, 0, c // 64, 0)
        d.a_frame_mul, d.b_frame_mul = 1, 0
        self._epilogue(d, t[p + ".proj_attn.bias"], x.t, out, None, stats, 1, T, c)
        self._conv_plan(d, p + ".proj+res", 2.0 * n * T * c * c)
        self.pool.put(o)
        self.pool.put(x.t)
        return _Act(out, stats)

    def mid(self, x: _Act, p: str) -> _Act:
        x = self.resnet(x, p + ".resnets.0")
        x = self.attention(x, p + ".attentions.0")
        return self.resnet(x, p + ".resnets.1")

    # ------------------------------------------------------------------ networks
    def _build_encoder(self, shape) -> torch.Tensor:
        eng, t = self.eng, self.eng.w.t
        n, cin, H, W = shape
        nb = len(eng.boc)
        if cin != eng.in_ch or H % (1 << (nb - 1)) or W % (1 << (nb - 1)):
            raise ValueError(f"encode: bad input shape {shape}")
        c0 = eng.boc[0]
        s0 = self.pool.get((n, H, W, c0))
        st = self._new_stats()
        if eng.stem_tc and "encoder.conv_in.w_tc" in t and c0 % 128 == 0 and (c0 // eng.groups) in (4, 8, 16):
            self._add(self.lib.wfk_conv3x3_stem_tc,
                      (self.input.data_ptr(), n, cin, H, W, 0, t["encoder.conv_in.w_tc"].data_ptr(),
                       t["encoder.conv_in.bias"].data_ptr(), c0, s0.data_ptr(), st.data_ptr(), c0 // eng.groups, self.bf),
                      "encoder.conv_in")
        else:
            self._add(self.lib.wfk_conv3x3_small_cin,
                      (self.input.data_ptr(), n, cin, H, W, None, None, t["encoder.conv_in.w_direct"].data_ptr(),
                       t["encoder.conv_in.bias"].data_ptr(), c0, s0.data_ptr(), st.data_ptr(), c0 // eng.groups),
                      "encoder.conv_in")
        x = _Act(s0, st)
        for i in range(nb):
            for j in range(eng.lpb):
                x = self.resnet(x, f"encoder.down_blocks.{i}.resnets.{j}")
            if i != nb - 1:
                p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                y = self.downsample(x, p + ".w", t[p + ".bias"], what=p)
                self.pool.put(x.t)
                x = y
        x = self.mid(x, "encoder.mid_block")
        _, h, w, c = x.shape
        out = torch.empty((n, 2 * eng.lc, h, w), dtype=torch.float32, device=self.dev)
        if eng.fuse_gn and h >= 2 and "encoder.conv_out.w_folded" in t and (2 * eng.lc) % 8 == 0:
            # conv_norm_out + SiLU fused into the halo staging, conv_out (+ folded quant_conv) on the tensor cores
            # (N padded to 8: a sliver of a 128-wide tile, still 5x faster than the CUDA-core gather), fp32 NHWC
            # result transposed to the NCHW moments tensor
            tab = None if eng.gn_inline else self.gn_table(x, "encoder.conv_norm_out", what="encoder.conv_norm_out(table)")
            npad = t["encoder.conv_out.bias_folded"].numel()
            mom = self.pool.get((n, h, w, npad), torch.float32)
            d = ConvDesc()
            self._view_nhwc(d.a[0], x.t, n, h, w, c)
            self._view_w(d.b[0], t["encoder.conv_out.w_folded"], c, npad, 9)
            d.n_frames, d.tile_h, d.tile_w, d.n_total = n, h, w, npad
            d.num_phases, d.taps_per_phase = 1, 9
            for r in range(3):
                for s_ in range(3):
                    d.taps[r * 3 + s_] = Tap(s_ - 1, r - 1, 0, 0, 0, r * 3 + s_, c // 64, 0)
            d.a_frame_mul, d.b_frame_mul = 1, 0
            self._epilogue(d, t["encoder.conv_out.bias_folded"], None, None, mom, None, h, w, npad)
            d.gn_table = _ptr(tab)
            if eng.gn_inline:
                self._gn_inline(d, x, "encoder.conv_norm_out")
            self._conv_plan(d, "encoder.gn+conv_out+quant_conv", 2.0 * n * h * w * (2 * eng.lc) * 9 * c)
            self._add(self.lib.wfk_nhwc_to_nchw_f32, (mom.data_ptr(), n, h * w, npad, out.data_ptr()), "moments->NCHW")
            self.keep.append((x.t, mom, tab))
            return out
        a = self.gn(x, "encoder.conv_norm_out", what="encoder.conv_norm_out")
        self.pool.put(x.t)
        self._add(self.lib.wfk_conv3x3_small_cout,
                  (a.data_ptr(), n, h, w, c, t["encoder.conv_out.w_direct"].data_ptr(),
                   t["encoder.conv_out.bias"].data_ptr(), 2 * eng.lc, t["quant_conv.w"].data_ptr(),
                   t["quant_conv.bias"].data_ptr(), out.data_ptr()), "encoder.conv_out+quant_conv")
        return out

    def _build_decoder(self, shape) -> torch.Tensor:
        eng, t = self.eng, self.eng.w.t
        n, lc, h, w = shape
        nb = len(eng.boc)
        if lc != eng.lc:
            raise ValueError(f"decode: bad latent shape {shape}")
        rev = list(reversed(eng.boc))
        c0 = rev[0]
        s0 = self.pool.get((n, h, w, c0))
        st = self._new_stats()
        if eng.stem_tc and "decoder.conv_in.w_tc" in t and c0 % 128 == 0 and (c0 // eng.groups) in (4, 8, 16):
            self._add(self.lib.wfk_conv3x3_stem_tc,
                      (self.input.data_ptr(), n, lc, h, w, 1, t["decoder.conv_in.w_tc"].data_ptr(),
                       t["decoder.conv_in.bias"].data_ptr(), c0, s0.data_ptr(), st.data_ptr(), c0 // eng.groups, self.bf),
                      "post_quant_conv+decoder.conv_in")
        else:
            self._add(self.lib.wfk_conv3x3_small_cin,
                      (self.input.data_ptr(), n, lc, h, w, t["post_quant_conv.w"].data_ptr(),
                       t["post_quant_conv.bias"].data_ptr(), t["decoder.conv_in.w_direct"].data_ptr(),
                       t["decoder.conv_in.bias"].data_ptr(), c0, s0.data_ptr(), st.data_ptr(), c0 // eng.groups),
                      "post_quant_conv+decoder.conv_in")
        x = _Act(s0, st)
        x = self.mid(x, "decoder.mid_block")
        for i in range(nb):
            for j in range(eng.lpb + 1):
                x = self.resnet(x, f"decoder.up_blocks.{i}.resnets.{j}")
            if i != nb - 1:
                p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                y = self.upsample(x, p + ".w_phase", t[p + ".bias"], what=p)
                self.pool.put(x.t)
                x = y
        _, H, W, c = x.shape
        out = torch.empty((n, eng.out_ch, H, W), dtype=torch.float32, device=self.dev)
        if eng.out_ch == 1:
            # fused GroupNorm + SiLU + conv_out: one read of the raw stream, no normalised copy
            self._add(self.lib.wfk_gn_silu_conv3x3_c1,
                      (x.t.data_ptr(), x.stats.data_ptr(), t["decoder.conv_norm_out.weight"].data_ptr(),
                       t["decoder.conv_norm_out.bias"].data_ptr(), n, H, W, c, eng.groups, GN_EPS,
                       t["decoder.conv_out.w_tap"].data_ptr(), float(t["decoder.conv_out.bias"][0].item()),
                       out.data_ptr(), self.bf), "decoder.conv_norm_out+silu+conv_out")
            self.keep.append(x.t)
            return out
        a = self.gn(x, "decoder.conv_norm_out", what="decoder.conv_norm_out")
        self.pool.put(x.t)
        self._add(self.lib.wfk_conv3x3_small_cout,
                  (a.data_ptr(), n, H, W, c, t["decoder.conv_out.w_direct"].data_ptr(),
                   t["decoder.conv_out.bias"].data_ptr(), eng.out_ch, None, None, out.data_ptr()),
                  "decoder.conv_out")
        return out

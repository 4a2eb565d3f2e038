"""ctypes binding of libwfk_b200.so (the C ABI in ``include/wfk_b200.h``).

This is the reference-side FFI stub INTEGRATION.md describes. There is deliberately no fallback:
if the shared library is missing, or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# WFK_LIB_PATH selects another build of the same ABI (same-box A/B runs of two kernel versions)
_LIB_PATH = Path(os.environ.get("WFK_LIB_PATH") or (Path(__file__).resolve().parent / "libwfk_b200.so"))

WFK_MAX_THRESHOLDS = 8
WFK_NUM_POOLS = 3
WFK_MAX_TAPS = 40
ABI_VERSION = 3          # WFK_ABI_VERSION of include/wfk_b200.h these bindings mirror


class MetricPartials(C.Structure):
    _fields_ = [
        ("counts", C.c_int64 * 4 * WFK_MAX_THRESHOLDS * WFK_NUM_POOLS),
        ("n_elems", C.c_int64 * WFK_NUM_POOLS),
        ("n_frames", C.c_int64),
        ("abs_sum", C.c_double * WFK_NUM_POOLS),
        ("sq_sum", C.c_double),
        ("ssim_sum", C.c_double),
        ("psnr_sum", C.c_double),
        ("reserved", C.c_double * 2),
    ]


class Tap(C.Structure):
    _fields_ = [("dx", C.c_int8), ("dy", C.c_int8), ("q", C.c_int8), ("src", C.c_int8),
                ("c_off", C.c_int16), ("b_slab", C.c_int16), ("kblocks", C.c_int16), ("reserved", C.c_int16)]


class View5(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dim", C.c_int64 * 5), ("stride", C.c_int64 * 5)]


class View3(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dim", C.c_int64 * 3), ("stride", C.c_int64 * 3)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("a", View5 * 2), ("b", View3 * 2),
        ("n_frames", C.c_int32), ("tile_h", C.c_int32), ("tile_w", C.c_int32), ("n_total", C.c_int32),
        ("num_phases", C.c_int32), ("taps_per_phase", C.c_int32),
        ("taps", Tap * WFK_MAX_TAPS),
        ("a_frame_mul", C.c_int32), ("b_frame_mul", C.c_int32),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("out_h", C.c_void_p), ("out_f", C.c_void_p),
        ("stats", C.c_void_p),
        ("out_rows", C.c_int32), ("out_cols", C.c_int32), ("out_sy", C.c_int32), ("out_sx", C.c_int32),
        ("ldc", C.c_int32), ("cpg", C.c_int32), ("operand_bf16", C.c_int32),
        ("gn_table", C.c_void_p),
        ("gn_stats", C.c_void_p), ("gn_gamma", C.c_void_p), ("gn_beta", C.c_void_p), ("gn_eps", C.c_float),
        ("gn_groups", C.c_int32),
        ("act", C.c_int32), ("act2", C.c_int32), ("act_slope", C.c_float),
        ("out2_h", C.c_void_p), ("scale2", C.c_void_p), ("shift2", C.c_void_p),
        ("row_in", C.c_void_p), ("row_out", C.c_void_p), ("row_ld", C.c_int32), ("row_scale", C.c_float),
    ]


ACT_NONE, ACT_LEAKY_RELU, ACT_GELU, ACT_SIGMOID, ACT_SILU, ACT_ROW_MAX, ACT_ROW_EXP, ACT_ROW_NORM = range(8)


_PROTOTYPES = {
    "wfk_strerror": (C.c_char_p, [C.c_int]),
    "wfk_last_error": (C.c_char_p, []),
    "wfk_abi_version": (C.c_int, []),
    "wfk_init": (C.c_int, [C.c_int]),
    "wfk_launch_count": (C.c_int64, []),
    "wfk_nonfinite_status": (C.c_int, [C.c_int, C.c_int]),
    "wfk_stage_vil_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "wfk_stage_vil_windows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_void_p, C.c_int, C.c_void_p]),
    "wfk_stage_vil_windows_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                           C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "wfk_convattn_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "wfk_render_panels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 9),
    "wfk_predict_linear": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wfk_dlinear": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                              C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p]),
    "wfk_convmodel_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "wfk_metrics_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "wfk_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int,
                              C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wfk_pooled_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                    C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wfk_crps_ensemble": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p]),
    "wfk_ensemble_mean": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "wfk_conv_plan_create": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(C.c_void_p)]),
    "wfk_conv_plan_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "wfk_conv_plan_destroy": (None, [C.c_void_p]),
    "wfk_groupnorm_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "wfk_gn_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                               C.c_void_p, C.c_void_p]),
    "wfk_conv3x3_small_cin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "wfk_conv3x3_stem_tc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "wfk_conv3x3_small_cout": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wfk_conv3x3_small_cout_act": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "wfk_conv4x4s2_c1in": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wfk_conv1x1_cout1": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_int,
                                    C.c_void_p, C.c_void_p]),
    "wfk_logit_sums": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "wfk_patchify16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wfk_unpatchify16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wfk_mha_small": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wfk_cross_encode_attn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p]),
    "wfk_layernorm_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "wfk_bcast_add_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wfk_gn_silu_conv3x3_c1": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "wfk_softmax_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "wfk_gaussian_posterior": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "wfk_nhwc_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wfk_f32_to_f16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None
_initialised_devices = set()   # devices registered with wfk_init (any number per process; the library is per-device)


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """dlopen the library and attach prototypes. Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python -m weatherforecastingtoolkit_b200.build` "
                "(there is no CPU / PyTorch fallback for this path)")
        lib = C.CDLL(os.fspath(_LIB_PATH))
        for name, (res, args) in _PROTOTYPES.items():  # debug probes (wfk_debug_*) are bound ad hoc by scripts/
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        got = lib.wfk_abi_version()
        if got != ABI_VERSION:   # a stale library would read the ctypes mirrors of the structs with the wrong layout
            raise RuntimeError(f"{_LIB_PATH} has ABI version {got}, the Python bindings expect {ABI_VERSION}: rebuild it "
                               "with `python -m weatherforecastingtoolkit_b200.build --force`")
        _lib = lib
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        lib = load()
        raise RuntimeError(f"libwfk_b200 {what}: {lib.wfk_strerror(status).decode()} ({status}): "
                           f"{lib.wfk_last_error().decode()}")


def init(device: int = 0) -> C.CDLL:
    """Load the library and register ``device`` (a B200) with it. Raises RuntimeError when it is not an sm_100 GPU.
    Any number of devices may be registered; every call then runs on the device of the stream it is given and
    leaves the caller's current device untouched."""
    lib = load()
    device = int(device)
    if device not in _initialised_devices:
        check(lib.wfk_init(device), "wfk_init")
        _initialised_devices.add(device)
    return lib


def launch_count() -> int:
    return int(load().wfk_launch_count())

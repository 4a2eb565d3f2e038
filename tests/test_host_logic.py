"""Host-side logic that needs no GPU: scores from partials, weight packing, synthetic data,
drop-in module surface."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import metric_case_inputs
from oracle import metrics_oracle as MO


def _partials_from_oracle(p, t):
    from weatherforecastingtoolkit_b200.metrics import MetricPartials
    pr = MO.partials(p, t)
    ints = np.zeros(100, dtype=np.int64)
    c = np.zeros((3, 8, 4), dtype=np.int64)
    c[:, :6] = pr["counts"]
    ints[:96] = c.reshape(-1)
    ints[96:99] = pr["n_elems"]
    ints[99] = pr["n_frames"]
    fl = np.zeros(8)
    fl[0:3] = pr["abs_sum"]
    fl[3], fl[4], fl[5] = pr["sq_sum"], pr["ssim_sum"], pr["psnr_sum"]
    return MetricPartials(ints, fl, 6)


@pytest.mark.parametrize("name", ["rand_2x10x64", "unclamped_1x3x50x70"])
def test_scores_from_partials_match_reference_dict(golden_metrics, name):
    """The host half of calc_metrics: exact integer counts + float64 sums -> the reference's 56 keys."""
    from weatherforecastingtoolkit_b200.metrics import scores_from_partials
    p, t = metric_case_inputs(name)
    got = scores_from_partials(_partials_from_oracle(p, t), extended=True)
    want = golden_metrics[name]["metrics"]
    assert list(want) == [k for k in got if k in want]
    for k, v in want.items():
        if k.startswith(("CSI", "HSS", "paper_CSI", "paper_HSS")):
            assert got[k] == v, k
        else:
            assert got[k] == pytest.approx(v, rel=1e-5, abs=1e-6), k
    assert got["MAE"] == got["CRPS"] and got["MSE"] > 0
    c = MO.integer_counts(p, t)[0, 0]
    assert got["POD_0"] == pytest.approx(c[0] / (c[0] + c[1]))
    assert got["FAR_0"] == pytest.approx(c[2] / (c[0] + c[2]))


def test_partials_are_additive_and_roundtrip_through_f64():
    from weatherforecastingtoolkit_b200.metrics import MetricPartials
    p, t = metric_case_inputs("rand_2x10x64")
    whole = _partials_from_oracle(p, t)
    a, b = _partials_from_oracle(p[:1], t[:1]), _partials_from_oracle(p[1:], t[1:])
    s = a + b
    assert np.array_equal(s.ints, whole.ints)
    assert np.allclose(s.floats, whole.floats, rtol=1e-12)
    big = MetricPartials(whole.ints * (2 ** 40 // 1000), whole.floats, 6)   # counts far above 2**24
    back = MetricPartials.from_f64_vector(big.as_f64_vector(), 6)
    assert np.array_equal(back.ints, big.ints)


def test_float32_threshold_rounding_H2():
    """pred >= python-float threshold compares in float32: the bit patterns the kernel gets."""
    want = [0x3D808081, 0x3E949495, 0x3F058586, 0x3F20A0A1, 0x3F35B5B6, 0x3F5BDBDC]
    got = [int(np.float32(th).view(np.uint32)) for th in MO.THRESHOLDS]
    assert got == want
    x = torch.tensor([np.float32(74 / 255)])
    assert bool((x >= 74 / 255).item())


def test_phase_weights_equal_nearest_upsample_conv():
    """The 4-phase 2x2 decomposition used for Upsample2D is the same linear map as
    F.interpolate(x2, nearest) + conv3x3 (resnet.py:128,137-139)."""
    from weatherforecastingtoolkit_b200.engine import PackedAKL, _PHASE_OFFS
    torch.manual_seed(0)
    c = 8
    w = torch.randn(c, c, 3, 3, dtype=torch.float64)
    x = torch.randn(1, c, 5, 6, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1)
    pw = PackedAKL._phase_weights(w.float()).double().reshape(2, 2, 2, 2, c, c)
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for a in (0, 1):
        for b in (0, 1):
            acc = torch.zeros(1, c, 5, 6, dtype=torch.float64)
            for i in (0, 1):
                for j in (0, 1):
                    dy, dx = _PHASE_OFFS[a][i], _PHASE_OFFS[b][j]
                    patch = xp[:, :, 1 + dy:1 + dy + 5, 1 + dx:1 + dx + 6]
                    acc += torch.einsum("oc,nchw->nohw", pw[a, b, i, j], patch)
            out[:, :, a::2, b::2] = acc
    assert torch.allclose(out, ref, atol=2e-2)   # weights were rounded to fp16 in the pack
    assert ((out - ref).norm() / ref.norm()).item() < 1e-3


def test_synthetic_generators_are_deterministic():
    from weatherforecastingtoolkit_b200.synthetic import (PATHB_AKL_CONFIG, akl_param_shapes, make_akl_state_dict,
                                                          make_predictor_params, make_vil_sequences)
    a = make_vil_sequences(2, 64, 64, 25, seed=3)
    b = make_vil_sequences(2, 64, 64, 25, seed=3)
    assert a.dtype == torch.uint8 and a.shape == (2, 64, 64, 25) and torch.equal(a, b)
    assert 0.2 < (a == 0).float().mean().item() < 0.8
    for th in (16, 74, 133, 160, 181, 219):
        frac = (make_vil_sequences(4, 128, 128, 5, seed=1) >= th).float().mean().item()
        assert 0.0 < frac < 1.0
    sd = make_akl_state_dict(PATHB_AKL_CONFIG, 0)
    assert list(sd) == list(akl_param_shapes(PATHB_AKL_CONFIG))
    assert sum(v.numel() for v in sd.values()) == 83_649_253      # 83.65 M params
    w, bias = make_predictor_params()
    assert w.shape == (48, 52) and bias.shape == (48,)
    assert torch.equal(make_akl_state_dict(PATHB_AKL_CONFIG, 0)["quant_conv.weight"], sd["quant_conv.weight"])


def test_drop_in_module_surface_cpu():
    """Constructor signature, state_dict names and error behaviour mirror the reference class."""
    from weatherforecastingtoolkit_b200.models.autoencoderkl import AutoencoderKL, DiagonalGaussianDistribution
    from weatherforecastingtoolkit_b200.synthetic import PATHB_AKL_CONFIG, make_akl_state_dict
    m = AutoencoderKL(**PATHB_AKL_CONFIG)
    sd = make_akl_state_dict(PATHB_AKL_CONFIG, 0)
    assert m.load_state_dict(sd, strict=True).missing_keys == []
    assert not any(p.requires_grad for p in m.parameters())
    with pytest.raises(RuntimeError):
        m.encode(torch.zeros(1, 1, 64, 64))          # CPU tensor: refuses, no fallback
    with pytest.raises(ValueError):
        AutoencoderKL(down_block_types=("Foo",), up_block_types=("UpDecoderBlock2D",))
    with pytest.raises(RuntimeError):                 # the posterior arithmetic is a CUDA kernel too: no CPU path
        DiagonalGaussianDistribution(torch.randn(2, 8, 4, 4))


def test_rollout_wrappers_refuse_cpu():
    from weatherforecastingtoolkit_b200.rollout import LatentLinearPredictor, stage_vil
    with pytest.raises(RuntimeError):
        stage_vil(torch.zeros(1, 8, 8, 25, dtype=torch.uint8))
    p = LatentLinearPredictor()
    assert p.weight.shape == (48, 52)
    assert p(torch.zeros(3, 52)).shape == (3, 48)       # nn.Linear forward kept for training code
    with pytest.raises(RuntimeError):
        p.rollout(torch.zeros(1, 25, 4, 8, 8))


def test_sequent_windows_matches_reference_sampler():
    """(event, first frame) order of SEVIRDataLoader._idx_sample (pipeline/datasets/sevir/sevir.py:864-877),
    restated: num_seq_per_event = 1 + (raw_seq_len - seq_len) // stride (:327-328)."""
    from weatherforecastingtoolkit_b200.rollout import sequent_windows
    raw, seq, stride, bs = 49, 25, 12, 4
    per_event = 1 + (raw - seq) // stride
    for index in range(3):
        event_idx, seq_idx = (index * bs) // per_event, (index * bs) % per_event
        want = []
        while len(want) < bs:
            want.append((event_idx, seq_idx * stride))
            seq_idx += 1
            if seq_idx >= per_event:
                event_idx, seq_idx = event_idx + 1, 0
        assert sequent_windows(10, raw, seq, stride, start=index * bs, count=bs) == want
    assert len(sequent_windows(2, 49, 25, 12)) == 6 and sequent_windows(1, 25, 25, 12) == [(0, 0)]


def test_metric_accumulator_refuses_cpu_and_empty():
    from weatherforecastingtoolkit_b200 import metrics as M
    acc = M.MetricAccumulator()
    with pytest.raises(RuntimeError):
        acc.compute()
    with pytest.raises(RuntimeError):
        acc.update(torch.rand(1, 2, 1, 32, 32), torch.rand(1, 2, 1, 32, 32))


def test_log_metrics_matches_reference_function(monkeypatch):
    """rollout.log_metrics against the UNMODIFIED ``pipeline.helpers.log_metrics`` (helpers.py:142-153; extracted
    with ast because helpers.py imports torch_lr_finder / lightning / matplotlib, none of which exist here): same
    detach, same tag prefixing, same log_dict keyword arguments. With ``process_group`` the counts were already summed
    over ranks, so Lightning must NOT average the scalars again: sync_dist=False."""
    import ast
    import os

    from conftest import REFERENCE, has_reference
    from weatherforecastingtoolkit_b200 import rollout

    seen = {}

    def fake_calc_metrics(pred, target, **kw):
        seen["args"] = (pred, target, kw)
        return {"CSI_0": 0.25, "SSIM": 0.5, "paper_CRPS": 0.125}

    class PL:
        def __init__(self):
            self.calls = []

        def log_dict(self, d, **kw):
            self.calls.append((dict(d), kw))

    monkeypatch.setattr(rollout.wf_metrics, "calc_metrics", fake_calc_metrics)
    p = torch.rand(1, 2, 1, 16, 16, requires_grad=True)
    t = torch.rand(1, 2, 1, 16, 16)
    mine = PL()
    rollout.log_metrics(p * 1.0, t, "val", mine)
    assert mine.calls == [({"val_CSI_0": 0.25, "val_SSIM": 0.5, "val_paper_CRPS": 0.125},
                           {"on_step": True, "on_epoch": True, "sync_dist": True})]
    assert not seen["args"][0].requires_grad and seen["args"][2] == {"process_group": None}
    grp = PL()
    rollout.log_metrics(p, t, "test", grp, process_group="fake-group")
    assert grp.calls[0][1]["sync_dist"] is False and list(grp.calls[0][0]) == ["test_CSI_0", "test_SSIM", "test_paper_CRPS"]
    assert seen["args"][2] == {"process_group": "fake-group"}
    # non-tensor inputs pass through untouched, as in the reference (isinstance checks)
    rollout.log_metrics([1.0], [2.0], "x", PL())
    assert seen["args"][0] == [1.0]
    if not has_reference():
        return
    src = open(os.path.join(REFERENCE, "pipeline", "helpers.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "log_metrics")
    ns = {"torch": torch, "calc_metrics": lambda a, b: fake_calc_metrics(a, b)}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "helpers.py", "exec"), ns)
    ref = PL()
    ns["log_metrics"](p * 1.0, t, "val", ref)
    assert ref.calls == mine.calls
    assert not seen["args"][0].requires_grad

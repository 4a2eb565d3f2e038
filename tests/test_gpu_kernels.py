"""-m gpu: every sm_100a kernel, called through the C ABI, against the CPU oracle / golden vectors."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import METRIC_CASES, metric_case_inputs

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def lib():
    from weatherforecastingtoolkit_b200 import _cabi
    return _cabi.init(0)


# ------------------------------------------------------------------ staging (bit-exact)
@pytest.mark.parametrize("shape", [(2, 384, 384, 25), (1, 50, 70, 7), (1, 16, 16, 1), (3, 33, 17, 25)])
def test_stage_vil_bitexact(lib, shape):
    from oracle import akl_oracle as O
    from weatherforecastingtoolkit_b200.rollout import stage_vil
    g = torch.Generator().manual_seed(sum(shape))
    u8 = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
    want = O.stage_vil(u8).permute(0, 3, 1, 2).unsqueeze(2).contiguous()
    got = stage_vil(u8.to(DEV))
    assert got.shape == want.shape
    assert torch.equal(got.cpu(), want)
    got16 = stage_vil(u8.to(DEV), dtype=torch.float16)
    assert torch.equal(got16.cpu(), want.to(torch.float16))


def test_stage_vil_rejects_bad_input(lib):
    from weatherforecastingtoolkit_b200.rollout import stage_vil
    with pytest.raises(TypeError):
        stage_vil(torch.zeros(1, 8, 8, 2, device=DEV))
    with pytest.raises(RuntimeError):
        stage_vil(torch.zeros(1, 8, 8, 2, dtype=torch.uint8))


# ------------------------------------------------------------------ predictor
@pytest.mark.parametrize("b,hw", [(1, 8), (3, 48)])
def test_predictor_vs_oracle(lib, b, hw):
    from oracle import akl_oracle as O
    from weatherforecastingtoolkit_b200.rollout import LatentLinearPredictor
    from weatherforecastingtoolkit_b200.synthetic import make_predictor_params
    w, bias = make_predictor_params(seed=3)
    torch.manual_seed(b)
    lat = torch.randn(b, 25, 4, hw, hw)
    pred_o, tgt_o, loss_o = O.predictor_rollout(lat, w, bias)
    p = LatentLinearPredictor()
    p.weight.data.copy_(w)
    p.bias.data.copy_(bias)
    pred, tgt, loss = p.rollout(lat.to(DEV))
    assert torch.allclose(pred.cpu(), pred_o, atol=2e-6, rtol=1e-5)
    assert torch.equal(tgt.cpu(), tgt_o)  # (x - last) + last, same fp32 ops
    assert loss.item() == pytest.approx(loss_o.item(), rel=1e-5)


def test_predictor_autoregressive_vs_oracle(lib):
    """north_star: the linear latent predictor stepped autoregressively (blocks of 12 frames from the last 13)."""
    from oracle import akl_oracle as O
    from weatherforecastingtoolkit_b200.rollout import LatentLinearPredictor
    from weatherforecastingtoolkit_b200.synthetic import make_predictor_params
    w, bias = make_predictor_params(seed=5)
    torch.manual_seed(11)
    inp = torch.randn(2, 13, 4, 12, 12)
    want = O.predictor_autoregressive(inp, w, bias, blocks=3)
    p = LatentLinearPredictor()
    p.weight.data.copy_(w)
    p.bias.data.copy_(bias)
    got = p.rollout_autoregressive(inp.to(DEV), blocks=3)
    assert got.shape == (2, 36, 4, 12, 12)
    assert torch.allclose(got.cpu(), want, atol=2e-5, rtol=1e-4)
    # the first block is the one-shot rollout of the reference step
    one, _, _ = p.rollout(torch.cat([inp, torch.zeros(2, 12, 4, 12, 12)], 1).to(DEV))
    assert torch.equal(one, got[:, :12])


# ------------------------------------------------------------------ metrics
@pytest.mark.parametrize("name", METRIC_CASES)
def test_metrics_vs_golden(lib, golden_metrics, name):
    from weatherforecastingtoolkit_b200 import metrics as M
    p, t = metric_case_inputs(name)
    mp = M.metric_partials(p.to(DEV), t.to(DEV))
    # bit-exact integer contingency counts (tp, fn, fp, tn) for 3 pools x 6 thresholds
    assert mp.counts.tolist() == golden_metrics[name]["counts"]
    got = M.scores_from_partials(mp)
    want = golden_metrics[name]["metrics"]
    assert list(got) == list(want)
    for k, v in want.items():
        if k.startswith(("CSI", "HSS", "paper_CSI", "paper_HSS")):
            assert got[k] == v, k                      # bit-exact: same counts, same float32 ratio ops
        elif "SSIM" in k:
            assert got[k] == pytest.approx(v, abs=1e-3), k   # north_star tolerance; observed ~1e-6
            assert got[k] == pytest.approx(v, abs=2e-5), k
        elif "PSNR" in k:
            assert got[k] == pytest.approx(v, rel=1e-5), k
        else:  # CRPS == MAE
            assert got[k] == pytest.approx(v, abs=1e-6), k


@pytest.mark.parametrize("name", METRIC_CASES)
def test_metrics_ssim_psnr_vs_independent(lib, name):
    """SSIM / PSNR of the fused kernel against the INDEPENDENT float64 implementation (oracle/ssim_independent.py:
    scipy separable filtering over the valid region, no code shared with the torchmetrics restatement)."""
    from oracle import ssim_independent as SI
    from weatherforecastingtoolkit_b200 import metrics as M
    p, t = metric_case_inputs(name)
    got = M.calc_metrics(p.to(DEV), t.to(DEV))
    pc, tc = p.clamp(0, 1).numpy(), t.clamp(0, 1).numpy()
    assert got["SSIM"] == pytest.approx(SI.ssim(pc, tc), abs=1e-4)
    assert got["PSNR"] == pytest.approx(SI.psnr_per_frame_mean(pc, tc), rel=1e-5)
    # unclamped stand-alone functions, incl. negative targets (PSNR range through zero)
    pd, td = (p - 0.3).to(DEV), (t - 0.3).to(DEV)
    assert M.ssim(pd, td) == pytest.approx(SI.ssim((p - 0.3).numpy(), (t - 0.3).numpy()), abs=1e-4)
    assert M.psnr(pd, td) == pytest.approx(SI.psnr_per_frame_mean((p - 0.3).numpy(), (t - 0.3).numpy()), rel=1e-5)


def test_metrics_ssim_analytic_anchors(lib):
    from weatherforecastingtoolkit_b200 import metrics as M
    a, b = 0.25, 0.75
    pa, tb = torch.full((1, 2, 1, 40, 56), a, device=DEV), torch.full((1, 2, 1, 40, 56), b, device=DEV)
    closed = (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)
    assert M.ssim(pa, tb) == pytest.approx(closed, abs=1e-3)     # float32 noise floor of a constant image: ~3e-4 of C2
    t = torch.rand(1, 1, 1, 32, 32)
    t[0, 0, 0, 0, 0], t[0, 0, 0, 0, 1] = 0.0, 1.0
    assert M.psnr((t + 0.1).to(DEV), t.to(DEV)) == pytest.approx(20.0, abs=1e-4)
    tp = t * 0.5 + 0.25     # strictly positive target: range = t.max() - 0
    assert M.psnr((tp + 0.1).to(DEV), tp.to(DEV)) == pytest.approx(10 * math.log10(0.75 ** 2 / 0.01), abs=1e-4)


@pytest.mark.parametrize("shape", [(1, 2, 1, 17, 23), (1, 1, 1, 33, 130), (2, 1, 1, 129, 67), (1, 1, 1, 70, 400), (1, 1, 1, 40, 800)])
def test_metrics_ragged_and_wide_shapes(lib, shape):
    """Odd widths (scalar-load path), widths beyond one 384-column strip (multi-strip path with 16-column overlaps),
    heights that are not multiples of the 8-row chunk / 32-row segment."""
    from oracle import metrics_oracle as MO
    from weatherforecastingtoolkit_b200 import metrics as M
    torch.manual_seed(sum(shape))
    p, t = torch.rand(*shape) * 1.2 - 0.1, torch.rand(*shape) * 1.2 - 0.1
    mp = M.metric_partials(p.to(DEV), t.to(DEV))
    assert mp.counts.tolist() == MO.integer_counts(p, t).tolist()
    want = MO.partials(p, t)
    assert mp.n_elems.tolist() == list(want["n_elems"])
    assert mp.ssim_sum == pytest.approx(want["ssim_sum"], abs=2e-5 * shape[0] * shape[1])
    assert mp.psnr_sum == pytest.approx(want["psnr_sum"], rel=1e-5)
    assert np.allclose(mp.abs_sum, want["abs_sum"], rtol=1e-5)
    assert mp.sq_sum == pytest.approx(want["sq_sum"], rel=1e-5)


def test_metrics_unsorted_thresholds(lib):
    """The kernel sorts thresholds internally (early exit of its compare loop); counts come back in the caller's order."""
    from oracle import metrics_oracle as MO
    from weatherforecastingtoolkit_b200 import metrics as M
    p, t = metric_case_inputs("vil_2x12x384")
    thr = [0.6, 16 / 255, 0.4, 219 / 255, 0.05]
    mp = M.metric_partials(p[:1, :4].to(DEV), t[:1, :4].to(DEV), thr)
    assert mp.counts.tolist() == MO.integer_counts(p[:1, :4], t[:1, :4], thr).tolist()


def test_log_metrics_gpu(lib):
    """rollout.log_metrics == pipeline.helpers.log_metrics (helpers.py:142-153): tag-prefixed calc_metrics dict handed
    to pl_module.log_dict(on_step=True, on_epoch=True, sync_dist=True)."""
    from weatherforecastingtoolkit_b200 import metrics as M
    from weatherforecastingtoolkit_b200.rollout import log_metrics

    class PL:
        def __init__(self):
            self.calls = []

        def log_dict(self, d, **kw):
            self.calls.append((d, kw))

    p, t = metric_case_inputs("rand_2x10x64")
    pd, td = p.to(DEV).requires_grad_(True), t.to(DEV)
    pl = PL()
    log_metrics(pd, td, "val", pl)
    (d, kw), = pl.calls
    want = M.calc_metrics(p.to(DEV), td)
    assert list(d) == [f"val_{k}" for k in want] and len(d) == 56
    assert all(d[f"val_{k}"] == v or (v != v and d[f"val_{k}"] != d[f"val_{k}"]) for k, v in want.items())
    assert all(isinstance(v, float) for v in d.values())
    assert kw == {"on_step": True, "on_epoch": True, "sync_dist": True}


def test_metrics_standalone_functions(lib):
    from oracle import metrics_oracle as MO
    from weatherforecastingtoolkit_b200 import metrics as M
    p, t = metric_case_inputs("unclamped_1x3x50x70")   # NOT clamped by the stand-alone functions
    pd, td = p.to(DEV), t.to(DEV)
    th = 74 / 255
    assert M.csi(pd, td, th) == MO.csi(p, t, th)
    assert M.csi(pd, td, th, "avg", 4) == MO.csi(p, t, th, "avg", 4)
    assert M.hss(pd, td, th, "avg", 16) == MO.hss(p, t, th, "avg", 16)
    assert M.crps(pd, td) == pytest.approx(MO.crps(p, t), abs=1e-6)
    assert M.crps(pd, td, "avg", 4) == pytest.approx(MO.crps(p, t, "avg", 4), abs=1e-6)
    assert M.ssim(pd, td) == pytest.approx(MO.ssim(p, t), abs=2e-5)
    assert M.psnr(pd, td) == pytest.approx(MO.psnr(p, t), rel=1e-5)
    got4 = M._hit_miss_fa_cn(pd, td, th)
    assert all(isinstance(v, torch.Tensor) and v.ndim == 0 and v.dtype == torch.float32 and v.is_cuda for v in got4)
    assert [float(v) for v in MO._hit_miss_fa_cn(p, t, th)] == [float(v) for v in got4]
    with pytest.raises(ValueError):
        M.csi(pd, td, th, "median", 4)


def test_metrics_generic_pools_and_ensembles_vs_reference_golden(lib):
    """pool_type='max', any scale, ensemble forecasts: the generic kernels against values computed by the UNMODIFIED
    reference module (tests/golden/make_golden_metrics_extra.py). CSI / HSS bit-exact (exact counts below 2**24 are
    what the reference's float32 sums hold too), CRPS to float32 rounding."""
    import json
    import os
    import sys

    from conftest import GOLDEN
    from weatherforecastingtoolkit_b200 import metrics as M
    sys.path.insert(0, GOLDEN)
    from make_golden_metrics_extra import THS, inputs
    with open(os.path.join(GOLDEN, "metrics_extra_golden.json")) as f:
        gold = json.load(f)
    p, t, ens, gt = inputs()
    pd, td, ed, gd = p.to(DEV), t.to(DEV), ens.to(DEV), gt.to(DEV)
    for row in gold["pooled"]:
        pool, scale = row["pool_type"], row["scale"]
        assert M.crps(pd, td, pool, scale) == pytest.approx(row["crps"], rel=2e-6), (pool, scale)
        for th in THS:
            assert M.csi(pd, td, th, pool, scale) == row[f"csi_{th:.6f}"], (pool, scale, th)
            assert M.hss(pd, td, th, pool, scale) == row[f"hss_{th:.6f}"], (pool, scale, th)
    for row in gold["crps_ensemble"]:
        assert M.crps(ed, gd, row["pool_type"], row["scale"]) == pytest.approx(row["crps"], rel=2e-5), row
    assert [float(v) for v in M._hit_miss_fa_cn(pd, td, THS[1])] == gold["hit_miss_fa_cn"]
    got = M.calc_metrics(ed, gd)
    want = gold["ensemble_calc_metrics"]
    assert list(got) == list(want)
    for k, v in want.items():
        if k.startswith(("CSI", "HSS", "paper_CSI", "paper_HSS")):
            assert got[k] == v, k          # ensemble mean bit-identical to torch's pred.mean(dim=1), then exact counts
        elif "SSIM" in k:
            assert got[k] == pytest.approx(v, abs=2e-5), k
        else:
            assert got[k] == pytest.approx(v, rel=2e-5), k
    # the ensemble mean itself
    assert torch.equal(M.ensemble_mean(ed, clamp=True).cpu(), ens.clamp(0, 1).mean(dim=1))


def test_metric_accumulator_reference_semantics(lib):
    from weatherforecastingtoolkit_b200 import metrics as M
    p, t = metric_case_inputs("vil_2x12x384")
    acc = M.MetricAccumulator(reference_semantics=True)
    per = []
    for i in range(2):
        acc.update(p[i:i + 1].to(DEV), t[i:i + 1].to(DEV))
        per.append(M.calc_metrics(p[i:i + 1].to(DEV), t[i:i + 1].to(DEV)))
    got = acc.compute()
    for k in per[0]:
        assert got[k] == pytest.approx((per[0][k] + per[1][k]) / 2, rel=1e-12, abs=1e-15), k
    ratio_of_sums = M.calc_metrics(p.to(DEV), t.to(DEV))
    assert got["CSI_3"] != ratio_of_sums["CSI_3"]       # mean of per-batch ratios != ratio of summed counts


def test_metrics_edge_cases(lib):
    from oracle import metrics_oracle as MO
    from weatherforecastingtoolkit_b200 import metrics as M
    # smallest legal image, one frame
    torch.manual_seed(1)
    p, t = torch.rand(1, 1, 1, 16, 19), torch.rand(1, 1, 1, 16, 19)
    mp = M.metric_partials(p.to(DEV), t.to(DEV))
    assert mp.counts.tolist() == MO.integer_counts(p, t).tolist()
    assert mp.n_elems.tolist() == [16 * 19, 16, 1]
    assert mp.ssim_sum == pytest.approx(MO.partials(p, t)["ssim_sum"], abs=1e-5)
    # all-zero target: range 0 -> PSNR is -inf in torchmetrics' formula too
    z = torch.zeros(1, 2, 1, 64, 64)
    got = M.calc_metrics(torch.rand(1, 2, 1, 64, 64).to(DEV), z.to(DEV))
    assert got["PSNR"] == -math.inf and got["CSI_0"] == 0.0
    # identical inputs: SSIM 1, MAE 0
    x = torch.rand(2, 3, 1, 96, 80)
    got = M.calc_metrics(x.to(DEV), x.clone().to(DEV))
    assert got["SSIM"] == pytest.approx(1.0, abs=1e-6) and got["CRPS"] == 0.0 and got["CSI_2"] == pytest.approx(1.0)
    with pytest.raises(RuntimeError):
        M.calc_metrics(torch.rand(1, 1, 1, 8, 8).to(DEV), torch.rand(1, 1, 1, 8, 8).to(DEV))  # < 11 px
    with pytest.raises(RuntimeError):
        M.calc_metrics(x, x)  # CPU tensors: no fallback


def test_metrics_full_size_additivity(lib):
    """BASELINE size (32 sequences x 12 frames x 384^2 = 56.6 M pixels, beyond float32-exact counting,
    hazard H1): the counts of the whole batch equal the int64 sum of per-sequence counts, each of which is
    checked against the oracle, and tp+fn+fp+tn == number of cells."""
    from oracle import metrics_oracle as MO
    from weatherforecastingtoolkit_b200 import metrics as M
    from weatherforecastingtoolkit_b200.synthetic import make_vil_sequences
    u8 = make_vil_sequences(32, 384, 384, 13, seed=77).to(DEV)
    x = (u8.float() / 255).permute(0, 3, 1, 2).unsqueeze(2)
    p, t = x[:, :12].contiguous(), x[:, 1:13].contiguous()
    whole = M.metric_partials(p, t)
    parts = [M.metric_partials(p[i:i + 1], t[i:i + 1]) for i in range(32)]
    tot = parts[0]
    for q in parts[1:]:
        tot = tot + q
    assert np.array_equal(whole.counts, tot.counts)
    assert whole.counts[0].sum(axis=1).tolist() == [32 * 12 * 384 * 384] * 6
    assert whole.counts[2].sum(axis=1).tolist() == [32 * 12 * 24 * 24] * 6
    assert whole.counts[0].max() > 2 ** 24          # the regime where the reference's float32 sums round
    for i in (0, 31):
        assert parts[i].counts.tolist() == MO.integer_counts(p[i:i + 1].cpu(), t[i:i + 1].cpu()).tolist()
    assert whole.ssim_sum == pytest.approx(tot.ssim_sum, rel=1e-9)
    assert whole.abs_sum[0] == pytest.approx(tot.abs_sum[0], rel=1e-9)


def test_metrics_deterministic(lib):
    from weatherforecastingtoolkit_b200 import metrics as M
    p, t = metric_case_inputs("rand_2x10x64")
    a = M.metric_partials(p.to(DEV), t.to(DEV))
    b = M.metric_partials(p.to(DEV), t.to(DEV))
    assert np.array_equal(a.ints, b.ints) and np.array_equal(a.floats, b.floats)


# ------------------------------------------------------------------ row softmax (attention.py:171)
@pytest.mark.parametrize("rows,cols", [(37, 2304), (5, 4096), (9, 50), (3, 4100), (13, 256), (21, 1024), (8, 2048),
                                       (2 * 2304, 2304)])
def test_softmax_rows(lib, rows, cols):
    from weatherforecastingtoolkit_b200 import _cabi
    torch.manual_seed(rows + cols)
    s = torch.randn(rows, cols, device=DEV) * 7.0
    out = torch.empty(rows, cols, dtype=torch.float16, device=DEV)
    scale = 1.0 / math.sqrt(512)
    _cabi.check(lib.wfk_softmax_rows(s.data_ptr(), rows, cols, scale, out.data_ptr(), 0,
                                     torch.cuda.current_stream().cuda_stream), "softmax")
    want = torch.softmax(s.float().cpu() * scale, dim=-1)
    assert (out.float().cpu() - want).abs().max().item() < 1e-3   # fp16 output
    assert (out.float().sum(-1).cpu() - 1).abs().max().item() < 2e-3


def test_nhwc_to_nchw_f32(lib):
    from weatherforecastingtoolkit_b200 import _cabi
    x = torch.randn(3, 5, 7, 8, device=DEV)
    out = torch.empty(3, 8, 5, 7, device=DEV)
    _cabi.check(lib.wfk_nhwc_to_nchw_f32(x.data_ptr(), 3, 5 * 7, 8, out.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "transpose")
    assert torch.equal(out, x.permute(0, 3, 1, 2).contiguous())


# ------------------------------------------------------------------ tensor-core stem convolutions
@pytest.mark.parametrize("n,cin,h,w,cout,ones,cpg", [(2, 1, 64, 64, 128, 0, 4), (1, 1, 50, 70, 128, 0, 4),
                                                      (2, 4, 48, 48, 512, 1, 16), (1, 4, 13, 9, 256, 1, 8),
                                                      (1, 3, 20, 24, 128, 0, 8),
                                                      # single-plane lean kernel (w % 16 == 0): every group size, several
                                                      # rows per warp, two 128-channel chunks, the production frame
                                                      (1, 1, 32, 48, 256, 0, 8), (2, 1, 16, 16, 128, 0, 16),
                                                      (1, 1, 7, 16, 128, 0, 4), (1, 1, 384, 384, 128, 0, 4)])
def test_stem_tc_vs_conv2d(lib, n, cin, h, w, cout, ones, cpg):
    """wfk_conv3x3_stem_tc == conv3x3(pad 1) of (optionally) a 1x1 pre-convolution, on identical fp16 operands;
    GroupNorm sums of the fp32 result."""
    from weatherforecastingtoolkit_b200 import _cabi
    torch.manual_seed(n * 1000 + h * 10 + cin)
    x = torch.rand(n, cin, h, w)
    wt = torch.randn(cout, cin, 3, 3) / math.sqrt(9 * cin)
    bias = torch.randn(cout) * 0.1
    if ones:
        pw, pb = torch.randn(cin, cin) * 0.5, torch.randn(cin) * 0.3
        wk = torch.cat([torch.einsum("oirs,ij->ojrs", wt, pw), torch.einsum("oirs,i->ors", wt, pb).unsqueeze(1)], 1)
    else:
        wk = wt
    kk = wk.shape[1] * 9
    wp = torch.zeros((kk + 15) // 16 * 16, cout)
    wp[:kk] = wk.permute(1, 2, 3, 0).reshape(kk, cout)
    wp16 = wp.half()
    out = torch.empty(n, h, w, cout, dtype=torch.float16, device=DEV)
    groups = cout // cpg
    stats = torch.zeros(n, groups, 2, dtype=torch.float64, device=DEV)
    xd, wd, bd = x.to(DEV), wp16.to(DEV), bias.to(DEV)
    _cabi.check(lib.wfk_conv3x3_stem_tc(xd.data_ptr(), n, cin, h, w, ones, wd.data_ptr(), bd.data_ptr(), cout,
                                        out.data_ptr(), stats.data_ptr(), cpg, 0, torch.cuda.current_stream().cuda_stream), "stem")
    # reference on the SAME rounded operands: fp16 input planes (+ the ones plane), fp16 packed weights
    planes = x.half().float()
    if ones:
        planes = torch.cat([planes, torch.ones(n, 1, h, w)], 1)
    wref = wp16.float()[:kk].reshape(wk.shape[1], 3, 3, cout).permute(3, 0, 1, 2).contiguous()
    want = F.conv2d(planes, wref, bias, padding=1).permute(0, 2, 3, 1)
    assert rel_l2(out.float(), want) < 2e-3
    torch.cuda.synchronize()
    ws = want.reshape(n, h * w, groups, cpg).double()
    assert torch.allclose(stats[..., 0].cpu(), ws.sum(dim=(1, 3)), rtol=1e-3, atol=5e-2)
    assert torch.allclose(stats[..., 1].cpu(), (ws * ws).sum(dim=(1, 3)), rtol=1e-3, atol=5e-2)
    if ones:  # against the unfolded fp32 definition: conv3x3(zero-padded (1x1 conv of x))
        z = F.conv2d(x, pw.reshape(cin, cin, 1, 1), pb)
        want32 = F.conv2d(z, wt, bias, padding=1).permute(0, 2, 3, 1)
        assert rel_l2(out.float(), want32) < 5e-3


# ------------------------------------------------------------------ decoder tail: GroupNorm + SiLU + conv3x3(C -> 1)
@pytest.mark.parametrize("n,h,w,c,groups", [(2, 40, 72, 128, 32), (1, 16, 16, 128, 32), (1, 21, 35, 64, 32),
                                             (1, 33, 18, 256, 32), (1, 96, 160, 128, 32)])
def test_gn_silu_conv3x3_c1_vs_fp32(lib, n, h, w, c, groups):
    """wfk_gn_silu_conv3x3_c1 (vae.py:162-164: conv_norm_out, SiLU, conv_out) against F.group_norm / F.silu /
    F.conv2d in fp32 on the same fp16 input: interior tiles (table addressing), border tiles and ragged sizes."""
    from weatherforecastingtoolkit_b200 import _cabi
    torch.manual_seed(h * 100 + w + c)
    x = (torch.randn(n, h, w, c) * 1.5 + 0.3).half()
    gamma, beta = torch.randn(c) * 0.5 + 1.0, torch.randn(c) * 0.3
    wt = torch.randn(1, c, 3, 3) / math.sqrt(9 * c)
    bias = 0.1
    xf = x.float().permute(0, 3, 1, 2)
    cpg = c // groups
    gsum = xf.double().reshape(n, groups, cpg * h * w)
    stats = torch.stack([gsum.sum(-1), (gsum * gsum).sum(-1)], dim=-1).contiguous()
    want = F.conv2d(F.silu(F.group_norm(xf, groups, gamma, beta, 1e-6)), wt, torch.tensor([bias]), padding=1)
    out = torch.empty(n, 1, h, w, device=DEV)
    xd, sd, gd, bd = x.to(DEV), stats.to(DEV), gamma.to(DEV), beta.to(DEV)
    wd = wt[0].permute(1, 2, 0).reshape(9, c).contiguous().to(DEV)          # [tap][c] fp32
    _cabi.check(lib.wfk_gn_silu_conv3x3_c1(xd.data_ptr(), sd.data_ptr(), gd.data_ptr(), bd.data_ptr(), n, h, w, c, groups,
                                           1e-6, wd.data_ptr(), bias, out.data_ptr(), 0,
                                           torch.cuda.current_stream().cuda_stream), "tail")
    torch.cuda.synchronize()
    assert rel_l2(out, want) < 2e-3
    assert (out.cpu() - want).abs().max().item() < 2e-2


# ------------------------------------------------------------------ event windowing + staging
def test_stage_vil_windows_bitexact(lib):
    """Windows cut from resident uint8 events == the reference's slicing (sevir.py:879-889) + staging formula."""
    from oracle import akl_oracle as O
    from weatherforecastingtoolkit_b200.rollout import sequent_windows, stage_vil, stage_vil_windows
    g = torch.Generator().manual_seed(9)
    events = torch.randint(0, 256, (3, 48, 40, 49), generator=g, dtype=torch.uint8)
    wins = sequent_windows(3, raw_seq_len=49, seq_len=25, stride=12)
    assert wins == [(0, 0), (0, 12), (0, 24), (1, 0), (1, 12), (1, 24), (2, 0), (2, 12), (2, 24)]
    got = stage_vil_windows(events.to(DEV), wins, seq_len=25)
    ref_batch = torch.stack([events[e, :, :, t0:t0 + 25] for e, t0 in wins])          # the reference's sampled_seq
    want = O.stage_vil(ref_batch).permute(0, 3, 1, 2).unsqueeze(2).contiguous()
    assert torch.equal(got.cpu(), want)
    assert torch.equal(got, stage_vil(ref_batch.to(DEV)))                               # == slice-then-stage
    with pytest.raises(ValueError):
        stage_vil_windows(events.to(DEV), [(0, 30)], seq_len=25)
    with pytest.raises(ValueError):
        stage_vil_windows(events.to(DEV), [(3, 0)], seq_len=25)


# ------------------------------------------------------------------ epoch accumulator
def test_metric_accumulator_equals_concatenated_batch(lib):
    from weatherforecastingtoolkit_b200 import metrics as M
    p, t = metric_case_inputs("rand_2x10x64")
    acc = M.MetricAccumulator()
    acc.update(p[:1].to(DEV), t[:1].to(DEV))
    acc.update(p[1:].to(DEV), t[1:].to(DEV))
    whole = M.metric_partials(p.to(DEV), t.to(DEV))
    got = acc.partials()
    assert np.array_equal(got.ints, whole.ints)                       # exact integer counts, additive over batches
    assert np.allclose(got.floats[:6], whole.floats[:6], rtol=1e-9, atol=0), (got.floats, whole.floats)  # [6:] reserved
    a, b = acc.compute(extended=True), M.calc_metrics(p.to(DEV), t.to(DEV), extended=True)
    for k in b:
        both_nan = a[k] != a[k] and b[k] != b[k]
        assert both_nan or a[k] == b[k] or abs(a[k] - b[k]) <= 1e-7 * max(1.0, abs(b[k])), (k, a[k], b[k])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process(lib):
    """the library holds no 'current device'. Kernels that need > 48 KB of dynamic shared memory (the fused
    skill scores, the conv-GEMM) run on cuda:1 while cuda:0 stays the caller's current device, from the main thread
    and from a worker thread, and the caller's device is restored after every call."""
    import threading

    from oracle import metrics_oracle as MO
    from weatherforecastingtoolkit_b200 import _cabi
    from weatherforecastingtoolkit_b200 import metrics as M
    from weatherforecastingtoolkit_b200.models.autoencoderkl import AutoencoderKL
    from weatherforecastingtoolkit_b200.rollout import stage_vil
    from weatherforecastingtoolkit_b200.synthetic import PATHB_AKL_CONFIG, make_akl_state_dict
    torch.cuda.set_device(0)
    p, t = metric_case_inputs("rand_2x10x64")
    want = MO.integer_counts(p, t).tolist()
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        got = M.metric_partials(p.to(dev), t.to(dev))
        assert got.counts.tolist() == want
        assert torch.cuda.current_device() == 0
    u8 = torch.randint(0, 256, (1, 32, 32, 25), dtype=torch.uint8)
    assert torch.equal(stage_vil(u8.to("cuda:1")).cpu(), stage_vil(u8.to("cuda:0")).cpu())
    sd = make_akl_state_dict(PATHB_AKL_CONFIG, 0)
    outs = {}
    for dev in ("cuda:0", "cuda:1"):
        m = AutoencoderKL(**PATHB_AKL_CONFIG)
        m.load_state_dict(sd, strict=True)
        m = m.to(dev)
        x = (u8[..., :2].float() / 255).permute(0, 3, 1, 2).reshape(2, 1, 32, 32).to(dev)
        outs[dev] = m.decode(m.encode(x).mode()).cpu()
        assert torch.cuda.current_device() == 0
    assert torch.equal(outs["cuda:0"], outs["cuda:1"])     # same kernels, same fixed-order reductions
    res = {}

    def worker():
        res["counts"] = M.metric_partials(p.to("cuda:1"), t.to("cuda:1")).counts.tolist()

    th = threading.Thread(target=worker)
    th.start()
    th.join()
    assert res["counts"] == want
    assert _cabi.load().wfk_nonfinite_status(1, 0) == 0

"""SEVIR sampling contract: oracle vs the reference's own outputs (golden), the product's host
logic vs the oracle on CPU, and the staged device batches vs the golden bit for bit on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.loader_oracle import LoaderOracle  # noqa: E402
from weatherforecastingtoolkit_b200.datastage import DeviceSEVIRLoader, DeviceSEVIRTorchDataset, SamplePlan  # noqa: E402
from weatherforecastingtoolkit_b200.synthetic import make_loader_events  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "loader_golden.npz"))
E, H, W, T_RAW = 5, 8, 12, 49
CASES = {
    "b4_uneven_ntchw": dict(batch_size=4, layout="NTCHW", split_mode="uneven"),
    "b4_floor_nhwt": dict(batch_size=4, layout="NHWT", split_mode="floor"),
    "b2_ceil_nthwc_shard1of2": dict(batch_size=2, layout="NTHWC", split_mode="ceil", num_shard=2, rank=1),
    "b3_uneven_tnhw_shard0of2": dict(batch_size=3, layout="TNHW", split_mode="uneven", num_shard=2, rank=0),
    "b4_sevir_rescale": dict(batch_size=4, layout="NTHW", split_mode="uneven", rescale_method="sevir"),
    "b5_stride6_len13": dict(batch_size=5, layout="TNCHW", split_mode="uneven", seq_len=13, stride=6),
    "b4_shuffled": dict(batch_size=4, layout="NTCHW", split_mode="uneven", shuffle=True, shuffle_seed=3),
}


def events():
    return make_loader_events(E, H, W, T_RAW, seed=7).numpy()


def oracle_for(name):
    kw = dict(CASES[name])
    kw.pop("shuffle", None), kw.pop("shuffle_seed", None)
    if "rescale_method" in kw:
        kw["rescale"] = kw.pop("rescale_method")
    return LoaderOracle(events(), raw_seq_len=T_RAW, order=GOLD[f"{name}/order"].tolist(), **kw)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_loader(name):
    o = oracle_for(name)
    assert len(o) == int(GOLD[f"{name}/len"])
    nb = 0
    for batch, mask in o:
        np.testing.assert_array_equal(batch, GOLD[f"{name}/batch{nb}"])
        assert (mask is None) == bool(GOLD[f"{name}/mask{nb}_is_none"])
        if mask is not None:
            assert mask == GOLD[f"{name}/mask{nb}"].tolist()
        nb += 1
    assert nb == int(GOLD[f"{name}/num_batches"])
    np.testing.assert_array_equal(o.idx_sample(1), GOLD[f"{name}/idx1"])


@pytest.mark.parametrize("name", list(CASES))
def test_host_plans_match_oracle(name):
    """No GPU: order, shard bounds, len, use_up and the (event, window, mask) plan of every batch."""
    ld = DeviceSEVIRLoader(events(), raw_seq_len=T_RAW, **CASES[name])
    o = oracle_for(name)
    assert ld._order.tolist() == GOLD[f"{name}/order"].tolist()   # pandas' shuffle reproduced
    assert (ld.start_event_idx, ld.end_event_idx, len(ld)) == (o.start_event_idx, o.end_event_idx, len(o))
    e, s, nb = ld.start_event_idx, 0, 0
    while not ld._used_up_at(e, s):
        assert not o.use_up
        plan, e, s = ld.plan_sequent(e, s)
        picks, oe, os_ = o._walk(o.curr_event_idx, o.curr_seq_idx)
        o.curr_event_idx, o.curr_seq_idx = oe, os_
        assert plan.picks == picks and (e, s) == (oe, os_)
        gm = GOLD[f"{name}/mask{nb}"].tolist()
        assert (plan.mask is None) == bool(GOLD[f"{name}/mask{nb}_is_none"])
        assert plan.mask is None or plan.mask == gm
        for (slot, t0), (ev, sq) in zip(plan.windows, picks):
            assert plan.events[slot] == ev and t0 == sq * ld.stride
        nb += 1
    assert o.use_up and nb == int(GOLD[f"{name}/num_batches"])


def test_sample_plan_dedups_events():
    p = SamplePlan([(3, 1), (3, 2), (4, 0), (3, 0), (7, 2)], end_event_idx=5, stride=12)
    assert p.events == [3, 4, 7] and p.real == [True, True, False]
    assert p.windows == [(0, 12), (0, 24), (1, 0), (0, 0), (2, 24)]
    assert p.mask == [True, True, True, True, False]


def test_constructor_errors():
    ev = events()
    with pytest.raises(ValueError):
        DeviceSEVIRLoader(ev, raw_seq_len=T_RAW, layout="NCHW")
    with pytest.raises(ValueError):
        DeviceSEVIRLoader(ev, raw_seq_len=T_RAW, split_mode="round")
    with pytest.raises(AssertionError):
        DeviceSEVIRLoader(ev, raw_seq_len=T_RAW, seq_len=50)
    with pytest.raises(ValueError):
        DeviceSEVIRLoader(ev, raw_seq_len=25)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("prefetch", [True, False])
def test_device_batches_bitexact(name, prefetch):
    ld = DeviceSEVIRLoader(events(), raw_seq_len=T_RAW, prefetch=prefetch, **CASES[name])
    nb = 0
    for d in ld:
        got = d["vil"]
        assert got.is_cuda and got.dtype == torch.float32
        np.testing.assert_array_equal(got.cpu().numpy(), GOLD[f"{name}/batch{nb}"])
        assert (d["mask"] is None) == bool(GOLD[f"{name}/mask{nb}_is_none"])
        nb += 1
    assert nb == int(GOLD[f"{name}/num_batches"])
    np.testing.assert_array_equal(ld[1]["vil"].cpu().numpy(), GOLD[f"{name}/idx1"])
    # a second pass after reset() gives the same batches (unshuffled cases)
    if not CASES[name].get("shuffle"):
        ld.reset()
        np.testing.assert_array_equal(next(ld)["vil"].cpu().numpy(), GOLD[f"{name}/batch0"])
    # every event crossed PCIe as uint8, at most once per batch that uses it
    assert ld.h2d_bytes % (H * W * T_RAW) == 0


@pytest.mark.gpu
def test_device_dataset_items():
    ds = DeviceSEVIRTorchDataset(events(), raw_seq_len=T_RAW)
    assert len(ds) == int(GOLD["dataset/len"])
    for i in (0, 4, len(ds) - 1):
        got = ds[i]
        assert got.is_contiguous()
        np.testing.assert_array_equal(got.cpu().numpy(), GOLD[f"dataset/item{i}"])


@pytest.mark.gpu
def test_random_mode_draws():
    """Reference `_random_sample` does not terminate (sevir.py:786-794); the batch its two randint draws describe is
    checked against the oracle under the same numpy seed."""
    kw = dict(batch_size=6, layout="NTCHW", sample_mode="random")
    ld = DeviceSEVIRLoader(events(), raw_seq_len=T_RAW, **kw)
    o = LoaderOracle(events(), raw_seq_len=T_RAW, **kw)
    np.random.seed(11)
    got = next(ld)["vil"].cpu().numpy()
    np.random.seed(11)
    want, _ = next(o)
    np.testing.assert_array_equal(got, want)

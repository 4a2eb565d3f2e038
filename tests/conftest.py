import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def has_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "pipeline"))


@pytest.fixture(scope="session")
def golden_akl():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "akl_golden.npz")))


@pytest.fixture(scope="session")
def golden_metrics():
    import json
    with open(os.path.join(GOLDEN, "metrics_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def akl_weights():
    from weatherforecastingtoolkit_b200.synthetic import PATHB_AKL_CONFIG, make_akl_state_dict
    return PATHB_AKL_CONFIG, make_akl_state_dict(PATHB_AKL_CONFIG, seed=0, affine_jitter=0.1)


def metric_case_inputs(name):
    """Regenerate the (pred, target) pair of a metrics golden case from its seed."""
    import torch
    from weatherforecastingtoolkit_b200.synthetic import make_vil_sequences
    if name == "rand_2x10x64":
        torch.manual_seed(0)
        return torch.rand(2, 10, 1, 64, 64), torch.rand(2, 10, 1, 64, 64)
    if name == "rand_2x12x384":
        torch.manual_seed(0)
        return torch.rand(2, 12, 1, 384, 384), torch.rand(2, 12, 1, 384, 384)
    if name == "unclamped_1x3x50x70":
        torch.manual_seed(5)
        return torch.randn(1, 3, 1, 50, 70) * 0.6 + 0.4, torch.randn(1, 3, 1, 50, 70) * 0.6 + 0.4
    if name == "vil_2x12x384":
        u8 = make_vil_sequences(2, 384, 384, 13, seed=31)
        x = ((1 / 255) * u8.float()).permute(0, 3, 1, 2).unsqueeze(2)
        return x[:, :12].contiguous(), x[:, 1:13].contiguous()
    raise KeyError(name)


METRIC_CASES = ["rand_2x10x64", "rand_2x12x384", "unclamped_1x3x50x70", "vil_2x12x384"]

"""Validation-panel rendering: byte colour tables vs the oracle's restatement of matplotlib on CPU,
the render kernel vs the oracle bit for bit on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import render_oracle as R  # noqa: E402
from weatherforecastingtoolkit_b200 import render  # noqa: E402

ALL_BYTES = np.arange(256, dtype=np.uint8)


def test_vil_table_matches_oracle_and_reference_constants():
    cmap, norm, vmin, vmax = render.vil_cmap()
    assert norm is None and vmin is None and vmax is None
    np.testing.assert_array_equal(cmap(ALL_BYTES), R.vil_rgba(ALL_BYTES))
    assert render.VIL_COLORS == R.VIL_COLORS and render.VIL_LEVELS == R.VIL_LEVELS
    # known answers straight from the reference's tables (sevir.py:1237-1250): bin edges are inclusive on the left,
    # values below 16 take the `under` colour (= first listed colour), 255 the `over` colour (= last)
    t = cmap(ALL_BYTES)
    assert t[0].tolist() == [77, 77, 77, 255] and t[15].tolist() == t[0].tolist()
    assert t[16].tolist() == [40, 190, 40, 255] and t[30].tolist() == t[16].tolist()
    assert t[219].tolist() == [231, 0, 255, 255] and t[255].tolist() == t[219].tolist()
    assert len({tuple(r) for r in t}) == 10


def test_diff_table_matches_oracle():
    t = render.diff_cmap()(ALL_BYTES)
    np.testing.assert_array_equal(t, R.reds_rgba(ALL_BYTES))
    assert t[0].tolist() == [255, 245, 240, 255] and t[255].tolist() == [103, 0, 12, 255]
    assert (np.diff(t[:, 1].astype(int)) <= 0).all()  # green channel of Reds decreases monotonically


def test_oracle_quantisation_edges():
    x = np.array([-1.0, 0.0, 0.5, 1.0, 2.0, np.nan, 16 / 255, 15.999 / 255], dtype=np.float32)
    assert R.quantise(x).tolist() == [0, 0, 127, 255, 255, 0, int(np.float32(16 / 255) * np.float32(255)), 15]
    assert R.abs_diff(np.array([3, 200], np.uint8), np.array([250, 10], np.uint8)).tolist() == [247, 190]


def _inputs(shape, seed):
    g = torch.Generator().manual_seed(seed)
    p = torch.rand(shape, generator=g) * 1.4 - 0.2
    t = torch.rand(shape, generator=g) * 1.4 - 0.2
    flat_t = t.view(-1)
    k = min(flat_t.numel(), 256)
    flat_t[:k] = torch.arange(k, dtype=torch.float32) / 255.0   # exact k/255 values sit on the truncation edge
    p.view(-1)[:: max(1, p.numel() // 7)] = float("nan")
    return p, t


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 3, 1, 48, 40), (1, 1, 7, 9), (5,), (3, 12, 384, 384)])
def test_render_panels_bitexact(shape):
    p, t = _inputs(shape, seed=sum(shape))
    got = render.render_panels(p.cuda(), t.cuda())
    want = R.render_panels(p.numpy(), t.numpy())
    for k, v in want.items():
        assert got[k].shape == v.shape and got[k].dtype == torch.uint8
        np.testing.assert_array_equal(got[k].cpu().numpy(), v, err_msg=k)
    only_u8 = render.render_panels(p.cuda(), t.cuda(), rgba=False)
    assert set(only_u8) == {"target_u8", "pred_u8", "diff_u8"}
    np.testing.assert_array_equal(only_u8["diff_u8"].cpu().numpy(), want["diff_u8"])


@pytest.mark.gpu
def test_log_wandb_images_mosaics():
    p, t = _inputs((6, 4, 1, 32, 24), seed=5)

    class Experiment:
        def __init__(self):
            self.logged = []

        def log(self, d):
            self.logged.append(d)

    class Logger:
        experiment = Experiment()

    class Module:
        logger, global_step = Logger(), 17

    mos = render.log_wandb_images(p.cuda(), t.cuda(), "val", Module(), batch_idxs=2)
    assert len(mos) == 2 and mos[0].shape == (3 * 32, 4 * 24, 4) and mos[0].dtype == np.uint8
    want = R.render_panels(p.numpy()[:2, :, 0], t.numpy()[:2, :, 0])
    for b in range(2):
        for row, key in enumerate(("target_rgba", "pred_rgba", "diff_rgba")):
            for ti in range(4):
                np.testing.assert_array_equal(mos[b][row * 32:(row + 1) * 32, ti * 24:(ti + 1) * 24], want[key][b, ti])
    assert len(Module.logger.experiment.logged) == 2 and Module.logger.experiment.logged[0]["global_step"] == 17
    with pytest.raises(ValueError):
        render.render_panels(p.cuda(), t.cuda()[:1])
    with pytest.raises(RuntimeError):
        render.render_panels(p, t)

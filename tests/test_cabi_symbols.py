"""The C-ABI library builds, loads and exports every symbol include/wfk_b200.h declares
(no compute calls: this runs without a GPU)."""
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from weatherforecastingtoolkit_b200 import build, _cabi
    build.build()
    return _cabi.load()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "wfk_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wfk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/wfk_b200.h but not exported"


def test_binding_covers_header(lib):
    from weatherforecastingtoolkit_b200 import _cabi
    assert sorted(_cabi.EXPORTED_SYMBOLS) == _declared_functions()


def test_abi_version(lib):
    from weatherforecastingtoolkit_b200 import _cabi
    assert lib.wfk_abi_version() == _cabi.ABI_VERSION == 3
    # the header, the library and the ctypes mirror must move together
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "wfk_b200.h")).read()
    assert int(re.search(r"#define WFK_ABI_VERSION (\d+)", hdr).group(1)) == _cabi.ABI_VERSION
    assert lib.wfk_strerror(0) == b"ok"
    assert b"invalid" in lib.wfk_strerror(-1)


def test_struct_sizes():
    import ctypes as C
    from weatherforecastingtoolkit_b200 import _cabi
    assert C.sizeof(_cabi.MetricPartials) == 108 * 8
    assert C.sizeof(_cabi.Tap) == 12
    assert C.sizeof(_cabi.View5) == 88 and C.sizeof(_cabi.View3) == 56


def test_no_silent_fallback_without_gpu(lib):
    """Without a B200 the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from weatherforecastingtoolkit_b200 import _cabi, metrics
    with pytest.raises(RuntimeError):
        _cabi.init(0)
    with pytest.raises(RuntimeError):
        metrics.calc_metrics(torch.rand(1, 2, 1, 64, 64), torch.rand(1, 2, 1, 64, 64))
    # calls before wfk_init are refused by the library itself
    assert lib.wfk_softmax_rows(None, 1, 8, 1.0, None, 0, None) < 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "weatherforecastingtoolkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"

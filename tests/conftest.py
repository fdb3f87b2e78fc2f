import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


_device_state = {}


def _device_error():
    """None if the CUDA library finds an sm_100 device, else the reason (checked once per session)."""
    if "err" not in _device_state:
        try:
            from mlb_odds_engine_b200._lib import Context

            Context(0).close()
            _device_state["err"] = None
        except Exception as e:  # MlbsimError (no device / wrong arch) or ImportError (library not built)
            _device_state["err"] = str(e)
    return _device_state["err"]


def pytest_runtest_setup(item):
    """GPU tests never silently pass or skip on a CPU box: they are deselected by `-m "not gpu"` and FAIL when selected
    without a device (there is no CPU fallback to fall back to)."""
    if item.get_closest_marker("gpu") is not None:
        err = _device_error()
        if err is not None:
            pytest.fail(f"test is marked gpu and no usable CUDA device / library was found: {err}", pytrace=False)

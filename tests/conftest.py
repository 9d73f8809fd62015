import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(autouse=True)
def _reset_recogniser_dropout_epoch(request):
    """The recogniser trainers advance a device-side dropout epoch inside their CUDA graphs (include/focr.h: focr_recog_epoch_*); a
    later test in the same process that compares masks with the oracle's RNG twin needs the documented default, epoch 0."""
    if "gpu" in request.keywords:
        try:
            import torch
            if torch.cuda.is_available():
                from fudanocr_b200 import _lib as L
                L.check(L.lib.focr_recog_epoch_set(0, L.cur_stream()))
        except Exception:  # pragma: no cover - the test itself will report a missing library
            pass
    yield

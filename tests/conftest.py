import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _occupancy_checked_with_every_grid_comparison(request, monkeypatch):
    """Whenever a GPU test takes the device grid (dump or digest) to compare it with the oracle's, the occupancy tiles
    and their summary words - what the march reads instead of the cell lists - are checked against the cell entries too
    (stg_rt_debug_occupancy_check): no bit without an entry, no entry without its bits, summaries equal to the masks."""
    if "gpu" not in request.keywords:
        return
    from stage_b200 import rt

    def checked(method):
        def wrapper(self, *a, **k):
            bad_bits, bad_summaries, _ = self.occupancy_check()
            assert (bad_bits, bad_summaries) == (0, 0), "occupancy tiles / summaries disagree with the cell entries"
            return method(self, *a, **k)
        return wrapper

    for name in ("grid_dump", "dump_grid", "grid_hash"):
        if hasattr(rt.Context, name):
            monkeypatch.setattr(rt.Context, name, checked(getattr(rt.Context, name)))

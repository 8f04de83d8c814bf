"""The device's sin/cos (the port of the host libm in stage_b200/csrc/rt_libm.cuh, as the ray march
uses it) against the C library of the box the test runs on, bit for bit."""
import numpy as np
import pytest

import oracle_lib
from stage_b200 import rt

pytestmark = pytest.mark.gpu


def test_device_trig_equals_host_libm():
    ctx = rt.Context(50.0, -1, -1, 2, 2, max_models=4, max_blocks=4, max_cell_entries=1024)
    rng = np.random.RandomState(2)
    x = np.concatenate([rng.uniform(-7, 7, 2_000_000), np.float32(rng.uniform(-7, 7, 2_000_000)).astype(np.float64),
                        rng.uniform(-1e3, 1e3, 500_000), 10.0 ** rng.uniform(-12, 8, 500_000),
                        [0.0, 1e-12, np.pi, -np.pi, np.pi / 2, 0.126, 0.855469, 2.426265]])
    dev = ctx.debug_trig(x)
    host = oracle_lib.libm_trig(x)
    bad = int(np.count_nonzero(dev.view(np.uint64) != host.view(np.uint64)))
    assert bad == 0, "%d of %d cos/sin values differ from the host libm" % (bad, dev.size)
    ctx.close()

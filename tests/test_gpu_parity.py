"""Parity of the CUDA path (through the C ABI) with the reference.

Three anchors, in decreasing strength:
  * the T1 oracle run live on this host (bit-exact in strict-trig mode: ranges, hit model, hit
    cell, and the C / R step counters, which proves the kernel walks the reference's own path),
  * the golden answers recorded from the unmodified reference (hit model exact, range <= 1e-5 m;
    bit-exact when this host's libm is the one the traces were recorded with),
  * the device grid dumped cell for cell against the oracle's grid.
"""
import os

import numpy as np
import pytest

import oracle_lib
import replay
import trace_io
from stage_b200 import rt

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RANGE_TOL = 1e-5  # metres; BASELINE.json north_star


@pytest.fixture(scope="module", params=["simple_120", "fasr_300", "everything_40", "pioneer_flocking_12"])
def trace(request):
    return trace_io.load(os.path.join(GOLDEN, request.param + ".trc.xz"))


@pytest.fixture(scope="module")
def t1(trace):
    return replay.T1Replay(trace)


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def test_strict_trig_is_bit_exact_against_oracle_and_golden(trace, t1):
    g = replay.GpuReplay(trace, strict=True)
    n = len(trace.rays)
    assert g.n_ticks == trace.n_ticks and n > 30000
    # vs the oracle on this host: everything identical
    assert np.array_equal(g.hit_model, t1.hits["hit_model"])
    assert np.array_equal(bits(g.ranges), bits(t1.hits["range"]))
    hit = g.hit_model >= 0
    assert hit.sum() > n // 4
    assert np.array_equal(g.hit_cell[hit, 0], t1.hits["hit_x"][hit])
    assert np.array_equal(g.hit_cell[hit, 1], t1.hits["hit_y"][hit])
    assert np.array_equal(g.steps[:, 0], t1.hits["cell_visits"])
    assert np.array_equal(g.steps[:, 1], t1.hits["region_probes"])
    # vs the reference's recorded answers
    assert np.array_equal(g.hit_model, trace.rays["hit_model"])
    assert np.max(np.abs(g.ranges - trace.rays["res_range"])) <= RANGE_TOL
    # whole ranger scans: headings recurrence + fused epilogue
    assert len(g.scans) == len(t1.scans) > 0
    for (ev, sc, out), (rg, it, be, hits) in zip(g.scans, t1.scans):
        assert np.array_equal(bits(out.ranges), bits(rg))
        assert np.array_equal(bits(out.intensities), bits(it))
        assert np.array_equal(bits(out.bearings), bits(be))
        assert np.array_equal(out.hit_model, hits["hit_model"])
        assert np.max(np.abs(out.ranges - sc["ranges"])) <= RANGE_TOL
        assert np.array_equal(out.intensities, sc["intensities"])
        assert np.array_equal(bits(out.bearings), bits(sc["bearings"]))
    g.ctx.close()


def test_device_trig_matches_model_and_cell_exactly(trace, t1):
    """Default mode: the device computes sin/cos itself with the port of the host libm
    (rt_libm.cuh), so on a glibc x86-64 FMA host it is bit-identical too (asserted by
    test_gpu_trig.py and test_gpu_synthetic.py); here only the north-star bar is required."""
    g = replay.GpuReplay(trace, strict=False)
    assert np.array_equal(g.hit_model, trace.rays["hit_model"])
    hit = g.hit_model >= 0
    assert np.array_equal(g.hit_cell[hit, 0], t1.hits["hit_x"][hit])
    assert np.array_equal(g.hit_cell[hit, 1], t1.hits["hit_y"][hit])
    err = np.abs(g.ranges - trace.rays["res_range"])
    assert err.max() <= RANGE_TOL
    # informational: how many ranges differ in the last bits when the device computes sin/cos
    differing = int(np.count_nonzero(bits(g.ranges) != bits(trace.rays["res_range"])))
    print("device trig: %d of %d ranges differ in bits, max abs err %.3g m" % (differing, len(err), err.max()))
    for (ev, sc, out) in g.scans:
        assert np.max(np.abs(out.ranges - sc["ranges"])) <= RANGE_TOL
        assert np.array_equal(out.intensities, sc["intensities"])
    g.ctx.close()


def test_device_grid_equals_oracle_grid_cell_for_cell(trace, t1):
    g = replay.GpuReplay(trace, strict=True, want=rt.WANT_RANGES)
    rec_g, cnt_g = g.ctx.grid_dump()
    rec_g, cnt_g = oracle_lib.sort_grid(rec_g), oracle_lib.sort_counts(cnt_g)
    rec_o, cnt_o = t1.world.dump_grid()
    assert len(rec_o) > 10000
    assert np.array_equal(rec_g, rec_o)
    assert np.array_equal(cnt_g, cnt_o)
    g.ctx.close()


def test_one_sync_per_run_gives_the_same_answers(trace, t1):
    """Batching across ticks must not change results: map/unmap issued while fans are queued flush
    them first (issue-order rule of include/stg_rt.h)."""
    g = replay.GpuReplay(trace, strict=True, max_ticks=30, sync_every_tick=False, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL)
    n = 0
    ticks = 0
    for ev in trace.events:
        if ev["type"] == trace_io.EV_RAYS:
            n = int(ev["first"]) + int(ev["model"])
        if ev["type"] == trace_io.EV_TICK:
            ticks += 1
            if ticks >= 30:
                break
    assert np.array_equal(g.hit_model[:n], t1.hits["hit_model"][:n])
    assert np.array_equal(bits(g.ranges[:n]), bits(t1.hits["range"][:n]))
    g.ctx.close()


def test_large_world_extent_replays_exactly():
    """large.world, the widest extent the reference ships (50 x 36 superregions = 1.8 M regions, 1.9 M cell entries per
    layer from one traced bitmap, outlines of up to 1765 vertices), 20 ticks of its ten laser robots: rays, scans and
    the device grid against the oracle and the reference's recorded answers. Most of these rays end in open space
    after hundreds of empty-region crossings far from the origin."""
    trace = trace_io.load(os.path.join(GOLDEN, "large_20.trc.xz"))
    t1 = replay.T1Replay(trace)
    g = replay.GpuReplay(trace, strict=True)
    n = len(trace.rays)
    assert n == 36000 and g.n_ticks == 20
    assert np.array_equal(g.hit_model, t1.hits["hit_model"]) and np.array_equal(g.hit_model, trace.rays["hit_model"])
    assert np.array_equal(bits(g.ranges), bits(t1.hits["range"]))
    assert np.max(np.abs(g.ranges - trace.rays["res_range"])) <= RANGE_TOL
    hit = g.hit_model >= 0
    assert 0.05 < hit.mean() < 0.5
    assert np.array_equal(g.hit_cell[hit, 0], t1.hits["hit_x"][hit]) and np.array_equal(g.hit_cell[hit, 1], t1.hits["hit_y"][hit])
    assert np.array_equal(g.steps[:, 0], t1.hits["cell_visits"]) and np.array_equal(g.steps[:, 1], t1.hits["region_probes"])
    assert len(g.scans) == len(t1.scans) == 200
    for (ev, sc, out), (rg, it, be, hits) in zip(g.scans, t1.scans):
        assert np.array_equal(bits(out.ranges), bits(rg)) and np.array_equal(bits(out.intensities), bits(it))
        assert np.array_equal(out.hit_model, hits["hit_model"])
        assert np.max(np.abs(out.ranges - sc["ranges"])) <= RANGE_TOL
    rec_g, cnt_g = g.ctx.grid_dump()
    rec_o, cnt_o = t1.world.dump_grid()
    assert len(rec_o) > 3000000
    assert np.array_equal(oracle_lib.sort_grid(rec_g), rec_o) and np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
    g.ctx.close()

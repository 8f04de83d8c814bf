"""BASELINE configs 4 and 5 in miniature, through the same code bench.py runs at full size (bench_swarm.py): a swarm with
bodies in the grid, sonar rings and fiducial finders; pucks that all move every tick under scanning lasers. The records'
`parity` blocks compare the device with the T1 oracle: rays bit for bit, fiducial records, the digest of the whole grid."""
import argparse
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def small_args(**kw):
    a = argparse.Namespace(steps=3, warmup=3, rects=300, half_extent=20.48, robots=40, c4_robots=3000, c4_max_fiducials=64, c5_pucks=800, deltas="poses")
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def make_env():
    import bench_swarm
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    return bench_swarm.Env(torch, None, 0, 1, 0, stream, 6450.3)


@pytest.mark.parametrize("deltas", ["poses", "outlines"])
def test_config4_swarm_record_is_exact(deltas):
    import bench_swarm
    # (9000 robots by pose: the library's host threads queue, size and write a swarm's tick side by side from 8192 models up)
    rec = bench_swarm.run_c4(make_env(), small_args(deltas=deltas, **({"c4_robots": 9000, "rects": 100} if deltas == "poses" else {})))
    p = rec["parity"]
    assert p["mismatches_total"] == 0, p
    assert p["rays_checked"] >= 10000 and p["fiducials_checked"] > 200 and p["grid_digest_equals_oracle_on_every_rank"]
    assert p["fiducial_max_abs_err"] <= p["fiducial_tolerance"]
    assert rec["value"] > 0 and rec["e2e"]["value"] > 0 and rec["roofline"]["frac"] > 0 and rec["gpu_launches"] > 0
    assert rec["e2e"]["fiducials_seen_per_step"] > 1000


@pytest.mark.parametrize("deltas", ["poses", "outlines"])
def test_config5_moving_pucks_record_is_exact(deltas):
    import bench_swarm
    rec = bench_swarm.run_c5(make_env(), small_args(deltas=deltas), 270.0, 1080, 30.0)
    p = rec["parity"]
    assert p["mismatches_total"] == 0, p
    assert p["rays_checked"] >= 10000 and p["grid_digest_equals_oracle_on_every_rank"] and p["rays_that_hit_a_puck"] > 0
    assert rec["value"] > 0 and rec["e2e"]["value"] > 0 and rec["k1"]["blocks_per_s"] > 0 and rec["roofline"]["frac"] > 0

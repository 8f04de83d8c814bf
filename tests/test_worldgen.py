"""The synthetic-world generator must rasterise exactly what the reference would: the same world
is loaded into the unmodified reference as world-file text and the polygons its Block::Map calls
received are compared with worldgen.pixel_polygon vertex for vertex."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib
import trace_io
from stage_b200 import worldgen

TESTS = os.path.dirname(os.path.abspath(__file__))

CHILD = r"""
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle_lib
from stage_b200 import worldgen
w = worldgen.make_world(seed=int(sys.argv[1]), n_rects=int(sys.argv[2]), n_robots=int(sys.argv[3]), half_extent_m=float(sys.argv[4]))
ref = oracle_lib.RefWorld(text=worldgen.to_worldfile_text(w))
ref.trace_end(sys.argv[5])
sys.stdout.flush(); os._exit(0)
"""


def test_sreg_extent():
    w = worldgen.make_world(seed=3, n_rects=10, n_robots=5, half_extent_m=81.92)
    assert w.sreg_extent() == (-4, -4, 8, 8)
    w = worldgen.make_world(seed=3, n_rects=10, n_robots=5, half_extent_m=10.0)
    assert w.sreg_extent() == (-1, -1, 2, 2)


def test_robots_are_clear_of_obstacles_and_each_other():
    w = worldgen.make_world(seed=1, n_rects=300, n_robots=200, half_extent_m=20.48)
    o, r = w.obstacles, w.robots
    for x, y, _ in r:
        inside = (abs(x - o[:, 0]) < o[:, 2] / 2 + 0.5) & (abs(y - o[:, 1]) < o[:, 3] / 2 + 0.5)
        assert not inside.any()
    d = np.hypot(r[:, None, 0] - r[None, :, 0], r[:, None, 1] - r[None, :, 1]) + np.eye(len(r)) * 10
    assert d.min() >= 0.6


@pytest.mark.skipif(not oracle_lib.ref_available(), reason="oracle/_ref not built")
def test_polygons_and_ids_equal_the_reference(tmp_path):
    args = ("7", "150", "60", "20.48")
    out = tmp_path / "syn.trc"
    subprocess.run([sys.executable, "-c", CHILD % (TESTS, os.path.dirname(TESTS)), *args, str(out)], check=True,
                   stdout=subprocess.DEVNULL)
    tr = trace_io.load(out)
    w = worldgen.make_world(seed=7, n_rects=150, n_robots=60, half_extent_m=20.48)
    assert tr.ppm == w.ppm
    polys = worldgen.obstacle_polygons(w) + worldgen.robot_polygons(w)
    ids = np.concatenate([w.obstacle_ids, w.robot_ids])
    zmax = [w.obstacle_height] * len(w.obstacles) + [w.robot_size[2]] * len(w.robots)
    maps = tr.events[tr.events["type"] == trace_io.EV_MAP]
    by_model = {}
    for ev in maps:
        by_model.setdefault(int(ev["model"]), []).append(ev)
    assert sorted(by_model) == sorted(int(i) for i in ids)
    for mid, poly, zm in zip(ids, polys, zmax):
        evs = by_model[int(mid)]
        # a model is re-rendered several times while loading (SetGeom, SetPose, the final
        # UnMap/Map of LoadWorldPostHook); what stays in the grid is the last Map per layer
        last = {int(e["layer"]): e for e in evs}
        assert sorted(last) == [0, 1]
        for e in last.values():
            assert np.array_equal(tr.polygon(e), poly), (mid, tr.polygon(e), poly)
            assert e["zmin"] == 0.0 and e["zmax"] == zm
    # the reference's model table: every body is its own root with ranger_return 0.5
    m = {int(x["id"]): x for x in tr.models}
    for mid in ids:
        assert m[int(mid)]["root"] == mid and m[int(mid)]["ranger_return"] == 0.5

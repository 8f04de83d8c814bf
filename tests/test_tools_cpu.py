"""Small host-side pieces the measurements lean on, checked without a GPU: the launch-list summariser against the committed
ncu launch list, and the K1 roofline's cell-visit count against the oracle's own MapPoly."""
import os
import subprocess
import sys

import numpy as np

import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_summary_reads_the_committed_launch_list():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(ROOT, "profiles", "r02b_launches.csv")],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    text = out.stdout
    assert "**whole run**" in text and "**part 3**" in text          # config 3, then 4, then 5
    for kernel in ("k2_ray_march<0>", "k1_pose_phase0", "k1_summaries", "k3_fiducials_grid", "k3_fid_pack"):
        assert "`%s`" % kernel in text, kernel
    # every kernel of the list is one of this repository's, apart from torch's fill (the bench's L2 flush)
    names = {ln.split("`")[1] for ln in text.splitlines() if ln.startswith("| `")}
    assert all(n.startswith(("k1_", "k2_", "k3_", "k4_", "k5_", "at::")) for n in names), names


def test_outline_cell_visits_is_what_map_poly_visits():
    sys.path.insert(0, ROOT)
    import bench_swarm
    from stage_b200 import worldgen
    w = worldgen.make_swarm_world(seed=3, n_rects=50, n_robots=400, half_extent_m=10.24)
    boxes = worldgen.robot_boxes(w, w.robots)
    t1 = oracle_lib.T1World(w.ppm)
    for j in range(len(boxes)):
        t1.model_upsert(1 + j, 1 + j, 1.0)
        t1.block_map(j, 1 + j, 0, 0.0, 0.1, boxes[j])
    rec, cnt = t1.dump_grid()
    assert int(cnt[:, 2].sum()) == bench_swarm.outline_cell_visits(boxes) == len(rec)   # Region::count sums the visits (duplicates included)

"""A small pass through every kernel the round's second session touched - a quick exercise after a kernel change, and
what to run under `compute-sanitizer --tool memcheck` / racecheck where the pool allows it (this one does not): a swarm
moved by pose (K1 pose phases with three warps per record, k1_summaries on its stream), laser fans through walls and
through the robots' own regions (the march's closed forms and jumps, the per-stretch streamed delivery), sonar rings
(k2_expand_rangers), fiducial finders through the grid path (bins, row cut, sift / sort / trace, k3_fid_fans, pack),
collisions, the grid digest and the occupancy check. Prints what it ran; exit code 0 = the library saw no error.

    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_check.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("STG_RT_FID_GRID", "1")
from stage_b200 import rt, worldgen  # noqa: E402


def main():
    w = worldgen.make_world(seed=2, n_rects=60, n_robots=130, half_extent_m=10.24)
    x0, y0, nx, ny = w.sreg_extent()
    n_ob, n = len(w.obstacles), len(w.robots)
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + 1, max_blocks=n_ob + n, max_cell_entries=1 << 17)
    worldgen.load_into_context(ctx, w)
    worldgen.define_boxes(ctx, n_ob, w.robot_ids, w.robot_size)
    for k, mid in enumerate(w.robot_ids):
        ctx.model_upsert(int(mid), int(mid), 0.1 + 0.05 * (k % 7))
    ctx.sync()
    poses = w.robots
    out = rt.Outputs(n * 1080, rt.WANT_RANGES | rt.WANT_INTENSITIES, alloc=ctx.pinned_empty)   # 140,400 beams: streamed delivery
    for t in range(1, 4):
        poses = worldgen.step_robots(w, poses, t, speed_m=0.2)
        ctx.models_pose_update(w.robot_ids, worldgen.gposes(poses), t % 2)
        fans = worldgen.ranger_fans(w, layer=t % 2, samples=1080, range_max=12.0, robots=poses)
        ctx.fans_submit(fans, out)
        ctx.sync()
        tg, fd = worldgen.fiducial_records(w, poses, w.robot_ids, layer=t % 2, range_max=4.0, range_max_id=2.0)
        rec, cnt = ctx.fiducials_submit(tg, fd, max_per_finder=8)
        ctx.sync()
    g = ctx.trace(fans[:20], want=rt.WANT_RANGES | rt.WANT_HIT_MODEL | rt.WANT_HIT_CELL | rt.WANT_STEPS)
    sensors = worldgen.p2dx_sonar_sensors()
    so = rt.Outputs(n * len(sensors), rt.WANT_RANGES | rt.WANT_INTENSITIES)
    ctx.rangers_submit(sensors, worldgen.ranger_poses(poses, w.robot_ids, layer=1), so)
    ctx.sync()
    blocks = (n_ob + np.arange(n)).astype(np.uint32)
    hit = ctx.collisions_test(1, (np.arange(n + 1, dtype=np.uint32), blocks))
    ctx.sync()
    bad = ctx.occupancy_check()
    digest = ctx.grid_hash()
    print("beams %d hit %.2f | fiducials seen %d | sonars %d | collisions %d | occupancy (bad bits, bad summaries, cells) %s | entries %s"
          % (len(out.ranges), float((out.ranges < 12.0).mean()), int(np.minimum(cnt, 8).sum()), len(so.ranges), int((np.asarray(hit) >= 0).sum()),
             bad, digest[1]))
    assert bad[0] == 0 and bad[1] == 0 and int(g.steps[:, 0].sum()) > 0
    ctx.close()


if __name__ == "__main__":
    main()

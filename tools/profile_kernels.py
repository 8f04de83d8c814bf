"""One launch of every kernel of the path at its benchmark size, between cudaProfilerStart / Stop, for
`ncu --profile-from-start off --set full`: the config-3 world (1000 robots: K1 through edge records and through poses,
the 1.08 M-beam march, K4 / K4b / K5 over all robots) and the config-4 swarm (100 k robots: K1 through poses, the 1.6 M
one-beam sonar march, K3 over 100 k finders x 100 k targets). Prints what it launched; run without ncu it is a smoke test.

    python tools/profile_kernels.py [--c4-robots 100000]
"""
import argparse
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stage_b200 import rt, worldgen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c4-robots", type=int, default=100000)
    ap.add_argument("--skip-c4", action="store_true")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # ---- config 3
    w = worldgen.make_world(seed=1, n_rects=4096, n_robots=1000, half_extent_m=81.92)
    x0, y0, nx, ny = w.sreg_extent()
    n_ob = len(w.obstacles)
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + 1, max_blocks=n_ob + 1000, max_cell_entries=1 << 22)
    worldgen.load_into_context(ctx, w)
    worldgen.define_boxes(ctx, n_ob, w.robot_ids, w.robot_size)
    blocks = (n_ob + np.arange(1000)).astype(np.uint32)
    z = np.stack([np.zeros(1000), np.full(1000, w.robot_size[2])], 1)
    first = np.arange(1001, dtype=np.uint32) * 4
    poses = [w.robots]
    for t in range(6):
        poses.append(worldgen.step_robots(w, poses[-1], t))
    fans = worldgen.ranger_fans(w, layer=1, robots=poses[2])
    fset = ctx.set_create(fans, want=rt.WANT_RANGES | rt.WANT_INTENSITIES | rt.WANT_HIT_MODEL | rt.WANT_HIT_CELL)
    # warm-up of every path (outside the profiled range)
    ctx.blocks_remap(blocks, w.robot_ids, 0, z, first, worldgen.robot_boxes(w, poses[1]).reshape(-1, 2))
    ctx.sync()
    ctx.models_pose_update(w.robot_ids, worldgen.gposes(poses[2]), 0)
    ctx.sync()
    fset.trace()
    ctx.sync()
    torch.cuda.synchronize()
    flush.fill_(1)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fset.trace()                                                                   # k2_fan_headings, k2_ray_march
    ctx.sync()
    ctx.blocks_remap(blocks, w.robot_ids, 1, z, first, worldgen.robot_boxes(w, poses[3]).reshape(-1, 2))
    ctx.sync()                                                                     # k1_phase0, k1_phase1 (1000 robots, edge records)
    ctx.models_pose_update(w.robot_ids, worldgen.gposes(poses[4]), 0)
    ctx.sync()                                                                     # k1_pose_phase0, k1_pose_phase1 (1000 robots, poses)
    hit = ctx.collisions_test(0, (np.arange(1001, dtype=np.uint32), blocks))       # k4_collisions
    ctx.sync()
    ctx.touching_models(0, [[int(b)] for b in blocks[:1000]])                      # k4_touchers
    ctx.sync()
    torch.cuda.profiler.stop()
    print("config 3: 1000 robots, %d beams; %d collisions" % (len(fans) * 1080, int((hit >= 0).sum())))
    ctx.close()

    if args.skip_c4:
        return
    # ---- config 4
    n = args.c4_robots
    w = worldgen.make_swarm_world(seed=1, n_rects=4096, n_robots=n, half_extent_m=81.92)
    n_ob = len(w.obstacles)
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + 1, max_blocks=n_ob + n, max_cell_entries=1 << 23, max_delta_bytes=64 << 20)
    worldgen.load_swarm_into_context(ctx, w, fiducial_return=1)
    worldgen.define_boxes(ctx, n_ob, w.robot_ids, w.robot_size)
    sensors = worldgen.p2dx_sonar_sensors()
    p = [w.robots]
    for t in range(4):
        p.append(worldgen.step_robots(w, p[-1], t, speed_m=0.02))
    out = rt.Outputs(n * 16, rt.WANT_RANGES | rt.WANT_INTENSITIES, alloc=ctx.pinned_empty)
    fid_out, fid_cnt = np.zeros(n * 32, rt.FIDUCIAL_DTYPE), np.zeros(n, np.uint32)

    def tick(t):
        ctx.models_pose_update(w.robot_ids, worldgen.gposes(p[t]), t % 2)
        ctx.rangers_submit(sensors, worldgen.ranger_poses(p[t], w.robot_ids, (t + 1) % 2), out)
        tg, fd = worldgen.fiducial_records(w, p[t], w.robot_ids, layer=(t + 1) % 2)
        ctx.fiducials_submit_packed(tg, fd, 64, n * 32, into=(fid_out, fid_cnt))
        ctx.sync()
    tick(1)
    tick(2)
    torch.cuda.synchronize()
    flush.fill_(1)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    tick(3)      # k1_pose_phase0/1 (100 k robots), k2_expand_rangers, k2_fan_headings_short, k2_ray_march (1.6 M sonar beams), k3_fid_* (100 k finders)
    torch.cuda.profiler.stop()
    print("config 4: %d robots, %d sonar beams, %d fiducials seen" % (n, n * 16, int(fid_cnt.sum())))
    ctx.close()


if __name__ == "__main__":
    main()

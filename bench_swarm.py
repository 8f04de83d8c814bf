"""BASELINE.json configs 4 and 5 (SURVEY.md 8d "C4", "C5") as sub-records of bench.py's JSON line.

C4  100,000 robots in total (0.1 m square bodies rendered into the grid, fiducial_return 1) on the config-3 map, each
    with the 16 one-beam sonars of worlds/pioneer.inc's p2dx_sonar (range 5 m) and a fiducial finder (fov pi,
    range_max 8, range_max_id 5). The robots are sharded over the GPUs (strong scaling: the total is fixed); every
    tick each rank re-rasterises its own robots, the deltas and the fiducial target records are all-gathered, every
    rank applies all deltas (K1), traces its own sonars (K2) and runs its own finders against the world's targets (K3).
C5  the config-3 world (1000 laser robots per GPU) plus 10,000 pucks (0.1 m squares) in total that ALL move every
    tick: 10 k Block::UnMap + 10 k Block::Map per tick through delta export -> all-gather -> K1, then the lasers scan.

Each record carries `value` (device time of the kernels, CUDA events on the launching stream, L2 flushed between
steps), `e2e` (whole ticks through the C ABI from host buffers), the kernel split, the K1 roofline on SURVEY 8(d)'s
per-block bytes, and `parity`: every rank checks a sample of its own rays (and fiducials) of the last tick against the
T1 oracle fed the same maps in rank order - 0 mismatches is the only passing value - and the digest of its device grid
against the oracle's and against every other rank's.

The oracle (tests/oracle_lib.py -> oracle/libt1_oracle.so) is used here as the checker only, after the timed regions.
"""
import ctypes
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
SONAR_RANGE = 5.0


class Env:
    """what bench.py hands over: the process group (or None), this rank's stream, reductions over the ranks"""

    def __init__(self, torch, dist, rank, n_gpus, local_rank, stream, peak_gbs, clock_sampler=None):
        self.torch, self.dist, self.rank, self.n_gpus, self.local_rank, self.stream = torch, dist, rank, n_gpus, local_rank, stream
        self.peak_gbs = peak_gbs
        self.clock_sampler = clock_sampler
        self.flush_buf = None

    def clocks_start(self):
        """nvidia-smi clock / throttle sampling over a timed region (rank 0 only)"""
        return self.clock_sampler(self.local_rank) if (self.clock_sampler and self.rank == 0) else None

    def barrier(self):
        if self.n_gpus > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.n_gpus == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(float(x), self.dist.ReduceOp.MAX if self.dist else None)

    def min(self, x):
        return self._reduce(float(x), self.dist.ReduceOp.MIN if self.dist else None)

    def sum(self, x):
        return self._reduce(float(x), self.dist.ReduceOp.SUM if self.dist else None)

    def same_everywhere(self, values):
        """True iff every rank holds the same list of 64-bit integers"""
        if self.n_gpus == 1:
            return True
        t = self.torch.tensor([int(v) & 0x7fffffffffffffff for v in values] + [int(v) >> 63 for v in values], dtype=self.torch.int64, device="cuda")
        lo, hi = t.clone(), t.clone()
        self.dist.all_reduce(lo, op=self.dist.ReduceOp.MIN)
        self.dist.all_reduce(hi, op=self.dist.ReduceOp.MAX)
        return bool(self.torch.equal(lo, hi))

    def flush_l2(self):
        if self.flush_buf is None:
            self.flush_buf = self.torch.empty(256 << 20, dtype=self.torch.uint8, device="cuda")
        self.flush_buf.fill_(1)


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    return oracle_lib


def build_oracle_world(oracle_lib, w, boxes_by_layer, extra=None, fiducial_return=0):
    """The T1 oracle's world after the same history: every static body mapped on both layers at load, then the movers'
    latest outline per layer in id order (= rank order, shards are contiguous; every mover is re-rendered every tick, so
    only each layer's last tick is left in the grid). extra = (ids, size_z, boxes_by_layer) of a second group of movers
    rendered after the first (C5's pucks: their blocks follow the robots')."""
    from stage_b200 import worldgen
    t1 = oracle_lib.T1World(w.ppm)
    t1.model_upsert(0, 0, 1.0)
    for mid in w.obstacle_ids:
        t1.model_upsert(int(mid), int(mid), 0.5)
    for mid in w.robot_ids:
        t1.model_upsert(int(mid), int(mid), 0.5, fiducial_return)
    ob = np.array(worldgen.obstacle_polygons(w), np.int32).reshape(-1, 4, 2)
    n_ob, n_rb = len(ob), len(w.robots)

    def remap(first_block, ids, zmax, boxes, layer):
        n = len(boxes)
        t1.blocks_remap(first_block + np.arange(n, dtype=np.uint32), ids, layer, np.stack([np.zeros(n), np.full(n, zmax)], 1),
                        np.arange(n + 1, dtype=np.uint32) * 4, np.ascontiguousarray(boxes.reshape(-1, 2)))
    for layer in (0, 1):
        remap(0, w.obstacle_ids, w.obstacle_height, ob, layer)
    for layer in (0, 1):
        remap(n_ob, w.robot_ids, w.robot_size[2], boxes_by_layer[layer], layer)
    if extra is not None:
        ids, size_z, pboxes = extra
        for mid in ids:
            t1.model_upsert(int(mid), int(mid), 0.5)
        for layer in (0, 1):
            remap(n_ob + n_rb, ids, size_z, pboxes[layer], layer)
    return t1


def host_sonar_fans(sensors, poses, rt):
    """the fan records ModelRanger::Sensor::Update sets up (model_ranger.cc:217-230) for every sensor of every ranger
    model, with the host C library - what stg_rt_rangers_submit derives on the device"""
    fans = rt.make_fans(len(poses) * len(sensors))
    k = 0
    for P in poses:
        cosa, sina = math.cos(P["a"]), math.sin(P["a"])
        for sn in sensors:
            n = int(sn["n_samples"])
            start = -sn["fov"] / 2.0 if n > 1 else 0.0
            la, lz = sn["pa"] + start, sn["pz"] + sn["size_z"] / 2.0
            a = P["a"] + la
            while a < -math.pi:
                a += 2.0 * math.pi
            while a > math.pi:
                a -= 2.0 * math.pi
            f = fans[k]
            f["finder"], f["n_samples"], f["pred"], f["ztest"], f["layer"], f["angles"] = P["finder"], n, rt.PRED_RANGER, 1, P["layer"], rt.ANGLES_RANGER
            f["ox"] = P["x"] + sn["px"] * cosa - sn["py"] * sina
            f["oy"] = P["y"] + sn["px"] * sina + sn["py"] * cosa
            f["oz"], f["oa"], f["range"], f["fov"] = P["z"] + lz, a, sn["range_max"], sn["fov"]
            k += 1
    return fans


def outline_cell_visits(boxes):
    """cell visits World::MapPoly makes for closed 4-point outlines ((n, 4, 2) pixel arrays): |dx| + |dy| per edge. With poses in,
    the device derives the outlines and the host never counts them: the roofline's numerator is taken from the same boxes."""
    d = np.abs(np.roll(boxes.astype(np.int64), -1, axis=1) - boxes.astype(np.int64))
    return int(d.sum())


def k1_roofline(edge_steps, blocks_moved, n_pts, k1_ms, peak_gbs):
    """B_blk = 8 n_pts + 16 + 4 (cells unmapped + cells mapped) per moved block"""
    algo = blocks_moved * (8 * n_pts + 16) + 4 * edge_steps
    achieved = algo / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
    return {"bound": "hbm", "kernel": "k1_phase0 + k1_phase1 (delta rasteriser)", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
            "frac": achieved / peak_gbs, "traffic": None, "algorithmic_bytes_per_launch": algo, "blocks_per_launch": blocks_moved,
            "cell_visits_per_launch": edge_steps, "ms_per_launch": k1_ms,
            "note": "bytes = 8*n_pts + 16 + 4*(cells unmapped + cells mapped) per moved block (SURVEY.md 8d); the pass is bound by L2 atomics "
                    "latency, not by DRAM (DESIGN.md section 3 K1)"}


def laser_parity(env, ctx, w, movers, T, fans, timed_ranges, samples, first_puck_id=None):
    """Per-rank parity of a laser workload (C3, C5) after tick T. movers = [(first block, model ids, size z, boxes_at(t), n)] groups
    that every rank re-renders each tick, its own contiguous share of each, group after group. The T1 oracle gets the load
    (everything at tick 0 on both layers) and the last two ticks' re-renderings rank by rank - every mover is re-rendered
    every tick, so that is the grid's whole content - then traces a strided sample of this rank's fans (>= 10,000 rays)."""
    from stage_b200 import multi, rt, worldgen
    oracle_lib = _oracle()
    N = env.n_gpus
    t1 = oracle_lib.T1World(w.ppm)
    t1.model_upsert(0, 0, 1.0)
    for mid in w.obstacle_ids:
        t1.model_upsert(int(mid), int(mid), 0.5)
    ob = np.array(worldgen.obstacle_polygons(w), np.int32).reshape(-1, 4, 2)

    def remap(first_block, ids, zmax, boxes, layer):
        n = len(boxes)
        t1.blocks_remap((first_block + np.arange(n)).astype(np.uint32), ids, layer, np.stack([np.zeros(n), np.full(n, zmax)], 1),
                        np.arange(n + 1, dtype=np.uint32) * 4, np.ascontiguousarray(boxes.reshape(-1, 2)))
    for layer in (0, 1):
        remap(0, w.obstacle_ids, w.obstacle_height, ob, layer)
    for first_block, ids, zmax, boxes_at, n in movers:
        for mid in ids:
            t1.model_upsert(int(mid), int(mid), 0.5)
    for layer in (0, 1):   # the load: group after group (worldgen.load_into_context, then the pucks)
        for first_block, ids, zmax, boxes_at, n in movers:
            remap(first_block, ids, zmax, boxes_at(0), layer)
    for tt in (T - 1, T):
        boxes = [boxes_at(tt) for _, _, _, boxes_at, _ in movers]
        for r in range(N):
            for (first_block, ids, zmax, _, n), bx in zip(movers, boxes):
                a, b = multi.shard_bounds(n, r, N)
                remap(first_block + a, ids[a:b], zmax, bx[a:b], tt % 2)
    dg, do = ctx.grid_hash(), t1.grid_hash()
    grid_ok = dg == do
    replicas_equal = env.same_everywhere(dg[0] + dg[1] + [dg[2]])
    n_chk = max(1, min(len(fans), (10000 + samples - 1) // samples))
    idx = np.linspace(0, len(fans) - 1, n_chk).astype(int)
    chk = fans[idx]
    g = ctx.trace(chk, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL | rt.WANT_INTENSITIES)
    bad_range = bad_model = tied = 0
    for i, f in enumerate(chk):
        rg, it, _, hits = t1.ranger_fan(int(f["finder"]), int(f["layer"]), [f["ox"], f["oy"], f["oz"], f["oa"]], float(f["range"]), float(f["fov"]), samples)
        sl = slice(i * samples, (i + 1) * samples)
        bad_range += int(np.count_nonzero(g.ranges[sl].view(np.uint64) != rg.view(np.uint64)))
        bad_model += int(np.count_nonzero(g.hit_model[sl] != hits["hit_model"])) + int(np.count_nonzero(g.intensities[sl] != it))
        tied += int(np.count_nonzero(timed_ranges[idx[i] * samples:(idx[i] + 1) * samples].view(np.uint64) != g.ranges[sl].view(np.uint64)))
    mism = bad_range + bad_model + tied + (0 if grid_ok else 1)
    out = {
        "rays_checked_per_rank": int(env.min(len(chk) * samples)), "rays_checked": int(env.sum(len(chk) * samples)),
        "range_bits_mismatch": int(env.sum(bad_range)), "hit_model_or_intensity_mismatch": int(env.sum(bad_model)),
        "timed_tick_ranges_differ_from_checked_tick": int(env.sum(tied)),
        "grid_digest_equals_oracle_on_every_rank": bool(env.min(1 if grid_ok else 0) == 1), "grid_digest_equal_across_ranks": bool(replicas_equal),
        "grid_entries": dg[1], "mismatches_total": int(env.sum(mism)) + (0 if replicas_equal else 1),
        "how": "per rank: the T1 oracle (CPU restatement pinned to the unmodified reference, tests/test_oracle_vs_reference.py) is fed the load and the "
               "last two ticks' re-renderings rank by rank; every beam of a strided sample of the rank's lasers (>= 10,000 rays; device trig vs host libm) "
               "bit for bit + hit model + intensity, tied to the timed tick's delivered ranges; the 5-word digest of the device grid "
               "(stg_rt_grid_hash: entries, list order, region counts) against the oracle's and all ranks'",
    }
    if first_puck_id is not None:
        out["rays_that_hit_a_puck"] = int(env.sum(int(np.count_nonzero(g.hit_model >= first_puck_id))))
    return out


# ---------------------------------------------------------------------------------------------------
# C4
# ---------------------------------------------------------------------------------------------------

def run_c4(env, args):
    from stage_b200 import multi, rt, worldgen
    torch, dist, rank, N = env.torch, env.dist, env.rank, env.n_gpus
    n_total = args.c4_robots
    steps, warm = args.steps, max(args.warmup, 3)
    w = worldgen.make_swarm_world(seed=1, n_rects=args.rects, n_robots=n_total, half_extent_m=args.half_extent)
    x0, y0, nx, ny = w.sreg_extent()
    n_ob = len(w.obstacles)
    lo, hi = multi.shard_bounds(n_total, rank, N)
    n_mine = hi - lo
    slot_bytes = 64 + (multi.shard_bounds(n_total, 0, N)[1]) * (8 * 32 + 40) + 4096   # a robot: 4 edges out, 4 in, one block update
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + 1, max_blocks=n_ob + n_total, max_cell_entries=1 << 23,
                     device=env.local_rank, max_delta_bytes=slot_bytes)
    ctx.set_stream(env.stream.cuda_stream)
    worldgen.load_swarm_into_context(ctx, w, fiducial_return=1)
    poses_in = getattr(args, "deltas", "poses") == "poses"
    if poses_in:
        worldgen.define_boxes(ctx, n_ob, w.robot_ids, w.robot_size)
    sensors = worldgen.p2dx_sonar_sensors()
    n_beams = n_mine * len(sensors)
    total_beams = n_total * len(sensors)
    my_ids = w.robot_ids[lo:hi]
    my_blocks = (n_ob + np.arange(lo, hi)).astype(np.uint32)

    n_ticks = 2 * (warm + steps) + 1
    poses = [w.robots]
    for t in range(n_ticks):
        poses.append(worldgen.step_robots(w, poses[-1], t, speed_m=0.02))
    first4 = np.arange(n_mine + 1, dtype=np.uint32) * 4
    z_mine = np.ascontiguousarray(np.stack([np.zeros(n_mine), np.full(n_mine, w.robot_size[2])], 1))

    def prepare(t):
        """what libstage's Move() / sensor Update()s hand over at tick t (caller-side work, outside the timed regions)"""
        p = poses[t][lo:hi]
        tg, fd = worldgen.fiducial_records(w, p, my_ids, layer=(t + 1) % 2)
        return {"t": t, "layer_w": t % 2, "xy": np.ascontiguousarray(worldgen.robot_boxes(w, p).reshape(-1, 2)), "gpose": worldgen.gposes(p),
                "rposes": worldgen.ranger_poses(p, my_ids, layer=(t + 1) % 2), "targets": tg, "finders": fd}

    ticks = [prepare(t) for t in range(1, n_ticks + 1)]
    exchange = multi.DeltaExchange(ctx, rank, N, dist=dist) if N > 1 else None
    # the fiducial target records of every rank's robots, all-gathered on the device: only a rank's own share crosses PCIe
    tex = multi.TargetExchange(rank, N, max_per_rank=multi.shard_bounds(n_total, 0, N)[1], record_dtype=rt.FID_TARGET_DTYPE, dist=dist)
    out = rt.Outputs(n_beams, rt.WANT_RANGES | rt.WANT_INTENSITIES, alloc=ctx.pinned_empty)
    max_per_finder, fid_cap = args.c4_max_fiducials, max(n_mine * args.c4_max_fiducials // 2, 1 << 16)
    fid_out = np.zeros(fid_cap, rt.FIDUCIAL_DTYPE)
    fid_cnt = np.zeros(n_mine, np.uint32)
    split = {"remap_ms": 0.0, "delta_exchange_ms": 0.0, "rangers_submit_ms": 0.0, "target_exchange_ms": 0.0, "fiducials_submit_ms": 0.0, "sync_ms": 0.0}

    def tick(tk):
        t0 = time.perf_counter()
        if poses_in:
            ctx.models_pose_update(my_ids, tk["gpose"], tk["layer_w"])
        else:
            ctx.blocks_remap(my_blocks, my_ids, tk["layer_w"], z_mine, first4, tk["xy"])
        t1 = time.perf_counter()
        if exchange is not None:
            exchange.step()
        t2 = time.perf_counter()
        ctx.rangers_submit(sensors, tk["rposes"], out)
        t3 = time.perf_counter()
        dev_targets, n_targets = tex.step_device(tk["targets"], ctx)
        t4 = time.perf_counter()
        ctx.fiducials_submit_device_targets(dev_targets, n_targets, tk["finders"], max_per_finder, fid_cap, into=(fid_out, fid_cnt))
        t5 = time.perf_counter()
        ctx.sync()
        t6 = time.perf_counter()
        for k, v in zip(split, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
            split[k] += 1e3 * v

    # ---- value: the same ticks, device time of the kernels (events on the launching stream), L2 flushed between steps
    for tk in ticks[:warm]:
        tick(tk)
    env.barrier()
    sampler = env.clocks_start()
    st0 = ctx.stats()
    kms = {"deltas_ms": 0.0, "headings_ms": 0.0, "march_ms": 0.0, "fiducials_ms": 0.0}
    for tk in ticks[warm:warm + steps]:
        env.flush_l2()
        tick(tk)
        kt = ctx.kernel_times()
        for k in kms:
            kms[k] += kt[k]
    env.barrier()
    st1 = ctx.stats()
    ray_ms = env.max(kms["headings_ms"] + kms["march_ms"])
    all_ms = env.max(sum(kms.values()))
    value = total_beams * steps / (ray_ms * 1e-3)
    edge_steps = env.sum((st1["edge_steps"] - st0["edge_steps"]) / steps)   # one K1 pass applies every rank's deltas

    # ---- e2e: whole ticks from host buffers, host clock around K ticks
    for tk in ticks[warm + steps:2 * warm + steps]:
        tick(tk)
    for k in split:
        split[k] = 0.0
    env.barrier()
    st2 = ctx.stats()
    t0 = time.perf_counter()
    for tk in ticks[2 * warm + steps:2 * warm + 2 * steps]:
        tick(tk)
    torch.cuda.synchronize()
    t_e2e = env.max(time.perf_counter() - t0)
    env.barrier()
    st3 = ctx.stats()
    clocks = sampler.stop() if sampler else None
    last = ticks[2 * warm + 2 * steps - 1]
    T = last["t"]
    fid_seen = int(np.minimum(fid_cnt, max_per_finder).sum())
    fid_seen_total = env.sum(fid_seen)
    ranges_timed = out.ranges.copy()

    # ---- parity of the last tick, per rank, against the T1 oracle (after the timed regions)
    # the gathered target array as every rank's finders saw it (padding records included), rebuilt on the host for the checker
    all_t, _ = worldgen.fiducial_records(w, poses[T], w.robot_ids, layer=(T + 1) % 2)
    gathered_targets = tex.host_view([all_t[a:b] for a, b in (multi.shard_bounds(n_total, r, N) for r in range(N))])
    parity = c4_parity(env, ctx, w, poses, T, lo, hi, sensors, ranges_timed, fid_out, fid_cnt, max_per_finder, gathered_targets, last, rt, worldgen)
    k1_ms_avg = kms["deltas_ms"] / steps
    if poses_in:   # an unmap of the outline of two ticks ago (same layer) and a map of this tick's, for every robot of every rank
        edge_steps = outline_cell_visits(worldgen.robot_boxes(w, poses[T])) + outline_cell_visits(worldgen.robot_boxes(w, poses[T - 2]))
    rec = {
        "workload": "C4 synthetic: config-3 map (%d rectangles + walls, %dx%d cells), %d robots in total with 0.1 m bodies in the grid, each 16 one-beam "
                    "p2dx sonars (5 m) + a fiducial finder (fov pi, range_max 8, range_max_id 5); robots sharded over %d GPU(s), grid replicated"
                    % (args.rects, int(2 * args.half_extent * 50), int(2 * args.half_extent * 50), n_total, N),
        "metric": "rays/sec", "unit": "rays/s", "scaling": "strong", "robots_total": n_total, "robots_per_gpu": n_mine, "sonar_rays_per_step": total_beams,
        "value": value, "ms_per_step": ray_ms / steps, "ticks_per_s_kernels": steps / (all_ms * 1e-3),
        "kernel_ms": {k: v / steps for k, v in kms.items()},
        "e2e": {"value": total_beams * steps / t_e2e, "unit": "rays/s", "ms_per_step": 1e3 * t_e2e / steps, "ticks_per_s": steps / t_e2e,
                "h2d_bytes_per_step": (st3["h2d_bytes"] - st2["h2d_bytes"]) // steps, "d2h_bytes_per_step": (st3["d2h_bytes"] - st2["d2h_bytes"]) // steps,
                "delivers": ["ranges", "intensities", "fiducials (packed records + counts)"],
                "host_call_ms_per_step": {k: v / steps for k, v in split.items()},
                "fiducials_seen_per_step": fid_seen_total, "timing": "host clock around K ticks, synchronised both sides, max over ranks"},
        "roofline": k1_roofline(edge_steps, n_total, 4, k1_ms_avg, env.peak_gbs),
        "gpu_launches": int(sum(st3[k] - st0[k] for k in ("launches_rasterise", "launches_march", "launches_post", "launches_other"))),
        "parity": parity, "clocks": clocks,
    }
    ctx.close()
    return rec


def c4_parity(env, ctx, w, poses, T, lo, hi, sensors, ranges_timed, fid_out, fid_cnt, max_per_finder, targets, last, rt, worldgen):
    oracle_lib = _oracle()
    n_mine = hi - lo
    boxes = {T % 2: worldgen.robot_boxes(w, poses[T]), (T - 1) % 2: worldgen.robot_boxes(w, poses[T - 1])}
    t1 = build_oracle_world(oracle_lib, w, boxes, fiducial_return=1)
    # (1) the grid: this replica against the oracle, and every replica against every other
    dg, do = ctx.grid_hash(), t1.grid_hash()
    grid_ok = dg == do
    replicas_equal = env.same_everywhere(dg[0] + dg[1] + [dg[2]])
    # (2) sonar rays of a sample of this rank's robots: the same tick again with hit models, tied to the timed tick's ranges
    out2 = rt.Outputs(n_mine * len(sensors), rt.WANT_RANGES | rt.WANT_HIT_MODEL | rt.WANT_INTENSITIES)
    ctx.rangers_submit(sensors, last["rposes"], out2)
    ctx.sync()
    same_as_timed = bool(np.array_equal(out2.ranges.view(np.uint64), ranges_timed.view(np.uint64)))
    stride = max(1, n_mine // 700)
    sample = np.arange(0, n_mine, stride)
    fans = host_sonar_fans(sensors, last["rposes"][sample], rt)
    S = len(sensors)
    bad_range = bad_model = 0
    hit = 0
    for i, f in enumerate(fans):
        rg, it, _, hits = t1.ranger_fan(int(f["finder"]), int(f["layer"]), [f["ox"], f["oy"], f["oz"], f["oa"]], float(f["range"]), float(f["fov"]), 1)
        b = int(sample[i // S]) * S + i % S
        bad_range += int(out2.ranges[b:b + 1].view(np.uint64)[0] != rg.view(np.uint64)[0])
        bad_model += int(out2.hit_model[b] != hits["hit_model"][0]) + int(out2.intensities[b] != it[0])
        hit += int(hits["hit_model"][0] >= 0)
    # (3) fiducials of a sample of this rank's finders
    L = oracle_lib.t1()
    L.t1_fiducial_update.restype = ctypes.c_uint32
    L.t1_fiducial_update.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
    at = np.concatenate([[0], np.cumsum(np.minimum(fid_cnt, max_per_finder).astype(np.int64))])
    valid = targets["model"] != 0xffffffff                # the oracle gets the records without the slots' padding ...
    index_of = np.cumsum(valid) - 1                        # ... so the device's target indices are mapped onto that array
    targets = np.ascontiguousarray(targets[valid])
    fsample = np.arange(0, n_mine, max(1, n_mine // 250))
    fid_bad = fid_checked = 0
    fid_err = 0.0
    for i in fsample:
        ref = np.zeros(max_per_finder, rt.FIDUCIAL_DTYPE)
        f1 = np.ascontiguousarray(last["finders"][i:i + 1])
        n = int(L.t1_fiducial_update(t1.h, len(targets), targets.ctypes.data, f1.ctypes.data, max_per_finder, ref.ctypes.data))
        got = fid_out[int(at[i]):int(at[i + 1])]
        fid_checked += n
        if n != int(fid_cnt[i]) or not np.array_equal(index_of[got["target"]], ref[:n]["target"]) or not np.array_equal(got["id"], ref[:n]["id"]):
            fid_bad += 1
        elif n:
            fid_err = max(fid_err, float(np.max(np.abs(got["range"] - ref[:n]["range"]))), float(np.max(np.abs(got["bearing"] - ref[:n]["bearing"]))))
    mism = bad_range + bad_model + fid_bad + (0 if grid_ok else 1) + (0 if same_as_timed else 1)
    return {
        "rays_checked_per_rank": int(env.min(len(fans))), "rays_checked": int(env.sum(len(fans))), "rays_hit_fraction": hit / max(len(fans), 1),
        "range_bits_mismatch": int(env.sum(bad_range)), "hit_model_or_intensity_mismatch": int(env.sum(bad_model)),
        "timed_tick_ranges_equal_checked_tick": bool(env.min(1 if same_as_timed else 0) == 1),
        "fiducial_finders_checked": int(env.sum(len(fsample))), "fiducials_checked": int(env.sum(fid_checked)), "fiducial_mismatch": int(env.sum(fid_bad)),
        "fiducial_max_abs_err": env.max(fid_err), "fiducial_tolerance": 1e-9,
        "grid_digest_equals_oracle_on_every_rank": bool(env.min(1 if grid_ok else 0) == 1), "grid_digest_equal_across_ranks": bool(replicas_equal),
        "grid_entries": dg[1], "mismatches_total": int(env.sum(mism)) + (0 if replicas_equal else 1),
        "how": "per rank: the T1 oracle (CPU restatement pinned to the unmodified reference, tests/test_oracle_vs_reference.py) is fed every static body and "
               "the robots' last two ticks in rank order; sonar ranges bit for bit + hit model + intensity for every 16 beams of a strided sample of the "
               "rank's robots (device trig vs host libm), fiducial records (target, id exact; range, bearing to 1e-9) of a strided sample of its finders, "
               "and the 5-word digest of the device grid (stg_rt_grid_hash: entries, list order, region counts) against the oracle's and all ranks'",
    }


# ---------------------------------------------------------------------------------------------------
# C5
# ---------------------------------------------------------------------------------------------------

def step_pucks(pucks, tick, half_extent):
    """a seeded random step for every puck: up to 5 cm in x and y, up to 5 degrees, on the 1 mm / 0.001 degree lattice"""
    rng = np.random.RandomState(7000 + tick)
    out = pucks.copy()
    n = len(out)
    out[:, 0] = np.clip(np.round((out[:, 0] + rng.uniform(-0.05, 0.05, n)) * 1000) / 1000.0, -half_extent + 1.0, half_extent - 1.0)
    out[:, 1] = np.clip(np.round((out[:, 1] + rng.uniform(-0.05, 0.05, n)) * 1000) / 1000.0, -half_extent + 1.0, half_extent - 1.0)
    out[:, 2] = np.round(((out[:, 2] + rng.uniform(-5.0, 5.0, n) + 180.0) % 360.0 - 180.0) * 1000) / 1000.0
    return out


def run_c5(env, args, fov_deg, samples, range_max):
    from stage_b200 import multi, rt, worldgen
    torch, dist, rank, N = env.torch, env.dist, env.rank, env.n_gpus
    steps, warm = args.steps, max(args.warmup, 3)
    n_robots, n_pucks = args.robots * N, args.c5_pucks
    w = worldgen.make_world(seed=1, n_rects=args.rects, n_robots=n_robots, half_extent_m=args.half_extent)
    pw = worldgen.make_swarm_world(seed=1, n_rects=args.rects, n_robots=n_pucks, half_extent_m=args.half_extent)   # same seed: the same rectangles
    assert np.array_equal(pw.obstacles, w.obstacles)
    puck_size = (0.1, 0.1, 0.6)   # tall enough for the lasers (mounted at z = 0.5) to see them
    x0, y0, nx, ny = w.sreg_extent()
    n_ob = len(w.obstacles)
    puck_ids = (w.n_models + np.arange(n_pucks)).astype(np.uint32)
    lo, hi = multi.shard_bounds(n_robots, rank, N)
    plo, phi = multi.shard_bounds(n_pucks, rank, N)
    n_pm = phi - plo
    slot_bytes = 64 + (multi.shard_bounds(n_pucks, 0, N)[1] + args.robots) * (8 * 32 + 40) + 4096
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + n_pucks + 1, max_blocks=n_ob + n_robots + n_pucks, max_cell_entries=1 << 22,
                     device=env.local_rank, max_delta_bytes=slot_bytes if N > 1 else 0)
    ctx.set_stream(env.stream.cuda_stream)
    worldgen.load_into_context(ctx, w)
    for mid in puck_ids:
        ctx.model_upsert(int(mid), int(mid), 0.5)
    pucks0 = pw.robots
    puck_blocks = (n_ob + n_robots + np.arange(n_pucks)).astype(np.uint32)
    pz = np.ascontiguousarray(np.stack([np.zeros(n_pucks), np.full(n_pucks, puck_size[2])], 1))

    def puck_boxes(p):
        return worldgen.box_polygons(p[:, 0], p[:, 1], p[:, 2] * (math.pi / 180.0), puck_size[0], puck_size[1], w.ppm)
    b0 = np.ascontiguousarray(puck_boxes(pucks0).reshape(-1, 2))
    for layer in (0, 1):
        ctx.blocks_remap(puck_blocks, puck_ids, layer, pz, np.arange(n_pucks + 1, dtype=np.uint32) * 4, b0)
    ctx.sync()
    poses_in = getattr(args, "deltas", "poses") == "poses"
    if poses_in:
        worldgen.define_boxes(ctx, n_ob, w.robot_ids, w.robot_size)
        worldgen.define_boxes(ctx, n_ob + n_robots, puck_ids, puck_size)

    my_ids = w.robot_ids[lo:hi]
    my_blocks = (n_ob + np.arange(lo, hi)).astype(np.uint32)
    n_beams = (hi - lo) * samples
    total_rays = n_beams * N
    n_ticks = 2 * (warm + steps) + 1
    rposes, pposes = [w.robots], [pucks0]
    for t in range(n_ticks):
        rposes.append(worldgen.step_robots(w, rposes[-1], t))
        pposes.append(step_pucks(pposes[-1], t, args.half_extent))
    rz = np.ascontiguousarray(np.stack([np.zeros(hi - lo), np.full(hi - lo, w.robot_size[2])], 1))
    rfirst, pfirst = np.arange(hi - lo + 1, dtype=np.uint32) * 4, np.arange(n_pm + 1, dtype=np.uint32) * 4

    def prepare(t):
        return {"t": t, "layer_w": t % 2, "rgpose": worldgen.gposes(rposes[t][lo:hi]), "pgpose": worldgen.gposes(pposes[t][plo:phi]),
                "rxy": np.ascontiguousarray(worldgen.robot_boxes(w, rposes[t][lo:hi]).reshape(-1, 2)),
                "pxy": np.ascontiguousarray(puck_boxes(pposes[t][plo:phi]).reshape(-1, 2)),
                "fans": worldgen.ranger_fans(w, layer=(t + 1) % 2, fov_deg=fov_deg, samples=samples, range_max=range_max,
                                             robots=rposes[t][lo:hi], robot_ids=my_ids)}
    ticks = [prepare(t) for t in range(1, n_ticks + 1)]
    exchange = multi.DeltaExchange(ctx, rank, N, dist=dist) if N > 1 else None
    out = rt.Outputs(n_beams, rt.WANT_RANGES | rt.WANT_INTENSITIES, alloc=ctx.pinned_empty)
    split = {"remap_ms": 0.0, "delta_exchange_ms": 0.0, "fans_submit_ms": 0.0, "flush_ms": 0.0, "sync_ms": 0.0}

    def tick(tk):
        t0 = time.perf_counter()
        if poses_in:
            ctx.models_pose_update(my_ids, tk["rgpose"], tk["layer_w"])                      # this rank's robots move ...
            ctx.models_pose_update(puck_ids[plo:phi], tk["pgpose"], tk["layer_w"])           # ... and its share of the pucks
        else:
            ctx.blocks_remap(my_blocks, my_ids, tk["layer_w"], rz, rfirst, tk["rxy"])
            ctx.blocks_remap(puck_blocks[plo:phi], puck_ids[plo:phi], tk["layer_w"], pz[plo:phi], pfirst, tk["pxy"])
        t1 = time.perf_counter()
        if exchange is not None:
            exchange.step()
        t2 = time.perf_counter()
        ctx.fans_submit(tk["fans"], out)
        t3 = time.perf_counter()
        ctx.flush()
        t4 = time.perf_counter()
        ctx.sync()
        t5 = time.perf_counter()
        for k, v in zip(split, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            split[k] += 1e3 * v

    for tk in ticks[:warm]:
        tick(tk)
    env.barrier()
    sampler = env.clocks_start()
    st0 = ctx.stats()
    kms = {"deltas_ms": 0.0, "headings_ms": 0.0, "march_ms": 0.0}
    for tk in ticks[warm:warm + steps]:
        env.flush_l2()
        tick(tk)
        kt = ctx.kernel_times()
        for k in kms:
            kms[k] += kt[k]
    env.barrier()
    st1 = ctx.stats()
    ray_ms = env.max(kms["headings_ms"] + kms["march_ms"])
    k1_ms = env.max(kms["deltas_ms"])
    value = total_rays * steps / (ray_ms * 1e-3)
    edge_steps = (st1["edge_steps"] - st0["edge_steps"]) / steps
    edge_steps_all = env.sum(edge_steps) if N > 1 else edge_steps       # what one K1 pass applies: every rank's deltas
    blocks_all = n_robots + n_pucks

    for tk in ticks[warm + steps:2 * warm + steps]:
        tick(tk)
    for k in split:
        split[k] = 0.0
    env.barrier()
    st2 = ctx.stats()
    t0 = time.perf_counter()
    for tk in ticks[2 * warm + steps:2 * warm + 2 * steps]:
        tick(tk)
    torch.cuda.synchronize()
    t_e2e = env.max(time.perf_counter() - t0)
    env.barrier()
    st3 = ctx.stats()
    clocks = sampler.stop() if sampler else None
    last = ticks[2 * warm + 2 * steps - 1]
    T = last["t"]

    # ---- parity (after the timed regions): rank order interleaves the two groups - rank r's robots, then rank r's pucks
    movers = [(n_ob, w.robot_ids, w.robot_size[2], lambda t: worldgen.robot_boxes(w, rposes[t]), n_robots),
              (n_ob + n_robots, puck_ids, puck_size[2], lambda t: puck_boxes(pposes[t]), n_pucks)]
    parity = laser_parity(env, ctx, w, movers, T, last["fans"], out.ranges, samples, first_puck_id=int(puck_ids[0]))
    k1_avg = k1_ms / steps
    if poses_in:   # every robot and every puck of every rank: the outline of two ticks ago (same layer) out, this tick's in
        edge_steps_all = sum(outline_cell_visits(worldgen.robot_boxes(w, rposes[t])) + outline_cell_visits(puck_boxes(pposes[t])) for t in (T, T - 2))
    rec = {
        "workload": "C5 moving-block stress: the config-3 world (%d laser robots per GPU, %d beams each) + %d pucks (0.1 m squares) in total that all "
                    "take a seeded random step every tick, sharded over %d GPU(s): delta export -> all-gather -> K1 on every replica, then the lasers scan"
                    % (args.robots, samples, n_pucks, N),
        "metric": "rays/sec", "unit": "rays/s", "scaling": "weak in robots (%d per GPU), strong in pucks (%d in total)" % (args.robots, n_pucks),
        "pucks_total": n_pucks, "blocks_rerendered_per_step": blocks_all, "rays_per_step": total_rays,
        "value": value, "ms_per_step": ray_ms / steps, "kernel_ms": {k: v / steps for k, v in kms.items()},
        "k1": {"ms_per_step": k1_avg, "blocks_per_s": blocks_all / (k1_avg * 1e-3) if k1_avg > 0 else None, "cell_visits_per_step": edge_steps_all},
        "e2e": {"value": total_rays * steps / t_e2e, "unit": "rays/s", "ms_per_step": 1e3 * t_e2e / steps, "ticks_per_s": steps / t_e2e,
                "blocks_per_s": blocks_all * steps / t_e2e,
                "h2d_bytes_per_step": (st3["h2d_bytes"] - st2["h2d_bytes"]) // steps, "d2h_bytes_per_step": (st3["d2h_bytes"] - st2["d2h_bytes"]) // steps,
                "delivers": ["ranges", "intensities"], "host_call_ms_per_step": {k: v / steps for k, v in split.items()},
                "delta_bytes_exported_per_step": (exchange.bytes_exported // (2 * (warm + steps))) if exchange else None,
                "timing": "host clock around K ticks, synchronised both sides, max over ranks"},
        "roofline": k1_roofline(edge_steps_all, blocks_all, 4, k1_avg, env.peak_gbs),
        "gpu_launches": int(sum(st3[k] - st0[k] for k in ("launches_rasterise", "launches_march", "launches_post", "launches_other"))),
        "parity": parity, "clocks": clocks,
    }
    ctx.close()
    return rec

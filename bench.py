#!/usr/bin/env python
"""Benchmark of the sensor-raycast path on BASELINE.json's config 3 ("C3"):
synthetic 8192 x 8192-cell world (ppm 50, 4096 random rectangles + walls), 1000 robots per GPU each
with a 1080-beam, 270 degree, 30 m laser.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # the reference's CPU path
    torchrun --nproc-per-node N bench.py --gpus N ...              # N > 1: one rank per GPU

A "step" is one simulation tick's worth of sensor rays for the robots of one GPU.
  value  rays/s of the march (fan headings K2a + ray march K2 with the K3 epilogue fused), fans
         and grid resident in HBM, L2 flushed before every timed step.
  e2e    rays/s of a whole tick through the C ABI from host buffers: re-rasterise every robot body
         (Block::UnMap + Block::Map deltas, K1), submit the fans, trace, copy ranges + hit model
         back to page-locked host memory.  N > 1: the moved-block deltas of all ranks are
         all-gathered with NCCL and applied out of the gathered buffer.
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FOV_DEG, SAMPLES, RANGE_MAX = 270.0, 1080, 30.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--robots", type=int, default=1000, help="robots per GPU")
    ap.add_argument("--rects", type=int, default=4096)
    ap.add_argument("--half-extent", type=float, default=81.92)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--only", default="all", choices=["all", "c3", "c4", "c5"],
                    help="all (default): the config-3 line with configs 4 and 5 as sub-records under `configs`; c3 / c4 / c5: that one alone")
    ap.add_argument("--c4-robots", type=int, default=100000, help="config 4: robots in total (sharded over the GPUs)")
    ap.add_argument("--c4-max-fiducials", type=int, default=64, help="config 4: records kept per finder (the counts are exact beyond it)")
    ap.add_argument("--c5-pucks", type=int, default=10000, help="config 5: pucks in total, all moved every tick")
    ap.add_argument("--deltas", default="poses", choices=["poses", "outlines"],
                    help="how a tick's moved bodies reach the grid: poses (default) = stg_rt_models_pose_update, the outlines are derived on the "
                         "device (Model::LocalToPixels); outlines = stg_rt_blocks_remap with outlines rasterised by the caller")
    ap.add_argument("--workload", default="c3", choices=["c3", "cave"],
                    help="c3 = BASELINE config 3 (the default, the judged line); cave = the denser variant of SURVEY.md 8(d): "
                         "worlds/bitmaps/cave.png tiled 10x10 (no reference arm: value / e2e / roofline only)")
    return ap.parse_args()


def make_workload(args, n_gpus):
    from stage_b200 import worldgen
    if args.workload == "cave":
        return worldgen.make_cave_world(seed=1, tiles=10, n_robots=args.robots * n_gpus)
    return worldgen.make_world(seed=1, n_rects=args.rects, n_robots=args.robots * n_gpus, half_extent_m=args.half_extent)


def workload_config(args, n_gpus):
    if args.workload == "cave":
        return {
            "workload": "C3-cave: worlds/bitmaps/cave.png (16 m floorplan of simple.world) tiled 10x10 = 8000x8000 cells (ppm 50), "
                        "%d robots/GPU x %d-beam %g-degree %g m laser, z-test, ranger predicate" % (args.robots, SAMPLES, FOV_DEG, RANGE_MAX),
            "robots_total": args.robots * n_gpus, "beams_per_robot": SAMPLES, "rays_per_step": args.robots * n_gpus * SAMPLES,
            "parallelism": "robots sharded over %d GPU(s), grid replicated" % n_gpus,
            "l2": "flushed (256 MiB write) before every timed step of `value`",
        }
    return {
        "workload": "C3 synthetic: %dx%d-cell world (ppm 50), %d rectangles + walls, %d robots/GPU x %d-beam "
                    "%g-degree %g m laser, z-test, ranger predicate" % (
                        int(2 * args.half_extent * 50), int(2 * args.half_extent * 50), args.rects, args.robots,
                        SAMPLES, FOV_DEG, RANGE_MAX),
        "robots_total": args.robots * n_gpus, "beams_per_robot": SAMPLES, "rays_per_step": args.robots * n_gpus * SAMPLES,
        "parallelism": "robots sharded over %d GPU(s), grid replicated" % n_gpus,
        "l2": "flushed (256 MiB write) before every timed step of `value`",
    }


def leave(code):
    """End the process without running C++ static destructors (a reference World must never be destroyed: ~World
    crashes in the reference itself) but WITH everything Python registered for exit - the driver's hook
    that records which shared libraries this process loaded lives there."""
    import atexit
    sys.stdout.flush()
    sys.stderr.flush()
    try:
        atexit._run_exitfuncs()
    finally:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(code)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            if len(p) >= 6:
                self.samples.append(p)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref = the unmodified reference built headless)
# ---------------------------------------------------------------------------------------------

def host_rays(fans):
    """Expand ranger fans into one record per beam exactly as ModelRanger::Sensor::Update walks them
    (float-rounded running heading, model_ranger.cc:237-241,254), for World::Raytrace on the host."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    m = int(fans["n_samples"][0])
    assert np.all(fans["n_samples"] == m)
    nf = len(fans)
    incr = fans["fov"] / max(m - 1, 1)
    a = fans["oa"].astype(np.float64)
    h = np.empty((nf, m))
    for t in range(m):  # sequential in t (the recurrence), vectorised over fans
        fa = a.astype(np.float32).astype(np.float64)
        h[:, t] = fa
        a = fa + incr
    rays = np.zeros((nf, m), oracle_lib.REF_RAY)
    for k in ("ox", "oy", "oz", "range", "finder", "ztest"):
        rays[k] = fans[k][:, None]
    rays["kind"] = fans["pred"][:, None]
    rays["oa"] = h
    return rays.reshape(-1)


def reference_world(world, record=False):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from stage_b200 import worldgen
    if not oracle_lib.ref_available():
        return None, None
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)   # the reference prints its banner to stdout; the JSON line must stay alone
    try:
        ref = oracle_lib.RefWorld(text=worldgen.to_worldfile_text(world), record=record)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    assert ref.model_id("o0") == int(world.obstacle_ids[0]) and ref.model_id("r0") == int(world.robot_ids[0])
    return ref, oracle_lib


def cpu_reference_throughput(world, fans, threads, reps, warm=1, record=False):
    """rays/s of the reference's own World::Raytrace over the given fans, `threads` host threads."""
    ref, _ = reference_world(world, record=record)
    if ref is None:
        return None
    rays = host_rays(fans)
    best, per_rep, hits = None, [], None
    for i in range(warm + reps):
        hits, secs = ref.raytrace(rays, nthreads=threads)
        if i >= warm:
            per_rep.append(secs)
    best = min(per_rep)
    return {"rays": len(rays), "best_s": best, "mean_s": float(np.mean(per_rep)), "rays_per_s": len(rays) / best,
            "hits": hits, "ref": ref, "rays_arr": rays}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from stage_b200 import worldgen
    n_gpus = args.gpus
    if args.workload != "c3":
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm runs BASELINE config 3 only (--workload c3)"}))
        return 0
    world = make_workload(args, n_gpus)
    threads = args.cpu_threads or os.cpu_count()
    ref, _ = reference_world(world)
    if ref is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built on this box"}))
        return 0
    # The same ticks as this repo's e2e arm: every robot takes its next pose (what ModelPosition::Move does to the grid:
    # UnMapWithChildren + MapWithChildren on the tick's layer, one robot after the other on one thread, as
    # World::Update runs its movers, world.cc:647-648), then every laser scans from the new pose (World::Raytrace,
    # split over all host threads). Poses and rays are prepared outside the timed region, like the e2e arm's.
    n_ticks = args.warmup + args.steps
    poses = [world.robots]
    for t in range(n_ticks):
        poses.append(worldgen.step_robots(world, poses[-1], t))
    ticks = []
    for t in range(1, n_ticks + 1):
        p = poses[t]
        fans = worldgen.ranger_fans(world, layer=(t + 1) % 2, fov_deg=FOV_DEG, samples=SAMPLES, range_max=RANGE_MAX, robots=p)
        xyza = np.stack([p[:, 0], p[:, 1], np.zeros(len(p)), np.array([worldgen.heading_rad(d) for d in p[:, 2]])], 1)
        ticks.append((t % 2, xyza, host_rays(fans)))
    t_move = t_rays = 0.0
    n_rays = len(ticks[0][2])
    for k, (layer, xyza, rays) in enumerate(ticks):
        s_move = ref.models_remap(world.robot_ids, xyza, layer)
        _, s_rays = ref.raytrace(rays, nthreads=threads)
        if k >= args.warmup:
            t_move += s_move
            t_rays += s_rays
    t_total = t_move + t_rays
    v = n_rays * args.steps / t_total
    line = {
        "impl": "reference", "metric": "rays/sec", "value": v, "unit": "rays/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "ticks_per_s": args.steps / t_total,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n_gpus),
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": threads, "kind": "reference",
                         "sample": "every step = one tick of the workload through the unmodified reference (oracle/_ref): %d robots "
                                   "re-rendered at their next pose (UnMapWithChildren + MapWithChildren, serial as in World::Update), "
                                   "then all %d rays through World::Raytrace split over %d host threads" % (len(world.robots), n_rays, threads)},
        "tick_split_ms": {"moves_1_thread": 1e3 * t_move / args.steps, "rays_%d_threads" % threads: 1e3 * t_rays / args.steps},
        "rays_only_value": n_rays * args.steps / t_rays,
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    leave(0)  # a reference World must never be destroyed


# ---------------------------------------------------------------------------------------------
# this repo's path
# ---------------------------------------------------------------------------------------------

def peak_hbm_gbs():
    """(GB/s, where it came from): the driver-measured copy bandwidth of this pool's B200s, else the profiling guide's fallback"""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6650 (B200_PROFILING.md)"


def kernel_source_digest():
    """what the committed ncu capture of the march kernel is tied to: the sources the kernel is compiled from"""
    import hashlib
    h = hashlib.sha256()
    for f in ("rt_trace.cuh", "rt_walk.h", "rt_march.cu", "rt_device.h", "rt_libm.cuh"):
        h.update(open(os.path.join(ROOT, "stage_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def march_traffic():
    """DRAM bytes per launch of the march kernel (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of
    this workload, profiles/r02_traffic.json) - reported only while the kernel's sources are the ones that capture was taken
    from; a bench run cannot measure it itself (no profiler inside a timed run)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return None, "no capture committed"
    if t.get("kernel_source_digest") != kernel_source_digest():
        return None, "the committed capture (profiles/r02_traffic.json) is of an older build of the kernel"
    return t["dram_bytes_per_launch"], "ncu --set full capture of this kernel build, %s" % t.get("capture", "profiles/")


def pin_to_gpu_numa(local_rank):
    """Restrict this process to the CPUs of the NUMA node the GPU is attached to; returns the node or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pci = pynvml.nvmlDeviceGetPciInfo(hnd).busId
        pci = pci.decode() if isinstance(pci, bytes) else pci
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % pci.lower()[-12:]).read())
        if node < 0:
            return None
        cpus = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        ids = []
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(ids) & os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from stage_b200 import multi, rt, worldgen

    n_gpus = args.gpus
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    # the library's host threads (stg_rt_sync expands intensity palette indices with them): the box's cores are shared by
    # the ranks, one of each rank's share stays with the thread that calls the ABI
    os.environ.setdefault("STG_RT_HOST_THREADS", str(max(1, min(8, (os.cpu_count() or 8) // max(1, n_gpus) - 1))))
    if world_size != n_gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torchrun --nproc-per-node %d" % (n_gpus, world_size, n_gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # keep the process (and the page-locked result buffers it first-touches) on the NUMA node its GPU hangs off:
    # the results travel GPU -> host every tick and a cross-socket hop costs bandwidth (matters on multi-socket hosts;
    # a no-op on a single-node box)
    affinity0 = os.sched_getaffinity(0)
    numa_node = pin_to_gpu_numa(local_rank)
    if n_gpus > 1:
        # NCCL prints its version banner on stdout when the communicator is created; stdout must carry
        # nothing but the JSON line, so it is pointed at stderr until the first collective is through
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    if args.only in ("c4", "c5"):   # development aid: one of the sub-records alone, without the config-3 line around it
        import bench_swarm
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        env = bench_swarm.Env(torch, dist if n_gpus > 1 else None, rank, n_gpus, local_rank, stream, peak_hbm_gbs()[0], ClockSampler)
        rec = bench_swarm.run_c4(env, args) if args.only == "c4" else bench_swarm.run_c5(env, args, FOV_DEG, SAMPLES, RANGE_MAX)
        if rank == 0:
            print(json.dumps(dict(rec, n_gpus=n_gpus, steps=args.steps, warmup=max(args.warmup, 3), higher_is_better=True, vs_baseline=None,
                                  dtype="f64", data="synthetic", config={"workload": rec["workload"]})))
        if n_gpus > 1:
            dist.barrier()
            dist.destroy_process_group()
        leave(0)

    world = make_workload(args, n_gpus)
    x0, y0, nx, ny = world.sreg_extent()
    n_bodies = world.n_static_blocks + len(world.robots)
    ctx = rt.Context(world.ppm, x0, y0, nx, ny, max_models=world.n_models + 1, max_blocks=n_bodies,
                     max_cell_entries=(1 << 23) if args.workload == "cave" else (1 << 22), device=local_rank, max_delta_bytes=(1 << 20) if n_gpus > 1 else 0)
    # all of the context's work goes on a torch-owned stream so torch.cuda.Event can bracket it
    # (the default stream's handle is NULL, which stg_rt_set_stream reads as "use your own")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    worldgen.load_into_context(ctx, world)
    if args.deltas == "poses":   # every replica knows every robot's body: the moved-block deltas are poses (48 B per block)
        worldgen.define_boxes(ctx, world.n_static_blocks, world.robot_ids, world.robot_size)

    lo, hi = multi.shard_bounds(args.robots * n_gpus, rank, n_gpus)
    my_ids = world.robot_ids[lo:hi]
    my_blocks = (world.n_static_blocks + np.arange(lo, hi)).astype(np.uint32)
    n_beams = args.robots * SAMPLES
    total_rays = n_beams * n_gpus

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if n_gpus == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: resident fans, kernels only -------------------------------------------------
    want_value = rt.WANT_RANGES | rt.WANT_INTENSITIES | rt.WANT_HIT_MODEL | rt.WANT_HIT_CELL
    fans0 = worldgen.ranger_fans(world, layer=1, fov_deg=FOV_DEG, samples=SAMPLES, range_max=RANGE_MAX,
                                 robots=world.robots[lo:hi], robot_ids=my_ids)
    fset = ctx.set_create(fans0, want=want_value)
    # algorithmic bytes: the kernel's own C / R counters (tests prove them equal to the oracle's)
    cset = ctx.set_create(fans0, want=rt.WANT_STEPS | rt.WANT_HIT_MODEL)
    cset.trace()
    cnt = cset.download()
    C, R = int(cnt.steps[:, 0].sum(dtype=np.int64)), int(cnt.steps[:, 1].sum(dtype=np.int64))
    hit_frac = float(np.mean(cnt.hit_model >= 0))
    cset.destroy()
    algo_bytes = 4 * C + 4 * R + 40 * n_beams

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    launches0 = ctx.stats()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        fset.trace()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    march_ms, head_ms = [], []
    barrier()
    for s, e in ev:
        flush_buf.fill_(1)       # evict the grid and the fans from L2 (same stream, before the start event)
        s.record(stream)
        fset.trace()
        e.record(stream)
        e.synchronize()
        kt = ctx.kernel_times()
        march_ms.append(kt["march_ms"])
        head_ms.append(kt["headings_ms"])
    barrier()
    step_ms = [s.elapsed_time(e) for s, e in ev]
    t_value_ms = max_over_ranks(float(np.sum(step_ms)))
    value = total_rays * args.steps / (t_value_ms * 1e-3)
    launches1 = ctx.stats()

    # ---- e2e: one whole tick per step through the C ABI, host buffers ---------------------------
    # what ModelRanger::Sensor::Update fills (model_ranger.cc:243-251): ranges and intensities every tick; bearings are a
    # function of the sensor's geometry alone (start_angle + t * sample_incr), so a caller fills them once per sensor
    # type - here from one submit before the timed ticks - and keeps them
    want_e2e = rt.WANT_RANGES | rt.WANT_INTENSITIES
    out = rt.Outputs(n_beams, want_e2e, alloc=ctx.pinned_empty)
    once = rt.Outputs(n_beams, rt.WANT_BEARINGS)
    ctx.fans_submit(fans0, once)
    ctx.sync()
    bearings = once.bearings
    # e2e ticks take under a millisecond each: W of them are over before the host's clocks, caches and page tables
    # have settled on a fresh box (first run of a box seen 0.47 ms per tick, the second 0.44), so the untimed lead-in
    # is at least 20 ticks (said in `config.e2e_warmup_ticks`); the timed region is exactly K ticks either way
    nw_e2e = max(args.warmup, 3, 20)
    n_e2e = args.steps + nw_e2e
    poses_all = [world.robots]
    for t in range(n_e2e):
        poses_all.append(worldgen.step_robots(world, poses_all[-1], t))
    poses = [p[lo:hi] for p in poses_all]
    ticks = []   # what libstage's Move() / Sensor::Update would hand over each tick (caller-side work)
    for t in range(1, n_e2e + 1):
        polys = worldgen.robot_polygons(world, poses[t])
        layer_w = t % 2
        ticks.append({
            "xy": np.ascontiguousarray(np.concatenate(polys, 0), np.int32),
            "first": np.arange(len(polys) + 1, dtype=np.uint32) * 4,
            "z": np.ascontiguousarray(np.stack([np.zeros(len(polys)), np.full(len(polys), world.robot_size[2])], 1)),
            "layer_w": layer_w, "gpose": worldgen.gposes(poses[t]),
            "fans": worldgen.ranger_fans(world, layer=(t + 1) % 2, fov_deg=FOV_DEG, samples=SAMPLES, range_max=RANGE_MAX,
                                         robots=poses[t], robot_ids=my_ids),
        })
    exchange = multi.DeltaExchange(ctx, rank, n_gpus, dist=dist) if n_gpus > 1 else None

    host_split = {"remap_ms": 0.0, "exchange_ms": 0.0, "submit_ms": 0.0, "flush_ms": 0.0, "sync_ms": 0.0}

    def tick(tk):
        t0 = time.perf_counter()
        if args.deltas == "poses":
            ctx.models_pose_update(my_ids, tk["gpose"], tk["layer_w"])
        else:
            ctx.blocks_remap(my_blocks, my_ids, tk["layer_w"], tk["z"], tk["first"], tk["xy"])
        t1 = time.perf_counter()
        if exchange is not None:  # all-gather the moved-block deltas, apply all ranks' in rank order
            exchange.step()
        t2 = time.perf_counter()
        ctx.fans_submit(tk["fans"], out)
        t3 = time.perf_counter()
        ctx.flush()          # host side of the tick: build + upload the delta blob, enqueue every kernel
        t3b = time.perf_counter()
        ctx.sync()           # wait for the GPU
        t4 = time.perf_counter()
        host_split["remap_ms"] += 1e3 * (t1 - t0)
        host_split["exchange_ms"] += 1e3 * (t2 - t1)
        host_split["submit_ms"] += 1e3 * (t3 - t2)
        host_split["flush_ms"] += 1e3 * (t3b - t3)
        host_split["sync_ms"] += 1e3 * (t4 - t3b)

    nw = nw_e2e
    for tk in ticks[:nw]:
        tick(tk)
    st0 = ctx.stats()
    for k in host_split:
        host_split[k] = 0.0
    barrier()
    t0 = time.perf_counter()
    for tk in ticks[nw:]:
        tick(tk)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    st1 = ctx.stats()
    clocks = sampler.stop() if sampler else None
    e2e_value = total_rays * args.steps / t_e2e
    e2e_kt = ctx.kernel_times()

    # ---- CPU baseline + parity of the last e2e tick against the unmodified reference ----------
    cpu = None
    parity = None
    os.sched_setaffinity(0, affinity0)  # the reference gets every host core
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline and args.workload == "c3":
        threads = args.cpu_threads or os.cpu_count()
        sample_fans = fans0
        cpu_reps = 20  # about a second of wall clock on 16 threads: ~15 s of CPU work
        res = cpu_reference_throughput(world, sample_fans, threads, reps=cpu_reps, record=True)
        if res is not None:
            one = res["ref"].raytrace(res["rays_arr"][:SAMPLES * 50], nthreads=1)[1]
            cpu = {"value": res["rays_per_s"], "unit": "rays/s", "cores": threads, "kind": "reference",
                   "sample": "all %d rays of one tick (static robots), best of %d passes (mean %.3g rays/s), unmodified reference "
                             "World::Raytrace via oracle/_ref on %d threads; 1 thread: %.3g rays/s. Rays only: "
                             "`--impl reference` times whole ticks (the robots' re-rendering too, as `e2e` does)"
                             % (res["rays"], cpu_reps, res["rays"] / res["mean_s"], threads, SAMPLES * 50 / one)}
            # Parity of this very workload with the reference's answers. Which of two overlapping
            # bodies a ray reports depends on the order the reference mapped them in while loading
            # (its std::set<Model*> order, world.cc:459-465), so a second context is loaded from the
            # reference's own recorded Map/UnMap calls and traces the same fans.
            import tempfile
            import replay
            import trace_io
            with tempfile.TemporaryDirectory() as td:
                res["ref"].trace_end(os.path.join(td, "load.trc"))
                tr = trace_io.load(os.path.join(td, "load.trc"))
            ctx2 = replay.make_context(tr, max_cell_entries=1 << 22)
            for ev in tr.events:
                if ev["type"] == trace_io.EV_MAP:
                    ctx2.block_map(int(ev["block"]), int(ev["model"]), int(ev["layer"]), float(ev["zmin"]), float(ev["zmax"]), tr.polygon(ev))
                elif ev["type"] == trace_io.EV_UNMAP:
                    ctx2.block_unmap(int(ev["block"]), int(ev["layer"]))
            g = ctx2.trace(fans0, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL)
            ctx2.close()
            h = res["hits"]
            parity = {"rays": int(len(h)), "hit_model_mismatch": int(np.count_nonzero(g.hit_model != h["hit_model"])),
                      "range_bits_mismatch": int(np.count_nonzero(g.ranges.view(np.uint64) != h["range"].view(np.uint64))),
                      "max_range_err_m": float(np.max(np.abs(g.ranges - h["range"]))),
                      "how": "device trig; grid loaded from the reference's recorded Map/UnMap calls; vs the unmodified reference's World::Raytrace on this host"}

    peak, peak_source = peak_hbm_gbs()
    import bench_swarm
    env = bench_swarm.Env(torch, dist if n_gpus > 1 else None, rank, n_gpus, local_rank, stream, peak, ClockSampler)
    # every rank: a sample of its own rays of the last e2e tick against the T1 oracle, and its grid against the oracle's
    # and every other rank's (the only parity there is at N > 1; at N = 1 it stands next to the whole-tick check against
    # the unmodified reference above)
    parity_ranks = None
    if args.workload == "c3":
        T = n_e2e
        n_all = args.robots * n_gpus
        movers = [(world.n_static_blocks, world.robot_ids, world.robot_size[2], lambda t: worldgen.robot_boxes(world, poses_all[t]), n_all)]
        parity_ranks = bench_swarm.laser_parity(env, ctx, world, movers, T, ticks[-1]["fans"], out.ranges, SAMPLES)
    ctx.close()
    configs = {}
    if args.only in ("all", "c4") and args.workload == "c3":
        configs["c4"] = bench_swarm.run_c4(env, args)
    if args.only in ("all", "c5") and args.workload == "c3":
        configs["c5"] = bench_swarm.run_c5(env, args, FOV_DEG, SAMPLES, RANGE_MAX)

    # what bounds an e2e tick: the march stores its results into host memory as it goes, so its duration inside the e2e tick
    # against the bytes it delivers is the link rate this run saw; profiles/r02_link_ceiling.json holds what the host takes
    # from N GPUs at once (tools/link_ceiling.py on the 8-GPU box)
    e2e_march_ms = env.max(e2e_kt["march_ms"])
    if rank == 0:
        traffic, traffic_note = march_traffic()
        bytes_per_tick = (st1["d2h_bytes"] - st0["d2h_bytes"]) / args.steps
        limiter = {"delivered_gb_per_s_per_gpu": bytes_per_tick / (e2e_march_ms * 1e-3) / 1e9, "march_ms_inside_the_tick": e2e_march_ms,
                   "march_ms_results_in_hbm": float(np.mean(march_ms))}
        try:
            ceil = json.load(open(os.path.join(ROOT, "profiles", "r02_link_ceiling.json")))
            c = ceil["kernel_stores"].get(str(n_gpus))
            if c:
                limiter["host_link_ceiling_gb_per_s_per_gpu"] = c
                limiter["ceiling_source"] = "profiles/r02_link_ceiling.json (tools/link_ceiling.py, %d GPUs storing at once)" % n_gpus
                floor_ms = bytes_per_tick / (c * 1e9) * 1e3
                limiter["link_floor_ms"] = floor_ms
                hbm_ms = limiter["march_ms_results_in_hbm"]
                if e2e_march_ms > 1.4 * hbm_ms:
                    limiter["what"] = ("result delivery into host memory: the march takes %.2f ms storing its %.1f MB of results to the host against %.2f ms "
                                       "into HBM; what this host takes from %d GPU(s) storing at once (%.1f GB/s each, measured) allows no less than %.2f ms"
                                       % (e2e_march_ms, bytes_per_tick / 1e6, hbm_ms, n_gpus, c, floor_ms))
                else:
                    limiter["what"] = "kernels and host calls (delivery costs the march less than 40 % extra)"
        except Exception:
            pass
        march_avg_ms = float(np.mean(march_ms))
        achieved = algo_bytes / (march_avg_ms * 1e-3) / 1e9
        line = {
            "metric": "rays/sec", "value": value, "unit": "rays/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t_value_ms / args.steps,
            "ticks_per_s": args.steps / (t_value_ms * 1e-3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n_gpus),
            "host": {"numa_node": numa_node, "e2e_warmup_ticks": nw_e2e, "cpus": os.cpu_count()},
            "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_step": 1e3 * t_e2e / args.steps,
                    "ticks_per_s": args.steps / t_e2e,
                    "h2d_bytes_per_step": (st1["h2d_bytes"] - st0["h2d_bytes"]) // args.steps,
                    "d2h_bytes_per_step": (st1["d2h_bytes"] - st0["d2h_bytes"]) // args.steps,
                    "delivers": ["ranges (f64, written by the march kernel straight into the caller's page-locked array)",
                                 "intensities (f64 in the caller's array: one palette byte per beam over PCIe, expanded by stg_rt_sync)",
                                 "bearings (f64, filled once per sensor type before the timed ticks: they depend on the sensor's geometry only)"],
                    "bearings_checked": bool(np.array_equal(bearings[:SAMPLES], -np.deg2rad(FOV_DEG) / 2.0 + np.arange(SAMPLES) * (np.deg2rad(FOV_DEG) / (SAMPLES - 1)))),
                    "limiter": limiter, "last_step_kernel_ms": e2e_kt, "host_call_ms_per_step": {k: v / args.steps for k, v in host_split.items()}, "timing": "host clock around K ticks, synchronised both sides, max over ranks"},
            "gpu_launches": int(sum(st1[k] - launches0[k] for k in ("launches_rasterise", "launches_march", "launches_post", "launches_other"))),
            "kernel_ms": {"ray_march": march_avg_ms, "fan_headings": float(np.mean(head_ms)),
                          "share_of_step": march_avg_ms / (t_value_ms / args.steps) if n_gpus == 1 else None},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_source,
                         "algorithmic_bytes_per_launch": algo_bytes, "cell_visits": C, "region_probes": R,
                         "beams": n_beams, "hit_fraction": hit_frac,
                         "note": "bytes = 4*C + 4*R + 40 per beam (SURVEY.md 8d), C/R counted by the kernel and proven equal to the oracle's in tests"},
            "cpu_baseline": cpu, "parity_vs_reference": parity, "parity": parity_ranks, "clocks": clocks,
            "configs": configs,
        }
        print(json.dumps(line))
    if n_gpus > 1:
        dist.barrier()
        dist.destroy_process_group()
    leave(0)  # the reference World loaded for the CPU baseline must never be destroyed


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        sys.exit(run_reference(a))
    run_ours(a)

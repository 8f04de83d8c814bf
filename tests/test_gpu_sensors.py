"""K3 parity: the fiducial-finder and blobfinder post-processing kernels against what the unmodified
reference produced (golden traces of worlds/fasr.world and worlds/lsp_test.world)."""
import os

import numpy as np
import pytest

import replay
import trace_io
from stage_b200 import rt

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module", params=["fasr_300", "lsp_test_20", "everything_40", "pioneer_flocking_12"])
def run(request):
    tr = trace_io.load(os.path.join(GOLDEN, request.param + ".trc.xz"))
    g = replay.GpuReplay(tr, strict=False, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL)
    yield request.param, tr, g
    g.ctx.close()


def test_fiducials_match_the_reference(run):
    name, tr, g = run
    assert len(g.fiducials) == (40 if name == "lsp_test_20" else len([e for e in tr.events if e["type"] == trace_io.EV_FIDUCIAL]))
    seen = 0
    for ev, d, out, cnt in g.fiducials:
        ref = d["found"]
        assert cnt == len(ref), (ev, cnt, ref)
        for o, r in zip(out, ref):
            assert int(d["targets"][o["target"], 0]) == int(r[0])   # which model (Fiducial::mod), same order
            assert o["id"] == int(r[1])
            assert abs(o["range"] - r[2]) <= 1e-9 and abs(o["bearing"] - r[3]) <= 1e-9
            assert np.allclose(o["geom"], r[4:8], rtol=0, atol=1e-12)
            assert np.array_equal(o["pose"], r[8:12])
            seen += 1
    assert seen >= (20 if name == "lsp_test_20" else 0)  # fasr's robots are not near a charger in the first 300 ticks


def test_blobs_match_the_reference(run):
    name, tr, g = run
    if name in ("fasr_300", "pioneer_flocking_12"):
        pytest.skip("no blobfinder in this world")
    colors = {int(m["id"]): m["color"][:3] for m in tr.models}
    assert len(g.blobs) == 40
    for ev, d, out, cnt in g.blobs:
        ref = d["blobs"]
        assert cnt == len(ref)
        for o, r in zip(out, ref):
            assert np.array_equal(colors[int(o["model"])], r[:3])
            assert (o["left"], o["top"], o["right"], o["bottom"]) == tuple(int(x) for x in r[3:7])
            assert abs(o["range"] - r[7]) <= 1e-5
    assert sum(c for *_, c in g.blobs) >= (60 if name == "lsp_test_20" else 0)


def test_fiducials_synthetic_against_oracle():
    """Many finders x many targets on a synthetic world: candidate window, key / range / field-of-view
    culls, line-of-sight rays and the order of the records, against the T1 restatement."""
    import ctypes
    import oracle_lib
    from stage_b200 import worldgen
    from test_gpu_synthetic import build_both
    w = worldgen.make_world(seed=11, n_rects=150, n_robots=300, half_extent_m=20.48)
    ctx, t1 = build_both(w)
    rng = np.random.RandomState(5)
    n_t = 200
    targets = np.zeros(n_t, rt.FID_TARGET_DTYPE)
    targets["model"] = w.robot_ids[:n_t]
    targets["fiducial_return"] = rng.randint(1, 9, n_t)
    targets["fiducial_key"] = rng.randint(0, 2, n_t)
    targets["x"], targets["y"] = w.robots[:n_t, 0], w.robots[:n_t, 1]
    targets["z"] = 0.0
    targets["a"] = w.robots[:n_t, 2] * (np.pi / 180)
    targets["sx"], targets["sy"], targets["sz"] = w.robot_size
    n_f = 100
    finders = np.zeros(n_f, rt.FID_FINDER_DTYPE)
    fid_robots = np.arange(n_t - 50, n_t + 50)  # half of the finders sit on robots that are targets themselves
    finders["finder"] = w.robot_ids[fid_robots]
    finders["fiducial_key"] = rng.randint(0, 2, n_f)
    finders["ignore_zloc"] = rng.randint(0, 2, n_f)
    finders["layer"] = 1
    for k, col in (("x", 0), ("y", 1)):
        finders[k] = w.robots[fid_robots, col]
        finders["o" + k] = w.robots[fid_robots, col]
    finders["a"] = finders["oa"] = w.robots[fid_robots, 2] * (np.pi / 180)
    finders["z"], finders["oz"] = 0.0, rng.choice([0.3, 0.9], n_f)  # 0.9 is above every body: only ignore_zloc sees
    finders["max_range_anon"] = rng.uniform(3.0, 9.0, n_f)
    finders["max_range_id"] = rng.uniform(1.0, 6.0, n_f)
    finders["fov"] = rng.uniform(0.5, 2 * np.pi, n_f)
    out, cnt = ctx.fiducials_submit(targets, finders, max_per_finder=64)
    ctx.sync()
    L = oracle_lib.t1()
    L.t1_fiducial_update.restype = ctypes.c_uint32
    L.t1_fiducial_update.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
    total = 0
    for i in range(n_f):
        ref = np.zeros(64, rt.FIDUCIAL_DTYPE)
        f = np.ascontiguousarray(finders[i:i + 1])
        n = L.t1_fiducial_update(t1.h, n_t, targets.ctypes.data, f.ctypes.data, 64, ref.ctypes.data)
        assert n == cnt[i], (i, n, cnt[i])
        for o, r in zip(out[i, :n], ref[:n]):
            assert o["target"] == r["target"] and o["id"] == r["id"]
            assert abs(o["range"] - r["range"]) <= 1e-9 and abs(o["bearing"] - r["bearing"]) <= 1e-9
            assert np.allclose(o["geom"], r["geom"], rtol=0, atol=1e-12) and np.array_equal(o["pose"], r["pose"])
        total += n
    assert total > 50
    ctx.close()


def _fid_world(n_robots, n_targets, n_finders, half, seed=21):
    import oracle_lib  # noqa: F401
    from stage_b200 import worldgen
    from test_gpu_synthetic import build_both
    w = worldgen.make_world(seed=seed, n_rects=int(150 * (half / 20.48) ** 2), n_robots=n_robots, half_extent_m=half)
    ctx, t1 = build_both(w)
    rng = np.random.RandomState(seed)
    targets = np.zeros(n_targets, rt.FID_TARGET_DTYPE)
    targets["model"] = w.robot_ids[:n_targets]
    targets["fiducial_return"] = rng.randint(1, 9, n_targets)
    targets["x"], targets["y"] = w.robots[:n_targets, 0], w.robots[:n_targets, 1]
    targets["a"] = w.robots[:n_targets, 2] * (np.pi / 180)
    targets["sx"], targets["sy"], targets["sz"] = w.robot_size
    fr = np.arange(n_robots - n_finders, n_robots)
    finders = np.zeros(n_finders, rt.FID_FINDER_DTYPE)
    finders["finder"] = w.robot_ids[fr]
    finders["layer"] = 1
    for k, col in (("x", 0), ("y", 1)):
        finders[k] = finders["o" + k] = w.robots[fr, col]
    finders["a"] = finders["oa"] = w.robots[fr, 2] * (np.pi / 180)
    finders["oz"] = 0.3
    finders["max_range_anon"] = rng.uniform(2.0, 6.0, n_finders)
    finders["max_range_id"] = 3.0
    finders["fov"] = rng.uniform(1.0, 2 * np.pi, n_finders)
    return w, ctx, t1, targets, finders


def test_fiducial_grid_query_equals_brute_force_and_oracle(monkeypatch):
    """Row f-3: the uniform-grid candidate query gives exactly the brute-force result (records and
    their order), and both match the T1 restatement."""
    import ctypes
    import oracle_lib
    w, ctx, t1, targets, finders = _fid_world(3000, 2500, 1500, 40.96)
    monkeypatch.setenv("STG_RT_FID_GRID", "0")
    out_b, cnt_b = ctx.fiducials_submit(targets, finders, max_per_finder=48)
    ctx.sync()
    monkeypatch.setenv("STG_RT_FID_GRID", "1")
    out_g, cnt_g = ctx.fiducials_submit(targets, finders, max_per_finder=48)
    ctx.sync()
    assert np.array_equal(cnt_b, cnt_g) and cnt_b.max() < 48 and cnt_b.sum() > 500
    for i in range(len(finders)):
        assert out_b[i, :cnt_b[i]].tobytes() == out_g[i, :cnt_g[i]].tobytes()
    L = oracle_lib.t1()
    L.t1_fiducial_update.restype = ctypes.c_uint32
    L.t1_fiducial_update.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
    for i in range(0, len(finders), 5):
        ref = np.zeros(48, rt.FIDUCIAL_DTYPE)
        f = np.ascontiguousarray(finders[i:i + 1])
        n = L.t1_fiducial_update(t1.h, len(targets), targets.ctypes.data, f.ctypes.data, 48, ref.ctypes.data)
        assert n == cnt_g[i]
        assert np.array_equal(out_g[i, :n]["target"], ref[:n]["target"]) and np.array_equal(out_g[i, :n]["id"], ref[:n]["id"])
        assert np.max(np.abs(out_g[i, :n]["range"] - ref[:n]["range"]), initial=0) <= 1e-9
    ctx.close()


def test_fiducial_query_at_config4_scale_timing():
    """BASELINE config 4's fiducial load: 100k robots, each a target and a finder, on the config-3
    obstacle field (detection range cut from 8 m to 3 m: at 3.7 robots/m^2 the +-3 m window already
    holds ~130 candidates per finder). The robots carry no bodies here (ignore_zloc finders: a target
    is seen unless an obstacle is in the way). Times the device path - the oracle would need hours -
    and checks sanity: order, ranges, and agreement with the oracle on a few finders."""
    import ctypes
    import time
    import oracle_lib
    from stage_b200 import worldgen
    from test_gpu_synthetic import build_both
    n = 100000
    w = worldgen.make_world(seed=1, n_rects=4096, n_robots=1, half_extent_m=81.92)
    x0, y0, nx, ny = w.sreg_extent()
    first_id = w.n_models + 1
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=first_id + n + 1, max_blocks=len(w.obstacles) + 2, max_cell_entries=1 << 22)
    worldgen.load_into_context(ctx, w)
    t1 = oracle_lib.T1World(w.ppm)
    for mid in [0] + list(w.obstacle_ids) + list(w.robot_ids):
        t1.model_upsert(int(mid), int(mid), 0.5)
    polys = worldgen.obstacle_polygons(w) + worldgen.robot_polygons(w)
    for b, (p, m) in enumerate(zip(polys, np.concatenate([w.obstacle_ids, w.robot_ids]))):
        t1.block_map(b, int(m), 1, 0.0, 1.0, p)
    rng = np.random.RandomState(33)
    ids = first_id + np.arange(n)
    for m in ids:
        ctx.model_upsert(int(m), int(m), 0.5, 1)
        t1.model_upsert(int(m), int(m), 0.5, 1)
    xy = rng.uniform(-80, 80, (n, 2))
    targets = np.zeros(n, rt.FID_TARGET_DTYPE)
    targets["model"], targets["fiducial_return"] = ids, rng.randint(1, 9, n)
    targets["x"], targets["y"], targets["a"] = xy[:, 0], xy[:, 1], rng.uniform(-np.pi, np.pi, n)
    targets["sx"] = targets["sy"] = targets["sz"] = 0.1
    finders = np.zeros(n, rt.FID_FINDER_DTYPE)
    finders["finder"], finders["layer"], finders["ignore_zloc"] = ids, 1, 1
    for k in ("x", "y", "a"):
        finders[k] = finders["o" + k] = targets[k]
    finders["oz"] = 0.5
    finders["max_range_anon"], finders["max_range_id"], finders["fov"] = 3.0, 2.0, np.pi
    # warm-up: sizes the device and pinned buffers, and faults in the caller's (845 MB, fixed-stride) result array
    out, cnt = ctx.fiducials_submit(targets, finders, max_per_finder=96)
    ctx.sync()
    first = out.copy()
    t0 = time.perf_counter()
    ctx.fiducials_submit(targets, finders, max_per_finder=96, into=(out, cnt))
    t1s = time.perf_counter()
    ctx.sync()
    dt = time.perf_counter() - t0
    seen = int(cnt.sum())
    assert np.array_equal(out.view(np.uint8), first.view(np.uint8))
    print("config-4 scale fiducial query: 100k finders x 100k targets, %d fiducials seen (max %d per finder), %.1f ms end to end "
          "(submit call %.1f ms; only the %.0f MB of records cross PCIe, packed)"
          % (seen, cnt.max(), 1e3 * dt, 1e3 * (t1s - t0), seen * 88 / 1e6))
    assert seen > 100000 and cnt.max() <= 96
    # packed hand-over: the same records, finder after finder
    pk, pc = ctx.fiducials_submit_packed(targets, finders, 96, 1 << 20)
    ctx.sync()
    t0 = time.perf_counter()
    ctx.fiducials_submit_packed(targets, finders, 96, 1 << 20, into=(pk, pc))
    ctx.sync()
    dt_packed = time.perf_counter() - t0
    assert np.array_equal(pc, cnt)
    at = np.concatenate([[0], np.cumsum(pc.astype(np.int64))]).astype(np.int64)
    for f in list(range(0, n, 997)) + [n - 1]:
        assert np.array_equal(pk[int(at[f]):int(at[f + 1])].view(np.uint8), out[f, :int(cnt[f])].view(np.uint8))
    print("  packed hand-over (stg_rt_fiducials_submit_packed): %.1f ms end to end" % (1e3 * dt_packed))
    L = oracle_lib.t1()
    L.t1_fiducial_update.restype = ctypes.c_uint32
    L.t1_fiducial_update.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
    for i in range(0, n, 9973):
        tg = out[i, :cnt[i]]["target"]
        assert np.all(np.diff(tg.astype(np.int64)) > 0)  # ascending target order
        ref = np.zeros(96, rt.FIDUCIAL_DTYPE)
        f = np.ascontiguousarray(finders[i:i + 1])
        k = L.t1_fiducial_update(t1.h, n, targets.ctypes.data, f.ctypes.data, 96, ref.ctypes.data)
        assert k == cnt[i] and np.array_equal(ref[:k]["target"], tg) and np.array_equal(ref[:k]["id"], out[i, :k]["id"])
    ctx.close()


def test_config4_sonar_rays_full_size_timing():
    """BASELINE config 4's rangers at full size: 100,000 robots x 16 one-beam sonars (range 5 m, the pioneer ring of
    pioneer.inc:24-39: bearings -90..90 and back) = 1.6 M single-beam fans per tick on the config-3 map. Times a tick
    through the ABI and checks a sample of the beams against the oracle bit for bit."""
    import time
    import oracle_lib
    from stage_b200 import worldgen
    w = worldgen.make_world(seed=1, n_rects=4096, n_robots=1, half_extent_m=81.92)
    x0, y0, nx, ny = w.sreg_extent()
    n_robots, n_sonar = 100000, 16
    first_id = w.n_models + 1
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=first_id + n_robots + 1, max_blocks=len(w.obstacles) + 2, max_cell_entries=1 << 22)
    worldgen.load_into_context(ctx, w)
    t1 = oracle_lib.T1World(w.ppm)
    for mid in [0] + list(w.obstacle_ids) + list(w.robot_ids):
        t1.model_upsert(int(mid), int(mid), 0.5)
    polys = worldgen.obstacle_polygons(w) + worldgen.robot_polygons(w)
    for b, (p, m) in enumerate(zip(polys, np.concatenate([w.obstacle_ids, w.robot_ids]))):
        t1.block_map(b, int(m), 1, 0.0, 1.0, p)
    rng = np.random.RandomState(44)
    ids = first_id + np.arange(n_robots)
    for m in ids[:: n_robots // 200]:   # the oracle only needs the finders it is asked about
        t1.model_upsert(int(m), int(m), 0.5)
    for m in ids:
        ctx.model_upsert(int(m), int(m), 0.5)
    pose = np.stack([rng.uniform(-80, 80, n_robots), rng.uniform(-80, 80, n_robots), rng.uniform(-np.pi, np.pi, n_robots)], 1)
    ring = np.deg2rad([90, 50, 30, 10, -10, -30, -50, -90, -90, -130, -150, -170, 170, 150, 130, 90])
    fans = rt.make_fans(n_robots * n_sonar)
    fans["finder"] = np.repeat(ids, n_sonar)
    fans["n_samples"], fans["pred"], fans["ztest"], fans["layer"], fans["angles"] = 1, rt.PRED_RANGER, 1, 1, rt.ANGLES_RANGER
    a = (pose[:, 2:3] + ring[None, :]).reshape(-1)
    fans["ox"] = np.repeat(pose[:, 0], n_sonar) + 0.15 * np.cos(a)
    fans["oy"] = np.repeat(pose[:, 1], n_sonar) + 0.15 * np.sin(a)
    fans["oz"], fans["oa"], fans["range"], fans["fov"] = 0.3, np.arctan2(np.sin(a), np.cos(a)), 5.0, 0.26
    out = rt.Outputs(len(fans), rt.WANT_RANGES | rt.WANT_HIT_MODEL, alloc=ctx.pinned_empty)
    ctx.fans_submit(fans, out)
    ctx.sync()  # warm-up: sizes the buffers
    t0 = time.perf_counter()
    ctx.fans_submit(fans, out)
    t_sub = time.perf_counter() - t0
    ctx.sync()
    dt = time.perf_counter() - t0
    kt = ctx.kernel_times()
    for i in range(0, len(fans), len(fans) // 3000):
        f = fans[i]
        if int(f["finder"]) not in set(int(m) for m in ids[:: n_robots // 200]):
            t1.model_upsert(int(f["finder"]), int(f["finder"]), 0.5)
        rg, it, be, hits = t1.ranger_fan(int(f["finder"]), 1, [f["ox"], f["oy"], f["oz"], f["oa"]], 5.0, 0.26, 1)
        assert out.ranges[i] == rg[0] and out.hit_model[i] == hits["hit_model"][0]
    # the same sensors as whole ranger models: 40 bytes per robot instead of 64 per sonar
    sensors = np.zeros(n_sonar, rt.RANGER_SENSOR_DTYPE)
    sensors["px"], sensors["py"], sensors["pa"] = 0.15 * np.cos(ring), 0.15 * np.sin(ring), ring
    sensors["pz"], sensors["size_z"], sensors["range_max"], sensors["fov"], sensors["n_samples"] = 0.3, 0.0, 5.0, 0.26, 1
    rp = np.zeros(n_robots, rt.RANGER_POSE_DTYPE)
    rp["finder"], rp["layer"], rp["x"], rp["y"], rp["a"] = ids, 1, pose[:, 0], pose[:, 1], pose[:, 2]
    out2 = rt.Outputs(len(fans), rt.WANT_RANGES | rt.WANT_HIT_MODEL, alloc=ctx.pinned_empty)
    ctx.rangers_submit(sensors, rp, out2)
    ctx.sync()
    t0 = time.perf_counter()
    ctx.rangers_submit(sensors, rp, out2)
    ctx.sync()
    dt_rig = time.perf_counter() - t0
    same = np.mean((out2.hit_model == out.hit_model) & (np.abs(out2.ranges - out.ranges) < 1e-9))
    assert same > 0.999  # the fan origins above were put together with numpy's cos / sin, not operation for operation
    print("config-4 sonars as 100 k ranger models (stg_rt_rangers_submit): %.2f ms per tick end to end, %.0f M rays/s" % (1e3 * dt_rig, len(fans) / dt_rig / 1e6))
    print("config-4 sonars: 1.6 M one-beam fans: %.2f ms per tick end to end (submit call %.2f ms: 102 MB of fan records staged and "
          "uploaded; headings %.3f ms, march %.3f ms on the device), %.0f M rays/s" % (1e3 * dt, 1e3 * t_sub, kt["headings_ms"], kt["march_ms"], len(fans) / dt / 1e6))
    ctx.close()


def host_ranger_fans(sensors, poses):
    """the fan records ModelRanger::Sensor::Update sets up (model_ranger.cc:217-230), with the host C library"""
    import math
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


def test_whole_ranger_models_equal_their_fans():
    """stg_rt_rangers_submit derives on the device the fans a caller of stg_rt_fans_submit would have set up on the
    host: same beams, same bits. A sonar ring of 16 one-beam sensors and a rig mixing a 180-beam laser with sonars."""
    from stage_b200 import worldgen
    w = worldgen.make_world(seed=6, n_rects=600, n_robots=300, half_extent_m=30.72)
    x0, y0, nx, ny = w.sreg_extent()
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + 1, max_blocks=len(w.obstacles) + len(w.robots), max_cell_entries=1 << 20)
    worldgen.load_into_context(ctx, w)
    rng = np.random.RandomState(2)
    poses = np.zeros(len(w.robots), rt.RANGER_POSE_DTYPE)
    poses["finder"], poses["layer"] = w.robot_ids, 1
    poses["x"], poses["y"], poses["z"] = w.robots[:, 0], w.robots[:, 1], 0.1
    poses["a"] = np.deg2rad(w.robots[:, 2])
    poses["a"][:5] = [0.0, np.pi, -np.pi, 3.1, -3.1]  # the normalisation's edges
    ring = np.deg2rad([90, 50, 30, 10, -10, -30, -50, -90, -90, -130, -150, -170, 170, 150, 130, 90])
    sonar = np.zeros(16, rt.RANGER_SENSOR_DTYPE)
    sonar["px"], sonar["py"], sonar["pa"] = 0.15 * np.cos(ring), 0.15 * np.sin(ring), ring
    sonar["size_z"], sonar["range_max"], sonar["fov"], sonar["n_samples"] = 0.04, 5.0, 0.26, 1
    mixed = np.zeros(4, rt.RANGER_SENSOR_DTYPE)
    mixed["px"], mixed["py"], mixed["pa"] = [0.1, 0.0, -0.1, 0.05], [0.0, 0.1, 0.0, -0.1], [0.0, 1.5, 3.0, -1.5]
    mixed["pz"], mixed["size_z"], mixed["range_max"] = 0.2, 0.1, [8.0, 5.0, 5.0, 2.0]
    mixed["fov"], mixed["n_samples"] = [np.pi, 0.26, 0.5, 0.26], [180, 1, 7, 1]
    for sensors in (sonar, mixed):
        n_beams = int(sensors["n_samples"].sum()) * len(poses)
        a = rt.Outputs(n_beams, rt.WANT_RANGES | rt.WANT_INTENSITIES | rt.WANT_BEARINGS | rt.WANT_HIT_MODEL)
        ctx.rangers_submit(sensors, poses, a)
        ctx.sync()
        b = ctx.trace(host_ranger_fans(sensors, poses), want=rt.WANT_RANGES | rt.WANT_INTENSITIES | rt.WANT_BEARINGS | rt.WANT_HIT_MODEL)
        assert np.array_equal(a.ranges.view(np.uint64), b.ranges.view(np.uint64)) and np.array_equal(a.hit_model, b.hit_model)
        assert np.array_equal(a.intensities, b.intensities) and np.array_equal(a.bearings.view(np.uint64), b.bearings.view(np.uint64))
        assert (a.hit_model >= 0).mean() > 0.3
    ctx.close()

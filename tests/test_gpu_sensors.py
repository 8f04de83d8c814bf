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


@pytest.fixture(scope="module", params=["fasr_300", "lsp_test_20"])
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
    if name != "lsp_test_20":
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
    assert sum(c for *_, c in g.blobs) >= 60


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

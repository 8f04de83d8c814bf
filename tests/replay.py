"""Drive the C ABI (stage_b200.rt.Context) and/or the T1 oracle with a recorded reference trace, in
the order a patched libstage would issue the calls. Test infrastructure."""
import numpy as np

import oracle_lib
import trace_io
from stage_b200 import rt


def extent_of(trace, margin_sregs=0):
    """smallest superregion rectangle (x0, y0, nx, ny) holding every mapped polygon"""
    v = trace.verts
    x0, y0 = int(v[:, 0].min()) >> 10, int(v[:, 1].min()) >> 10
    x1, y1 = int(v[:, 0].max()) >> 10, int(v[:, 1].max()) >> 10
    return x0 - margin_sregs, y0 - margin_sregs, x1 - x0 + 1 + 2 * margin_sregs, y1 - y0 + 1 + 2 * margin_sregs


def entries_bound(trace):
    """upper bound of the (cell, block) entries one layer ever holds: the outline lengths of everything mapped on it"""
    ev = trace.events[trace.events["type"] == trace_io.EV_MAP]
    worst = 0
    for layer in (0, 1):
        e = ev[ev["layer"] == layer]
        if not len(e):
            continue
        n = e["n_pts"].astype(np.int64)
        first = e["first"].astype(np.int64)
        idx = np.repeat(first, n) + (np.arange(int(n.sum())) - np.repeat(np.cumsum(n) - n, n))  # every vertex of every polygon
        last = np.repeat(first + n - 1, n)
        nxt = np.where(idx == last, np.repeat(first, n), idx + 1)                                 # its successor, cyclic
        v = trace.verts.astype(np.int64)
        worst = max(worst, int(np.abs(v[nxt] - v[idx]).sum()))
    return worst


def make_context(trace, max_cell_entries=1 << 20, **kw):
    x0, y0, nx, ny = extent_of(trace)
    need = entries_bound(trace)
    while max_cell_entries < need and max_cell_entries < (1 << 23):  # big maps (large.world: 1.9 M entries per layer)
        max_cell_entries <<= 1
    n_blocks = int(trace.events["block"][trace.events["type"] == trace_io.EV_MAP].max()) + 1
    ctx = rt.Context(trace.ppm, x0, y0, nx, ny, max_models=int(trace.models["id"].max()) + 1,
                     max_blocks=n_blocks, max_cell_entries=max_cell_entries, **kw)
    for m in trace.models:
        ctx.model_upsert(int(m["id"]), int(m["root"]), float(m["ranger_return"]), int(m["fiducial_return"]),
                         int(m["fiducial_key"]), int(m["flags"]), m["color"])
    return ctx


def make_t1(trace):
    w = oracle_lib.T1World(trace.ppm)
    for m in trace.models:
        w.model_upsert(int(m["id"]), int(m["root"]), float(m["ranger_return"]), int(m["fiducial_return"]),
                       int(m["fiducial_key"]), int(m["flags"]))
    return w


def rays_to_fans(rays):
    fans = rt.make_fans(len(rays))
    fans["finder"] = rays["finder"]
    fans["n_samples"] = 1
    fans["pred"] = rays["kind"]
    fans["ztest"] = rays["ztest"]
    fans["layer"] = rays["layer"]
    fans["angles"] = rt.ANGLES_EXPLICIT
    fans["first_heading"] = np.arange(len(rays), dtype=np.uint32)
    for k in ("ox", "oy", "oz", "oa", "range"):
        fans[k] = rays[k]
    return fans, np.ascontiguousarray(rays["oa"])


def rays_to_t1(rays):
    r = np.zeros(len(rays), oracle_lib.T1_RAY)
    for k in ("ox", "oy", "oz", "oa", "range", "finder", "kind", "ztest", "layer"):
        r[k] = rays[k]
    return r


def scan_to_fan(ev, sc):
    fan = rt.make_fans(1)
    fan["finder"] = ev["model"]
    fan["n_samples"] = sc["n"]
    fan["pred"] = rt.PRED_RANGER
    fan["ztest"] = 1
    fan["layer"] = ev["layer"]
    fan["angles"] = rt.ANGLES_RANGER
    fan["ox"], fan["oy"], fan["oz"], fan["oa"] = sc["origin"]
    fan["range"] = sc["range_max"]
    fan["fov"] = sc["fov"]
    return fan


def fid_records(ev, d):
    """FIDUCIAL trace event -> (targets, finder) arrays for stg_rt_fiducials_submit"""
    t = d["targets"]
    targets = np.zeros(len(t), rt.FID_TARGET_DTYPE)
    targets["model"], targets["fiducial_return"], targets["fiducial_key"] = t[:, 0], t[:, 1], t[:, 2]
    for i, k in enumerate(("x", "y", "z", "a", "sx", "sy", "sz")):
        targets[k] = t[:, 3 + i]
    f = np.zeros(1, rt.FID_FINDER_DTYPE)
    f["finder"], f["fiducial_key"], f["ignore_zloc"], f["layer"] = ev["model"], int(d["fiducial_key"]), int(d["ignore_zloc"]), ev["layer"]
    for k in ("x", "y", "z", "a", "ox", "oy", "oz", "oa", "max_range_anon", "max_range_id", "fov"):
        f[k] = d[k]
    return targets, f


def blob_records(ev, d):
    f = np.zeros(1, rt.BLOB_FINDER_DTYPE)
    f["finder"], f["layer"] = ev["model"], ev["layer"]
    f["scan_width"], f["scan_height"] = int(d["scan_width"]), int(d["scan_height"])
    f["first_color"], f["n_colors"] = 0, len(d["colors"])
    for k in ("ox", "oy", "oz", "oa", "range", "fov"):
        f[k] = d[k]
    return f, d["colors"]


class GpuReplay:
    """Results of pushing a whole trace through the C ABI."""

    def __init__(self, trace, strict, max_ticks=None, sync_every_tick=True, want=rt.WANT_ALL):
        self.trace = trace
        ctx = self.ctx = make_context(trace)
        n = len(trace.rays)
        self.ranges = np.full(n, np.nan)
        self.hit_model = np.full(n, -2, np.int32)
        self.hit_cell = np.zeros((n, 2), np.int32)
        self.steps = np.zeros((n, 2), np.uint32)
        self.scans = []  # (event, Outputs)
        self.fiducials = []  # (event, dict, out, count)
        self.blobs = []
        self.collisions = []  # one hit model id (-1: none) per COLLISION event
        self.touchers = []  # one ascending list of model ids per TOUCHERS event
        pending = []
        ticks = 0
        for ev in trace.events:
            t = ev["type"]
            if t == trace_io.EV_MAP:
                ctx.block_map(int(ev["block"]), int(ev["model"]), int(ev["layer"]), float(ev["zmin"]),
                              float(ev["zmax"]), trace.polygon(ev))
            elif t == trace_io.EV_UNMAP:
                ctx.block_unmap(int(ev["block"]), int(ev["layer"]))
            elif t == trace_io.EV_COLLISION:
                # the reference tests right after re-mapping the mover (model_position.cc:549-554),
                # against the grid as it stands then: one query, answered before the next event
                _, layer, blocks, _ = trace.collision(ev)
                hit = ctx.collisions_test(layer, [blocks])
                ctx.sync()
                self._collect(pending)
                self.collisions.append(int(hit[0]))
            elif t == trace_io.EV_TOUCHERS:
                _, layer, blocks, _ = trace.touchers(ev)
                ids, counts = ctx.touching_models(layer, [blocks], max_per_query=16)
                ctx.sync()
                self._collect(pending)
                self.touchers.append([int(x) for x in ids[0, :counts[0]]])
            elif t == trace_io.EV_RAYS:
                first, cnt = int(ev["first"]), int(ev["model"])
                rays = trace.rays[first:first + cnt]
                fans, headings = rays_to_fans(rays)
                out = rt.Outputs(cnt, want)
                ctx.fans_submit(fans, out, headings=headings,
                                trig=oracle_lib.libm_trig(headings) if strict else None)
                pending.append((first, cnt, out))
            elif t == trace_io.EV_RANGER_SCAN:
                sc = trace.scan(ev)
                fan = scan_to_fan(ev, sc)
                out = rt.Outputs(sc["n"], want)
                trig = None
                if strict:
                    trig = oracle_lib.libm_trig(oracle_lib.ranger_headings(sc["origin"][3], sc["fov"], sc["n"]))
                ctx.fans_submit(fan, out, trig=trig)
                self.scans.append((ev, sc, out))
            elif t == trace_io.EV_FIDUCIAL:
                d = trace.fiducial(ev)
                targets, f = fid_records(ev, d)
                out, cnt = ctx.fiducials_submit(targets, f, max_per_finder=max(len(targets), 1))
                ctx.sync()  # one fiducial batch in flight at a time
                self.fiducials.append((ev, d, out[0, :cnt[0]].copy(), int(cnt[0])))
            elif t == trace_io.EV_BLOB:
                d = trace.blob(ev)
                f, colors = blob_records(ev, d)
                out, cnt = ctx.blobs_submit(f, colors)
                ctx.sync()
                self.blobs.append((ev, d, out[0, :cnt[0]].copy(), int(cnt[0])))
            elif t == trace_io.EV_TICK:
                ticks += 1
                if sync_every_tick:
                    ctx.sync()
                    self._collect(pending)
                if max_ticks is not None and ticks >= max_ticks:
                    break
        ctx.sync()
        self._collect(pending)
        self.n_ticks = ticks

    def _collect(self, pending):
        for first, cnt, out in pending:
            if out.ranges is not None:
                self.ranges[first:first + cnt] = out.ranges
            if out.hit_model is not None:
                self.hit_model[first:first + cnt] = out.hit_model
            if out.hit_cell is not None:
                self.hit_cell[first:first + cnt] = out.hit_cell
            if out.steps is not None:
                self.steps[first:first + cnt] = out.steps
        pending.clear()


class T1Replay:
    """The same trace through the T1 oracle, keeping the per-ray hit cells and step counters."""

    def __init__(self, trace, max_ticks=None):
        w = self.world = make_t1(trace)
        self.hits = np.zeros(len(trace.rays), oracle_lib.T1_HIT)
        self.scans = []
        self.collisions = []
        self.touchers = []
        ticks = 0
        for ev in trace.events:
            t = ev["type"]
            if t == trace_io.EV_MAP:
                w.block_map(int(ev["block"]), int(ev["model"]), int(ev["layer"]), float(ev["zmin"]),
                            float(ev["zmax"]), trace.polygon(ev))
            elif t == trace_io.EV_UNMAP:
                w.block_unmap(int(ev["block"]), int(ev["layer"]))
            elif t == trace_io.EV_COLLISION:
                _, layer, blocks, _ = trace.collision(ev)
                self.collisions.append(w.test_collision(layer, blocks))
            elif t == trace_io.EV_TOUCHERS:
                _, layer, blocks, _ = trace.touchers(ev)
                self.touchers.append([int(x) for x in w.touching_models(layer, blocks)])
            elif t == trace_io.EV_RAYS:
                first, cnt = int(ev["first"]), int(ev["model"])
                self.hits[first:first + cnt] = w.raytrace(rays_to_t1(trace.rays[first:first + cnt]))
            elif t == trace_io.EV_RANGER_SCAN:
                sc = trace.scan(ev)
                self.scans.append(w.ranger_fan(int(ev["model"]), int(ev["layer"]), sc["origin"], sc["range_max"],
                                               sc["fov"], sc["n"]))
            elif t == trace_io.EV_TICK:
                ticks += 1
                if max_ticks is not None and ticks >= max_ticks:
                    break
        self.n_ticks = ticks

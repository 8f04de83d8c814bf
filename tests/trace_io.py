"""Reader for the reference event traces written by oracle/ref_harness.cc (format:
oracle/trace_format.h). Test infrastructure only."""
import gzip
import numpy as np

MAGIC = b"STGTRC02"
EV_MAP, EV_UNMAP, EV_RAYS, EV_TICK = 1, 2, 3, 4
KIND_RANGER, KIND_UNRELATED, KIND_GRIPPER = 0, 1, 2

HEADER = np.dtype([("magic", "S8"), ("ppm", "<f8"), ("n_models", "<u4"), ("n_events", "<u4"),
                   ("n_verts", "<u4"), ("n_rays", "<u4"), ("n_ticks", "<u4"), ("reserved", "<u4")])
MODEL = np.dtype([("id", "<u4"), ("parent", "<u4"), ("root", "<u4"), ("fiducial_return", "<i4"),
                  ("fiducial_key", "<i4"), ("flags", "<u4"), ("ranger_return", "<f8"),
                  ("color", "<f8", (4,)), ("type", "S16")])
EVENT = np.dtype([("type", "u1"), ("layer", "u1"), ("n_pts", "<u2"), ("block", "<u4"),
                  ("model", "<u4"), ("first", "<u4"), ("zmin", "<f8"), ("zmax", "<f8")])
RAY = np.dtype([("ox", "<f8"), ("oy", "<f8"), ("oz", "<f8"), ("oa", "<f8"), ("range", "<f8"),
                ("res_range", "<f8"), ("hit_model", "<i4"), ("finder", "<u4"), ("kind", "u1"),
                ("ztest", "u1"), ("layer", "u1"), ("pad", "u1", (5,))])
assert HEADER.itemsize == 40 and MODEL.itemsize == 80 and EVENT.itemsize == 32 and RAY.itemsize == 64


class Trace:
    def __init__(self, header, models, events, verts, rays):
        self.ppm = float(header["ppm"])
        self.n_ticks = int(header["n_ticks"])
        self.models, self.events, self.verts, self.rays = models, events, verts, rays

    def polygon(self, ev):
        """(n_pts, 2) int32 pixel-space vertices of a MAP event."""
        return self.verts[ev["first"]:ev["first"] + ev["n_pts"]]


def load(path):
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        buf = f.read()
    h = np.frombuffer(buf, HEADER, 1)[0]
    if h["magic"] != MAGIC:
        raise ValueError("not a %s trace: %r" % (MAGIC.decode(), path))
    off = HEADER.itemsize
    models = np.frombuffer(buf, MODEL, h["n_models"], off); off += models.nbytes
    events = np.frombuffer(buf, EVENT, h["n_events"], off); off += events.nbytes
    verts = np.frombuffer(buf, "<i4", 2 * int(h["n_verts"]), off).reshape(-1, 2); off += verts.nbytes
    rays = np.frombuffer(buf, RAY, h["n_rays"], off); off += rays.nbytes
    if off != len(buf):
        raise ValueError("trailing bytes in trace %r" % (path,))
    return Trace(h, models, events, verts, rays)

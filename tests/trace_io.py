"""Reader for the reference event traces written by oracle/ref_harness.cc (format:
oracle/trace_format.h). Test infrastructure only."""
import gzip
import lzma

import numpy as np

MAGIC = b"STGTRC05"
EV_MAP, EV_UNMAP, EV_RAYS, EV_TICK, EV_RANGER_SCAN, EV_FIDUCIAL, EV_BLOB, EV_COLLISION = 1, 2, 3, 4, 5, 6, 7, 8
EV_TOUCHERS = 9
SCAN_HEADER_DOUBLES = 6
KIND_RANGER, KIND_UNRELATED, KIND_GRIPPER = 0, 1, 2

HEADER = np.dtype([("magic", "S8"), ("ppm", "<f8"), ("n_models", "<u4"), ("n_events", "<u4"),
                   ("n_verts", "<u4"), ("n_rays", "<u4"), ("n_ticks", "<u4"), ("n_aux", "<u4")])
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
    def __init__(self, header, models, events, verts, rays, aux):
        self.ppm = float(header["ppm"])
        self.n_ticks = int(header["n_ticks"])
        self.models, self.events, self.verts, self.rays, self.aux = models, events, verts, rays, aux

    def scan(self, ev):
        """RANGER_SCAN event -> dict(origin[4], range_max, fov, n, ranges, intensities, bearings)."""
        n, a = int(ev["n_pts"]), self.aux[ev["first"]:]
        h = SCAN_HEADER_DOUBLES
        return dict(origin=a[0:4], range_max=float(a[4]), fov=float(a[5]), n=n, ranges=a[h:h + n],
                    intensities=a[h + n:h + 2 * n], bearings=a[h + 2 * n:h + 3 * n])

    def fiducial(self, ev):
        """FIDUCIAL event -> dict(finder header fields, targets[n,10], found[m,12])."""
        a = self.aux[ev["first"]:]
        nt, nf = int(a[13]), int(ev["block"])
        keys = ("x y z a ox oy oz oa max_range_anon max_range_id fov fiducial_key ignore_zloc").split()
        d = {k: float(a[i]) for i, k in enumerate(keys)}
        d["targets"] = a[14:14 + 10 * nt].reshape(nt, 10)
        d["found"] = a[14 + 10 * nt:14 + 10 * nt + 12 * nf].reshape(nf, 12)
        return d

    def blob(self, ev):
        """BLOB event -> dict(header fields, colors[n,3], blobs[m,8])."""
        a = self.aux[ev["first"]:]
        nc, nb = int(a[8]), int(ev["block"])
        keys = "ox oy oz oa range fov scan_width scan_height".split()
        d = {k: float(a[i]) for i, k in enumerate(keys)}
        d["colors"] = a[9:9 + 3 * nc].reshape(nc, 3)
        d["blobs"] = a[9 + 3 * nc:9 + 3 * nc + 8 * nb].reshape(nb, 8)
        return d

    def collision(self, ev):
        """COLLISION event -> (model, layer, blocks in visiting order, hit model id or -1)"""
        n = int(ev["n_pts"])
        return int(ev["model"]), int(ev["layer"]), self.aux[ev["first"]:ev["first"] + n].astype(np.uint32), int(ev["block"]) - 1

    def touchers(self, ev):
        """TOUCHERS event -> (model, layer, the model's own blocks, ids of the touching models, ascending)"""
        n, m, f = int(ev["n_pts"]), int(ev["block"]), int(ev["first"])
        return int(ev["model"]), int(ev["layer"]), self.aux[f:f + n].astype(np.uint32), self.aux[f + n:f + n + m].astype(np.uint32)

    def polygon(self, ev):
        """(n_pts, 2) int32 pixel-space vertices of a MAP event."""
        return self.verts[ev["first"]:ev["first"] + ev["n_pts"]]


def load(path):
    path = str(path)
    opener = gzip.open if path.endswith(".gz") else lzma.open if path.endswith(".xz") else open
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
    aux = np.frombuffer(buf, "<f8", h["n_aux"], off); off += aux.nbytes
    if off != len(buf):
        raise ValueError("trailing bytes in trace %r" % (path,))
    return Trace(h, models, events, verts, rays, aux)

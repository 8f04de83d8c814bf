"""Reader for the binary dumps integration/stage_run.cc writes (one record per tick: every model's id and global pose,
every ranger sensor's ranges / intensities / bearings, every fiducial finder's records, every blobfinder's blobs)."""
import struct

import numpy as np


def read(path):
    """-> list of ticks; a tick is {model id: {"pose": (4,), "sensors": [(ranges, intensities, bearings)] | None,
    "fiducials": [(range, bearing, geom(4), pose(4), model id, fiducial id)] | None}}"""
    raw = open(path, "rb").read()
    off, ticks = 0, []
    cur = None
    while off < len(raw):
        (word,) = struct.unpack_from("<I", raw, off)
        if word == 0x5449434b:
            cur = {}
            ticks.append(cur)
            off += 8
            continue
        mid, kind = struct.unpack_from("<II", raw, off)
        pose = struct.unpack_from("<4d", raw, off + 8)
        off += 40
        rec = {"pose": pose, "sensors": None, "fiducials": None, "blobs": None}
        if kind == 1:
            (ns,) = struct.unpack_from("<I", raw, off)
            off += 4
            rec["sensors"] = []
            for _ in range(ns):
                (n,) = struct.unpack_from("<I", raw, off)
                off += 4
                a = np.frombuffer(raw, np.float64, 3 * n, off).reshape(3, n)
                off += 24 * n
                rec["sensors"].append((a[0], a[1], a[2]))
        elif kind == 2:
            (nf,) = struct.unpack_from("<I", raw, off)
            off += 4
            rec["fiducials"] = []
            for _ in range(nf):
                d = struct.unpack_from("<10d", raw, off)
                mod, fid = struct.unpack_from("<Ii", raw, off + 80)
                off += 88
                rec["fiducials"].append((d[0], d[1], d[2:6], d[6:10], mod, fid))
        elif kind == 3:
            (nb,) = struct.unpack_from("<I", raw, off)
            off += 4
            rec["blobs"] = []
            for _ in range(nb):
                d = struct.unpack_from("<5d", raw, off)
                box = struct.unpack_from("<4I", raw, off + 40)
                off += 56
                rec["blobs"].append((d[:4], d[4], box))   # (rgba, range, (left, top, right, bottom))
        cur[mid] = rec
    return ticks


def describe_first_difference(a_path, b_path):
    """where two dumps part: (tick, model id, what) or None"""
    A, B = read(a_path), read(b_path)
    for t, (ta, tb) in enumerate(zip(A, B)):
        for mid in sorted(ta):
            ra, rb = ta[mid], tb.get(mid)
            if rb is None:
                return t, mid, "model missing"
            if ra["pose"] != rb["pose"]:
                return t, mid, "pose %r vs %r" % (ra["pose"], rb["pose"])
            if ra["sensors"] is not None:
                for k, (sa, sb) in enumerate(zip(ra["sensors"], rb["sensors"])):
                    for name, xa, xb in zip(("ranges", "intensities", "bearings"), sa, sb):
                        if len(xa) != len(xb) or not np.array_equal(xa.view(np.uint64), xb.view(np.uint64)):
                            bad = np.nonzero(xa.view(np.uint64) != xb.view(np.uint64))[0] if len(xa) == len(xb) else []
                            return t, mid, "sensor %d %s: %d of %d differ, first at beam %s: %r vs %r" % (
                                k, name, len(bad), len(xa), bad[:1], xa[bad[:1]], xb[bad[:1]])
            if ra["fiducials"] is not None and ra["fiducials"] != rb["fiducials"]:
                return t, mid, "fiducials %r vs %r" % (ra["fiducials"][:3], rb["fiducials"][:3])
            if ra["blobs"] is not None and ra["blobs"] != rb["blobs"]:
                return t, mid, "blobs %r vs %r" % (ra["blobs"][:3], rb["blobs"][:3])
    if len(A) != len(B):
        return min(len(A), len(B)), None, "tick counts %d vs %d" % (len(A), len(B))
    return None


def first_difference(a_path, b_path):
    """byte offset of the first difference, or None when the files are identical"""
    a, b = np.fromfile(a_path, np.uint8), np.fromfile(b_path, np.uint8)
    n = min(len(a), len(b))
    d = np.nonzero(a[:n] != b[:n])[0]
    if len(d):
        return int(d[0])
    return None if len(a) == len(b) else n


def tick_of_offset(path, offset):
    """index of the tick record that contains a byte offset (tick records start with the marker 'KCIT' + index)"""
    raw = np.fromfile(path, np.uint8)
    marks = np.nonzero((raw[:-3] == 0x4b) & (raw[1:-2] == 0x43) & (raw[2:-1] == 0x49) & (raw[3:] == 0x54))[0]
    marks = marks[marks % 4 == 0] if len(marks) else marks
    k = int(np.searchsorted(marks, offset, side="right")) - 1
    if k < 0:
        return None
    return struct.unpack("<I", raw[marks[k] + 4:marks[k] + 8].tobytes())[0]

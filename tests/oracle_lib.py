"""ctypes access to the two oracles. TEST INFRASTRUCTURE ONLY (see oracle/README.md).

T1 = oracle/libt1_oracle.so  (restatement, built by `make -C oracle`; always available)
T0 = oracle/_ref/libref_harness.so + libstage_ref.so  (the unmodified reference, built by
     oracle/build_ref.sh where /root/reference exists; prebuilt files travel to the GPU box)
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
REF = os.path.join(ORACLE, "_ref")

T1_RAY = np.dtype([("ox", "<f8"), ("oy", "<f8"), ("oz", "<f8"), ("oa", "<f8"), ("range", "<f8"),
                   ("finder", "<u4"), ("kind", "u1"), ("ztest", "u1"), ("layer", "u1"), ("pad", "u1")])
T1_HIT = np.dtype([("range", "<f8"), ("hit_model", "<i4"), ("hit_x", "<i4"), ("hit_y", "<i4"),
                   ("cell_visits", "<u4"), ("region_probes", "<u4"), ("pad", "<u4")])
REF_RAY = np.dtype([("ox", "<f8"), ("oy", "<f8"), ("oz", "<f8"), ("oa", "<f8"), ("range", "<f8"),
                    ("finder", "<u4"), ("kind", "u1"), ("ztest", "u1"), ("pad", "u1", (2,))])
REF_HIT = np.dtype([("range", "<f8"), ("hit_model", "<i4"), ("pad", "<i4")])
assert T1_RAY.itemsize == 48 and T1_HIT.itemsize == 32 and REF_RAY.itemsize == 48 and REF_HIT.itemsize == 16


class ReplayStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in (
        "rays", "range_mismatch", "model_mismatch", "maps", "unmaps", "ticks", "cell_visits",
        "region_probes", "scans", "scan_mismatch", "fiducial_updates", "fiducial_mismatch", "blob_updates",
        "blob_mismatch")] + [("max_range_err", ctypes.c_double)] + [(n, ctypes.c_uint64) for n in (
        "collision_tests", "collision_mismatch", "collision_hits", "toucher_calls", "toucher_mismatch", "toucher_models")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_t1 = None


def t1():
    global _t1
    if _t1 is None:
        so = os.path.join(ORACLE, "libt1_oracle.so")
        src = os.path.join(ORACLE, "t1_oracle.cc")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-C", ORACLE, "libt1_oracle.so"], check=True, capture_output=True)
        L = ctypes.CDLL(so)
        vp, u32, i32, f64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int32, ctypes.c_double
        L.t1_create.restype, L.t1_create.argtypes = vp, [f64]
        L.t1_destroy.argtypes = [vp]
        L.t1_model_upsert.argtypes = [vp, u32, u32, f64, i32, i32, u32]
        L.t1_block_map.argtypes = [vp, u32, u32, ctypes.c_int, f64, f64, ctypes.c_int, vp]
        L.t1_block_unmap.argtypes = [vp, u32, ctypes.c_int]
        L.t1_raytrace.argtypes = [vp, ctypes.c_uint64, vp, vp, vp]
        L.t1_libm_trig.argtypes = [ctypes.c_uint64, vp, vp]
        L.t1_ranger_fan.argtypes = [vp, u32, ctypes.c_int, vp, f64, f64, u32, vp, vp, vp, vp]
        L.t1_even_fan.argtypes = [vp, u32, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, f64, f64, u32, vp]
        L.t1_dump_grid.restype = ctypes.c_int64
        L.t1_dump_grid.argtypes = [vp, vp, ctypes.c_int64, vp, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
        L.t1_replay.restype, L.t1_replay.argtypes = vp, [ctypes.c_char_p, ctypes.POINTER(ReplayStats)]
        L.t1_blocks_remap.restype, L.t1_blocks_remap.argtypes = ctypes.c_int, [vp, u32, vp, vp, ctypes.c_int, vp, vp, vp]
        L.t1_grid_hash.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64 * 5)]
        L.t1_test_collision.restype, L.t1_test_collision.argtypes = ctypes.c_int32, [vp, ctypes.c_int, u32, vp]
        L.t1_touching_models.restype, L.t1_touching_models.argtypes = u32, [vp, ctypes.c_int, u32, vp, u32, vp]
        _t1 = L
    return _t1


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def libm_trig(headings):
    """(n,2) array of (cos, sin) from the host libm, with World::Raytrace's zero-heading rule."""
    headings = np.ascontiguousarray(headings, np.float64)
    out = np.zeros((len(headings), 2), np.float64)
    t1().t1_libm_trig(len(headings), _p(headings), _p(out))
    return out


def ranger_headings(oa, fov, n):
    """Per-beam headings of ModelRanger::Sensor::Update (model_ranger.cc:232-255): the running angle
    goes through a float every beam."""
    incr = fov / max(n - 1, 1)
    out = np.zeros(n, np.float64)
    a = float(oa)
    for t in range(n):
        f = float(np.float32(a))
        out[t] = f
        a = f + incr
    return out


def sort_grid(rec):
    """canonical order of {layer,x,y,pos,block} records"""
    if len(rec) == 0:
        return rec
    idx = np.lexsort((rec[:, 3], rec[:, 1], rec[:, 2], rec[:, 0]))
    return rec[idx]


def first_occurrences(rec):
    """{layer,x,y,pos,block} records with the later copies of a block in a cell dropped and the positions renumbered:
    what decides which block a ray meets first when a block was rendered more than once (Map on a mapped block)"""
    rec = sort_grid(rec)
    out, seen, cell, pos = [], set(), None, 0
    for r in rec:
        key = (r[0], r[1], r[2])
        if key != cell:
            cell, seen, pos = key, set(), 0
        if r[4] in seen:
            continue
        seen.add(r[4])
        out.append((r[0], r[1], r[2], pos, r[4]))
        pos += 1
    return np.array(out, np.int32).reshape(-1, 5)


def entry_multiset(rec):
    """the (layer, x, y, block) entries, duplicates included, in canonical order"""
    e = rec[:, [0, 1, 2, 4]]
    return e[np.lexsort((e[:, 3], e[:, 1], e[:, 2], e[:, 0]))]


def sort_counts(cnt):
    if len(cnt) == 0:
        return cnt
    return cnt[np.lexsort((cnt[:, 0], cnt[:, 1]))]


class T1World:
    """The restated CPU world (oracle/t1_oracle.cc)."""

    def __init__(self, ppm, handle=None):
        self.L = t1()
        self.h = handle if handle is not None else self.L.t1_create(ppm)

    def model_upsert(self, mid, root, ranger_return=1.0, fiducial_return=0, fiducial_key=0, flags=2):
        self.L.t1_model_upsert(self.h, mid, root, ranger_return, fiducial_return, fiducial_key, flags)

    def block_map(self, block, model, layer, zmin, zmax, xy):
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
        assert self.L.t1_block_map(self.h, block, model, layer, zmin, zmax, len(xy), _p(xy)) == 0

    def block_unmap(self, block, layer):
        self.L.t1_block_unmap(self.h, block, layer)

    def raytrace(self, rays, trig=None):
        rays = np.ascontiguousarray(rays, T1_RAY)
        out = np.zeros(len(rays), T1_HIT)
        trig = None if trig is None else np.ascontiguousarray(trig, np.float64)
        self.L.t1_raytrace(self.h, len(rays), _p(rays), _p(trig), _p(out))
        return out

    def ranger_fan(self, finder, layer, origin, range_max, fov, n):
        origin = np.ascontiguousarray(origin, np.float64)
        rg, it, be = np.zeros(n), np.zeros(n), np.zeros(n)
        hits = np.zeros(n, T1_HIT)
        self.L.t1_ranger_fan(self.h, finder, layer, _p(origin), range_max, fov, n, _p(rg), _p(it), _p(be), _p(hits))
        return rg, it, be, hits

    def even_fan(self, finder, layer, kind, ztest, gpose, rng, fov, n):
        gpose = np.ascontiguousarray(gpose, np.float64)
        hits = np.zeros(n, T1_HIT)
        self.L.t1_even_fan(self.h, finder, layer, kind, ztest, _p(gpose), rng, fov, n, _p(hits))
        return hits

    def test_collision(self, layer, blocks):
        """Model::TestCollision over these blocks (own, then descendants', depth first) -> model id or -1"""
        blocks = np.ascontiguousarray(blocks, np.uint32)
        return int(self.L.t1_test_collision(self.h, layer, len(blocks), _p(blocks)))

    def touching_models(self, layer, blocks, cap=256):
        """Model::AppendTouchingModels over the model's own blocks -> ascending ids of the unrelated models touched"""
        blocks = np.ascontiguousarray(blocks, np.uint32)
        out = np.zeros(cap, np.uint32)
        n = int(self.L.t1_touching_models(self.h, layer, len(blocks), _p(blocks), cap, _p(out)))
        assert n <= cap
        return out[:n]

    def blocks_remap(self, blocks, models, layer, z, first_pt, xy):
        """n x (Block::UnMap + Block::Map): the oracle's side of stg_rt_blocks_remap"""
        blocks = np.ascontiguousarray(blocks, np.uint32)
        models = np.ascontiguousarray(models, np.uint32)
        z = np.ascontiguousarray(z, np.float64)
        first_pt = np.ascontiguousarray(first_pt, np.uint32)
        xy = np.ascontiguousarray(xy, np.int32)
        assert self.L.t1_blocks_remap(self.h, len(blocks), _p(blocks), _p(models), layer, _p(z), _p(first_pt), _p(xy)) == 0

    def grid_hash(self):
        """(hash[2], entries[2], region_count_hash): the digest stg_rt_grid_hash computes, from the oracle's own lists"""
        out = (ctypes.c_uint64 * 5)()
        self.L.t1_grid_hash(ctypes.c_void_p(self.h), out)
        return [int(out[0]), int(out[1])], [int(out[2]), int(out[3])], int(out[4])

    def dump_grid(self):
        nc = ctypes.c_int64(0)
        n = self.L.t1_dump_grid(self.h, None, 0, None, 0, ctypes.byref(nc))
        rec = np.zeros((max(n, 1), 5), np.int32)
        cnt = np.zeros((max(nc.value, 1), 3), np.int64)
        self.L.t1_dump_grid(self.h, _p(rec), n, _p(cnt), nc.value, ctypes.byref(nc))
        return sort_grid(rec[:n]), sort_counts(cnt[:nc.value])


def t1_replay(path):
    """Replay a raw (uncompressed) trace file through T1; returns (stats dict, T1World)."""
    st = ReplayStats()
    h = t1().t1_replay(str(path).encode(), ctypes.byref(st))
    if not h:
        raise RuntimeError("t1_replay failed on %s" % path)
    return st.as_dict(), T1World(0.0, handle=h)


# ---------------------------------------------------------------------------------------------
# T0: the unmodified reference
# ---------------------------------------------------------------------------------------------

def ref_available():
    return os.path.exists(os.path.join(REF, "libref_harness.so")) and os.path.exists(os.path.join(REF, "libstage_ref.so"))


_ref = None


def ref():
    """Load the harness (which pulls in libstage_ref.so AFTER itself, so its interposers bind)."""
    global _ref
    if _ref is None:
        if not ref_available():
            raise RuntimeError("oracle/_ref is not built (run oracle/build_ref.sh where /root/reference exists)")
        os.environ["STAGEPATH"] = "%s/plugins:%s/share/stage" % (REF, REF)
        L = ctypes.CDLL(os.path.join(REF, "libref_harness.so"))
        vp = ctypes.c_void_p
        L.ref_world_load_file.restype, L.ref_world_load_file.argtypes = vp, [ctypes.c_char_p]
        L.ref_world_load_text.restype, L.ref_world_load_text.argtypes = vp, [ctypes.c_char_p, ctypes.c_char_p]
        L.ref_world_update.argtypes = [vp, ctypes.c_int]
        L.ref_world_ppm.restype, L.ref_world_ppm.argtypes = ctypes.c_double, [vp]
        L.ref_world_updates.restype, L.ref_world_updates.argtypes = ctypes.c_uint64, [vp]
        L.ref_trace_end.argtypes = [vp, ctypes.c_char_p]
        L.ref_raytrace_batch.restype = ctypes.c_double
        L.ref_raytrace_batch.argtypes = [vp, ctypes.c_uint64, vp, vp, ctypes.c_int]
        L.ref_dump_grid.restype = ctypes.c_int64
        L.ref_dump_grid.argtypes = [vp, vp, ctypes.c_int64, vp, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
        L.ref_model_id.restype, L.ref_model_id.argtypes = ctypes.c_int32, [vp, ctypes.c_char_p]
        L.ref_model_pose.argtypes = [vp, ctypes.c_uint32, vp]
        L.ref_ranger_scan.argtypes = [vp, ctypes.c_uint32, ctypes.c_int, vp, vp, vp, ctypes.c_int]
        L.ref_model_subscribe.argtypes = [vp, ctypes.c_char_p]
        L.ref_fiducials_on_main_thread.argtypes = [ctypes.c_int]
        L.ref_model_set_speed.argtypes = [vp, ctypes.c_char_p, ctypes.c_double, ctypes.c_double, ctypes.c_double]
        L.ref_models_remap.restype = ctypes.c_double
        L.ref_models_remap.argtypes = [vp, ctypes.c_uint32, vp, vp, ctypes.c_int]
        L.ref_probe_models.restype, L.ref_probe_models.argtypes = ctypes.c_int, [vp, ctypes.c_uint32, vp]
        L.ref_block_geometry.restype = ctypes.c_int64
        L.ref_block_geometry.argtypes = [vp, vp, vp, ctypes.c_int64, vp, vp, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
        L.ref_models_remap2.restype = ctypes.c_double
        L.ref_models_remap2.argtypes = [vp, ctypes.c_uint32, vp, vp, ctypes.c_int, ctypes.c_int]
        _ref = L
    return _ref


class RefWorld:
    """A live World of the unmodified reference. Never destroyed (the reference's ~World crashes)."""

    def __init__(self, path=None, text=None, record=True, fiducials_on_main_thread=True, path_hint="synthetic.world"):
        self.L = ref()
        # on worker threads fiducial updates race with Move(): not reproducible
        self.L.ref_fiducials_on_main_thread(1 if fiducials_on_main_thread else 0)
        if record:
            self.L.ref_trace_begin()
        if path is not None:
            self.h = self.L.ref_world_load_file(str(path).encode())
        else:
            # path_hint: where the text pretends to live (its includes and bitmaps are looked up next to it)
            self.h = self.L.ref_world_load_text(text.encode(), str(path_hint).encode())
        if not self.h:
            raise RuntimeError("reference failed to load the world")

    @property
    def ppm(self):
        return self.L.ref_world_ppm(self.h)

    def update(self, n=1):
        return self.L.ref_world_update(self.h, n)

    def trace_end(self, path):
        assert self.L.ref_trace_end(self.h, str(path).encode()) == 0

    def raytrace(self, rays, nthreads=1):
        """returns (hits, seconds spent inside World::Raytrace loops)"""
        rays = np.ascontiguousarray(rays, REF_RAY)
        out = np.zeros(len(rays), REF_HIT)
        secs = self.L.ref_raytrace_batch(self.h, len(rays), _p(rays), _p(out), nthreads)
        return out, secs

    def dump_grid(self):
        nc = ctypes.c_int64(0)
        n = self.L.ref_dump_grid(self.h, None, 0, None, 0, ctypes.byref(nc))
        rec = np.zeros((max(n, 1), 5), np.int32)
        cnt = np.zeros((max(nc.value, 1), 3), np.int64)
        self.L.ref_dump_grid(self.h, _p(rec), n, _p(cnt), nc.value, ctypes.byref(nc))
        return sort_grid(rec[:n]), sort_counts(cnt[:nc.value])

    def block_geometry(self):
        """every block as the reference holds it now: dict(model[nb], first[nb + 1], gpose[nb, 4], local_z[nb, 2], pts[np, 2] in
        model coordinates, pix[np, 2] = Model::LocalToPixels of them) - models by id, blocks in BlockGroup order"""
        npts = ctypes.c_int64(0)
        nb = self.L.ref_block_geometry(self.h, None, None, 0, None, None, 0, ctypes.byref(npts))
        hdr, num = np.zeros((max(nb, 1), 4), np.int64), np.zeros((max(nb, 1), 6), np.float64)
        pts, pix = np.zeros((max(npts.value, 1), 2), np.float64), np.zeros((max(npts.value, 1), 2), np.int32)
        assert self.L.ref_block_geometry(self.h, _p(hdr), _p(num), nb, _p(pts), _p(pix), npts.value, ctypes.byref(npts)) == nb
        first = np.concatenate([hdr[:nb, 2], [npts.value]]).astype(np.int64)
        return {"model": hdr[:nb, 0].astype(np.uint32), "first": first, "gpose": num[:nb, :4].copy(), "local_z": num[:nb, 4:6].copy(),
                "pts": pts[:npts.value], "pix": pix[:npts.value]}

    def subscribe(self, name):
        assert self.L.ref_model_subscribe(self.h, name.encode()) == 0, name

    def models_remap(self, ids, xyza, layer, record=False):
        """what Move() does to the grid for these models, one after the other; returns the seconds it took.
        record: the Map / UnMap calls are logged into the trace (off for the timed benchmark ticks)"""
        ids = np.ascontiguousarray(ids, np.uint32)
        xyza = np.ascontiguousarray(xyza, np.float64).reshape(len(ids), 4)
        return self.L.ref_models_remap2(self.h, len(ids), _p(ids), _p(xyza), layer, 1 if record else 0)

    def probe_models(self, ids):
        """Model::TestCollision and Model::AppendTouchingModels of these models, recorded into the trace; -> collisions found"""
        ids = np.ascontiguousarray(ids, np.uint32)
        return int(self.L.ref_probe_models(self.h, len(ids), _p(ids)))

    def set_speed(self, name, x, y, a):
        assert self.L.ref_model_set_speed(self.h, name.encode(), x, y, a) == 0, name

    def model_id(self, name):
        return self.L.ref_model_id(self.h, name.encode())

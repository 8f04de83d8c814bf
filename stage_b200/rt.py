"""ctypes binding of libstg_rt.so (include/stg_rt.h): the same entry points a patched libstage
calls, with numpy arrays for the bulk arguments. Nothing here computes: every call goes to the
CUDA library, and importing fails loudly if that library cannot be built or loaded.
"""
import ctypes
import os

import numpy as np

from . import build as _build

PRED_RANGER, PRED_UNRELATED, PRED_GRIPPER = 0, 1, 2
ANGLES_RANGER, ANGLES_EVEN_FOV, ANGLES_EXPLICIT = 0, 1, 2
VIS_GRIPPER_RETURN, VIS_OBSTACLE_RETURN, VIS_BLOB_RETURN = 1, 2, 4
CFG_HOST_EXACT_FIDUCIALS = 1
WANT_RANGES, WANT_INTENSITIES, WANT_BEARINGS, WANT_HIT_MODEL, WANT_HIT_CELL, WANT_STEPS = 1, 2, 4, 8, 16, 32
WANT_ALL = 63

# stg_rt_fan, 64 bytes
FAN_DTYPE = np.dtype([
    ("finder", "<u4"), ("n_samples", "<u4"), ("pred", "u1"), ("ztest", "u1"), ("layer", "u1"),
    ("angles", "u1"), ("first_heading", "<u4"), ("ox", "<f8"), ("oy", "<f8"), ("oz", "<f8"),
    ("oa", "<f8"), ("range", "<f8"), ("fov", "<f8")])
assert FAN_DTYPE.itemsize == 64

# stg_rt_fid_target (72 B), stg_rt_fid_finder (104 B), stg_rt_fiducial (88 B)
FID_TARGET_DTYPE = np.dtype([("model", "<u4"), ("fiducial_return", "<i4"), ("fiducial_key", "<i4"), ("reserved", "<u4"),
                             ("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("a", "<f8"), ("sx", "<f8"), ("sy", "<f8"), ("sz", "<f8")])
FID_FINDER_DTYPE = np.dtype([("finder", "<u4"), ("fiducial_key", "<i4"), ("ignore_zloc", "u1"), ("layer", "u1"), ("pad", "u1", (6,)),
                             ("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("a", "<f8"), ("ox", "<f8"), ("oy", "<f8"), ("oz", "<f8"),
                             ("oa", "<f8"), ("max_range_anon", "<f8"), ("max_range_id", "<f8"), ("fov", "<f8")])
FIDUCIAL_DTYPE = np.dtype([("target", "<u4"), ("id", "<i4"), ("range", "<f8"), ("bearing", "<f8"), ("geom", "<f8", (4,)),
                           ("pose", "<f8", (4,))])
# stg_rt_blob_finder (72 B), stg_rt_blob (32 B)
BLOB_FINDER_DTYPE = np.dtype([("finder", "<u4"), ("scan_width", "<u4"), ("scan_height", "<u4"), ("layer", "u1"), ("pad", "u1", (3,)),
                              ("first_color", "<u4"), ("n_colors", "<u4"), ("ox", "<f8"), ("oy", "<f8"), ("oz", "<f8"),
                              ("oa", "<f8"), ("range", "<f8"), ("fov", "<f8")])
RANGER_SENSOR_DTYPE = np.dtype([("px", "<f8"), ("py", "<f8"), ("pz", "<f8"), ("pa", "<f8"), ("size_z", "<f8"), ("range_max", "<f8"),
                                ("fov", "<f8"), ("n_samples", "<u4"), ("reserved", "<u4")])
RANGER_POSE_DTYPE = np.dtype([("finder", "<u4"), ("layer", "u1"), ("pad", "u1", (3,)), ("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("a", "<f8")])
assert RANGER_SENSOR_DTYPE.itemsize == 64 and RANGER_POSE_DTYPE.itemsize == 40
FAN_NOISE_DTYPE = np.dtype([("range_noise_const", "<f8"), ("range_noise", "<f8"), ("angle_noise", "<f8"), ("seed", "<u8")])
BLOB_DTYPE = np.dtype([("model", "<u4"), ("left", "<u4"), ("top", "<u4"), ("right", "<u4"), ("bottom", "<u4"), ("pad", "<u4"),
                       ("range", "<f8")])
assert (FID_TARGET_DTYPE.itemsize, FID_FINDER_DTYPE.itemsize, FIDUCIAL_DTYPE.itemsize) == (72, 104, 88)
assert (BLOB_FINDER_DTYPE.itemsize, BLOB_DTYPE.itemsize) == (72, 32)


class Config(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("device", ctypes.c_int32), ("ppm", ctypes.c_double),
                ("sreg_x0", ctypes.c_int32), ("sreg_y0", ctypes.c_int32), ("sreg_nx", ctypes.c_int32),
                ("sreg_ny", ctypes.c_int32), ("max_models", ctypes.c_uint32), ("max_blocks", ctypes.c_uint32),
                ("max_cell_entries", ctypes.c_uint32), ("flags", ctypes.c_uint32),
                ("max_delta_bytes", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class FanOut(ctypes.Structure):
    _fields_ = [("ranges", ctypes.c_void_p), ("intensities", ctypes.c_void_p), ("bearings", ctypes.c_void_p),
                ("hit_model", ctypes.c_void_p), ("hit_cell", ctypes.c_void_p), ("steps", ctypes.c_void_p)]


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in (
        "launches_rasterise", "launches_march", "launches_post", "launches_other", "beams", "fans",
        "map_calls", "unmap_calls", "edge_steps", "h2d_bytes", "d2h_bytes", "extent_grows")]


class StgRtError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("stg_rt error %d: %s" % (code, text))
        self.code = code


_lib = None


def lib():
    """Load (building first if stale and nvcc is here) the CUDA library. No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    # STG_RT_LIBRARY: load this build of the library instead (kernel experiments; same ABI, checked below)
    path = os.environ.get("STG_RT_LIBRARY") or _build.build()
    L = ctypes.CDLL(path)
    vp, u32, i32, f64, u64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int32, ctypes.c_double, ctypes.c_uint64
    sz = ctypes.c_size_t
    sigs = {
        "stg_rt_abi_version": (ctypes.c_int, []),
        "stg_rt_create": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.POINTER(vp)]),
        "stg_rt_destroy": (ctypes.c_int, [vp]),
        "stg_rt_last_error": (ctypes.c_char_p, [vp]),
        "stg_rt_model_upsert": (ctypes.c_int, [vp, u32, u32, f64, i32, i32, u32, vp]),
        "stg_rt_block_map": (ctypes.c_int, [vp, u32, u32, ctypes.c_int, f64, f64, ctypes.c_int, vp]),
        "stg_rt_block_unmap": (ctypes.c_int, [vp, u32, ctypes.c_int]),
        "stg_rt_blocks_remap": (ctypes.c_int, [vp, u32, vp, vp, ctypes.c_int, vp, vp, vp]),
        "stg_rt_block_define": (ctypes.c_int, [vp, u32, u32, f64, f64, ctypes.c_int, vp]),
        "stg_rt_models_pose_update": (ctypes.c_int, [vp, u32, vp, vp, ctypes.c_int]),
        "stg_rt_debug_block_outline": (ctypes.c_int, [vp, u32, ctypes.c_int, vp, ctypes.c_int]),
        "stg_rt_fans_submit": (ctypes.c_int, [vp, u32, vp, vp, vp, ctypes.POINTER(FanOut)]),
        "stg_rt_fiducials_submit": (ctypes.c_int, [vp, u32, vp, u32, vp, u32, vp, vp]),
        "stg_rt_fiducials_submit_packed": (ctypes.c_int, [vp, u32, vp, u32, vp, u32, u64, vp, vp]),
        "stg_rt_fiducials_submit_device_targets": (ctypes.c_int, [vp, u32, vp, u32, vp, u32, u64, vp, vp]),
        "stg_rt_blobs_submit": (ctypes.c_int, [vp, u32, vp, u32, vp, u32, vp, vp]),
        "stg_rt_rangers_submit": (ctypes.c_int, [vp, u32, vp, u32, vp, ctypes.POINTER(FanOut)]),
        "stg_rt_collisions_test": (ctypes.c_int, [vp, ctypes.c_int, u32, vp, vp, u32, vp]),
        "stg_rt_touching_models": (ctypes.c_int, [vp, ctypes.c_int, u32, vp, vp, u32, vp, vp]),
        "stg_rt_models_move": (ctypes.c_int, [vp, ctypes.c_int, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "stg_rt_fans_submit_noisy": (ctypes.c_int, [vp, u32, vp, vp, ctypes.POINTER(FanOut)]),
        "stg_rt_flush": (ctypes.c_int, [vp]),
        "stg_rt_sync": (ctypes.c_int, [vp]),
        "stg_rt_host_alloc": (vp, [vp, sz]),
        "stg_rt_host_free": (ctypes.c_int, [vp, vp]),
        "stg_rt_set_create": (ctypes.c_int, [vp, u32, vp, vp, vp, u32, ctypes.POINTER(vp)]),
        "stg_rt_set_destroy": (ctypes.c_int, [vp, vp]),
        "stg_rt_set_update_origins": (ctypes.c_int, [vp, vp, vp, ctypes.c_int]),
        "stg_rt_set_trace": (ctypes.c_int, [vp, vp]),
        "stg_rt_set_download": (ctypes.c_int, [vp, vp, ctypes.POINTER(FanOut)]),
        "stg_rt_set_device_ptrs": (ctypes.c_int, [vp, vp, ctypes.POINTER(FanOut)]),
        "stg_rt_set_beams": (u64, [vp]),
        "stg_rt_grid_dump": (ctypes.c_int64, [vp, vp, ctypes.c_int64, vp, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]),
        "stg_rt_debug_region_owners": (ctypes.c_int64, [vp, vp, ctypes.c_int64]),
        "stg_rt_debug_occupancy_check": (ctypes.c_int, [vp, ctypes.POINTER(u64 * 3)]),
        "stg_rt_grid_hash": (ctypes.c_int, [vp, ctypes.POINTER(u64 * 5)]),
        "stg_rt_delta_export": (ctypes.c_int, [vp, ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(sz)]),
        "stg_rt_delta_slot_bytes": (sz, [vp]),
        "stg_rt_delta_apply_gathered": (ctypes.c_int, [vp, vp, ctypes.c_int, sz]),
        "stg_rt_stream_signal": (ctypes.c_int, [vp, vp]),
        "stg_rt_stream_wait": (ctypes.c_int, [vp, vp]),
        "stg_rt_ranger_rand_advance": (u32, [u32, u32]),
        "stg_rt_set_stream": (ctypes.c_int, [vp, vp]),
        "stg_rt_get_stream": (vp, [vp]),
        "stg_rt_get_stats": (ctypes.c_int, [vp, ctypes.POINTER(Stats)]),
        "stg_rt_kernel_times": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_float * 3)]),
        "stg_rt_kernel_times_ex": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_float * 6), ctypes.c_int]),
        "stg_rt_debug_link_bandwidth": (ctypes.c_int, [vp, ctypes.c_int, sz, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
        "stg_rt_debug_trig": (ctypes.c_int, [vp, u64, vp, vp]),
        "stg_rt_debug_headings": (ctypes.c_int, [vp, u32, vp, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)  # AttributeError if the library does not export what the header declares
        fn.restype, fn.argtypes = res, args
    if L.stg_rt_abi_version() != 1:
        raise RuntimeError("libstg_rt.so ABI version mismatch")
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "stg_rt_abi_version stg_rt_create stg_rt_destroy stg_rt_last_error stg_rt_model_upsert stg_rt_block_map "
    "stg_rt_block_unmap stg_rt_blocks_remap stg_rt_block_define stg_rt_models_pose_update stg_rt_debug_block_outline stg_rt_fiducials_submit stg_rt_fiducials_submit_packed stg_rt_fiducials_submit_device_targets stg_rt_rangers_submit stg_rt_collisions_test stg_rt_touching_models stg_rt_models_move stg_rt_blobs_submit stg_rt_fans_submit_noisy stg_rt_fans_submit stg_rt_flush stg_rt_sync stg_rt_host_alloc stg_rt_host_free "
    "stg_rt_set_create stg_rt_set_destroy stg_rt_set_update_origins stg_rt_set_trace stg_rt_set_download "
    "stg_rt_set_device_ptrs stg_rt_set_beams stg_rt_grid_dump stg_rt_delta_export stg_rt_delta_slot_bytes "
    "stg_rt_delta_apply_gathered stg_rt_stream_signal stg_rt_stream_wait stg_rt_ranger_rand_advance stg_rt_set_stream stg_rt_get_stream stg_rt_get_stats stg_rt_kernel_times stg_rt_kernel_times_ex stg_rt_debug_link_bandwidth stg_rt_debug_trig stg_rt_debug_headings stg_rt_debug_region_owners stg_rt_debug_occupancy_check stg_rt_grid_hash").split()


_BYTES0 = ctypes.c_char * 0


def _ptr(a):
    """address of a contiguous numpy array (as an int: every pointer argument is declared c_void_p). Going through
    the buffer protocol costs a tenth of ndarray.ctypes.data_as, which matters for calls made every tick."""
    if a is None:
        return None
    try:
        return ctypes.addressof(_BYTES0.from_buffer(a))
    except (TypeError, ValueError, BufferError):  # read-only buffer
        return a.ctypes.data


def make_fans(n):
    return np.zeros(n, FAN_DTYPE)


class Outputs:
    """numpy result arrays for a batch of fans, plus the FanOut struct pointing at them."""

    def __init__(self, n_beams, want=WANT_ALL, alloc=None):
        def mk(flag, dtype, shape):
            if not (want & flag):
                return None
            if alloc is None:
                return np.empty(shape, dtype)
            return alloc(shape, dtype)
        self.ranges = mk(WANT_RANGES, np.float64, (n_beams,))
        self.intensities = mk(WANT_INTENSITIES, np.float64, (n_beams,))
        self.bearings = mk(WANT_BEARINGS, np.float64, (n_beams,))
        self.hit_model = mk(WANT_HIT_MODEL, np.int32, (n_beams,))
        self.hit_cell = mk(WANT_HIT_CELL, np.int32, (n_beams, 2))
        self.steps = mk(WANT_STEPS, np.uint32, (n_beams, 2))
        self.struct = FanOut(*[_ptr(a) for a in (self.ranges, self.intensities, self.bearings,
                                                 self.hit_model, self.hit_cell, self.steps)])


class Context:
    """One stg_rt_ctx: the replicated world state of one GPU plus its queues."""

    def __init__(self, ppm, sreg_x0, sreg_y0, sreg_nx, sreg_ny, max_models=1 << 16, max_blocks=1 << 16,
                 max_cell_entries=0, device=0, max_delta_bytes=0, flags=0):
        self._L = lib()
        cfg = Config(ctypes.sizeof(Config), device, ppm, sreg_x0, sreg_y0, sreg_nx, sreg_ny, max_models,
                     max_blocks, max_cell_entries, flags, max_delta_bytes, 0)
        h = ctypes.c_void_p()
        rc = self._L.stg_rt_create(ctypes.byref(cfg), ctypes.byref(h))
        if rc:
            raise StgRtError(rc, (self._L.stg_rt_last_error(None) or b"").decode())
        self._h = h
        self._keep = []  # arrays the library writes into asynchronously
        self._pinned = []

    def close(self):
        if self._h:
            self._L.stg_rt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise StgRtError(rc, (self._L.stg_rt_last_error(self._h) or b"").decode())
        return rc

    # ---- world state
    def model_upsert(self, mid, root, ranger_return=1.0, fiducial_return=0, fiducial_key=0, vis_flags=VIS_OBSTACLE_RETURN,
                     rgba=None):
        c = None if rgba is None else np.ascontiguousarray(rgba, np.float64)
        self._check(self._L.stg_rt_model_upsert(self._h, mid, root, ranger_return, fiducial_return, fiducial_key,
                                                vis_flags, _ptr(c)))

    def block_map(self, block, model, layer, zmin, zmax, xy):
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
        self._check(self._L.stg_rt_block_map(self._h, block, model, layer, zmin, zmax, len(xy), _ptr(xy)))

    def block_unmap(self, block, layer):
        self._check(self._L.stg_rt_block_unmap(self._h, block, layer))

    def blocks_remap(self, blocks, models, layer, z, first_pt, xy):
        blocks = np.ascontiguousarray(blocks, np.uint32)
        models = np.ascontiguousarray(models, np.uint32)
        z = np.ascontiguousarray(z, np.float64)
        first_pt = np.ascontiguousarray(first_pt, np.uint32)
        xy = np.ascontiguousarray(xy, np.int32)
        assert len(first_pt) == len(blocks) + 1 and len(models) == len(blocks) and z.size == 2 * len(blocks)
        self._check(self._L.stg_rt_blocks_remap(self._h, len(blocks), _ptr(blocks), _ptr(models), layer, _ptr(z),
                                                _ptr(first_pt), _ptr(xy)))

    def block_define(self, block, model, local_zmin, local_zmax, pts):
        """register a block's polygon in model coordinates (Block::pts after CalcSize) and its local z range"""
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 2)
        self._check(self._L.stg_rt_block_define(self._h, block, model, local_zmin, local_zmax, len(pts), _ptr(pts)))

    def models_pose_update(self, models, gpose_xyza, layer):
        """UnMap + Map of every defined block of these models at gpose = GetGlobalPose() + geom.pose, outlines derived on the device"""
        models = np.ascontiguousarray(models, np.uint32)
        gp = np.ascontiguousarray(gpose_xyza, np.float64).reshape(len(models), 4)
        self._check(self._L.stg_rt_models_pose_update(self._h, len(models), _ptr(models), _ptr(gp), layer))

    def block_outline(self, block, layer):
        """(n, 2) int32 pixel outline the block is rendered with on that layer now (empty: not mapped)"""
        n = self._check(self._L.stg_rt_debug_block_outline(self._h, block, layer, None, 0))
        xy = np.zeros((max(n, 1), 2), np.int32)
        if n:
            self._check(self._L.stg_rt_debug_block_outline(self._h, block, layer, _ptr(xy), n))
        return xy[:n]

    # ---- rays
    def fans_submit(self, fans, out, headings=None, trig=None):
        fans = np.ascontiguousarray(fans, FAN_DTYPE)
        headings = None if headings is None else np.ascontiguousarray(headings, np.float64)
        trig = None if trig is None else np.ascontiguousarray(trig, np.float64)
        self._keep.append(out)
        self._check(self._L.stg_rt_fans_submit(self._h, len(fans), _ptr(fans), _ptr(headings), _ptr(trig),
                                               ctypes.byref(out.struct)))

    def rangers_submit(self, sensors, poses, out):
        """whole ranger models: sensors = RANGER_SENSOR_DTYPE array shared by all, poses = RANGER_POSE_DTYPE array"""
        sensors = np.ascontiguousarray(sensors, RANGER_SENSOR_DTYPE)
        poses = np.ascontiguousarray(poses, RANGER_POSE_DTYPE)
        self._keep.append(out)
        self._check(self._L.stg_rt_rangers_submit(self._h, len(sensors), _ptr(sensors), len(poses), _ptr(poses), ctypes.byref(out.struct)))

    def fiducials_submit(self, targets, finders, max_per_finder=16, into=None):
        """-> (out[n_finders, max_per_finder] of FIDUCIAL_DTYPE, count[n_finders]); valid after sync().
        into = (out, count) arrays of those shapes to reuse (only out[f, :count[f]] is written)."""
        targets = np.ascontiguousarray(targets, FID_TARGET_DTYPE)
        finders = np.ascontiguousarray(finders, FID_FINDER_DTYPE)
        if into is None:
            out = np.zeros((len(finders), max_per_finder), FIDUCIAL_DTYPE)
            cnt = np.zeros(len(finders), np.uint32)
        else:
            out, cnt = into
            assert out.shape == (len(finders), max_per_finder) and out.dtype == FIDUCIAL_DTYPE and out.flags.c_contiguous
            assert cnt.shape == (len(finders),) and cnt.dtype == np.uint32
        self._keep.append((out, cnt))
        self._check(self._L.stg_rt_fiducials_submit(self._h, len(targets), _ptr(targets), len(finders), _ptr(finders),
                                                    max_per_finder, _ptr(out), _ptr(cnt)))
        return out, cnt

    def collisions_test(self, layer, queries, ground_model=0):
        """queries: one sequence of block ids per model (own blocks, then descendants', depth first)
        -> int32 array of hit model ids (-1: none); valid after sync()."""
        if isinstance(queries, tuple):  # already flattened: (query_first[n + 1], blocks[])
            first, blocks = (np.ascontiguousarray(a, np.uint32) for a in queries)
        else:
            first = np.zeros(len(queries) + 1, np.uint32)
            first[1:] = np.cumsum([len(q) for q in queries])
            blocks = np.ascontiguousarray(np.concatenate([np.asarray(q, np.uint32) for q in queries]) if len(queries) else np.zeros(0), np.uint32)
        hit = np.full(len(first) - 1, -2, np.int32)
        self._keep.append((first, blocks, hit))
        self._check(self._L.stg_rt_collisions_test(self._h, layer, len(first) - 1, _ptr(first), _ptr(blocks), ground_model, _ptr(hit)))
        return hit

    def touching_models(self, layer, queries, max_per_query=32):
        """queries: one sequence of block ids per model (its own blocks) -> (ids[n, max_per_query] uint32 ascending,
        counts[n] uint32); valid after sync()."""
        first = np.zeros(len(queries) + 1, np.uint32)
        first[1:] = np.cumsum([len(q) for q in queries])
        blocks = np.ascontiguousarray(np.concatenate([np.asarray(q, np.uint32) for q in queries]) if len(queries) else np.zeros(0), np.uint32)
        ids = np.zeros((len(queries), max_per_query), np.uint32)
        counts = np.zeros(len(queries), np.uint32)
        self._keep.append((first, blocks, ids, counts))
        self._check(self._L.stg_rt_touching_models(self._h, layer, len(queries), _ptr(first), _ptr(blocks), max_per_query, _ptr(ids), _ptr(counts)))
        return ids, counts

    def models_move(self, layer, movers):
        """movers: per mover a list of (block id, (zmin, zmax, polygon) wanted, (zmin, zmax, polygon) present) - own
        blocks, then descendants', depth first - in the order the reference moves them. Synchronous.
        -> uint8 array: 1 = stalled (re-rendered at the present pose)."""
        first = np.zeros(len(movers) + 1, np.uint32)
        first[1:] = np.cumsum([len(m) for m in movers])
        flat = [b for m in movers for b in m]
        blocks = np.array([b[0] for b in flat], np.uint32)

        def pack(which):
            z = np.ascontiguousarray([[b[which][0], b[which][1]] for b in flat], np.float64).reshape(-1, 2)
            polys = [np.ascontiguousarray(b[which][2], np.int32).reshape(-1, 2) for b in flat]
            fp = np.zeros(len(flat) + 1, np.uint32)
            fp[1:] = np.cumsum([len(p) for p in polys])
            return z, fp, np.ascontiguousarray(np.concatenate(polys, 0) if polys else np.zeros((0, 2)), np.int32)
        z, fp, xy = pack(1)
        zs, fps, xys = pack(2)
        stalled = np.zeros(len(movers), np.uint8)
        self._check(self._L.stg_rt_models_move(self._h, layer, len(movers), _ptr(first), _ptr(blocks), _ptr(z), _ptr(fp), _ptr(xy),
                                               _ptr(zs), _ptr(fps), _ptr(xys), _ptr(stalled)))
        return stalled

    def fiducials_submit_packed(self, targets, finders, max_per_finder, capacity, into=None):
        """-> (records[capacity] packed finder after finder, count[n_finders]); valid after sync()."""
        targets = np.ascontiguousarray(targets, FID_TARGET_DTYPE)
        finders = np.ascontiguousarray(finders, FID_FINDER_DTYPE)
        out, cnt = into if into is not None else (np.zeros(capacity, FIDUCIAL_DTYPE), np.zeros(len(finders), np.uint32))
        self._keep.append((out, cnt))
        self._check(self._L.stg_rt_fiducials_submit_packed(self._h, len(targets), _ptr(targets), len(finders), _ptr(finders),
                                                           max_per_finder, capacity, _ptr(out), _ptr(cnt)))
        return out, cnt

    def fiducials_submit_device_targets(self, dev_ptr, n_targets, finders, max_per_finder, capacity, into=None):
        """fiducials_submit_packed with the target records already on the device (the all-gathered array)"""
        finders = np.ascontiguousarray(finders, FID_FINDER_DTYPE)
        out, cnt = into if into is not None else (np.zeros(capacity, FIDUCIAL_DTYPE), np.zeros(len(finders), np.uint32))
        self._keep.append((out, cnt))
        self._check(self._L.stg_rt_fiducials_submit_device_targets(self._h, n_targets, dev_ptr, len(finders), _ptr(finders), max_per_finder,
                                                                   capacity, _ptr(out), _ptr(cnt)))
        return out, cnt

    def blobs_submit(self, finders, colors_rgb=None, max_per_finder=None):
        """-> (out[n_finders, max_per_finder] of BLOB_DTYPE, count[n_finders]); valid after sync()."""
        finders = np.ascontiguousarray(finders, BLOB_FINDER_DTYPE)
        colors = np.zeros((0, 3)) if colors_rgb is None else np.ascontiguousarray(colors_rgb, np.float64).reshape(-1, 3)
        if max_per_finder is None:
            max_per_finder = int(finders["scan_width"].max())
        out = np.zeros((len(finders), max_per_finder), BLOB_DTYPE)
        cnt = np.zeros(len(finders), np.uint32)
        self._keep.append((out, cnt))
        self._check(self._L.stg_rt_blobs_submit(self._h, len(finders), _ptr(finders), len(colors),
                                                _ptr(colors) if len(colors) else None, max_per_finder, _ptr(out), _ptr(cnt)))
        return out, cnt

    def fans_submit_noisy(self, fans, noise, out):
        fans = np.ascontiguousarray(fans, FAN_DTYPE)
        noise = np.ascontiguousarray(noise, FAN_NOISE_DTYPE)
        assert len(noise) == len(fans)
        self._keep.append(out)
        self._check(self._L.stg_rt_fans_submit_noisy(self._h, len(fans), _ptr(fans), _ptr(noise), ctypes.byref(out.struct)))

    def flush(self):
        self._check(self._L.stg_rt_flush(self._h))

    def sync(self):
        self._check(self._L.stg_rt_sync(self._h))
        self._keep.clear()

    def trace(self, fans, want=WANT_ALL, headings=None, trig=None):
        """submit + sync convenience: returns the Outputs of this batch."""
        out = Outputs(int(np.sum(fans["n_samples"])), want)
        self.fans_submit(fans, out, headings, trig)
        self.sync()
        return out

    def pinned_empty(self, shape, dtype):
        """numpy array over page-locked memory the GPU writes results into directly."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = self._L.stg_rt_host_alloc(self._h, max(n, 1))
        if not p:
            raise StgRtError(-2, (self._L.stg_rt_last_error(self._h) or b"").decode())
        buf = (ctypes.c_char * max(n, 1)).from_address(p)
        arr = np.frombuffer(buf, dtype, int(np.prod(shape))).reshape(shape)
        self._pinned.append(p)
        return arr

    # ---- resident sets
    def set_create(self, fans, want=WANT_ALL, headings=None, trig=None):
        return FanSet(self, fans, want, headings, trig)

    # ---- inspection
    def grid_dump(self):
        """(records[n,5] int32 {layer,x,y,pos,block}, region_counts[m,3] int64 {rx,ry,count})."""
        nc = ctypes.c_int64(0)
        n = self._check(self._L.stg_rt_grid_dump(self._h, None, 0, None, 0, ctypes.byref(nc)))
        rec = np.zeros((max(n, 1), 5), np.int32)
        cnt = np.zeros((max(nc.value, 1), 3), np.int64)
        n2 = self._check(self._L.stg_rt_grid_dump(self._h, _ptr(rec), n, _ptr(cnt), nc.value, ctypes.byref(nc)))
        assert n2 == n
        return rec[:n], cnt[:nc.value]

    def grid_hash(self):
        """(hash[2], entries[2], region_count_hash): the digest stg_rt_grid_hash computes on the device"""
        out = (ctypes.c_uint64 * 5)()
        self._check(self._L.stg_rt_grid_hash(self._h, ctypes.byref(out)))
        return [int(out[0]), int(out[1])], [int(out[2]), int(out[3])], int(out[4])

    def region_owners(self):
        """int64[m,3] {rx, ry, tag} of the regions with an owner tag (root + 1, or 0xffffffff = several trees)."""
        n = self._check(self._L.stg_rt_debug_region_owners(self._h, None, 0))
        rec = np.zeros((max(n, 1), 3), np.int64)
        assert self._check(self._L.stg_rt_debug_region_owners(self._h, _ptr(rec), n)) == n
        return rec[:n]

    def occupancy_check(self):
        """(cells whose mask bits disagree with the cell entries, summary bits that disagree with the masks, occupied
        cells seen): the first two are 0 on a sound grid."""
        out = (ctypes.c_uint64 * 3)()
        self._check(self._L.stg_rt_debug_occupancy_check(self._h, ctypes.byref(out)))
        return int(out[0]), int(out[1]), int(out[2])

    def stats(self):
        s = Stats()
        self._check(self._L.stg_rt_get_stats(self._h, ctypes.byref(s)))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def kernel_times(self):
        """device time of the most recent launch of each kernel group, in ms (-1: never launched)"""
        ms = (ctypes.c_float * 6)()
        self._check(self._L.stg_rt_kernel_times_ex(self._h, ctypes.byref(ms), 6))
        return {"deltas_ms": ms[0], "headings_ms": ms[1], "march_ms": ms[2], "fiducials_ms": ms[3], "blobs_ms": ms[4],
                "collisions_ms": ms[5]}

    def link_bandwidth(self, mode, nbytes, reps=20):
        """GB/s into page-locked host memory: mode 0 = kernel stores (zero-copy), 1 = cudaMemcpyAsync D2H"""
        v = ctypes.c_double(0)
        self._check(self._L.stg_rt_debug_link_bandwidth(self._h, mode, nbytes, reps, ctypes.byref(v)))
        return v.value

    def debug_trig(self, headings):
        """(n, 2) array of the device's (cos, sin) for the given headings"""
        h = np.ascontiguousarray(headings, np.float64)
        out = np.empty((len(h), 2), np.float64)
        self._check(self._L.stg_rt_debug_trig(self._h, len(h), _ptr(h), _ptr(out)))
        return out

    def debug_headings(self, fans, headings=None):
        """the per-beam headings the device derives for these fans, concatenated in fan order"""
        fans = np.ascontiguousarray(fans, FAN_DTYPE)
        h = None if headings is None else np.ascontiguousarray(headings, np.float64)
        out = np.empty(int(fans["n_samples"].sum()), np.float64)
        self._check(self._L.stg_rt_debug_headings(self._h, len(fans), _ptr(fans), _ptr(h), _ptr(out)))
        return out

    # ---- multi-GPU deltas / plumbing
    def delta_slot_bytes(self):
        return self._L.stg_rt_delta_slot_bytes(self._h)

    def delta_export(self, rank):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        self._check(self._L.stg_rt_delta_export(self._h, rank, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def delta_apply_gathered(self, dev_ptr, n_ranks, slot_bytes):
        self._check(self._L.stg_rt_delta_apply_gathered(self._h, dev_ptr, n_ranks, slot_bytes))

    def stream_signal(self, consumer_stream):
        self._check(self._L.stg_rt_stream_signal(self._h, consumer_stream))

    def stream_wait(self, producer_stream):
        self._check(self._L.stg_rt_stream_wait(self._h, producer_stream))

    def set_stream(self, cuda_stream):
        self._check(self._L.stg_rt_set_stream(self._h, cuda_stream))

    def get_stream(self):
        return self._L.stg_rt_get_stream(self._h)


class FanSet:
    def __init__(self, ctx, fans, want, headings, trig):
        self.ctx = ctx
        fans = np.ascontiguousarray(fans, FAN_DTYPE)
        headings = None if headings is None else np.ascontiguousarray(headings, np.float64)
        trig = None if trig is None else np.ascontiguousarray(trig, np.float64)
        h = ctypes.c_void_p()
        ctx._check(ctx._L.stg_rt_set_create(ctx._h, len(fans), _ptr(fans), _ptr(headings), _ptr(trig), want,
                                            ctypes.byref(h)))
        self._h = h
        self.want = want
        self.n_fans = len(fans)
        self.n_beams = int(ctx._L.stg_rt_set_beams(h))

    def update_origins(self, xyza, layer):
        xyza = np.ascontiguousarray(xyza, np.float64).reshape(self.n_fans, 4)
        self.ctx._check(self.ctx._L.stg_rt_set_update_origins(self.ctx._h, self._h, _ptr(xyza), layer))

    def trace(self):
        self.ctx._check(self.ctx._L.stg_rt_set_trace(self.ctx._h, self._h))

    def download(self, want=None):
        out = Outputs(self.n_beams, self.want if want is None else want)
        self.ctx._check(self.ctx._L.stg_rt_set_download(self.ctx._h, self._h, ctypes.byref(out.struct)))
        return out

    def device_ptrs(self):
        d = FanOut()
        self.ctx._check(self.ctx._L.stg_rt_set_device_ptrs(self.ctx._h, self._h, ctypes.byref(d)))
        return d

    def destroy(self):
        if self._h:
            self.ctx._L.stg_rt_set_destroy(self.ctx._h, self._h)
            self._h = None

"""Python-side mirror of the C ABI in include/lmot.h (ctypes; used by tests/, bench.py and __graft_entry__).

The product is liblmot.so (CUDA kernels for sm_100a behind `extern "C"`).  This module only loads it and
marshals numpy arrays / raw device pointers; it contains NO implementation of any pipeline stage and there is
no CPU fallback: if the library is missing or no CUDA device is usable, construction raises.

Entry points mirror the reference's four free functions:
    Lmot.ground_remove(points)            <- groundRemove           ground_removal.cpp:177
    Lmot.component_cluster(elevated)      <- componentClustering    component_clustering.cpp:260
    Lmot.box_fit(elevated, grid, k)       <- boxFitting             box_fitting.cpp:422
    Lmot.track_step(boxes, ts, v, yaw)    <- getOriginPoints + immUkfJpdaf   imm_ukf_jpda.cpp:74,704
    Lmot.frame(points, ts, v, yaw)        <- all four, device resident between stages

The directory name is not a Python identifier; import it with
    importlib.import_module("3d-lidar-multi-object-tracking_b200")
"""
from __future__ import annotations

import ctypes as C
import os

# The frame pipeline uses up to 8 detection streams + 1 tracker stream per context; with the default of 8 hardware
# work queues streams alias and serialise each other.  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblmot.so")
TRACK_DUMP_DOUBLES = 236

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_STATE = 0, -1, -2, -3, -4
WARN_TRACK_TABLE_FULL = 1
RULE_INTENDED, RULE_GCC13_O2_COMPAT = 0, 1


class LmotError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"lmot status {status}: {msg}")
        self.status = status


class Params(C.Structure):
    _fields_ = [
        ("r_min", C.c_float), ("r_max", C.c_float), ("t_hmin", C.c_float), ("t_hmax", C.c_float),
        ("t_hdiff", C.c_float), ("h_sensor", C.c_float), ("ground_tolerance", C.c_double),
        ("roi_m", C.c_float),
        ("ram_points", C.c_int), ("l_slope_dist", C.c_int), ("l_num_points", C.c_int), ("sensor_height", C.c_float),
        ("t_height_min", C.c_float), ("t_height_max", C.c_float), ("t_width_min", C.c_float), ("t_width_max", C.c_float),
        ("t_len_min", C.c_float), ("t_len_max", C.c_float), ("t_area_max", C.c_float),
        ("t_ratio_min", C.c_float), ("t_ratio_max", C.c_float), ("min_len_ratio", C.c_float), ("t_pt_per_m3", C.c_float),
        ("min_cluster_points", C.c_int), ("rule_filter", C.c_int),
        ("oracle_compat_first_frame", C.c_int),
        ("max_points", C.c_int), ("max_clusters", C.c_int), ("max_boxes", C.c_int), ("max_tracks", C.c_int),
        ("node_prefilter", C.c_int), ("filter_z_min", C.c_float), ("filter_z_max", C.c_float),
        ("filter_x_min", C.c_float), ("filter_x_max", C.c_float), ("filter_y_min", C.c_float), ("filter_y_max", C.c_float),
        ("global_frame", C.c_int),
        ("pipeline_depth", C.c_int), ("result_ring", C.c_int),
    ]


class TrackOut(C.Structure):
    _fields_ = [
        ("cap", C.c_int), ("n_tracks", C.c_int), ("n_vis", C.c_int),
        ("targets", C.POINTER(C.c_float)), ("vandyaw", C.POINTER(C.c_double)), ("track_manage", C.POINTER(C.c_int32)),
        ("is_static", C.POINTER(C.c_uint8)), ("is_vis", C.POINTER(C.c_uint8)), ("vis_bb", C.POINTER(C.c_float)),
    ]


MAX_BATCH = 8


class BatchOut(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int),
        ("n_elevated", C.c_int * MAX_BATCH), ("n_ground", C.c_int * MAX_BATCH), ("num_cluster", C.c_int * MAX_BATCH), ("n_boxes", C.c_int * MAX_BATCH),
        ("n_boxes_total", C.c_int), ("boxes", C.POINTER(C.c_float)), ("max_boxes", C.c_int), ("tracks", TrackOut),
    ]


class FrameOut(C.Structure):
    _fields_ = [
        ("n_elevated", C.c_int), ("n_ground", C.c_int), ("num_cluster", C.c_int), ("n_boxes", C.c_int),
        ("boxes", C.POINTER(C.c_float)), ("max_boxes", C.c_int), ("tracks", TrackOut),
    ]


# every symbol include/lmot.h declares (tests assert the library exports all of them)
ABI_SYMBOLS = [
    "lmot_default_params", "lmot_create", "lmot_destroy", "lmot_get_params", "lmot_strerror", "lmot_last_error", "lmot_build_info",
    "lmot_set_stream", "lmot_pinned_alloc", "lmot_pinned_free", "lmot_ground_remove", "lmot_component_cluster", "lmot_cluster_outputs", "lmot_box_fit", "lmot_track_step", "lmot_frame",
    "lmot_frame_dev", "lmot_frame_fetch", "lmot_frame_submit", "lmot_frame_collect", "lmot_frames_in_flight", "lmot_frame_ready", "lmot_flush",
    "lmot_ground_remove_dev", "lmot_detect_dev", "lmot_sync",
    "lmot_batch_submit", "lmot_batch_dev", "lmot_batch_detect_dev", "lmot_batch_collect", "lmot_batch_fetch", "lmot_batch", "lmot_batch_ground_ccl_dev",
    "lmot_debug_stage_clocks", "lmot_detect_boxes_dev", "lmot_track_step_lists_dev", "lmot_tracker_counters_dev", "lmot_tracker_table_received",
    "lmot_origin_points", "lmot_tracker_table", "lmot_tracker_set_num_tracks", "lmot_tracker_reset", "lmot_tracker_num_tracks", "lmot_tracker_dump", "lmot_tracker_load", "lmot_tracker_get_ego", "lmot_tracker_set_ego",
    "lmot_debug_polar_grid", "lmot_debug_cell_index", "lmot_debug_label_grid", "lmot_debug_phase_clock", "lmot_debug_timeline", "lmot_debug_tracker_trace", "lmot_selftest_atan2f",
    "lmot_enable_timing", "lmot_last_stage_ms", "lmot_last_kernel_ms", "lmot_debug_host_ns",
]

_lib = None


def load_library() -> C.CDLL:
    """Load liblmot.so (raises OSError when it has not been built: there is nothing to fall back to)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} not built -- run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(LIB_PATH)
        lib.lmot_strerror.restype = C.c_char_p
        lib.lmot_last_error.restype = C.c_char_p
        lib.lmot_build_info.restype = C.c_char_p
        lib.lmot_last_error.argtypes = [C.c_void_p]
        lib.lmot_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Params), C.c_int]
        lib.lmot_destroy.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def default_params() -> Params:
    p = Params()
    load_library().lmot_default_params(C.byref(p))
    return p


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _pts(points):
    p = np.ascontiguousarray(points, dtype=np.float32)
    if p.ndim != 2 or p.shape[1] < 3:
        raise ValueError("points must be (N, >=3) float32")
    return p, int(p.shape[0]), int(p.shape[1])


class Lmot:
    """One context = one sensor stream on one GPU (not thread-safe, like the reference's globals)."""

    def __init__(self, params: Params | None = None, device: int = 0):
        self.lib = load_library()
        self.params = params if params is not None else default_params()
        h = C.c_void_p()
        st = self.lib.lmot_create(C.byref(h), C.byref(self.params), device)
        if st != OK:
            raise LmotError(st, self.lib.lmot_strerror(st).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.lmot_destroy(self.h)
            self.h = None
            self.lib.lmot_pinned_free.argtypes = [C.c_void_p]
            for ptr in getattr(self, "_pinned", []):
                self.lib.lmot_pinned_free(ptr)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, st: int):
        if st < 0:
            raise LmotError(st, self.lib.lmot_strerror(st).decode() + " | " + self.lib.lmot_last_error(self.h).decode())
        self.last_warning = st          # > 0: LMOT_WARN_* (the call's outputs are valid)

    def set_stream(self, cuda_stream_handle: int | None):
        self._chk(self.lib.lmot_set_stream(self.h, C.c_void_p(cuda_stream_handle or 0)))

    def sync(self):
        self._chk(self.lib.lmot_sync(self.h))

    # ---- groundRemove ------------------------------------------------------------------------------------
    def ground_remove(self, points):
        """-> dict(labels (N,) u8, elevated (Ne,4) f32, ground (Ng,4) f32); clouds keep input order."""
        p, n, s = _pts(points)
        labels = np.zeros(max(n, 1), np.uint8)
        elev = np.zeros((max(n, 1), 4), np.float32)
        grnd = np.zeros((max(n, 1), 4), np.float32)
        ne, ng = C.c_int(0), C.c_int(0)
        self._chk(self.lib.lmot_ground_remove(self.h, _fp(p), n, s, labels.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(elev),
                                              C.byref(ne), _fp(grnd), C.byref(ng)))
        return dict(labels=labels[:n], elevated=elev[: ne.value], ground=grnd[: ng.value])

    def ground_remove_dev(self, d_ptr: int, n: int):
        self._chk(self.lib.lmot_ground_remove_dev(self.h, C.c_void_p(d_ptr), n))

    def cluster_outputs(self, cap: int | None = None):
        """Side outputs of the cluster node for the last clustering: (clustered (n,4), obstacles (m,4) [x,y,z,cluster], cost_map (50,50))"""
        cap = cap or int(self.params.max_points)
        cl = np.zeros((cap, 4), np.float32); ob = np.zeros((62500, 4), np.float32); cm = np.zeros(2500, np.int32)
        ncl = C.c_int(0); nob = C.c_int(0)
        self._chk(self.lib.lmot_cluster_outputs(self.h, _fp(cl), cap, C.byref(ncl), _fp(ob), 62500, C.byref(nob),
                                                cm.ctypes.data_as(C.POINTER(C.c_int32))))
        return cl[: ncl.value].copy(), ob[: nob.value].copy(), cm.reshape(50, 50)

    # ---- componentClustering ------------------------------------------------------------------------------
    def component_cluster(self, elevated):
        """-> (grid (250,250) int32 x-major, num_cluster)"""
        if len(elevated) == 0:
            p, n, s = np.zeros((1, 4), np.float32), 0, 4
        else:
            p, n, s = _pts(elevated)
        grid = np.zeros(250 * 250, np.int32)
        nc = C.c_int(0)
        self._chk(self.lib.lmot_component_cluster(self.h, _fp(p), n, s, grid.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(nc)))
        return grid.reshape(250, 250), nc.value

    # ---- boxFitting ---------------------------------------------------------------------------------------
    def box_fit(self, elevated, grid, num_cluster, max_boxes: int = 1024):
        """-> (boxes (B,8,3) f32, markers (B,6) f32) in cluster-id order"""
        if len(elevated) == 0:
            p, n, s = np.zeros((1, 4), np.float32), 0, 4
        else:
            p, n, s = _pts(elevated)
        g = np.ascontiguousarray(grid, np.int32).reshape(-1)
        boxes = np.zeros((max_boxes, 8, 3), np.float32)
        markers = np.zeros((max_boxes, 6), np.float32)
        nb = C.c_int(0)
        self._chk(self.lib.lmot_box_fit(self.h, _fp(p), n, s, g.ctypes.data_as(C.POINTER(C.c_int32)), int(num_cluster), _fp(boxes),
                                        max_boxes, C.byref(nb), _fp(markers)))
        return boxes[: nb.value].copy(), markers[: nb.value].copy()

    # ---- getOriginPoints + immUkfJpdaf -------------------------------------------------------------------
    def _track_out(self, cap):
        """Output buffers are allocated once per (context, cap) and reused: the hot loop must not allocate megabytes."""
        cache = self.__dict__.setdefault("_out_cache", {})
        if cap not in cache:
            bufs = dict(targets=np.zeros((cap, 3), np.float32), vandyaw=np.zeros((cap, 2), np.float64),
                        track_manage=np.zeros(cap, np.int32), is_static=np.zeros(cap, np.uint8), is_vis=np.zeros(cap, np.uint8),
                        vis_bb=np.zeros((cap, 8, 3), np.float32), boxes=np.zeros((self.params.max_boxes, 8, 3), np.float32))
            to = TrackOut()
            to.cap = cap
            to.targets = _fp(bufs["targets"]); to.vandyaw = bufs["vandyaw"].ctypes.data_as(C.POINTER(C.c_double))
            to.track_manage = bufs["track_manage"].ctypes.data_as(C.POINTER(C.c_int32))
            to.is_static = bufs["is_static"].ctypes.data_as(C.POINTER(C.c_uint8))
            to.is_vis = bufs["is_vis"].ctypes.data_as(C.POINTER(C.c_uint8)); to.vis_bb = _fp(bufs["vis_bb"])
            fo = FrameOut()
            fo.boxes = _fp(bufs["boxes"]); fo.max_boxes = self.params.max_boxes
            fo_nb = FrameOut()
            fo_nb.max_boxes = self.params.max_boxes
            cache[cap] = (to, bufs, fo, fo_nb)
        to, bufs, _, _ = cache[cap]
        return to, bufs

    def _frame_out(self, cap, want_boxes=True):
        to, bufs = self._track_out(cap)
        _, _, fo, fo_nb = self._out_cache[cap]
        f = fo if want_boxes else fo_nb
        f.tracks = to
        return f, bufs

    @staticmethod
    def _track_result(to, bufs):
        t, v = to.n_tracks, to.n_vis
        return dict(targets=bufs["targets"][:t].copy(), vandyaw=bufs["vandyaw"][:t].copy(),
                    track_manage=bufs["track_manage"][:t].copy(), is_static=bufs["is_static"][:t].copy(),
                    is_vis=bufs["is_vis"][:t].copy(), vis_bb=bufs["vis_bb"][:v].copy())

    @staticmethod
    def _frame_result(fo, bufs, want_boxes=True):
        r = Lmot._track_result(fo.tracks, bufs)
        r.update(n_elevated=fo.n_elevated, n_ground=fo.n_ground, num_cluster=fo.num_cluster,
                 boxes=bufs["boxes"][: fo.n_boxes].copy() if want_boxes else bufs["boxes"][:0])
        return r

    def track_step(self, boxes, timestamp_us, v_gps=0.0, yaw_gps=0.0, cap: int | None = None):
        b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 8, 3)
        m = b.shape[0]
        if m == 0:
            b = np.zeros((1, 8, 3), np.float32)
        cap = cap or self.params.max_tracks
        to, bufs = self._track_out(cap)
        self._chk(self.lib.lmot_track_step(self.h, _fp(b), m, C.c_double(timestamp_us), C.c_double(v_gps), C.c_double(yaw_gps), C.byref(to)))
        return self._track_result(to, bufs)

    def frame(self, points, timestamp_us, v_gps=0.0, yaw_gps=0.0, cap: int | None = None):
        """The whole hot path on one host frame -> dict(n_elevated, n_ground, num_cluster, boxes, tracks...)."""
        p, n, s = _pts(points)
        fo, bufs = self._frame_out(cap or self.params.max_tracks)
        self._chk(self.lib.lmot_frame(self.h, _fp(p), n, s, C.c_double(timestamp_us), C.c_double(v_gps), C.c_double(yaw_gps), C.byref(fo)))
        return self._frame_result(fo, bufs)

    def frame_dev(self, d_ptr: int, n: int, timestamp_us, v_gps=0.0, yaw_gps=0.0):
        """Asynchronous: device-resident XYZI frame (stride 4) through all four stages on the context's stream."""
        self._chk(self.lib.lmot_frame_dev(self.h, C.c_void_p(d_ptr), n, C.c_double(timestamp_us), C.c_double(v_gps), C.c_double(yaw_gps)))

    def flush(self):
        """Device-side join: the caller stream waits for everything submitted so far."""
        self._chk(self.lib.lmot_flush(self.h))

    def pinned_array(self, shape, dtype=np.float32) -> np.ndarray:
        """numpy array in page-locked host memory (lmot_pinned_alloc); freed when the context closes."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.lib.lmot_pinned_alloc.restype = C.c_void_p
        self.lib.lmot_pinned_alloc.argtypes = [C.c_size_t]
        ptr = self.lib.lmot_pinned_alloc(nbytes)
        if not ptr:
            raise LmotError(-1, "lmot_pinned_alloc failed")
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        buf = (C.c_char * nbytes).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def frame_submit(self, points, timestamp_us, v_gps=0.0, yaw_gps=0.0):
        """Asynchronous host frame (pinned memory recommended); collect results later with frame_collect()."""
        p, n, s = _pts(points)
        self._chk(self.lib.lmot_frame_submit(self.h, _fp(p), n, s, C.c_double(timestamp_us), C.c_double(v_gps), C.c_double(yaw_gps)))

    def frame_collect(self, cap: int | None = None, want_boxes: bool = True):
        """Results of the oldest submitted frame (blocks until it is done)."""
        fo, bufs = self._frame_out(cap or self.params.max_tracks, want_boxes)
        self._chk(self.lib.lmot_frame_collect(self.h, C.byref(fo)))
        return self._frame_result(fo, bufs, want_boxes)

    def frame_ready(self) -> bool:
        return self.lib.lmot_frame_ready(self.h) == 1

    def frames_in_flight(self) -> int:
        n = C.c_int(0)
        self._chk(self.lib.lmot_frames_in_flight(self.h, C.byref(n)))
        return n.value

    def detect_dev(self, d_ptr: int, n: int):
        self._chk(self.lib.lmot_detect_dev(self.h, C.c_void_p(d_ptr), n))

    def frame_fetch(self, cap: int | None = None, want_boxes: bool = True):
        fo, bufs = self._frame_out(cap or self.params.max_tracks, want_boxes)
        self._chk(self.lib.lmot_frame_fetch(self.h, C.byref(fo)))
        return self._frame_result(fo, bufs, want_boxes)

    # ---- batched ticks: F sensor streams, one frame each, one shared track table (BASELINE.json configs[3]) --------
    def batch_prepare(self, dev_frames):
        """Marshal a tick's (device pointer, n) list ONCE: the returned handle can be passed to batch_dev / batch_detect_dev /
        batch_ground_ccl_dev in place of the list (a timing loop must not spend 20 us per call building ctypes arrays)."""
        return ("prepared",) + tuple(self._batch_args(dev_frames, True))

    def _batch_args(self, frames, device: bool):
        if isinstance(frames, tuple) and frames and frames[0] == "prepared":
            return frames[1:]
        F = len(frames)
        ptrs = (C.c_void_p * F)()
        ns = (C.c_int * F)()
        keep = []
        stride = 4
        for i, f in enumerate(frames):
            if device:
                ptrs[i], ns[i] = int(f[0]), int(f[1])
            else:
                p, n, stride = _pts(f)
                keep.append(p)
                ptrs[i], ns[i] = p.ctypes.data, n
        return ptrs, ns, F, stride, keep

    def _batch_result(self, bo, bufs, want_boxes=True):
        r = Lmot._track_result(bo.tracks, bufs)
        F = bo.n_frames
        r.update(n_frames=F, n_elevated=list(bo.n_elevated[:F]), n_ground=list(bo.n_ground[:F]), num_cluster=list(bo.num_cluster[:F]),
                 n_boxes=list(bo.n_boxes[:F]), boxes=bufs["boxes"][: bo.n_boxes_total].copy() if want_boxes else bufs["boxes"][:0])
        return r

    def _batch_out(self, cap, want_boxes=True):
        to, bufs = self._track_out(cap)
        bo = BatchOut()
        if want_boxes:
            bo.boxes = _fp(bufs["boxes"])
        bo.max_boxes = self.params.max_boxes
        bo.tracks = to
        return bo, bufs

    def batch(self, frames, timestamp_us, v_gps=0.0, yaw_gps=0.0, cap: int | None = None):
        """One tick, synchronous: `frames` = list of (N_i, >=3) host arrays (one per sensor stream)."""
        ptrs, ns, F, stride, keep = self._batch_args(frames, False)
        bo, bufs = self._batch_out(cap or self.params.max_tracks)
        self._chk(self.lib.lmot_batch(self.h, ptrs, ns, F, stride, C.c_double(timestamp_us), C.c_double(v_gps), C.c_double(yaw_gps), C.byref(bo)))
        return self._batch_result(bo, bufs)

    def batch_submit(self, frames, timestamp_us, v_gps=0.0, yaw_gps=0.0):
        ptrs, ns, F, stride, keep = self._batch_args(frames, False)
        self._keep = keep          # the copies are asynchronous: the host arrays must outlive the call
        self._chk(self.lib.lmot_batch_submit(self.h, ptrs, ns, F, stride, C.c_double(timestamp_us), C.c_double(v_gps), C.c_double(yaw_gps)))

    def batch_dev(self, dev_frames, timestamp_us, v_gps=0.0, yaw_gps=0.0):
        """dev_frames = list of (device pointer, n) of stride-4 float frames."""
        ptrs, ns, F, _, _ = self._batch_args(dev_frames, True)
        self._chk(self.lib.lmot_batch_dev(self.h, ptrs, ns, F, C.c_double(timestamp_us), C.c_double(v_gps), C.c_double(yaw_gps)))

    def batch_detect_dev(self, dev_frames):
        ptrs, ns, F, _, _ = self._batch_args(dev_frames, True)
        self._chk(self.lib.lmot_batch_detect_dev(self.h, ptrs, ns, F))

    def batch_ground_ccl_dev(self, dev_frames):
        ptrs, ns, F, _, _ = self._batch_args(dev_frames, True)
        self._chk(self.lib.lmot_batch_ground_ccl_dev(self.h, ptrs, ns, F))

    def batch_collect(self, cap: int | None = None, want_boxes: bool = True):
        bo, bufs = self._batch_out(cap or self.params.max_tracks, want_boxes)
        self._chk(self.lib.lmot_batch_collect(self.h, C.byref(bo)))
        return self._batch_result(bo, bufs, want_boxes)

    def batch_fetch(self, cap: int | None = None, want_boxes: bool = True):
        bo, bufs = self._batch_out(cap or self.params.max_tracks, want_boxes)
        self._chk(self.lib.lmot_batch_fetch(self.h, C.byref(bo)))
        return self._batch_result(bo, bufs, want_boxes)

    def debug_stage_clocks(self, which: int):
        """which = 1: (frames, 16) stamps of the last ccl_bitmap_kernel; 2: (CTAs, 8) of the last box_fit_kernel (phase clock armed)."""
        buf = np.zeros((4096, 16), np.uint64)
        n = C.c_int(0); w = C.c_int(0)
        self._chk(self.lib.lmot_debug_stage_clocks(self.h, int(which), buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), 4096, C.byref(n), C.byref(w)))
        return buf.reshape(-1)[: n.value * w.value].reshape(n.value, w.value).copy()

    def tracker_reset(self):
        self._chk(self.lib.lmot_tracker_reset(self.h))

    def tracker_num_tracks(self) -> int:
        n = C.c_int(0)
        self._chk(self.lib.lmot_tracker_num_tracks(self.h, C.byref(n)))
        return n.value

    def tracker_dump(self) -> np.ndarray:
        t = self.tracker_num_tracks()
        out = np.zeros((max(t, 1), TRACK_DUMP_DOUBLES), np.float64)
        n = C.c_int(0)
        self._chk(self.lib.lmot_tracker_dump(self.h, out.ctypes.data_as(C.POINTER(C.c_double)), max(t, 1), C.byref(n)))
        return out[: n.value]

    def tracker_load(self, dumps, init, timestamp_us, ego_velo=0.0, ego_yaw=0.0, ego_pre_yaw=0.0, ego_point_yaw=-np.pi / 2):
        d = np.ascontiguousarray(dumps, np.float64).reshape(-1, TRACK_DUMP_DOUBLES)
        n = d.shape[0]
        if n == 0:
            d = np.zeros((1, TRACK_DUMP_DOUBLES), np.float64)
        self._chk(self.lib.lmot_tracker_load(self.h, d.ctypes.data_as(C.POINTER(C.c_double)), n, int(init), C.c_double(timestamp_us),
                                             C.c_double(ego_velo), C.c_double(ego_yaw), C.c_double(ego_pre_yaw), C.c_double(ego_point_yaw)))

    def origin_points(self, timestamp_us, v_gps=0.0, yaw_gps=0.0) -> np.ndarray:
        """getOriginPoints for this frame WITHOUT advancing the tracker: (x, y, yaw, x, y, yaw + pi/2)."""
        o = np.zeros(6, np.float64)
        self._chk(self.lib.lmot_origin_points(self.h, C.c_double(timestamp_us), C.c_double(v_gps), C.c_double(yaw_gps), o.ctypes.data_as(C.POINTER(C.c_double))))
        return o

    def tracker_get_ego(self) -> np.ndarray:
        e = np.zeros(8, np.float64)
        self._chk(self.lib.lmot_tracker_get_ego(self.h, e.ctypes.data_as(C.POINTER(C.c_double))))
        return e

    def tracker_set_ego(self, ego8):
        e = np.ascontiguousarray(ego8, np.float64)
        assert e.shape == (8,)
        self._chk(self.lib.lmot_tracker_set_ego(self.h, e.ctypes.data_as(C.POINTER(C.c_double))))

    def tracker_table(self):
        """-> (device pointer, bytes per track, capacity) of the track table (for NCCL broadcast, see shared_tracker.py)"""
        ptr = C.c_void_p(); bpt = C.c_int(0); cap = C.c_int(0)
        self._chk(self.lib.lmot_tracker_table(self.h, C.byref(ptr), C.byref(bpt), C.byref(cap)))
        return int(ptr.value), bpt.value, cap.value

    def tracker_set_num_tracks(self, n: int):
        self._chk(self.lib.lmot_tracker_set_num_tracks(self.h, int(n)))

    def enable_timing(self, on: bool = True):
        self._chk(self.lib.lmot_enable_timing(self.h, int(on)))

    def last_stage_ms(self):
        ms = (C.c_float * 4)()
        self._chk(self.lib.lmot_last_stage_ms(self.h, ms))
        return [float(x) for x in ms]

    def debug_host_ns(self, reset=True):
        ns = (C.c_double * 4)()
        self._chk(self.lib.lmot_debug_host_ns(self.h, ns, int(reset)))
        return [float(x) for x in ns]

    def last_kernel_ms(self):
        ms = (C.c_float * 32)(); n = C.c_int(0)
        self._chk(self.lib.lmot_last_kernel_ms(self.h, ms, 32, C.byref(n)))
        return [float(ms[i]) for i in range(min(n.value, 32))]

    def debug_label_grid(self):
        grid = np.zeros(250 * 250, np.int32)
        nc = C.c_int(0)
        self._chk(self.lib.lmot_debug_label_grid(self.h, grid.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(nc)))
        return grid.reshape(250, 250), nc.value

    def debug_timeline(self):
        """(frames, 5 + kernels) float32 ms: completion of the stage boundaries and kernels of the frames still in the result ring."""
        buf = np.full((64, 64), -1.0, np.float32)
        n = C.c_int(0); stride = C.c_int(0)
        self._chk(self.lib.lmot_debug_timeline(self.h, buf.ctypes.data_as(C.POINTER(C.c_float)), 64, C.byref(n), C.byref(stride)))
        flat = buf.reshape(-1)[: n.value * stride.value].reshape(n.value, stride.value) if n.value else buf[:0]
        return flat

    def debug_tracker_trace(self):
        """(32, 8) uint64 ns, oldest step first: TA start, TA end, TB start, TB end, TC start, TC end, latest start of a working TA CTA, same for TB."""
        raw = np.zeros(32 * 8 + 64, np.uint64); nxt = C.c_int(0)
        self._chk(self.lib.lmot_debug_tracker_trace(self.h, raw.ctypes.data_as(C.POINTER(C.c_ulonglong)), C.byref(nxt)))
        buf = raw[:256].reshape(32, 8).copy()
        self.last_tc_phases = raw[256:272].copy()
        self.last_tb_phases = raw[272:282].copy()
        self.last_ta_phases = raw[288:304].copy()    # CTA 0 of imm_predict_gate_kernel
        self.last_gate_phases = raw[282:288].copy()     # [0] gate kernel entry, [1] gate kernel after its wait, [2] TA CTA 0 entry (before its wait)        # same for the first track's warp of imm_update_kernel       # %globaltimer stamps inside the last spawn_output_kernel (fast path)
        for k in (0, 2, 4):
            buf[:, k] = ~buf[:, k]        # starts are stored complemented (see trace_start in tracker.cu)
        return np.roll(buf, -nxt.value, axis=0)

    def debug_phase_clock(self):
        """First call arms the phase clock of ground_fused_kernel; later calls -> (n_ctas, 8) uint64 ns stamps of the last launch."""
        buf = np.zeros((1024, 8), np.uint64)
        n = C.c_int(0)
        self._chk(self.lib.lmot_debug_phase_clock(self.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), 1024, C.byref(n)))
        return buf[: n.value]

    def debug_polar_grid(self):
        names = ["minz", "height", "smoothed", "hdiff", "hground"]
        arrs = [np.zeros(80 * 120, np.float32) for _ in names]
        isg = np.zeros(80 * 120, np.uint8)
        self._chk(self.lib.lmot_debug_polar_grid(self.h, *[_fp(a) for a in arrs], isg.ctypes.data_as(C.POINTER(C.c_uint8))))
        d = {k: v.reshape(80, 120) for k, v in zip(names, arrs)}
        d["isground"] = isg.reshape(80, 120)
        return d

    def debug_cell_index(self, n: int):
        ch = np.zeros(max(n, 1), np.int32); b = np.zeros(max(n, 1), np.int32)
        self._chk(self.lib.lmot_debug_cell_index(self.h, ch.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 b.ctypes.data_as(C.POINTER(C.c_int32)), n))
        return ch[:n], b[:n]


def selftest_atan2f(y, x) -> np.ndarray:
    """Host-side evaluation of the library's bit-exact atan2f restatement (no GPU needed)."""
    lib = load_library()
    y = np.ascontiguousarray(y, np.float32); x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    st = lib.lmot_selftest_atan2f(_fp(y), _fp(x), int(y.size), _fp(out))
    if st != OK:
        raise LmotError(st, "selftest")
    return out


# ---- several GPUs, one track table: ctypes mirror of include/lmot_shared.h (liblmot_shared.so = host/shared_tracker.cpp + NCCL) ----
SHARED_LIB_PATH = os.path.join(HERE, "liblmot_shared.so")
SHARED_STREAMS, SHARED_FRAMES = 0, 1
SHARED_ABI_SYMBOLS = ["lmot_shared_unique_id", "lmot_shared_create", "lmot_shared_destroy", "lmot_shared_tick_dev", "lmot_shared_last_us", "lmot_shared_last_error"]
_shared_lib = None


def load_shared_library() -> C.CDLL:
    global _shared_lib
    if _shared_lib is None:
        load_library()
        if not os.path.exists(SHARED_LIB_PATH):
            raise OSError(f"{SHARED_LIB_PATH} not built")
        lib = C.CDLL(SHARED_LIB_PATH)
        lib.lmot_shared_last_error.restype = C.c_char_p
        lib.lmot_shared_last_error.argtypes = [C.c_void_p]
        lib.lmot_shared_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p]
        lib.lmot_shared_destroy.argtypes = [C.c_void_p]
        lib.lmot_shared_tick_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.POINTER(TrackOut)]
        lib.lmot_shared_last_us.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        _shared_lib = lib
    return _shared_lib


def shared_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    st = load_shared_library().lmot_shared_unique_id(buf)
    if st != OK:
        raise LmotError(st, "lmot_shared_unique_id")
    return buf.raw


class SharedTracker:
    """One rank of `world` GPUs feeding ONE track table (owner rank runs the tracker; NCCL all-gather of boxes, broadcast of the table)."""

    def __init__(self, ctx: Lmot, rank: int, world: int, owner: int, unique_id: bytes):
        self.lib = load_shared_library()
        self.ctx, self.rank, self.world, self.owner = ctx, rank, world, owner
        h = C.c_void_p()
        st = self.lib.lmot_shared_create(C.byref(h), ctx.h, rank, world, owner, unique_id)
        if st != OK:
            raise LmotError(st, "lmot_shared_create")
        self.h = h

    def tick_dev(self, d_ptr: int, n: int, timestamp_us, v_gps=0.0, yaw_gps=0.0, mode=SHARED_STREAMS, frame_dt_us=0.0, cap: int | None = None):
        to, bufs = self.ctx._track_out(cap or self.ctx.params.max_tracks)
        st = self.lib.lmot_shared_tick_dev(self.h, C.c_void_p(d_ptr), int(n), float(timestamp_us), float(v_gps), float(yaw_gps), int(mode), float(frame_dt_us), C.byref(to))
        if st < 0:
            raise LmotError(st, self.lib.lmot_shared_last_error(self.h).decode())
        if self.rank == self.owner:
            return Lmot._track_result(to, bufs)
        return dict(n_tracks=to.n_tracks)

    def last_us(self):
        us = (C.c_float * 4)()
        self.lib.lmot_shared_last_us(self.h, us)
        return dict(zip(("detect", "allgather", "tracker", "broadcast"), (float(x) for x in us)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.lmot_shared_destroy(self.h)
            self.h = None

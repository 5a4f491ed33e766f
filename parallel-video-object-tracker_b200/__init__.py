"""ctypes binding of libpvt.so (include/pvt.h) -- the B200-native NCC tracking hot path.

The directory name contains hyphens, so import it with
    pvt = importlib.import_module("parallel-video-object-tracker_b200")
(tests/, bench.py and __graft_entry__.py do).  This module holds NO compute: every call goes through
the C ABI into hand-written sm_100a CUDA kernels, and loading fails loudly when the library has not
been built (`make -C parallel-video-object-tracker_b200/csrc`, or __graft_entry__.build()).  There is
no CPU fallback here or in the library.

Reference surface mirrored (paths under /root/reference/tracker):
    include/baseline_kernel.hpp:8-17  ncc_match_naive_cuda / _shared_cuda / _const / _const_tiled / _batched
    include/utils.hpp:5-14            toGrayF32
    src/main.cpp:6-20                 SEARCH_RADIUS_X/Y, NCC_MIN/STRONG_CONFIDENCE, TEMPLATE_UPDATE_LR, BATCH_SIZE
    src/main.cpp:93-169               the per-frame loop  -> Tracker.step / Tracker.submit
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpvt.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE, ERR_NOMEM = 0, -1, -2, -3, -4, -5
MODE_NAIVE, MODE_CPU, MODE_SHARED, MODE_CONST, MODE_CONST_TILED, MODE_BATCH = range(6)
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_TC, KERNEL_TC_GLOBAL = range(4)
FMT_BGR8, FMT_GRAY8, FMT_GRAYF32 = range(3)
MEM_HOST, MEM_DEVICE, MEM_HOST_PINNED = range(3)
INGEST_AUTO, INGEST_FULL, INGEST_ROI = range(3)
FORMULA_CCOEFF_NORMED, FORMULA_EPS = range(2)   # pvt_formula: the reference's CPU operator / its CUDA kernels' eps formula
MODE_FLAG_EPS = 0x100

# every symbol include/pvt.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = [
    "pvt_version", "pvt_last_error", "pvt_device_count", "pvt_device_info", "pvt_default_params", "pvt_create",
    "pvt_destroy", "pvt_set_params", "pvt_alloc_pinned", "pvt_free_pinned", "pvt_track_init", "pvt_track_remove",
    "pvt_step", "pvt_submit", "pvt_collect", "pvt_submit_sequence", "pvt_sync", "pvt_get_state", "pvt_set_state", "pvt_get_window_map",
    "pvt_to_gray_f32", "pvt_ncc_match", "pvt_ncc_match_batched", "pvt_profile_enable", "pvt_profile_get",
    "pvt_launch_count", "pvt_timer_start", "pvt_timer_stop", "pvt_trace_enable", "pvt_trace_get",
    "pvt_default_params_ghc", "pvt_get_lost_state", "pvt_set_lost_state", "pvt_plan_query", "pvt_tc_plan_query", "pvt_ncc_match_batched_f",
    "pvt_search_kind", "pvt_draw_boxes",
]


class Params(C.Structure):
    _fields_ = [("search_radius_x", C.c_int), ("search_radius_y", C.c_int),
                ("ncc_min_confidence", C.c_double), ("ncc_strong_confidence", C.c_double),
                ("template_update_lr", C.c_double), ("batch_size", C.c_int), ("mode", C.c_int),
                ("kernel", C.c_int), ("keep_maps", C.c_int), ("ingest", C.c_int), ("lost_frame_threshold", C.c_int),
                ("formula", C.c_int), ("reserved", C.c_int), ("ncc_global_confidence", C.c_double)]


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("frame_w", C.c_int), ("frame_h", C.c_int), ("max_streams", C.c_int),
                ("max_tracks", C.c_int), ("max_templ_w", C.c_int), ("max_templ_h", C.c_int),
                ("max_radius_x", C.c_int), ("max_radius_y", C.c_int), ("reserved", C.c_int * 7)]


class Frame(C.Structure):
    _fields_ = [("stream", C.c_int), ("format", C.c_int), ("memory", C.c_int), ("reserved", C.c_int),
                ("data", C.c_void_p), ("step", C.c_size_t)]


class Result(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("conf", C.c_float),
                ("moved", C.c_uint8), ("updated", C.c_uint8), ("searched", C.c_uint8), ("valid", C.c_uint8),
                ("track", C.c_int32), ("step", C.c_int32)]


RESULT_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("conf", "<f4"),
                         ("moved", "u1"), ("updated", "u1"), ("searched", "u1"), ("valid", "u1"),
                         ("track", "<i4"), ("step", "<i4")])
assert RESULT_DTYPE.itemsize == C.sizeof(Result) == 32


class Profile(C.Structure):
    _fields_ = [("ingest_ms", C.c_double), ("stats_ms", C.c_double), ("ncc_ms", C.c_double), ("update_ms", C.c_double),
                ("ingest_launches", C.c_int64), ("stats_launches", C.c_int64), ("ncc_launches", C.c_int64),
                ("update_launches", C.c_int64), ("steps", C.c_int64), ("ncc_macs", C.c_double), ("ingest_bytes", C.c_double),
                ("search_kernel_ms", C.c_double), ("search_kernel_macs", C.c_double)]


class PvtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpvt error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libpvt.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built: run `make -C {os.path.join(_HERE, 'csrc')}` "
                          "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    L.pvt_last_error.restype = C.c_char_p
    L.pvt_launch_count.restype = C.c_int64
    L.pvt_launch_count.argtypes = [C.c_void_p]
    L.pvt_draw_boxes.restype = C.c_int
    L.pvt_draw_boxes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_void_p]
    L.pvt_search_kind.restype = C.c_int
    L.pvt_search_kind.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.pvt_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Params), C.POINTER(Config)]
    L.pvt_destroy.argtypes = [C.c_void_p]
    L.pvt_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
    L.pvt_alloc_pinned.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.pvt_free_pinned.argtypes = [C.c_void_p]
    L.pvt_track_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(Frame), C.c_int, C.c_int, C.c_int, C.c_int]
    L.pvt_track_remove.argtypes = [C.c_void_p, C.c_int]
    L.pvt_step.argtypes = [C.c_void_p, C.c_int, C.POINTER(Frame), C.c_void_p]
    L.pvt_submit.argtypes = [C.c_void_p, C.c_int, C.POINTER(Frame)]
    L.pvt_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.pvt_sync.argtypes = [C.c_void_p]
    L.pvt_submit_sequence.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(Frame), C.c_int, C.c_int, C.c_void_p]
    L.pvt_get_state.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_void_p, C.c_size_t]
    L.pvt_set_state.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_void_p, C.c_size_t]
    L.pvt_get_lost_state.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.pvt_set_lost_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.pvt_get_window_map.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int32)]
    L.pvt_to_gray_f32.argtypes = [C.c_void_p, C.POINTER(Frame), C.c_void_p, C.c_size_t, C.c_int]
    L.pvt_ncc_match.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int,
                                C.c_size_t, C.c_void_p, C.c_size_t]
    L.pvt_ncc_match_batched.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_size_t, C.c_void_p,
                                        C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_size_t]
    L.pvt_ncc_match_batched_f.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_size_t, C.c_void_p,
                                          C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_size_t]
    L.pvt_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.pvt_profile_get.argtypes = [C.c_void_p, C.POINTER(Profile), C.c_int]
    L.pvt_trace_enable.argtypes = [C.c_void_p, C.c_int]
    L.pvt_trace_get.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.pvt_timer_start.argtypes = [C.c_void_p]
    L.pvt_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.pvt_device_info.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    _lib = L
    return L


def _ck(rc):
    if rc < 0:
        raise PvtError(rc, lib().pvt_last_error().decode())
    return rc


def default_params(ghc=False, **kw) -> Params:
    """tracker/src/main.cpp:6-20 constants, or (ghc=True) those of tracker_ghc/src/main.cpp:9-23 (lost-object mode on)."""
    p = Params()
    (lib().pvt_default_params_ghc if ghc else lib().pvt_default_params)(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


PLAN_FIELDS = ("G", "C", "GB", "bands", "ctas_per_band", "span", "boxW", "boxH", "pj", "pd", "ctas_per_track", "n_full", "n_tail",
               "tail_parts", "smem", "fringe")


def plan_query(n_tracks, templ_w, templ_h, frame_w, frame_h, radius_x=80, radius_y=80, sm_count=148) -> dict:
    """The k_ncc_search plan pvt_create would derive (host logic only: works without a GPU)."""
    out = (C.c_int32 * 16)()
    _ck(lib().pvt_plan_query(sm_count, n_tracks, templ_w, templ_h, frame_w, frame_h, radius_x, radius_y, out))
    return dict(zip(PLAN_FIELDS, out))


TC_PLAN_FIELDS = ("xw", "xtiles", "mtiles", "ksteps", "groups", "tmem_cols", "stages", "smem")


def tc_plan_query(n_tracks, templ_w, templ_h, frame_w, frame_h, radius_x=80, radius_y=80, whole_frame_pass=False, sm_count=148) -> dict:
    """The k_ncc_tc geometry pvt_create would derive for the local windows / the whole-frame pass (host logic only)."""
    out = (C.c_int32 * 8)()
    _ck(lib().pvt_tc_plan_query(sm_count, n_tracks, templ_w, templ_h, frame_w, frame_h, radius_x, radius_y, 1 if whole_frame_pass else 0, out))
    return dict(zip(TC_PLAN_FIELDS, out))


def device_count() -> int:
    return _ck(lib().pvt_device_count())


def device_info(device=0) -> dict:
    sm, clk, mclk, mem, cc = C.c_int(), C.c_int(), C.c_int(), C.c_size_t(), C.c_int()
    _ck(lib().pvt_device_info(device, C.byref(sm), C.byref(clk), C.byref(mclk), C.byref(mem), C.byref(cc)))
    return dict(sm_count=sm.value, sm_clock_khz=clk.value, mem_clock_khz=mclk.value, mem_bytes=mem.value, cc=cc.value)


def _fmt_of(a: np.ndarray) -> int:
    if a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3:
        return FMT_BGR8
    if a.dtype == np.uint8 and a.ndim == 2:
        return FMT_GRAY8
    if a.dtype == np.float32 and a.ndim == 2:
        return FMT_GRAYF32
    raise ValueError("frame must be HxWx3 u8 (BGR), HxW u8 (gray) or HxW f32 (toGrayF32 output)")


def host_frame(a: np.ndarray, stream=0) -> Frame:
    """pvt_frame over a numpy array (kept alive by the caller); rows may be strided, pixels must be dense."""
    fmt = _fmt_of(a)
    px = 3 if fmt == FMT_BGR8 else 1 if fmt == FMT_GRAY8 else 4
    if a.strides[1] != px or (a.ndim == 3 and a.strides[2] != 1):
        a = np.ascontiguousarray(a)
    f = Frame(stream, fmt, MEM_HOST, 0, a.ctypes.data, a.strides[0])
    f._keep = a
    return f


def device_frame(ptr: int, step: int, fmt=FMT_BGR8, stream=0) -> Frame:
    return Frame(stream, fmt, MEM_DEVICE, 0, ptr, step)


class RingArray:
    """A frame ring as the contiguous pvt_frame array pvt_submit_sequence takes (ring_len x n_frames), built once."""

    def __init__(self, ring):
        self.n_frames = len(ring[0])
        self.ring_len = len(ring)
        flat = [f for st in ring for f in st]
        self.arr = (Frame * len(flat))(*flat)
        self._keep = flat

    def __len__(self):
        return self.ring_len


class PinnedBuffer:
    """cudaHostAlloc'ed bytes exposed as a numpy array (for frames streamed from the host)."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        _ck(lib().pvt_alloc_pinned(C.byref(self.ptr), nbytes))
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def free(self):
        if self.ptr:
            lib().pvt_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Tracker:
    """One pvt_ctx: the reference's per-frame loop (main.cpp:93-169) for any number of tracks on one GPU."""

    def __init__(self, frame_w, frame_h, max_templ_w, max_templ_h, max_streams=1, max_tracks=1, device=0,
                 max_radius_x=0, max_radius_y=0, **params):
        self.params = default_params(**params)
        self.cfg = Config(device, frame_w, frame_h, max_streams, max_tracks, max_templ_w, max_templ_h, max_radius_x, max_radius_y)
        self._h = C.c_void_p()
        _ck(lib().pvt_create(C.byref(self._h), C.byref(self.params), C.byref(self.cfg)))
        self.max_tracks = max_tracks

    def close(self):
        if self._h:
            lib().pvt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_params(self, **kw):
        for k, v in kw.items():
            setattr(self.params, k, v)
        _ck(lib().pvt_set_params(self._h, C.byref(self.params)))

    def init_track(self, track, frame0, roi, stream=0):
        """main.cpp:70-71.  frame0: numpy frame, a Frame, or None (reuse the stream's current image)."""
        x, y, w, h = (int(v) for v in roi)
        f = None
        if frame0 is not None:
            f = frame0 if isinstance(frame0, Frame) else host_frame(frame0, stream)
            f.stream = stream
        _ck(lib().pvt_track_init(self._h, track, stream, C.byref(f) if f is not None else None, x, y, w, h))

    def remove_track(self, track):
        _ck(lib().pvt_track_remove(self._h, track))

    def _frames(self, frames):
        if isinstance(frames, np.ndarray):
            frames = [frames]
        for f in frames:
            # pvt_frame carries no width/height (the geometry is the context's, like the reference's fixed-size
            # Mats); numpy inputs are checked here the way the reference CV_Asserts them (baseline_kernel.cu:419-422)
            if not isinstance(f, Frame) and tuple(f.shape[:2]) != (self.cfg.frame_h, self.cfg.frame_w):
                raise PvtError(ERR_INVALID, f"frame is {f.shape[1]}x{f.shape[0]}, context expects {self.cfg.frame_w}x{self.cfg.frame_h}")
        fl = [f if isinstance(f, Frame) else host_frame(f, i) for i, f in enumerate(frames)]
        arr = (Frame * len(fl))(*fl)
        arr._keep = fl
        return arr

    def step(self, frames) -> np.ndarray:
        """One synchronous time step; returns a RESULT_DTYPE array with one entry per track slot."""
        arr = self._frames(frames)
        out = np.zeros(self.max_tracks, RESULT_DTYPE)
        _ck(lib().pvt_step(self._h, len(arr), arr, out.ctypes.data))
        return out

    def submit(self, frames):
        arr = self._frames(frames)
        _ck(lib().pvt_submit(self._h, len(arr), arr))
        return arr  # keep host buffers alive until collect()

    def submit_sequence(self, n_steps, ring, collect_every=0, want_results=False):
        """main.cpp:93-169 as one call: n_steps steps cycling over `ring`, a list of per-step frame lists -- or a RingArray built
        once from such a list (a C caller passes the same pvt_frame array every time; building it is not part of the call)."""
        if isinstance(ring, RingArray):
            arr, n_frames, ring = ring.arr, ring.n_frames, ring
        else:
            n_frames = len(ring[0])
            flat = [f for st in ring for f in st]
            arr = (Frame * len(flat))(*flat)
        out = np.zeros((n_steps, self.max_tracks), RESULT_DTYPE) if (want_results and collect_every > 0) else None
        _ck(lib().pvt_submit_sequence(self._h, n_steps, n_frames, arr, len(ring), collect_every,
                                      out.ctypes.data if out is not None else None))
        return out

    def collect(self, max_steps=64) -> np.ndarray:
        out = np.zeros((max_steps, self.max_tracks), RESULT_DTYPE)
        n = _ck(lib().pvt_collect(self._h, out.ctypes.data, max_steps))
        return out[:n]

    def sync(self):
        _ck(lib().pvt_sync(self._h))

    def get_lost_state(self, track=0):
        """(lost_frame_count, use_global_search) of tracker_ghc/src/main.cpp:143-144."""
        a, b = C.c_int(), C.c_int()
        _ck(lib().pvt_get_lost_state(self._h, track, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_lost_state(self, track, lost_frame_count, use_global_search):
        _ck(lib().pvt_set_lost_state(self._h, track, int(lost_frame_count), int(use_global_search)))

    def get_state(self, track=0):
        bbox = (C.c_int32 * 4)()
        _ck(lib().pvt_get_state(self._h, track, bbox, None, 0))
        templ = np.empty((bbox[3], bbox[2]), np.float32)
        _ck(lib().pvt_get_state(self._h, track, bbox, templ.ctypes.data, templ.strides[0]))
        return tuple(bbox), templ

    def set_state(self, track, bbox, templ=None):
        b = (C.c_int32 * 4)(*[int(v) for v in bbox])
        if templ is not None:
            templ = np.ascontiguousarray(templ, np.float32)
            _ck(lib().pvt_set_state(self._h, track, b, templ.ctypes.data, templ.strides[0]))
        else:
            _ck(lib().pvt_set_state(self._h, track, b, None, 0))

    def window_map(self, track=0):
        win = (C.c_int32 * 4)()
        _ck(lib().pvt_get_window_map(self._h, track, None, 0, win))
        out = np.empty((win[3], win[2]), np.float32)
        _ck(lib().pvt_get_window_map(self._h, track, out.ctypes.data, out.strides[0], win))
        return out, tuple(win)

    def to_gray_f32(self, frame, stream=0) -> np.ndarray:
        f = frame if isinstance(frame, Frame) else host_frame(frame, stream)
        out = np.empty((self.cfg.frame_h, self.cfg.frame_w), np.float32)
        _ck(lib().pvt_to_gray_f32(self._h, C.byref(f), out.ctypes.data, out.strides[0], MEM_HOST))
        return out

    def profile_enable(self, on=True):
        _ck(lib().pvt_profile_enable(self._h, 1 if on else 0))

    def profile_get(self, reset=True) -> dict:
        p = Profile()
        _ck(lib().pvt_profile_get(self._h, C.byref(p), 1 if reset else 0))
        return {k: getattr(p, k) for k, _ in Profile._fields_}

    def trace_enable(self, on=True):
        _ck(lib().pvt_trace_enable(self._h, 1 if on else 0))

    def trace_get(self, max_steps=64) -> np.ndarray:
        """[steps, 8 kernel slots, 2] globaltimer ns (start of first CTA, end of last CTA); slots: ingest, colprefix,
        rowsum, ncc_search, ncc_finalize, update."""
        out = np.zeros((max_steps, 8, 2), np.uint64)
        n = _ck(lib().pvt_trace_get(self._h, out.ctypes.data, max_steps))
        return out[:n]

    def launch_count(self) -> int:
        return int(lib().pvt_launch_count(self._h))

    def draw_boxes(self, frame, boxes, bgr=None):
        """main.cpp:166: cv::rectangle(frame, bbox, {0,255,0}, 2) for every box, in place.  frame: HxWx3 u8 numpy array or a Frame."""
        f = frame if isinstance(frame, Frame) else host_frame(frame)
        b = np.ascontiguousarray(np.asarray(boxes, np.int32).reshape(-1, 4))
        col = (C.c_uint8 * 3)(*bgr) if bgr is not None else None
        _ck(lib().pvt_draw_boxes(self._h, C.byref(f), len(b), b.ctypes.data_as(C.POINTER(C.c_int32)), col))
        return frame

    def search_kind(self):
        """(name of the search kernel this context's plan runs, kernels per searched step)."""
        buf = C.create_string_buffer(64)
        n = _ck(lib().pvt_search_kind(self._h, buf, 64))
        return buf.value.decode(), n

    def timer_start(self):
        _ck(lib().pvt_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        _ck(lib().pvt_timer_stop(self._h, C.byref(ms)))
        return ms.value


# ---- map-level operators, reference names (tracker/include/baseline_kernel.hpp:8-17) ----------------
def _ncc_match(mode, frame_gray_f32, templ_gray_f32, device=0, formula=FORMULA_CCOEFF_NORMED):
    mode |= MODE_FLAG_EPS if formula == FORMULA_EPS else 0
    f = np.asarray(frame_gray_f32)
    t = np.asarray(templ_gray_f32)
    if f.dtype != np.float32 or t.dtype != np.float32 or f.ndim != 2 or t.ndim != 2:
        raise ValueError("CV_32FC1 inputs required (ncc_cpu.cpp:7-8)")
    if f.strides[1] != 4:
        f = np.ascontiguousarray(f)
    if t.strides[1] != 4:
        t = np.ascontiguousarray(t)
    if f.shape[1] < t.shape[1] or f.shape[0] < t.shape[0]:
        raise ValueError("frame smaller than template (ncc_cpu.cpp:9-10)")
    out = np.empty((f.shape[0] - t.shape[0] + 1, f.shape[1] - t.shape[1] + 1), np.float32)
    _ck(lib().pvt_ncc_match(device, mode, f.ctypes.data, f.shape[1], f.shape[0], f.strides[0], t.ctypes.data, t.shape[1], t.shape[0],
                            t.strides[0], out.ctypes.data, out.strides[0]))
    return out


def ncc_match_naive_cuda(frame_gray_f32, templ_gray_f32, device=0, formula=FORMULA_CCOEFF_NORMED):
    return _ncc_match(MODE_NAIVE, frame_gray_f32, templ_gray_f32, device, formula)


def ncc_match_shared_cuda(frame_gray_f32, templ_gray_f32, device=0, formula=FORMULA_CCOEFF_NORMED):
    return _ncc_match(MODE_SHARED, frame_gray_f32, templ_gray_f32, device, formula)


def ncc_match_const(frame_gray_f32, templ_gray_f32, device=0, formula=FORMULA_CCOEFF_NORMED):
    return _ncc_match(MODE_CONST, frame_gray_f32, templ_gray_f32, device, formula)


def ncc_match_const_tiled(frame_gray_f32, templ_gray_f32, device=0, formula=FORMULA_CCOEFF_NORMED):
    return _ncc_match(MODE_CONST_TILED, frame_gray_f32, templ_gray_f32, device, formula)


def ncc_match_cpu(frame_gray_f32, templ_gray_f32, device=0):
    """The reference's CPU operator is not part of this library: raises PvtError(ERR_UNSUPPORTED)."""
    return _ncc_match(MODE_CPU, frame_gray_f32, templ_gray_f32, device)


def ncc_match_naive_cuda_batched(frames_gray_f32, templ_gray_f32, device=0, formula=FORMULA_CCOEFF_NORMED):
    frames = [np.ascontiguousarray(f, np.float32) for f in frames_gray_f32]
    if not frames:
        raise ValueError("empty batch (baseline_kernel.cu:412)")
    t = np.ascontiguousarray(templ_gray_f32, np.float32)
    fh, fw = frames[0].shape
    if any(f.shape != (fh, fw) for f in frames):
        raise ValueError("all frames must share one geometry (baseline_kernel.cu:419-422)")
    outs = [np.empty((fh - t.shape[0] + 1, fw - t.shape[1] + 1), np.float32) for _ in frames]
    fp = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
    op = (C.c_void_p * len(frames))(*[o.ctypes.data for o in outs])
    _ck(lib().pvt_ncc_match_batched_f(device, formula, len(frames), fp, fw, fh, fw * 4, t.ctypes.data, t.shape[1], t.shape[0], t.strides[0],
                                      op, outs[0].strides[0]))
    return outs


def track_clip(frames, roi, device=0, **params):
    """Convenience twin of the reference main loop for one clip: returns records[n-1] (RESULT_DTYPE) and the final template."""
    n, H, W, _ = frames.shape
    x, y, w, h = roi
    with Tracker(W, H, w, h, device=device, **params) as tr:
        tr.init_track(0, frames[0], roi)
        recs = [tr.step([frames[k]])[0] for k in range(1, n)]
        _, templ = tr.get_state(0)
    return np.array(recs, RESULT_DTYPE), templ

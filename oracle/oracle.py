"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/ncc_oracle.c (see its header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ncc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


class Record(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int), ("conf", C.c_float),
                ("moved", C.c_int), ("updated", C.c_int), ("searched", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def num_threads() -> int:
    return lib().orc_num_threads()


def bgr2gray(bgr: np.ndarray) -> np.ndarray:
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_bgr2gray(_p(bgr), C.c_int(w), C.c_int(h), C.c_size_t(w * 3), _p(out), C.c_size_t(w))
    return out


def gray_to_f32(gray: np.ndarray) -> np.ndarray:
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    out = np.empty((h, w), np.float32)
    lib().orc_gray_to_f32(_p(gray), C.c_int(w), C.c_int(h), C.c_size_t(w), _p(out), C.c_size_t(w * 4))
    return out


def to_gray_f32(bgr: np.ndarray) -> np.ndarray:
    return gray_to_f32(bgr2gray(bgr))


def mean_stddev(t: np.ndarray):
    t = np.ascontiguousarray(t, np.float32)
    m, s = C.c_double(), C.c_double()
    lib().orc_mean_stddev(_p(t), C.c_int(t.shape[1]), C.c_int(t.shape[0]), C.c_size_t(t.shape[1] * 4), C.byref(m), C.byref(s))
    return m.value, s.value


def search_window(x, y, w, h, outW, outH, rx=80, ry=80):
    win = (C.c_int * 4)()
    lib().orc_search_window(C.c_int(x), C.c_int(y), C.c_int(w), C.c_int(h), C.c_int(outW), C.c_int(outH),
                            C.c_int(rx), C.c_int(ry), win)
    return tuple(win)


def ncc_window(frame: np.ndarray, templ: np.ndarray, x0, y0, ww, wh) -> np.ndarray:
    frame = np.ascontiguousarray(frame, np.float32)
    templ = np.ascontiguousarray(templ, np.float32)
    out = np.empty((wh, ww), np.float32)
    rc = lib().orc_ncc_window(_p(frame), C.c_int(frame.shape[1]), C.c_int(frame.shape[0]), C.c_size_t(frame.shape[1] * 4),
                              _p(templ), C.c_int(templ.shape[1]), C.c_int(templ.shape[0]), C.c_size_t(templ.shape[1] * 4),
                              C.c_int(x0), C.c_int(y0), C.c_int(ww), C.c_int(wh), _p(out), C.c_size_t(ww * 4))
    if rc:
        raise ValueError(f"orc_ncc_window rc={rc}")
    return out


def ncc_window_eps(frame: np.ndarray, templ: np.ndarray, x0, y0, ww, wh) -> np.ndarray:
    """The formula of the reference's CUDA kernels (baseline_kernel.cu:21-64), FP32 sequential like the kernel."""
    frame = np.ascontiguousarray(frame, np.float32)
    templ = np.ascontiguousarray(templ, np.float32)
    out = np.empty((wh, ww), np.float32)
    rc = lib().orc_ncc_window_eps(_p(frame), C.c_int(frame.shape[1]), C.c_int(frame.shape[0]), C.c_size_t(frame.shape[1] * 4),
                                  _p(templ), C.c_int(templ.shape[1]), C.c_int(templ.shape[0]), C.c_size_t(templ.shape[1] * 4),
                                  C.c_int(x0), C.c_int(y0), C.c_int(ww), C.c_int(wh), _p(out), C.c_size_t(ww * 4))
    if rc:
        raise ValueError(f"orc_ncc_window_eps rc={rc}")
    return out


def ncc_eps_exact(frame: np.ndarray, templ: np.ndarray, x0, y0, ww, wh) -> np.ndarray:
    """The same formula evaluated in float64 (numpy): what the FP32 kernel approximates."""
    f = np.asarray(frame, np.float64)
    t = np.asarray(templ, np.float64)
    th, tw = t.shape
    n = tw * th
    tm = float(np.float32(t.mean()))
    ts = float(np.float32(np.float32(t.std() + float(np.float32(1e-6))) + np.float32(1e-6)))
    win = np.lib.stride_tricks.sliding_window_view(f[y0:y0 + wh + th - 1, x0:x0 + ww + tw - 1], (th, tw))
    mean = win.mean(axis=(2, 3))
    var = (win * win).mean(axis=(2, 3)) - mean * mean
    sd = np.sqrt(np.maximum(var, float(np.float32(1e-6))))
    cov = np.einsum("yxij,ij->yx", win, t - tm) - mean * (t - tm).sum()
    return cov / ((sd + float(np.float32(1e-6))) * ts * n)


class formula:
    """with oracle.formula(1): ...  -- the tracker loops search the eps map (what the reference's GPU modes hand main.cpp)."""

    def __init__(self, f):
        self.f = int(f)

    def __enter__(self):
        lib().orc_set_formula(C.c_int(self.f))

    def __exit__(self, *a):
        lib().orc_set_formula(C.c_int(0))


def ncc_match_cpu(frame: np.ndarray, templ: np.ndarray) -> np.ndarray:
    fh, fw = frame.shape
    th, tw = templ.shape
    return ncc_window(frame, templ, 0, 0, fw - tw + 1, fh - th + 1)


def max_loc(m: np.ndarray):
    m = np.ascontiguousarray(m, np.float32)
    b, x, y = C.c_double(), C.c_int(), C.c_int()
    lib().orc_max_loc(_p(m), C.c_int(m.shape[1]), C.c_int(m.shape[0]), C.c_size_t(m.shape[1] * 4), C.byref(b), C.byref(x), C.byref(y))
    return b.value, x.value, y.value


def add_weighted(templ: np.ndarray, patch: np.ndarray, lr: float = 0.10) -> np.ndarray:
    t = np.ascontiguousarray(templ, np.float32).copy()
    p = np.ascontiguousarray(patch, np.float32)
    lib().orc_add_weighted(_p(t), _p(p), C.c_int(t.size), C.c_double(1 - lr), C.c_double(lr))
    return t


def track_step(gray, templ, bx, by, rx=80, ry=80, min_conf=0.40, strong_conf=0.70, lr=0.10, want_map=False):
    """One tracked frame; templ (contiguous f32) is updated IN PLACE. Returns (record, window, map|None)."""
    gray = np.ascontiguousarray(gray, np.float32)
    assert templ.flags.c_contiguous and templ.dtype == np.float32
    th, tw = templ.shape
    x, y = C.c_int(bx), C.c_int(by)
    rec, win = Record(), (C.c_int * 4)()
    mp = np.empty(((2 * ry + 1), (2 * rx + 1)), np.float32) if want_map else None
    rc = lib().orc_track_step(_p(gray), C.c_int(gray.shape[1]), C.c_int(gray.shape[0]), _p(templ), C.c_int(tw), C.c_int(th),
                              C.byref(x), C.byref(y), C.c_int(rx), C.c_int(ry), C.c_double(min_conf), C.c_double(strong_conf),
                              C.c_double(lr), C.byref(rec), _p(mp) if want_map else None, win)
    if rc:
        raise ValueError(f"orc_track_step rc={rc}")
    w = tuple(win)
    if want_map:
        mp = mp.ravel()[: w[2] * w[3]].reshape(w[3], w[2]).copy()
    return rec, w, mp


def track_clip(frames: np.ndarray, roi, rx=80, ry=80, min_conf=0.40, strong_conf=0.70, lr=0.10, batch=1):
    """Whole clip; returns (records[n-1,8] float64: x y w h conf moved updated searched, final template)."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n, fh, fw, _ = frames.shape
    x, y, tw, th = roi
    recs = (Record * (n - 1))()
    templ = np.empty((th, tw), np.float32)
    rc = lib().orc_track_clip(_p(frames), C.c_int(n), C.c_int(fw), C.c_int(fh), C.c_int(x), C.c_int(y), C.c_int(tw), C.c_int(th),
                              C.c_int(rx), C.c_int(ry), C.c_double(min_conf), C.c_double(strong_conf), C.c_double(lr),
                              C.c_int(batch), recs, _p(templ))
    if rc:
        raise ValueError(f"orc_track_clip rc={rc}")
    out = np.array([(r.x, r.y, r.w, r.h, r.conf, r.moved, r.updated, r.searched) for r in recs], np.float64)
    return out, templ


class RecordGhc(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int), ("conf", C.c_float),
                ("moved", C.c_int), ("updated", C.c_int), ("searched", C.c_int), ("lost_count", C.c_int), ("use_global", C.c_int)]


def track_clip_ghc(frames: np.ndarray, roi, rx=60, ry=60, min_conf=0.40, global_conf=0.60, strong_conf=0.70, lr=0.10,
                   lost_threshold=50):
    """tracker_ghc/src/main.cpp:145-239 over a whole clip; returns (records[n-1,10] float64: x y w h conf moved updated
    searched(1 local, 2 whole map) lost_count use_global, final template)."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n, fh, fw, _ = frames.shape
    x, y, tw, th = roi
    recs = (RecordGhc * (n - 1))()
    templ = np.empty((th, tw), np.float32)
    rc = lib().orc_track_clip_ghc(_p(frames), C.c_int(n), C.c_int(fw), C.c_int(fh), C.c_int(x), C.c_int(y), C.c_int(tw), C.c_int(th),
                                  C.c_int(rx), C.c_int(ry), C.c_double(min_conf), C.c_double(global_conf), C.c_double(strong_conf),
                                  C.c_double(lr), C.c_int(lost_threshold), recs, _p(templ))
    if rc:
        raise ValueError(f"orc_track_clip_ghc rc={rc}")
    out = np.array([(r.x, r.y, r.w, r.h, r.conf, r.moved, r.updated, r.searched, r.lost_count, r.use_global) for r in recs], np.float64)
    return out, templ


def draw_rectangle(img: np.ndarray, box, bgr=(0, 255, 0)) -> np.ndarray:
    """tracker/src/main.cpp:166  cv::rectangle(frame, bbox, {0,255,0}, 2) for a box inside the frame, restated: the 3-pixel band
    around the lines (x0,y0)-(x1,y1), x1 = x+w-1, y1 = y+h-1, minus the four outer corner pixels (pinned to cv2 4.13.0 in
    tests/test_oracle_golden.py).  Paints in place and returns img."""
    H, W = img.shape[:2]
    x, y, w, h = (int(v) for v in box)
    x0, y0, x1, y1 = x, y, x + w - 1, y + h - 1
    ys, xs = np.mgrid[max(0, y0 - 1):min(H, y1 + 2), max(0, x0 - 1):min(W, x1 + 2)]
    hole = (xs >= x0 + 2) & (xs <= x1 - 2) & (ys >= y0 + 2) & (ys <= y1 - 2)
    corner = ((xs == x0 - 1) | (xs == x1 + 1)) & ((ys == y0 - 1) | (ys == y1 + 1))
    m = ~hole & ~corner
    img[ys[m], xs[m]] = np.asarray(bgr, img.dtype)
    return img


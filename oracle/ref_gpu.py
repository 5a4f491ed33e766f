"""TEST / BENCH INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libref_baseline.so: the reference's OWN GPU operators
(/root/reference/tracker/src/baseline_kernel.cu, compiled unmodified by oracle/ref_build/Makefile).  Used to pin
PVT_FORMULA_EPS against the real kernels and as bench.py's `ref_gpu_baseline`.  Needs a GPU to run;
nothing under parallel-video-object-tracker_b200/ imports this."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_baseline.so")
MODES = {"naive": 0, "shared": 2, "const": 3, "const_tiled": 4}
_lib = None


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
        _lib.ref_ncc_match.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.ref_ncc_match_batched.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    return _lib


def ncc_match(mode: str, frame: np.ndarray, templ: np.ndarray) -> np.ndarray:
    """baseline::ncc_match_{naive_cuda, shared_cuda, const, const_tiled}: full (fh-th+1) x (fw-tw+1) map."""
    f = np.ascontiguousarray(frame, np.float32)
    t = np.ascontiguousarray(templ, np.float32)
    out = np.empty((f.shape[0] - t.shape[0] + 1, f.shape[1] - t.shape[1] + 1), np.float32)
    rc = lib().ref_ncc_match(MODES[mode], f.ctypes.data, f.shape[1], f.shape[0], t.ctypes.data, t.shape[1], t.shape[0], out.ctypes.data)
    if rc:
        raise RuntimeError(f"reference operator {mode} failed / asserted (rc={rc})")
    return out


def ncc_match_batched(frames, templ: np.ndarray):
    fs = [np.ascontiguousarray(f, np.float32) for f in frames]
    t = np.ascontiguousarray(templ, np.float32)
    fh, fw = fs[0].shape
    outs = [np.empty((fh - t.shape[0] + 1, fw - t.shape[1] + 1), np.float32) for _ in fs]
    fp = (C.c_void_p * len(fs))(*[f.ctypes.data for f in fs])
    op = (C.c_void_p * len(fs))(*[o.ctypes.data for o in outs])
    rc = lib().ref_ncc_match_batched(len(fs), fp, fw, fh, t.ctypes.data, t.shape[1], t.shape[0], op)
    if rc:
        raise RuntimeError(f"reference batched operator failed (rc={rc})")
    return outs

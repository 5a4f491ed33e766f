"""TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product library.

Call-for-call Python restatement of the reference tracker's `--cpu` loop with
the image's cv2 4.13.0 -- the exact OpenCV version the primary tracker/ links
(opencv_core4130.lib, /root/reference/tracker/Makefile:19-27).  The arithmetic
itself (cvtColor, convertTo, matchTemplate, minMaxLoc, addWeighted) is
OpenCV's; only the driver loop is restated:

    toGrayF32            tracker/include/utils.hpp:5-14
    ncc_match_cpu        tracker/src/ncc_cpu.cpp:5-13  (TM_CCOEFF_NORMED, full frame)
    window clamp         tracker/src/main.cpp:135-146
    minMaxLoc on the ROI tracker/src/main.cpp:147-151
    gates + EMA          tracker/src/main.cpp:153-161
    batch hold semantics tracker/src/main.cpp:115-130  (only the N-th frame is searched)
    lost-object logic    tracker_ghc/src/main.cpp:145-239 (track_clip_ghc: whole-frame re-acquisition)

Uses: (1) generating the golden fixtures in tests/golden/ (make_golden.py),
(2) the `cpu_baseline` / `--impl reference` legs of bench.py, timed on the
host cores with the reference's own `t_tot` convention (main.cpp:101,163-164).
"""
from __future__ import annotations

import time

import numpy as np

try:  # cv2 is present in the build image and on the GPU boxes (same image)
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

# reference parameters, tracker/src/main.cpp:6-20
SEARCH_RADIUS_X = 80
SEARCH_RADIUS_Y = 80
NCC_MIN_CONFIDENCE = 0.40
NCC_STRONG_CONFIDENCE = 0.70
TEMPLATE_UPDATE_LR = 0.10
BATCH_SIZE = 4


def available() -> bool:
    return cv2 is not None


def to_gray_f32(bgr: np.ndarray) -> np.ndarray:
    """tracker/include/utils.hpp:5-14: cvtColor(BGR2GRAY) then convertTo(CV_32F, 1.0f/255.0f).

    Python cv2 does not bind Mat::convertTo; float32(g) * float32(1/255) is the same single
    rounding of the exact product g * 0x1.010102p-8 that convertTo performs (the product of an
    8-bit and a 24-bit significand is exact in double, and exact-then-rounded in the float SIMD
    path), and make_golden.py cross-checks it against cv2.multiply(..., dtype=CV_32F).
    """
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY) if bgr.ndim == 3 else bgr
    return gray.astype(np.float32) * (np.float32(1.0) / np.float32(255.0))


def search_window(x, y, w, h, outW, outH, rx=SEARCH_RADIUS_X, ry=SEARCH_RADIUS_Y):
    """tracker/src/main.cpp:135-146, C `int` arithmetic ( / truncates; all operands >= 0 here)."""
    cx = x + w // 2
    cy = y + h // 2
    minTx = max(0, cx - rx - w // 2)
    maxTx = min(outW - 1, cx + rx - w // 2)
    minTy = max(0, cy - ry - h // 2)
    maxTy = min(outH - 1, cy + ry - h // 2)
    return minTx, minTy, maxTx - minTx + 1, maxTy - minTy + 1


def track_clip(frames, roi, rx=SEARCH_RADIUS_X, ry=SEARCH_RADIUS_Y,
               min_conf=NCC_MIN_CONFIDENCE, strong_conf=NCC_STRONG_CONFIDENCE,
               lr=TEMPLATE_UPDATE_LR, batch=1, keep_maps=(), timing=None):
    """Run the reference `--cpu` loop over frames[1:], template cut from frames[0] at roi.

    Returns dict(records=[n-1, 7] (x, y, w, h, conf, moved, updated) float64,
                 maps={frame_index: window map f32}, templ=final template f32).
    batch>1 reproduces main.cpp:115-130: only every batch-th frame is searched; the
    others keep the stale box (conf reported as NaN, moved=updated=0).
    """
    x, y, w, h = roi
    g = to_gray_f32(frames[0])
    templ = g[y:y + h, x:x + w].copy()
    recs, maps = [], {}
    t_tot = 0.0
    t_gray = 0.0
    pending = 0
    for k in range(1, len(frames)):
        t0 = time.perf_counter()
        g = to_gray_f32(frames[k])
        t_gray += time.perf_counter() - t0
        if batch > 1:
            pending += 1
            if pending < batch:
                recs.append((x, y, w, h, np.nan, 0, 0))
                continue
            pending = 0
        t1 = time.perf_counter()
        ncc = cv2.matchTemplate(g, templ, cv2.TM_CCOEFF_NORMED)
        outH, outW = ncc.shape
        minTx, minTy, ww, wh = search_window(x, y, w, h, outW, outH, rx, ry)
        view = ncc[minTy:minTy + wh, minTx:minTx + ww]
        _, bestVal, _, bestLoc = cv2.minMaxLoc(view)
        bx, by = bestLoc[0] + minTx, bestLoc[1] + minTy
        moved = updated = 0
        if bestVal >= min_conf:
            x, y = bx, by
            moved = 1
            if bestVal >= strong_conf:
                patch = g[y:y + h, x:x + w].copy()
                cv2.addWeighted(templ, 1 - lr, patch, lr, 0.0, templ)
                updated = 1
        t_tot += time.perf_counter() - t1
        if k in keep_maps:
            maps[k] = view.copy()
        recs.append((x, y, w, h, bestVal, moved, updated))
    if timing is not None:
        timing["t_tot"] = t_tot          # NCC + peak + update (main.cpp:101,163-164)
        timing["t_gray"] = t_gray        # toGrayF32, outside t_tot in the reference
        timing["frames"] = len(frames) - 1
    return {"records": np.array(recs, dtype=np.float64), "maps": maps, "templ": templ}


# tracker_ghc/src/main.cpp:9-23
GHC_SEARCH_RADIUS = 60
NCC_GLOBAL_CONFIDENCE = 0.60
LOST_FRAME_THRESHOLD = 50


def _bbox_outside(x, y, w, h, fw, fh):
    """tracker_ghc/src/main.cpp:49-55 isBboxOutsideFrame."""
    cx, cy = x + w // 2, y + h // 2
    return (cx < 0 or cx >= fw or cy < 0 or cy >= fh) or (x + w < 0 or x >= fw or y + h < 0 or y >= fh)


def track_clip_ghc(frames, roi, rx=GHC_SEARCH_RADIUS, ry=GHC_SEARCH_RADIUS, min_conf=NCC_MIN_CONFIDENCE,
                   global_conf=NCC_GLOBAL_CONFIDENCE, strong_conf=NCC_STRONG_CONFIDENCE, lr=TEMPLATE_UPDATE_LR,
                   lost_threshold=LOST_FRAME_THRESHOLD, start_box=None, lost0=0, use_global0=False):
    """tracker_ghc/src/main.cpp:145-239 (mode "cpu") with cv2: full-frame matchTemplate, then either the local window
    or the whole map.  records = [n-1, 10]: x y w h conf moved updated searched(1 local, 2 whole map) lost_count use_global.
    start_box / lost0 / use_global0: resume from a checkpointed state (the loop's variables :143-144 and bbox after the
    template has been cut from frames[0] at roi); the defaults are the reference's initial state."""
    x, y, w, h = roi
    g = to_gray_f32(frames[0])
    templ = g[y:y + h, x:x + w].copy()
    lost, use_global, recs = int(lost0), bool(use_global0), []
    if start_box is not None:
        x, y = int(start_box[0]), int(start_box[1])
    for k in range(1, len(frames)):
        g = to_gray_f32(frames[k])
        fh, fw = g.shape
        ncc = cv2.matchTemplate(g, templ, cv2.TM_CCOEFF_NORMED)
        outH, outW = ncc.shape
        if _bbox_outside(x, y, w, h, fw, fh) or lost >= lost_threshold:
            use_global = True
        if use_global:
            _, bestVal, _, bestLoc = cv2.minMaxLoc(ncc)
            bx, by = bestLoc
        else:
            minTx, minTy, ww, wh = search_window(x, y, w, h, outW, outH, rx, ry)
            if ww > 0 and wh > 0:
                _, bestVal, _, bestLoc = cv2.minMaxLoc(ncc[minTy:minTy + wh, minTx:minTx + ww])
                bx, by = bestLoc[0] + minTx, bestLoc[1] + minTy
            else:
                _, bestVal, _, bestLoc = cv2.minMaxLoc(ncc)
                bx, by = bestLoc
        searched = 2 if use_global else 1
        thr = global_conf if use_global else min_conf
        moved = updated = 0
        if bestVal >= thr:
            x, y = bx, by
            moved = 1
            lost = 0
            if not _bbox_outside(x, y, w, h, fw, fh):
                use_global = False
            if bestVal >= strong_conf:
                patch = g[y:y + h, x:x + w].copy()
                cv2.addWeighted(templ, 1 - lr, patch, lr, 0.0, templ)
                updated = 1
        else:
            lost += 1
        recs.append((x, y, w, h, bestVal, moved, updated, searched, lost, int(use_global)))
    return {"records": np.array(recs, dtype=np.float64), "templ": templ}

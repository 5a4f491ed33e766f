"""Frame source and sink of the reference loop, around the GPU hot path.

tracker/src/main.cpp:52-60   cv::VideoCapture cap(INPUT_VIDEO); "Cannot open video."; the first frame carries the ROI
tracker/src/main.cpp:73-82   fps = CAP_PROP_FPS (30 if <= 1); cv::VideoWriter(out, fourcc('m','p','4','v'), fps, frame size)
tracker/src/main.cpp:93-97   cap >> frame until empty
tracker/src/main.cpp:166-167 cv::rectangle(frame, bbox, {0,255,0}, 2); writer.write(frame)

Decode and encode are OpenCV videoio (FFmpeg) on the host, exactly as in the reference -- that is I/O, not the path; what sits
between them is the library: every decoded BGR8 frame goes through pvt_step (ingest -> NCC search -> peak -> gates -> EMA on the
GPU) and the box is painted by pvt_draw_boxes (k_overlay, cv::rectangle's thickness-2 coverage bit for bit) before the frame is
handed to the encoder.  Only BGR8 / GRAY8 sources are parity-safe (an NV12 luma plane is not cvtColor(BGR2GRAY)), so frames are
taken as cap.read() delivers them.  tracker_ghc's loop (tracker_ghc/src/main.cpp:83-107, 147, 241-247) is the same with
lost_frame_threshold > 0.

No CPU fallback: the Tracker raises when libpvt.so / a CUDA device is missing.  cv2 is imported lazily and only here.
"""
from __future__ import annotations

import time

import numpy as np

from . import MODE_BATCH, MODE_CONST, MODE_CONST_TILED, MODE_NAIVE, MODE_SHARED, RESULT_DTYPE, Tracker


def _cv2():
    try:
        import cv2
    except ImportError as e:   # the reference needs OpenCV for the same job
        raise ImportError("track_video needs OpenCV's videoio (cv2) to decode / encode, as the reference does") from e
    return cv2


def track_video(video_in, roi, video_out=None, *, fourcc="mp4v", max_frames=None, draw=True, **params):
    """The reference main loop on a video file.  roi = (x, y, w, h) on the FIRST frame (the reference asks cv::selectROI).
    Returns (records[n-1] RESULT_DTYPE, final template, summary dict with the reference's summary fields).
    Errors follow the reference: IOError("Cannot open video.") / ("Cannot open video writer."), ValueError("No ROI selected.")."""
    cv2 = _cv2()
    cap = cv2.VideoCapture(str(video_in))
    if not cap.isOpened():
        raise IOError("Cannot open video.")                                    # main.cpp:53-56
    ok, frame = cap.read()
    if not ok or frame is None:
        raise IOError("Cannot open video.")                                    # main.cpp:59-60 (empty first frame)
    x, y, w, h = (int(v) for v in roi)
    if w == 0 or h == 0:
        raise ValueError("No ROI selected.")                                   # main.cpp:66-69
    H, W = frame.shape[:2]
    fps_video = cap.get(cv2.CAP_PROP_FPS)
    if not fps_video or fps_video <= 1:
        fps_video = 30                                                         # main.cpp:73-74
    writer = None
    if video_out is not None:
        writer = cv2.VideoWriter(str(video_out), cv2.VideoWriter_fourcc(*fourcc), fps_video, (W, H))   # main.cpp:76-77
        if not writer.isOpened():
            raise IOError("Cannot open video writer.")                         # main.cpp:79-82
    recs = []
    t_tot = 0.0
    t_start = time.perf_counter()
    try:
        with Tracker(W, H, w, h, **params) as tr:
            tr.init_track(0, np.ascontiguousarray(frame), (x, y, w, h))        # main.cpp:70-71
            while max_frames is None or len(recs) < max_frames:
                ok, frame = cap.read()                                         # main.cpp:95-96
                if not ok or frame is None:
                    break
                frame = np.ascontiguousarray(frame)
                t1 = time.perf_counter()
                r = tr.step([frame])[0]                                        # main.cpp:98-161 on the GPU
                t_tot += time.perf_counter() - t1
                recs.append(r)
                if writer is not None:
                    if draw:
                        tr.draw_boxes(frame, [(int(r["x"]), int(r["y"]), int(r["w"]), int(r["h"]))])   # main.cpp:166
                    writer.write(frame)                                        # main.cpp:167
            _, templ = tr.get_state(0)
    finally:
        cap.release()
        if writer is not None:
            writer.release()
    elapsed = time.perf_counter() - t_start
    n = len(recs)
    summary = {"frames": n, "time_s": elapsed, "computation_time_s": t_tot, "fps": n / elapsed if elapsed > 0 else 0.0,
               "fps_video": float(fps_video), "frame_size": (W, H)}           # main.cpp:171-182
    return np.array(recs, RESULT_DTYPE), templ, summary


def main(argv=None) -> int:
    """Command line of the reference (tracker/src/main.cpp:23-49, 171-182) on a video file:
    python -m parallel-video-object-tracker_b200.video VIDEO --roi x,y,w,h [--out OUT.mp4] [--shared|--const|--const_tiled|--batch=N]
    The mode flags select what they select in the reference build of this library: nothing but the hold semantics of --batch=N
    (every mode runs the same kernels); --cpu is rejected (the library has no CPU path).  The ROI comes from --roi instead of
    cv::selectROI (no display here)."""
    import sys
    argv = list(sys.argv[1:] if argv is None else argv)
    mode, batch, roi, out, path = "naive", 4, None, None, None
    i = 0
    while i < len(argv):
        a = argv[i]
        if a == "--cpu": mode = "cpu"
        elif a == "--shared": mode = "shared"
        elif a == "--const": mode = "const"
        elif a == "--const_tiled": mode = "const_tiled"
        elif a.startswith("--batch="):
            mode = "batch"
            try:
                batch = max(1, int(a[8:]))
            except ValueError:
                batch = 1                                                      # std::atoi of a non-number is 0 -> max(1, 0)
        elif a == "--roi" and i + 1 < len(argv):
            i += 1
            try:
                roi = tuple(int(v) for v in argv[i].split(","))
            except ValueError:
                roi = None
        elif a == "--out" and i + 1 < len(argv):
            i += 1
            out = argv[i]
        elif not a.startswith("--") and path is None:
            path = a
        i += 1
    print("--------\nNCC Tracker Starting\nInput video : %s\nMode        : %s" % (path, mode))
    if mode == "batch":
        print("Batch size  : %d" % batch)
    print("--------\n")
    if mode == "cpu":
        print("--cpu is not available: this library has no CPU path", file=sys.stderr)
        return -1
    if path is None:
        print("Cannot open video.", file=sys.stderr)
        return -1
    if roi is None or len(roi) != 4:
        roi = (0, 0, 0, 0)                                                     # "No ROI selected." -- after the video has been opened, as in the reference
    kw = {"mode": {"naive": MODE_NAIVE, "shared": MODE_SHARED, "const": MODE_CONST, "const_tiled": MODE_CONST_TILED, "batch": MODE_BATCH}[mode]}
    if mode == "batch":
        kw["batch_size"] = batch
    try:
        recs, _, s = track_video(path, roi, out, **kw)
    except IOError as e:
        print(str(e), file=sys.stderr)
        return -1
    except ValueError as e:
        print(" " + str(e), file=sys.stderr)
        return -1
    print("\n--------\n Tracking Complete\n Mode       : %s\n Frames     : %d\n Time (sec) : %g\n Computation Time (sec)  : %g\n FPS        : %g\n--------"
          % (mode, s["frames"], s["time_s"], s["computation_time_s"], s["fps"]))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())

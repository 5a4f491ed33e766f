"""Measurement / QA tool: a short pass over every kernel path (local K-split, throughput + fringe via PVT_PLAN, lost-object
whole-frame pass) for compute-sanitizer.   compute-sanitizer --tool memcheck python tools/sanitize_probe.py"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools import synth
pvt = importlib.import_module("parallel-video-object-tracker_b200")

def run(tag, clip, n_tracks=1, frames=6, **kw):
    fr, roi = clip["frames"], clip["roi"]
    n, H, W, _ = fr.shape
    with pvt.Tracker(W, H, roi[2], roi[3], max_tracks=n_tracks, **kw) as tr:
        for t in range(n_tracks):
            tr.init_track(t, fr[0] if t == 0 else None, roi if t == 0 else (4 + 3 * t, 5 + 2 * t, roi[2], roi[3]))
        for k in range(1, min(n, frames)):
            r = tr.step([fr[k]])
        print(tag, "ok", r[0]["x"], r[0]["y"], float(r[0]["conf"]), flush=True)

small = synth.make_clip(seed=1, W=320, H=240, tw=32, th=32, n_frames=8, R=40)
odd = synth.make_clip(seed=7, W=301, H=233, tw=37, th=29, n_frames=8, R=40)
run("local+winstats", small, search_radius_x=40, search_radius_y=40)            # round 2 default for one track: k_ncc_local beside k_winstats
os.environ["PVT_LOCAL_STATS"] = "1"; os.environ["PVT_LOCAL_UPDATE"] = "1"
run("local, statistics + update in the CTA", odd, search_radius_x=40, search_radius_y=24, keep_maps=1)
del os.environ["PVT_LOCAL_STATS"], os.environ["PVT_LOCAL_UPDATE"]
os.environ["PVT_NO_LOCAL"] = "1"
run("ksplit", small, search_radius_x=40, search_radius_y=40)
os.environ["PVT_FUSED"] = "1"
run("fused", small, search_radius_x=40, search_radius_y=40)
del os.environ["PVT_FUSED"], os.environ["PVT_NO_LOCAL"]
run("tensor-core", small, n_tracks=3, search_radius_x=40, search_radius_y=40, kernel=pvt.KERNEL_TC)
run("tensor-core-odd", odd, search_radius_x=40, search_radius_y=24, keep_maps=1, kernel=pvt.KERNEL_TC)
os.environ["PVT_STATS_LEGACY"] = "1"
run("legacy statistics", small, n_tracks=2, search_radius_x=40, search_radius_y=40)
del os.environ["PVT_STATS_LEGACY"]
run("ksplit-odd", odd, search_radius_x=40, search_radius_y=24, keep_maps=1)
os.environ["PVT_PLAN"] = "16,1,1,1"
run("unsplit+fringe", small, n_tracks=6, search_radius_x=40, search_radius_y=40)
os.environ["PVT_PLAN"] = "16,2,2,1"
run("deferred-fringe", small, n_tracks=3, search_radius_x=40, search_radius_y=40)
del os.environ["PVT_PLAN"]
lost = synth.make_clip(seed=31, W=320, H=240, tw=32, th=32, n_frames=30, R=40, variant="lost")
run("lost-mode", lost, frames=30, search_radius_x=12, search_radius_y=12, lost_frame_threshold=4)
f = pvt.ncc_match_naive_cuda(np.random.default_rng(0).random((80, 96), np.float32), np.random.default_rng(1).random((13, 17), np.float32))
print("map", f.shape)
with pvt.Tracker(320, 240, 8, 8) as tr:
    img = np.zeros((240, 320, 3), np.uint8)
    tr.draw_boxes(img, [(0, 0, 8, 8), (300, 220, 20, 20), (100, 50, 1, 1)])
    print("overlay", int(img.sum()))

"""k_ncc_tc over column tiles on the 1080p whole map, as a LOCAL window (search radius = the frame): the same kernel, geometry and
CTA count as the whole-frame pass of the lost-object mode, in a graph without a conditional node (ncu cannot profile kernel
nodes of graphs that hold one).  python tools/wf_probe_local.py [steps]"""
import importlib
import sys
import time

import numpy as np

sys.path.insert(0, "/root/repo")
from tools import synth

pvt = importlib.import_module("parallel-video-object-tracker_b200")
W, H, TW, TH = 1920, 1080, 64, 64
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
sc = synth.Scene(synth.ClipSpec(seed=41, W=W, H=H, tw=TW, th=TH, n_frames=4, R=80, period=32))
frames = [sc.frame(k) for k in range(4)]
with pvt.Tracker(W, H, TW, TH, search_radius_x=W, search_radius_y=H, kernel=pvt.KERNEL_TC) as tr:
    tr.init_track(0, frames[0], (*sc.obj_pos(0), TW, TH))
    for k in range(4):
        tr.step([frames[1 + k % 3]])
    t0 = time.perf_counter()
    for k in range(steps):
        r = tr.step([frames[1 + k % 3]])[0]
    dt = (time.perf_counter() - t0) / steps
    print("search kind", tr.search_kind(), "| host-timed ms per step (incl. the frame's H2D copy):", round(1e3 * dt, 4), "| conf", float(r["conf"]))

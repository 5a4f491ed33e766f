import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import helpers as Hp
pvt = importlib.import_module("parallel-video-object-tracker_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "lost"
(c, tk) = Hp.clip(name)
frames, roi = c["frames"], c["roi"]
g = Hp.golden(f"clip_{name}.npz")["records"]
H, W = frames.shape[1:3]
with pvt.Tracker(W, H, roi[2], roi[3], search_radius_x=tk.get("rx", 80), search_radius_y=tk.get("ry", 80)) as tr:
    tr.init_track(0, frames[0], roi)
    for k in range(1, len(frames)):
        r = tr.step([frames[k]])[0]
        print(k, r, g[k - 1][:5], flush=True)

"""Measurement tool: the whole-frame (lost-object) pass on one 1080p stream, for an ncu launch list."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
pvt = importlib.import_module("parallel-video-object-tracker_b200")
wl = dict(bench.WORKLOADS["C2"])
W, H, tw, th, L = wl["W"], wl["H"], wl["tw"], wl["th"], wl["ring"]
scenes, host, dev = bench.build_rings(wl, 0, torch)
ring = bench.ring_descs(pvt, wl, dev, True)
tr = pvt.Tracker(W, H, tw, th, search_radius_x=80, search_radius_y=80, lost_frame_threshold=50, ncc_global_confidence=2.0)
tr.init_track(0, pvt.device_frame(dev[0, 0].data_ptr(), W * 3, stream=0), bench.rois_for(wl, scenes[0])[0], stream=0)
tr.set_lost_state(0, 1000, 1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
tr.submit_sequence(n, ring[1:] + ring[:1]); tr.sync()
tr.timer_start(); tr.submit_sequence(n, ring[1:] + ring[:1]); print("ms/step", tr.timer_stop() / n)
tr.close()
# the same stream with nothing lost: what lost-object mode costs a step that only runs the local pass
for thr in (0, 50):
    t2 = pvt.Tracker(W, H, tw, th, search_radius_x=80, search_radius_y=80, lost_frame_threshold=thr)
    t2.init_track(0, pvt.device_frame(dev[0, 0].data_ptr(), W * 3, stream=0), bench.rois_for(wl, scenes[0])[0], stream=0)
    t2.submit_sequence(64, ring[1:] + ring[:1]); t2.sync()
    t2.timer_start(); t2.submit_sequence(400, ring[1:] + ring[:1]); ms = t2.timer_stop()
    print("lost_frame_threshold", thr, "us/step", 1e3 * ms / 400, "last", t2.collect(1)[0][0]["conf"])
    t2.close()

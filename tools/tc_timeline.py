"""Measurement tool: phase stamps of k_ncc_tc's CTA 0 (device globaltimer, warm, no profiler): prologue, tile landed,
MMA issue loop, all MMAs complete, epilogue.   usage: python tools/tc_timeline.py [C5|C4|C2]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
pvt = importlib.import_module("parallel-video-object-tracker_b200")
wname = sys.argv[1] if len(sys.argv) > 1 else "C5"
wl = dict(bench.WORKLOADS[wname])
W, H, tw, th, R, L, S = wl["W"], wl["H"], wl["tw"], wl["th"], wl["R"], wl["ring"], wl["streams"]
scenes, host, dev = bench.build_rings(wl, 0, torch, want_host=False)
ring = bench.ring_descs(pvt, wl, dev, True)
tr = pvt.Tracker(W, H, tw, th, max_streams=S, max_tracks=S * wl["rois"], search_radius_x=R, search_radius_y=R, kernel=pvt.KERNEL_TC)
t = 0
for s in range(S):
    for j, roi in enumerate(bench.rois_for(wl, scenes[s % len(scenes)])):
        tr.init_track(t, pvt.device_frame(dev[s, 0].data_ptr(), W * 3, stream=s) if j == 0 else None, roi, stream=s); t += 1
sh = lambda st: ring[st % L:] + ring[:st % L]
tr.trace_enable(True)
tr.submit_sequence(32, sh(1)); tr.sync()
tr.timer_start(); tr.submit_sequence(32, sh(33)); ms = tr.timer_stop()
T = tr.trace_get(32).astype(np.int64)[4:]
b = T[:, 3, 0]                                           # k_ncc_tc: first CTA start
med = lambda x: float(np.median(x - b)) / 1e3
print("%s PVT_KERNEL_TC: %.2f us/step (events); k_ncc_tc first start .. last end %.2f us" % (wname, 1e3 * ms / 32, med(T[:, 3, 1])))
print("  CTA 0: prologue done +%.2f | tile landed +%.2f | MMAs all issued +%.2f | MMAs complete +%.2f | epilogue done +%.2f us" %
      (med(T[:, 4, 0]), med(T[:, 6, 0]), med(T[:, 4, 1]), med(T[:, 7, 0]), med(T[:, 7, 1])))

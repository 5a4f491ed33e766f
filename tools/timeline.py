"""Measurement tool: warm device-side timeline of one workload's step (globaltimer stamps, no profiler attached)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
pvt = importlib.import_module("parallel-video-object-tracker_b200")
wname = sys.argv[1] if len(sys.argv) > 1 else "C2"
wl = dict(bench.WORKLOADS[wname])
W, H, tw, th, R, L, S = wl["W"], wl["H"], wl["tw"], wl["th"], wl["R"], wl["ring"], wl["streams"]
scenes, host, dev = bench.build_rings(wl, 0, torch)
on_host = len(sys.argv) > 2 and sys.argv[2] == "host"   # pinned host ring: the ROI ingest reads the tiles zero-copy over PCIe
ring = bench.ring_descs(pvt, wl, host if on_host else dev, not on_host)
tr = pvt.Tracker(W, H, tw, th, max_streams=S, max_tracks=S * wl["rois"], search_radius_x=R, search_radius_y=R)
t = 0
for s in range(S):
    for j, roi in enumerate(bench.rois_for(wl, scenes[s % len(scenes)])):
        tr.init_track(t, pvt.device_frame(dev[s, 0].data_ptr(), W * 3, stream=s) if j == 0 else None, roi, stream=s); t += 1
sh = lambda st: ring[st % L:] + ring[:st % L]
tr.trace_enable(True)
tr.submit_sequence(64, sh(1)); tr.sync()
tr.timer_start(); tr.submit_sequence(64, sh(65)); ms = tr.timer_stop()
T = tr.trace_get(64).astype(np.int64)
names = ["ingest", "colprefix", "rowsum", "ncc_search", "ncc_finalize", "update", "ncc_fringe", "ncc_tail_finalize"]
T = T[8:]                                     # skip the first steps
t0 = T[:, 0, 0:1]
print("%s: %.2f us/step (events); step-to-step %.2f us (trace)" % (wname, 1e3 * ms / 64, np.median(np.diff(T[:, 0, 0])) / 1e3))
prev_end = None
local = T[:, 7, 0].any() and T[:, 6, 0].any() and not T[:, 4, 0].any()
for k, nm in enumerate(names):
    if not T[:, k, 0].any() or (local and k in (1, 2, 6, 7)): continue
    st = np.median(T[:, k, 0] - T[:, 0, 0]) / 1e3; en = np.median(T[:, k, 1] - T[:, 0, 0]) / 1e3
    print("  %-13s start %7.2f  end %7.2f  dur %6.2f us  gap-before %6.2f" % (nm, st, en, en - st, st - prev_end if prev_end is not None else 0.0))
    prev_end = en
if T[:, 7, 0].any() and T[:, 6, 0].any() and not T[:, 4, 0].any():
    b = T[:, 3, 0]
    print("  k_ncc_local, CTA 0 (from the kernel's first CTA start): staged +%.2f | column partials +%.2f | column slides +%.2f | statistics done +%.2f | FMA loop done +%.2f | reduced, peak sent +%.2f us" %
          tuple(np.median(x - b) / 1e3 for x in (T[:, 6, 0], T[:, 1, 0], T[:, 1, 1], T[:, 7, 0], T[:, 6, 1], T[:, 7, 1])))
elif T[:, 7, 0].any() and T[:, 6, 0].any():
    b = T[:, 3, 0]
    print("  K-split search, CTA 0 (from the kernel's first CTA start): tile requested +%.2f | landed +%.2f | loop done +%.2f | partial sums stored +%.2f us" %
          tuple(np.median(x - b) / 1e3 for x in (T[:, 6, 0], T[:, 6, 1], T[:, 7, 0], T[:, 7, 1])))
if T[:, 2, 0].any() and local:
    print("  k_prefetch_roi (statistics branch): start +%.2f  end +%.2f us from the step's start" % (np.median(T[:, 2, 0] - T[:, 0, 0]) / 1e3, np.median(T[:, 2, 1] - T[:, 0, 0]) / 1e3))
nxt = np.median(T[1:, 0, 0] - T[:-1, 5, 1]) / 1e3
print("  gap to next step's ingest: %.2f us" % nxt)

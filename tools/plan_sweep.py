"""Measurement tool: time one workload under several k_ncc_search plans (PVT_PLAN=GB,pj,pd) in one process."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
pvt = importlib.import_module("parallel-video-object-tracker_b200")
wname = sys.argv[1] if len(sys.argv) > 1 else "C2"
plans = sys.argv[2:] or [""]
wl = dict(bench.WORKLOADS[wname])
W, H, tw, th, R, L, S = wl["W"], wl["H"], wl["tw"], wl["th"], wl["R"], wl["ring"], wl["streams"]
scenes, host, dev = bench.build_rings(wl, 0, torch)
ring = bench.ring_descs(pvt, wl, dev, True)
n_tracks = S * wl["rois"]
K = int(os.environ.get("SWEEP_STEPS", "200"))
for plan in plans:
    if plan: os.environ["PVT_PLAN"] = plan
    else: os.environ.pop("PVT_PLAN", None)
    tr = pvt.Tracker(W, H, tw, th, max_streams=S, max_tracks=n_tracks, search_radius_x=R, search_radius_y=R)
    t = 0
    for s in range(S):
        for j, roi in enumerate(bench.rois_for(wl, scenes[s % len(scenes)])):
            tr.init_track(t, pvt.device_frame(dev[s, 0].data_ptr(), W * 3, stream=s) if j == 0 else None, roi, stream=s); t += 1
    sh = lambda st: ring[st % L:] + ring[:st % L]
    tr.submit_sequence(16, sh(1)); tr.sync()
    tr.timer_start(); tr.submit_sequence(K, sh(17)); ms = tr.timer_stop()
    last = tr.collect(1)[0][0]
    tx, ty = scenes[0].obj_pos((16 + K) % L)
    tr.profile_enable(True); tr.profile_get(True); tr.submit_sequence(50, sh(17 + K)); p = tr.profile_get(True); tr.profile_enable(False)
    print("plan=%-10s ms/step %.4f  ok=%s | ingest %.1f stats %.1f ncc %.1f update %.1f us" % (plan or "auto", ms / K, bool(last["x"] == tx and last["y"] == ty),
          1e3 * p["ingest_ms"] / 50, 1e3 * p["stats_ms"] / 50, 1e3 * p["ncc_ms"] / 50, 1e3 * p["update_ms"] / 50), flush=True)
    tr.close()

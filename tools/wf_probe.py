"""Whole-frame search (lost-object mode) step times: FP32 plan, PVT_KERNEL_TC, PVT_KERNEL_TC_GLOBAL.  python tools/wf_probe.py [fp32|tc|tc_global ...]"""
import importlib, json, sys, torch
sys.path.insert(0, "/root/repo")
import bench
pvt = importlib.import_module("parallel-video-object-tracker_b200")
torch.cuda.set_device(0)
out = {}
for name, kw in (("fp32", {}), ("tc", {"tc": True}), ("tc_global", {"tc": True, "tc_kernel": pvt.KERNEL_TC_GLOBAL})):
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    r = bench.whole_frame_leg(pvt, torch, {"ms_per_step": 0.0}, **kw)
    print(name, "whole-frame step ms", round(r["ms_per_step"], 5), "| local step ms (lost mode on)", round(r["local_step_ms_with_lost_mode_on"], 5), flush=True)

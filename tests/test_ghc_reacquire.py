"""Lost-object re-acquisition (the reference's second tracker, tracker_ghc/src/main.cpp:145-239).

CPU part: oracle/ncc_oracle.c:orc_track_clip_ghc against the cv2 4.13.0 fixtures of tests/golden/make_golden_ghc.py.
GPU part (-m gpu): the library in lost-object mode (pvt_params.lost_frame_threshold > 0: local pass + whole-frame pass per
step, all state on the device) against the same fixtures and the oracle -- identical trajectory, flags, search kind per
frame, lost counters and final template; confidences within 1e-4.  Nothing here reads /root/reference."""
import importlib
import json
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as Hp
from tools import synth

pvt = importlib.import_module("parallel-video-object-tracker_b200")

with open(os.path.join(Hp.GOLD, "meta_ghc.json")) as fh:
    META = json.load(fh)["clips"]
NAMES = sorted(META)


def load(name):
    m = META[name]
    c = synth.make_clip(synth.ClipSpec(**m["spec"]))
    assert zlib.crc32(np.ascontiguousarray(c["frames"]).tobytes()) & 0xFFFFFFFF == m["frames_crc"]
    return c, m["track"], np.load(os.path.join(Hp.GOLD, f"ghc_{name}.npz"))


def check(got, want, what):
    """got/want: [n, 10] x y w h conf moved updated searched lost_count use_global"""
    assert np.array_equal(got[:, :4].astype(np.int64), want[:, :4].astype(np.int64)), f"{what}: bbox trajectory differs"
    assert np.array_equal(got[:, 5:8].astype(np.int64), want[:, 5:8].astype(np.int64)), f"{what}: moved/updated/search-kind differ"
    d = np.abs(got[:, 4] - want[:, 4]).max()
    assert d <= Hp.TOL_SCORE, f"{what}: confidence differs by {d}"


@pytest.mark.parametrize("name", NAMES)
def test_oracle_ghc_vs_cv2_golden(name):
    c, tk, g = load(name)
    rec, templ = O.track_clip_ghc(c["frames"], c["roi"], **tk)
    want = g["records"]
    check(rec, want, name)
    assert np.array_equal(rec[:, 8:10].astype(np.int64), want[:, 8:10].astype(np.int64)), "lost counter / global flag differ"
    assert np.array_equal(templ, g["templ"]), "final template differs (EMA recurrence is bit-exact)"
    m = META[name]
    assert m["reacquired"] >= 1 and m["global_frames"] > m["reacquired"]   # the fixture exercises failed AND successful whole-frame searches


def test_oracle_ghc_without_losses_equals_plain_tracker():
    """With a threshold that is never reached the ghc loop is the tracker/ loop (same window clamp, gates, EMA)."""
    c, tk = Hp.clip("small")
    a, ta = O.track_clip(c["frames"], c["roi"])
    b, tb = O.track_clip_ghc(c["frames"], c["roi"], rx=80, ry=80, lost_threshold=1000)
    assert np.array_equal(a[:, :7], b[:, :7]) and np.array_equal(ta, tb)


# ------------------------------------------------------------------------------------------------- GPU
def gpu_records(frames, roi, tk, **extra):
    n, H, W, _ = frames.shape
    x, y, w, h = roi
    kw = dict(search_radius_x=tk.get("rx", 60), search_radius_y=tk.get("ry", 60), lost_frame_threshold=tk.get("lost_threshold", 50),
              ncc_global_confidence=tk.get("global_conf", 0.60))
    kw.update(extra)
    out = []
    with pvt.Tracker(W, H, w, h, **kw) as tr:
        tr.init_track(0, frames[0], roi)
        for k in range(1, n):
            r = tr.step([frames[k]])[0]
            lost, glob = tr.get_lost_state(0)
            out.append((r["x"], r["y"], r["w"], r["h"], float(r["conf"]), r["moved"], r["updated"], r["searched"], lost, glob))
        _, templ = tr.get_state(0)
    return np.array(out, np.float64), templ


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_ghc_vs_golden_and_oracle(name):
    c, tk, g = load(name)
    rec, templ = gpu_records(c["frames"], c["roi"], tk)
    want = g["records"]
    check(rec, want, name + " vs cv2")
    # the golden's lost counter / flag describe the state AFTER the frame, like pvt_get_lost_state
    assert np.array_equal(rec[:, 8].astype(np.int64), want[:, 8].astype(np.int64)), "lost_frame_count differs"
    orec, otempl = O.track_clip_ghc(c["frames"], c["roi"], **tk)
    check(rec, orec, name + " vs oracle")
    assert np.array_equal(templ, otempl), "final template differs from the oracle's (bit-exact EMA given identical peaks)"
    assert (rec[:, 7] == 2).sum() == META[name]["global_frames"]


@pytest.mark.gpu
def test_gpu_ghc_async_sequence_and_two_tracks():
    """No host round trip: the whole clip is enqueued at once (resident frames), two tracks on one stream -- one of them
    follows the object (gets lost and is re-acquired), the other sits on static background and never leaves the local pass."""
    import torch
    name = "reacquire"
    c, tk, g = load(name)
    frames, roi = c["frames"], c["roi"]
    n, H, W, _ = frames.shape
    dev = torch.from_numpy(frames).cuda()
    with pvt.Tracker(W, H, 32, 32, max_streams=1, max_tracks=2, search_radius_x=tk["rx"], search_radius_y=tk["ry"],
                     lost_frame_threshold=tk["lost_threshold"]) as tr:
        tr.init_track(0, pvt.device_frame(dev[0].data_ptr(), W * 3, stream=0), roi, stream=0)
        tr.init_track(1, None, (8, 8, 32, 32), stream=0)
        ring = [[pvt.Frame(0, pvt.FMT_BGR8, pvt.MEM_DEVICE, 0, dev[k].data_ptr(), W * 3)] for k in range(1, n)]
        tr.submit_sequence(n - 1, ring)
        res = tr.collect(n - 1)
    got = np.array([(r["x"], r["y"], r["w"], r["h"], float(r["conf"]), r["moved"], r["updated"], r["searched"], 0, 0) for r in res[:, 0]], np.float64)
    check(got, g["records"], "async two-track")
    assert np.all(res[:, 1]["searched"] == 1) and np.all(res[:, 1]["valid"] == 1)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["auto", "tc_global", "tc"])
@pytest.mark.parametrize("name", ["reacquire", "reacquire_odd"])
def test_gpu_ghc_async_sequence_single_track_n_step_graphs(name, kernel):
    """One track, the whole clip enqueued at once: sequences of up to 16 steps are ONE graph launch in this mode too (a conditional
    node per step), the local pass in the latency shape; the track gets lost and is re-acquired INSIDE such a graph.  Same cv2
    golden; kernel "tc_global" / "tc": the whole-frame pass on the tensor cores."""
    import torch
    c, tk, g = load(name)
    frames, roi = c["frames"], c["roi"]
    n, H, W, _ = frames.shape
    dev = torch.from_numpy(frames).cuda()
    kern = {"auto": pvt.KERNEL_AUTO, "tc_global": pvt.KERNEL_TC_GLOBAL, "tc": pvt.KERNEL_TC}[kernel]
    with pvt.Tracker(W, H, roi[2], roi[3], search_radius_x=tk["rx"], search_radius_y=tk["ry"], lost_frame_threshold=tk["lost_threshold"],
                     ncc_global_confidence=tk.get("global_conf", 0.60), kernel=kern) as tr:
        tr.init_track(0, pvt.device_frame(dev[0].data_ptr(), W * 3, stream=0), roi, stream=0)
        ring = [[pvt.Frame(0, pvt.FMT_BGR8, pvt.MEM_DEVICE, 0, dev[k].data_ptr(), W * 3)] for k in range(1, n)]
        tr.submit_sequence(n - 1, ring)
        res = tr.collect(n - 1)
        _, templ = tr.get_state(0)
    got = np.array([(r["x"], r["y"], r["w"], r["h"], float(r["conf"]), r["moved"], r["updated"], r["searched"], 0, 0) for r in res[:, 0]], np.float64)
    check(got, g["records"], f"async single track ({kernel})")
    assert (got[:, 7] == 2).sum() == META[name]["global_frames"] > 0
    assert np.array_equal(templ, g["templ"])


@pytest.mark.gpu
def test_gpu_ghc_checkpoint_resume_in_lost_state():
    name = "reacquire"
    c, tk, g = load(name)
    frames, roi = c["frames"], c["roi"]
    want = g["records"]
    cut = int(np.argmax(want[:, 7] == 2)) + 2          # two frames into the whole-frame search phase
    n, H, W, _ = frames.shape
    kw = dict(search_radius_x=tk["rx"], search_radius_y=tk["ry"], lost_frame_threshold=tk["lost_threshold"])
    with pvt.Tracker(W, H, 32, 32, **kw) as tr:
        tr.init_track(0, frames[0], roi)
        for k in range(1, cut + 1):
            tr.step([frames[k]])
        bbox, templ = tr.get_state(0)
        lost, glob = tr.get_lost_state(0)
    assert glob == 1 and lost == int(want[cut - 1, 8])
    out = []
    with pvt.Tracker(W, H, 32, 32, **kw) as tr:
        tr.init_track(0, frames[0], roi)
        tr.set_state(0, bbox, templ)
        tr.set_lost_state(0, lost, glob)
        for k in range(cut + 1, n):
            r = tr.step([frames[k]])[0]
            out.append((r["x"], r["y"], r["w"], r["h"], float(r["conf"]), r["moved"], r["updated"], r["searched"], 0, 0))
    check(np.array(out, np.float64), want[cut:], "resume")


@pytest.mark.gpu
def test_gpu_lost_mode_is_creation_time_choice():
    with pvt.Tracker(64, 64, 8, 8) as tr:
        with pytest.raises(pvt.PvtError) as e:
            tr.set_params(lost_frame_threshold=5)       # the whole-frame scratch is sized at pvt_create
        assert e.value.code == pvt.ERR_INVALID
        tr.params.lost_frame_threshold = 0
        tr.init_track(0, np.zeros((64, 64, 3), np.uint8), (4, 4, 8, 8))
        with pytest.raises(pvt.PvtError) as e:
            tr.set_lost_state(0, 1, 1)
        assert e.value.code == pvt.ERR_STATE
    with pvt.Tracker(64, 64, 8, 8, lost_frame_threshold=5) as tr:
        tr.set_params(lost_frame_threshold=9, ncc_global_confidence=0.5)   # values may change while the mode stays on
        with pytest.raises(pvt.PvtError):
            tr.set_params(lost_frame_threshold=0)

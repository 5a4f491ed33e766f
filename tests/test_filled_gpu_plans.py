"""Parity on the FILLED-GPU plans: the shapes BASELINE.json's configs[3] / configs[4] and the 1080p whole-frame pass
actually run -- unsplit search + k_ncc_fringe behind it (programmatic dependent launch) + TAIL SPLIT with
k_ncc_tail_finalize, and the 6-band whole-frame plan -- under the planner's own choice (no PVT_PLAN override).
Every test asserts from pvt_plan_query that the plan under test is the one with n_tail > 0 (and fringe == 3), so it
cannot silently stop covering that path.

GPU tests (-m gpu) compare, through the C ABI, with (1) the cv2 4.13.0 golden records / final-template CRCs of
tests/golden/make_golden_filled.py and (2) the CPU oracle's track_step incl. the whole window map of every track
(G1..G5 of SURVEY.md 8(c)).  The CPU tests hold the oracle to the same goldens on a subset.
Reference: tracker/src/main.cpp:98-161 per track; tracker_ghc/src/main.cpp:183-239 for the whole-frame case."""
import importlib
import json
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as O
from tests import filled_cases as FC
from tests import helpers as Hp

pvt = importlib.import_module("parallel-video-object-tracker_b200")


def _meta():
    with open(os.path.join(Hp.GOLD, "meta_filled.json")) as fh:
        return json.load(fh)


def _tcrc(t):
    return zlib.crc32(np.ascontiguousarray(t, np.float32).tobytes()) & 0xFFFFFFFF


def _rec7(r):
    return np.array([r["x"], r["y"], r["w"], r["h"], float(r["conf"]), r["moved"], r["updated"]], np.float64)


def _check_track_vs_oracle(gray, templ, box, res, m, win, what):
    """One track, one step: oracle step from the same state; G1/G2/G5 on the record, G3/G4 on the whole window map."""
    rec, owin, want = O.track_step(gray, templ, box[0], box[1], rx=FC.R, ry=FC.R, want_map=True)   # templ updated in place
    assert win == owin, what
    assert m.shape == want.shape, what
    sig = Hp.window_sigma(gray, FC.TW, FC.TH, owin)
    d = np.abs(m - want)
    assert d[sig >= 0.002].max(initial=0) <= Hp.TOL_SCORE, (what, float(d.max()))
    assert d.max() <= Hp.TOL_LOWVAR, (what, float(d.max()))
    deg = (want == 0) | (np.abs(want) == 1)
    assert np.array_equal(m[deg], want[deg]), what
    top = np.sort(want.ravel())[-2:]
    if want.size == 1 or top[1] - top[0] >= Hp.AMBIGUOUS_GAP:
        assert np.argmax(m) == np.argmax(want), what
        assert (res["x"], res["y"], res["w"], res["h"]) == (rec.x, rec.y, rec.w, rec.h), what
        assert (res["moved"], res["updated"]) == (rec.moved, rec.updated), what
        assert abs(float(res["conf"]) - rec.conf) <= Hp.TOL_SCORE, what
    return rec


# ---- CPU: the oracle against the cv2 goldens of these cases (subset; the full sets run on the GPU box) -----------------
def test_oracle_matches_filled_goldens_subset():
    meta = _meta()
    c5 = FC.C5()
    g = Hp.golden("filled_c5_64x1080p.npz")
    for s in (3, 20, 63):
        fr = np.stack([c5.frame(s, k) for k in range(c5.n_frames)])
        rec, templ = O.track_clip(fr, c5.roi(s), rx=FC.R, ry=FC.R)
        Hp.check_records(rec[:, :7], g["records"][s], f"c5 stream {s}")
        assert _tcrc(templ) == int(g["templ_crc"][s])
        if f"templ_{s}" in g.files:
            assert np.array_equal(templ, g[f"templ_{s}"])
    assert meta["c5"]["rois"][63] == list(c5.roi(63))
    c4 = FC.C4()
    assert c4.crc() == meta["c4"]["frames_crc"]
    g = Hp.golden("filled_c4_256roi.npz")
    rois = c4.rois()
    for t in (0, 100, 241, 255):
        rec, templ = O.track_clip(np.stack(c4.frames), rois[t], rx=FC.R, ry=FC.R)
        Hp.check_records(rec[:, :7], g["records"][t], f"c4 roi {t}")
        assert np.array_equal(templ, g[f"templ_{t}"])


def test_planner_gives_the_filled_plans_these_tests_are_about():
    p5 = pvt.plan_query(64, FC.TW, FC.TH, FC.W, FC.H, FC.R, FC.R)
    p4 = pvt.plan_query(256, FC.TW, FC.TH, FC.W, FC.H, FC.R, FC.R)
    pw = pvt.plan_query(1, FC.TW, FC.TH, FC.W, FC.H, FC.W, FC.H)
    for p in (p5, p4):
        assert p["pj"] * p["pd"] == 1 and p["n_tail"] > 0 and p["tail_parts"] >= 2 and p["fringe"] == 3
    assert pw["pj"] * pw["pd"] == 1 and pw["bands"] == 6 and pw["n_tail"] > 0 and pw["tail_parts"] >= 2


# ---- GPU ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["auto", "tc"])
def test_c5_64_streams_1080p_tail_split_and_fringe(kernel):
    """kernel "tc": the same case on the tensor-core search (k_ncc_tc; two 128-row tiles per track, band-aware MMAs, every
    window-origin alignment 0..15 occurs among the 64 streams) -- same gates, same goldens."""
    sm = pvt.device_info(0)["sm_count"]
    plan = pvt.plan_query(64, FC.TW, FC.TH, FC.W, FC.H, FC.R, FC.R, sm_count=sm)
    assert plan["pj"] * plan["pd"] == 1 and plan["n_tail"] > 0 and plan["tail_parts"] >= 2 and plan["fringe"] == 3, plan
    c5 = FC.C5()
    assert c5.crc() == _meta()["c5"]["frames_crc"], "synthetic c5 frames are not byte-identical to the golden run's"
    gold = Hp.golden("filled_c5_64x1080p.npz")
    n = c5.n_streams
    with pvt.Tracker(FC.W, FC.H, FC.TW, FC.TH, max_streams=n, max_tracks=n, keep_maps=1, search_radius_x=FC.R, search_radius_y=FC.R,
                     kernel=pvt.KERNEL_TC if kernel == "tc" else pvt.KERNEL_AUTO) as tr:
        f0 = c5.frames_at(0)
        templs, boxes = [], []
        for s in range(n):
            roi = c5.roi(s)
            tr.init_track(s, f0[s], roi, stream=s)
            templs.append(O.to_gray_f32(f0[s])[roi[1]:roi[1] + FC.TH, roi[0]:roi[0] + FC.TW].copy())
            boxes.append((roi[0], roi[1]))
        seen_fringe = seen_clamped = 0
        for k in range(1, c5.n_frames):
            fk = c5.frames_at(k)
            res = tr.step(fk)
            for s in range(n):
                m, win = tr.window_map(s)
                gray = O.to_gray_f32(fk[s])
                rec = _check_track_vs_oracle(gray, templs[s], boxes[s], res[s], m, win, f"c5 stream {s} frame {k}")
                boxes[s] = (rec.x, rec.y)
                seen_fringe += win[2] == 2 * FC.R + 1 and win[3] == 2 * FC.R + 1
                seen_clamped += win[2] < 2 * FC.R + 1 or win[3] < 2 * FC.R + 1
                # cv2 4.13.0 golden: identical box and flags, confidence within 1e-4
                assert np.array_equal(_rec7(res[s])[[0, 1, 2, 3, 5, 6]], gold["records"][s, k - 1][[0, 1, 2, 3, 5, 6]]), (s, k)
                assert abs(float(res[s]["conf"]) - gold["records"][s, k - 1, 4]) <= Hp.TOL_SCORE, (s, k)
        assert seen_fringe >= 100 and seen_clamped >= 20
        for s in range(n):
            _, templ = tr.get_state(s)
            assert np.array_equal(templ, templs[s]), f"c5 stream {s}: final template differs from the oracle's"
            assert _tcrc(templ) == int(gold["templ_crc"][s]), f"c5 stream {s}: final template not bit-identical to cv2's"


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["auto", "tc"])
def test_c4_256_rois_one_stream_tail_split_and_fringe(kernel):
    sm = pvt.device_info(0)["sm_count"]
    plan = pvt.plan_query(256, FC.TW, FC.TH, FC.W, FC.H, FC.R, FC.R, sm_count=sm)
    assert plan["pj"] * plan["pd"] == 1 and plan["n_tail"] > 0 and plan["tail_parts"] >= 2 and plan["fringe"] == 3, plan
    c4 = FC.C4()
    assert c4.crc() == _meta()["c4"]["frames_crc"]
    gold = Hp.golden("filled_c4_256roi.npz")
    rois = c4.rois()
    n = len(rois)
    g0 = O.to_gray_f32(c4.frames[0])
    first_tail_track = plan["n_full"] // plan["ctas_per_track"]
    with pvt.Tracker(FC.W, FC.H, FC.TW, FC.TH, max_streams=1, max_tracks=n, keep_maps=1, search_radius_x=FC.R, search_radius_y=FC.R,
                     kernel=pvt.KERNEL_TC if kernel == "tc" else pvt.KERNEL_AUTO) as tr:
        for t, roi in enumerate(rois):
            tr.init_track(t, c4.frames[0] if t == 0 else None, roi)
        templs = [g0[r[1]:r[1] + FC.TH, r[0]:r[0] + FC.TW].copy() for r in rois]
        boxes = [(r[0], r[1]) for r in rois]
        for k in range(1, c4.n_frames):
            res = tr.step([c4.frames[k]])
            gray = O.to_gray_f32(c4.frames[k])
            # frame 1: every track against the oracle incl. its whole map; frame 2: the tail-item tracks, the border boxes and
            # a sample of the rest (the cv2 golden below covers every track on both frames)
            for t in range(n):
                full_check = k == 1 or t >= first_tail_track - 1 or t % 16 == 0
                if full_check:
                    m, win = tr.window_map(t)
                    rec = _check_track_vs_oracle(gray, templs[t], boxes[t], res[t], m, win, f"c4 roi {t} frame {k}")
                    boxes[t] = (rec.x, rec.y)
                assert np.array_equal(_rec7(res[t])[[0, 1, 2, 3, 5, 6]], gold["records"][t, k - 1][[0, 1, 2, 3, 5, 6]]), (t, k)
                assert abs(float(res[t]["conf"]) - gold["records"][t, k - 1, 4]) <= Hp.TOL_SCORE, (t, k)
        for t in range(n):
            _, templ = tr.get_state(t)
            assert _tcrc(templ) == int(gold["templ_crc"][t]), f"c4 roi {t}: final template not bit-identical to cv2's"
            if t >= first_tail_track - 1 or t % 16 == 0:
                assert np.array_equal(templ, templs[t]), t
        # mid-run additions: this context ingests whole frames (256 tiles cover the frame), so a new track may still be cut
        # from the stream's current image
        tr.init_track(5, None, rois[5])


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["auto", "tc", "tc_global"])
def test_whole_frame_1080p_six_band_tail_split_plan(kernel):
    """(1) map operator on a 1080p frame: the 1857 x 1017 map under the 6-band tail-split plan against the oracle's;
    (2) lost-object mode with the track forced lost: the same plan inside the conditional graph node, then back to local.
    kernel "tc": (2) with the whole-frame pass on the tensor cores (k_ncc_tc over column tiles of the 1857-wide map; the whole
    map of that shape is compared score by score in tests/test_tc_kernel.py::test_tc_whole_map_in_column_tiles_vs_oracle)."""
    sm = pvt.device_info(0)["sm_count"]
    plan = pvt.plan_query(1, FC.TW, FC.TH, FC.W, FC.H, FC.W, FC.H, sm_count=sm)
    assert plan["pj"] * plan["pd"] == 1 and plan["bands"] == 6 and plan["n_tail"] > 0 and plan["tail_parts"] >= 2, plan
    if kernel != "auto":   # the whole-frame pass on the tensor cores: 17 x 8 column / row tiles meet the map
        tcp = pvt.tc_plan_query(1, FC.TW, FC.TH, FC.W, FC.H, whole_frame_pass=True, sm_count=sm)
        assert tcp["xtiles"] >= 2 and tcp["xw"] <= 256, tcp
    wf = FC.WF()
    assert wf.crc() == _meta()["wf"]["frames_crc"]
    gold = Hp.golden("filled_wf_1080p.npz")
    roi = wf.roi()
    g0, g1 = O.to_gray_f32(wf.frames[0]), O.to_gray_f32(wf.frames[1])
    templ = g0[roi[1]:roi[1] + FC.TH, roi[0]:roi[0] + FC.TW].copy()
    want = O.ncc_match_cpu(g1, templ)
    m = pvt.ncc_match_naive_cuda(g1, templ)
    assert m.shape == want.shape == (FC.H - FC.TH + 1, FC.W - FC.TW + 1)
    sig = Hp.window_sigma(g1, FC.TW, FC.TH, (0, 0, want.shape[1], want.shape[0]))
    d = np.abs(m - want)
    assert d[sig >= 0.002].max(initial=0) <= Hp.TOL_SCORE and d.max() <= Hp.TOL_LOWVAR, float(d.max())
    assert np.argmax(m) == np.argmax(want)
    best, bx, by = O.max_loc(want)

    with pvt.Tracker(FC.W, FC.H, FC.TW, FC.TH, search_radius_x=FC.R, search_radius_y=FC.R, lost_frame_threshold=50,
                     ncc_global_confidence=0.60, kernel={"auto": pvt.KERNEL_AUTO, "tc": pvt.KERNEL_TC, "tc_global": pvt.KERNEL_TC_GLOBAL}[kernel]) as tr:
        tr.init_track(0, wf.frames[0], roi)
        tr.set_state(0, wf.stale_box, None)
        tr.set_lost_state(0, 1000, 1)
        r1 = tr.step([wf.frames[1]])[0]
        lost1 = tr.get_lost_state(0)
        r2 = tr.step([wf.frames[2]])[0]
        _, t_end = tr.get_state(0)
    # frame 1: whole-map arg-max (tracker_ghc/src/main.cpp:186-193), accepted at NCC_GLOBAL_CONFIDENCE (:217), EMA at 0.70
    assert int(r1["searched"]) == 2 and (r1["x"], r1["y"]) == (bx, by) and abs(float(r1["conf"]) - best) <= Hp.TOL_SCORE
    assert r1["moved"] == 1 and r1["updated"] == 1 and lost1 == (0, 0)
    t1 = O.add_weighted(templ, g1[by:by + FC.TH, bx:bx + FC.TW])
    rec2, _, _ = O.track_step(O.to_gray_f32(wf.frames[2]), t1, bx, by, rx=FC.R, ry=FC.R)
    assert int(r2["searched"]) == 1 and (r2["x"], r2["y"], r2["moved"], r2["updated"]) == (rec2.x, rec2.y, rec2.moved, rec2.updated)
    assert abs(float(r2["conf"]) - rec2.conf) <= Hp.TOL_SCORE
    assert np.array_equal(t_end, t1), "template after the whole-frame re-acquisition + one local step differs from the oracle's"
    # cv2 4.13.0 golden (cv2_harness.track_clip_ghc resumed from the same lost state)
    for r, g in ((r1, gold["records"][0]), (r2, gold["records"][1])):
        assert (r["x"], r["y"], r["moved"], r["updated"], int(r["searched"])) == tuple(int(v) for v in g[[0, 1, 5, 6, 7]])
        assert abs(float(r["conf"]) - g[4]) <= Hp.TOL_SCORE
    assert np.array_equal(t_end, gold["templ"]), "final template not bit-identical to cv2's"

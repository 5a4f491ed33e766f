"""Host logic of the search-kernel planner (pvt_plan_query: choose_plan + item grid / tail splitting), CPU only.
The plan decides tile geometry, K-split and tail split from the context's geometry; these tests pin the plans of the
BASELINE configurations and check the invariants the kernels rely on over a sweep of geometries."""
import importlib
import itertools

import pytest

pvt = importlib.import_module("parallel-video-object-tracker_b200")
SMS = 148


def q(n, tw, th, W, H, rx, ry=None):
    return pvt.plan_query(n, tw, th, W, H, rx, rx if ry is None else ry, sm_count=SMS)


def test_filled_gpu_plans_use_the_window_origin_grid_with_fringe():
    for n in (256, 64):                                 # C4 (256 ROIs), C5's per-GPU share (64 streams)
        p = q(n, 64, 64, 1920, 1080, 80)
        # 161 x 161 window = (20*8 + 1) x (32*5 + 1): 20 x 32 thread tiles = exactly 5 CTAs of 128 per track; the remainder
        # column and row go to k_ncc_fringe (bits 1 | 2); unsplit
        assert (p["G"], p["C"], p["GB"], p["ctas_per_track"], p["pj"], p["pd"], p["fringe"]) == (32, 20, 32, 5, 1, 1, 3)
        assert p["span"] == 4 and p["boxW"] == 8 * 4 + 64 + 4 and p["boxH"] == 32 * 5 + 64 - 1
        assert p["n_full"] % (2 * SMS) == 0 and p["n_full"] + p["n_tail"] == n * 5
        assert p["n_tail"] > 0 and 2 <= p["tail_parts"] <= 8 and p["n_tail"] * p["tail_parts"] <= 2 * SMS


def test_single_stream_plans_are_k_split_without_fringe_and_leave_sms_for_the_statistics():
    for (tw, th, W, H, R) in ((64, 64, 1920, 1080, 80), (128, 128, 3840, 2160, 160)):      # C2, C3
        p = q(1, tw, th, W, H, R)
        parts = p["pj"] * p["pd"]
        assert parts > 1 and p["fringe"] == 0 and p["n_tail"] == 0
        assert p["G"] == -(-(2 * R + 1) // 5) and p["C"] == -(-(2 * R + 1) // 8)          # remainder stays in the grid
        ctas = p["ctas_per_track"] * parts
        assert SMS // 2 <= ctas <= 2 * (SMS - 8)                                           # fills the GPU, 8 SMs stay free


def test_whole_frame_pass_is_unsplit_with_tail_splitting():
    p = q(1, 64, 64, 1920, 1080, 1920, 1080)            # lost-object mode: window == the whole map
    assert p["pj"] * p["pd"] == 1 and p["bands"] > 1 and p["boxH"] <= 256
    assert p["n_full"] == 2 * SMS and p["n_tail"] > 0 and p["tail_parts"] >= 2


GEOMS = list(itertools.product((1, 3, 64, 300), ((8, 8), (17, 13), (32, 32), (37, 29), (64, 64), (128, 128), (200, 40)),
                               ((320, 240), (1920, 1080), (3840, 2160)), (4, 12, 40, 80, 160)))


@pytest.mark.parametrize("n,templ,frame,R", GEOMS)
def test_plan_invariants(n, templ, frame, R):
    tw, th = templ
    W, H = frame
    if tw > W or th > H:
        pytest.skip("template larger than the frame")
    p = q(n, tw, th, W, H, R)
    Wmax, Hmax = min(2 * R + 1, W), min(2 * R + 1, H)
    nch = (tw + 7) // 8
    parts = p["pj"] * p["pd"]
    # TMA box: <= 256 per dimension, rows of 16-byte multiples, pitch == 4 (mod 8) floats (conflict-free LDS.128)
    assert 0 < p["boxW"] <= 256 and 0 < p["boxH"] <= 256 and p["boxW"] % 8 == 4
    assert 2 * (p["smem"] + 1024) <= 228 * 1024                                            # two CTAs per SM
    # the grid covers the window, minus at most one fringe row / column
    assert 8 * p["C"] >= Wmax - (p["fringe"] & 1) and 5 * p["G"] >= Hmax - (p["fringe"] >> 1 & 1)
    assert (p["fringe"] & 1) == 0 or Wmax % 8 == 1
    assert (p["fringe"] & 2) == 0 or Hmax % 5 == 1
    assert p["bands"] * p["GB"] >= p["G"] and p["ctas_per_band"] * 128 >= p["GB"] * p["C"]
    assert p["ctas_per_track"] == p["bands"] * p["ctas_per_band"]
    assert 8 * p["span"] + 8 * -(-nch // p["pj"]) + 4 == p["boxW"]
    # K-split parts are never empty; K-split never meets the fringe kernel; tail splitting only without K-split
    assert 1 <= p["pj"] <= nch and 1 <= p["pd"] <= min(th, 32)
    assert parts == 1 or (p["fringe"] == 0 and p["n_tail"] == 0)
    assert p["n_full"] + p["n_tail"] == n * p["ctas_per_track"]
    assert p["n_tail"] == 0 or (2 <= p["tail_parts"] <= nch and p["n_tail"] * p["tail_parts"] <= 2 * SMS)


def test_plan_query_rejects_bad_geometry():
    with pytest.raises(pvt.PvtError):
        q(1, 64, 64, 32, 32, 80)
    with pytest.raises(pvt.PvtError):
        q(0, 8, 8, 64, 64, 8)

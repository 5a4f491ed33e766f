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


def test_tc_plan_column_tiles_host_logic():
    """k_ncc_tc geometry (pvt_tc_plan_query, no device): one accumulator where the window fits (the round-2 plans of C2/C4/C5 are
    unchanged), column tiles for 4K windows and for the whole-frame pass; the limits the kernel's issue code is unrolled for hold."""
    c5 = pvt.tc_plan_query(64, 64, 64, 1920, 1080, 80, 80)
    assert (c5["xw"], c5["xtiles"], c5["mtiles"], c5["ksteps"], c5["tmem_cols"]) == (176, 1, 2, 8, 512)
    small = pvt.tc_plan_query(1, 32, 32, 320, 240, 12, 12)
    assert (small["xw"], small["xtiles"], small["mtiles"]) == (32, 1, 1)
    wf = pvt.tc_plan_query(1, 64, 64, 1920, 1080, whole_frame_pass=True)
    assert wf["xw"] == 112 and wf["xtiles"] == 18 and wf["mtiles"] == 9 and wf["tmem_cols"] == 256   # 17 x 8 of them meet the 1857 x 1017 map: 136 CTAs, one wave
    assert -(-(1920 - 64 + 1) // wf["xw"]) * -(-(1080 - 64 + 1) // 128) <= 148
    c3 = pvt.tc_plan_query(1, 128, 128, 3840, 2160, 160, 160)
    assert c3["xtiles"] >= 2 and c3["xw"] * c3["xtiles"] >= 321 and c3["mtiles"] == 3
    for p in (c5, small, wf, c3, pvt.tc_plan_query(8, 17, 129, 333, 130, 333, 130), pvt.tc_plan_query(1, 260, 40, 1920, 1080, whole_frame_pass=True)):
        assert p["xw"] % 16 == 0 and 16 <= p["xw"] <= 256 and p["ksteps"] <= 10 and p["tmem_cols"] >= 2 * p["xw"] and p["smem"] <= 227 * 1024
        assert p["groups"] == p["xw"] // 8 and p["stages"] >= 2
    with pytest.raises(pvt.PvtError) as e:                      # 128 candidate rows + th - 1 exceed the 256-row TMA box
        pvt.tc_plan_query(1, 160, 160, 3840, 2160, 160, 160)
    assert e.value.code == pvt.ERR_UNSUPPORTED

"""PVT_KERNEL_TC -- the tensor-core search (csrc/ncc_tc.cuh: tcgen05.mma kind::i8, exact integer cross term of the 16-bit
fixed-point centred template) against the same oracle, cv2 4.13.0 goldens and gates (G1..G5) as the FP32
kernels.  The filled-GPU shapes (64 x 1080p streams, 256 ROIs) run in tests/test_filled_gpu_plans.py[tc].
Reference semantics: tracker/src/ncc_cpu.cpp:12, main.cpp:135-161."""
import importlib
import json
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as Hp

pvt = importlib.import_module("parallel-video-object-tracker_b200")
pytestmark = pytest.mark.gpu


def records_of(res):
    return np.stack([res["x"], res["y"], res["w"], res["h"], res["conf"].astype(np.float64), res["moved"], res["updated"]], 1).astype(np.float64)


@pytest.mark.parametrize("name", ["small", "lowtex", "lost", "fade", "border", "flat", "oddsize", "c1_standin", "c2_1080p"])
def test_tc_clip_trajectory_vs_cv2_golden(name):
    (c, tk) = Hp.clip(name)
    g = Hp.golden(f"clip_{name}.npz")
    recs, templ = pvt.track_clip(c["frames"], c["roi"], search_radius_x=tk.get("rx", 80), search_radius_y=tk.get("ry", 80), kernel=pvt.KERNEL_TC)
    Hp.check_records(records_of(recs), g["records"], name + " (tc)")
    assert np.array_equal(templ, g["templ"]), f"{name}: final template not bit-identical"


@pytest.mark.parametrize("name,k", [("small", 1), ("small", 7), ("lowtex", 2), ("border", 3), ("flat", 1), ("oddsize", 2), ("c2_1080p", 1)])
def test_tc_window_map(name, k):
    (c, tk) = Hp.clip(name)
    g = Hp.golden(f"clip_{name}.npz")
    frames, roi = c["frames"], c["roi"]
    H, W = frames.shape[1:3]
    templ, bbox = g[f"map{k}_templ"], g[f"map{k}_bbox"]
    with pvt.Tracker(W, H, roi[2], roi[3], keep_maps=1, search_radius_x=tk.get("rx", 80), search_radius_y=tk.get("ry", 80), kernel=pvt.KERNEL_TC) as tr:
        tr.init_track(0, frames[0], roi)
        tr.set_state(0, (int(bbox[0]), int(bbox[1]), roi[2], roi[3]), templ)
        tr.step([frames[k]])
        m, win = tr.window_map(0)
    assert win == tuple(int(v) for v in g[f"map{k}_win"])
    off, on = g[f"map{k}_ipp_off"], g[f"map{k}_ipp_on"]
    sig = Hp.window_sigma(O.to_gray_f32(frames[k]), roi[2], roi[3], win)
    d = np.abs(m - off)
    assert d[sig >= 0.002].max(initial=0) <= Hp.TOL_SCORE, float(d.max())      # G3 vs the exact (IPP-off) oracle
    assert d[sig < 0.002].max(initial=0) <= Hp.TOL_LOWVAR
    assert np.abs(m - on)[sig >= 0.02].max(initial=0) <= Hp.TOL_SCORE
    deg = (off == 0) | (np.abs(off) == 1)
    assert np.array_equal(m[deg], off[deg])                                     # G4
    assert np.argmax(m) == np.argmax(on)                                        # G1
    d_dir = os.path.join(Hp.ROOT, "gpurun_out")
    if os.path.isdir(d_dir):
        with open(os.path.join(d_dir, "tc_diffs.txt"), "a") as fh:
            fh.write(f"{name}/{k}: max |k_ncc_tc - cv2 IPP-off| = {d.max():.3e}\n")


def test_tc_exact_ties_pick_lowest_index_and_score_identically():
    """Integer accumulation is order-free: identical windows get identical cross terms wherever they sit in the tile (any
    x mod 16, either 128-row tile), so an exactly periodic frame ties exactly and the first maximum in row-major order wins
    (cv::minMaxLoc, main.cpp:150)."""
    rng = np.random.default_rng(5)
    pw, ph, W, H = 48, 36, 400, 330
    patch = rng.integers(0, 256, (ph, pw), dtype=np.uint8)
    frame = np.tile(patch, (H // ph + 1, W // pw + 1))[:H, :W].copy()
    tw, th, R = 20, 18, 100
    with pvt.Tracker(W, H, tw, th, keep_maps=1, search_radius_x=R, search_radius_y=R, kernel=pvt.KERNEL_TC) as tr:
        tr.init_track(0, frame, (7, 5, tw, th))
        tr.set_state(0, (150, 140, tw, th), None)
        r = tr.step([frame])[0]
        m, win = tr.window_map(0)
    gray = O.gray_to_f32(frame)
    rec, owin, want = O.track_step(gray, gray[5:5 + th, 7:7 + tw].copy(), 150, 140, rx=R, ry=R, want_map=True)
    assert win == owin and np.abs(m - want).max() <= Hp.TOL_SCORE
    ys, xs = np.nonzero(want == want.max())
    assert len(ys) >= 12                                         # the period fits the window many times
    vals = m[ys, xs]
    assert np.all(vals == vals[0]) and vals[0] == m.max()        # bit-identical scores at every repeat
    assert (r["x"], r["y"]) == (rec.x, rec.y) == (win[0] + xs[0], win[1] + ys[0])


@pytest.mark.parametrize("seed", range(10))
def test_tc_random_geometries_vs_oracle(seed):
    rng = np.random.default_rng(777 + seed)
    W, H = int(rng.integers(120, 460)), int(rng.integers(90, 330))
    tw, th = int(rng.integers(3, min(90, W // 2))), int(rng.integers(3, min(90, H // 2)))
    rx, ry = int(rng.integers(1, min(100, (290 - tw) // 2))), int(rng.integers(1, 100))
    n_tracks = int(rng.choice([1, 3, 9]))
    from scipy.ndimage import gaussian_filter
    base = gaussian_filter(rng.random((H, W, 3)), (1.2, 1.2, 0))
    f0 = np.clip((base - base.min()) / (base.max() - base.min()) * 255, 0, 255).astype(np.uint8)
    f1 = np.clip(np.roll(f0, (int(rng.integers(-2, 3)), int(rng.integers(-2, 3))), (0, 1)).astype(np.int16) + rng.integers(-3, 4, f0.shape), 0, 255).astype(np.uint8)
    outW, outH = W - tw + 1, H - th + 1
    boxes = [(int(rng.integers(0, outW)), int(rng.integers(0, outH))) for _ in range(n_tracks)]
    boxes[0] = [(0, 0), (outW - 1, outH - 1), (outW // 2, outH // 2)][seed % 3]
    g0, g1 = O.to_gray_f32(f0), O.to_gray_f32(f1)
    with pvt.Tracker(W, H, tw, th, max_streams=1, max_tracks=n_tracks, keep_maps=1, search_radius_x=rx, search_radius_y=ry, kernel=pvt.KERNEL_TC) as tr:
        for t, (x, y) in enumerate(boxes):
            tr.init_track(t, f0 if t == 0 else None, (x, y, tw, th))
        res = tr.step([f1])
        for t, (x, y) in enumerate(boxes):
            templ = g0[y:y + th, x:x + tw].copy()
            rec, win, want = O.track_step(g1, templ, x, y, rx=rx, ry=ry, want_map=True)
            m, w = tr.window_map(t)
            assert w == win, (seed, t)
            sig = Hp.window_sigma(g1, tw, th, win)
            d = np.abs(m - want)
            assert d[sig >= 0.002].max(initial=0) <= Hp.TOL_SCORE and d.max() <= Hp.TOL_LOWVAR, (seed, t, float(d.max()))
            gap = np.sort(want.ravel())[-2:]
            if want.size == 1 or gap[1] - gap[0] >= 1e-4:
                assert (res[t]["x"], res[t]["y"]) == (rec.x, rec.y), (seed, t)
                assert (res[t]["moved"], res[t]["updated"]) == (rec.moved, rec.updated)
                _, got_t = tr.get_state(t)
                assert np.array_equal(got_t, templ), (seed, t)


def test_tc_contract_errors_and_gray_sources():
    (c, _) = Hp.clip("small")
    frames, roi = c["frames"], c["roi"]
    H, W = frames.shape[1:3]
    want = records_of(pvt.track_clip(frames[:6], roi)[0])
    # 8-bit gray frames are accepted (utils.hpp:9-10: a single-channel input skips cvtColor) ...
    g8 = np.stack([O.bgr2gray(f) for f in frames[:6]])
    with pvt.Tracker(W, H, 32, 32, kernel=pvt.KERNEL_TC) as tr:
        tr.init_track(0, g8[0], roi)
        got = np.array([tr.step([g8[k]])[0] for k in range(1, 6)])
        Hp.check_records(records_of(got), want, "tc gray8")
        # ... float frames are not: the search runs on the gray LEVELS
        with pytest.raises(pvt.PvtError) as e:
            tr.step([O.to_gray_f32(frames[1])])
        assert e.value.code == pvt.ERR_INVALID
        tr.set_params(kernel=pvt.KERNEL_AUTO)                     # the FP32 kernels remain available on a TC context
        tr.set_params(kernel=pvt.KERNEL_TC)
    with pvt.Tracker(W, H, 32, 32) as tr:                         # but not the other way round: buffers are sized at creation
        with pytest.raises(pvt.PvtError) as e:
            tr.set_params(kernel=pvt.KERNEL_TC)
        assert e.value.code == pvt.ERR_INVALID
    with pytest.raises(pvt.PvtError) as e:                        # templates higher than 129 rows: 128 candidate rows + th - 1 exceed the 256-row TMA box
        pvt.Tracker(3840, 2160, 160, 160, search_radius_x=160, search_radius_y=160, kernel=pvt.KERNEL_TC)
    assert e.value.code == pvt.ERR_UNSUPPORTED


def test_tc_lost_object_mode_and_batch_hold():
    """kernel TC + lost-object mode: local windows AND the whole-frame pass (column tiles) on the tensor cores; + batch hold."""
    from tools import synth
    with open(os.path.join(Hp.GOLD, "meta_ghc.json")) as fh:
        m = json.load(fh)["clips"]["reacquire"]
    c = synth.make_clip(synth.ClipSpec(**m["spec"]))
    assert zlib.crc32(np.ascontiguousarray(c["frames"]).tobytes()) & 0xFFFFFFFF == m["frames_crc"]
    want = np.load(os.path.join(Hp.GOLD, "ghc_reacquire.npz"))["records"]
    tk = m["track"]
    recs, _ = pvt.track_clip(c["frames"], c["roi"], search_radius_x=tk["rx"], search_radius_y=tk["ry"], lost_frame_threshold=tk["lost_threshold"],
                             ncc_global_confidence=0.60, kernel=pvt.KERNEL_TC)
    got = records_of(recs)
    assert np.array_equal(got[:, :4], want[:, :4]) and np.array_equal(got[:, 5:7], want[:, 5:7])
    assert np.array_equal(recs["searched"], want[:, 7].astype(np.uint8))
    assert np.abs(got[:, 4] - want[:, 4]).max() <= Hp.TOL_SCORE
    (cb, _) = Hp.clip("batch4")
    g = Hp.golden("clip_batch4.npz")
    rb, tb = pvt.track_clip(cb["frames"], cb["roi"], mode=pvt.MODE_BATCH, batch_size=4, kernel=pvt.KERNEL_TC)
    Hp.check_records(records_of(rb), g["records"], "batch4 (tc)")
    assert np.array_equal(tb, g["templ"])


# ---- column tiles: windows wider than one accumulator (4K windows, whole-frame maps) are cut into tiles of TcCfg.XW candidate
# ---- columns; a CTA = (track, 128 rows, XW columns) with its own origin alignment, partial last tile and band
def _whole_map_case(W, H, tw, th, seed):
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    base = gaussian_filter(rng.random((H, W, 3)), (1.5, 1.5, 0))
    f0 = np.clip((base - base.min()) / (base.max() - base.min()) * 255, 0, 255).astype(np.uint8)
    f1 = np.clip(np.roll(f0, (2, -3), (0, 1)).astype(np.int16) + rng.integers(-3, 4, f0.shape), 0, 255).astype(np.uint8)
    f1[: H // 5, : W // 4] = 77                                  # a flat corner: degenerate windows (G4) in the first tiles
    return f0, f1


@pytest.mark.parametrize("W,H,tw,th,xw", [(700, 300, 40, 36, None), (700, 300, 40, 36, 64), (700, 300, 40, 36, 240), (523, 417, 64, 64, 112),
                                          (333, 130, 17, 129, None), (1920, 1080, 64, 64, None)])
def test_tc_whole_map_in_column_tiles_vs_oracle(W, H, tw, th, xw, monkeypatch):
    """Search radius = the frame: the local window is the whole (W - tw + 1) x (H - th + 1) map, wider than one accumulator.
    Every score against the oracle's map (G3/G4), the arg-max and the record (G1/G2/G5); PVT_TC_XW forces tile widths."""
    if xw is not None:
        monkeypatch.setenv("PVT_TC_XW", str(xw))
    else:
        monkeypatch.delenv("PVT_TC_XW", raising=False)
    plan = pvt.tc_plan_query(1, tw, th, W, H, W, H)             # reads PVT_TC_XW like pvt_create does
    assert plan["xtiles"] >= 2 and (xw is None or plan["xw"] == xw), plan   # the test cannot silently stop covering the column tiles
    f0, f1 = _whole_map_case(W, H, tw, th, 1000 + W + tw)
    x, y = (W - tw) // 3, (H - th) // 2
    g0, g1 = O.to_gray_f32(f0), O.to_gray_f32(f1)
    templ = g0[y:y + th, x:x + tw].copy()
    with pvt.Tracker(W, H, tw, th, keep_maps=1, search_radius_x=W, search_radius_y=H, kernel=pvt.KERNEL_TC) as tr:
        tr.init_track(0, f0, (x, y, tw, th))
        r = tr.step([f1])[0]
        m, win = tr.window_map(0)
        _, t_end = tr.get_state(0)
    assert win == (0, 0, W - tw + 1, H - th + 1)
    want = O.ncc_match_cpu(g1, templ)
    assert m.shape == want.shape
    sig = Hp.window_sigma(g1, tw, th, win)
    d = np.abs(m - want)
    assert d[sig >= 0.002].max(initial=0) <= Hp.TOL_SCORE and d.max() <= Hp.TOL_LOWVAR, float(d.max())
    deg = (want == 0) | (np.abs(want) == 1)
    assert np.array_equal(m[deg], want[deg])
    assert np.argmax(m) == np.argmax(want)
    rec, _, _ = O.track_step(g1, templ, x, y, rx=W, ry=H)       # templ updated in place
    assert (r["x"], r["y"], r["moved"], r["updated"]) == (rec.x, rec.y, rec.moved, rec.updated)
    assert abs(float(r["conf"]) - rec.conf) <= Hp.TOL_SCORE
    assert np.array_equal(t_end, templ)


def test_tc_c3_4k_windows_in_column_tiles_vs_cv2_golden():
    """BASELINE.json configs[2] (4K, 128 x 128, R 160): the 321-wide window is cut into column tiles."""
    (c, tk) = Hp.clip("c3_4k")
    g = Hp.golden("clip_c3_4k.npz")
    H, W = c["frames"].shape[1:3]
    assert pvt.tc_plan_query(1, c["roi"][2], c["roi"][3], W, H, tk.get("rx", 160), tk.get("ry", 160))["xtiles"] >= 2
    recs, templ = pvt.track_clip(c["frames"], c["roi"], search_radius_x=tk.get("rx", 160), search_radius_y=tk.get("ry", 160), kernel=pvt.KERNEL_TC)
    Hp.check_records(records_of(recs), g["records"], "c3_4k (tc)")
    assert np.array_equal(templ, g["templ"])


@pytest.mark.parametrize("kernel", ["tc", "tc_global"])
@pytest.mark.parametrize("name", ["ghc_defaults", "reacquire", "reacquire_odd"])
def test_tc_lost_object_clips_whole_frame_pass_on_tensor_cores(name, kernel):
    """tracker_ghc/src/main.cpp:183-239 with kernel TC: the local windows and the whole-frame pass (column tiles) on the
    tensor cores, against the cv2 4.13.0 goldens of the lost-object clips.  kernel "tc_global" (PVT_KERNEL_TC_GLOBAL): FP32
    kernels for the local windows, k_ncc_tc for the whole-frame pass only."""
    from tools import synth
    with open(os.path.join(Hp.GOLD, "meta_ghc.json")) as fh:
        clips = json.load(fh)["clips"]
    if name not in clips:
        pytest.skip(f"no lost-object golden named {name}")
    m = clips[name]
    c = synth.make_clip(synth.ClipSpec(**m["spec"]))
    assert zlib.crc32(np.ascontiguousarray(c["frames"]).tobytes()) & 0xFFFFFFFF == m["frames_crc"]
    z = np.load(os.path.join(Hp.GOLD, f"ghc_{name}.npz"))
    want = z["records"]
    tk = m["track"]
    recs, templ = pvt.track_clip(c["frames"], c["roi"], search_radius_x=tk["rx"], search_radius_y=tk["ry"], lost_frame_threshold=tk["lost_threshold"],
                                 ncc_global_confidence=tk.get("global_conf", 0.60), kernel=pvt.KERNEL_TC if kernel == "tc" else pvt.KERNEL_TC_GLOBAL)
    got = records_of(recs)
    assert np.array_equal(got[:, :4], want[:, :4]) and np.array_equal(got[:, 5:7], want[:, 5:7])
    assert np.array_equal(recs["searched"], want[:, 7].astype(np.uint8))
    assert np.abs(got[:, 4] - want[:, 4]).max() <= Hp.TOL_SCORE
    if "templ" in z.files:
        assert np.array_equal(templ, z["templ"])


def test_tc_global_runs_the_fp32_plan_on_local_windows():
    """PVT_KERNEL_TC_GLOBAL without a lost track is the planner's FP32 path bit for bit (records, final template), and reports
    the FP32 search kernel; GRAYF32 frames are rejected like with PVT_KERNEL_TC (the whole-frame pass reads the gray levels)."""
    (c, tk) = Hp.clip("small")
    frames, roi = c["frames"], c["roi"]
    H, W = frames.shape[1:3]
    ra, ta = pvt.track_clip(frames[:8], roi, lost_frame_threshold=50)
    rg, tg = pvt.track_clip(frames[:8], roi, lost_frame_threshold=50, kernel=pvt.KERNEL_TC_GLOBAL)
    assert np.array_equal(records_of(ra), records_of(rg)) and np.array_equal(ta, tg)
    with pvt.Tracker(W, H, roi[2], roi[3], lost_frame_threshold=50, kernel=pvt.KERNEL_TC_GLOBAL) as tr, \
         pvt.Tracker(W, H, roi[2], roi[3], lost_frame_threshold=50) as ta_:
        assert tr.search_kind()[0] == ta_.search_kind()[0] != "k_ncc_tc"
        tr.init_track(0, frames[0], roi)
        with pytest.raises(pvt.PvtError) as e:
            tr.step([O.to_gray_f32(frames[1])])
        assert e.value.code == pvt.ERR_INVALID


def test_tc_global_to_tc_switch_rebuilds_the_template_digits():
    """A PVT_KERNEL_TC_GLOBAL context keeps no template digits in its FP32 local pass; switching it to PVT_KERNEL_TC
    (pvt_set_params) must derive them for the templates as they are NOW (after EMA updates), not as they were at init."""
    (c, tk) = Hp.clip("small")
    g = Hp.golden("clip_small.npz")
    frames, roi = c["frames"], c["roi"]
    H, W = frames.shape[1:3]
    n = len(frames)
    cut = n // 2
    with pvt.Tracker(W, H, roi[2], roi[3], search_radius_x=tk.get("rx", 80), search_radius_y=tk.get("ry", 80), kernel=pvt.KERNEL_TC_GLOBAL) as tr:
        tr.init_track(0, frames[0], roi)
        got = [tr.step([frames[k]])[0] for k in range(1, cut)]
        assert any(int(r["updated"]) for r in got)              # the template has moved on since init
        tr.set_params(kernel=pvt.KERNEL_TC)
        assert tr.search_kind()[0] == "k_ncc_tc"
        got += [tr.step([frames[k]])[0] for k in range(cut, n)]
        _, templ = tr.get_state(0)
    Hp.check_records(records_of(np.array(got)), g["records"], "small (tc_global -> tc)")
    assert np.array_equal(templ, g["templ"])

"""GPU (-m gpu): parity of the CUDA path, called through the C ABI, with the CPU oracle and the
cv2 4.13.0 golden vectors -- gates G1..G6 of SURVEY.md §8(c).  Nothing here reads /root/reference."""
import importlib

import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as Hp

pvt = importlib.import_module("parallel-video-object-tracker_b200")
pytestmark = pytest.mark.gpu

CLIPS = ["small", "lowtex", "lost", "fade", "border", "flat", "oddsize", "c1_standin", "c2_1080p", "c3_4k"]


def records_of(res):
    """RESULT_DTYPE[n] -> [n, 7] float64 like the golden records (x y w h conf moved updated)."""
    return np.stack([res["x"], res["y"], res["w"], res["h"], res["conf"].astype(np.float64), res["moved"], res["updated"]], 1).astype(np.float64)


def run_clip(frames, roi, **params):
    recs, templ = pvt.track_clip(frames, roi, **params)
    return records_of(recs), templ


# ---- G6 ingest ---------------------------------------------------------------------------------
def test_ingest_bit_exact_golden_vectors():
    g = Hp.golden("ingest.npz")
    with pvt.Tracker(8192, 1, 1, 1) as tr:
        out = tr.to_gray_f32(g["bgr"])
    assert np.array_equal(out, g["lut"][g["gray"]])
    with pvt.Tracker(256, 1, 1, 1) as tr:
        assert np.array_equal(tr.to_gray_f32(np.arange(256, dtype=np.uint8).reshape(1, 256))[0], g["lut"])


@pytest.mark.parametrize("W,H", [(320, 240), (301, 233), (1920, 1080), (67, 5)])
def test_ingest_matches_oracle(W, H):
    rng = np.random.default_rng(W * 7 + H)
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    with pvt.Tracker(W, H, 1, 1) as tr:
        assert np.array_equal(tr.to_gray_f32(bgr), O.to_gray_f32(bgr))
        # strided rows (cv::Mat::step > cols*3) and the gray / f32 input formats
        wide = np.zeros((H, W + 5, 3), np.uint8)
        wide[:, :W] = bgr
        assert np.array_equal(tr.to_gray_f32(wide[:, :W]), O.to_gray_f32(bgr))
        g8 = O.bgr2gray(bgr)
        assert np.array_equal(tr.to_gray_f32(g8), O.gray_to_f32(g8))
        f32 = O.to_gray_f32(bgr)
        assert np.array_equal(tr.to_gray_f32(f32), f32)


# ---- map-level operators (a3..a8) -----------------------------------------------------------------
@pytest.mark.parametrize("fn", ["ncc_match_naive_cuda", "ncc_match_shared_cuda", "ncc_match_const", "ncc_match_const_tiled"])
def test_map_level_operators_vs_cv2_and_oracle(fn):
    g = Hp.golden("maps.npz")
    m = getattr(pvt, fn)(g["frame"], g["templ"])
    assert m.shape == g["full_ipp_off"].shape
    assert np.abs(m - g["full_ipp_off"]).max() <= Hp.TOL_SCORE
    assert np.abs(m - g["full_ipp_on"]).max() <= Hp.TOL_SCORE
    assert np.abs(m - O.ncc_match_cpu(g["frame"], g["templ"])).max() <= 2e-5
    assert np.argmax(m) == np.argmax(g["full_ipp_on"])


def test_map_level_degenerate_cells_identical():
    g = Hp.golden("maps.npz")
    f, t = g["frame"], g["templ"]
    ones = pvt.ncc_match_naive_cuda(f, np.full((13, 17), 0.25, np.float32))
    assert np.array_equal(ones, g["flat_templ_on"])                       # flat template: all ones
    z = pvt.ncc_match_naive_cuda(np.full_like(f, 0.5), t)
    assert np.array_equal(z, g["flat_frame_on"]) and np.all(z == 0)       # flat windows: exactly 0
    s = pvt.ncc_match_naive_cuda(f, f.copy())                            # template == frame: 1x1 map
    assert s.shape == (1, 1) and abs(float(s[0, 0]) - 1.0) <= 1e-6


def test_map_level_exact_ties_pick_lowest_index():
    g = Hp.golden("maps.npz")
    m = pvt.ncc_match_naive_cuda(g["tie_frame"], g["tie_templ"])
    assert np.abs(m - g["tie_map"]).max() <= Hp.TOL_SCORE
    # identical windows give bit-identical scores on the GPU too (same operation order for every candidate)
    ties = np.argwhere(g["tie_map"] == g["tie_map"].max())
    vals = m[ties[:, 0], ties[:, 1]]
    assert np.all(vals == vals[0]) and vals[0] == m.max()
    # ...and the tracker-level peak pick returns the first of them in row-major order
    with pvt.Tracker(96, 80, 17, 13, search_radius_x=96, search_radius_y=80, ncc_min_confidence=2.0) as tr:
        tr.init_track(0, g["tie_frame"], (5, 3, 17, 13))
        tr.set_state(0, (40, 30, 17, 13), g["tie_templ"])
        r = tr.step([g["tie_frame"]])[0]
    assert r["moved"] == 0 and (r["x"], r["y"]) == (40, 30)
    with pvt.Tracker(96, 80, 17, 13, search_radius_x=96, search_radius_y=80) as tr:
        tr.init_track(0, g["tie_frame"], (5, 3, 17, 13))
        tr.set_state(0, (40, 30, 17, 13), g["tie_templ"])
        r = tr.step([g["tie_frame"]])[0]
    assert (r["x"], r["y"]) == (int(g["tie_best"][1]), int(g["tie_best"][2])) == (5, 3)


def test_map_level_batched_and_large_templates():
    g = Hp.golden("maps.npz")
    outs = pvt.ncc_match_naive_cuda_batched([g["frame"], g["frame"][::-1].copy(), g["frame"]], g["templ"])
    assert np.array_equal(outs[0], outs[2])
    assert np.abs(outs[0] - g["full_ipp_off"]).max() <= Hp.TOL_SCORE
    assert np.abs(outs[1] - O.ncc_match_cpu(g["frame"][::-1].copy(), g["templ"])).max() <= 2e-5
    # 72x72 = 5184 px > the reference's 4096-px constant-memory limit (baseline_kernel.cu:499-500)
    rng = np.random.default_rng(3)
    f = O.gray_to_f32(rng.integers(0, 256, (150, 170), dtype=np.uint8))
    t = f[20:92, 31:103].copy()
    m = pvt.ncc_match_const(f, t)
    assert np.abs(m - O.ncc_match_cpu(f, t)).max() <= 2e-5 and np.unravel_index(np.argmax(m), m.shape) == (20, 31)


# ---- tracker level: G1 / G2 / G5 on every clip -----------------------------------------------------
@pytest.mark.parametrize("name", CLIPS)
def test_clip_trajectory_vs_cv2_golden(name):
    (c, tk) = Hp.clip(name)
    g = Hp.golden(f"clip_{name}.npz")
    rec, templ = run_clip(c["frames"], c["roi"], search_radius_x=tk.get("rx", 80), search_radius_y=tk.get("ry", 80))
    Hp.check_records(rec, g["records"], name)
    assert np.array_equal(templ, g["templ"]), f"{name}: final template not bit-identical"


@pytest.mark.parametrize("name", ["small", "lost", "fade", "oddsize"])
def test_clip_trajectory_vs_oracle_and_direct_kernel(name):
    (c, tk) = Hp.clip(name)
    kw = dict(search_radius_x=tk.get("rx", 80), search_radius_y=tk.get("ry", 80))
    want, wt = O.track_clip(c["frames"], c["roi"], rx=kw["search_radius_x"], ry=kw["search_radius_y"])
    rec, templ = run_clip(c["frames"], c["roi"], **kw)
    d = Hp.check_records(rec, want[:, :7], name)
    assert d <= 2e-5 and np.array_equal(templ, wt)
    rec2, templ2 = run_clip(c["frames"], c["roi"], kernel=pvt.KERNEL_DIRECT, **kw)
    # the verification kernel: same boxes, flags and template; scores equal up to the summation order (a single-track
    # context splits the template over several CTAs -- K-split -- which regroups the FP32 partial sums)
    assert np.array_equal(rec[:, [0, 1, 2, 3, 5, 6]], rec2[:, [0, 1, 2, 3, 5, 6]]) and np.array_equal(templ, templ2)
    assert np.abs(rec[:, 4] - rec2[:, 4]).max() <= 2e-6


def test_batch_mode_hold_semantics():
    (c, tk) = Hp.clip("batch4")
    g = Hp.golden("clip_batch4.npz")
    rec, templ = run_clip(c["frames"], c["roi"], mode=pvt.MODE_BATCH, batch_size=4)
    Hp.check_records(rec, g["records"], "batch4")
    assert np.array_equal(templ, g["templ"])


# ---- G3 / G4 window maps ---------------------------------------------------------------------------
@pytest.mark.parametrize("name,k", [("small", 1), ("small", 7), ("lowtex", 2), ("border", 3), ("flat", 1), ("oddsize", 2), ("c2_1080p", 1)])
def test_window_map(name, k):
    (c, tk) = Hp.clip(name)
    g = Hp.golden(f"clip_{name}.npz")
    frames, roi = c["frames"], c["roi"]
    H, W = frames.shape[1:3]
    templ, bbox = g[f"map{k}_templ"], g[f"map{k}_bbox"]
    with pvt.Tracker(W, H, roi[2], roi[3], keep_maps=1, search_radius_x=tk.get("rx", 80), search_radius_y=tk.get("ry", 80)) as tr:
        tr.init_track(0, frames[0], roi)
        tr.set_state(0, (int(bbox[0]), int(bbox[1]), roi[2], roi[3]), templ)
        tr.step([frames[k]])
        m, win = tr.window_map(0)
        for kern in (pvt.KERNEL_DIRECT,):
            tr.set_params(kernel=kern)
            tr.set_state(0, (int(bbox[0]), int(bbox[1]), roi[2], roi[3]), templ)
            tr.step([frames[k]])
            m2, _ = tr.window_map(0)
            assert np.abs(m - m2).max() <= 5e-5 and np.argmax(m) == np.argmax(m2)
    assert win == tuple(int(v) for v in g[f"map{k}_win"])
    off, on = g[f"map{k}_ipp_off"], g[f"map{k}_ipp_on"]
    gray = O.to_gray_f32(frames[k])
    sig = Hp.window_sigma(gray, roi[2], roi[3], win)
    d = np.abs(m - off)
    assert d[sig >= 0.002].max(initial=0) <= Hp.TOL_SCORE            # G3 vs the exact (IPP-off) oracle
    assert d[sig < 0.002].max(initial=0) <= Hp.TOL_LOWVAR
    assert np.abs(m - on)[sig >= 0.02].max(initial=0) <= Hp.TOL_SCORE  # G3 vs the default build
    deg = (off == 0) | (np.abs(off) == 1)
    assert np.array_equal(m[deg], off[deg])                           # G4
    assert np.argmax(m) == np.argmax(on)                              # G1


# ---- gates are compared in double (main.cpp:153,157) ------------------------------------------------
def test_thresholds_compare_in_double():
    (c, _) = Hp.clip("small")
    frames, roi = c["frames"], c["roi"]
    conf = run_clip(frames[:2], roi)[0][0, 4]
    cf = float(np.float32(conf))
    up = float(np.nextafter(np.float64(cf), 2.0))          # just above the score, rounds to the same float
    assert np.float32(up) == np.float32(cf)
    r = run_clip(frames[:2], roi, ncc_strong_confidence=cf)[0][0]
    assert r[5] == 1 and r[6] == 1                         # conf >= conf
    r = run_clip(frames[:2], roi, ncc_strong_confidence=up)[0][0]
    assert r[5] == 1 and r[6] == 0                         # a float compare would still update here
    r = run_clip(frames[:2], roi, ncc_min_confidence=up, ncc_strong_confidence=up)[0][0]
    assert r[5] == 0 and r[6] == 0 and (r[0], r[1]) == (roi[0], roi[1])


# ---- batching across tracks / streams, async submission ----------------------------------------------
def test_multi_stream_multi_track_matches_single_runs():
    names = ["small", "lost", "border"]
    clips = [Hp.clip(n)[0] for n in names]
    n = min(len(c["frames"]) for c in clips)
    H, W = clips[0]["frames"].shape[1:3]
    singles = [run_clip(c["frames"][:n], c["roi"])[0] for c in clips]
    with pvt.Tracker(W, H, 32, 32, max_streams=3, max_tracks=5) as tr:
        # tracks 0..2: one per stream; tracks 3,4: second and third object on stream 0 (multi-ROI, config C4 style)
        for i, c in enumerate(clips):
            tr.init_track(i, c["frames"][0], c["roi"], stream=i)
        x, y, w, h = clips[0]["roi"]
        tr.init_track(3, None, (x + 9, y - 7, w, h), stream=0)
        tr.init_track(4, None, (x, y, w, h), stream=0)
        out = []
        for k in range(1, n):
            out.append(tr.step([c["frames"][k] for c in clips]))
    out = np.array(out)
    def same(a, b):   # boxes and flags identical; scores up to the summation order (contexts of different size plan
        # the template split differently, which regroups the FP32 partial sums)
        return np.array_equal(a[:, [0, 1, 2, 3, 5, 6]], b[:, [0, 1, 2, 3, 5, 6]]) and np.abs(a[:, 4] - b[:, 4]).max() <= 2e-6
    for i in range(3):
        assert same(records_of(out[:, i]), singles[i])
    assert np.array_equal(records_of(out[:, 4]), records_of(out[:, 0]))   # same ROI, same stream, same context -> bit-identical
    assert same(records_of(out[:, 4]), singles[0])
    assert np.all(out[:, 3]["valid"] == 1)
    # a stream that gets no frame in a step is not stepped
    with pvt.Tracker(W, H, 32, 32, max_streams=2, max_tracks=2) as tr:
        tr.init_track(0, clips[0]["frames"][0], clips[0]["roi"], stream=0)
        tr.init_track(1, clips[1]["frames"][0], clips[1]["roi"], stream=1)
        f = pvt.host_frame(clips[0]["frames"][1], 0)
        r = tr.step([f])
    assert r[0]["searched"] == 1 and r[1]["searched"] == 0 and r[1]["valid"] == 1 and np.isnan(r[1]["conf"])


def test_async_submit_collect_equals_step():
    (c, _) = Hp.clip("small")
    frames, roi = c["frames"], c["roi"]
    want = run_clip(frames, roi)[0]
    H, W = frames.shape[1:3]
    with pvt.Tracker(W, H, 32, 32) as tr:
        tr.init_track(0, frames[0], roi)
        keep = [tr.submit([frames[k]]) for k in range(1, len(frames))]
        got = tr.collect(len(frames) - 1)
        assert tr.launch_count() >= 2 * (len(frames) - 1)
    assert np.array_equal(records_of(got[:, 0]), want)
    assert list(got[:, 0]["step"]) == list(range(len(frames) - 1))


def test_device_resident_frames():
    torch = pytest.importorskip("torch")
    (c, _) = Hp.clip("small")
    frames, roi = c["frames"], c["roi"]
    want = run_clip(frames, roi)[0]
    H, W = frames.shape[1:3]
    dev = torch.from_numpy(frames).cuda()
    torch.cuda.synchronize()
    with pvt.Tracker(W, H, 32, 32) as tr:
        tr.init_track(0, pvt.device_frame(dev[0].data_ptr(), W * 3), roi)
        for k in range(1, len(frames)):
            tr.submit([pvt.device_frame(dev[k].data_ptr(), W * 3)])
        got = tr.collect(len(frames) - 1)
    assert np.array_equal(records_of(got[:, 0]), want)


@pytest.mark.parametrize("memory,collect_every", [("device", 0), ("device", 5), ("device", 23), ("pinned", 0), ("pinned", 6), ("pageable", 4)])
def test_submit_sequence_fast_paths_equal_step_by_step(memory, collect_every):
    """pvt_submit_sequence over a frame ring: device-resident and pinned-host rings take the bare-graph-launch path (several
    time steps per launch, pinned frames read zero-copy by the ROI ingest), pageable frames the staged path; read-backs may
    fall anywhere relative to the multi-step launches.  All must reproduce the step-by-step records."""
    torch = pytest.importorskip("torch")
    (c, _) = Hp.clip("small")
    frames, roi = c["frames"], c["roi"]
    want = run_clip(frames, roi)[0]
    n, H, W, _ = frames.shape
    if memory == "device":
        buf = torch.from_numpy(frames).cuda()
        torch.cuda.synchronize()
    elif memory == "pinned":
        buf = torch.from_numpy(frames).pin_memory()
    else:
        buf = torch.from_numpy(frames.copy())
    mem = pvt.MEM_DEVICE if memory == "device" else pvt.MEM_HOST
    ring = [[pvt.Frame(0, pvt.FMT_BGR8, mem, 0, buf[k].data_ptr(), W * 3)] for k in range(1, n)]
    with pvt.Tracker(W, H, 32, 32, ingest=pvt.INGEST_ROI) as tr:
        tr.init_track(0, frames[0], roi)
        if collect_every:
            k = ((n - 1) // collect_every) * collect_every
            res = tr.submit_sequence(k, ring, collect_every=collect_every, want_results=True)
            assert np.array_equal(records_of(res[:, 0]), want[:k])
            if k < n - 1:
                tr.submit_sequence(n - 1 - k, ring[k:] + ring[:k])
        else:
            tr.submit_sequence(n - 1, ring)
        got = tr.collect(n - 1)
    assert np.array_equal(records_of(got[:, 0]), want)
    assert list(got[:, 0]["step"]) == list(range(n - 1))


@pytest.mark.parametrize("radius", [80, 24])
def test_pinned_ring_prefetch_and_its_fallback(radius):
    """Pinned host rings: k_prefetch_roi stages the next step's tile (current tile grown by R/2) while the step computes.
    radius 80: the object moves 20 px per frame, the staged region always covers the next tile.  radius 24: it moves more
    than R/2, the coverage check fails and the ingest falls back to the zero-copy read.  Same records either way."""
    torch = pytest.importorskip("torch")
    (c, _) = Hp.clip("small")
    frames, roi = c["frames"], c["roi"]
    want = run_clip(frames, roi, search_radius_x=radius, search_radius_y=radius)[0]
    assert want[:, 5].all()                      # the object stays inside the window in both cases
    n, H, W, _ = frames.shape
    buf = torch.from_numpy(frames).pin_memory()
    ring = [[pvt.Frame(0, pvt.FMT_BGR8, pvt.MEM_HOST, 0, buf[k].data_ptr(), W * 3)] for k in range(1, n)]
    with pvt.Tracker(W, H, 32, 32, ingest=pvt.INGEST_ROI, search_radius_x=radius, search_radius_y=radius) as tr:
        tr.init_track(0, frames[0], roi)
        tr.submit_sequence(n - 1, ring)
        got = tr.collect(n - 1)
        # a second pass over the same ring from a box that was moved by hand: the staged region of the last step does not
        # cover it; nothing may be taken from the stale staging buffer
        tr.set_state(0, (roi[0], roi[1], 32, 32), None)
        tr.submit_sequence(3, ring)
        again = tr.collect(3)
    assert np.array_equal(records_of(got[:, 0]), want)
    assert np.array_equal(records_of(again[:, 0])[:, :4], want[:3, :4])


def test_pinned_buffers_refilled_between_sequences_are_not_served_from_the_stage():
    """The staging buffer is keyed by (step, frame pointer): a caller that refills its pinned buffers between two calls, or
    goes on with per-frame submits, must get the NEW pixels."""
    torch = pytest.importorskip("torch")
    (c, _) = Hp.clip("small")
    frames, roi = c["frames"], c["roi"]
    want = run_clip(frames, roi)[0]
    n, H, W, _ = frames.shape
    buf = torch.empty((4, H, W, 3), dtype=torch.uint8).pin_memory()      # a 4-frame pinned ring, refilled by the caller
    ring = [[pvt.Frame(0, pvt.FMT_BGR8, pvt.MEM_HOST, 0, buf[k].data_ptr(), W * 3)] for k in range(4)]
    got = []
    with pvt.Tracker(W, H, 32, 32, ingest=pvt.INGEST_ROI) as tr:
        tr.init_track(0, frames[0], roi)
        k = 1
        while k + 4 <= n:
            buf.copy_(torch.from_numpy(frames[k:k + 4]))                   # same pointers, new content
            tr.submit_sequence(4, ring)
            got.append(tr.collect(4))
            k += 4
        for j in range(k, n):                                              # then frame by frame through the same buffer
            buf[0].copy_(torch.from_numpy(frames[j]))
            got.append(tr.step([ring[0][0]])[None])
    got = np.concatenate(got)
    assert np.array_equal(records_of(got[:, 0]), want)


def test_state_roundtrip_and_errors():
    (c, _) = Hp.clip("small")
    frames, roi = c["frames"], c["roi"]
    H, W = frames.shape[1:3]
    with pvt.Tracker(W, H, 32, 32, max_tracks=2) as tr:
        with pytest.raises(pvt.PvtError) as e:
            tr.get_state(0)
        assert e.value.code == pvt.ERR_STATE
        tr.init_track(0, frames[0], roi)
        bbox, templ = tr.get_state(0)
        assert bbox == tuple(roi)
        assert np.array_equal(templ, O.to_gray_f32(frames[0])[roi[1]:roi[1] + roi[3], roi[0]:roi[0] + roi[2]])
        for bad in [(-1, 0, 32, 32), (0, 0, 33, 32), (W - 31, 0, 32, 32)]:
            with pytest.raises(pvt.PvtError) as e:
                tr.init_track(1, None, bad)
            assert e.value.code == pvt.ERR_INVALID
        with pytest.raises(pvt.PvtError) as e:
            tr.step([frames[1][:, : W - 1]])           # wrong geometry is caught as a short row
        assert e.value.code == pvt.ERR_INVALID
        with pytest.raises(pvt.PvtError):
            tr.window_map(0)                            # keep_maps was not requested
        # checkpoint / resume: state moved to a fresh context continues identically
        tr.step([frames[1]])
        bbox, templ = tr.get_state(0)
        a = tr.step([frames[2]])[0]
    with pvt.Tracker(W, H, 32, 32, max_tracks=2) as tr2:   # same context shape -> same plan -> bit-identical scores
        tr2.init_track(0, frames[0], roi)
        tr2.set_state(0, bbox, templ)
        b = tr2.step([frames[2]])[0]
    assert records_of(np.array([a])).tolist() == records_of(np.array([b])).tolist()


# ---- size-independent properties at BASELINE.json's full sizes ------------------------------------
@pytest.mark.parametrize("W,H,tw,th,R", [(1920, 1080, 64, 64, 80), (3840, 2160, 128, 128, 160)])
def test_full_size_properties(W, H, tw, th, R):
    rng = np.random.default_rng(W)
    base = rng.integers(0, 256, (H + 64, W + 64), dtype=np.uint8)
    # smooth a little so the correlation surface has a clear single peak
    img = ((base[:-1, :-1].astype(np.uint16) + base[1:, :-1] + base[:-1, 1:] + base[1:, 1:]) // 4).astype(np.uint8)
    f0 = np.ascontiguousarray(img[:H, :W])
    with pvt.Tracker(W, H, tw, th, keep_maps=1, search_radius_x=R, search_radius_y=R) as tr:
        for (x, y) in [(W // 2, H // 2), (3, 5), (W - tw - 2, H - th - 1)]:
            tr.init_track(0, f0, (x, y, tw, th))
            # self-match: the same frame again -> peak stays put with score 1 (within rounding)
            r = tr.step([f0])[0]
            assert (r["x"], r["y"]) == (x, y) and abs(r["conf"] - 1.0) <= 1e-5 and r["updated"] == 1
            m, win = tr.window_map(0)
            assert m.shape == (win[3], win[2]) and np.all(np.abs(m) <= 1.0) and not np.isnan(m).any()
            assert m[y - win[1], x - win[0]] == m.max()
            # shift property: a translated frame moves the peak by the same vector
            dx, dy = (17, -11) if 100 < x < W - tw - 100 else ((9, 6) if x < 100 else (-9, -6))
            f1 = np.zeros_like(f0)
            f1[max(dy, 0):H + min(dy, 0), max(dx, 0):W + min(dx, 0)] = f0[max(-dy, 0):H - max(dy, 0), max(-dx, 0):W - max(dx, 0)]
            r = tr.step([f1])[0]
            assert (r["x"], r["y"]) == (x + dx, y + dy) and r["conf"] > 0.99


# ---- ingest modes: whole frames (the reference's toGrayF32) vs search tiles only, incl. zero-copy pinned host frames
@pytest.mark.parametrize("name", ["small", "border", "lost"])
def test_roi_ingest_equals_full_ingest(name):
    (c, tk) = Hp.clip(name)
    frames, roi = c["frames"], c["roi"]
    full = run_clip(frames, roi, ingest=pvt.INGEST_FULL)
    part = run_clip(frames, roi, ingest=pvt.INGEST_ROI)
    assert np.array_equal(full[0], part[0]) and np.array_equal(full[1], part[1])
    # pinned host frames: k_ingest_roi reads them over PCIe directly (no staging copy)
    n, H, W, _ = frames.shape
    pin = pvt.PinnedBuffer(frames.nbytes)
    host = pin.array.reshape(frames.shape)
    host[...] = frames
    with pvt.Tracker(W, H, roi[2], roi[3], ingest=pvt.INGEST_ROI) as tr:
        tr.init_track(0, host[0], roi)
        keep = [tr.submit([host[k]]) for k in range(1, n)]
        got = tr.collect(n - 1)
    assert np.array_equal(records_of(got[:, 0]), full[0])
    pin.free()


# ---- window origin / clamp / fringe sweep -------------------------------------------------------------
# k_ncc_search lays its thread tiles out from the window origin: the TMA tile is fetched from the origin rounded down to
# 4 pixels and shifted in shared memory (all four alignments must give the same map), a remainder of exactly one row /
# column of the window is computed by k_ncc_fringe (present only when the window is not clamped), and the K-split shape
# (one track) and the throughput shape (many tracks) sum in different places.  Every position is checked against the
# oracle's window map, in both shapes.
@pytest.mark.parametrize("R,tw,th", [(40, 32, 32), (20, 24, 19), (12, 37, 29)])
def test_window_origin_clamp_and_fringe_sweep(R, tw, th):
    W, H = 333, 251
    rng = np.random.default_rng(R * 1000 + tw)
    base = rng.integers(0, 256, (H + 16, W + 16), dtype=np.uint8)
    frame = np.stack([np.roll(base, s, axis=(0, 1))[:H, :W] for s in (0, 1, 2)], -1).copy()
    nxt = np.clip(frame.astype(np.int16) + rng.integers(-3, 4, frame.shape), 0, 255).astype(np.uint8)
    gray = O.to_gray_f32(nxt)
    outW, outH = W - tw + 1, H - th + 1
    xs = [0, 1, 2, 3, R - 1, R, R + 1, R + 2, R + 3, 150, 151, 152, 153, outW - 1 - R, outW - R, outW - 1]
    ys = [0, 3, R, R + 1, 100, outH - 1 - R, outH - R + 5, outH - 1]
    pos = [(x, y) for x in xs for y in ys if 0 <= x < outW and 0 <= y < outH]
    n = len(pos)

    def run(max_tracks, chunk):
        maps = []
        with pvt.Tracker(W, H, tw, th, max_streams=1, max_tracks=max_tracks, keep_maps=1, search_radius_x=R, search_radius_y=R) as tr:
            for i0 in range(0, n, chunk):
                grp = pos[i0:i0 + chunk]
                for t, (x, y) in enumerate(grp):
                    tr.init_track(t, frame if t == 0 else None, (x, y, tw, th))
                for t in range(len(grp), max_tracks):
                    if i0:
                        tr.remove_track(t)
                tr.step([nxt])
                for t in range(len(grp)):
                    maps.append(tr.window_map(t))
        return maps

    import os
    single = run(1, 1)          # latency shape (K-split, the remainder stays in the grid as masked lanes)
    os.environ["PVT_PLAN"] = "%d,1,1,1" % ((2 * R + 1) // 5)   # throughput shape as planned for a filled GPU: unsplit, remainder
    try:                                                        # row / column in k_ncc_fringe behind the search (PDL)
        many = run(48, 48)
        os.environ["PVT_PLAN"] = "%d,2,3,1" % ((2 * R + 1) // 5)   # K-split ON TOP of the fringe geometry (deferred fringe)
        forced = run(48, 48)
    finally:
        del os.environ["PVT_PLAN"]
    templ_of = lambda x, y: O.to_gray_f32(frame)[y:y + th, x:x + tw].copy()
    seen_fringe = seen_clamped = 0
    for (x, y), (m1, w1), (m2, w2), (m3, w3) in zip(pos, single, many, forced):
        win = O.search_window(x, y, tw, th, outW, outH, R, R)
        assert w1 == win and w2 == win and w3 == win
        want = O.ncc_window(gray, templ_of(x, y), *win)
        sig = Hp.window_sigma(gray, tw, th, win)
        for m in (m1, m2, m3):
            d = np.abs(m - want)
            assert d[sig >= 0.002].max(initial=0) <= Hp.TOL_SCORE, (x, y)
            assert np.argmax(m) == np.argmax(want), (x, y)
        assert np.abs(m1 - m2).max() <= 2e-5
        seen_fringe += win[2] == 2 * R + 1 and win[3] == 2 * R + 1
        seen_clamped += win[2] < 2 * R + 1 or win[3] < 2 * R + 1
    assert seen_fringe >= 4 and seen_clamped >= 20


# ---- §8(f) n4: the eps formula of the reference's CUDA kernels (pvt_formula) ----------------------------------------
def _textured(rng, H, W):
    from scipy.ndimage import gaussian_filter
    f = gaussian_filter(rng.random((H, W)), 2.0)
    return ((f - f.min()) / (f.max() - f.min())).astype(np.float32)


@pytest.mark.parametrize("fn", ["ncc_match_naive_cuda", "ncc_match_shared_cuda", "ncc_match_const", "ncc_match_const_tiled"])
def test_eps_formula_map_operators(fn):
    rng = np.random.default_rng(11)
    f = _textured(rng, 150, 210)
    t = f[40:72, 90:130].copy() + rng.normal(0, 0.01, (32, 40)).astype(np.float32)
    m = getattr(pvt, fn)(f, t, formula=pvt.FORMULA_EPS)
    exact = O.ncc_eps_exact(f, t, 0, 0, m.shape[1], m.shape[0])
    seq = O.ncc_window_eps(f, t, 0, 0, m.shape[1], m.shape[0])           # FP32 sequential, like baseline_kernel.cu:21-64
    assert np.abs(m - exact).max() <= 2e-5                                # FP64 window sums + blocked FP32 cross term
    assert np.abs(m - seq).max() <= Hp.TOL_SCORE
    assert np.argmax(m) == np.argmax(seq) == np.argmax(exact)
    # and the default formula is untouched by a context of the other kind living next to it
    assert np.abs(getattr(pvt, fn)(f, t) - O.ncc_match_cpu(f, t)).max() <= 2e-5


def test_eps_formula_degenerate_cells_follow_the_kernel_not_opencv():
    rng = np.random.default_rng(12)
    f = _textured(rng, 90, 120)
    flat_t = pvt.ncc_match_naive_cuda(f, np.full((16, 24), 0.25, np.float32), formula=pvt.FORMULA_EPS)
    assert np.abs(flat_t).max() <= 1e-3 and np.all(np.isfinite(flat_t))    # cov == 0 (OpenCV's rule says 1 everywhere)
    t = f[10:26, 30:54].copy()
    flat_w = pvt.ncc_match_naive_cuda(np.full_like(f, 0.5), t, formula=pvt.FORMULA_EPS)
    want = O.ncc_eps_exact(np.full_like(f, 0.5), t, 0, 0, flat_w.shape[1], flat_w.shape[0])
    assert np.abs(flat_w - want).max() <= 1e-4 and np.abs(flat_w).max() <= 1e-2   # the 1e-3 floor on sigma_w, not a 0/0
    outs = pvt.ncc_match_naive_cuda_batched([f, f[::-1].copy()], t, formula=pvt.FORMULA_EPS)
    assert np.abs(outs[0] - O.ncc_eps_exact(f, t, 0, 0, outs[0].shape[1], outs[0].shape[0])).max() <= 2e-5
    assert np.abs(outs[1] - O.ncc_eps_exact(f[::-1], t, 0, 0, outs[1].shape[1], outs[1].shape[0])).max() <= 2e-5


@pytest.mark.parametrize("name", ["small", "oddsize", "c1_standin"])
def test_eps_formula_tracker_follows_the_gpu_mode_loop(name):
    """main.cpp:103-161 fed by a GPU-mode map: same peaks, gates and EMA as the oracle loop run on the eps map."""
    (c, tk) = Hp.clip(name)
    kw = dict(search_radius_x=tk.get("rx", 80), search_radius_y=tk.get("ry", 80))
    with O.formula(1):
        want, wt = O.track_clip(c["frames"], c["roi"], rx=kw["search_radius_x"], ry=kw["search_radius_y"])
    rec, templ = run_clip(c["frames"], c["roi"], formula=pvt.FORMULA_EPS, **kw)
    Hp.check_records(rec, want[:, :7], name + " (eps)")
    assert np.array_equal(templ, wt)
    # the formula is a creation-time choice
    with pvt.Tracker(64, 64, 8, 8) as tr:
        with pytest.raises(pvt.PvtError):
            tr.set_params(formula=pvt.FORMULA_EPS)


# ---- random geometries under the planner's own choices (no PVT_PLAN): unequal radii, odd templates, boxes anywhere -----
@pytest.mark.parametrize("seed", range(24))
def test_random_geometries_vs_oracle(seed):
    rng = np.random.default_rng(4242 + seed)
    W, H = int(rng.integers(90, 420)), int(rng.integers(70, 300))
    tw, th = int(rng.integers(3, min(72, W // 2))), int(rng.integers(3, min(72, H // 2)))
    rx, ry = int(rng.integers(1, 70)), int(rng.integers(1, 70))
    n_tracks = int(rng.choice([1, 1, 3, 7]))
    from scipy.ndimage import gaussian_filter
    base = gaussian_filter(rng.random((H, W, 3)), (1.2, 1.2, 0))
    f0 = np.clip((base - base.min()) / (base.max() - base.min()) * 255, 0, 255).astype(np.uint8)
    f1 = np.clip(np.roll(f0, (int(rng.integers(-2, 3)), int(rng.integers(-2, 3))), (0, 1)).astype(np.int16)
                 + rng.integers(-3, 4, f0.shape), 0, 255).astype(np.uint8)
    outW, outH = W - tw + 1, H - th + 1
    boxes = [(int(rng.integers(0, outW)), int(rng.integers(0, outH))) for _ in range(n_tracks)]
    boxes[0] = [(0, 0), (outW - 1, outH - 1), (outW // 2, outH // 2)][seed % 3]        # corners and the middle
    g0, g1 = O.to_gray_f32(f0), O.to_gray_f32(f1)
    with pvt.Tracker(W, H, tw, th, max_streams=1, max_tracks=n_tracks, keep_maps=1, search_radius_x=rx, search_radius_y=ry) as tr:
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
            if want.size == 1 or gap[1] - gap[0] >= Hp.AMBIGUOUS_GAP:                          # G1, unless the oracle itself is ambiguous
                assert (res[t]["x"], res[t]["y"]) == (rec.x, rec.y), (seed, t)
                assert (res[t]["moved"], res[t]["updated"]) == (rec.moved, rec.updated)
                assert abs(float(res[t]["conf"]) - rec.conf) <= Hp.TOL_SCORE
                _, got_t = tr.get_state(t)
                assert np.array_equal(got_t, templ), (seed, t)                                  # EMA bit-exact given the same peak


# ---- round-2 regressions ------------------------------------------------------------------------------------
def test_prefetch_branch_that_starts_late_still_stages_the_right_frame(monkeypatch):
    """k_prefetch_roi runs on a low-priority branch that joins at the END of the step, i.e. after the update has advanced the
    device step counter.  It must take its step from what k_ingest_roi recorded, not from the counter: a CTA that starts late
    (here: forced by a 60 us spin, longer than a whole step) would otherwise stage frame k+2 under the tag of frame k+1."""
    torch = pytest.importorskip("torch")
    (c, _) = Hp.clip("small")
    frames, roi = c["frames"], c["roi"]
    want = run_clip(frames, roi)[0]
    n, H, W, _ = frames.shape
    buf = torch.from_numpy(frames).pin_memory()
    ring = [[pvt.Frame(0, pvt.FMT_BGR8, pvt.MEM_HOST_PINNED, 0, buf[k].data_ptr(), W * 3)] for k in range(1, n)]
    monkeypatch.setenv("PVT_DEBUG_PREFETCH_DELAY_US", "60")
    with pvt.Tracker(W, H, 32, 32, ingest=pvt.INGEST_ROI) as tr:
        tr.init_track(0, frames[0], roi)
        tr.submit_sequence(n - 1, ring)
        got = tr.collect(n - 1)
    assert np.array_equal(records_of(got[:, 0]), want)


def test_track_init_from_current_image_needs_a_complete_plane():
    """main.cpp:70-71 cuts the template from a fully converted frame.  On the ROI ingest only the search tiles are refreshed,
    so pvt_track_init(frame0 = NULL) after a step must fail (PVT_ERR_STATE) instead of cutting stale pixels; on the full ingest
    it keeps working and cuts from the CURRENT frame."""
    (c, _) = Hp.clip("small")
    frames, roi = c["frames"], c["roi"]
    H, W = frames.shape[1:3]
    with pvt.Tracker(W, H, 32, 32, max_tracks=2, ingest=pvt.INGEST_ROI) as tr:
        tr.init_track(0, frames[0], roi)
        tr.init_track(1, None, (5, 5, 32, 32))            # right after a full ingest: fine
        tr.step([frames[1]])
        with pytest.raises(pvt.PvtError) as e:
            tr.init_track(1, None, (200, 150, 32, 32))
        assert e.value.code == pvt.ERR_STATE
        tr.init_track(1, frames[1], (200, 150, 32, 32))   # with the frame: fine
        _, t = tr.get_state(1)
        assert np.array_equal(t, O.to_gray_f32(frames[1])[150:182, 200:232])
    with pvt.Tracker(W, H, 32, 32, max_tracks=2, ingest=pvt.INGEST_FULL) as tr:
        tr.init_track(0, frames[0], roi)
        tr.step([frames[1]])
        tr.init_track(1, None, (200, 150, 32, 32))
        _, t = tr.get_state(1)
        assert np.array_equal(t, O.to_gray_f32(frames[1])[150:182, 200:232])


def test_rejected_submit_does_not_shift_the_batch_cadence():
    (c, tk) = Hp.clip("batch4")
    g = Hp.golden("clip_batch4.npz")
    frames, roi = c["frames"], c["roi"]
    H, W = frames.shape[1:3]
    with pvt.Tracker(W, H, 32, 32, mode=pvt.MODE_BATCH, batch_size=4) as tr:
        tr.init_track(0, frames[0], roi)
        out = []
        for k in range(1, len(frames)):
            if k in (2, 4, 7):                             # a bad call in every phase of the cadence
                with pytest.raises(pvt.PvtError):
                    tr.step([frames[k][:, : W - 1]])
            out.append(tr.step([frames[k]])[0])
        rec = records_of(np.array(out))
        Hp.check_records(rec, g["records"], "batch4 with rejected submits")
        # a new cadence (set_params) starts from an empty batch: the next three frames are held, the fourth searched
        tr.set_params(batch_size=3)
        tr.set_params(batch_size=4)
        kinds = [int(tr.step([frames[1]])[0]["searched"]) for _ in range(4)]
        assert kinds == [0, 0, 0, 1]

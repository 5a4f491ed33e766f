"""k_ncc_local -- the single-stream search with the K-split and the window statistics inside the CTA (no partial sums in
global memory, no second-stage kernel): ingest ~> [k_winstats ||] k_ncc_local -> k_update (or, opt-in, the update by the track's last CTA).
Its cross terms follow the accumulation order of a K-split part of k_ncc_search exactly, so against the K-split path forced to
the same parts (PVT_PLAN) the scores may only differ where the FP64 normaliser rounds differently (sum of squares in another
order, ~1e-15 relative): a few cells by one float ulp.  Everything else is held to the same gates as every path.
Reference semantics: tracker/src/main.cpp:135-161, ncc_cpu.cpp:12."""
import importlib
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as Hp
from tests.test_fused_step import env, records_of
from tools import synth

pvt = importlib.import_module("parallel-video-object-tracker_b200")
pytestmark = pytest.mark.gpu


def run(frames, roi, R, n_tracks=1, **e):
    H, W = frames.shape[1:3]
    with env(**e):
        with pvt.Tracker(W, H, roi[2], roi[3], max_tracks=n_tracks, keep_maps=1, search_radius_x=R, search_radius_y=R) as tr:
            for t in range(n_tracks):
                tr.init_track(t, frames[0] if t == 0 else None, (roi[0] + 5 * t, roi[1] - 3 * t, roi[2], roi[3]))
            l0 = tr.launch_count()
            recs, maps = [], []
            for k in range(1, len(frames)):
                recs.append(tr.step([frames[k]]).copy())
                maps.append([tr.window_map(t) for t in range(n_tracks)])
            per_step = (tr.launch_count() - l0) / (len(frames) - 1)
            templ = [tr.get_state(t)[1].copy() for t in range(n_tracks)]
    return np.stack(recs), maps, templ, per_step


def test_local_c2_geometry_vs_ksplit_same_parts_and_oracle():
    c = synth.make_clip(synth.ClipSpec(seed=21, W=1920, H=1080, tw=64, th=64, n_frames=4, R=80))
    frames, roi = c["frames"], c["roi"]
    a, ma, ta, ka = run(frames, roi, 80, PVT_NO_LOCAL=None, PVT_PLAN=None, PVT_LOCAL_STATS="1", PVT_LOCAL_UPDATE="1")   # statistics and update inside the CTA: TR 5, 8 x 6 parts
    b, mb, tb, kb = run(frames, roi, 80, PVT_NO_LOCAL="1", PVT_PLAN="33,8,6,0")   # K-split with the local plan's parts (8 chunks x 6 row parts)
    a2, ma2, ta2, ka2 = run(frames, roi, 80, PVT_NO_LOCAL=None, PVT_PLAN=None)    # default: statistics from k_winstats beside the search (TR 6, 8 x 5 parts)
    b2, mb2, tb2, kb2 = run(frames, roi, 80, PVT_NO_LOCAL="1", PVT_PLAN="33,8,5,0")
    assert ka2 == 4 and kb2 == 4, (ka2, kb2)     # ingest + k_winstats + k_ncc_local + k_update
    assert np.array_equal(a2["conf"].view(np.uint32), b2["conf"].view(np.uint32)) and np.array_equal(ta2[0], tb2[0])
    for k in range(len(ma2)):
        assert np.array_equal(ma2[k][0][0].view(np.uint32), mb2[k][0][0].view(np.uint32))   # same normalisers, same sums: same bits
    assert ka == 2 and kb == 4, (ka, kb)          # ingest + k_ncc_local (statistics and update inside)   vs   the four K-split kernels
    for f in ("x", "y", "moved", "updated"):
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(ta[0], tb[0])
    for k in range(len(ma)):
        (m1, w1), (m2, w2) = ma[k][0], mb[k][0]
        assert w1 == w2
        d = np.abs(m1 - m2)
        assert d.max() <= 1.2e-7 and (d > 0).mean() < 1e-3, (float(d.max()), float((d > 0).mean()))
    # ... and against the oracle on the first searched frame
    g = O.to_gray_f32(frames[1])
    templ = np.ascontiguousarray(O.to_gray_f32(frames[0])[roi[1]:roi[1] + 64, roi[0]:roi[0] + 64])
    (m1, w1) = ma[0][0]
    want = O.ncc_window(g, templ, *w1)
    sig = Hp.window_sigma(g, 64, 64, w1)
    d = np.abs(m1 - want)
    assert d[sig >= 0.002].max(initial=0) <= Hp.TOL_SCORE and d[sig < 0.002].max(initial=0) <= Hp.TOL_LOWVAR
    assert np.argmax(m1) == np.argmax(want)


# (W, H, tw, th, R, tracks): odd template sizes, a window clamped at the frame border (R > margins), a wide template, two tracks
GEOMS = [(320, 240, 32, 32, 80, 1), (400, 300, 37, 29, 40, 1), (640, 360, 100, 24, 60, 1), (320, 240, 24, 48, 100, 1), (1280, 720, 48, 40, 80, 1),
         (640, 480, 32, 32, 40, 2), (1920, 1080, 64, 64, 80, 1)]


@pytest.mark.parametrize("W,H,tw,th,R,n_tracks", GEOMS)
@pytest.mark.parametrize("stats", ["winstats", "in_cta"])
def test_local_equals_default_ksplit_path(W, H, tw, th, R, n_tracks, stats):
    c = synth.make_clip(synth.ClipSpec(seed=31 + tw + n_tracks, W=W, H=H, tw=tw, th=th, n_frames=5, R=R))
    frames, roi = c["frames"], c["roi"]
    a, ma, ta, ka = run(frames, roi, R, n_tracks, PVT_NO_LOCAL=None, PVT_LOCAL_STATS="1" if stats == "in_cta" else None)
    b, mb, tb, kb = run(frames, roi, R, n_tracks, PVT_NO_LOCAL="1")
    if ka != (3 if stats == "in_cta" else 4) or kb != 4:
        pytest.skip("no k_ncc_local plan for this geometry (%s / %s kernels per step)" % (ka, kb))
    for f in ("x", "y", "moved", "updated", "searched", "valid"):
        assert np.array_equal(a[f], b[f]), f
    assert np.abs(a["conf"] - b["conf"]).max() <= 2e-6
    for t in range(n_tracks):
        assert np.array_equal(ta[t], tb[t])
        for k in range(len(ma)):
            (m1, w1), (m2, w2) = ma[k][t], mb[k][t]
            assert w1 == w2 and np.abs(m1 - m2).max() <= 2e-6
            deg = (m2 == 0) | (np.abs(m2) == 1)
            assert np.array_equal(m1[deg], m2[deg])


@pytest.mark.parametrize("name", ["small", "lowtex", "lost", "fade", "border", "flat", "oddsize", "c1_standin", "c2_1080p"])
def test_local_clip_vs_cv2_golden(name):
    (c, tk) = Hp.clip(name)
    g = Hp.golden(f"clip_{name}.npz")
    recs, templ = pvt.track_clip(c["frames"], c["roi"], search_radius_x=tk.get("rx", 80), search_radius_y=tk.get("ry", 80))
    Hp.check_records(records_of(recs), g["records"], name + " (default plan)")
    assert np.array_equal(templ, g["templ"])

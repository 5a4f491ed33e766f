"""k_winstats (one statistics kernel: sliding FP64 column sums + a CTA-wide prefix scan, no prefix arrays in HBM) against the
two-kernel statistics it replaces (k_colprefix + k_rowsum, still built behind PVT_STATS_LEGACY=1) and against the oracle.
The window sum of a u8-sourced frame is exact in double in any order, so the normalisers differ at most in the last bits of
the sum of squares: maps must agree to ~1e-7, degenerate cells and peaks exactly.
Reference semantics: cv::matchTemplate(TM_CCOEFF_NORMED) normalisation (tracker/src/ncc_cpu.cpp:12), SURVEY.md 8(c)."""
import importlib
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as Hp
from tools import synth

pvt = importlib.import_module("parallel-video-object-tracker_b200")
pytestmark = pytest.mark.gpu


def one_map(frames, roi, R, legacy, **kw):
    H, W = frames.shape[1:3]
    old = os.environ.pop("PVT_STATS_LEGACY", None)
    old_nl = os.environ.get("PVT_NO_LOCAL")
    os.environ["PVT_NO_LOCAL"] = "1"      # both runs on the K-split search with the same plan: only the statistics kernels differ
    if legacy:
        os.environ["PVT_STATS_LEGACY"] = "1"
    try:
        with pvt.Tracker(W, H, roi[2], roi[3], keep_maps=1, search_radius_x=R, search_radius_y=R, **kw) as tr:
            tr.init_track(0, frames[0], roi)
            res = tr.step([frames[1]])
            m, win = tr.window_map(0)
    finally:
        os.environ.pop("PVT_STATS_LEGACY", None)
        os.environ.pop("PVT_NO_LOCAL", None)
        if old_nl is not None:
            os.environ["PVT_NO_LOCAL"] = old_nl
        if old is not None:
            os.environ["PVT_STATS_LEGACY"] = old
    return m, win, res


# (W, H, tw, th, R): small / odd sizes / wide template (57 candidates per CTA along x: 3 x-tiles) / a template too wide for
# k_winstats (falls back to the two-kernel path: both runs identical) / 1080p C2 / a window clamped at the frame corner
GEOMS = [(320, 240, 32, 32, 80), (400, 300, 37, 29, 40), (640, 360, 200, 40, 60), (640, 360, 240, 16, 30), (1920, 1080, 64, 64, 80),
         (320, 240, 24, 48, 100)]


@pytest.mark.parametrize("W,H,tw,th,R", GEOMS)
@pytest.mark.parametrize("kernel", ["auto", "tc"])
def test_winstats_equals_two_kernel_statistics(W, H, tw, th, R, kernel):
    if kernel == "tc" and (tw > 260 or th > 129):
        pytest.skip("outside PVT_KERNEL_TC's geometry")      # (wide windows / templates run in column tiles: the 200-wide template here)
    c = synth.make_clip(synth.ClipSpec(seed=7 + tw, W=W, H=H, tw=tw, th=th, n_frames=3, R=R))
    frames, roi = c["frames"], c["roi"]
    kw = {"kernel": pvt.KERNEL_TC} if kernel == "tc" else {}
    new, win_n, rn = one_map(frames, roi, R, False, **kw)
    old, win_o, ro = one_map(frames, roi, R, True, **kw)
    assert win_n == win_o
    assert np.array_equal(np.isnan(new), np.isnan(old)) and not np.isnan(new).any()
    assert np.abs(new - old).max() <= 2e-7, float(np.abs(new - old).max())
    deg = (old == 0) | (np.abs(old) == 1)
    assert np.array_equal(new[deg], old[deg])                      # degenerate cells identical
    assert np.argmax(new) == np.argmax(old)
    assert (rn[0]["x"], rn[0]["y"]) == (ro[0]["x"], ro[0]["y"]) and abs(float(rn[0]["conf"]) - float(ro[0]["conf"])) <= 2e-7


def test_winstats_flat_and_low_variance_windows_match_the_oracle():
    """flat windows (normaliser 0 -> score 0) and near-flat ones sit on OpenCV's diff2 <= min(0.5, 10 eps wsq) threshold:
    the sliding sums must decide exactly like the oracle's integral images"""
    rng = np.random.default_rng(3)
    W, H, tw, th, R = 320, 240, 32, 32, 60
    f0 = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    f1 = f0.copy()
    f1[40:160, 60:220] = 77                                         # a flat region larger than the template
    f1[100:110, 100:180, :] = np.arange(80, dtype=np.uint8)[None, :, None] // 40 + 77   # ... with a faint step inside
    roi = (120, 90, tw, th)
    frames = np.stack([f0, f1])
    m, win, _ = one_map(frames, roi, R, False)
    g = O.to_gray_f32(f1)
    templ = O.to_gray_f32(f0)[roi[1]:roi[1] + th, roi[0]:roi[0] + tw]
    want = O.ncc_window(g, np.ascontiguousarray(templ), *win)
    assert np.array_equal(m == 0, want == 0)                        # the same cells are degenerate
    sig = Hp.window_sigma(g, tw, th, win)                           # 1e-4 where sigma_w >= 0.002, 5e-4 below
    d = np.abs(m - want)
    assert d[sig >= 0.002].max(initial=0) <= Hp.TOL_SCORE and d[sig < 0.002].max(initial=0) <= Hp.TOL_LOWVAR
    assert (want == 0).sum() > 100

"""CPU: pin the C oracle (oracle/ncc_oracle.c) against the cv2 4.13.0 golden vectors.

These are the reference's arithmetic calls; the GPU parity tests then compare
the CUDA path with this oracle at sizes the oracle finishes in seconds.
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as Hp


def test_ingest_lut_and_gray_bit_exact():
    g = Hp.golden("ingest.npz")
    lut = O.gray_to_f32(np.arange(256, dtype=np.uint8).reshape(1, 256))[0]
    assert np.array_equal(lut, g["lut"])                      # G6: f32 bit-exact
    assert np.array_equal(O.bgr2gray(g["bgr"]), g["gray"])    # G6: gray u8 bit-exact
    assert lut[0] == 0.0 and lut[255] == 1.0


def test_add_weighted_bit_exact_recurrence():
    g = Hp.golden("addweighted.npz")
    t = g["a"].copy()
    for i in range(1, 151):
        t = O.add_weighted(t, np.roll(g["b"], i, 0), 0.10)
        if i in (1, 10, 150):
            assert np.array_equal(t, g[f"after{i}"]), f"addWeighted differs after {i} steps"


def test_full_map_vs_cv2():
    g = Hp.golden("maps.npz")
    m = O.ncc_match_cpu(g["frame"], g["templ"])
    assert m.shape == g["full_ipp_off"].shape
    assert np.abs(m - g["full_ipp_off"]).max() <= 1.5e-7       # IPP-off == exact formula
    assert np.abs(m - g["full_ipp_on"]).max() <= Hp.TOL_SCORE
    assert np.argmax(m) == np.argmax(g["full_ipp_on"]) == np.argmax(g["full_ipp_off"])


def test_degenerate_cells_identical():
    g = Hp.golden("maps.npz")
    f, t = g["frame"], g["templ"]
    ones = O.ncc_match_cpu(f, np.full((13, 17), 0.25, np.float32))
    assert np.array_equal(ones, g["flat_templ_on"]) and np.all(ones == 1.0)       # flat template -> all ones
    self_ = O.ncc_match_cpu(f, f.copy())
    assert self_.shape == (1, 1) and abs(float(self_[0, 0]) - float(g["self_on"][0, 0])) <= 1e-6
    z = O.ncc_match_cpu(np.full_like(f, 0.5), t)
    assert np.array_equal(z, g["flat_frame_on"]) and np.all(z == 0.0)             # flat window -> exactly 0


def test_exact_ties_lowest_index():
    g = Hp.golden("maps.npz")
    m = O.ncc_match_cpu(g["tie_frame"], g["tie_templ"])
    b, x, y = O.max_loc(m)
    # periodic frame -> 16 exactly tied maxima; first one in row-major order wins (a11)
    assert np.array_equal(m, g["tie_map"])
    assert (m == m.max()).sum() == 16
    assert (x, y) == (int(g["tie_best"][1]), int(g["tie_best"][2])) == (5, 3)
    # the default IPP build breaks these ties with ~1e-5 noise: every tied cell stays within tolerance
    assert np.abs(m - g["tie_map_ipp_on"]).max() <= Hp.TOL_SCORE
    # cv::minMaxLoc itself on tied arrays, full and as an ROI view
    t = g["tied"]
    assert O.max_loc(t) == (float(g["tied_full"][0]), int(g["tied_full"][1]), int(g["tied_full"][2])) == (float(np.float32(0.9)), 4, 1)
    assert O.max_loc(t[1:, 2:]) == (float(g["tied_view"][0]), int(g["tied_view"][1]), int(g["tied_view"][2]))
    # NaN never wins, first occurrence wins
    a = np.array([[np.nan, 0.5, 0.5], [0.5, np.nan, 0.25]], np.float32)
    assert O.max_loc(a) == (0.5, 1, 0)


def test_search_window_matches_reference_arithmetic():
    from oracle import cv2_harness as H
    rng = np.random.default_rng(0)
    for _ in range(500):
        w, h = int(rng.integers(1, 130)), int(rng.integers(1, 130))
        W, H_ = int(rng.integers(w, 400)), int(rng.integers(h, 400))
        x, y = int(rng.integers(0, W - w + 1)), int(rng.integers(0, H_ - h + 1))
        rx, ry = int(rng.integers(0, 170)), int(rng.integers(0, 170))
        assert O.search_window(x, y, w, h, W - w + 1, H_ - h + 1, rx, ry) == \
            H.search_window(x, y, w, h, W - w + 1, H_ - h + 1, rx, ry)


@pytest.mark.parametrize("name", ["small", "lowtex", "lost", "fade", "border", "flat", "oddsize", "batch4",
                                  "c1_standin", "c2_1080p", "c3_4k"])
def test_clip_trajectory(name):
    (c, tk) = Hp.clip(name)
    g = Hp.golden(f"clip_{name}.npz")
    rec, templ = O.track_clip(c["frames"], c["roi"], rx=tk.get("rx", 80), ry=tk.get("ry", 80), batch=tk.get("batch", 1))
    Hp.check_records(rec, g["records"], name)
    # G5: with identical peaks the EMA recurrence is bit-identical
    assert np.array_equal(templ, g["templ"]), f"{name}: final template differs"


@pytest.mark.parametrize("name,k", [("small", 1), ("small", 7), ("lowtex", 2), ("border", 3), ("flat", 1),
                                    ("oddsize", 2), ("c2_1080p", 1)])
def test_window_map(name, k):
    (c, tk) = Hp.clip(name)
    g = Hp.golden(f"clip_{name}.npz")
    win = tuple(int(v) for v in g[f"map{k}_win"])
    gray = O.to_gray_f32(c["frames"][k])
    templ = g[f"map{k}_templ"]
    m = O.ncc_window(gray, templ, *win)
    off, on = g[f"map{k}_ipp_off"], g[f"map{k}_ipp_on"]
    assert m.shape == off.shape
    assert np.abs(m - off).max() <= 1.5e-7, "oracle vs IPP-off cv2"
    sig = Hp.window_sigma(gray, templ.shape[1], templ.shape[0], win)
    hi = sig >= 0.02
    assert np.abs(m - on)[hi].max() <= Hp.TOL_SCORE                 # G3 vs the default (IPP-on) cv2
    deg = (off == 0) | (np.abs(off) == 1)
    assert np.array_equal(m[deg], off[deg])                          # G4 degenerate cells identical
    assert np.argmax(m) == np.argmax(on)                             # G1


# ---- the eps formula of the reference's CUDA kernels (parity unpinned, see oracle/ncc_oracle.c) -----
def test_eps_formula_restatement_vs_exact_arithmetic():
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(5)
    f = gaussian_filter(rng.random((100, 140)), 2.0)
    f = ((f - f.min()) / (f.max() - f.min())).astype(np.float32)
    t = f[20:52, 60:92].copy() + rng.normal(0, 0.01, (32, 32)).astype(np.float32)
    seq = O.ncc_window_eps(f, t, 3, 2, 90, 60)
    exact = O.ncc_eps_exact(f, t, 3, 2, 90, 60)
    assert np.abs(seq - exact).max() <= 5e-5            # sequential FP32 sums vs float64
    assert np.unravel_index(np.argmax(seq), seq.shape) == (18, 57)
    # close to, but not the same as, the CPU operator ("NOT the oracle formula")
    cv = O.ncc_window(f, t, 3, 2, 90, 60)
    assert 0 < np.abs(seq - cv).max() <= 1e-3
    # flat template: cov == 0 -> 0 (OpenCV: 1); flat windows: the 1e-3 floor on sigma_w keeps the quotient finite
    assert np.abs(O.ncc_window_eps(f, np.full((8, 8), 0.5, np.float32), 0, 0, 20, 20)).max() <= 1e-3
    assert np.all(np.isfinite(O.ncc_window_eps(np.full_like(f, 0.5), t, 0, 0, 20, 20)))


def test_eps_formula_switch_drives_the_tracker_loop_and_resets():
    (c, tk) = Hp.clip("small")
    base, _ = O.track_clip(c["frames"], c["roi"])
    with O.formula(1):
        eps, _ = O.track_clip(c["frames"], c["roi"])
    again, _ = O.track_clip(c["frames"], c["roi"])
    assert np.array_equal(base, again)
    assert np.array_equal(base[:, :4], eps[:, :4]) and not np.array_equal(base[:, 4], eps[:, 4])

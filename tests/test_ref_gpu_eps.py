"""SURVEY.md 8(f) n4 pinned against the reference's OWN kernels: oracle/_ref/libref_baseline.so is
/root/reference/tracker/src/baseline_kernel.cu compiled unmodified for sm_100a (oracle/ref_build/Makefile) and run here on the
same inputs as libpvt's PVT_FORMULA_EPS operators.  The reference accumulates everything sequentially in FP32 (two passes per
window, baseline_kernel.cu:34-60); libpvt takes window sums from FP64 integrals and a blocked FP32 cross term, so the two agree
to FP32 summation noise, not bit for bit: the gate is the path's 1e-4 with identical peaks, and the tracker loop fed by either
map (main.cpp:103-161) follows the same trajectory."""
import importlib
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_gpu as RG
from tests import helpers as Hp

pvt = importlib.import_module("parallel-video-object-tracker_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _textured(rng, H, W):
    from scipy.ndimage import gaussian_filter
    f = gaussian_filter(rng.random((H, W)), 2.0)
    return ((f - f.min()) / (f.max() - f.min())).astype(np.float32)


def test_recipe_and_shim_are_committed_and_the_adapter_compiles_against_cv_mat(tmp_path):
    rb = os.path.join(ROOT, "oracle", "ref_build")
    for f in ("Makefile", "ref_capi.cu", os.path.join("opencv2", "opencv.hpp")):
        assert os.path.exists(os.path.join(rb, f)), f
    mk = open(os.path.join(rb, "Makefile")).read()
    assert "$(REF)/src/baseline_kernel.cu" in mk and "../_ref/" in mk          # sources where they lie, outputs into oracle/_ref only
    assert "oracle/_ref/" in open(os.path.join(ROOT, ".gitignore")).read()
    # -DPVT_WITH_OPENCV: the cv::Mat branch of the host adapter with the reference's exact signatures (no GPU needed to compile)
    pvt.lib()
    exe = str(tmp_path / "ops_demo_cv")
    pkg = os.path.join(ROOT, "parallel-video-object-tracker_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DPVT_WITH_OPENCV", "-I" + rb, "-I" + os.path.join(pkg, "host"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "ops_demo_cv.cpp"), "-L" + pkg, "-lpvt", "-Wl,-rpath," + pkg])
    assert subprocess.run([exe, "--signatures"], capture_output=True, text=True).stdout.strip() == "6"


@pytest.mark.gpu
def test_cv_mat_branch_of_the_adapter_runs(tmp_path):
    g = Hp.golden("maps.npz")
    f, t = g["frame"], g["templ"]
    rb = os.path.join(ROOT, "oracle", "ref_build")
    pkg = os.path.join(ROOT, "parallel-video-object-tracker_b200")
    exe = str(tmp_path / "ops_demo_cv")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DPVT_WITH_OPENCV", "-I" + rb, "-I" + os.path.join(pkg, "host"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "ops_demo_cv.cpp"), "-L" + pkg, "-lpvt", "-Wl,-rpath," + pkg])
    f.tofile(tmp_path / "f.f32"); t.tofile(tmp_path / "t.f32")
    r = subprocess.run([exe, str(tmp_path / "f.f32"), str(f.shape[1]), str(f.shape[0]), str(tmp_path / "t.f32"), str(t.shape[1]), str(t.shape[0]),
                        str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    ref = g["full_ipp_off"]
    for k in ("naive", "const_tiled", "batched1"):
        m = np.fromfile(tmp_path / f"o.{k}.f32", np.float32).reshape(ref.shape)
        assert np.abs(m - ref).max() <= Hp.TOL_SCORE, k


def _note(line):
    """measured differences for DESIGN.md / profiles (only when the scratch directory of a GPU run exists)"""
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "ref_gpu_diffs.txt"), "a") as fh:
            fh.write(line + "\n")


needs_ref = pytest.mark.skipif(not RG.available(), reason="oracle/_ref/libref_baseline.so not built (needs /root/reference at build time)")


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("mode", ["naive", "shared", "const", "const_tiled"])
def test_eps_formula_matches_the_reference_kernels(mode):
    rng = np.random.default_rng(21)
    # const_tiled returns its out-of-range threads BEFORE the cooperative tile load (baseline_kernel.cu:235 vs :251-263), so
    # edge CTAs of a map whose size is not a multiple of the 32 x 8 block read a partly unloaded tile: that defect is not
    # reproduced -- the tiled operator is compared on a map of 160 x 104 (whole blocks only)
    H, W, th, tw = (135, 199, 32, 40) if mode == "const_tiled" else (150, 210, 32, 40)
    f = _textured(rng, H, W)
    t = f[40:40 + th, 90:90 + tw].copy() + rng.normal(0, 0.01, (th, tw)).astype(np.float32)
    ref = RG.ncc_match(mode, f, t)
    got = getattr(pvt, {"naive": "ncc_match_naive_cuda", "shared": "ncc_match_shared_cuda", "const": "ncc_match_const",
                        "const_tiled": "ncc_match_const_tiled"}[mode])(f, t, formula=pvt.FORMULA_EPS)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= Hp.TOL_SCORE, float(np.abs(got - ref).max())
    assert np.argmax(got) == np.argmax(ref)
    # the repo's own restatement of the kernel (FP32 sequential, oracle/ncc_oracle.c) against the real thing: this is the pin
    seq = O.ncc_window_eps(f, t, 0, 0, ref.shape[1], ref.shape[0])
    _note(f"{mode}: |libpvt eps - reference kernel| max {np.abs(got - ref).max():.3e}; |oracle FP32-sequential restatement - reference kernel| max "
          f"{np.abs(seq - ref).max():.3e}; |float64 formula - reference kernel| max {np.abs(O.ncc_eps_exact(f, t, 0, 0, ref.shape[1], ref.shape[0]) - ref).max():.3e}")
    assert np.array_equal(seq, ref), float(np.abs(seq - ref).max())   # measured on B200: bit-identical
    # and the default formula is NOT what these kernels compute (flat template: OpenCV says 1 everywhere, the kernels ~0)
    flat = np.full((th, tw), 0.25, np.float32)
    rf = RG.ncc_match(mode, f, flat)
    gf = pvt.ncc_match_naive_cuda(f, flat, formula=pvt.FORMULA_EPS)
    assert np.abs(rf).max() <= 1e-3 and np.abs(gf - rf).max() <= Hp.TOL_SCORE


@pytest.mark.gpu
@needs_ref
def test_eps_batched_and_1080p_against_the_reference():
    rng = np.random.default_rng(22)
    f0, f1 = _textured(rng, 120, 160), _textured(rng, 120, 160)
    t = f0[30:54, 50:82].copy()
    ref = RG.ncc_match_batched([f0, f1], t)
    got = pvt.ncc_match_naive_cuda_batched([f0, f1], t, formula=pvt.FORMULA_EPS)
    for a, b in zip(got, ref):
        assert np.abs(a - b).max() <= Hp.TOL_SCORE and np.argmax(a) == np.argmax(b)
    # BASELINE.json configs[1] geometry: 1920x1080 frame, 64x64 template (N = 4096, the const modes' limit, baseline_kernel.cu:500)
    (c, _) = Hp.clip("c2_1080p")
    g0, g1 = O.to_gray_f32(c["frames"][0]), O.to_gray_f32(c["frames"][1])
    x, y, w, h = c["roi"]
    t = g0[y:y + h, x:x + w].copy()
    ref = RG.ncc_match("const", g1, t)
    got = pvt.ncc_match_const(g1, t, formula=pvt.FORMULA_EPS)
    d = np.abs(got - ref)
    sig = Hp.window_sigma(g1, w, h, (0, 0, ref.shape[1], ref.shape[0]))
    _note(f"1080p const: |libpvt eps - reference kernel| max {d.max():.3e} (sigma_w >= 0.02: {d[sig >= 0.02].max():.3e})")
    # At N = 4096 pixels the reference kernel's own sequential FP32 sums (two passes, baseline_kernel.cu:34-60) sit up to ~2e-4
    # from exact arithmetic (measured 2.15e-4 here; 2.2e-5 at N = 1280 above); libpvt's eps path stays within 2e-5 of the exact
    # formula.  So at this size the gate between the two is the reference's noise, 5e-4, with identical peaks -- and the pin is the
    # crop below, where the FP32-sequential restatement reproduces the reference kernel BIT FOR BIT.
    assert d.max() <= 5e-4 and np.argmax(got) == np.argmax(ref), float(d.max())
    crop = np.ascontiguousarray(g1[400:640, 800:1120])
    rc = RG.ncc_match("naive", crop, t)
    seq = O.ncc_window_eps(crop, t, 0, 0, rc.shape[1], rc.shape[0])
    assert np.array_equal(seq, rc), float(np.abs(seq - rc).max())
    gc = pvt.ncc_match_naive_cuda(crop, t, formula=pvt.FORMULA_EPS)
    exact = O.ncc_eps_exact(crop, t, 0, 0, rc.shape[1], rc.shape[0])
    _note(f"1080p crop 64x64: restatement == reference kernel bit-exact; |libpvt - float64 formula| max {np.abs(gc - exact).max():.3e}; "
          f"|reference kernel - float64 formula| max {np.abs(rc - exact).max():.3e}")
    assert np.abs(gc - exact).max() <= 2e-5 and np.argmax(gc) == np.argmax(rc)
    with pytest.raises(RuntimeError):                      # 72 x 72 > 4096 px: the reference asserts, libpvt does not
        RG.ncc_match("const", g1, g0[100:172, 100:172].copy())
    assert pvt.ncc_match_const(g1[:300, :300].copy(), g0[100:172, 100:172].copy(), formula=pvt.FORMULA_EPS).shape == (229, 229)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("name", ["small", "oddsize"])
def test_tracker_on_reference_gpu_maps_follows_the_same_trajectory(name):
    """main.cpp:103-161 fed by the REAL reference kernel's map (what `tracker --naive` does) vs Tracker(formula=EPS)."""
    (c, tk) = Hp.clip(name)
    frames, roi = c["frames"], c["roi"]
    rx, ry = tk.get("rx", 80), tk.get("ry", 80)
    x, y, w, h = roi
    templ = O.to_gray_f32(frames[0])[y:y + h, x:x + w].copy()
    want = []
    for k in range(1, len(frames)):
        g = O.to_gray_f32(frames[k])
        m = RG.ncc_match("naive", g, templ)
        win = O.search_window(x, y, w, h, m.shape[1], m.shape[0], rx, ry)
        best, bx, by = O.max_loc(np.ascontiguousarray(m[win[1]:win[1] + win[3], win[0]:win[0] + win[2]]))
        bx, by = bx + win[0], by + win[1]
        moved = updated = 0
        if best >= 0.40:
            x, y, moved = bx, by, 1
            if best >= 0.70:
                templ = O.add_weighted(templ, g[y:y + h, x:x + w])
                updated = 1
        want.append((x, y, w, h, best, moved, updated))
    want = np.array(want, np.float64)
    recs, t_end = pvt.track_clip(frames, roi, search_radius_x=rx, search_radius_y=ry, formula=pvt.FORMULA_EPS)
    got = np.stack([recs["x"], recs["y"], recs["w"], recs["h"], recs["conf"].astype(np.float64), recs["moved"], recs["updated"]], 1).astype(np.float64)
    Hp.check_records(got, want, name + " (reference GPU maps)")
    assert np.array_equal(t_end, templ)

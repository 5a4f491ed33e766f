"""The C++ host layer (reference interface mirror + CLI twin) over the C ABI.
CPU part: it builds and rejects what it must.  GPU part: same trajectories / maps as the golden vectors."""
import importlib
import os
import struct
import subprocess

import numpy as np
import pytest

from tests import helpers as Hp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "parallel-video-object-tracker_b200", "host")


@pytest.fixture(scope="module")
def built():
    importlib.import_module("parallel-video-object-tracker_b200").lib()
    subprocess.check_call(["make", "-s", "-C", HOST])
    return HOST


def write_clip(path, frames):
    n, h, w, _ = frames.shape
    with open(path, "wb") as f:
        f.write(b"PVTBGR1\n" + struct.pack("<iii", w, h, n))
        f.write(np.ascontiguousarray(frames).tobytes())


def test_cli_builds_and_rejects_cpu_mode(built, tmp_path):
    r = subprocess.run([os.path.join(built, "tracker"), "--cpu", "nothing.bgr"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in r.stderr
    assert "Mode        : cpu" in r.stdout                      # the reference banner, main.cpp:43-49
    r = subprocess.run([os.path.join(built, "tracker"), str(tmp_path / "missing.bgr"), "--roi", "1,1,4,4"], capture_output=True, text=True)
    assert r.returncode != 0 and "Cannot open video." in r.stderr   # main.cpp:54
    (c, _) = Hp.clip("small")
    write_clip(tmp_path / "c.bgr", c["frames"][:2])
    r = subprocess.run([os.path.join(built, "tracker"), str(tmp_path / "c.bgr")], capture_output=True, text=True)
    assert r.returncode != 0 and "No ROI selected." in r.stderr      # main.cpp:66-69


@pytest.mark.gpu
@pytest.mark.parametrize("name,flags", [("small", []), ("small", ["--const_tiled"]), ("lost", ["--shared"]), ("batch4", ["--batch=4"])])
def test_cli_trajectory_matches_golden(built, tmp_path, name, flags):
    (c, tk) = Hp.clip(name)
    g = Hp.golden(f"clip_{name}.npz")["records"]
    write_clip(tmp_path / "c.bgr", c["frames"])
    roi = ",".join(str(v) for v in c["roi"])
    r = subprocess.run([os.path.join(built, "tracker"), *flags, str(tmp_path / "c.bgr"), "--roi", roi, "--out", str(tmp_path / "o.csv")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert " Tracking Complete" in r.stdout and f" Frames     : {len(c['frames']) - 1}" in r.stdout   # main.cpp:175-182
    rows = np.genfromtxt(tmp_path / "o.csv", delimiter=",", skip_header=1)
    got = np.column_stack([rows[:, 1:5], rows[:, 5], rows[:, 6:8]])
    Hp.check_records(got, g, f"cli {name} {flags}")


@pytest.mark.gpu
def test_cli_gpu_formula_flag_follows_the_eps_oracle(built, tmp_path):
    from oracle import oracle as O
    (c, tk) = Hp.clip("small")
    with O.formula(1):
        want, _ = O.track_clip(c["frames"], c["roi"])
    write_clip(tmp_path / "c.bgr", c["frames"])
    roi = ",".join(str(v) for v in c["roi"])
    r = subprocess.run([os.path.join(built, "tracker"), "--gpu-formula", str(tmp_path / "c.bgr"), "--roi", roi, "--out", str(tmp_path / "o.csv")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = np.genfromtxt(tmp_path / "o.csv", delimiter=",", skip_header=1)
    got = np.column_stack([rows[:, 1:5], rows[:, 5], rows[:, 6:8]])
    Hp.check_records(got, want[:, :7], "cli --gpu-formula")
    base = Hp.golden("clip_small.npz")["records"]
    assert not np.array_equal(got[:, 4].astype(np.float32), base[:, 4].astype(np.float32))   # it is a different score


def test_ghc_cli_builds_and_rejects(built, tmp_path):
    exe = os.path.join(built, "tracker_ghc")
    r = subprocess.run([exe, "nothing.bgr", "--cpu"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in r.stderr
    r = subprocess.run([exe, str(tmp_path / "missing.bgr")], capture_output=True, text=True)
    assert r.returncode != 0 and "Cannot open video:" in r.stderr         # tracker_ghc/src/main.cpp:84
    (c, _) = Hp.clip("small")
    write_clip(tmp_path / "c.bgr", c["frames"][:2])
    r = subprocess.run([exe, str(tmp_path / "c.bgr"), "--first"], capture_output=True, text=True)
    assert r.returncode != 0 and "No template selected" in r.stderr        # :117-120


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--tc-global"]])
def test_ghc_cli_reacquires_like_the_golden(built, tmp_path, extra):
    """tracker_ghc twin: local search -> lost -> whole-frame search -> re-acquired, as cv2 4.13.0 does it.
    --tc-global: the whole-frame search on the tensor cores (PVT_KERNEL_TC_GLOBAL), same golden."""
    import json
    import zlib
    from tools import synth
    with open(os.path.join(Hp.GOLD, "meta_ghc.json")) as fh:
        m = json.load(fh)["clips"]["reacquire"]
    c = synth.make_clip(synth.ClipSpec(**m["spec"]))
    assert zlib.crc32(np.ascontiguousarray(c["frames"]).tobytes()) & 0xFFFFFFFF == m["frames_crc"]
    want = np.load(os.path.join(Hp.GOLD, "ghc_reacquire.npz"))["records"]
    write_clip(tmp_path / "c.bgr", c["frames"])
    roi = ",".join(str(v) for v in c["roi"])
    tk = m["track"]
    r = subprocess.run([os.path.join(built, "tracker_ghc"), str(tmp_path / "c.bgr"), "--first", "--roi", roi, "--out", str(tmp_path / "o.csv"),
                        "--radius", str(tk["rx"]), "--lost", str(tk["lost_threshold"])] + extra, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Tracking mode: cuda" in r.stdout and f"Interactive tracking summary: frames={len(c['frames']) - 1}," in r.stdout
    rows = np.genfromtxt(tmp_path / "o.csv", delimiter=",", skip_header=1)
    assert np.array_equal(rows[:, 1:5], want[:, :4]) and np.array_equal(rows[:, 6:10], want[:, 5:9])
    # use_global_search: the reference raises the flag at the START of the next frame (tracker_ghc/src/main.cpp:183-185), the
    # library reports it as it will be evaluated there -- so it leads the golden's end-of-frame value by one frame
    assert np.array_equal(rows[:-1, 10], np.maximum(want[:-1, 9], (want[1:, 7] == 2)))
    assert np.abs(rows[:, 5] - want[:, 4]).max() <= Hp.TOL_SCORE
    assert (rows[:, 8] == 2).sum() == m["global_frames"]


@pytest.mark.gpu
def test_cpp_operators_match_golden(built, tmp_path):
    g = Hp.golden("maps.npz")
    f, t = g["frame"], g["templ"]
    f.tofile(tmp_path / "f.f32"); t.tofile(tmp_path / "t.f32")
    r = subprocess.run([os.path.join(built, "ops_demo"), str(tmp_path / "f.f32"), str(f.shape[1]), str(f.shape[0]),
                        str(tmp_path / "t.f32"), str(t.shape[1]), str(t.shape[0]), str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    ref = g["full_ipp_off"]
    maps = {k: np.fromfile(tmp_path / f"o.{k}.f32", np.float32).reshape(ref.shape)
            for k in ("naive", "shared", "const", "const_tiled", "view", "batched0", "batched1")}
    for k, m in maps.items():
        assert np.abs(m - ref).max() <= Hp.TOL_SCORE, k
        assert np.array_equal(m, maps["naive"]), k          # every mode runs the same kernels

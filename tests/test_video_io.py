"""the frame source and sink either side of the path -- tracker/src/main.cpp:52-60 (cv::VideoCapture),
:73-82 (cv::VideoWriter, mp4v, the capture's fps), :95-96 (cap >> frame), :166-167 (cv::rectangle + writer.write).
`track_video` decodes / encodes with OpenCV's videoio on the host exactly as the reference does; the GPU test compares the
whole pipeline (decode -> pvt_step -> pvt_draw_boxes -> encode) with the same loop done by cv2 4.13.0 alone on the same file:
identical trajectory and flags, scores within 1e-4, final template bit-identical, and an output clip that decodes to the same
frames.  CPU tests: the reference's error behaviour, and the container round trip the fixture relies on."""
import importlib

import numpy as np
import pytest

from oracle import cv2_harness as H
from tests import helpers as Hp

pvt = importlib.import_module("parallel-video-object-tracker_b200")
video = importlib.import_module("parallel-video-object-tracker_b200.video")
cv2 = pytest.importorskip("cv2")


def write_video(path, frames, fps=25.0, fourcc="mp4v"):
    n, Hh, W, _ = frames.shape
    w = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*fourcc), fps, (W, Hh))
    assert w.isOpened(), "this OpenCV build cannot encode " + fourcc
    for f in frames:
        w.write(np.ascontiguousarray(f))
    w.release()


def read_video(path):
    cap = cv2.VideoCapture(str(path))
    assert cap.isOpened()
    out = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        out.append(f)
    fps = cap.get(cv2.CAP_PROP_FPS)
    cap.release()
    return np.stack(out), fps


def test_track_video_errors_follow_the_reference(tmp_path):
    with pytest.raises(IOError, match="Cannot open video"):                  # main.cpp:53-56
        video.track_video(tmp_path / "missing.mp4", (10, 10, 32, 32))
    (c, _) = Hp.clip("small")
    write_video(tmp_path / "in.mp4", c["frames"][:4])
    with pytest.raises(ValueError, match="No ROI selected"):                 # main.cpp:66-69 (checked before any device work)
        video.track_video(tmp_path / "in.mp4", (10, 10, 0, 32))


def test_video_cli_flags_and_errors(tmp_path, capsys):
    """main.cpp:23-49: banner, --batch=N parsing, --cpu rejected; :53-56 / :66-69 error order (video first, then the ROI)."""
    assert video.main(["--cpu", "x.mp4"]) == -1
    out = capsys.readouterr()
    assert "NCC Tracker Starting" in out.out and "Mode        : cpu" in out.out and "no CPU path" in out.err
    assert video.main([str(tmp_path / "missing.mp4"), "--batch=3"]) == -1
    out = capsys.readouterr()
    assert "Mode        : batch" in out.out and "Batch size  : 3" in out.out and "Cannot open video." in out.err
    (c, _) = Hp.clip("small")
    write_video(tmp_path / "in.mp4", c["frames"][:3])
    assert video.main([str(tmp_path / "in.mp4")]) == -1                      # opened, but no --roi
    assert "No ROI selected." in capsys.readouterr().err


def test_container_round_trip_is_deterministic(tmp_path):
    """The fixture of the GPU test: the same frames through the same encoder give the same decoded frames (so the two
    pipelines' outputs can be compared frame by frame), the frame count and the fps survive the container."""
    (c, _) = Hp.clip("small")
    fr = c["frames"][:12]
    write_video(tmp_path / "a.mp4", fr, fps=25.0)
    write_video(tmp_path / "b.mp4", fr, fps=25.0)
    a, fa = read_video(tmp_path / "a.mp4")
    b, fb = read_video(tmp_path / "b.mp4")
    assert a.shape == fr.shape and np.array_equal(a, b) and abs(fa - 25.0) < 1e-6 and fa == fb


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small", "border"])
def test_track_video_equals_the_cv2_loop_on_the_same_file(name, tmp_path):
    (c, tk) = Hp.clip(name)
    rx, ry = tk.get("rx", 80), tk.get("ry", 80)
    roi = tuple(int(v) for v in c["roi"])
    write_video(tmp_path / "in.mp4", c["frames"], fps=25.0)
    dec, fps_in = read_video(tmp_path / "in.mp4")                            # what BOTH pipelines see: the lossy decode, BGR8
    assert dec.shape == c["frames"].shape
    # the reference loop with the real library: cv2_harness.track_clip restates main.cpp:93-161 call for call
    want = H.track_clip(dec, roi, rx=rx, ry=ry)
    ann = dec.copy()
    for k, r in enumerate(want["records"]):
        cv2.rectangle(ann[k + 1], tuple(int(v) for v in r[:4]), (0, 255, 0), 2)      # main.cpp:166
    write_video(tmp_path / "ref_out.mp4", ann[1:], fps=fps_in)
    # the library between the same decoder and encoder
    recs, templ, summary = video.track_video(tmp_path / "in.mp4", roi, tmp_path / "out.mp4", search_radius_x=rx, search_radius_y=ry)
    got = np.stack([recs["x"], recs["y"], recs["w"], recs["h"], recs["conf"].astype(np.float64), recs["moved"], recs["updated"]], 1).astype(np.float64)
    Hp.check_records(got, want["records"], name + " (video)")
    assert np.array_equal(templ, want["templ"]), "final template not bit-identical"
    assert summary["frames"] == len(dec) - 1 and summary["frame_size"] == (dec.shape[2], dec.shape[1]) and abs(summary["fps_video"] - fps_in) < 1e-9
    ours, fps_o = read_video(tmp_path / "out.mp4")
    ref, fps_r = read_video(tmp_path / "ref_out.mp4")
    assert ours.shape == ref.shape == ann[1:].shape and fps_o == fps_r
    assert np.array_equal(ours, ref), "annotated output clip differs from the cv2 pipeline's"


@pytest.mark.gpu
def test_video_cli_batch_mode_equals_the_cv2_loop(tmp_path, capsys):
    """The command line on a file with --batch=2 (main.cpp:115-130: every second frame searched, the stale box drawn on the
    others): return code, banner / summary, and an output clip that decodes to the cv2 pipeline's frames."""
    (c, tk) = Hp.clip("small")
    roi = tuple(int(v) for v in c["roi"])
    write_video(tmp_path / "in.mp4", c["frames"], fps=30.0)
    dec, fps_in = read_video(tmp_path / "in.mp4")
    want = H.track_clip(dec, roi, batch=2)
    ann = dec.copy()
    for k, r in enumerate(want["records"]):
        cv2.rectangle(ann[k + 1], tuple(int(v) for v in r[:4]), (0, 255, 0), 2)
    write_video(tmp_path / "ref_out.mp4", ann[1:], fps=fps_in)
    rc = video.main([str(tmp_path / "in.mp4"), "--roi", ",".join(str(v) for v in roi), "--out", str(tmp_path / "out.mp4"), "--batch=2"])
    out = capsys.readouterr().out
    assert rc == 0 and "Mode        : batch" in out and "Batch size  : 2" in out and "Tracking Complete" in out
    assert " Frames     : %d" % (len(dec) - 1) in out
    ours, _ = read_video(tmp_path / "out.mp4")
    ref, _ = read_video(tmp_path / "ref_out.mp4")
    assert np.array_equal(ours, ref), "annotated output clip differs from the cv2 pipeline's"

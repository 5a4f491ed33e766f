"""The sink side of the reference loop: tracker/src/main.cpp:166  cv::rectangle(frame, bbox, {0,255,0}, 2).
CPU: the oracle's restatement of cv::rectangle's thickness-2 coverage against the real cv2 4.13.0 over random in-frame boxes
(borders, 1-pixel boxes).  GPU: pvt_draw_boxes (k_overlay) against cv2, host and device frames, several boxes, and the CLI twin's
--video-out clip."""
import importlib
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as Hp

pvt = importlib.import_module("parallel-video-object-tracker_b200")
cv2 = pytest.importorskip("cv2")


def boxes_for(rng, W, H, n):
    out = []
    for t in range(n):
        w, h = int(rng.integers(1, min(48, W) + 1)), int(rng.integers(1, min(48, H) + 1))
        x, y = int(rng.integers(0, W - w + 1)), int(rng.integers(0, H - h + 1))
        if t % 3 == 0: x = 0
        if t % 5 == 0: y = H - h
        if t % 7 == 0: x = W - w
        if t % 11 == 0: y = 0
        out.append((x, y, w, h))
    return out


def test_oracle_rectangle_equals_cv2():
    rng = np.random.default_rng(1)
    for _ in range(400):
        W, H = int(rng.integers(20, 120)), int(rng.integers(20, 90))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        for box in boxes_for(rng, W, H, 3):
            a, b = img.copy(), img.copy()
            cv2.rectangle(a, box, (0, 255, 0), 2)
            O.draw_rectangle(b, box)
            assert np.array_equal(a, b), box


@pytest.mark.gpu
@pytest.mark.parametrize("W,H", [(320, 240), (1920, 1080), (101, 67)])
def test_draw_boxes_equals_cv2(W, H):
    rng = np.random.default_rng(W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    boxes = boxes_for(rng, W, H, 9)
    want = img.copy()
    for b in boxes:
        cv2.rectangle(want, b, (0, 255, 0), 2)
    with pvt.Tracker(W, H, 8, 8) as tr:
        got = tr.draw_boxes(img.copy(), boxes)
        assert np.array_equal(got, want)
        red = tr.draw_boxes(img.copy(), boxes[:1], bgr=(0, 0, 255))
        w2 = img.copy(); cv2.rectangle(w2, boxes[0], (0, 0, 255), 2)
        assert np.array_equal(red, w2)
        with pytest.raises(pvt.PvtError):
            tr.draw_boxes(img.copy(), [(W - 4, 0, 8, 8)])          # leaves the frame: rejected, not painted differently
        torch = pytest.importorskip("torch")
        d = torch.from_numpy(img).cuda()
        tr.draw_boxes(pvt.device_frame(d.data_ptr(), W * 3), boxes)
        assert np.array_equal(d.cpu().numpy(), want)


@pytest.mark.gpu
def test_cli_video_out_is_the_reference_annotated_clip(tmp_path):
    """tracker --video-out: every frame annotated exactly like main.cpp:166-167 does it (cv::rectangle at the frame's new box)."""
    from tests.test_host_cpp import HOST, write_clip
    pvt.lib()
    subprocess.check_call(["make", "-s", "-C", HOST])
    (c, tk) = Hp.clip("small")
    g = Hp.golden("clip_small.npz")["records"]
    frames = c["frames"]
    write_clip(tmp_path / "c.bgr", frames)
    roi = ",".join(str(v) for v in c["roi"])
    r = subprocess.run([os.path.join(HOST, "tracker"), str(tmp_path / "c.bgr"), "--roi", roi, "--video-out", str(tmp_path / "o.bgr")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(tmp_path / "o.bgr", "rb").read()
    assert raw[:8] == b"PVTBGR1\n"
    W, H, N = np.frombuffer(raw[8:20], np.int32)
    assert (W, H, N) == (frames.shape[2], frames.shape[1], len(frames) - 1)
    out = np.frombuffer(raw[20:], np.uint8).reshape(N, H, W, 3)
    for k in range(N):
        want = frames[k + 1].copy()
        x, y, w, h = (int(v) for v in g[k, :4])
        cv2.rectangle(want, (x, y, w, h), (0, 255, 0), 2)
        assert np.array_equal(out[k], want), k

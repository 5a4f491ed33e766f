"""Deterministic inputs of the filled-GPU parity cases (BASELINE.json configs[3] / configs[4] shapes and the 1080p
whole-frame pass), shared by tests/golden/make_golden_filled.py (cv2 4.13.0, build container) and
tests/test_filled_gpu_plans.py (GPU box).  Integer-only generation through tools/synth.py: byte-identical everywhere,
guarded by CRCs stored in tests/golden/meta_filled.json.

    c5: 64 tracks on 64 DISTINCT 1920x1080 streams, 64x64 templates, R = 80, 3 tracked frames.  8 seeded scenes x 8
        variants (vertical / horizontal flip, channel swap) give 64 different image sequences; three of four streams
        track the moving object, the fourth a background patch near a border or corner (clamped windows, every x-alignment).
    c4: 256 ROIs on ONE 1920x1080 stream (the moving object, a 16 x 15 grid of background patches, 15 border / corner
        boxes), 64x64, R = 80, 2 tracked frames.
    wf: one 1920x1080 stream, 64x64, lost-object parameters (tracker_ghc): the track is forced into the lost state, so
        frame 1 is searched over the WHOLE map (1857 x 1017 candidates), frame 2 locally again.
"""
from __future__ import annotations

import zlib

import numpy as np

from tools import synth

W, H, TW, TH, R = 1920, 1080, 64, 64, 80


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def _variant(frame: np.ndarray, v: int) -> np.ndarray:
    if v & 1:
        frame = frame[::-1]
    if v & 2:
        frame = frame[:, ::-1]
    if v & 4:
        frame = frame[:, :, ::-1]
    return np.ascontiguousarray(frame)


def _variant_pos(x: int, y: int, v: int):
    return (W - TW - x if v & 2 else x), (H - TH - y if v & 1 else y)


# background boxes for every fourth stream of c5 / the border boxes of c4: corners, edges, all four x alignments
BORDER_BOXES = [(0, 0), (1, 2), (2, 0), (3, 7), (W - TW, H - TH), (W - TW - 1, H - TH - 3), (W - TW, 0), (0, H - TH),
                (70, 5), (5, 70), (W - TW - 37, 300), (901, H - TH - 2), (79, 79), (81, 82), (W - TW - 80, H - TH - 81), (640, 0)]


class C5:
    n_streams, n_frames = 64, 4

    def __init__(self):
        self.scenes = [synth.Scene(synth.ClipSpec(seed=21 + b, W=W, H=H, tw=TW, th=TH, n_frames=self.n_frames, R=R, period=32))
                       for b in range(8)]
        self._base = {}

    def frame(self, s: int, k: int) -> np.ndarray:
        b, v = s % 8, s // 8
        if (b, k) not in self._base:
            self._base[(b, k)] = self.scenes[b].frame(k)
        return _variant(self._base[(b, k)], v)

    def frames_at(self, k: int):
        return [self.frame(s, k) for s in range(self.n_streams)]

    def roi(self, s: int):
        b, v = s % 8, s // 8
        if s % 4 == 3:
            x, y = BORDER_BOXES[s // 4]
        else:
            x, y = _variant_pos(*self.scenes[b].obj_pos(0), v)
        return (int(x), int(y), TW, TH)

    def crc(self) -> int:
        c = 0
        for k in range(self.n_frames):
            for s in range(self.n_streams):
                c = zlib.crc32(self.frame(s, k).tobytes(), c)
        return c & 0xFFFFFFFF


class C4:
    n_rois, n_frames = 256, 3

    def __init__(self):
        self.scene = synth.Scene(synth.ClipSpec(seed=31, W=W, H=H, tw=TW, th=TH, n_frames=self.n_frames, R=R, period=32))
        self.frames = [self.scene.frame(k) for k in range(self.n_frames)]

    def rois(self):
        ox, oy = self.scene.obj_pos(0)
        out = [(ox, oy, TW, TH)]
        for gy in range(15):
            for gx in range(16):
                out.append((48 + gx * 116 + (gy % 4), 41 + gy * 66 + (gx % 3), TW, TH))
        out += [(x, y, TW, TH) for (x, y) in BORDER_BOXES[:15]]
        assert len(out) == self.n_rois
        return [(int(a), int(b), c, d) for (a, b, c, d) in out]

    def crc(self) -> int:
        return crc(np.stack(self.frames))


class WF:
    n_frames = 3

    def __init__(self):
        self.scene = synth.Scene(synth.ClipSpec(seed=41, W=W, H=H, tw=TW, th=TH, n_frames=self.n_frames, R=R, period=32))
        self.frames = [self.scene.frame(k) for k in range(self.n_frames)]

    def roi(self):
        return (*self.scene.obj_pos(0), TW, TH)

    # where the (lost) track believes the object is before frame 1: far from it, so only the whole-map search finds it
    stale_box = (100, 900, TW, TH)

    def crc(self) -> int:
        return crc(np.stack(self.frames))

#!/usr/bin/env python
"""Golden fixtures for the lost-object re-acquisition path (the reference's second tracker,
/root/reference/tracker_ghc/src/main.cpp:145-239), generated with the REAL OpenCV (cv2 4.13.0) through
oracle/cv2_harness.track_clip_ghc, which restates that loop call for call.

Run in the build container only (needs cv2):  python tests/golden/make_golden_ghc.py
Inputs are NOT stored: the clips are regenerated from their seeds by tools/synth.py (variant "lost": the object
disappears for a fifth of the clip while it keeps moving, so it re-appears outside the local search window).
"""
from __future__ import annotations

import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import cv2_harness as H  # noqa: E402
from tools import synth  # noqa: E402

# name -> (ClipSpec kwargs, tracker_ghc parameters)
CLIPS = {
    # short threshold: lost -> 4 lost frames -> whole-frame search fails while hidden -> re-acquired -> local again
    "reacquire": (dict(seed=31, W=320, H=240, tw=32, th=32, n_frames=48, R=40, variant="lost"),
                  dict(rx=12, ry=12, lost_threshold=4)),
    # the reference's own constants (radius 60, LOST_FRAME_THRESHOLD 50, NCC_GLOBAL_CONFIDENCE 0.6)
    "ghc_defaults": (dict(seed=32, W=320, H=240, tw=32, th=32, n_frames=300, R=40, variant="lost", period=64),
                     dict(rx=60, ry=60, lost_threshold=50)),
    # odd geometry, template not a multiple of 8, window clamped at the borders
    "reacquire_odd": (dict(seed=33, W=301, H=233, tw=37, th=29, n_frames=40, R=40, variant="lost"),
                      dict(rx=10, ry=14, lost_threshold=3, global_conf=0.55)),
}


def main():
    meta = {"cv2": H.cv2.__version__, "clips": {}}
    for name, (skw, tkw) in CLIPS.items():
        clip = synth.make_clip(synth.ClipSpec(**skw))
        frames, roi = clip["frames"], clip["roi"]
        r = H.track_clip_ghc(frames, roi, **tkw)
        rec = r["records"]
        np.savez_compressed(os.path.join(HERE, f"ghc_{name}.npz"), records=rec, templ=r["templ"], roi=np.array(roi, np.int32),
                            truth=clip["truth"])
        meta["clips"][name] = {
            "spec": skw, "track": tkw, "frames_crc": zlib.crc32(np.ascontiguousarray(frames).tobytes()) & 0xFFFFFFFF,
            "n": int(len(frames)), "moved": int(rec[:, 5].sum()), "updated": int(rec[:, 6].sum()),
            "global_frames": int((rec[:, 7] == 2).sum()), "reacquired": int(((rec[:, 7] == 2) & (rec[:, 5] == 1)).sum()),
            "lost_max": int(rec[:, 8].max()),
        }
        print(name, meta["clips"][name])
    with open(os.path.join(HERE, "meta_ghc.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

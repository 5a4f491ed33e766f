#!/usr/bin/env python
"""Golden fixtures for the filled-GPU parity cases (tests/filled_cases.py) with the REAL OpenCV (cv2 4.13.0) through
oracle/cv2_harness.py: per-track records of every tracked frame and the CRC of every final template (bit-exact gate),
plus a few final templates in full.  Build container only:  python tests/golden/make_golden_filled.py
Inputs are regenerated from seeds on the GPU box; meta_filled.json holds their CRCs."""
from __future__ import annotations

import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from oracle import cv2_harness as H  # noqa: E402
from tests import filled_cases as FC  # noqa: E402


def tcrc(t):
    return zlib.crc32(np.ascontiguousarray(t, np.float32).tobytes()) & 0xFFFFFFFF


def main():
    assert cv2.__version__.startswith("4.13"), cv2.__version__
    meta = {"cv2": cv2.__version__, "ipp": bool(cv2.ipp.useIPP())}

    c5 = FC.C5()
    recs, crcs, templs = [], [], {}
    for s in range(c5.n_streams):
        fr = np.stack([c5.frame(s, k) for k in range(c5.n_frames)])
        r = H.track_clip(fr, c5.roi(s), rx=FC.R, ry=FC.R)
        recs.append(r["records"]); crcs.append(tcrc(r["templ"]))
        if s % 16 == 3 or s == 63:
            templs[f"templ_{s}"] = r["templ"]
    np.savez_compressed(os.path.join(HERE, "filled_c5_64x1080p.npz"), records=np.array(recs), templ_crc=np.array(crcs, np.uint32), **templs)
    meta["c5"] = {"frames_crc": c5.crc(), "rois": [c5.roi(s) for s in range(c5.n_streams)]}
    print("c5: moved %d/%d updated %d/%d" % (np.array(recs)[:, :, 5].sum(), np.array(recs)[:, :, 5].size, np.array(recs)[:, :, 6].sum(), np.array(recs)[:, :, 6].size))

    c4 = FC.C4()
    fr = np.stack(c4.frames)
    recs, crcs, templs = [], [], {}
    for t, roi in enumerate(c4.rois()):
        r = H.track_clip(fr, roi, rx=FC.R, ry=FC.R)
        recs.append(r["records"]); crcs.append(tcrc(r["templ"]))
        if t in (0, 100, 241, 255):
            templs[f"templ_{t}"] = r["templ"]
    np.savez_compressed(os.path.join(HERE, "filled_c4_256roi.npz"), records=np.array(recs), templ_crc=np.array(crcs, np.uint32), **templs)
    meta["c4"] = {"frames_crc": c4.crc()}
    print("c4: moved %d/%d updated %d/%d" % (np.array(recs)[:, :, 5].sum(), np.array(recs)[:, :, 5].size, np.array(recs)[:, :, 6].sum(), np.array(recs)[:, :, 6].size))

    wf = FC.WF()
    r = H.track_clip_ghc(np.stack(wf.frames), wf.roi(), rx=FC.R, ry=FC.R, start_box=wf.stale_box, lost0=1000, use_global0=True)
    np.savez_compressed(os.path.join(HERE, "filled_wf_1080p.npz"), records=r["records"], templ=r["templ"])
    meta["wf"] = {"frames_crc": wf.crc()}
    print("wf records:\n", r["records"])

    with open(os.path.join(HERE, "meta_filled.json"), "w") as fh:
        json.dump(meta, fh, indent=1)


if __name__ == "__main__":
    main()

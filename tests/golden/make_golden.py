#!/usr/bin/env python
"""Generate the golden fixtures in this directory with the REAL OpenCV (cv2 4.13.0, the version
the reference links: /root/reference/tracker/Makefile:19-27) through oracle/cv2_harness.py, which
restates /root/reference/tracker/src/main.cpp:93-169 call for call.

The reference ships no tests or golden vectors; these files are the pin.
Run in the build container only (needs cv2):  python tests/golden/make_golden.py
Inputs are NOT stored: every clip is regenerated from its seed by tools/synth.py (integer-only,
byte-identical on every host); `frames_crc` guards that assumption.
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

import cv2  # noqa: E402

from oracle import cv2_harness as H  # noqa: E402
from tools import synth  # noqa: E402

# name -> (ClipSpec kwargs, tracker kwargs, frames whose window map is stored)
CLIPS = {
    "small":    (dict(seed=1, W=320, H=240, tw=32, th=32, n_frames=24, R=80), dict(), (1, 7)),
    "lowtex":   (dict(seed=2, W=320, H=240, tw=32, th=32, n_frames=16, R=80, variant="lowtex"), dict(), (2,)),
    "lost":     (dict(seed=3, W=320, H=240, tw=32, th=32, n_frames=30, R=80, variant="lost"), dict(), ()),
    "fade":     (dict(seed=4, W=320, H=240, tw=32, th=32, n_frames=20, R=80, variant="fade"), dict(), ()),
    "border":   (dict(seed=5, W=320, H=240, tw=32, th=32, n_frames=24, R=80, variant="border"), dict(), (3,)),
    "flat":     (dict(seed=6, W=320, H=240, tw=32, th=32, n_frames=12, R=80, variant="flat"), dict(), (1,)),
    "oddsize":  (dict(seed=7, W=301, H=233, tw=37, th=29, n_frames=16, R=40), dict(rx=40, ry=24), (2,)),
    "batch4":   (dict(seed=1, W=320, H=240, tw=32, th=32, n_frames=24, R=80), dict(batch=4), ()),
    "c1_standin": (dict(seed=8, W=1280, H=720, tw=48, th=40, n_frames=151, R=80, period=60), dict(), ()),
    "c2_1080p": (dict(seed=9, W=1920, H=1080, tw=64, th=64, n_frames=13, R=80, period=48), dict(), (1,)),
    "c3_4k":    (dict(seed=10, W=3840, H=2160, tw=128, th=128, n_frames=5, R=160, period=16), dict(rx=160, ry=160), ()),
}


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def window_maps(frames, roi, recs, k, rx, ry):
    """window map of frame k with the tracker state BEFORE frame k, IPP on and off."""
    # replay to the state before frame k
    r = H.track_clip(frames[:k], roi, rx=rx, ry=ry)
    templ = r["templ"]
    x, y = (int(r["records"][-1, 0]), int(r["records"][-1, 1])) if k > 1 else (roi[0], roi[1])
    g = H.to_gray_f32(frames[k])
    out = {}
    for ipp in (True, False):
        cv2.ipp.setUseIPP(ipp)
        m = cv2.matchTemplate(g, templ, cv2.TM_CCOEFF_NORMED)
        win = H.search_window(x, y, roi[2], roi[3], m.shape[1], m.shape[0], rx, ry)
        out["ipp_on" if ipp else "ipp_off"] = m[win[1]:win[1] + win[3], win[0]:win[0] + win[2]].copy()
    cv2.ipp.setUseIPP(True)
    out["win"] = np.array(win, np.int32)
    out["templ"] = templ
    out["bbox"] = np.array([x, y], np.int32)
    return out


def main():
    meta = {"cv2": cv2.__version__, "ipp": bool(cv2.ipp.useIPP()), "clips": {}}
    assert cv2.__version__.startswith("4.13"), cv2.__version__

    # --- ingest vectors (G6): LUT and BGR->gray
    g = np.arange(256, dtype=np.uint8).reshape(1, 256)
    lut = H.to_gray_f32(g)[0]
    alt = cv2.multiply(g, 1.0, scale=float(np.float32(1.0) / np.float32(255.0)), dtype=cv2.CV_32F)[0]
    assert np.array_equal(lut, alt)
    px = (synth.hash_u32(np.arange(3 * 8192, dtype=np.uint32), 77) >> np.uint32(24)).astype(np.uint8).reshape(1, 8192, 3)
    # make sure the extremes are present
    px[0, :4] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 0, 255]]
    gray = cv2.cvtColor(px, cv2.COLOR_BGR2GRAY)
    np.savez_compressed(os.path.join(HERE, "ingest.npz"), lut=lut, bgr=px, gray=gray)

    # --- addWeighted recurrences (G5, a13)
    a = (synth.hash_u32(np.arange(4096, dtype=np.uint32), 5) >> np.uint32(8)).astype(np.float32) / np.float32(1 << 24)
    b = (synth.hash_u32(np.arange(4096, dtype=np.uint32), 6) >> np.uint32(24)).astype(np.uint8)
    b = H.to_gray_f32(b.reshape(64, 64))
    a = a.reshape(64, 64).copy()
    t = a.copy()
    snaps = {}
    for i in range(1, 151):
        cv2.addWeighted(t, 1 - H.TEMPLATE_UPDATE_LR, np.roll(b, i, 0), H.TEMPLATE_UPDATE_LR, 0.0, t)
        if i in (1, 10, 150):
            snaps[f"after{i}"] = t.copy()
    np.savez_compressed(os.path.join(HERE, "addweighted.npz"), a=a, b=b, **snaps)

    # --- map-level vectors (a3, ncc_match_cpu): small full maps, IPP on/off, plus edge cases
    sc = synth.make_clip(seed=21, W=96, H=80, tw=17, th=13, n_frames=2, R=20)
    f = H.to_gray_f32(sc["frames"][1])
    x, y, w, h = sc["roi"]
    t = H.to_gray_f32(sc["frames"][0])[y:y + h, x:x + w].copy()
    maps = {"frame": f, "templ": t}
    for ipp in (True, False):
        cv2.ipp.setUseIPP(ipp)
        maps["full_ipp_on" if ipp else "full_ipp_off"] = cv2.matchTemplate(f, t, cv2.TM_CCOEFF_NORMED)
        # flat template -> all ones ; template == frame -> 1x1 map ; flat frame -> zeros
        maps["flat_templ_" + ("on" if ipp else "off")] = cv2.matchTemplate(f, np.full((13, 17), 0.25, np.float32), cv2.TM_CCOEFF_NORMED)
        maps["self_" + ("on" if ipp else "off")] = cv2.matchTemplate(f, f.copy(), cv2.TM_CCOEFF_NORMED)
        maps["flat_frame_" + ("on" if ipp else "off")] = cv2.matchTemplate(np.full_like(f, 0.5), t, cv2.TM_CCOEFF_NORMED)
    cv2.ipp.setUseIPP(True)
    # exact ties: a frame that is periodic with period (24, 20) -> identical windows -> identical scores
    tile = H.to_gray_f32(sc["frames"][0])[:20, :24]
    per = np.tile(tile, (4, 4)).copy()
    tt = per[3:3 + 13, 5:5 + 17].copy()
    # (IPP's float noise breaks exact ties in the default build -- SURVEY.md §8(c) -- so the tie
    #  pin is the IPP-off map, plus minMaxLoc itself on hand-made tied arrays incl. an ROI view)
    cv2.ipp.setUseIPP(False)
    mp = cv2.matchTemplate(per, tt, cv2.TM_CCOEFF_NORMED)
    cv2.ipp.setUseIPP(True)
    mp_on = cv2.matchTemplate(per, tt, cv2.TM_CCOEFF_NORMED)
    _, bv, _, bl = cv2.minMaxLoc(mp)
    tied = np.zeros((5, 7), np.float32)
    tied[3, 2] = tied[1, 4] = tied[1, 5] = 0.9
    r1 = cv2.minMaxLoc(tied)
    r2 = cv2.minMaxLoc(tied[1:, 2:])
    maps.update(tie_frame=per, tie_templ=tt, tie_map=mp, tie_map_ipp_on=mp_on, tie_best=np.array([bv, bl[0], bl[1]]),
                tied=tied, tied_full=np.array([r1[1], r1[3][0], r1[3][1]]), tied_view=np.array([r2[1], r2[3][0], r2[3][1]]))
    np.savez_compressed(os.path.join(HERE, "maps.npz"), **maps)

    # --- clips (G1, G2, G5; window maps for G3/G4)
    for name, (skw, tkw, keep) in CLIPS.items():
        spec = synth.ClipSpec(**skw)
        clip = synth.make_clip(spec)
        frames, roi = clip["frames"], clip["roi"]
        r = H.track_clip(frames, roi, **tkw)
        out = {"records": r["records"], "templ": r["templ"], "roi": np.array(roi, np.int32),
               "truth": clip["truth"]}
        for k in keep:
            wm = window_maps(frames, roi, r["records"], k, tkw.get("rx", 80), tkw.get("ry", 80))
            for key, v in wm.items():
                out[f"map{k}_{key}"] = v
        np.savez_compressed(os.path.join(HERE, f"clip_{name}.npz"), **out)
        rec = r["records"]
        srch = ~np.isnan(rec[:, 4])
        meta["clips"][name] = {
            "spec": skw, "track": tkw, "frames_crc": crc(frames), "n": int(len(frames)),
            "conf_min": float(np.nanmin(rec[:, 4])), "conf_max": float(np.nanmax(rec[:, 4])),
            "moved": int(rec[:, 5].sum()), "updated": int(rec[:, 6].sum()), "searched": int(srch.sum()),
            "track_err_max": int(np.abs(rec[:, :2] - clip["truth"][1:]).max()),
        }
        print(name, meta["clips"][name])
    with open(os.path.join(HERE, "meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

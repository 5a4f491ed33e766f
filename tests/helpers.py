"""Shared helpers for the parity tests: golden loading, clip regeneration, the parity gates of
SURVEY.md §8(c) (G1..G6)."""
from __future__ import annotations

import json
import os
import zlib
from functools import lru_cache

import numpy as np

from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

TOL_SCORE = 1e-4          # BASELINE.json north_star: "NCC scores must agree within 1e-4 absolute"
TOL_LOWVAR = 5e-4         # windows with sigma_w < 0.002 (oracle's own f32 rounding of cc dominates)
AMBIGUOUS_GAP = 2e-5      # G1: a frame is "ambiguous" iff oracle top1 - top2 < 2e-5


@lru_cache(maxsize=None)
def meta():
    with open(os.path.join(GOLD, "meta.json")) as fh:
        return json.load(fh)


def golden(name: str):
    return np.load(os.path.join(GOLD, name))


@lru_cache(maxsize=4)
def clip(name: str):
    m = meta()["clips"][name]
    c = synth.make_clip(synth.ClipSpec(**m["spec"]))
    crc = zlib.crc32(np.ascontiguousarray(c["frames"]).tobytes()) & 0xFFFFFFFF
    assert crc == m["frames_crc"], f"synthetic clip {name} is not byte-identical to the one the goldens were made from"
    return c, m["track"]


def window_sigma(gray: np.ndarray, tw: int, th: int, win) -> np.ndarray:
    """population std-dev of every candidate window (float64), for the sigma_w-conditioned gates."""
    g = gray.astype(np.float64)
    S = np.zeros((g.shape[0] + 1, g.shape[1] + 1)); S[1:, 1:] = g.cumsum(0).cumsum(1)
    Q = np.zeros_like(S); Q[1:, 1:] = (g * g).cumsum(0).cumsum(1)
    x0, y0, ww, wh = win
    ys, xs = np.arange(y0, y0 + wh)[:, None], np.arange(x0, x0 + ww)[None, :]
    box = lambda I: I[ys + th, xs + tw] - I[ys, xs + tw] - I[ys + th, xs] + I[ys, xs]
    n = tw * th
    var = np.maximum(box(Q) / n - (box(S) / n) ** 2, 0)
    return np.sqrt(var)


def check_records(got: np.ndarray, want: np.ndarray, what: str):
    """G1/G2/G5: identical bbox trajectory and flags, confidence within 1e-4 on searched frames."""
    assert got.shape[0] == want.shape[0], what
    assert np.array_equal(got[:, :4].astype(np.int64), want[:, :4].astype(np.int64)), f"{what}: bbox trajectory differs"
    assert np.array_equal(got[:, 5:7].astype(np.int64), want[:, 5:7].astype(np.int64)), f"{what}: moved/updated flags differ"
    s = ~np.isnan(want[:, 4])
    assert np.array_equal(np.isnan(got[:, 4]), ~s), f"{what}: searched/held pattern differs"
    d = np.abs(got[s, 4] - want[s, 4]).max() if s.any() else 0.0
    assert d <= TOL_SCORE, f"{what}: confidence differs by {d}"
    return d

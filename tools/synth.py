"""Deterministic synthetic video clips for the NCC tracker hot path.

Everything here is integer arithmetic on numpy arrays (a counter-based 32-bit
hash, box blurs through integer cumulative sums, integer min/max stretch,
triangle-wave motion), so a (seed, geometry) pair yields byte-identical frames
on every host: the golden fixtures under tests/golden/ were produced from these
frames by the cv2 harness in the build container and are compared on the GPU
box against frames regenerated there from the same seeds.

This is test / bench infrastructure (SURVEY.md §8(d) "synthetic inputs"); it is
not part of the product library and does not touch oracle/.

Layout of a clip: frame 0 is the initialisation frame (the reference cuts the
template from it, tracker/src/main.cpp:70-71), frames 1..n-1 are tracked.
Frames are BGR u8, HWC contiguous, like cv::VideoCapture output
(tracker/src/main.cpp:95).
"""
from __future__ import annotations

import numpy as np

__all__ = ["hash_u32", "noise_u8", "box_blur_u8", "stretch_u8", "tri", "make_clip",
           "clip_frame", "ClipSpec"]


def hash_u32(idx: np.ndarray, seed: int) -> np.ndarray:
    """lowbias32-style avalanche hash of a uint32 counter array (exact, wraps mod 2^32)."""
    x = idx.astype(np.uint32, copy=True)
    x += np.uint32((seed * 0x9E3779B1 + 0x7F4A7C15) & 0xFFFFFFFF)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def noise_u8(shape, seed: int) -> np.ndarray:
    n = int(np.prod(shape))
    return (hash_u32(np.arange(n, dtype=np.uint32), seed) >> np.uint32(24)).astype(np.uint8).reshape(shape)


def _box1d(a: np.ndarray, r: int, axis: int) -> np.ndarray:
    """Integer box filter of radius r along axis with edge replication; floor division."""
    if r <= 0:
        return a
    a = np.moveaxis(a, axis, 0).astype(np.int64)
    n = a.shape[0]
    pad = np.concatenate([np.repeat(a[:1], r + 1, 0), a, np.repeat(a[-1:], r, 0)], 0)
    c = np.cumsum(pad, 0)
    out = (c[2 * r + 1:2 * r + 1 + n] - c[:n]) // (2 * r + 1)
    return np.moveaxis(out, 0, axis)


def box_blur_u8(img: np.ndarray, r: int, passes: int = 3) -> np.ndarray:
    a = img.astype(np.int64)
    for _ in range(passes):
        a = _box1d(a, r, 0)
        a = _box1d(a, r, 1)
    return a


def stretch_u8(a: np.ndarray, lo_out: int = 0, hi_out: int = 255) -> np.ndarray:
    """Per-channel integer min/max stretch to [lo_out, hi_out]."""
    a = a.astype(np.int64)
    red = tuple(range(a.ndim - 1)) if a.ndim == 3 else None
    mn = a.min(axis=red, keepdims=True)
    mx = a.max(axis=red, keepdims=True)
    span = np.maximum(mx - mn, 1)
    return (lo_out + ((a - mn) * (hi_out - lo_out)) // span).astype(np.uint8)


def tri(k: int, period: int, amp: int) -> int:
    """Integer triangle wave: period `period` frames, range [-amp, amp], tri(0)=0."""
    if period <= 0 or amp == 0:
        return 0
    q = period // 4
    if q == 0:
        return 0
    k = k % (4 * q)
    if k < q:
        v = k
    elif k < 3 * q:
        v = 2 * q - k
    else:
        v = k - 4 * q
    return (v * amp) // q


class ClipSpec:
    """Parameters of one synthetic clip; see make_clip()."""

    def __init__(self, seed=0, W=320, H=240, tw=32, th=32, n_frames=20, R=80,
                 variant="normal", period=None, margin=24, noise=3):
        self.seed, self.W, self.H, self.tw, self.th = seed, W, H, tw, th
        self.n_frames, self.R, self.variant = n_frames, R, variant
        self.period = period if period else max(8, 4 * ((n_frames + 3) // 4))
        self.margin, self.noise = margin, noise


class _Scene:
    def __init__(self, s: ClipSpec):
        self.s = s
        m = s.margin
        bg = noise_u8((s.H + 2 * m, s.W + 2 * m, 3), 11 + 7919 * s.seed)
        bg = stretch_u8(box_blur_u8(bg, 4))
        obj = noise_u8((s.th, s.tw, 3), 23 + 104729 * s.seed)
        obj = stretch_u8(box_blur_u8(obj, 1, passes=2))
        if s.variant == "lowtex":
            # contrast ramp left->right on the background: 0.4 % .. 100 % around mid grey
            x = np.arange(bg.shape[1], dtype=np.int64)
            num = 1 + (x * 255) // max(bg.shape[1] - 1, 1)          # 1..256
            bg = (128 + ((bg.astype(np.int64) - 128) * num[None, :, None]) // 256).clip(0, 255).astype(np.uint8)
        if s.variant == "flat":
            # large constant patches: exact-zero NCC cells (flat windows)
            bg[: bg.shape[0] // 2, : bg.shape[1] // 2] = 97
        self.bg, self.obj = bg, obj
        # object path: triangle-wave Lissajous, per-frame displacement <= R/4
        per = s.period
        step = max(1, s.R // 4)
        self.ax = min((s.W - s.tw) // 2 - 2, step * max(per // 4, 1))
        self.ay = min((s.H - s.th) // 2 - 2, (step * max(per // 4, 1) * 2) // 3)
        if s.variant == "border":
            # path hugs the top-left corner so the search window clamps at two borders
            self.cx0, self.cy0 = self.ax, self.ay
        else:
            self.cx0, self.cy0 = (s.W - s.tw) // 2, (s.H - s.th) // 2

    def obj_pos(self, k: int):
        s = self.s
        x = self.cx0 + tri(k, s.period, self.ax)
        y = self.cy0 + tri(k + s.period // 8, s.period, self.ay)
        if s.variant == "border":
            x, y = max(0, x - self.ax), max(0, y - self.ay)
        return int(min(max(x, 0), s.W - s.tw)), int(min(max(y, 0), s.H - s.th))

    def frame(self, k: int) -> np.ndarray:
        s = self.s
        m = s.margin
        px = m + tri(k, 2 * s.period, m - 4)
        py = m + tri(k + s.period // 2, 2 * s.period, m - 4)
        f = self.bg[py:py + s.H, px:px + s.W].astype(np.int16)
        ox, oy = self.obj_pos(k)
        show = True
        if s.variant == "lost":
            lo, hi = s.n_frames // 3, s.n_frames // 3 + max(2, s.n_frames // 5)
            show = not (lo <= (k % max(s.n_frames, 1)) < hi)
        if show:
            o = self.obj.astype(np.int16)
            if s.variant == "fade":
                # object fades toward the background: exercises the 0.40..0.70 branch
                a = max(0, 256 - 24 * k)
                o = (o * a + f[oy:oy + s.th, ox:ox + s.tw] * (256 - a)) // 256
            f[oy:oy + s.th, ox:ox + s.tw] = o
        if s.noise:
            n = hash_u32(np.arange(f.size, dtype=np.uint32), 1000 + k + 65537 * s.seed)
            n = ((n >> np.uint32(20)) % np.uint32(2 * s.noise + 1)).astype(np.int16) - s.noise
            f = f + n.reshape(f.shape)
        return np.ascontiguousarray(f.clip(0, 255).astype(np.uint8))


def make_clip(spec: ClipSpec | None = None, **kw):
    """Return dict(frames=u8[n,H,W,3], roi=(x,y,w,h), truth=int[n,2])."""
    s = spec or ClipSpec(**kw)
    sc = _Scene(s)
    frames = np.stack([sc.frame(k) for k in range(s.n_frames)])
    truth = np.array([sc.obj_pos(k) for k in range(s.n_frames)], dtype=np.int32)
    x0, y0 = sc.obj_pos(0)
    return {"frames": frames, "roi": (x0, y0, s.tw, s.th), "truth": truth, "spec": s}


def clip_frame(spec: ClipSpec, k: int) -> np.ndarray:
    """One frame of a clip without materialising the rest (bench ring construction)."""
    return _Scene(spec).frame(k)


class Scene(_Scene):
    """Public handle: build once, call .frame(k) / .obj_pos(k) many times."""

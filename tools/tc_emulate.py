#!/usr/bin/env python
"""Numerics emulation of a tcgen05 (kind::f16) implicit-GEMM cross term for the NCC search -- BASELINE.json north_star item 3,
VERDICT r1 "decide the tensor-core variant with data", step 1.  CPU only (numpy); uses the test oracle as the yardstick, so
this is a measurement tool, not product code.

Formulation emulated (see DESIGN.md "Tensor-core variant"):
    frame operand   A = g, the 8-bit gray level itself (the path's frame is f = fl32(g * 1/255), utils.hpp:8,12): EXACT in fp16/bf16
    template operand B = the centred template tc = fl32(t - mean_t), scaled by 2^12 and split into TWO fp16 terms t1 + t2
                        (|residual| <= 2^-23 |tc|); every product g * t_i is exact in FP32 (8 + 11 significant bits)
    accumulation     FP32 in TMEM.  Per template row dy, per split term, the 64 taps of a candidate fall into 4-5 MMA K-blocks of
                     16 image columns (banded-Toeplitz B operand; which taps share a block depends on the candidate's x mod 16).
                     Pessimistic hardware model: a K-block's 16 products are summed exactly, then every add into the accumulator
                     TRUNCATES toward zero (tensor-core accumulation is not round-to-nearest).
    cross term       cc = acc / (255 * 2^12); normalisation exactly as the FP32 path (OpenCV's operation order in FP64).
Reported per golden map / clip: max |score - oracle| (the 1e-4 gate), degenerate cells, peak identity, trajectories.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from tests import helpers as Hp  # noqa: E402

SCALE = 4096.0


def rz32(x: np.ndarray) -> np.ndarray:
    """float64 -> float32 rounding toward zero."""
    r = x.astype(np.float32)
    over = np.abs(r.astype(np.float64)) > np.abs(x)
    r[over] = np.nextafter(r[over], np.float32(0))
    return r


def split_fp16(tc: np.ndarray):
    s = tc.astype(np.float64) * SCALE
    t1 = s.astype(np.float16)
    t2 = (s - t1.astype(np.float64)).astype(np.float16)
    return t1.astype(np.float64), t2.astype(np.float64)


def tc_cross(gray_u8: np.ndarray, tc: np.ndarray, win, truncate=True) -> np.ndarray:
    """Emulated tensor-core cross term  sum f * tc  over the window `win` = (x0, y0, ww, wh)."""
    x0, y0, ww, wh = win
    th, tw = tc.shape
    t1, t2 = split_fp16(tc)
    G = gray_u8[y0:y0 + wh + th - 1, x0:x0 + ww + tw - 1].astype(np.float64)
    acc = np.zeros((wh, ww), np.float32)
    rnd = rz32 if truncate else (lambda v: v.astype(np.float32))
    for a in range(16):                                   # candidates x = a (mod 16) share their K-block boundaries
        xs = np.arange(a, ww, 16)
        if xs.size == 0:
            continue
        cuts = [0] + [c for c in range(16 - a, tw, 16) if c > 0]
        cuts = sorted(set(cuts))
        A = np.zeros((wh, xs.size), np.float32)
        for dy in range(th):
            rows = G[dy:dy + wh]                                              # [wh, tileW]
            W = np.lib.stride_tricks.sliding_window_view(rows, tw, axis=1)[:, xs, :]   # [wh, nx, tw]
            for t in (t1, t2):
                P = W * t[dy][None, None, :]                                  # exact products
                blocks = np.add.reduceat(P, cuts, axis=2)                     # exact K-block sums (<= 16 products each)
                for b in range(blocks.shape[2]):
                    A = rnd(A.astype(np.float64) + rnd(blocks[:, :, b]).astype(np.float64))
        acc[:, xs] = A
    return (acc.astype(np.float64) / (255.0 * SCALE)).astype(np.float32)


def quantise_i8x2(tc: np.ndarray):
    """The product kernel's template operand (kind::i8): q = round(tc * 2^k) with 2^k * max|tc| <= 127 * 256, split into two
    signed 8-bit digits q = 256 * d1 + d0.  Returns (q as int64, 2^-k, sum(q) * 2^-k - sum(tc) in float64)."""
    m = float(np.abs(tc).max())
    if m == 0.0:
        return np.zeros(tc.shape, np.int64), 1.0, 0.0
    k = int(np.floor(np.log2(32512.0 / m)))
    q = np.rint(tc.astype(np.float64) * 2.0 ** k).astype(np.int64)
    d1 = (q + 128) >> 8
    d0 = q - 256 * d1
    assert d1.min() >= -128 and d1.max() <= 127 and d0.min() >= -128 and d0.max() <= 127
    return q, 2.0 ** -k, float(q.sum()) * 2.0 ** -k - float(tc.astype(np.float64).sum())


def i8_cross(gray_u8: np.ndarray, gray: np.ndarray, tc: np.ndarray, win, n_templ: int) -> np.ndarray:
    """EXACT integer cross term of the quantised template (what int32 accumulation in TMEM gives, in any order), scaled by
    the ingest constant fl32(1/255), with the quantisation's DC part removed through the window sum:
        cc = c255 * 2^-k * sum g q  -  (wsum / N) * (sum(q) 2^-k - sum(tc))"""
    x0, y0, ww, wh = win
    th, tw = tc.shape
    q, inv, dc = quantise_i8x2(tc)
    G = gray_u8[y0:y0 + wh + th - 1, x0:x0 + ww + tw - 1].astype(np.int64)
    W = np.lib.stride_tricks.sliding_window_view(G, (th, tw))
    acc = np.einsum("yxij,ij->yx", W, q)                                   # exact in int64
    g = gray.astype(np.float64)
    S = np.zeros((g.shape[0] + 1, g.shape[1] + 1)); S[1:, 1:] = g.cumsum(0).cumsum(1)
    ys, xs = np.arange(y0, y0 + wh)[:, None], np.arange(x0, x0 + ww)[None, :]
    wsum = S[ys + th, xs + tw] - S[ys, xs + tw] - S[ys + th, xs] + S[ys, xs]
    c255 = float(np.float32(1.0) / np.float32(255.0))
    return (acc.astype(np.float64) * (c255 * inv) - wsum * (dc / n_templ)).astype(np.float32)


def normalise(gray: np.ndarray, templ: np.ndarray, cc: np.ndarray, win) -> np.ndarray:
    """OpenCV TM_CCOEFF_NORMED finalisation with the cross term of the CENTRED template given."""
    x0, y0, ww, wh = win
    th, tw = templ.shape
    n = tw * th
    g = gray.astype(np.float64)
    S = np.zeros((g.shape[0] + 1, g.shape[1] + 1)); S[1:, 1:] = g.cumsum(0).cumsum(1)
    Q = np.zeros_like(S); Q[1:, 1:] = (g * g).cumsum(0).cumsum(1)
    ys, xs = np.arange(y0, y0 + wh)[:, None], np.arange(x0, x0 + ww)[None, :]
    box = lambda I: I[ys + th, xs + tw] - I[ys, xs + tw] - I[ys + th, xs] + I[ys, xs]
    wsum, wsq = box(S), box(Q)
    mean, sd = O.mean_stddev(templ)
    if sd * sd < np.finfo(np.float64).eps:
        return np.ones((wh, ww), np.float32)
    tn = np.sqrt(sd * sd) / np.sqrt(1.0 / n)
    diff2 = np.maximum(wsq - (wsum * wsum) * (1.0 / n), 0)
    t = np.where(diff2 <= np.minimum(0.5, 10 * np.finfo(np.float32).eps * wsq), 0.0, np.sqrt(diff2) * tn)
    num = cc.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(np.abs(num) < t, num / np.where(t == 0, 1, t), np.where(np.abs(num) < t * 1.125, np.sign(num), 0.0))
    return r.astype(np.float32)


def centred(templ):
    mean, _ = O.mean_stddev(templ)
    return (templ.astype(np.float64) - mean).astype(np.float32)


MODEL = "f16x2"   # or "i8x2"


def tc_map(frame_bgr, templ, win):
    g8 = O.bgr2gray(frame_bgr)
    gray = O.gray_to_f32(g8)
    if MODEL == "i8x2":
        return normalise(gray, templ, i8_cross(g8, gray, centred(templ), win, templ.size), win), gray
    return normalise(gray, templ, tc_cross(g8, centred(templ), win), win), gray


def main():
    global MODEL
    if len(sys.argv) > 1:
        MODEL = sys.argv[1]
    out = {"model": MODEL, "maps": [], "clips": []}
    # ---- golden window maps (cv2 IPP-off is the exact formula)
    for name, k in [("small", 1), ("small", 7), ("lowtex", 2), ("border", 3), ("flat", 1), ("oddsize", 2), ("c2_1080p", 1)]:
        (c, tk) = Hp.clip(name)
        g = Hp.golden(f"clip_{name}.npz")
        templ, win = g[f"map{k}_templ"], tuple(int(v) for v in g[f"map{k}_win"])
        m, gray = tc_map(c["frames"][k], templ, win)
        off, on = g[f"map{k}_ipp_off"], g[f"map{k}_ipp_on"]
        sig = Hp.window_sigma(gray, templ.shape[1], templ.shape[0], win)
        d = np.abs(m - off)
        deg = (off == 0) | (np.abs(off) == 1)
        rec = {"map": f"{name}/{k}", "max_abs_diff_vs_exact": float(d.max()), "max_where_sigma_ge_0.002": float(d[sig >= 0.002].max(initial=0)),
               "max_where_sigma_lt_0.002": float(d[sig < 0.002].max(initial=0)), "degenerate_cells_identical": bool(np.array_equal(m[deg], off[deg])),
               "peak_identical": bool(np.argmax(m) == np.argmax(on)),
               "fp32_oracle_path_max_diff": float(np.abs(O.ncc_window(gray, templ, *win) - off).max())}
        rec["G3"] = rec["max_where_sigma_ge_0.002"] <= Hp.TOL_SCORE and rec["max_where_sigma_lt_0.002"] <= Hp.TOL_LOWVAR
        out["maps"].append(rec)
        print(rec, flush=True)
    # ---- whole clips: the tracker loop (main.cpp:135-161) on emulated maps vs the cv2 golden records
    for name in ["small", "lowtex", "lost", "fade", "border", "flat", "oddsize", "c2_1080p"]:
        (c, tk) = Hp.clip(name)
        gold = Hp.golden(f"clip_{name}.npz")
        frames, roi = c["frames"], c["roi"]
        rx, ry = tk.get("rx", 80), tk.get("ry", 80)
        x, y, w, h = roi
        templ = O.to_gray_f32(frames[0])[y:y + h, x:x + w].copy()
        recs = []
        for k in range(1, len(frames)):
            gray = O.to_gray_f32(frames[k])
            win = O.search_window(x, y, w, h, gray.shape[1] - w + 1, gray.shape[0] - h + 1, rx, ry)
            m, _ = tc_map(frames[k], templ, win)
            best, bx, by = O.max_loc(m)
            bx, by = bx + win[0], by + win[1]
            moved = updated = 0
            if best >= 0.40:
                x, y, moved = bx, by, 1
                if best >= 0.70:
                    templ = O.add_weighted(templ, gray[y:y + h, x:x + w])
                    updated = 1
            recs.append((x, y, w, h, best, moved, updated))
        recs = np.array(recs, np.float64)
        want = gold["records"]
        same_box = bool(np.array_equal(recs[:, :4], want[:, :4]) and np.array_equal(recs[:, 5:7], want[:, 5:7]))
        rec = {"clip": name, "frames": len(recs), "trajectory_and_flags_identical": same_box, "max_conf_diff": float(np.abs(recs[:, 4] - want[:, 4]).max()),
               "final_template_bit_identical": bool(np.array_equal(templ, gold["templ"]))}
        out["clips"].append(rec)
        print(rec, flush=True)
    with open(os.path.join(ROOT, "profiles", "tc_emulation_%s_r2.json" % MODEL), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()

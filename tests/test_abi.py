"""CPU: the C-ABI library loads, exports every symbol include/pvt.h declares, and fails loudly
(never falls back) when no CUDA device is present.  No compute calls here."""
import ctypes as C
import importlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pvt = importlib.import_module("parallel-video-object-tracker_b200")


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pvt.h")).read()
    return sorted(set(re.findall(r"PVT_API\s+[\w\s\*]+?\b(pvt_\w+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(pvt.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = pvt.lib()
    for s in _declared_symbols():
        assert hasattr(L, s), f"libpvt.so does not export {s}"
    assert L.pvt_version() == 200


def test_default_params_are_the_reference_constants():
    p = pvt.default_params()          # tracker/src/main.cpp:6-20
    assert (p.search_radius_x, p.search_radius_y) == (80, 80)
    assert (p.ncc_min_confidence, p.ncc_strong_confidence, p.template_update_lr) == (0.40, 0.70, 0.10)
    assert p.batch_size == 4 and p.mode == pvt.MODE_NAIVE and p.kernel == pvt.KERNEL_AUTO
    assert p.lost_frame_threshold == 0                                    # tracker/ has no lost-object logic
    g = pvt.default_params(ghc=True)  # tracker_ghc/src/main.cpp:9-23
    assert (g.search_radius_x, g.search_radius_y, g.lost_frame_threshold) == (60, 60, 50)
    assert (g.ncc_min_confidence, g.ncc_global_confidence, g.ncc_strong_confidence, g.template_update_lr) == (0.40, 0.60, 0.70, 0.10)


def test_struct_layouts():
    assert C.sizeof(pvt.Result) == 32 and pvt.RESULT_DTYPE.itemsize == 32
    assert C.sizeof(pvt.Frame) == 32
    assert C.sizeof(pvt.Params) == 8 + 24 + 5 * 4 + 4 + 8 + 8   # + lost_frame_threshold, formula, reserved, ncc_global_confidence
    assert C.sizeof(pvt.Config) == 16 * 4


def test_cpu_mode_is_rejected_not_emulated():
    # argument validation happens before any CUDA call, so this holds with or without a GPU
    with pytest.raises(pvt.PvtError) as e:
        pvt.Tracker(64, 64, 8, 8, mode=pvt.MODE_CPU)
    assert e.value.code == pvt.ERR_UNSUPPORTED
    f = np.zeros((16, 16), np.float32)
    with pytest.raises(pvt.PvtError) as e:
        pvt.ncc_match_cpu(f, f[:4, :4])
    assert e.value.code == pvt.ERR_UNSUPPORTED


def test_no_gpu_fails_loudly():
    try:
        n = pvt.device_count()
    except pvt.PvtError as e:
        n = 0
        assert e.code == pvt.ERR_CUDA
    if n > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pvt.PvtError) as e:
        pvt.Tracker(64, 64, 8, 8)
    assert e.value.code == pvt.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_invalid_arguments_return_codes():
    for kw in (dict(frame_w=0, frame_h=8), dict(frame_w=8, frame_h=8, max_templ_w=9)):
        a = dict(frame_w=64, frame_h=64, max_templ_w=8, max_templ_h=8)
        a.update(kw)
        with pytest.raises(pvt.PvtError) as e:
            pvt.Tracker(**a)
        assert e.value.code == pvt.ERR_INVALID
    with pytest.raises(pvt.PvtError) as e:
        pvt.Tracker(64, 64, 8, 8, search_radius_x=-1)
    assert e.value.code == pvt.ERR_INVALID
    with pytest.raises(pvt.PvtError) as e:
        pvt.Tracker(64, 64, 8, 8, formula=7)                      # pvt_formula is validated before any device is touched
    assert e.value.code == pvt.ERR_INVALID and "formula" in str(e.value)
    f = np.zeros((16, 16), np.float32)
    with pytest.raises(pvt.PvtError) as e:
        pvt.ncc_match_naive_cuda_batched([f], f[:4, :4], formula=7)
    assert e.value.code == pvt.ERR_INVALID
    with pytest.raises(pvt.PvtError) as e:                        # the eps flag does not unlock the CPU mode either
        pvt._ncc_match(pvt.MODE_CPU, f, f[:4, :4], formula=pvt.FORMULA_EPS)
    assert e.value.code == pvt.ERR_UNSUPPORTED


def test_product_never_touches_the_oracle():
    """The product path (package + csrc) must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "parallel-video-object-tracker_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in src and "ncc_oracle" not in src, os.path.join(dp, f)
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), os.path.join(dp, f)

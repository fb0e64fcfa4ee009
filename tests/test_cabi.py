"""CPU: the C-ABI library loads and exports every symbol include/rtc.h declares; host-only entry
points work without a GPU; GPU entry points fail loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rtc.h")).read()
    return sorted(set(re.findall(r"RTC_API\s+[\w\s\*]+?\b(rtc_\w+)\s*\(", src)))


def test_header_symbols_exported(rtc):
    L = rtc.load_library()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"librtc_b200.so does not export {s}"
    assert sorted(rtc.EXPORTS) == syms               # the Python binding tracks the header


def test_header_constants_match_binding(rtc):
    """Flag, mode and gather values of the Python binding == include/rtc.h."""
    src = open(os.path.join(ROOT, "include", "rtc.h")).read()
    flags = {m.group(1): 1 << int(m.group(2)) for m in re.finditer(r"RTC_FLAG_(\w+)\s*=\s*1u\s*<<\s*(\d+)", src)}
    assert len(flags) >= 6
    for name, value in flags.items():
        assert getattr(rtc, "FLAG_" + name) == value, name
    assert len(set(flags.values())) == len(flags)
    modes = re.search(r"typedef enum rtc_mode \{(.*?)\}", src, re.S).group(1)
    values = dict(re.findall(r"RTC_([A-Z_0-9]+)\s*=\s*(\d+)", modes))
    assert sorted(values) == ["BIT_ASCII", "BIT_PIXEL", "RGB_ASCII", "RGB_NORMALS", "RGB_PIXEL", "SDL"]
    for name, value in values.items():
        assert getattr(rtc, name) == int(value), name
    for name in ("GATHER_HOST", "GATHER_P2P"):
        assert getattr(rtc, name) == int(re.search(r"RTC_%s\s*=\s*(\d+)" % name, src).group(1))


def test_pod_layouts(rtc):
    import ctypes
    assert ctypes.sizeof(rtc.RtcParams) == 96
    assert rtc.OBJECT_DTYPE.itemsize == 64


def test_host_only_entry_points(rtc):
    assert rtc.encode_capacity(7681, 4320, rtc.RGB_PIXEL) >= 20 * 7680 * 4320 + 4320
    assert rtc.encode_capacity(400, 150, rtc.BIT_ASCII) >= 12 * 399 * 150 + 150
    L = rtc.load_library()
    assert [L.rtc_mode_bpp(m) for m in range(5)] == [1, 1, 3, 3, 3]
    assert [L.rtc_mode_has_glyph(m) for m in range(5)] == [1, 0, 1, 0, 0]
    p = rtc.camera_params(400, 150, (0, 0, 0), (0, np.float32(np.pi), 0))
    assert abs(p.element1 - 0.866025388) < 1e-7 and abs(p.element2 - 0.577350259) < 1e-7


def test_no_cpu_fallback(rtc):
    """Without a CUDA device the product path must fail loudly."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(rtc.RtcError, match="no usable CUDA device|CUDA"):
        rtc.Context(0)
    with pytest.raises(rtc.RtcError, match="no usable CUDA device|CUDA"):
        rtc.MultiGpu([0, 0])                              # the multi-GPU frame driver has no CPU fallback either


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "raytracing-in-windows-console_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle." not in txt.replace("oracle.py", "") or f == "__init__.py" and "oracle" not in txt, f
                assert "rt_oracle" not in txt and "libref_cpu" not in txt, f

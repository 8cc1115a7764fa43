"""CPU-side checks of the drop-in boundary: libb2points.so loads, exports every symbol include/b2points.h declares, and the
product path fails LOUDLY without a CUDA device (there is no CPU fallback anywhere in the library)."""
import ctypes as C
import os
import re

import pytest

from gtsam_points_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b2points.h")).read()
    return sorted(set(re.findall(r"B2_API\s+[\w\s\*]+?\b(b2_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/b2points.h but not exported by libb2points.so"
    # the Python mirror binds exactly the declared surface (nothing undeclared, nothing missing)
    assert set(names) == set(capi.SIGNATURES)
    assert capi.lib().b2_version().decode().startswith("b2points")


def test_record_layout_matches_header():
    hdr = open(os.path.join(ROOT, "include", "b2points.h")).read()
    assert int(re.search(r"#define B2_LINEARIZED_DOUBLES (\d+)", hdr).group(1)) == capi.B2_LINEARIZED_DOUBLES == 128
    # H_target 36 | H_source 36 | H_target_source 36 | b_target 6 | b_source 6 | error | num_inliers | reserved 6
    assert 36 * 3 + 6 * 2 + 2 + 6 == capi.B2_LINEARIZED_DOUBLES


def test_no_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    st = capi.lib().b2_ctx_create(0, None, C.byref(h))
    assert st == 5  # B2_ERR_NO_DEVICE
    assert not h.value
    assert b"no CUDA device" in capi.lib().b2_last_error()
    import gtsam_points_b200 as g

    with pytest.raises(capi.B2Error):
        g.Context(0)

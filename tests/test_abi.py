"""The C-ABI library loads, exports every symbol include/silent_b200.h declares, and its host-side logic (plan geometry,
tap tables, argument validation) agrees with the oracle. No compute call needs a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from pysilent_b200 import _lib


def header_functions():
    text = open(os.path.join(ROOT, "include", "silent_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(silent_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    declared = header_functions()
    assert len(declared) >= 25
    assert sorted(_lib.SYMBOLS) == declared
    for name in declared:
        assert hasattr(built_lib, name), name
    assert built_lib.silent_abi_version() == 1


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "silent_b200.h"\nint main(void){silent_params p; (void)p; return SILENT_OK;}\n')
    import subprocess
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])


def _plan(L, H, W, cw, ch, scale, c=3, nc=3, dtype=0):
    p = _lib.SilentParams(H, W, c, nc, cw, ch, float(scale), dtype, 0)
    h = ctypes.c_void_p()
    rc = L.silent_plan_create(ctypes.byref(p), ctypes.byref(h))
    return rc, h


@pytest.mark.parametrize("H,W,cw,ch,scale", [(1080, 1920, 288, 192, 2 ** .5), (480, 640, 288, 192, 1.3),
                                             (480, 640, 288, 192, np.e ** .5), (720, 1280, 288, 192, 2 ** .5),
                                             (2160, 3840, 288, 192, 2 ** .5), (97, 131, 24, 16, 2 ** .5),
                                             (61, 53, 20, 12, 1.3), (50, 64, 16, 16, 1.5)])
def test_plan_geometry_and_tables_equal_oracle(built_lib, c_oracle, H, W, cw, ch, scale):
    L = built_lib
    rc, h = _plan(L, H, W, cw, ch, scale)
    assert rc == 0, L.silent_last_error()
    t = c_oracle.pyramid_tables((H, W), (cw, ch), float(scale))
    assert L.silent_plan_levels(h) == t["L"]
    from oracle import silent_oracle as lit
    for s in range(t["L"]):
        vals = [ctypes.c_int() for _ in range(6)]
        assert L.silent_plan_level_info(h, s, *[ctypes.byref(v) for v in vals]) == 0
        (y0, y1), (x0, x1) = lit.level_crop((H, W), (cw, ch), float(scale), s)
        assert [v.value for v in vals] == [y0, y1, x0, x1, t["valid"][s][0], t["valid"][s][1]]
        iy, wy = np.empty((ch, 6), np.int32), np.empty((ch, 6), np.float32)
        ix, wx = np.empty((cw, 6), np.int32), np.empty((cw, 6), np.float32)
        assert L.silent_plan_level_tables(h, s, iy.ctypes.data, wy.ctypes.data, ix.ctypes.data, wx.ctypes.data) == 0
        oky, okx = t["oky"][s].astype(bool), t["okx"][s].astype(bool)
        assert np.array_equal(iy[:, 0] >= 0, oky) and np.array_equal(ix[:, 0] >= 0, okx)
        assert np.array_equal(iy[oky], t["iy"][s][oky]) and np.array_equal(wy[oky], t["wy"][s][oky])   # bitwise
        assert np.array_equal(ix[okx], t["ix"][s][okx]) and np.array_equal(wx[okx], t["wx"][s][okx])
    L.silent_plan_destroy(h)


def test_algorithmic_bytes_match_survey(built_lib):
    # C2/C3 13.24 MB, C1 6.11 MB, C5 9.13 MB per frame (uint8 in, two fp32 output tensors)
    for (H, W, scale, want) in [(1080, 1920, 2 ** .5, 13.24e6), (480, 640, 1.3, 6.11e6), (720, 1280, 2 ** .5, 9.13e6)]:
        rc, h = _plan(built_lib, H, W, 288, 192, scale)
        assert rc == 0
        got = built_lib.silent_plan_algorithmic_bytes(h)
        assert abs(got - want) / want < 2e-3, (H, W, got)
        built_lib.silent_plan_destroy(h)


def test_argument_validation_without_gpu(built_lib):
    L = built_lib
    for bad in [dict(scale=1.0), dict(nc=0), dict(cw=0), dict(H=0), dict(nc=4), dict(dtype=7)]:
        kw = dict(H=100, W=100, cw=10, ch=10, scale=1.5, c=3, nc=3, dtype=0)
        kw.update(bad)
        rc, _ = _plan(L, kw["H"], kw["W"], kw["cw"], kw["ch"], kw["scale"], kw["c"], kw["nc"], kw["dtype"])
        assert rc in (-1, -2), bad
        assert len(L.silent_last_error()) > 0
    rc, h = _plan(L, 10, 10, 288, 192, 1.5)       # image smaller than the centre: zero levels, still a valid plan
    assert rc == 0 and L.silent_plan_levels(h) == 0
    L.silent_plan_destroy(h)
    assert L.silent_conv2d(None, 1, 4, 4, 3, None, 3, 3, 0, 0.0, None, None) == -1
    assert L.silent_conv2d(ctypes.c_void_p(8), 1, 4, 4, 3, ctypes.c_void_p(8), 4, 3, 0, 0.0, ctypes.c_void_p(8), None) == -2
    assert L.silent_conv2d(ctypes.c_void_p(8), 1, 4, 4, 9, ctypes.c_void_p(8), 3, 3, 0, 0.0, ctypes.c_void_p(8), None) == -2
    assert L.silent_selection_workspace_bytes(0, 4, 4) == 0 and L.silent_selection_workspace_bytes(6, 192, 288) > 0
    w = _lib.SilentStackWeights()
    assert L.silent_stack_fused(None, 1, 4, 4, ctypes.byref(w), None, None, None, None, 0, None) == -1
    assert L.silent_stack_workspace_bytes(6, 192, 288) >= 3 * 192 * 288 * 8 and L.silent_stack_workspace_bytes(0, 1, 1) == 0
    assert L.silent_stack_fused(ctypes.c_void_p(16), 2, 4, 4, ctypes.byref(w), None, None, None, None, 0, None) == -3
    assert L.silent_launch_count() >= 0


def test_fused_stack_structure_check_is_host_side(built_lib):
    from pysilent_b200 import LineEndPipeline
    f = LineEndPipeline().filters()
    blur = f["blur"].copy()
    blur[3, 3, 1, 2] *= 2
    w = _lib.make_stack_weights(f["rgc"], f["rgby"], f["stripe"], blur, f["end"])
    rc = built_lib.silent_stack_fused(ctypes.c_void_p(16), 1, 4, 4, ctypes.byref(w), None, None, None,
                                      ctypes.c_void_p(256), 1 << 20, None)
    assert rc == -5 and b"blur filter slices differ" in built_lib.silent_last_error()
    with pytest.raises(ValueError):
        _lib.make_stack_weights(f["rgc"], f["rgby"], f["stripe"], f["blur"][:5, :5], f["end"])

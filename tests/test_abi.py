"""CPU: the C-ABI library loads, exports every symbol include/gvv_b200.h declares, and rejects
bad arguments with the documented codes -- without ever touching a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from gvv_differentiable_cuda_renderer_b200 import _native


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "gvv_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gvv_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = _native.lib()
    syms = header_symbols()
    assert len(syms) >= 9
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/gvv_b200.h but not exported"
    assert sorted(_native.EXPORTS) == syms, "python binding and header disagree on the symbol list"


def test_sass_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def _desc(**kw):
    faces = np.array([0, 1, 2], np.int32)
    d = dict(faces=faces.ctypes.data, num_faces=1, texcoords=None, num_vertices=3, num_cameras=1, width=8, height=8,
             albedo_mode=0, shading_mode=0, image_filter_size=1, texture_filter_size=1, compute_normal_map=0, device=0)
    d.update(kw)
    return _native.gvv_desc(**d), faces


@pytest.mark.parametrize("kw,msg", [
    (dict(num_vertices=0), "number_of_vertices not set!"),
    (dict(num_cameras=0), "number_of_cameras not set!"),
    (dict(width=0), "render_resolution_u not set!"),
    (dict(height=-1), "render_resolution_v not set!"),
    (dict(albedo_mode=7), "INVALID ALBEDO MODE"),
    (dict(shading_mode=2), "INVALID SHADING MODE"),
    (dict(albedo_mode=1), "textured albedo needs texture_coordinates"),
    (dict(num_vertices=2), "references vertex"),
    (dict(width=8000, height=8000), "more than 40960 tiles"),          # rejected at create, not at the first launch
])
def test_create_rejects_bad_attributes(kw, msg):
    L = _native.lib()
    d, keep = _desc(**kw)
    h = ctypes.c_void_p()
    rc = L.gvv_create(ctypes.byref(d), ctypes.byref(h))
    assert rc == 1 and not h.value          # GVV_EINVAL, no handle
    assert msg in L.gvv_last_error().decode()


def test_null_handle_calls_fail_cleanly():
    L = _native.lib()
    assert L.gvv_forward(None, 1, 1, 1, *([None] * 14)) == 1
    assert L.gvv_backward(None, 1, 1, 1, *([None] * 17)) == 1
    assert L.gvv_destroy(None) == 0
    assert L.gvv_reserve(None, 1, None) == 1
    assert L.gvv_launch_count(None) == 0
    out = ctypes.c_double()
    assert L.gvv_bench_atomics(0, 5, 1, 1, 1, ctypes.byref(out)) == 1


def test_python_layer_rejects_bad_modes_before_touching_cuda():
    with pytest.raises(_native.GvvError, match="INVALID ALBEDO MODE"):
        _native.NativeRenderer([0, 1, 2], None, 3, 1, 8, 8, "phong", "shaded")
    with pytest.raises(_native.GvvError, match="INVALID SHADING MODE"):
        _native.NativeRenderer([0, 1, 2], None, 3, 1, 8, 8, "vertexColor", "flat")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gvv_differentiable_cuda_renderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower().replace("# oracle", ""), f"{f} mentions the oracle"

"""CPU: this repo's OBJ / calibration readers pinned against the REFERENCE's own readers on the
reference's bundled fixtures.

  1. tests/golden/io_readers.json -- digests of what python/utils/{OBJReader,CameraReader}.py read from
     python/data/{triangle,cone,magdalena}.obj and the two calibration files (tools/make_io_golden.py);
  2. live, when /root/reference is present (authoring container): the reference modules imported and run side by side.

The fixtures themselves are staged in tests/_refdata (tools/stage_ref_data.py); without them the tests skip.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import refdata
from gvv_differentiable_cuda_renderer_b200.io import CameraReader, OBJReader

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io_readers.json")
REF_ROOT = os.environ.get("GVV_REFERENCE", "/root/reference")
need_data = pytest.mark.skipif(not refdata.available(), reason="tests/_refdata not staged (tools/stage_ref_data.py)")


def sha(a, dtype):
    a = np.ascontiguousarray(np.asarray(a, dtype=dtype))
    return hashlib.sha1(a.astype("<" + np.dtype(dtype).str[1:]).tobytes()).hexdigest(), list(a.shape)


OBJ_FIELDS = (("facesVertexId", np.int32), ("facesTextureId", np.int32), ("vertexCoordinates", np.float32), ("vertexColors", np.float32),
              ("pertVertexTextureCoordinate", np.float32), ("textureCoordinates", np.float32), ("textureMap", np.float32), ("vertexLabels", np.int32))


@need_data
@pytest.mark.parametrize("name", ["triangle.obj", "cone.obj", "magdalena.obj"])
def test_obj_reader_equals_reference_reader_golden(name):
    g = json.load(open(GOLDEN))["obj"][name]
    r = OBJReader(refdata.DATA + "/" + name)
    assert r.numberOfVertices == g["numberOfVertices"]
    assert (r.texHeight, r.texWidth) == (g["texHeight"], g["texWidth"])
    for field, dt in OBJ_FIELDS:
        h, shape = sha(getattr(r, field), dt)
        assert shape == g[field]["shape"], field
        assert h == g[field]["sha1"], f"{name}: {field} differs from what the reference reader reads"
    if "numberOfEdges" in g:                      # the reference's adjacency only survives numpy 2 on triangle.obj
        assert r.numberOfEdges == g["numberOfEdges"] and r.maximumNumNeighbours == g["maximumNumNeighbours"]
        assert sha(r.numberOfNeigbours, np.float32)[0] == g["numberOfNeigbours"]["sha1"]
        assert [list(x) for x in r.compressedAdjacency] == g["neighbourSets"]


@need_data
@pytest.mark.parametrize("key", ["cameras.calibration@1024x1024", "cameras.calibration@512x512", "cameras.calibration@640x360",
                                 "monocular.calibration@1024x1024", "monocular.calibration@512x512", "monocular.calibration@640x360"])
def test_camera_reader_equals_reference_reader_golden(key):
    g = json.load(open(GOLDEN))["cam"][key]
    name, res = key.split("@")
    u, v = (int(x) for x in res.split("x"))
    c = CameraReader(os.path.join(refdata.DATA, name), u, v)
    assert c.numberOfCameras == g["numberOfCameras"]
    assert sha(c.extrinsics, np.float32)[0] == g["extrinsics"]["sha1"]
    assert sha(c.intrinsics, np.float32)[0] == g["intrinsics"]["sha1"]
    assert [float(x) for x in c.intrinsics] == g["intrinsics_f64"]            # same float64 values before the op casts them
    assert c.originalSizeU == g["originalSizeU"] and c.originalSizeV == g["originalSizeV"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "python", "utils")), reason="reference checkout not present")
@pytest.mark.parametrize("name", ["triangle.obj", "cone.obj", "magdalena.obj"])
def test_obj_reader_equals_reference_reader_live(name):
    import importlib.util
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import make_io_golden
    mod = make_io_golden.load_ref(REF_ROOT, "OBJReader")
    path = os.path.join(REF_ROOT, "python", "data") + "/" + name
    ref = make_io_golden.read_obj_with_reference(mod, path)
    mine = OBJReader(path)
    for field, dt in OBJ_FIELDS:
        a, b = np.asarray(getattr(mine, field), dt), np.asarray(getattr(ref, field), dt)
        assert a.shape == b.shape and np.array_equal(a, b), field
    cam_mod = make_io_golden.load_ref(REF_ROOT, "CameraReader")
    for cal in ("cameras.calibration", "monocular.calibration"):
        p = os.path.join(REF_ROOT, "python", "data", cal)
        a, b = CameraReader(p, 800, 600), cam_mod.CameraReader(p, 800, 600)
        assert a.numberOfCameras == b.numberOfCameras and list(a.extrinsics) == list(b.extrinsics)
        assert [float(x) for x in a.intrinsics] == [float(x) for x in b.intrinsics]


@need_data
def test_bundled_scenes_have_the_documented_shapes():
    """cone 5 v / 6 f, magdalena 5,118 v / 10,115 f / 5,926 vt with a 1024^2 texture and max
    vertex degree 19, one camera at 1024^2 with fx = 752.69, fy = 752.83, cx = 523.57, cy = 502.70."""
    cone = OBJReader(refdata.DATA + "/cone.obj")
    assert cone.numberOfVertices == 5 and len(cone.facesVertexId) == 18
    mag = OBJReader(refdata.DATA + "/magdalena.obj")
    assert mag.numberOfVertices == 5118 and len(mag.facesVertexId) == 3 * 10115 and len(mag.pertVertexTextureCoordinate) == 5926
    assert (mag.texHeight, mag.texWidth) == (1024, 1024)
    deg = np.bincount(np.asarray(mag.facesVertexId), minlength=5118)
    assert deg.max() == 19
    cam = CameraReader(os.path.join(refdata.DATA, "cameras.calibration"), 1024, 1024)
    K = np.asarray(cam.intrinsics).reshape(3, 3)
    assert cam.numberOfCameras == 1 and np.allclose([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], [752.6881, 752.8311, 523.5678, 502.7037])
    sc = refdata.magdalena_scene(cameras=3, width=64, height=64)
    assert sc["vertex_pos"].shape == (1, 5118, 3) and sc["extrinsics"].shape == (1, 36)

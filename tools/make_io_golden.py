#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- golden digests of the REFERENCE's own readers on its bundled fixtures.

Imports python/utils/OBJReader.py and python/utils/CameraReader.py from the reference checkout (authoring container
only) and runs them on python/data/{triangle,cone,magdalena}.obj and the two calibration files.  What they read is
written to tests/golden/io_readers.json as SHA-1 digests of the little-endian fp32 / int32 arrays (+ shapes and a
few leading values), so that tests/test_io_reference.py can pin this repo's readers where /root/reference does not exist.

The reference's OBJReader.__init__ dies inside computeAdjacency under numpy >= 1.24 (np.asarray of a ragged list,
OBJReader.py:158) on every mesh but triangle.obj; the fields are therefore collected by calling its methods one by
one on an instance made with object.__new__, and the adjacency fields are recorded only where it runs.

  python tools/make_io_golden.py [reference root]
"""
import contextlib
import hashlib
import importlib.util
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def digest(a, dtype):
    a = np.ascontiguousarray(np.asarray(a, dtype=dtype))
    return {"shape": list(a.shape), "sha1": hashlib.sha1(a.astype("<" + np.dtype(dtype).str[1:]).tobytes()).hexdigest(),
            "head": [float(x) for x in a.reshape(-1)[:6]]}


def load_ref(ref_root, name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(ref_root, "python", "utils", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def read_obj_with_reference(mod, path):
    """The reference reader's fields, method by method (see module docstring)."""
    r = object.__new__(mod.OBJReader)
    r.filename = path
    r.folderPath = path[0:path.rfind('/') + 1]
    with contextlib.redirect_stdout(io.StringIO()):
        r.readObjFile()
        r.numberOfVertices = len(r.vertexColors)
        r.computePerFaceTextureCoordinated()
        r.loadSegmentationWeights()
        try:
            r.computeAdjacency()
            r._adjacency_ok = True
        except Exception:
            r._adjacency_ok = False
        r.loadMtlTexture(r.mtlFilePathFull, r.mtlFilePath)
    return r


def obj_record(r):
    rec = {"numberOfVertices": int(r.numberOfVertices),
           "facesVertexId": digest(r.facesVertexId, np.int32), "facesTextureId": digest(r.facesTextureId, np.int32),
           "vertexCoordinates": digest(r.vertexCoordinates, np.float32), "vertexColors": digest(r.vertexColors, np.float32),
           "pertVertexTextureCoordinate": digest(r.pertVertexTextureCoordinate, np.float32),
           "textureCoordinates": digest(r.textureCoordinates, np.float32),
           "textureMap": digest(np.asarray(r.textureMap), np.float32), "texHeight": int(r.texHeight), "texWidth": int(r.texWidth),
           "vertexLabels": digest(r.vertexLabels, np.int32)}
    if getattr(r, "_adjacency_ok", False):
        rec["numberOfEdges"] = int(r.numberOfEdges)
        rec["numberOfNeigbours"] = digest(r.numberOfNeigbours, np.float32)
        rec["maximumNumNeighbours"] = int(r.maximumNumNeighbours)
        # the reference stores neighbour ids 1-based in insertion order; as a set per vertex, 0-based:
        rec["neighbourSets"] = [sorted(int(x) - 1 for x in row) for row in r.compressedAdjacency]
    return rec


def cam_record(c):
    return {"numberOfCameras": int(c.numberOfCameras), "extrinsics": digest(c.extrinsics, np.float32), "intrinsics": digest(c.intrinsics, np.float32),
            "intrinsics_f64": [float(x) for x in c.intrinsics], "originalSizeU": [float(x) for x in c.originalSizeU],
            "originalSizeV": [float(x) for x in c.originalSizeV]}


def main():
    ref_root = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("GVV_REFERENCE", "/root/reference")
    data = os.path.join(ref_root, "python", "data")
    OBJ, CAM = load_ref(ref_root, "OBJReader"), load_ref(ref_root, "CameraReader")
    out = {"source": "python/utils/OBJReader.py + python/utils/CameraReader.py of the reference, run by tools/make_io_golden.py", "obj": {}, "cam": {}}
    for name in ("triangle.obj", "cone.obj", "magdalena.obj"):
        out["obj"][name] = obj_record(read_obj_with_reference(OBJ, data + "/" + name))
    for name in ("cameras.calibration", "monocular.calibration"):
        for res in ((1024, 1024), (512, 512), (640, 360)):
            out["cam"][f"{name}@{res[0]}x{res[1]}"] = cam_record(CAM.CameraReader(os.path.join(data, name), res[0], res[1]))
    path = os.path.join(ROOT, "tests", "golden", "io_readers.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()

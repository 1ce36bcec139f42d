#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- stages the reference's bundled fixtures for the tests and the parity tools.

The reference ships its only pinned *inputs* under python/data: the cone / triangle / magdalena
meshes with their MTL + PNG textures, the Skeletool calibrations and the two tensor modules of its scripts.
/root/reference does not exist on the GPU box, so __graft_entry__.build() copies those DATA files -- never
sources -- into tests/_refdata/ (git-ignored, NOT gpurun-ignored: it travels to the box like the built .so
files), the way oracle/build_ref.sh stages the compiled reference kernels into oracle/_ref/.

  python tools/stage_ref_data.py [reference root]      (default $GVV_REFERENCE or /root/reference)
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "tests", "_refdata")
FILES = ["cone.obj", "cone.mtl", "triangle.obj", "untitled.mtl", "magdalena.obj", "magdalena.mtl", "textureMap.png",
         "textureMapEasy.png", "textureMapEasy2.png", "textureMap2.png", "textureMap5.png",
         "cameras.calibration", "monocular.calibration", "segmentation.txt", "test_mesh_tensor.py", "test_SH_tensor.py"]


def stage(ref_root=None):
    ref_root = ref_root or os.environ.get("GVV_REFERENCE", "/root/reference")
    src = os.path.join(ref_root, "python", "data")
    if not os.path.isdir(src):
        return None
    os.makedirs(DEST, exist_ok=True)
    n = 0
    for f in FILES:
        s = os.path.join(src, f)
        if os.path.exists(s):
            shutil.copyfile(s, os.path.join(DEST, f))
            n += 1
    return DEST, n


if __name__ == "__main__":
    r = stage(sys.argv[1] if len(sys.argv) > 1 else None)
    print("reference data not found" if r is None else f"staged {r[1]} files into {r[0]}")

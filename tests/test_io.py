"""CPU: OBJ / calibration readers on files written by the test itself."""
import numpy as np

from gvv_differentiable_cuda_renderer_b200.io import CameraReader, OBJReader


def write_scene(tmp_path, with_texture=True):
    from PIL import Image
    obj = tmp_path / "quad.obj"
    lines = ["# test mesh", "mtllib ./quad.mtl",
             "v 0 0 0 1 0 0", "v 100 0 0 0 1 0", "v 100 100 5 0 0 1", "v 0 100 5 1 1 0", "v 50 50 50 0.5 0.5 0.5",
             "vt 0 0", "vt 1 0", "vt 1 1", "vt 0 1", "vt 0.5 0.5",
             "f 1/1 2/2 3/3", "f 1/1 3/3 4/4", "f 1/1 2/2 5/5 3/3"]        # last one is a quad: first 3 corners kept
    obj.write_text("\n".join(lines) + "\n")
    (tmp_path / "quad.mtl").write_text("newmtl m0\nKd 0.8 0.8 0.8\nmap_Kd tex.png\n")
    if with_texture:
        img = (np.arange(8 * 4 * 3).reshape(8, 4, 3) % 256).astype(np.uint8)
        Image.fromarray(img).save(tmp_path / "tex.png")
    (tmp_path / "segmentation.txt").write_text("0\n1\n2\n3\n4\n")
    return str(obj)


def test_obj_reader_mirrors_reference_attributes(tmp_path):
    r = OBJReader(write_scene(tmp_path))
    assert r.numberOfVertices == 5
    assert r.facesVertexId == [0, 1, 2, 0, 2, 3, 0, 1, 4]
    assert r.facesTextureId == [0, 1, 2, 0, 2, 3, 0, 1, 4]
    assert r.vertexCoordinates[2] == [100.0, 100.0, 5.0] and r.vertexColors[1] == [0.0, 1.0, 0.0]
    assert len(r.textureCoordinates) == 3 * 3 * 2 and r.textureCoordinates[:6] == [0.0, 0.0, 1.0, 0.0, 1.0, 1.0]
    assert r.texHeight == 8 and r.texWidth == 4 and r.textureMap.shape == (8, 4, 3)
    assert abs(float(r.textureMap[0, 1, 0]) - 3 / 255.0) < 1e-6          # RGB order, [0,1] range
    assert r.compressedAdjacency[0] == [1, 2, 3, 4] and r.maximumNumNeighbours == 4
    assert r.numberOfEdges == int(r.numberOfNeigbours.sum()) and r.vertexLabels == [0, 1, 2, 3, 4]
    assert r.faces_array().shape == (3, 3) and r.texcoords_array().shape == (3, 3, 2)


def test_obj_reader_without_texture_or_colours(tmp_path):
    p = tmp_path / "plain.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    r = OBJReader(str(p))
    assert r.textureMap is None and r.vertexColors == [[0.5, 0.5, 0.5]] * 3 and r.textureCoordinates == [0.0] * 6


def test_camera_reader_rescales_intrinsics(tmp_path):
    cal = tmp_path / "cameras.calibration"
    cal.write_text("Skeletool Camera Calibration File V1.0\n"
                   "name          0\n  sensor      10 10\n  size        1024 2048\n  animated    0\n"
                   "  intrinsic   700 0 500 0 0 800 1000 0 0 0 1 0 0 0 0 1 \n"
                   "  extrinsic   1 0 0 10 0 1 0 20 0 0 1 3000 0 0 0 1 \n  radial      0\n"
                   "name          1\n  sensor      10 10\n  size        512 512\n  animated    0\n"
                   "  intrinsic   350 0 250 0 0 350 260 0 0 0 1 0 0 0 0 1 \n"
                   "  extrinsic   0 0 1 1 0 1 0 2 -1 0 0 3 0 0 0 1 \n  radial      0\n")
    c = CameraReader(str(cal), 512, 512)
    assert c.numberOfCameras == 2 and len(c.extrinsics) == 24 and len(c.intrinsics) == 18
    K = np.asarray(c.intrinsics).reshape(2, 3, 3)
    assert np.allclose(K[0], [[350, 0, 250], [0, 200, 250], [0, 0, 1]])       # x0.5 in u, x0.25 in v
    assert np.allclose(K[1], [[350, 0, 250], [0, 350, 260], [0, 0, 1]])
    assert c.extrinsics[:12] == [1, 0, 0, 10, 0, 1, 0, 20, 0, 0, 1, 3000]
    assert c.extrinsics_array().shape == (1, 24) and c.intrinsics_array().dtype == np.float32

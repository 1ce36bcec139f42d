"""TEST INFRASTRUCTURE -- scenes built from the reference's bundled fixtures (tests/_refdata, staged by
tools/stage_ref_data.py from /root/reference/python/data; see tests/_refdata's header there).

  cone_scene()        BASELINE.json config 1 as shipped: python/test_render.py:22-27,64-65 -- cone.obj, the one
                      camera of cameras.calibration, B = 2, 1024x1024, vertexColor + shaded, SH of test_SH_tensor.
  magdalena_scene()   config 3: python/test_gradients_Texture.py:31-58 -- magdalena.obj topology, vertices of
                      test_mesh_tensor.getGTMesh(), textureMap.png; `cameras` > 1 derives a ring from the bundled
                      camera by rotating it about the vertical axis through the mesh centroid.

Everything is read with THIS repo's readers (gvv_differentiable_cuda_renderer_b200.io), so these scenes also
exercise them on the real files.
"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "_refdata")


def available():
    return all(os.path.exists(os.path.join(DATA, f)) for f in ("cone.obj", "magdalena.obj", "cameras.calibration",
                                                                 "test_mesh_tensor.py", "test_SH_tensor.py", "textureMap.png"))


def _module(name):
    spec = importlib.util.spec_from_file_location("_refdata_" + name, os.path.join(DATA, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _readers():
    from gvv_differentiable_cuda_renderer_b200.io import CameraReader, OBJReader
    return CameraReader, OBJReader


def sh_coeff(batch, cams):
    return np.asarray(_module("test_SH_tensor").getSHCoeff(batch, cams), np.float32)


def _pack(obj, verts, cam_E, cam_K, batch, width, height, texture=None):
    """Scene dict in the layout of synthetic.make_scene.  cam_E [C,12], cam_K [C,9]."""
    N = obj.numberOfVertices
    C = cam_E.shape[0]
    tex = np.asarray(obj.textureMap if texture is None else texture, np.float32)
    rep = lambda a, shape: np.ascontiguousarray(np.broadcast_to(np.asarray(a, np.float32).reshape((1,) + shape), (batch,) + shape))
    return dict(faces=obj.faces_array(), texcoords=obj.texcoords_array(), num_vertices=N, num_cameras=C, width=width, height=height,
                vertex_pos=rep(verts, (N, 3)), vertex_color=rep(obj.vertexColors, (N, 3)),
                texture=rep(tex, tex.shape), sh_coeff=sh_coeff(batch, C),
                target_image=np.zeros((batch, C, height, width, 3), np.float32),
                extrinsics=rep(cam_E.reshape(-1), (C * 12,)), intrinsics=rep(cam_K.reshape(-1), (C * 9,)))


def cone_scene(width=1024, height=1024, batch=2):
    CameraReader, OBJReader = _readers()
    cam = CameraReader(os.path.join(DATA, "cameras.calibration"), width, height)
    obj = OBJReader(DATA + "/cone.obj")
    E = np.asarray(cam.extrinsics, np.float32).reshape(cam.numberOfCameras, 12)
    K = np.asarray(cam.intrinsics, np.float32).reshape(cam.numberOfCameras, 9)
    return _pack(obj, obj.vertexCoordinates, E, K, batch, width, height)


def ring_from_camera(E34, centre, n):
    """n world->camera matrices: the given camera rotated about the vertical (world y) axis through `centre` by
    2 pi k / n.  Camera k sees the scene as camera 0 would see it rotated by -angle about that axis:
    E_k = E_0 * T(centre) * R_y(angle) * T(-centre)."""
    E0 = np.eye(4)
    E0[:3] = np.asarray(E34, np.float64).reshape(3, 4)
    out = []
    for k in range(n):
        a = 2.0 * np.pi * k / n
        R = np.eye(4)
        R[0, 0], R[0, 2], R[2, 0], R[2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
        Tp, Tm = np.eye(4), np.eye(4)
        Tp[:3, 3], Tm[:3, 3] = centre, -np.asarray(centre)
        out.append((E0 @ Tp @ R @ Tm)[:3].reshape(-1))
    return np.asarray(out, np.float32)


def magdalena_scene(cameras=1, width=1024, height=1024, batch=1):
    CameraReader, OBJReader = _readers()
    cam = CameraReader(os.path.join(DATA, "cameras.calibration"), width, height)
    obj = OBJReader(DATA + "/magdalena.obj")
    verts = np.asarray(_module("test_mesh_tensor").getGTMesh(), np.float32).reshape(-1, 3)
    assert verts.shape[0] == obj.numberOfVertices
    E = ring_from_camera(cam.extrinsics[:12], verts.astype(np.float64).mean(0), cameras)
    K = np.tile(np.asarray(cam.intrinsics[:9], np.float32).reshape(1, 9), (cameras, 1))
    return _pack(obj, verts, E, K, batch, width, height)

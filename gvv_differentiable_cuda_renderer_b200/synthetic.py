"""Synthetic meshes and cameras for tests and bench (SURVEY.md 8d workloads).

numpy only; units are millimetres, as the reference assumes (RendererUtil.h:29 divides by 1000).
"""
import numpy as np


def uv_sphere(rings, segments, radius=150.0, noise=0.02, seed=0):
    """Closed UV sphere: `rings` latitude bands x `segments` longitude steps, with seeded radial
    noise so that no two fragments tie exactly.  Returns (verts [N,3] f32, faces [F,3] i32,
    texcoords [F,3,2] f32 per face corner)."""
    rng = np.random.default_rng(seed)
    lat = np.linspace(0.0, np.pi, rings + 1)[1:-1]
    lon = np.linspace(0.0, 2.0 * np.pi, segments, endpoint=False)
    la, lo = np.meshgrid(lat, lon, indexing="ij")
    ring = np.stack([np.sin(la) * np.cos(lo), np.cos(la), np.sin(la) * np.sin(lo)], -1).reshape(-1, 3)
    dirs = np.concatenate([[[0.0, 1.0, 0.0]], ring, [[0.0, -1.0, 0.0]]], 0)
    r = radius * (1.0 + noise * (2.0 * rng.random(len(dirs)) - 1.0))
    verts = (dirs * r[:, None]).astype(np.float32)
    uv_v = np.concatenate([[[0.5, 1.0]], np.stack([lo / (2 * np.pi), 1.0 - la / np.pi], -1).reshape(-1, 2), [[0.5, 0.0]]], 0)
    nr = rings - 1
    south = 1 + nr * segments
    faces, tcs = [], []

    def vid(i, j):
        return 1 + i * segments + (j % segments)

    def uv(v, j_hint):
        u, w = uv_v[v]
        if 0 < v < south and j_hint == segments:      # seam: keep u monotone inside the face
            u = 1.0
        return (u, w)

    for j in range(segments):
        a, b = vid(0, j), vid(0, j + 1)
        faces.append((0, b, a)); tcs.append(((lon[j] / (2 * np.pi) + 0.5 / segments, 1.0), uv(b, j + 1), uv(a, j)))
    for i in range(nr - 1):
        for j in range(segments):
            a, b, c, d = vid(i, j), vid(i, j + 1), vid(i + 1, j), vid(i + 1, j + 1)
            faces.append((a, b, c)); tcs.append((uv(a, j), uv(b, j + 1), uv(c, j)))
            faces.append((b, d, c)); tcs.append((uv(b, j + 1), uv(d, j + 1), uv(c, j)))
    for j in range(segments):
        a, b = vid(nr - 1, j), vid(nr - 1, j + 1)
        faces.append((a, b, south)); tcs.append((uv(a, j), uv(b, j + 1), (lon[j] / (2 * np.pi) + 0.5 / segments, 0.0)))
    return verts, np.asarray(faces, np.int32), np.asarray(tcs, np.float32)


def pyramid(size=300.0, height=400.0):
    """5 vertices / 6 faces (apex + square base): a few triangles with huge bounding boxes, the
    regime of the reference's test_render.py scene."""
    s = size / 2
    verts = np.array([[0, height / 2, 0], [-s, -height / 2, -s], [s, -height / 2, -s], [s, -height / 2, s], [-s, -height / 2, s]], np.float32)
    faces = np.array([[0, 2, 1], [0, 3, 2], [0, 4, 3], [0, 1, 4], [1, 2, 3], [1, 3, 4]], np.int32)
    tcs = np.tile(np.array([[0.1, 0.1], [0.9, 0.1], [0.5, 0.9]], np.float32), (6, 1, 1))
    return verts, faces, tcs


def single_triangle(size=200.0):
    verts = np.array([[-size, -size * 0.7, 0], [size, -size * 0.6, 30], [10, size, -20]], np.float32)
    faces = np.array([[0, 1, 2]], np.int32)
    tcs = np.array([[[0.0, 0.0], [1.0, 0.0], [0.5, 1.0]]], np.float32)
    return verts, faces, tcs


def look_at_extrinsics(eye, target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0)):
    """Row-major 3x4 world->camera [R|t], camera looking down +z with y down (image convention)."""
    eye, target, up = (np.asarray(v, np.float64) for v in (eye, target, up))
    z = target - eye
    z /= np.linalg.norm(z)
    x = np.cross(z, up)   # right-handed image frame with y down
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], 0)
    t = -R @ eye
    return np.concatenate([R, t[:, None]], 1).astype(np.float32)


def ring_cameras(n, distance, width, height, focal, elevation=0.0, phase=0.0):
    """n cameras on a ring around the origin.  Returns (extrinsics [n,12], intrinsics [n,9])."""
    E, K = [], []
    for k in range(n):
        a = 2.0 * np.pi * k / n + phase
        eye = (distance * np.sin(a) * np.cos(elevation), distance * np.sin(elevation), distance * np.cos(a) * np.cos(elevation))
        E.append(look_at_extrinsics(eye).reshape(-1))
        K.append(np.array([focal, 0, width / 2.0, 0, focal, height / 2.0, 0, 0, 1], np.float32))
    return np.stack(E).astype(np.float32), np.stack(K).astype(np.float32)


def base_sh():
    """SH coefficients of the reference's demo scenes: [0.7,0,0,-0.5,0,...] per channel
    (python/data/test_SH_tensor.py:4)."""
    return np.tile(np.array([0.7, 0, 0, -0.5, 0, 0, 0, 0, 0], np.float32), 3)


def make_scene(kind="sphere", rings=24, segments=32, cameras=2, width=128, height=128, batch=1, tex=64, seed=0,
               coverage_radius_frac=0.4, distance=1500.0, noise=0.02):
    """A complete set of op inputs (numpy, op layout).  `coverage_radius_frac` = silhouette
    radius / image width for the sphere (0.4 -> ~50 % coverage, the headline workload)."""
    rng = np.random.default_rng(seed + 1)
    if kind == "sphere":
        verts, faces, tcs = uv_sphere(rings, segments, noise=noise, seed=seed)
        radius = 150.0
    elif kind == "pyramid":
        verts, faces, tcs = pyramid()
        radius = 280.0
    elif kind == "triangle":
        verts, faces, tcs = single_triangle()
        radius = 250.0
    else:
        raise ValueError(kind)
    N = len(verts)
    focal = coverage_radius_frac * width * np.sqrt(distance ** 2 - radius ** 2) / radius
    E, K = ring_cameras(cameras, distance, width, height, focal, elevation=0.15, phase=0.3)
    vpos = np.stack([verts + (rng.normal(0, 0.5, verts.shape).astype(np.float32) if b else 0) for b in range(batch)])
    vcol = rng.random((batch, N, 3), dtype=np.float32)
    texture = rng.random((batch, tex, tex, 3), dtype=np.float32)
    sh = (base_sh()[None, None] + rng.random((batch, cameras, 27), dtype=np.float32) * 0.1).astype(np.float32)
    target = np.zeros((batch, cameras, height, width, 3), np.float32)
    return dict(faces=faces, texcoords=tcs, num_vertices=N, num_cameras=cameras, width=width, height=height,
                vertex_pos=vpos.astype(np.float32), vertex_color=vcol, texture=texture, sh_coeff=sh,
                target_image=target, extrinsics=np.tile(E.reshape(1, -1), (batch, 1)),
                intrinsics=np.tile(K.reshape(1, -1), (batch, 1)))

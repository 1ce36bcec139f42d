"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of Oracle 2 (oracle/gvv_oracle.cpp, CPU, OpenMP)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgvv_oracle.so")
LIB64_PATH = os.path.join(_HERE, "libgvv_oracle64.so")      # same source, -DGVVO_FP64: every float is a double
SRC = os.path.join(_HERE, "gvv_oracle.cpp")
ALBEDO = {"vertexColor": 0, "textured": 1, "normal": 2, "lighting": 3, "foregroundMask": 4}
SHADING = {"shaded": 0, "shadeless": 1}


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(SRC):
        # -ffp-contract=off: plain IEEE fp32, no FMA contraction (the GPU contracts; see header of the .cpp)
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c++17", "-o", LIB_PATH, SRC],
                       check=True)
    if force or not os.path.exists(LIB64_PATH) or os.path.getmtime(LIB64_PATH) < os.path.getmtime(SRC):
        subprocess.run(["g++", "-O2", "-DGVVO_FP64", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c++17", "-o", LIB64_PATH, SRC],
                       check=True)
    return LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        vp, i = ctypes.c_void_p, ctypes.c_int
        L.gvvo_forward.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, i] + [vp] * 6 + [vp] * 4 + [vp] * 3 + [i]
        L.gvvo_forward.restype = ctypes.c_longlong
        L.gvvo_backward.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, i, i] + [vp] * 12 + [vp] * 4 + [i]
        L.gvvo_backward.restype = i
        L.gvvo_max_threads.restype = i
        _lib = L
    return _lib


_lib64 = None


def _load64():
    global _lib64
    if _lib64 is None:
        build()
        L = ctypes.CDLL(LIB64_PATH)
        vp, i = ctypes.c_void_p, ctypes.c_int
        L.gvvo_backward.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, i, i] + [vp] * 12 + [vp] * 4 + [i]
        L.gvvo_backward.restype = i
        _lib64 = L
    return _lib64


def set_texture_bilinear(on):
    """Non-default variant: bilinear texture fetch + weighted 4-texel gradient scatter (commented out in the reference)."""
    _load().gvvo_set_texture_bilinear(int(bool(on)))


def max_threads():
    return int(_load().gvvo_max_threads())


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def _p(a):
    return None if a is None else a.ctypes.data


def forward(faces, texcoords, N, C, W, H, albedo, shading, vertex_pos, vertex_color, texture, sh_coeff, extrinsics,
            intrinsics, nthreads=0):
    """Returns dict(bary, face, render, vertex_normal, best_depth, second_depth, tie, fragments)."""
    f = _c(np.asarray(faces).reshape(-1), np.int32)
    t = _c(np.asarray(texcoords).reshape(-1), np.float32)
    F = f.size // 3
    vp_, vc_, tx_, sh_, ex_, in_ = (_c(a, np.float32) for a in (vertex_pos, vertex_color, texture, sh_coeff, extrinsics, intrinsics))
    B, texH, texW = tx_.shape[0], tx_.shape[1], tx_.shape[2]
    bary = np.zeros((B, C, H, W, 2), np.float32)
    face = np.zeros((B, C, H, W), np.int32)
    render = np.zeros((B, C, H, W, 3), np.float32)
    vn = np.zeros((B, C, N, 3), np.float32)
    best = np.zeros((B, C, H, W), np.int32)
    second = np.zeros((B, C, H, W), np.int32)
    tie = np.zeros((B, C, H, W), np.uint8)
    frag = _load().gvvo_forward(_p(f), F, _p(t), N, C, W, H, ALBEDO[albedo], SHADING[shading], B, texH, texW,
                                _p(vp_), _p(vc_), _p(tx_), _p(sh_), _p(ex_), _p(in_),
                                _p(bary), _p(face), _p(render), _p(vn), _p(best), _p(second), _p(tie), nthreads)
    return dict(bary=bary, face=face, render=render, vertex_normal=vn, best_depth=best, second_depth=second, tie=tie,
                fragments=int(frag))


def normal_map(faces, texcoords, N, C, vertex_pos, tex_h, tex_w):
    """compute_normal_map path.  Returns dict(normal_map [B,texH,texW,3], vertex_normal, covered, tie)."""
    L = _load()
    vp, i = ctypes.c_void_p, ctypes.c_int
    L.gvvo_normal_map.argtypes = [vp, i, vp, i, i, i, i, i, vp, vp, vp, vp, vp]
    L.gvvo_normal_map.restype = i
    f = _c(np.asarray(faces).reshape(-1), np.int32)
    t = _c(np.asarray(texcoords).reshape(-1), np.float32)
    pos = _c(vertex_pos, np.float32)
    B = pos.shape[0]
    vn = np.zeros((B, C, N, 3), np.float32)
    nm = np.zeros((B, tex_h, tex_w, 3), np.float32)
    cov = np.zeros((tex_h, tex_w), np.uint8)
    tie = np.zeros((tex_h, tex_w), np.uint8)
    L.gvvo_normal_map(_p(f), f.size // 3, _p(t), N, C, B, tex_h, tex_w, _p(pos), _p(vn), _p(nm), _p(cov), _p(tie))
    return dict(normal_map=nm, vertex_normal=vn, covered=cov, tie=tie)


def backward(faces, texcoords, N, C, W, H, albedo, shading, image_filter, render_grad, target_grad, vertex_pos,
             vertex_color, texture, sh_coeff, target_image, vertex_normal, bary, face, extrinsics, intrinsics, nthreads=0, fp64=False):
    """Returns (vertex_pos_grad, vertex_color_grad, texture_grad, sh_coeff_grad).
    fp64: evaluate the same formulas at the same (fp32-valued) inputs in double precision and return float64 arrays."""
    ft = np.float64 if fp64 else np.float32
    f = _c(np.asarray(faces).reshape(-1), np.int32)
    t = _c(np.asarray(texcoords).reshape(-1), ft)
    F = f.size // 3
    rg, tg, vp_, vc_, tx_, sh_, ti_, vn_, ba_, ex_, in_ = (_c(a, ft) for a in (
        render_grad, target_grad, vertex_pos, vertex_color, texture, sh_coeff, target_image, vertex_normal, bary, extrinsics, intrinsics))
    fb = _c(face, np.int32)
    B, texH, texW = tx_.shape[0], tx_.shape[1], tx_.shape[2]
    gpos = np.zeros((B, N, 3), ft)
    gcol = np.zeros((B, N, 3), ft)
    gtex = np.zeros((B, texH, texW, 3), ft)
    gsh = np.zeros((B, C, 27), ft)
    rc = (_load64() if fp64 else _load()).gvvo_backward(_p(f), F, _p(t), N, C, W, H, ALBEDO[albedo], SHADING[shading], image_filter, B, texH, texW,
                               _p(rg), _p(tg), _p(vp_), _p(vc_), _p(tx_), _p(sh_), _p(ti_), _p(vn_), _p(ba_), _p(fb),
                               _p(ex_), _p(in_), _p(gpos), _p(gcol), _p(gtex), _p(gsh), nthreads)
    if rc:
        raise RuntimeError("Unsupported color mode in renderer gradient!")
    return gpos, gcol, gtex, gsh

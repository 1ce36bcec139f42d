"""TEST INFRASTRUCTURE ONLY -- loader for Oracle 1 (oracle/_ref/libgvv_ref.so, built by
oracle/build_ref.sh from the UNMODIFIED reference sources + oracle/ref_harness.cu).

Runs the reference's own kernels on the GPU; used as the bit-exact comparator for the face buffer
and as the "reference CUDA on B200" timing arm.  Handles are never destroyed (the reference's
destructor frees caller-owned pointers, CUDABasedRasterization.cpp:112-121).
"""
import ctypes
import os

import numpy as np
import torch

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libgvv_ref.so")


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def _load():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(LIB_PATH)
        vp, i = ctypes.c_void_p, ctypes.c_int
        L.gvvref_create.argtypes = [vp, i, vp, i, i, i, i, ctypes.c_char_p, ctypes.c_char_p, i, i, i, i]
        L.gvvref_create.restype = vp
        L.gvvref_forward.argtypes = [vp, i, i, i] + [vp] * 7 + [vp] * 6 + [vp] * 4
        L.gvvref_forward.restype = i
        L.gvvref_backward.argtypes = [vp, i, i, i] + [vp] * 12 + [vp] * 4
        L.gvvref_backward.restype = i
        _lib = L
    return _lib


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class RefRenderer:
    def __init__(self, faces, texcoords, N, C, W, H, albedo, shading, image_filter=1, texture_filter=1,
                 compute_normal=False, with_backward=True):
        f = np.ascontiguousarray(np.asarray(faces, np.int32).reshape(-1))
        t = np.ascontiguousarray(np.asarray(texcoords, np.float32).reshape(-1))
        self.F, self.N, self.C, self.W, self.H = f.size // 3, N, C, W, H
        self.h = _load().gvvref_create(f.ctypes.data, self.F, t.ctypes.data, N, C, W, H, albedo.encode(), shading.encode(),
                                       image_filter, texture_filter, int(compute_normal), int(with_backward))

    def forward(self, vpos, vcol, tex, sh, target, extr, intr, intermediates=False):
        B, texH, texW = tex.shape[0], tex.shape[1], tex.shape[2]
        C, N, W, H = self.C, self.N, self.W, self.H
        dev = vpos.device
        o = dict(device=dev, dtype=torch.float32)
        bary = torch.zeros((B, C, H, W, 2), **o)
        face = torch.zeros((B, C, H, W), device=dev, dtype=torch.int32)
        render = torch.zeros((B, C, H, W, 3), **o)
        vnormal = torch.zeros((B, C, N, 3), **o)
        target_out = torch.zeros((B, C, H, W, 3), **o)
        nmap = torch.zeros((B, texH, texW, 3), **o)
        depth = torch.zeros((B, C, H, W), device=dev, dtype=torch.int32) if intermediates else None
        cam = torch.zeros((B, C, 32), **o) if intermediates else None
        proj = torch.zeros((B, C, N, 3), **o) if intermediates else None
        bbox = torch.zeros((B, C, self.F, 4), device=dev, dtype=torch.int32) if intermediates else None
        torch.cuda.synchronize()
        rc = _load().gvvref_forward(self.h, B, texH, texW, _p(vpos), _p(vcol), _p(tex), _p(sh), _p(target), _p(extr), _p(intr),
                                    _p(bary), _p(face), _p(render), _p(vnormal), _p(target_out), _p(nmap),
                                    _p(depth), _p(cam), _p(proj), _p(bbox))
        torch.cuda.synchronize()
        if rc:
            raise RuntimeError(f"reference forward failed: cuda error {rc}")
        out = dict(bary=bary, face=face, render=render, vertex_normal=vnormal, target_out=target_out, normal_map=nmap)
        if intermediates:
            out.update(depth=depth, cam=cam, proj=proj, bbox=bbox)
        return out

    def backward(self, render_grad, vpos, vcol, tex, sh, target, vnormal, bary, face, target_grad, extr, intr):
        B, texH, texW = tex.shape[0], tex.shape[1], tex.shape[2]
        dev = vpos.device
        o = dict(device=dev, dtype=torch.float32)
        gpos = torch.zeros((B, self.N, 3), **o)
        gcol = torch.zeros((B, self.N, 3), **o)
        gtex = torch.zeros((B, texH, texW, 3), **o)
        gsh = torch.zeros((B, self.C, 27), **o)
        if target_grad is None:
            target_grad = torch.zeros_like(render_grad)
        torch.cuda.synchronize()
        rc = _load().gvvref_backward(self.h, B, texH, texW, _p(render_grad), _p(vpos), _p(vcol), _p(tex), _p(sh), _p(target),
                                     _p(vnormal), _p(bary), _p(face), _p(target_grad), _p(extr), _p(intr),
                                     _p(gpos), _p(gcol), _p(gtex), _p(gsh))
        torch.cuda.synchronize()
        if rc:
            raise RuntimeError(f"reference backward failed: cuda error {rc}")
        return gpos, gcol, gtex, gsh

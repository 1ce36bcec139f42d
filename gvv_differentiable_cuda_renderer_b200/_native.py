"""ctypes binding of libgvv_b200.so (include/gvv_b200.h).

This is the only way the Python layer reaches the GPU: there is no CPU fallback and no
alternative backend.  If the shared library is missing the import of this module fails loudly.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("GVV_B200_LIB") or os.path.join(_PKG, "libgvv_b200.so")   # (the override serves A/B builds of the library, tools/build_variants.sh)
CSRC = os.path.join(_PKG, "csrc")
SOURCES = ["gvv_api.cu", "gvv_forward.cu", "gvv_backward.cu", "gvv_normalmap.cu", "gvv_helpers.cu", "gvv_microbench.cu"]

ALBEDO_MODES = {"vertexColor": 0, "textured": 1, "normal": 2, "lighting": 3, "foregroundMask": 4}
SHADING_MODES = {"shaded": 0, "shadeless": 1}


def build_library(force=False, verbose=False):
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    deps.append(os.path.join(_ROOT, "include", "gvv_b200.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB_PATH


class gvv_desc(ctypes.Structure):
    _fields_ = [("faces", ctypes.c_void_p), ("num_faces", ctypes.c_int32), ("texcoords", ctypes.c_void_p),
                ("num_vertices", ctypes.c_int32), ("num_cameras", ctypes.c_int32), ("width", ctypes.c_int32),
                ("height", ctypes.c_int32), ("albedo_mode", ctypes.c_int32), ("shading_mode", ctypes.c_int32),
                ("image_filter_size", ctypes.c_int32), ("texture_filter_size", ctypes.c_int32),
                ("compute_normal_map", ctypes.c_int32), ("device", ctypes.c_int32)]


class gvv_allreduce_desc(ctypes.Structure):
    _fields_ = [("peer_buffers", ctypes.c_void_p), ("signal_pads", ctypes.c_void_p), ("multicast_ptr", ctypes.c_uint64),
                ("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("offset_floats", ctypes.c_int64), ("count_floats", ctypes.c_int64),
                ("result", ctypes.c_void_p), ("mode", ctypes.c_int32), ("channels", ctypes.c_int32), ("first_channel", ctypes.c_int32),
                ("epoch_word", ctypes.c_int32), ("after_backward", ctypes.c_int32)]


_lib = None

# every symbol include/gvv_b200.h declares
EXPORTS = ["gvv_create", "gvv_destroy", "gvv_reserve", "gvv_forward", "gvv_backward", "gvv_last_error", "gvv_launch_count",
           "gvv_debug_copy", "gvv_debug_eval", "gvv_set_option", "gvv_bench_atomics",
           "gvv_kernel_count", "gvv_kernel_name", "gvv_kernel_times",
           "gvv_gaussian_smooth", "gvv_image_gradient", "gvv_set_target_gradient", "gvv_set_allreduce"]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the renderer has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        L.gvv_create.argtypes = [ctypes.POINTER(gvv_desc), ctypes.POINTER(vp)]
        L.gvv_create.restype = ctypes.c_int
        L.gvv_destroy.argtypes = [vp]
        L.gvv_destroy.restype = ctypes.c_int
        L.gvv_reserve.argtypes = [vp, i32, vp]
        L.gvv_reserve.restype = ctypes.c_int
        L.gvv_forward.argtypes = [vp, i32, i32, i32] + [vp] * 7 + [vp] * 6 + [vp]
        L.gvv_forward.restype = ctypes.c_int
        L.gvv_backward.argtypes = [vp, i32, i32, i32] + [vp] * 12 + [vp] * 4 + [vp]
        L.gvv_backward.restype = ctypes.c_int
        L.gvv_last_error.restype = ctypes.c_char_p
        L.gvv_launch_count.argtypes = [vp]
        L.gvv_launch_count.restype = i64
        L.gvv_debug_copy.argtypes = [vp, i32, vp, i64, vp]
        L.gvv_debug_copy.restype = i64
        L.gvv_debug_eval.argtypes = [vp, i32, vp, vp, vp, vp]
        L.gvv_debug_eval.restype = ctypes.c_int
        L.gvv_set_option.argtypes = [vp, ctypes.c_char_p, i32]
        L.gvv_set_option.restype = ctypes.c_int
        L.gvv_kernel_count.restype = i32
        L.gvv_kernel_name.argtypes = [i32]
        L.gvv_kernel_name.restype = ctypes.c_char_p
        L.gvv_kernel_times.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)]
        L.gvv_kernel_times.restype = ctypes.c_int
        L.gvv_bench_atomics.argtypes = [i32, i32, i64, i64, i32, ctypes.POINTER(ctypes.c_double)]
        L.gvv_bench_atomics.restype = ctypes.c_int
        L.gvv_gaussian_smooth.argtypes = [i32, i64, i32, i32, i32, vp, vp, vp, vp, vp]
        L.gvv_gaussian_smooth.restype = ctypes.c_int
        L.gvv_image_gradient.argtypes = [i32, i64, i32, i32, i32, vp, vp, vp, vp]
        L.gvv_image_gradient.restype = ctypes.c_int
        L.gvv_set_allreduce.argtypes = [vp, ctypes.POINTER(gvv_allreduce_desc)]
        L.gvv_set_allreduce.restype = ctypes.c_int
        L.gvv_set_target_gradient.argtypes = [vp, vp, vp]
        L.gvv_set_target_gradient.restype = ctypes.c_int
        _lib = L
    return _lib


class GvvError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise GvvError(f"{what}: [{rc}] {lib().gvv_last_error().decode()}")


def _ptr(t):
    return None if t is None else t.data_ptr()      # ctypes converts the int through argtypes = c_void_p


def _f32(t, name, device):
    if t is None:
        return None
    if t.dtype is torch.float32 and t.device == device and t.is_contiguous():   # fast path: nothing to do
        return t
    if t.device != device:
        raise GvvError(f"{name} is on {t.device}, the renderer lives on {device}")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def normalize_device(device):
    """torch.device with an explicit index: 'cuda' means the current device (torch never compares cuda == cuda:0)."""
    if device is None:
        return torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
    device = torch.device(device)
    if device.type != "cuda":
        raise GvvError(f"the renderer lives on a CUDA device, not on {device} (there is no CPU path)")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
    return device


def _numel(t, n, name):
    if t is not None and t.numel() != n:
        raise GvvError(f"{name} has {t.numel()} elements, expected {n}")


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NULL = _NullCtx()


def _device_ctx(dev):
    """torch.cuda.device(dev) only when dev is not already current (the context manager costs ~10 us per call)."""
    return _NULL if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)


class NativeRenderer:
    """One gvv_handle: immutable topology + scratch on one CUDA device."""

    def __init__(self, faces, texcoords, num_vertices, num_cameras, width, height, albedo_mode, shading_mode,
                 image_filter_size=1, texture_filter_size=1, compute_normal_map=False, device=None):
        if albedo_mode not in ALBEDO_MODES:
            raise GvvError("INVALID ALBEDO MODE")        # CudaRenderer.cpp:60-64
        if shading_mode not in SHADING_MODES:
            raise GvvError("INVALID SHADING MODE")       # CudaRenderer.cpp:67-71
        if not torch.cuda.is_available():
            raise GvvError("CUDA device required: the renderer has no CPU fallback")
        self.device = normalize_device(device)
        f = np.ascontiguousarray(np.asarray(faces, dtype=np.int32).reshape(-1))
        if f.size % 3:
            raise GvvError("No triangular faces!")        # CUDABasedRasterization.cpp:38-41
        self.F = f.size // 3
        t = None
        if texcoords is not None and len(texcoords):
            t = np.ascontiguousarray(np.asarray(texcoords, dtype=np.float32).reshape(-1))
            if t.size != self.F * 6:
                raise GvvError("Texture coordinates have wrong dimensionality!")   # CUDABasedRasterization.cpp:49-52
        self.N, self.C, self.W, self.H = int(num_vertices), int(num_cameras), int(width), int(height)
        self.albedo_mode, self.shading_mode = albedo_mode, shading_mode
        self.compute_normal_map = bool(compute_normal_map)
        d = gvv_desc(f.ctypes.data, self.F, t.ctypes.data if t is not None else None, self.N, self.C, self.W, self.H,
                     ALBEDO_MODES[albedo_mode], SHADING_MODES[shading_mode], int(image_filter_size),
                     int(texture_filter_size), int(self.compute_normal_map), self.device.index)
        h = ctypes.c_void_p()
        _check(lib().gvv_create(ctypes.byref(d), ctypes.byref(h)), "gvv_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().gvv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reserve(self, max_batch):
        """Allocate the scratch for calls of up to max_batch batch elements now (before CUDA-graph capture)."""
        with _device_ctx(self.device):
            _check(lib().gvv_reserve(self._h, int(max_batch), self._stream()), "gvv_reserve")

    def set_option(self, key, value):
        _check(lib().gvv_set_option(self._h, key.encode(), int(value)), "gvv_set_option")
        if key == "shared_batch_grads":
            self.shared_batch_grads = bool(value)

    @property
    def launch_count(self):
        return int(lib().gvv_launch_count(self._h))

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def forward(self, vertex_pos, vertex_color, texture, sh_coeff, target_image, extrinsics, intrinsics):
        dev = self.device
        vertex_pos = _f32(vertex_pos, "vertex_pos", dev)
        vertex_color = _f32(vertex_color, "vertex_color", dev)
        texture = _f32(texture, "texture", dev)
        sh_coeff = _f32(sh_coeff, "sh_coeff", dev)
        target_image = _f32(target_image, "target_image", dev)
        extrinsics = _f32(extrinsics, "extrinsics", dev)
        intrinsics = _f32(intrinsics, "intrinsics", dev)
        # B, texH, texW come from the texture tensor, as in CudaRenderer.cpp:207-209
        B, texH, texW = int(texture.shape[0]), int(texture.shape[1]), int(texture.shape[2])
        C, N, W, H = self.C, self.N, self.W, self.H
        if texture.dim() != 4 or texture.shape[3] != 3:
            raise GvvError(f"texture must be [B, texH, texW, 3], got {tuple(texture.shape)}")
        for t, n, name in ((vertex_pos, B * N * 3, "vertex_pos"), (sh_coeff, B * C * 27, "sh_coeff"),
                           (extrinsics, B * C * 12, "extrinsics"), (intrinsics, B * C * 9, "intrinsics"),
                           (vertex_color, B * N * 3, "vertex_color"), (target_image, B * C * H * W * 3, "target_image")):
            _numel(t, n, name)
        o = dict(device=dev, dtype=torch.float32)
        with _device_ctx(dev):
            bary = torch.empty((B, C, H, W, 2), **o)
            face = torch.empty((B, C, H, W), device=dev, dtype=torch.int32)
            render = torch.empty((B, C, H, W, 3), **o)
            vnormal = torch.empty((B, C, N, 3), **o)
            normal_map = torch.zeros((B, texH, texW, 3), **o) if self.compute_normal_map else torch.empty((0,), **o)
            if self.compute_normal_map:
                # rasterisation is skipped (CUDABasedRasterization.cu:463-466): the raster outputs stay
                # unwritten in the reference; we define them as background
                bary.zero_(); face.fill_(-1); render.zero_(); render[..., 1] = 1.0
            # out4 is a copy of in4 in the reference; aliasing it saves 12 B/px and keeps the gradient path
            target_out = target_image
            _check(lib().gvv_forward(self._h, B, texH, texW, _ptr(vertex_pos), _ptr(vertex_color), _ptr(texture),
                                     _ptr(sh_coeff), _ptr(target_image), _ptr(extrinsics), _ptr(intrinsics),
                                     _ptr(bary), _ptr(face), _ptr(render), _ptr(vnormal), _ptr(target_out),
                                     _ptr(normal_map) if self.compute_normal_map else None, self._stream()),
                   "gvv_forward")
        return bary, face, render, vnormal, target_out, normal_map

    def backward(self, render_grad, target_grad, vertex_pos, vertex_color, texture, sh_coeff, target_image,
                 vertex_normal, bary, face, extrinsics, intrinsics, out=None):
        """out: optional preallocated (vertex_pos_grad, vertex_color_grad, texture_grad, sh_coeff_grad) fp32 contiguous
        tensors (entries may be None) -- e.g. views into ONE flat buffer, so that the gradients of parameters shared
        across ranks can be all-reduced in place without packing (sharding.allreduce_shared_grads)."""
        dev = self.device
        # B comes from vertex_pos like in the reference's gradient op (CudaRendererGrad.cpp:195; the forward takes it
        # from the texture, CudaRenderer.cpp:207); where the reference would overrun on a mismatch, this raises
        C, N, W, H = self.C, self.N, self.W, self.H
        if vertex_pos.dim() < 2 or vertex_pos.numel() % (N * 3):
            raise GvvError(f"vertex_pos must be [B, {N}, 3], got {tuple(vertex_pos.shape)}")
        B, texH, texW = vertex_pos.numel() // (N * 3), int(texture.shape[1]), int(texture.shape[2])
        if texture.dim() != 4 or texture.shape[3] != 3 or int(texture.shape[0]) != B:
            raise GvvError(f"texture must be [{B}, texH, texW, 3] (batch of vertex_pos), got {tuple(texture.shape)}")
        if face.dtype != torch.int32:
            raise GvvError(f"face_buffer must be int32, got {face.dtype}")
        if face.device != dev:
            raise GvvError(f"face_buffer is on {face.device}, the renderer lives on {dev}")
        P = B * C * H * W
        for t, n, name in ((render_grad, P * 3, "render_buffer_grad"), (target_grad, P * 3, "target_buffer_grad"),
                           (vertex_pos, B * N * 3, "vertex_pos"), (vertex_color, B * N * 3, "vertex_color"),
                           (sh_coeff, B * C * 27, "sh_coeff"), (target_image, P * 3, "target_image"),
                           (vertex_normal, B * C * N * 3, "vertex_normal"), (bary, P * 2, "barycentric_buffer"),
                           (face, P, "face_buffer"), (extrinsics, B * C * 12, "extrinsics"), (intrinsics, B * C * 9, "intrinsics")):
            _numel(t, n, name)
        args = [_f32(t, n, dev) for t, n in [(render_grad, "render_buffer_grad"), (vertex_pos, "vertex_pos"),
                                              (vertex_color, "vertex_color"), (texture, "texture"),
                                              (sh_coeff, "sh_coeff"), (target_image, "target_image"),
                                              (vertex_normal, "vertex_normal"), (bary, "barycentric_buffer")]]
        face = face.contiguous()
        target_grad = _f32(target_grad, "target_buffer_grad", dev)
        extrinsics = _f32(extrinsics, "extrinsics", dev)
        intrinsics = _f32(intrinsics, "intrinsics", dev)
        o = dict(device=dev, dtype=torch.float32)
        with _device_ctx(dev):
            pre = tuple(out) if out is not None else (None, None, None, None)
            Bs = 1 if getattr(self, "shared_batch_grads", False) else B     # set_option("shared_batch_grads", 1)
            shapes = ((B, N, 3), (Bs, N, 3), (Bs, texH, texW, 3), (Bs, C, 27))
            outs = []
            for t, shp, name in zip(pre, shapes, ("vertex_pos_grad", "vertex_color_grad", "texture_grad", "sh_coeff_grad")):
                if t is None:
                    t = torch.empty(shp, **o)
                elif t.device != dev or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != int(np.prod(shp)):
                    raise GvvError(f"out[{name}] must be a contiguous fp32 tensor of {shp} on {dev}")
                outs.append(t)
            gpos, gcol, gtex, gsh = outs
            _check(lib().gvv_backward(self._h, B, texH, texW, *[_ptr(a) for a in args], _ptr(face), _ptr(target_grad),
                                      _ptr(extrinsics), _ptr(intrinsics), _ptr(gpos), _ptr(gcol), _ptr(gtex),
                                      _ptr(gsh), self._stream()), "gvv_backward")
        return gpos, gcol, gtex, gsh

    def set_allreduce(self, desc):
        """desc: gvv_allreduce_desc or None -- the one-shot all-reduce gvv_backward ends with (sharding.SymmetricGradBuffer)."""
        _check(lib().gvv_set_allreduce(self._h, ctypes.byref(desc) if desc is not None else None), "gvv_set_allreduce")

    def set_target_gradient(self, d_du, d_dv):
        """Hand the precomputed target-image gradient (image_gradient(target, image_filter_size)) to the
        backward's model-to-data term; (None, None) restores the per-pixel recomputation.  The tensors
        must stay alive (and unchanged) while they are set."""
        self._target_grad_cache = (d_du, d_dv)
        _check(lib().gvv_set_target_gradient(self._h, _ptr(d_du), _ptr(d_dv)), "gvv_set_target_gradient")

    def debug_copy(self, which, nbytes):
        buf = np.empty(nbytes, dtype=np.uint8)
        n = lib().gvv_debug_copy(self._h, which, buf.ctypes.data, nbytes, self._stream())
        if n < 0:
            raise GvvError("gvv_debug_copy failed")
        return buf[:min(n, nbytes)]


def _eval_pairs(self, queries):
    """queries int32 [n,4] = (view, x, y, face) -> (keys int32 [n], ab float32 [n,2]); key = INT32_MIN on a miss."""
    q = np.ascontiguousarray(queries, dtype=np.int32).reshape(-1, 4)
    keys = np.zeros(len(q), np.int32)
    ab = np.zeros((len(q), 2), np.float32)
    _check(lib().gvv_debug_eval(self._h, len(q), q.ctypes.data, keys.ctypes.data, ab.ctypes.data, self._stream()), "gvv_debug_eval")
    return keys, ab


NativeRenderer.eval_pairs = _eval_pairs


def _kernel_times(self):
    """{kernel name: (total device ms, launches)} since set_option('time_kernels', 1); clears the log."""
    n = lib().gvv_kernel_count()
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_int64 * n)()
    _check(lib().gvv_kernel_times(self._h, ms, cnt), "gvv_kernel_times")
    return {lib().gvv_kernel_name(i).decode(): (ms[i], cnt[i]) for i in range(n) if cnt[i]}


NativeRenderer.kernel_times = _kernel_times


def bench_atomics(kind, n_addr, n_ops, iters=10, device=0):
    out = ctypes.c_double(0.0)
    _check(lib().gvv_bench_atomics(device, kind, n_addr, n_ops, iters, ctypes.byref(out)), "gvv_bench_atomics")
    return out.value


def image_gradient(image, filter_size):
    """imageGradient (RendererUtil.h:566-620) of [..., H, W, 3] fp32 CUDA images -> (dI/du, dI/dv)."""
    image = image.contiguous().float()
    H, W = int(image.shape[-3]), int(image.shape[-2])
    n = image.numel() // (H * W * 3)
    du, dv = torch.empty_like(image), torch.empty_like(image)
    with torch.cuda.device(image.device):
        _check(lib().gvv_image_gradient(image.device.index, n, H, W, int(filter_size), _ptr(image), _ptr(du), _ptr(dv),
                                        ctypes.c_void_p(torch.cuda.current_stream(image.device).cuda_stream)), "gvv_image_gradient")
    return du, dv


def gaussian_smooth(image, taps):
    """Separable depthwise Gaussian (zero SAME padding) of [..., H, W, 3] fp32 CUDA images with the
    normalised 1-D kernel `taps` (odd length), see GaussianSmoothingGpu.smoothImage."""
    image = image.contiguous().float()
    taps = np.ascontiguousarray(taps, dtype=np.float32)
    if taps.size % 2 != 1:
        raise GvvError("gaussian_smooth needs an odd number of taps")
    H, W = int(image.shape[-3]), int(image.shape[-2])
    n = image.numel() // (H * W * 3)
    tmp, out = torch.empty_like(image), torch.empty_like(image)
    with torch.cuda.device(image.device):
        _check(lib().gvv_gaussian_smooth(image.device.index, n, H, W, taps.size // 2, taps.ctypes.data, _ptr(image), _ptr(tmp), _ptr(out),
                                         ctypes.c_void_p(torch.cuda.current_stream(image.device).cuda_stream)), "gvv_gaussian_smooth")
    return out

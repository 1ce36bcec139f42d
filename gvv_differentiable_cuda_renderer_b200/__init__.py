"""B200-native differentiable rasteriser: drop-in for GVV-Differentiable-CUDA-Renderer's op."""
from . import synthetic  # noqa: F401  (numpy only)

__all__ = ["CudaRendererGpu", "NativeRenderer", "synthetic"]


def __getattr__(name):   # torch / CUDA pieces are imported lazily so `import` works on CPU-only boxes
    if name == "CudaRendererGpu":
        from .CudaRenderer import CudaRendererGpu
        return CudaRendererGpu
    if name == "NativeRenderer":
        from ._native import NativeRenderer
        return NativeRenderer
    raise AttributeError(name)

"""B200-native per-frame rendering path of kdedeurw/Custom-GPGPU-Rasterizer.

The product is ``csrc/libb2r.so`` (hand-written CUDA for sm_100a behind the C ABI of
``include/b2r.h``, plus the C++ veneer that keeps the reference's class surface).  The Python
modules are a thin harness: ``scenefile``/``scenes`` describe inputs and need only numpy;
``renderer`` binds the C ABI with ctypes and fails at import if the library is not built.
"""
from .scenefile import Scene, MeshData  # noqa: F401

__all__ = ["Scene", "MeshData"]
